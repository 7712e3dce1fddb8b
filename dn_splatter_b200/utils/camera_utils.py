"""Camera-space helpers with the reference's names (/root/reference/dn_splatter/utils/camera_utils.py:70-172).
Plain torch: they are init/eval-time utilities (the training-step use, inside normal_from_depth_image, is the
CUDA stencil in normal_utils.py)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor


def get_camera_coords(img_size: tuple, pixel_offset: float = 0.5) -> Tensor:
    """[H*W,2] pixel centres, x fastest (reference :70-89)."""
    w, h = int(img_size[0]), int(img_size[1])
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return (torch.stack([xs, ys], dim=-1).reshape(-1, 2) + pixel_offset).float()


def get_means3d_backproj(depths: Tensor, fx: float, fy: float, cx: float, cy: float, img_size: tuple, c2w: Tensor,
                         device: torch.device, mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Back-projection of a z-depth map to world points (reference :92-144)."""
    d = depths.reshape(-1, 1).float().to(device)
    uv = get_camera_coords(img_size).to(device)
    pts = torch.cat([(uv[:, 0:1] - cx) * d / fx, (uv[:, 1:2] - cy) * d / fy, d], dim=-1)
    if mask is not None:
        mask = torch.as_tensor(mask, device=device)
        pts, uv = pts[mask], uv[mask]
    c2w = c2w.float().to(device)
    return pts @ torch.linalg.inv(c2w[..., :3, :3]) + c2w[..., :3, 3], uv


def project_pix(p: Tensor, fx: float, fy: float, cx: float, cy: float, c2w: Tensor, device: torch.device,
                return_z_depths: bool = False) -> Tensor:
    """World points -> pixel coordinates (reference :147-172)."""
    c2w = c2w.to(device)
    pc = (p.to(device) - c2w[..., :3, 3]) @ c2w[..., :3, :3]
    u, v = pc[:, 0] * fx / pc[:, 2] + cx, pc[:, 1] * fy / pc[:, 2] + cy
    return torch.stack([u, v, pc[:, 2]], dim=-1) if return_z_depths else torch.stack([u, v], dim=-1)


def get_colored_points_from_depth(depths: Tensor, rgbs: Tensor, c2w: Tensor, fx: float, fy: float, cx: float, cy: float,
                                  img_size: tuple, mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Coloured world points of a depth + rgb frame (reference :175-210)."""
    points, _ = get_means3d_backproj(depths=depths.float(), fx=fx, fy=fy, cx=cx, cy=cy, img_size=img_size, c2w=c2w.float(),
                                     device=depths.device)
    colors = rgbs.reshape(-1, 3)
    if mask is not None:
        mask = torch.as_tensor(mask, device=depths.device)
        return points[mask], colors[mask]
    return points, colors

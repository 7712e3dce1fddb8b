"""Duck-typed stand-in for nerfstudio.cameras.cameras.Cameras [EXT] used when nerfstudio is not installed.

Only what DNSplatterModel.get_outputs / get_loss_dict touch (reference dn_model.py:416-424, 473-479,
580-596): camera_to_worlds [B,3,4], fx/fy/cx/cy/width/height [B,1], shape, metadata,
get_intrinsics_matrices(), rescale_output_resolution().  Intrinsics live on the host so reading W/H
never synchronises the device; camera_to_worlds may live on either side."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

try:  # pragma: no cover - nerfstudio is optional
    from nerfstudio.cameras.cameras import Cameras as _NSCameras
except Exception:  # noqa: BLE001
    _NSCameras = None


class Cameras:
    def __init__(self, camera_to_worlds: Tensor, fx, fy, cx, cy, width, height, metadata: Optional[Dict] = None):
        c2w = torch.as_tensor(camera_to_worlds, dtype=torch.float32)
        if c2w.dim() == 2:
            c2w = c2w[None]
        self.camera_to_worlds = c2w[:, :3, :4]
        b = c2w.shape[0]

        def col(v, dt):
            t = torch.as_tensor(v, dtype=dt).reshape(-1, 1).cpu()
            return t.expand(b, 1).clone() if t.shape[0] == 1 else t

        self.fx, self.fy, self.cx, self.cy = (col(v, torch.float32) for v in (fx, fy, cx, cy))
        self.width, self.height = col(width, torch.int64), col(height, torch.int64)
        self.metadata = metadata

    @property
    def shape(self):
        return self.camera_to_worlds.shape[:1]

    @property
    def device(self):
        return self.camera_to_worlds.device

    def to(self, device):
        self.camera_to_worlds = self.camera_to_worlds.to(device)
        return self

    def __getitem__(self, i):
        sl = slice(i, i + 1) if isinstance(i, int) else i
        return Cameras(self.camera_to_worlds[sl], self.fx[sl], self.fy[sl], self.cx[sl], self.cy[sl], self.width[sl],
                       self.height[sl], self.metadata)

    def get_intrinsics_matrices(self) -> Tensor:
        K = torch.zeros(self.shape[0], 3, 3, dtype=torch.float32)
        K[:, 0, 0], K[:, 1, 1] = self.fx[:, 0], self.fy[:, 0]
        K[:, 0, 2], K[:, 1, 2] = self.cx[:, 0], self.cy[:, 0]
        K[:, 2, 2] = 1.0
        return K

    def rescale_output_resolution(self, scaling_factor: float) -> None:
        if scaling_factor == 1 or scaling_factor == 1.0:
            return
        self.fx, self.fy = self.fx * scaling_factor, self.fy * scaling_factor
        self.cx, self.cy = self.cx * scaling_factor, self.cy * scaling_factor
        self.width = (self.width * scaling_factor + 0.5).floor().to(torch.int64)
        self.height = (self.height * scaling_factor + 0.5).floor().to(torch.int64)


def is_camera(obj) -> bool:
    return isinstance(obj, Cameras) or (_NSCameras is not None and isinstance(obj, _NSCameras))

"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the SuGaR-style density / level-set queries of
the reference.  Follows, all under /root/reference/dn_splatter/:
  * utils/knn.py:29-43            knn_sk            (sklearn NearestNeighbors with k+1 neighbours, FIRST COLUMN DROPPED
                                                     — also when the queries are not data points: the nearest Gaussian
                                                     of a free sample is discarded, ranks 2..k+1 are returned)
  * dn_model.py:1603-1612         scale_rot_to_inv_cov3d(return_sqrt=True): R(q) diag(1 / clamp(s, 1e-3))
  * dn_model.py:1077-1135         get_density       (sum of sigmoid(opacity) exp(-d_M^2 / 2) over the 16 neighbours,
                                                     d_M^2 clamped to [0, 1e8]; d >= 1 -> d / (d + 1e-5); clamp 1e-4)
  * dn_model.py:1137-1158         get_sdf           sqrt(-2 log density)
  * dn_model.py:1160-1165         get_sdf_weight    mean over the neighbours of min_k exp(scales)
  * dn_model.py:1449-1494         get_density_grad  (the weights are the clamped SQUARED DISTANCES, not exp(-d^2/2))
  * dn_model.py:1207-1447         compute_level_surface_points
  * utils/camera_utils.py:70-144  get_camera_coords / get_means3d_backproj
PINNED: tests/golden/dn_sugar_a.npz was produced by the reference's own functions (tests/golden/make_golden_sugar.py)
and tests/test_sugar_golden.py checks this file against it.  Only tests/ may import this.
"""
from __future__ import annotations

import random
from typing import Dict, Optional, Sequence

import torch
from torch import Tensor

from . import gsplat_ref as G

KNN = 16


def knn_sk(x: Tensor, y: Tensor, k: int) -> Tensor:
    from sklearn.neighbors import NearestNeighbors

    model = NearestNeighbors(n_neighbors=k + 1, algorithm="auto", metric="euclidean").fit(x.detach().cpu().numpy())
    _, idx = model.kneighbors(y.detach().cpu().numpy())
    return torch.from_numpy(idx[:, 1:]).long()


def inv_scaled_rotation(log_scales: Tensor, quats: Tensor) -> Tensor:
    """R(q_hat) * (1 / clamp(exp(s), 1e-3)) broadcast over COLUMNS: the 'square root' of the inverse covariance."""
    inv = 1.0 / torch.exp(log_scales).clamp(min=1e-3)
    return G.quat_to_rotmat(quats) * inv[..., None, :]


def mahalanobis_sq(samples: Tensor, idx: Tensor, p: Dict[str, Tensor]):
    """(clamped squared Mahalanobis distance [M,K], M^T (x - mu) [M,K,3,1], M [M,K,3,3])."""
    Minv = inv_scaled_rotation(p["scales"][idx], p["quats"][idx])
    shift = samples[:, None, :] - p["means"][idx]
    man = Minv.transpose(-1, -2) @ shift[..., None]
    d2 = (man[..., 0] * man[..., 0]).sum(dim=-1).clamp(min=0.0, max=1e8)
    return d2, man, Minv


def raw_density(samples: Tensor, idx: Tensor, p: Dict[str, Tensor]) -> Tensor:
    d2, _, _ = mahalanobis_sq(samples, idx, p)
    w = torch.sigmoid(p["opacities"][idx])[..., 0] * torch.exp(-0.5 * d2)
    dens = w.sum(dim=-1)
    big = dens >= 1.0
    dens = torch.where(big, dens / (dens.detach() + 1e-5), dens)
    return dens


def get_density(samples: Tensor, p: Dict[str, Tensor], idx: Optional[Tensor] = None) -> Tensor:
    if idx is None:
        idx = knn_sk(p["means"], samples, KNN)
    return raw_density(samples, idx, p).clamp(min=1e-4)


def get_sdf(samples: Tensor, p: Dict[str, Tensor], idx: Optional[Tensor] = None) -> Tensor:
    return torch.sqrt(-2.0 * torch.log(get_density(samples, p, idx)))


def get_sdf_weight(idx: Tensor, p: Dict[str, Tensor]) -> Tensor:
    return torch.exp(p["scales"]).min(dim=-1)[0][idx].mean(dim=1)


def get_density_grad(samples: Tensor, p: Dict[str, Tensor], idx: Tensor) -> Tensor:
    d2, man, Minv = mahalanobis_sq(samples, idx, p)
    grad = (d2[..., None] * (Minv @ man)[..., 0]).sum(dim=-2)
    return -torch.nn.functional.normalize(grad, dim=-1)


def backproject(depth: Tensor, fx, fy, cx, cy, W: int, H: int, c2w_cv: Tensor) -> Tensor:
    """Pixel centres (+0.5), x = (u - cx) d / fx, y = (v - cy) d / fy, z = d, then p @ inv(R) + t; [H*W,3] row-major."""
    u = (torch.arange(W, dtype=torch.float32) + 0.5)[None, :].expand(H, W).reshape(-1)
    v = (torch.arange(H, dtype=torch.float32) + 0.5)[:, None].expand(H, W).reshape(-1)
    d = depth.reshape(-1).float()
    cam = torch.stack([(u - cx) * d / fx, (v - cy) * d / fy, d], dim=-1)
    return cam @ torch.linalg.inv(c2w_cv[:3, :3].float()) + c2w_cv[:3, 3].float()


def gaussian_std_along_view(p: Dict[str, Tensor], cam_pos: Tensor) -> Tensor:
    """|| exp(s) * (R(q_hat)^T v) ||, v the unit vector from the Gaussian to the camera (dn_model.py:1264-1271)."""
    view = cam_pos[None, :] - p["means"]
    view = view / view.norm(dim=-1, keepdim=True)
    Rt = G.quat_to_rotmat(p["quats"]).transpose(-1, -2)
    return (torch.exp(p["scales"]) * (Rt @ view[..., None])[..., 0]).norm(dim=-1)


def ray_densities(points: Tensor, idx: Tensor, p: Dict[str, Tensor], cam_pos: Tensor, n_range: int = 21, range_size: float = 3.0):
    """Densities at the 21 samples of every pixel ray (no clamp to 1e-4 here, unlike get_density).  Returns
    (densities [P,21], offsets t [P,21], unit ray directions [P,3])."""
    std = gaussian_std_along_view(p, cam_pos)[idx][..., 0]
    t = torch.linspace(-range_size, range_size, n_range).view(1, -1) * std[:, None]
    dirs = torch.nn.functional.normalize(points - cam_pos[None, :], dim=-1)
    samples = (points[:, None, :] + t[..., None] * dirs[:, None, :]).reshape(-1, 3)
    sidx = idx[:, None, :].expand(-1, n_range, -1).reshape(-1, idx.shape[1])
    return raw_density(samples, sidx, p).reshape(-1, n_range), t, dirs


def level_crossings(dens: Tensor, t: Tensor, level: float):
    """First sample above the level, linear interpolation with the sample before it; a ray is empty when its first
    sample is not under the level or when nothing is above (dn_model.py:1350-1377).  Returns (keep mask [P], t*)."""
    under, above = dens - level < 0, dens - level > 0
    first = above.float().argmax(dim=-1, keepdim=True)  # first True; 0 when none
    empty = ~under[:, 0] | (first[:, 0] == 0)
    keep = ~empty
    f = first[keep]
    d1, d0 = dens[keep].gather(1, f).view(-1), dens[keep].gather(1, f - 1).view(-1)
    t1, t0 = t[keep].gather(1, f).view(-1), t[keep].gather(1, f - 1).view(-1)
    return keep, (level - d0) / (d1 - d0) * (t1 - t0) + t0


def compute_level_surface_points(p: Dict[str, Tensor], gauss_normals: Tensor, depth: Tensor, rgb: Tensor, c2w: Tensor,
                                 fx, fy, cx, cy, W: int, H: int, num_samples: int, mask: Optional[Tensor] = None,
                                 surface_levels: Sequence[float] = (0.1, 0.3, 0.5), return_normal: str = "closest_gaussian",
                                 knn=knn_sk):
    """`depth` [H,W,1] / `rgb` [H,W,3]: outputs of get_outputs for this camera; `c2w` the nerfstudio [3,4] matrix;
    `gauss_normals` = gauss_params["normals"] as get_outputs left them.  Uses python's `random.sample` like the
    reference (seed it outside)."""
    flip = torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))
    pts = backproject(depth, fx, fy, cx, cy, W, H, c2w.float() @ flip).view(H, W, 3)
    cols = rgb.reshape(H, W, 3)
    if mask is not None:
        pts, depth = pts * mask, depth * mask
    has_depth = ~(depth <= 0.0)[..., 0]
    pts, cols = pts[has_depth], cols[has_depth]
    idx = knn(p["means"], pts, KNN)
    cam_pos = c2w[:3, 3].float()
    dens, t, dirs = ray_densities(pts, idx, p, cam_pos)
    out = {}
    for level in surface_levels:
        keep, ts = level_crossings(dens, t, level)
        xp = pts[keep] + ts[:, None] * dirs[keep]
        if return_normal == "analytical":
            d2, man, Minv = mahalanobis_sq(xp, idx[keep], p)
            w = torch.sigmoid(p["opacities"][idx[keep]])[..., 0] * torch.exp(-0.5 * d2)
            nrm = -torch.nn.functional.normalize((w[..., None] * (Minv @ man)[..., 0]).sum(dim=-2), dim=-1)
        elif return_normal == "closest_gaussian":
            nrm = gauss_normals[idx[keep][:, 0]]
        else:
            raise NotImplementedError
        n = xp.shape[0]
        pick = torch.tensor(random.sample(range(n), num_samples if num_samples < n else n), dtype=torch.long)
        out[level] = {"points": xp[pick], "normals": nrm[pick], "colors": cols[keep][pick]}
    return out

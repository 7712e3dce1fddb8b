"""Seeded synthetic scenes and ring cameras used by bench.py and the tests.

Mirrors the reference's own random initialisation (dn_splatter/dn_model.py:135,153-156,217-218,
1497-1509) with a closed-form stand-in for the 3-NN scale init (dn_model.py:186-189, 204-205).
All tensors are generated on the CPU generator so every device sees identical bits.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

BACKGROUND = (0.1490, 0.1647, 0.2157)  # dn_model.py:161-163


def random_quats(n: int, gen: torch.Generator) -> torch.Tensor:
    """Uniform random unit quaternions (Shoemake), wxyz slot order as dn_model.py:1497-1509."""
    u, v, w = (torch.rand(n, generator=gen) for _ in range(3))
    return torch.stack(
        [
            torch.sqrt(1 - u) * torch.sin(2 * math.pi * v),
            torch.sqrt(1 - u) * torch.cos(2 * math.pi * v),
            torch.sqrt(u) * torch.sin(2 * math.pi * w),
            torch.sqrt(u) * torch.cos(2 * math.pi * w),
        ],
        dim=-1,
    )


def make_scene(n: int, seed: int = 0, sh_degree: int = 3, opacity_profile: str = "trained",
               scale_mult: float = 1.0) -> Dict[str, torch.Tensor]:
    """Raw (pre-activation) Gaussian parameters with the reference's names and shapes."""
    gen = torch.Generator().manual_seed(seed)
    means = (torch.rand(n, 3, generator=gen) - 0.5) * 10
    quats = random_quats(n, gen)
    base = math.log(0.718 * (1000.0 / n) ** (1.0 / 3.0) * scale_mult)
    scales = base + 0.3 * torch.randn(n, 3, generator=gen)
    scales[:, 2] -= math.log(10.0)  # flat-disc init (dn_model.py:204-205)
    if opacity_profile == "trained":
        op = 0.05 + 0.9 * torch.rand(n, 1, generator=gen)
    elif opacity_profile == "init":
        op = torch.full((n, 1), 0.1)
    else:
        op = torch.full((n, 1), float(opacity_profile))
    opacities = torch.logit(op)
    k = (sh_degree + 1) ** 2
    features_dc = torch.rand(n, 3, generator=gen)
    features_rest = 0.1 * torch.randn(n, k - 1, 3, generator=gen)
    return {
        "means": means, "quats": quats, "scales": scales, "opacities": opacities,
        "features_dc": features_dc, "features_rest": features_rest,
    }


def look_at_c2w(pos: torch.Tensor, target: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """nerfstudio/OpenGL camera_to_world [3,4]: camera looks down -z, +y up, +x right."""
    fwd = target - pos
    fwd = fwd / fwd.norm()
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    true_up = torch.linalg.cross(right, fwd)
    return torch.stack([right, true_up, -fwd, pos], dim=1)


def ring_cameras(n_views: int, width: int, height: int) -> List[Dict]:
    """View i on a ring of radius 8 looking at the origin, fx = fy = 0.9 W."""
    cams = []
    for i in range(n_views):
        th = 2 * math.pi * i / n_views
        pos = torch.tensor([8 * math.cos(th), 8 * math.sin(th), 2 * math.sin(3 * th)], dtype=torch.float32)
        c2w = look_at_c2w(pos, torch.zeros(3), torch.tensor([0.0, 0.0, 1.0]))
        cams.append({
            "c2w": c2w, "fx": 0.9 * width, "fy": 0.9 * width, "cx": width / 2.0, "cy": height / 2.0,
            "width": width, "height": height,
        })
    return cams

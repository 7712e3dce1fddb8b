"""Shared helpers for the parity tests: run the oracle (CPU) and the CUDA path on the same seeded scene."""
import torch

from dn_splatter_b200.synthetic import BACKGROUND, make_scene, ring_cameras
from oracle import dn_ref


def scene_and_camera(n, width, height, view=1, n_views=5, seed=0, sh_degree=3, **kw):
    params = make_scene(n, seed=seed, sh_degree=sh_degree, **kw)
    cam = ring_cameras(n_views, width, height)[view]
    return params, cam


def oracle_outputs(params, cam, dtype=torch.float32, requires_grad=False, **kw):
    p = {k: v.detach().clone().to(dtype).requires_grad_(requires_grad) for k, v in params.items()}
    out = dn_ref.get_outputs(p, cam["c2w"].to(dtype), cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["width"],
                             cam["height"], torch.tensor(BACKGROUND, dtype=dtype), **kw)
    return p, out


def cuda_outputs(params, cam, requires_grad=False, device="cuda", viewmat=None, **kw):
    """`viewmat`: pass the oracle's own matrix when bit-exactness of integer outputs is asserted (a GPU matmul
    rounds the translation column differently from the CPU one)."""
    from dn_splatter_b200 import dn_rasterize, get_viewmat

    p = {k: v.detach().clone().to(device).requires_grad_(requires_grad) for k, v in params.items()}
    c2w = cam["c2w"].to(device)
    K = torch.tensor([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], dtype=torch.float32, device=device)
    out = dn_rasterize(p["means"], p["quats"], p["scales"], p["opacities"], p["features_dc"], p["features_rest"],
                       get_viewmat(c2w) if viewmat is None else viewmat.to(device), K, cam["width"], cam["height"],
                       background=BACKGROUND, c2w=c2w, **kw)
    return p, out


def frac_close(a, b, atol, rtol=0.0):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    ok = (a - b).abs() <= atol + rtol * b.abs()
    return float(ok.double().mean()), float((a - b).abs().max())

"""BASELINE.json configs[2] / configs[3] sizes (3M Gaussians @1080p with normals; 6M Gaussians @2560x1440): one
forward + backward each, checked through size-independent properties (finite outputs, alpha in [0,1], sorted tile
lists within int32 range, gradients finite and non-zero, precise-hit == exact lists on the images).  ~2 GB of scene + a few GB of
intersection buffers, a few seconds on a B200."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.mark.parametrize("n,W,H", [(3_000_000, 1920, 1080), (6_000_000, 2560, 1440)])
def test_large_scene_forward_backward(n, W, H):
    from dn_splatter_b200 import dn_rasterize, get_viewmat
    from dn_splatter_b200.synthetic import BACKGROUND, make_scene, ring_cameras

    p = {k: v.cuda().requires_grad_(True) for k, v in make_scene(n, seed=1).items()}
    cam = ring_cameras(8, W, H)[3]
    K = torch.tensor([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], dtype=torch.float32)
    out = dn_rasterize(p["means"], p["quats"], p["scales"], p["opacities"], p["features_dc"], p["features_rest"],
                       get_viewmat(cam["c2w"]), K, W, H, background=BACKGROUND, c2w=cam["c2w"])
    n_isects = out.info["n_isects"]
    assert 0 < n_isects < 2 ** 31
    offs = out.info["tile_offsets"].long()
    assert int(offs[0]) == 0 and int(offs[-1]) == n_isects and bool((offs[1:] >= offs[:-1]).all())
    ids = out.info["flatten_ids"]
    assert int(ids.min()) >= 0 and int(ids.max()) < n
    for name in ("rgb", "depth", "normal", "alpha", "surface_normal"):
        t = getattr(out, name)
        assert bool(torch.isfinite(t).all()), name
    assert float(out.alpha.min()) >= 0.0 and float(out.alpha.max()) <= 1.0
    loss = out.rgb.mean() + 0.1 * out.depth.mean() + out.normal.mean()
    loss.backward()
    for k, v in p.items():
        assert v.grad is not None and bool(torch.isfinite(v.grad).all()), k
    assert float(p["means"].grad.abs().sum()) > 0 and float(p["quats"].grad.abs().sum()) > 0
    with torch.no_grad():
        exact = dn_rasterize(p["means"], p["quats"], p["scales"], p["opacities"], p["features_dc"], p["features_rest"],
                             get_viewmat(cam["c2w"]), K, W, H, background=BACKGROUND, c2w=cam["c2w"], exact_lists=True)
    assert torch.equal(exact.rgb, out.rgb) and torch.equal(exact.normal, out.normal)
    del out, exact, loss
    torch.cuda.empty_cache()

"""Size-independent properties at BASELINE.json's full C2 size (1M Gaussians, 1920x1080), where the CPU oracle cannot
run: exact-vs-precise list equivalence, sortedness and tie order of the intersection lists, range of last_ids,
output ranges, linearity of the backward in the upstream gradient, and flat-bucket == autograd gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu
needs_cuda = pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")

N, W, H = 1_000_000, 1920, 1080


@pytest.fixture(scope="module")
def scene():
    from dn_splatter_b200 import get_viewmat
    from dn_splatter_b200.synthetic import BACKGROUND, make_scene, ring_cameras

    params = {k: v.cuda() for k, v in make_scene(N, seed=0).items()}
    cam = ring_cameras(200, W, H)[17]
    K = torch.tensor([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], dtype=torch.float32)
    c2w = cam["c2w"]
    return params, get_viewmat(c2w), K, c2w, BACKGROUND


def _render(scene, requires_grad=False, **kw):
    from dn_splatter_b200 import dn_rasterize

    params, vm, K, c2w, bg = scene
    p = {k: v.detach().clone().requires_grad_(requires_grad) for k, v in params.items()}
    out = dn_rasterize(p["means"], p["quats"], p["scales"], p["opacities"], p["features_dc"], p["features_rest"], vm, K, W, H,
                       background=bg, c2w=c2w, **kw)  # host camera (by value)
    return p, out


@needs_cuda
def test_fullsize_lists_and_images(scene):
    _, full = _render(scene, exact_lists=True)
    _, cut = _render(scene)
    for name in ("rgb", "depth", "normal", "alpha", "surface_normal"):
        assert torch.equal(getattr(full, name), getattr(cut, name)), name
    assert int(full.tiles_per_gauss.sum()) == full.info["n_isects"] > cut.info["n_isects"] > 0
    for out in (full, cut):
        ids, offs = out.info["flatten_ids"].long(), out.info["tile_offsets"].long()
        n_i = out.info["n_isects"]
        assert int(offs[0]) == 0 and int(offs[-1]) == n_i and bool((offs[1:] >= offs[:-1]).all())
        tile_of = torch.repeat_interleave(torch.arange(offs.numel() - 1, device=offs.device), offs[1:] - offs[:-1])
        d = out.depths[ids]
        same = tile_of[1:] == tile_of[:-1]
        assert not bool(((d[1:] < d[:-1]) & same).any()), "depth order inside a tile"
        ties = (d[1:] == d[:-1]) & same
        assert not bool(((ids[1:] <= ids[:-1]) & ties).any()), "ties keep ascending Gaussian index"
        assert bool((out.radii[ids] > 0).all())
        # last_ids of covered pixels point inside their own tile's slice
        last = out.info["last_ids"].long()
        lt = out.info["list_tile"]  # 16 for the exact lists, 16 << list_shift (supertile lists) otherwise
        ty = torch.arange(H, device=last.device)[:, None] // lt
        tx = torch.arange(W, device=last.device)[None, :] // lt
        t = ty * out.info["lists_x"] + tx
        cov = out.alpha[..., 0] > 0
        assert bool(((last >= offs[t]) & (last < offs[t + 1]))[cov].all())
    assert float(cut.alpha.min()) >= 0.0 and float(cut.alpha.max()) < 1.0
    assert float(cut.rgb.min()) >= 0.0 and float(cut.rgb.max()) <= 1.0
    nrm = (2 * cut.normal - 1).norm(dim=-1)
    assert float((nrm - 1).abs().max()) < 1e-4
    assert torch.isfinite(cut.depth).all() and torch.isfinite(cut.surface_normal).all()


@needs_cuda
def test_fullsize_backward_is_linear_and_bucket_matches_autograd(scene):
    g = torch.Generator().manual_seed(0)
    w = {k: torch.rand(s, generator=g).cuda() for k, s in (("rgb", (H, W, 3)), ("depth", (H, W, 1)), ("normal", (H, W, 3)),
                                                            ("alpha", (H, W, 1)))}
    p, out = _render(scene, requires_grad=True)
    loss = sum((getattr(out, k) * w[k]).sum() for k in w) * 1e-3
    names = list(p)
    g1 = torch.autograd.grad(loss, [p[k] for k in names], retain_graph=True)
    g3 = torch.autograd.grad(3.0 * loss, [p[k] for k in names])
    for k, a, b in zip(names, g1, g3):
        assert torch.isfinite(a).all(), k
        rel = float((3.0 * a - b).norm() / (b.norm() + 1e-30))
        assert rel < 1e-4, (k, rel)  # float atomics reorder, nothing else
    # invisible Gaussians get exactly zero gradient
    inv = out.radii <= 0
    assert float(g1[names.index("means")][inv].abs().max()) == 0.0
    # flat-bucket (grad-sink) path == autograd path
    from dn_splatter_b200.parallel import FlatGradBucket

    p2, _ = _render(scene, requires_grad=False)
    leaf = {k: torch.nn.Parameter(v) for k, v in p2.items()}
    bucket = FlatGradBucket(leaf)
    from dn_splatter_b200 import dn_rasterize

    _, vm, K, c2w, bg = scene
    out2 = dn_rasterize(leaf["means"], leaf["quats"], leaf["scales"], leaf["opacities"], leaf["features_dc"],
                        leaf["features_rest"], vm, K, W, H, background=bg, c2w=c2w, grad_sink=bucket.sink())
    (sum((getattr(out2, k) * w[k]).sum() for k in w) * 1e-3).backward()
    for k, a in zip(names, g1):
        rel = float((leaf[k].grad - a).norm() / (a.norm() + 1e-30))
        assert rel < 1e-4, (k, rel)


@needs_cuda
def test_fullsize_window_matches_oracle(scene):
    """at the full C2 size the CUDA path was only compared with itself.  Here a tile-aligned 256x192 window
    around the principal point of the 1M-Gaussian 1080p frame is rendered by the CPU oracle (same camera, principal point
    shifted into the window, the full frame's 1.3 tan(fov) clamp) and compared with the same window of the CUDA render:
    images to the parity tolerances of tests/test_gpu_parity.py, and the gradient of a window-supported loss w.r.t. every
    parameter of the Gaussians that reach the window."""
    from dn_splatter_b200.synthetic import BACKGROUND
    from oracle import dn_ref

    torch.set_num_threads(min(8, torch.get_num_threads()))
    params, vm, K, c2w, bg = scene
    x0, y0, ww, wh = 832, 448, 256, 192  # multiples of 16: the window's tiles are the frame's tiles
    g = torch.Generator().manual_seed(3)
    wts = {k: torch.rand(wh, ww, c, generator=g) for k, c in (("rgb", 3), ("depth", 1), ("normal", 3), ("alpha", 1))}

    p, out = _render(scene, requires_grad=True)
    win = lambda t: t[y0:y0 + wh, x0:x0 + ww]  # noqa: E731
    loss = sum((win(getattr(out, k)) * wts[k].cuda()).sum() for k in wts) * 1e-3
    loss.backward()

    pc = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in params.items()}
    cam = dict(fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]) - x0, cy=float(K[1, 2]) - y0)
    ref = dn_ref.get_outputs(pc, c2w.cpu(), cam["fx"], cam["fy"], cam["cx"], cam["cy"], ww, wh, torch.tensor(BACKGROUND),
                             fov_size=(W, H))
    from tests.helpers import frac_close

    for name, got, want in (("rgb", out.rgb, ref["rgb"]), ("normal", out.normal, ref["normal"]),
                            ("alpha", out.alpha, ref["accumulation"])):
        frac, mx = frac_close(win(got), want, atol=1e-4)
        assert frac >= 0.999 and mx <= 2e-2, (name, frac, mx)
    cov = ref["accumulation"] > 0  # the depth fill uses the frame-wide maximum: compare where something was composited
    frac, mx = frac_close(win(out.depth)[cov], ref["depth"][cov], atol=1e-4, rtol=1e-5)
    assert frac >= 0.999, ("depth", frac, mx)
    # gradients of the window loss (the fill value is detached, uncovered pixels contribute no depth gradient)
    lref = ((ref["rgb"] * wts["rgb"]).sum() + (torch.where(cov, ref["depth"], torch.zeros(())) * wts["depth"]).sum()
            + (ref["normal"] * wts["normal"]).sum() + (ref["accumulation"] * wts["alpha"]).sum()) * 1e-3
    # same masking on the CUDA side: recompute its loss with the uncovered depth pixels dropped
    for v in p.values():
        v.grad = None
    p2, out2 = _render(scene, requires_grad=True)
    covc = cov.cuda()
    loss2 = ((win(out2.rgb) * wts["rgb"].cuda()).sum() + (torch.where(covc, win(out2.depth), torch.zeros((), device="cuda")) * wts["depth"].cuda()).sum()
             + (win(out2.normal) * wts["normal"].cuda()).sum() + (win(out2.alpha) * wts["alpha"].cuda()).sum()) * 1e-3
    loss2.backward()
    lref.backward()
    for k in ("means", "quats", "scales", "opacities", "features_dc", "features_rest"):
        gc, gr = p2[k].grad.cpu(), pc[k].grad
        rel = float((gc - gr).norm() / (gr.norm() + 1e-30))
        assert rel < 5e-3, (k, rel)

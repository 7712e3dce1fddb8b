"""Round-2 first call: times every kernel that was written after round 1's GPU budget ran out against the path it is
meant to replace, so one gpurun call decides which of them becomes the default.

    python scripts/exp_bench.py [--n 1000000] [--reps 20] > gpurun_out/exp_bench.jsonl

Rows (one JSON line each):
  ssim        FusedSSIM fwd+bwd  vs  dn_model.ssim (torch convs + autograd) at 1920x1080x3
  adam        FusedAdam (1 launch) vs  7 x torch.optim.Adam (default foreach) and fused=True, N Gaussians, SH degree 3
  project_bwd default vs DNR_FLAG_COMPACT_BWD on the bench scene (stage events)
Each row also checks agreement with the reference path (max abs / rel error), so a faster-but-wrong kernel is visible.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--only", default="")
args = ap.parse_args()


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def emit(row):
    print(json.dumps(row), flush=True)


def bench_ssim():
    from dn_splatter_b200.dn_model import ssim
    from dn_splatter_b200.regularization_strategy import FusedSSIM

    H, W = 1080, 1920
    g = torch.Generator().manual_seed(0)
    x = torch.rand(H, W, 3, generator=g).cuda().requires_grad_(True)
    y = (x.detach().cpu() * 0.7 + 0.3 * torch.rand(H, W, 3, generator=g)).cuda()

    def ref():
        s = ssim(y.permute(2, 0, 1)[None], x.permute(2, 0, 1)[None])
        return s, torch.autograd.grad(s, x)[0]

    def fused():
        s = FusedSSIM.apply(x, y)
        return s, torch.autograd.grad(s, x)[0]

    (sr, gr), (sf, gf) = ref(), fused()
    emit({"row": "ssim", "torch_ms": timed(ref, args.reps), "fused_ms": timed(fused, args.reps),
          "value_abs_err": abs(float(sr) - float(sf)), "grad_rel_err": float((gr - gf).norm() / gr.norm())})


def bench_adam():
    from dn_splatter_b200.optim import FusedAdam

    n = args.n
    shapes = {"means": (n, 3), "scales": (n, 3), "quats": (n, 4), "features_dc": (n, 3), "features_rest": (n, 15, 3),
              "opacities": (n, 1)}
    lrs = {"means": 1.6e-4, "scales": 5e-3, "quats": 1e-3, "features_dc": 2.5e-3, "features_rest": 1.25e-4, "opacities": 5e-2}
    g = torch.Generator(device="cuda").manual_seed(0)

    def make():
        torch.manual_seed(0)
        ps = {k: torch.nn.Parameter(torch.randn(*s, device="cuda", generator=g)) for k, s in shapes.items()}
        for p in ps.values():
            p.grad = torch.randn(p.shape, device="cuda", generator=g) * 1e-3
        return ps

    a = make()
    fused = FusedAdam([{"params": [p], "lr": lrs[k], "eps": 1e-15, "name": k} for k, p in a.items()])
    b = make()
    for k in a:
        b[k].data.copy_(a[k].data)
        b[k].grad.copy_(a[k].grad)
    ref = [torch.optim.Adam([p], lr=lrs[k], eps=1e-15) for k, p in b.items()]
    fused.step()
    for o in ref:
        o.step()
    err = max(float((a[k] - b[k]).abs().max()) for k in a)
    c = make()
    ref_fused = [torch.optim.Adam([p], lr=lrs[k], eps=1e-15, fused=True) for k, p in c.items()]
    floats = sum(p.numel() for p in a.values())
    t_f = timed(fused.step, args.reps)
    emit({"row": "adam", "n_gauss": n, "floats": floats, "fused_ms": t_f,
          "torch_foreach_ms": timed(lambda: [o.step() for o in ref], args.reps),
          "torch_fused_ms": timed(lambda: [o.step() for o in ref_fused], args.reps),
          "fused_GBps": floats * 28 / t_f / 1e6, "param_abs_err_after_1_step": err})


def bench_project_bwd():
    import dn_splatter_b200.rasterize as R
    from dn_splatter_b200 import dn_rasterize, get_viewmat
    from dn_splatter_b200.synthetic import BACKGROUND, make_scene, ring_cameras

    W, H = 1920, 1080
    cam = ring_cameras(8, W, H)[3]
    K = torch.tensor([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], dtype=torch.float32)
    vm = get_viewmat(cam["c2w"])
    res = {}
    grads = {}
    for compact in (False, True):
        p = {k: v.cuda().requires_grad_(True) for k, v in make_scene(args.n, seed=0).items()}
        out = dn_rasterize(p["means"], p["quats"], p["scales"], p["opacities"], p["features_dc"], p["features_rest"], vm, K, W, H,
                           background=BACKGROUND, c2w=cam["c2w"], compact_bwd=compact)
        loss = (out.rgb.sum() + out.depth.sum() * 0.1 + out.normal.sum()) * 1e-3
        R.STAGE_EVENTS = []
        for _ in range(args.reps):
            gr = torch.autograd.grad(loss, list(p.values()), retain_graph=True)
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for name, a, b in R.STAGE_EVENTS if name == "project_bwd")
        R.STAGE_EVENTS = None
        res[compact] = t[len(t) // 2]
        grads[compact] = gr
    rel = max(float((x - y).norm() / (x.norm() + 1e-30)) for x, y in zip(grads[False], grads[True]))
    emit({"row": "project_bwd", "default_ms": res[False], "compact_ms": res[True], "grad_rel_err": rel})


def bench_knn():
    import time

    from dn_splatter_b200.sugar import KnnIndex
    from oracle import sugar_ref as S  # sklearn, the reference's backend: timed on the host as the baseline

    n, m = args.n, 200_000
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 3, generator=g) * torch.tensor([3.0, 3.0, 0.5])
    y = x[torch.randint(0, n, (m,), generator=g)] + 0.01 * torch.randn(m, 3, generator=g)
    xc, yc = x.cuda(), y.cuda()
    t_build = timed(lambda: KnnIndex(xc), 5)
    index = KnnIndex(xc)
    t_query = timed(lambda: index.query(yc, 16), 5)
    t0 = time.time()
    want = S.knn_sk(x, y[:20_000], 16)
    t_sk = (time.time() - t0) * 1e3 * (m / 20_000)
    got = index.query(yc[:20_000], 16).cpu()
    emit({"row": "knn", "n_points": n, "n_queries": m, "build_ms": t_build, "query_ms": t_query,
          "sklearn_ms_extrapolated": t_sk, "index_agreement": float((got == want).float().mean())})


def bench_render_service():
    """forward-only render-all-views loop (what gs-mesh / ns-eval do), maps kept on the device or copied
    to pinned host buffers."""
    import time

    from dn_splatter_b200.cameras import Cameras
    from dn_splatter_b200.dn_model import DNSplatterModelConfig
    from dn_splatter_b200.render_service import ViewRenderer
    from dn_splatter_b200.synthetic import make_scene, ring_cameras

    W, H, n_views = 1920, 1080, 48
    m = DNSplatterModelConfig(random_init=True, num_random=16, background_color="black", sync_free=True).setup(device="cuda")
    m.load_gaussians(make_scene(args.n, seed=0))
    m.step = 30000
    cams = [Cameras(c["c2w"][None], c["fx"], c["fy"], c["cx"], c["cy"], W, H) for c in ring_cameras(n_views, W, H)]
    row = {"row": "render_service", "n_gauss": args.n, "views": n_views, "resolution": f"{W}x{H}"}
    for to_host in (False, True):
        r = ViewRenderer(m, to_host=to_host)
        for _ in r.render(cams[:8]):  # warm-up (also seeds the sync-free capacity)
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in r.render(cams):
            pass
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_views
        row["ms_per_view_host" if to_host else "ms_per_view_device"] = dt * 1e3
        row["mpix_s_host" if to_host else "mpix_s_device"] = W * H / 1e6 / dt
    emit(row)


for name, fn in (("ssim", bench_ssim), ("adam", bench_adam), ("project_bwd", bench_project_bwd), ("knn", bench_knn),
                 ("render_service", bench_render_service)):
    if args.only and name not in args.only.split(","):
        continue
    try:
        fn()
    except Exception as e:  # one broken experimental kernel must not hide the others
        emit({"row": name, "error": f"{type(e).__name__}: {e}"})

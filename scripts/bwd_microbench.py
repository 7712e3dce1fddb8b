"""BASELINE.json configs[4]: backward microbench — 2M Gaussians, 1920x1080, one fixed view; sweep the
scale multiplier (-> intersections per Gaussian) and the opacity profile, time raster_bwd alone with CUDA events and
report the per-Gaussian gradient-scatter rate G * I_c / t against the measured HBM peak.

    python scripts/bwd_microbench.py [--n 2000000] [--reps 10]

Results: profiles/bwd_microbench_r2*.jsonl.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dn_splatter_b200.rasterize as R  # noqa: E402
from dn_splatter_b200 import dn_rasterize, get_viewmat  # noqa: E402
from dn_splatter_b200.synthetic import BACKGROUND, make_scene, ring_cameras  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
W, H = 1920, 1080
cam = ring_cameras(1, W, H)[0]
K = torch.tensor([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], dtype=torch.float32)
vm = get_viewmat(cam["c2w"])
peak = 6555.5
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
g = torch.Generator().manual_seed(0)
w = {k: torch.rand(s, generator=g).cuda() for k, s in (("rgb", (H, W, 3)), ("depth", (H, W, 1)), ("normal", (H, W, 3)))}
rows = []
for scale_mult in (0.25, 0.5, 1.0, 2.0, 4.0):
    for op in ("0.1", "trained", "0.99"):
        p = {k: v.cuda().requires_grad_(True) for k, v in make_scene(args.n, seed=0, opacity_profile=op, scale_mult=scale_mult).items()}
        stats = torch.zeros(4, dtype=torch.int64, device="cuda")
        out = dn_rasterize(p["means"], p["quats"], p["scales"], p["opacities"], p["features_dc"], p["features_rest"], vm, K, W, H,
                           background=BACKGROUND, c2w=cam["c2w"], stats=stats)
        loss = sum((getattr(out, k) * w[k]).sum() for k in w) * 1e-3
        R.STAGE_EVENTS = []
        for _ in range(args.reps):
            torch.autograd.grad(loss, list(p.values()), retain_graph=True)
        torch.cuda.synchronize()
        t = [a.elapsed_time(b) for name, a, b in R.STAGE_EVENTS if name == "raster_bwd"][2:]
        R.STAGE_EVENTS = None
        ms = sorted(t)[len(t) // 2]
        info = out.info
        # (tile, Gaussian) pairs the backward composites = list entries kept by the tile filter up to each tile's deepest
        # last id (device counter stats[3], summed over the backward launches of this configuration)
        i_c = int(stats[3]) // args.reps
        G = 60  # bytes of reduced gradient record per composited intersection (48 + 12 normals)
        rows.append({"scale_mult": scale_mult, "opacity": op, "n_isects": info["n_isects"], "n_composited": i_c,
                     "isects_per_gauss": info["n_isects"] / args.n, "raster_bwd_ms": ms,
                     "scatter_GBps": G * i_c / ms / 1e6, "frac_of_hbm_peak": G * i_c / ms / 1e6 / peak})
        print(json.dumps(rows[-1]), flush=True)

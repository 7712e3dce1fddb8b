"""How much of raster_bwd's (pixel, record) work is live?  For sampled tiles of the bench scene: the fraction of
(pixel, record <= tile's deepest last_id) pairs that contribute (alpha >= 1/255 and record <= the pixel's last_id), and
for several warp footprints the number of (warp, record) steps a pixel-parallel kernel executes (a step runs when any
pixel of the footprint is live; records deeper than the footprint's own deepest last_id are skipped outright).
    python scripts/pair_stats.py [--n-gauss 1000000] [--tiles 600]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dn_splatter_b200 import dn_rasterize, get_viewmat  # noqa: E402
from dn_splatter_b200.synthetic import BACKGROUND, make_scene, ring_cameras  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-gauss", type=int, default=1_000_000)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--tiles", type=int, default=600)
ap.add_argument("--view", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda")
p = {k: v.to(dev) for k, v in make_scene(args.n_gauss, seed=0).items()}
cam = ring_cameras(200, args.width, args.height)[args.view]
c2w = cam["c2w"].to(dev)
K = torch.tensor([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], device=dev)
with torch.no_grad():
    out = dn_rasterize(p["means"], p["quats"], p["scales"], p["opacities"], p["features_dc"], p["features_rest"],
                       get_viewmat(c2w), K, args.width, args.height, background=BACKGROUND, c2w=c2w)
info = out.info
ids, offs, last = info["flatten_ids"].long(), info["tile_offsets"].long(), info["last_ids"].long()
m2d, con, op = out.means2d, out.conics, info["opacities"]
H, W = args.height, args.width
tx = info["tile_width"]
g = torch.Generator().manual_seed(0)
n_tiles = offs.numel() - 1
sample = torch.randperm(n_tiles, generator=g)[: args.tiles].tolist()
acc = dict(pairs=0, live=0, recs=0, tiles=0, listed=0)
shapes = {"4x4": (4, 4), "8x4": (8, 4), "8x8": (8, 8), "16x4": (16, 4), "16x8": (16, 8), "16x16": (16, 16)}
steps = {k: 0 for k in shapes}
steps_in_range = {k: 0 for k in shapes}
for t in sample:
    ty, txi = divmod(t, tx)
    y0, x0 = ty * 16, txi * 16
    if y0 + 16 > H or x0 + 16 > W:
        continue
    s, e = int(offs[t]), int(offs[t + 1])
    lt = last[y0:y0 + 16, x0:x0 + 16]
    hi = min(int(lt.max()) + 1, e)
    if hi <= s:
        continue
    gi = ids[s:hi]
    idx = torch.arange(s, hi, device=dev)
    py = (torch.arange(16, device=dev) + y0 + 0.5)[:, None, None]
    px = (torch.arange(16, device=dev) + x0 + 0.5)[None, :, None]
    dx, dy = m2d[gi, 0][None, None] - px, m2d[gi, 1][None, None] - py
    sig = 0.5 * (con[gi, 0] * dx * dx + con[gi, 2] * dy * dy) + con[gi, 1] * dx * dy
    alpha = torch.clamp(op[gi][None, None] * torch.exp(-sig), max=0.999)
    live = (sig >= 0) & (alpha >= 1 / 255.0) & (idx[None, None] <= lt[:, :, None])  # [16,16,R]
    acc["pairs"] += live.numel()
    acc["live"] += int(live.sum())
    acc["recs"] += hi - s
    acc["listed"] += e - s
    acc["tiles"] += 1
    for name, (w, h) in shapes.items():
        lv = live.view(16 // h, h, 16 // w, w, -1).permute(0, 2, 1, 3, 4).reshape(-1, h * w, hi - s)
        steps[name] += int(lv.any(dim=1).sum())
        lw = lt.view(16 // h, h, 16 // w, w).permute(0, 2, 1, 3).reshape(-1, h * w).amax(dim=1)  # footprint's deepest
        steps_in_range[name] += int((idx[None] <= lw[:, None]).sum())
res = {"tiles": acc["tiles"], "listed_records_per_tile": acc["listed"] / acc["tiles"],
       "records_to_deepest_last_id_per_tile": acc["recs"] / acc["tiles"],
       "live_pair_fraction": acc["live"] / acc["pairs"], "live_pairs_per_tile": acc["live"] / acc["tiles"]}
for name, (w, h) in shapes.items():
    res[f"steps_per_tile_{name}"] = steps[name] / acc["tiles"]
    res[f"pixel_slots_per_live_pair_{name}"] = steps[name] * w * h / acc["live"]
    res[f"steps_in_range_per_tile_{name}"] = steps_in_range[name] / acc["tiles"]
print(json.dumps(res))

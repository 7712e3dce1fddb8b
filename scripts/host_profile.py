"""Host-side time per phase of one training step (how much Python/launch time hides behind the GPU)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sync", action="store_true")
ap.add_argument("--steps", type=int, default=12)
a = ap.parse_args()
sys.argv = [sys.argv[0]] + (["--sync"] if a.sync else [])
args = bench.parse()
dev = torch.device("cuda:0")
model, bucket, cams = bench.build_workload(args, dev, True)
sets = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_gt_sets(model, cams, args, True, 2)]
for s in range(4):
    bench.run_step(model, bucket, cams[s], sets[s % 2])
torch.cuda.synchronize()
acc = {}
t_all = time.perf_counter()
for s in range(a.steps):
    t0 = time.perf_counter()
    bucket.zero_()
    out = model.get_outputs(cams[s])
    t1 = time.perf_counter()
    ld = model.get_loss_dict(out, dict(sets[s % 2]))
    loss = ld["main_loss"] + ld["scale_reg"]
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    for k, v in (("get_outputs", t1 - t0), ("get_loss_dict", t2 - t1), ("backward", t3 - t2)):
        acc[k] = acc.get(k, 0) + v
torch.cuda.synchronize()
wall = (time.perf_counter() - t_all) / a.steps
print({k: round(1e3 * v / a.steps, 3) for k, v in acc.items()}, "host ms/step;", "wall ms/step", round(1e3 * wall, 3),
      "sync" if a.sync else "sync_free")

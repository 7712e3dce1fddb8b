"""Repeats bench.py's device-resident timed loop several times in one process to separate warm-up effects from noise."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
sys.argv = [sys.argv[0]]
args = bench.parse()
dev = torch.device("cuda:0")
model, bucket, cams = bench.build_workload(args, dev, True)
sets = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_gt_sets(model, cams, args, True, 4)]
with torch.no_grad():
    for c in cams:
        model.get_outputs(c)
for s in range(5):
    bench.run_step(model, bucket, cams[s], sets[s % 4])
torch.cuda.synchronize()
for trial in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    host = []
    for s in range(40):
        h0 = time.perf_counter()
        bench.run_step(model, bucket, cams[(trial * 40 + s) % len(cams)], sets[s % 4])
        host.append(time.perf_counter() - h0)
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    host.sort()
    print(f"trial {trial}: gpu {e0.elapsed_time(e1)/40:.3f} ms/step, host issue {1e3*t_issue/40:.3f} ms/step, "
          f"host median {1e3*host[20]:.3f} max {1e3*host[-1]:.3f}")

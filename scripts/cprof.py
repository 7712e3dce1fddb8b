import cProfile, pstats, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
extra = [a for a in sys.argv[1:]]
sys.argv = [sys.argv[0]] + extra
args = bench.parse()
dev = torch.device("cuda:0")
model, bucket, cams = bench.build_workload(args, dev, True)
sets = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_gt_sets(model, cams, args, True, 2)]
for s in range(4):
    bench.run_step(model, bucket, cams[s], sets[s % 2])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for s in range(10):
    bench.run_step(model, bucket, cams[s], sets[s % 2])
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)

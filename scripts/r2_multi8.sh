#!/bin/bash
# round 2, 8-GPU call: C2 at N=8 (peer-memory reduce fused into Adam vs dense NCCL all-reduce), N=4, and C4 (6M Gaussians, 1440p)
mkdir -p gpurun_out
run() { # tag nproc extra...
  tag=$1; np=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --skip-cpu-baseline "$@" > gpurun_out/r2m8_$tag.json 2> gpurun_out/r2m8_$tag.err
  echo "$tag rc=$?"
}
run peer8 8 --epochs 2
run nccl8 8 --epochs 0 --skip-e2e --nccl-allreduce
run c4_peer8 8 --epochs 0 --n-gauss 6000000 --width 2560 --height 1440 --views 1000 --gt-sets 4
python - <<'PY'
import json
for f in ('peer8','nccl8','c4_peer8'):
    try:
        d=json.loads(open(f'gpurun_out/r2m8_{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), (d['e2e'] or {}).get('value'), d['epochs'], d['cuda_graph'], {k:round(v,3) for k,v in d['stages_ms'].items()})
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2m8_peer8.err | cut -c1-300

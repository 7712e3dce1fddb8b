#!/bin/bash
# round 2, 4-GPU call: the 2-GPU peer-reduce test, then C2 at N=4 (peer-memory reduce fused into Adam vs dense NCCL all-reduce)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2m4_tests.log 2>&1; echo "multi tests rc=$?"; grep -v "^$" gpurun_out/r2m4_tests.log | grep -iv "warning" | tail -4 | cut -c1-300
run() { # tag nproc extra...
  tag=$1; np=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --skip-cpu-baseline "$@" > gpurun_out/r2m4_$tag.json 2> gpurun_out/r2m4_$tag.err
  echo "$tag rc=$?"
}
run peer4 4 --epochs 0
run nccl4 4 --epochs 0 --skip-e2e --nccl-allreduce
python - <<'PY'
import json
for f in ('peer4','nccl4'):
    try:
        d=json.loads(open(f'gpurun_out/r2m4_{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), (d['e2e'] or {}).get('value'), d['epochs'], d['cuda_graph'], {k:round(v,3) for k,v in d['stages_ms'].items()})
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2m4_peer4.err | cut -c1-300

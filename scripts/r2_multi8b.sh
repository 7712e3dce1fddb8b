#!/bin/bash
# round 2: re-validate the peer-memory reduce (all peer loads in flight) — 2-GPU equality test, then C2 at N=8 and N=2
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2m8b_tests.log 2>&1; echo "multi tests rc=$?"; grep -v "^$" gpurun_out/r2m8b_tests.log | grep -iv "warning" | tail -4 | cut -c1-300
run() { # tag nproc extra...
  tag=$1; np=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np --steps 20 --warmup 5 --skip-cpu-baseline "$@" > gpurun_out/r2m8b_$tag.json 2> gpurun_out/r2m8b_$tag.err
  echo "$tag rc=$?"
}
run peer8 8 --epochs 2
run peer4 4 --epochs 0 --skip-e2e
run peer2 2 --epochs 0 --skip-e2e
python - <<'PY'
import json
for f in ('peer8','peer4','peer2'):
    try:
        d=json.loads(open(f'gpurun_out/r2m8b_{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), (d['e2e'] or {}).get('value'), d['epochs'], d['cuda_graph'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2m8b_peer8.err | cut -c1-300

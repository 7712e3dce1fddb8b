#!/bin/bash
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2m2e_tests.log 2>&1; echo "multi tests rc=$?"; grep -v "^$" gpurun_out/r2m2e_tests.log | grep -iv "warning" | tail -12 | cut -c1-300
TR="timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu-baseline --epochs 2 > gpurun_out/r2m2e_peer.json 2> gpurun_out/r2m2e_peer.err; echo "peer rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m2e_peer.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d['e2e']['value'], d['epochs'], d['cuda_graph'])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2m2c_tests.log 2>&1; echo "multi tests rc=$?"; tail -8 gpurun_out/r2m2c_tests.log | cut -c1-250
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
DNR_DEBUG_CAPTURE=1 timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu-baseline --epochs 2 > gpurun_out/r2m2c_peer.json 2> gpurun_out/r2m2c_peer.err; echo "peer rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m2c_peer.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d['e2e']['value'], d['epochs'], d['cuda_graph'])
PY
grep -i "invalidated" gpurun_out/r2m2c_peer.err | head -3

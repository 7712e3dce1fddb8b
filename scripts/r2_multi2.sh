#!/bin/bash
# round 2, 2-GPU call: peer-memory reduce+Adam test, then the bench at N=2 with the fused peer reduction and with NCCL
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2m2_tests.log 2>&1; echo "multi tests rc=$?"; tail -15 gpurun_out/r2m2_tests.log | cut -c1-250
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu-baseline --epochs 2 > gpurun_out/r2m2_peer.json 2> gpurun_out/r2m2_peer.err; echo "peer rc=$?"; tail -c 1500 gpurun_out/r2m2_peer.json; tail -5 gpurun_out/r2m2_peer.err | cut -c1-300
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu-baseline --epochs 2 --nccl-allreduce > gpurun_out/r2m2_nccl.json 2> gpurun_out/r2m2_nccl.err; echo "nccl rc=$?"; tail -c 600 gpurun_out/r2m2_nccl.json
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --epochs 2 > gpurun_out/r2m2_n1.json 2> gpurun_out/r2m2_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for f in ('n1','peer','nccl'):
    try:
        d=json.loads(open(f'gpurun_out/r2m2_{f}.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d['e2e']['value'] if d['e2e'] else None, d['epochs'])
    except Exception as e: print(f,'ERR',e)
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/debug_capture2.py > gpurun_out/r2m2d_debug.log 2>&1; echo "debug rc=$?"
grep "rank" gpurun_out/r2m2d_debug.log | cut -c1-600
timeout 280 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2m2d_tests.log 2>&1; echo "multi tests rc=$?"; grep -v "^$" gpurun_out/r2m2d_tests.log | grep -iv "warning" | tail -25 | cut -c1-300

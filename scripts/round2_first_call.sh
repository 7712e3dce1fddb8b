#!/bin/bash
# Round 2, first gpurun call (about 12 GPU-minutes): everything that was written after round 1's GPU budget ran out.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh'
# Every stage runs under its own timeout and writes into gpurun_out/; nothing here changes the default path.
mkdir -p gpurun_out
export DNR_TEST_EXPERIMENTAL=1
timeout 300 python -m pytest tests -q -m gpu -k "experimental or fused_ssim or fused_adam or sugar" > gpurun_out/r2_experimental_tests.log 2>&1
echo "experimental tests rc=$?"
timeout 300 python scripts/exp_bench.py > gpurun_out/r2_exp_bench.jsonl 2> gpurun_out/r2_exp_bench.err
echo "exp_bench rc=$?"; cat gpurun_out/r2_exp_bench.jsonl
DNR_RUN_TRAINING_TEST=1 timeout 400 python -m pytest tests/test_gpu_training.py -q -m gpu > gpurun_out/r2_training_test.log 2>&1
echo "training test rc=$?"; tail -3 gpurun_out/r2_training_test.log
timeout 400 python scripts/bwd_microbench.py --reps 6 > gpurun_out/r2_bwd_microbench.jsonl 2> gpurun_out/r2_bwd_microbench.err
echo "bwd_microbench rc=$?"; tail -3 gpurun_out/r2_bwd_microbench.jsonl
DNR_RUN_SCALE_TEST=1 timeout 300 python -m pytest tests/test_gpu_scale.py -q -m gpu > gpurun_out/r2_scale_test.log 2>&1
echo "scale test rc=$?"; tail -3 gpurun_out/r2_scale_test.log

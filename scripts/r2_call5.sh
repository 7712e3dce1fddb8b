#!/bin/bash
# round 2, call 5: all GPU tests, ncu --set full of raster_bwd / raster_fwd, launch list of a training step, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c5_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2c5_tests.log | cut -c1-220
B="python bench.py --steps 2 --warmup 3 --views 24 --gt-sets 2 --skip-cpu-baseline --skip-e2e --no-graph --epochs 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:raster_bwd_kernel -s 4 -c 1 -f -o gpurun_out/prof_raster_bwd_r2a $B > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"; tail -2 gpurun_out/ncu_bwd.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:raster_fwd_kernel -s 30 -c 1 -f -o gpurun_out/prof_raster_fwd_r2a $B > gpurun_out/ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"; tail -2 gpurun_out/ncu_fwd.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 760 -c 600 --csv --log-file gpurun_out/launches_r2a.csv $B > gpurun_out/launches_r2a.log 2>&1; echo "launch list rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c5_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), d['e2e']['ms_per_step'], {k:round(v,3) for k,v in d['stages_ms'].items()}, d['epochs']['median_ms_per_step'])
PY
ls -la gpurun_out | grep r2a

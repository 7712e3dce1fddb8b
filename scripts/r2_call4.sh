#!/bin/bash
# round 2, call 4: all GPU tests (no -x), ncu --set full of the new raster kernels, launch list, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c4_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2c4_tests.log | cut -c1-220
B="python bench.py --steps 2 --warmup 3 --gt-sets 2 --skip-cpu-baseline --skip-e2e --no-graph --epochs 0"
for k in raster_bwd raster_fwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 210 -c 1 -o gpurun_out/prof_${k}_r2a $B > /dev/null 2>&1
  echo "ncu $k rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r2a.csv $B > gpurun_out/launches_r2a.log 2>&1; echo "launch list rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c4_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), d['e2e'], {k:round(v,3) for k,v in d['stages_ms'].items()}, d['epochs'])
PY

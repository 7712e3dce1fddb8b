#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c7_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2c7_tests.log | cut -c1-220
B="python bench.py --steps 2 --warmup 3 --views 24 --gt-sets 2 --skip-cpu-baseline --skip-e2e --no-graph --epochs 0"
for k in ssim_fwd ssim_bwd; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 4 -c 1 -f -o gpurun_out/prof_${k}_r2a $B > gpurun_out/ncu_${k}.log 2>&1; echo "ncu $k rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 600 --csv --log-file gpurun_out/launches_r2c.csv $B > gpurun_out/launches_r2c.log 2>&1; echo "launch list rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c7_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), d['e2e']['ms_per_step'], {k:round(v,3) for k,v in d['stages_ms'].items()}, d['epochs']['median_ms_per_step'])
PY

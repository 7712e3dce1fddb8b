#!/bin/bash
# final-tree check on one GPU: all GPU tests, smoke(), the default bench line, a short reference-arm run
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2z_smoke.log | cut -c1-300
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), d['e2e'], d['epochs'], d['roofline'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])
d=json.loads(open('gpurun_out/r2z_bench_ref.json').read().strip().splitlines()[-1])
print(d['value'], d['steps'], d['warmup'], d['ms_per_step'], d['cpu_baseline']['sample'][:160])
PY

#!/bin/bash
# round 2, evidence call (1 GPU): all GPU tests, ncu --set full per kernel + launch list, final bench, C3 config, C5 sweep, A/B rows
TAG=${1:-r2final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log | cut -c1-200
bash scripts/round2_ncu.sh $TAG
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --n-gauss 3000000 --width 1440 --height 1080 --views 300 --gt-sets 8 --epochs 2 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err; echo "c3 rc=$?"
timeout 400 python scripts/bwd_microbench.py --reps 6 > gpurun_out/${TAG}_bwd_microbench.jsonl 2> gpurun_out/${TAG}_bwd_microbench.err; echo "c5 rc=$?"
timeout 300 python scripts/exp_bench.py > gpurun_out/${TAG}_exp_bench.jsonl 2> gpurun_out/${TAG}_exp_bench.err; echo "exp rc=$?"
python - <<PY
import json
for f in ('bench','c3'):
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json' % f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), (d['e2e'] or {}).get('value'), d['epochs'], d['roofline']['frac'], {k:round(v,3) for k,v in d['stages_ms'].items()})
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/${TAG}_exp_bench.jsonl | cut -c1-300

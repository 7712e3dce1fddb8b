#!/bin/bash
mkdir -p gpurun_out
DNR_DEBUG_CAPTURE=1 timeout 300 python bench.py --steps 5 --warmup 3 --views 24 --gt-sets 2 --skip-cpu-baseline --skip-e2e --epochs 0 > gpurun_out/r2c8_dbg.json 2> gpurun_out/r2c8_dbg.err; echo "dbg rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c8_dbg.json').read().strip().splitlines()[-1]); print('graph:', d['cuda_graph'], round(d['ms_per_step'],3))
except Exception as e: print('ERR', e)
PY
grep -i "invalidated\|Error" gpurun_out/r2c8_dbg.err | head -5
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c8_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2c8_tests.log | cut -c1-220
B="timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-e2e --epochs 0"
$B > gpurun_out/r2c8_default.json 2> gpurun_out/r2c8_default.err; echo "default rc=$?"
$B --variant 4 > gpurun_out/r2c8_fwdquad.json 2> gpurun_out/r2c8_fwdquad.err; echo "fwdquad rc=$?"
python - <<'PY'
import json
for f in ('default','fwdquad'):
    try:
        d=json.loads(open(f'gpurun_out/r2c8_{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['cuda_graph'])
    except Exception as e: print(f,'ERR',e)
PY

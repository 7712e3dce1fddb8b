#!/bin/bash
# round 2, call 3: new raster kernels (supertile lists + tile filter, packed f32x2, transpose reduction, fused loss bwd),
# touched-only project_bwd, SSIM + Adam in the bench step.  Tests first, then the bench and A/B rows.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2c3_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2c3_tests.log
B="timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-e2e --epochs 0"
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2c3_bench.json; tail -3 gpurun_out/r2c3_bench.err
for v in 1 2 3; do $B --variant $v > gpurun_out/r2c3_variant$v.json 2> gpurun_out/r2c3_variant$v.err; echo "variant $v rc=$?"; done
for s in 0 1 3; do $B --list-shift $s > gpurun_out/r2c3_shift$s.json 2> gpurun_out/r2c3_shift$s.err; echo "shift $s rc=$?"; done
$B --no-fused-loss-bwd > gpurun_out/r2c3_unfusedloss.json 2> gpurun_out/r2c3_unfusedloss.err; echo "unfused rc=$?"
$B --no-ssim --no-optimizer > gpurun_out/r2c3_oldstep.json 2> gpurun_out/r2c3_oldstep.err; echo "oldstep rc=$?"
$B --no-graph > gpurun_out/r2c3_eager.json 2> gpurun_out/r2c3_eager.err; echo "eager rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2c3_')[1], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()}, d['roofline']['n_isects'], d['roofline']['n_isects_composited'], d['roofline'].get('list_entries_walked_fwd'), d['roofline'].get('list_entries_walked_bwd'))
    except Exception as e:
        print(f, 'ERR', e)
PY

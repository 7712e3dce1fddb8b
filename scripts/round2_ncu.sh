#!/bin/bash
# Round 2: one `ncu --set full` capture for every hand-written kernel of the step that has none yet (round 1 captured
# raster_fwd, raster_bwd, project_bwd).  About 1.5 GPU-minutes per kernel (bench setup dominates).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_ncu.sh'
# then here:  for f in gpurun_out/prof_*_r2.ncu-rep; do python scripts/ncu_summary.py $f > profiles/$(basename ${f%.ncu-rep} | sed s/prof_//).txt; done
mkdir -p gpurun_out
# -s skips the launches of the setup pass (bench.py pre-visits all 200 views) so the captured launch is a warm step
for k in project_fwd count emit pad offsets finalize_fwd loss_fwd loss_bwd; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 205 -c 1 \
    -o gpurun_out/prof_${k}_r2 python bench.py --steps 2 --warmup 3 --gt-sets 2 --skip-cpu-baseline --skip-e2e --no-graph > /dev/null 2>&1
  echo "$k rc=$?"
done
# launch list of one eager step (per-kernel times; cold-cache, serialised: use the SHARES)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 2 --warmup 3 --gt-sets 2 --skip-cpu-baseline --skip-e2e --no-graph > gpurun_out/launches_r2.log 2>&1
ls -la gpurun_out | grep r2

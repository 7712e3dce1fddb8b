#!/bin/bash
# One `ncu --set full` capture per hand-written kernel of the training step (BASELINE north_star: "each kernel ships with an
# ncu capture"), plus the launch list of one steady-state step.  ~1 GPU-minute per kernel (bench setup dominates).
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/round2_ncu.sh r2final'
# then here:  python scripts/collect_profiles.py r2final
TAG=${1:-r2final}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --views 24 --gt-sets 2 --skip-cpu-baseline --skip-e2e --no-graph --epochs 0"
# kernel regex : launches to skip (forward-only kernels also run in the 26 setup views)
for spec in raster_bwd_kernel:4 raster_fwd_kernel:30 project_fwd_kernel:30 count_kernel:30 emit_kernel:30 \
            finalize_fwd_kernel:30 loss_fwd_kernel:4 ssim_fwd_kernel:4 ssim_bwd_kernel:4 project_bwd_touched_kernel:4 \
            adam_kernel:4; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${k} -s ${skip} -c 1 -f -o gpurun_out/prof_${k%_kernel}_${TAG} $B > gpurun_out/ncu_${k}.log 2>&1
  echo "$k rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/launches_${TAG}.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | grep ${TAG}

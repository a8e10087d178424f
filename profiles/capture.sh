#!/bin/bash
# Run on the GPU box (under gpurun):  bash profiles/capture.sh <tag> [main|other]
#   main : launch list of a short bench run + one full capture of each hot kernel of the headline config
#   other: the kernels of the other configs (n=37 quaternion layout, C5 arm + ball)
# (two calls: gpurun brings back at most 64 MiB per call)
TAG=${1:-r1}
PART=${2:-main}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
FULL="ncu --set full --clock-control none --import-source on"
if [ "$PART" = main ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
  for K in backward_sym quad_fused rollout_quad8; do
    $FULL -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
  done
else
  $FULL -k regex:rollout_arm8 -s 0 -c 1 -f -o gpurun_out/prof_rollout_arm8_$TAG python scratch/c5_bench.py 8 > gpurun_out/ncu_rollout_arm8_$TAG.log 2>&1
  PB_MODEL=quadruped_quat PB_REPS=1 $FULL -k regex:quad_quat_fused -s 0 -c 1 -f -o gpurun_out/prof_quad_quat_fused_$TAG python scratch/phase_bench.py derivs > gpurun_out/ncu_quad_quat_fused_$TAG.log 2>&1
  PB_MODEL=quadruped_quat PB_REPS=1 $FULL -k regex:backward_sym -s 2 -c 1 -f -o gpurun_out/prof_backward_sym_n37_$TAG python scratch/phase_bench.py backward > gpurun_out/ncu_backward_sym_n37_$TAG.log 2>&1
fi
ls -la gpurun_out | tail -8

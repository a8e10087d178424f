#!/bin/bash
# Run on the GPU box (under gpurun):  bash profiles/capture.sh <tag>
# 1) launch list of a short bench run, 2) one full capture of each hot kernel of the headline
# config, 3) the rollout kernels of the other configs (n=37 quaternion layout, C5 arm + ball).
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in backward_sym quad_fused rollout_quad8; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:rollout_arm8 -s 0 -c 1 -f -o gpurun_out/prof_rollout_arm8_$TAG python scratch/c5_bench.py 8 > gpurun_out/ncu_rollout_arm8_$TAG.log 2>&1
PB_MODEL=quadruped_quat PB_REPS=1 ncu --set full --clock-control none --import-source on -k regex:rollout_quad8 -s 0 -c 1 -f -o gpurun_out/prof_rollout_quat_$TAG python scratch/phase_bench.py linesearch > gpurun_out/ncu_rollout_quat_$TAG.log 2>&1
ls -la gpurun_out | tail -8

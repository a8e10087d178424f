#!/bin/bash
# Run on the GPU box (under gpurun):  bash profiles/capture.sh <tag>
# 1) launch list of a short bench run, 2) one full capture of each hot kernel.
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in backward_sym quad_fused rollout_quad8; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out | tail -8

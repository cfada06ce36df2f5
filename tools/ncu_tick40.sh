#!/bin/bash
# full ncu capture of the step + obs kernels at tick ~40 of the bench workload. $1 = tag
TAG=${1:-t40}
CMD="python bench.py --steps 48 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nmmo_ -s 84 -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log

#!/bin/bash
# full ncu captures of the step kernel at tick ~30 (many players alive) and ~150 (few). $1 = tag
TAG=${1:-s}
CMD="python bench.py --steps 160 --warmup 3 --no-cpu --steady-steps 0"
ncu --set full --clock-control none --import-source on -k regex:nmmo_step -s 30 -c 1 -o gpurun_out/prof_${TAG}_t30 $CMD > gpurun_out/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nmmo_step -s 150 -c 1 -o gpurun_out/prof_${TAG}_t150 $CMD >> gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log

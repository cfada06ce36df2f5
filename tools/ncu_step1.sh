#!/bin/bash
# one full ncu capture of the step kernel at launch index $2 (default 150). $1 = tag
TAG=${1:-s}; IDX=${2:-150}
CMD="python bench.py --steps 160 --warmup 3 --no-cpu --steady-steps 0"
ncu --set full --clock-control none --import-source on -k regex:nmmo_step -s $IDX -c 1 -o gpurun_out/prof_${TAG}_t$IDX $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_${TAG}.log

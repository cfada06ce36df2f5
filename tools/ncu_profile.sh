#!/bin/bash
# ncu evidence for bench.py (run under gpurun). $1 = tag (e.g. r1a)
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nmmo_ -s 14 -c 3 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log | cut -c1-600
tail -3 gpurun_out/ncu_full_$TAG.log

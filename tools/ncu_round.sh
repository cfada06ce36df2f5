#!/bin/bash
# Round evidence (run under gpurun): launch list of the bench command + full captures of the
# step/obs kernels at tick ~40 (incremental writer) and in the dense-writer section.  $1 = tag
TAG=${1:-r1}
CMD="python bench.py --steps 48 --warmup 3 --no-cpu --steady-steps 0"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nmmo_ -s 84 -c 2 -o gpurun_out/prof_${TAG}_tick40 $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nmmo_obs -s 56 -c 1 -o gpurun_out/prof_${TAG}_dense $CMD >> gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-400
grep -c nmmo gpurun_out/launches_$TAG.csv

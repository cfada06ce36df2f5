#!/bin/bash
# Round-2 evidence (run under gpurun): the driver's bench command, its launch list, and full captures of
# both kernels INSIDE the window the driver times (ticks 5-25, every agent alive), plus the dense-writer point.
# $1 = tag
TAG=${1:-r2a}
CMD="python bench.py --gpus 1 --steps 20 --warmup 5"
FAST="$CMD --no-cpu --steady-steps 0"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches_$TAG.csv $FAST > gpurun_out/ncu_list_$TAG.log 2>&1
# launches: reset = 2 kernels, then 2 per tick; tick 12 of the run = launches 26, 27
ncu --set full --clock-control none --import-source on -k regex:nmmo_ -s 26 -c 2 -o gpurun_out/prof_${TAG}_alive $FAST > gpurun_out/ncu_full_$TAG.log 2>&1
# dense writer: obs launches so far 1 + 5 + 20 = 26, two dense warm-up ticks, then the timed dense ticks
ncu --set full --clock-control none --import-source on -k regex:nmmo_obs -s 29 -c 1 -o gpurun_out/prof_${TAG}_dense $FAST >> gpurun_out/ncu_full_$TAG.log 2>&1
tail -c 600 gpurun_out/plain_$TAG.log
grep -c nmmo gpurun_out/launches_$TAG.csv

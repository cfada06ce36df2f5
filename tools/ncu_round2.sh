#!/bin/bash
# Round-2 evidence (run under gpurun): the driver's bench command, its launch list, and full captures of
# both kernels INSIDE the window the driver times (ticks 5-25, every agent alive), plus the dense-writer point.
# $1 = tag
TAG=${1:-r2a}
CMD="python bench.py --gpus 1 --steps 20 --warmup 5"
FAST="$CMD --no-cpu --steady-steps 0"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_$TAG.csv $FAST > gpurun_out/ncu_list_$TAG.log 2>&1
# launches: reset = 2 kernels, 48 spin-up ticks (2 each), reset, then 2 per tick; tick 12 after the second reset = launches 124, 125
ncu --set full --clock-control none --import-source on -k regex:nmmo_ -s 124 -c 2 -o gpurun_out/prof_${TAG}_alive $FAST > gpurun_out/ncu_full_$TAG.log 2>&1
# dense writer: obs launches so far 1 + 48 + 1 + 5 + 20 = 75, two dense warm-up ticks, then the timed dense ticks
ncu --set full --clock-control none --import-source on -k regex:nmmo_obs -s 78 -c 1 -o gpurun_out/prof_${TAG}_dense $FAST >> gpurun_out/ncu_full_$TAG.log 2>&1
tail -c 600 gpurun_out/plain_$TAG.log
grep -c nmmo gpurun_out/launches_$TAG.csv

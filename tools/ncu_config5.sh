#!/bin/bash
# Config-5 evidence (run under gpurun): bench line, launch list and a full capture of the big kernels at tick 12.  $1 = tag
TAG=${1:-r2b}
CMD="python bench.py --gpus 1 --workload config5 --steps 20 --warmup 5"
FAST="$CMD --no-cpu --steady-steps 0"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_c5_$TAG.log 2> gpurun_out/plain_c5_$TAG.err || { tail -5 gpurun_out/plain_c5_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_c5_$TAG.csv $FAST > gpurun_out/ncu_list_c5_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nmmo_ -s 124 -c 2 -o gpurun_out/prof_${TAG}_c5 $FAST > gpurun_out/ncu_full_c5_$TAG.log 2>&1
tail -c 1500 gpurun_out/plain_c5_$TAG.log

#!/bin/bash
# one development iteration on the observation kernel (run under gpurun): parity subset, the driver-style bench line,
# then one full ncu capture of the observation kernel at tick 12.  $1 = tag
TAG=${1:-it}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_gpu_api.py tests/test_big_family_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu --steady-steps 0 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/tick', d['ms_per_step'], d['kernels_ms'], 'obs GB/s', d['roofline_obs_kernel']['achieved'], 'e2e', d['e2e']['value'])"
ncu --set full --clock-control none --import-source on -k regex:nmmo_obs -s 13 -c 1 -o gpurun_out/prof_${TAG}_obs python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --steady-steps 0 > gpurun_out/ncu_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_${TAG}.log | cut -c1-200

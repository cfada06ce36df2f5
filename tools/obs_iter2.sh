#!/bin/bash
# development iteration (run under gpurun): parity subset, driver-style bench lines of config 2 and config 5
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_gpu_api.py tests/test_big_family_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --steady-steps 0 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 ms/tick', d['ms_per_step'], d['kernels_ms'], 'obs GB/s', d['roofline_obs_kernel']['achieved'], 'e2e', d['e2e']['value'])"
python bench.py --no-cpu --steady-steps 0 --steps 20 --warmup 5 --workload config5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5 ms/tick', d['ms_per_step'], d['kernels_ms'], 'obs GB/s', d['roofline_obs_kernel']['achieved'], 'e2e', d['e2e']['value'])"

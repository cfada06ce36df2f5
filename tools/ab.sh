#!/bin/bash
# parity subset + default bench + phase totals: the A/B run after a kernel change
timeout 500 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --steady-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/tick', d['ms_per_step'], 'step GB/s', d['roofline_step_kernel']['achieved'], 'obs GB/s', d['roofline_obs_kernel']['achieved'], 'early ms', d['early_window']['ms_per_step'], 'e2e', d['e2e']['value'])"
python tools/profile_phases.py 4096 64 100 2>&1 | grep -E "total cyc"
python tools/profile_phases.py 4096 24 4 2>&1 | grep -E "total cyc"

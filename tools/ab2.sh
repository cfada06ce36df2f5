#!/bin/bash
# quick A/B: default bench + per-phase cycles in both regimes (no parity run)
python bench.py --no-cpu --steady-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/tick', d['ms_per_step'], 'step GB/s', d['roofline_step_kernel']['achieved'], 'obs GB/s', d['roofline_obs_kernel']['achieved'], 'early ms', d['early_window']['ms_per_step'], 'e2e', d['e2e']['value'])"
python tools/profile_phases.py 4096 24 4 2>&1 | sed -n "/^load/,/^total/p" | awk '{printf "%s=%s ", $1, $(NF-1)} END{print ""}'
python tools/profile_phases.py 4096 64 100 2>&1 | sed -n "/^load/,/^total/p" | awk '{printf "%s=%s ", $1, $(NF-1)} END{print ""}'

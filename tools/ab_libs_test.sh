#!/bin/bash
# like ab_libs.sh, but each build first has to pass a short parity subset
B=nmmo_b200/_build
cp $B/libnmmo_b200.so $B/libnmmo_keep.so
for so in "$@"; do
  cp $B/$so $B/libnmmo_b200.so
  timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "default_config_full_size or small_world_rich or shape_sweep or autosample" 2>&1 | tail -1
  for rep in 1 2; do
  python bench.py --no-cpu --steady-steps 0 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', 'ms/tick', round(d['ms_per_step'],4), d['kernels_ms']['step_kernel'], d['kernels_ms']['obs_kernel'])"
  done
done
cp $B/libnmmo_keep.so $B/libnmmo_b200.so

#!/bin/bash
# A/B of two builds of the library (run under gpurun): $@ = .so files under nmmo_b200/_build; each is copied over the
# library the package loads and the driver-style bench line is taken twice
B=nmmo_b200/_build
cp $B/libnmmo_b200.so $B/libnmmo_keep.so
for so in "$@"; do
  cp $B/$so $B/libnmmo_b200.so
  for rep in 1 2; do
  python bench.py --no-cpu --steady-steps 0 --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', 'ms/tick', round(d['ms_per_step'],4), d['kernels_ms']['step_kernel'], d['kernels_ms']['obs_kernel'])"
  done
done
cp $B/libnmmo_keep.so $B/libnmmo_b200.so

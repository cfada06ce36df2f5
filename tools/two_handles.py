"""Experiment: two half-size handles on two streams (step kernel of one overlapping the observation kernel of the other)."""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
from bench import world
from nmmo_b200.lib import Simulator
E = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
w = world("takeru")
sims = [Simulator(*w[:2], E, *w[2:], env_base=k * E) for k in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
for k, s in enumerate(sims):
    s.set_autosample(1, s.actions)
    s.reset(np.arange(E, dtype=np.uint64) + 1 + k * E)
torch.cuda.synchronize()
def run(n, offset):
    for t in range(n):
        for k, s in enumerate(sims):
            with torch.cuda.stream(streams[k]):
                s.step()
for mode in ("two streams", "one stream"):
    for k, s in enumerate(sims):
        s.reset(np.arange(E, dtype=np.uint64) + 1 + k * E)
    torch.cuda.synchronize()
    if mode == "one stream":
        streams = [torch.cuda.current_stream()] * 2
    run(5, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(20, 0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(mode, "ms per tick of 2 x %d envs: %.4f -> %.1f M slot-steps/s" % (E, dt / 20 * 1e3, 2 * E * 128 * 20 / dt / 1e6), sims[0].kernel_names())

"""A/B of the host-buffer step with int32 and int16 actions on one box (development aid)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bench import world
from nmmo_b200.lib import Simulator
E = 4096
w = world()
sim = Simulator(*w[:2], E, *w[2:])
P = sim.P
sim.set_autosample(1, sim.actions)
seeds = np.arange(E, dtype=np.uint64) + 1
K = 16
t32 = torch.empty((K, E, P, 12), dtype=torch.int32).pin_memory(); t16 = torch.empty((K, E, P, 12), dtype=torch.int16).pin_memory()
sim.reset(seeds)
for _ in range(100): sim.step()
for k in range(K):
    t32[k].copy_(sim.actions); t16[k].copy_(sim.actions.to(torch.int16)); sim.step()
torch.cuda.synchronize()
# raw copies
d = torch.empty((E, P, 12), dtype=torch.int32, device="cuda")
for name, src in (("h2d 25MB", t32[0]),):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): d.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(name, f"{dt*1e3:.3f} ms  {src.numel()*4/dt/1e9:.1f} GB/s")
scratch = torch.zeros_like(sim.actions)
for name, tape in (("int32", t32.numpy()), ("int16", t16.numpy()), ("int32", t32.numpy()), ("int16", t16.numpy())):
    sim.set_autosample(1, sim.actions); sim.reset(seeds)
    for _ in range(100): sim.step()
    sim.set_autosample(1, scratch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(K): sim.step_host(tape[k])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(name, f"{dt*1e3:.3f} ms/step  {E*P/dt/1e6:.0f} M slot-steps/s")
sim.close()

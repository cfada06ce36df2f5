"""Per-phase SM-clock shares of the step kernel at the bench workload (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from bench import world
from nmmo_b200.lib import Simulator

NAMES = ["load", "bookkeep+validate", "npc_decide", "update", "harvest(seq)", "use", "buy/give(seq)", "destroy",
         "attack(seq)", "move(seq)", "sell", "cull", "npc_spawn(seq)", "respawn+expire", "writeback issue", "fold events",
         "rewards/info", "tail"]
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 64
w = world()
sim = Simulator(*w[:2], E, *w[2:])
sim.reset(np.arange(E) + 1)
for _ in range(8):
    sim.sample_actions(1); sim.step()
sim.profile(True)
for _ in range(ticks):
    sim.sample_actions(1); sim.step()
torch.cuda.synchronize()
out = sim.profile(False).astype(np.float64)
tot = out.sum()
for n, v in zip(NAMES, out):
    print(f"{n:22s} {v / tot * 100:6.2f}%  {v / (E * ticks):10.0f} cycles/env-tick")
print("total cycles/env-tick", tot / (E * ticks))
dep = []
for e in (0, 1, 2, 3):
    ent, items, mp, sc = sim.snapshot(e)
    dep.append(int(np.isin(mp, (3, 6, 8, 10, 12, 14)).sum()))
    print("env", e, "tick", sc[0], "depleted tiles", dep[-1], "alive players", int((ent[:128, 43] == 1).sum()), "npcs", int((ent[128:, 43] == 1).sum()),
          "items", int((items[:, 0] > 0).sum()))

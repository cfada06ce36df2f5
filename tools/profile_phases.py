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
         "window preds", "rewards/info", "tail"]
CNT = {22: "attacks", 23: "attack rounds", 24: "movers", 25: "move rounds", 26: "events", 27: "alive players", 28: "died this tick"}
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 64
w = world()
sim = Simulator(*w[:2], E, *w[2:])
sim.reset(np.arange(E) + 1)
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 8
for _ in range(warm):
    sim.sample_actions(1); sim.step()
sim.profile(True)
for _ in range(ticks):
    sim.sample_actions(1); sim.step()
torch.cuda.synchronize()
raw = sim.profile(False).astype(np.float64)
out = raw[:len(NAMES)]
tot = out.sum()
for k, n in CNT.items():
    print(f"{n:22s} {raw[k] / (E * ticks):10.2f} per env-tick")
for n, v in zip(NAMES, out):
    print(f"{n:22s} {v / tot * 100:6.2f}%  {v / (E * ticks):10.0f} cycles/env-tick")
print("total cycles/env-tick", tot / (E * ticks))
print("obs kernel (warp 0 timeline): load %.0f, lists %.0f, market %.0f, worklist %.0f, records(warp0) %.0f; work items %.2f; mean warp busy in record loop %.0f cycles" % (
    *(raw[32:37] / (E * ticks)), raw[37] / (E * ticks), raw[40] / (E * ticks * 8)))
print("attack round parts (cycles/env-tick): clear %.0f, register %.0f, apply %.0f" % tuple(raw[44:47] / (E * ticks)))
dep = []
for e in (0, 1, 2, 3):
    ent, items, mp, sc = sim.snapshot(e)
    dep.append(int(np.isin(mp, (3, 6, 8, 10, 12, 14)).sum()))
    print("env", e, "tick", sc[0], "depleted tiles", dep[-1], "alive players", int((ent[:128, 43] == 1).sum()), "npcs", int((ent[128:, 43] == 1).sum()),
          "items", int((items[:, 0] > 0).sum()))

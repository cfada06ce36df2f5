"""Long lock-step run at full scale: error flags, event-ring overflows and counter sanity (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bench import world
from nmmo_b200.config import SPEC
from nmmo_b200.lib import Simulator
E, T = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 3000
for agent in ("takeru", "neurips23_start_kit", "yaofeng"):
    w = world(agent)
    sim = Simulator(*w[:2], E, *w[2:])
    sim.set_autosample(3, sim.actions)
    sim.reset(np.arange(E, dtype=np.uint64) + 5)
    for t in range(T):
        sim.step()
    torch.cuda.synchronize()
    sums, counts, counters = sim.stats()
    errs = sum(int(sim.snapshot(e)[3][7]) for e in range(0, E, 257))
    n = counts[SPEC["IN_LENGTH"]]
    print(agent, "ticks", T, "slot-steps", int(counters[0]), "alive-steps", int(counters[1]), "episodes", int(counters[2]), "ring overflows", int(counters[3]),
          "finished agents", int(n), "mean length %.1f" % (sums[SPEC["IN_LENGTH"]] / max(1, n)), "mean return %.4f" % (sums[SPEC["IN_RETURN"]] / max(1, n)),
          "error flags (sampled envs)", errs, "rewards finite", bool(torch.isfinite(sim.rewards).all()))
    assert counters[0] == E * 128 * T or counters[0] <= E * 128 * T
    sim.close()

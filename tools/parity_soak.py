"""Differential soak: many envs of several shapes / wrappers, CUDA simulator vs the CPU oracle, every tick bit for bit
(development aid; test infrastructure -- it drives the oracle)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from util import SMALL, build_world, run_parity
from nmmo_b200.lib import Simulator
from oracle.oracle import OracleEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
cases = [
    ("takeru rich", dict(agent="takeru"), dict(**SMALL, NC_HORIZON=150, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=2, NC_WEAPON_DROP_THR=1 << 29)),
    ("start kit", dict(agent="neurips23_start_kit"), dict(**SMALL, NC_HORIZON=120, NC_RES_DEPLETION=2)),
    ("yaofeng", dict(agent="yaofeng"), dict(**SMALL, NC_HORIZON=200, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=1)),
    ("full size", dict(agent="takeru"), dict(NC_HORIZON=1024)),
    ("mid size", dict(agent="takeru"), dict(NC_N_PLAYERS=40, NC_N_NPCS=120, NC_MAP_CENTER=48, NC_HORIZON=100, NC_RES_DEPLETION=2, NC_NPC_SPAWN_ATTEMPTS=32)),
]
for name, kw, over in cases:
    world = build_world(task_dim=None if "full" in name else 64, **kw, **over)
    n = E if "size" not in name else max(8, E // (8 if "mid" in name else 32))
    sim = Simulator(*world[:2], n, *world[2:])
    oracles = [OracleEnv(*world) for _ in range(n)]
    t0 = time.time()
    stats = run_parity(sim, oracles, seeds=np.arange(n) * 104729 + 31, ticks=T, check_state_every=50)
    print(name, "envs", n, "ticks", T, stats, "%.1f s" % (time.time() - t0), flush=True)
    sim.close()
print("parity soak ok")

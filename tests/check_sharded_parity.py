#!/usr/bin/env python
"""BASELINE.json config 3 as a check: envs sharded by global index over the GPUs of one box, full-size config,
and on every rank a sampled subset of its envs compared with the CPU oracle bit for bit (observations, rewards,
done flags, masks) every tick, with recorded draws injected into both sides.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/check_sharded_parity.py \\
        --envs 4096 --check 8 --ticks 256

(test infrastructure: uses the oracle; not part of the product path)
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from bench import world  # noqa: E402
from nmmo_b200.config import SPEC as S  # noqa: E402
from nmmo_b200.dist import global_seeds, reduce_stats  # noqa: E402
from nmmo_b200.lib import Simulator  # noqa: E402
from oracle.oracle import OracleEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--check", type=int, default=8, help="envs per GPU compared with the oracle")
    ap.add_argument("--ticks", type=int, default=256)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--policy", default="forage", choices=["forage", "random"],
                    help="forage: the scripted survival policy (episodes run to the horizon); random: uniform-random valid actions")
    ap.add_argument("--inject-ticks", type=int, default=1 << 20, help="ticks with injected draws (default: all)")
    ap.add_argument("--out", default=None, help="also write the JSON line to this file (rank 0)")
    a = ap.parse_args()
    ws, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    w = world("takeru")
    cfg = w[0]
    E, P = a.envs, int(cfg[S["NC_N_PLAYERS"]])
    base = rank * E
    sim = Simulator(*w[:2], E, *w[2:], device=lr, env_base=base)
    seeds = global_seeds(a.seed, base, E)
    pick = np.linspace(0, E - 1, a.check).astype(int)
    oracles = {int(e): OracleEnv(*w) for e in pick}
    # recorded draws: NPC wander directions and tile respawns of the first ticks, the same values on both sides
    rng = np.random.default_rng(1000 + rank)
    R = P + int(cfg[S["NC_N_NPCS"]]); Sz = int(cfg[S["NC_MAP_SIZE"]])
    for e, o in oracles.items():
        keys, vals = [], []
        for t in range(0, min(a.ticks, a.inject_ticks)):
            for row in rng.choice(np.arange(P, R), 8, replace=False):
                keys.append((t << 36) | (S["RS_NPC_DECIDE"] << 32) | (int(row) << 8)); vals.append(int(rng.integers(0, 2 ** 32)))
            for i in rng.choice(Sz * Sz, 32, replace=False):
                keys.append((t << 36) | (S["RS_RESPAWN"] << 32) | (int(i) << 8)); vals.append(int(rng.integers(0, 2 ** 28)))
        keys = np.array(keys, np.uint64); vals = np.array(vals, np.uint32)
        o.inject_rng(keys, vals); sim.inject_rng(e, keys, vals)
    sim.reset(seeds)
    for e, o in oracles.items():
        o.reset(int(seeds[e]))
    bad, t0, episodes, alive_sum = 0, time.perf_counter(), 0, 0
    pick_d = torch.as_tensor(pick, device="cuda")
    for t in range(a.ticks + 1):
        # the sampled envs' outputs come home in one gather per tensor
        g_obs = sim.obs.view(E, P, -1).index_select(0, pick_d).cpu().numpy()
        g_rew = sim.rewards.view(E, P).index_select(0, pick_d).cpu().numpy(); g_term = sim.terminated.view(E, P).index_select(0, pick_d).cpu().numpy()
        g_trunc = sim.truncated.view(E, P).index_select(0, pick_d).cpu().numpy(); g_mask = sim.mask.view(E, P).index_select(0, pick_d).cpu().numpy()
        g_done = sim.episode_done.index_select(0, pick_d).cpu().numpy()
        for k, (e, o) in enumerate(oracles.items()):
            same = (np.array_equal(g_obs[k], o.obs) and np.array_equal(g_rew[k].view(np.uint32), o.rewards.view(np.uint32))
                    and np.array_equal(g_term[k], o.terminated) and np.array_equal(g_trunc[k], o.truncated)
                    and np.array_equal(g_mask[k], o.mask) and bool(g_done[k]) == bool(o.episode_done))
            bad += 0 if same else 1
            episodes += int(o.episode_done)
            alive_sum += int(o.mask.sum())
        if t == a.ticks:
            break
        if a.policy == "forage":
            sim.forage_actions(a.seed)
        else:
            sim.sample_actions(a.seed)
        acts = sim.actions.index_select(0, pick_d).cpu().numpy()
        for k, (e, o) in enumerate(oracles.items()):
            if a.policy == "random":
                bad += 0 if np.array_equal(acts[k], o.sample_actions(a.seed + base + e)) else 1
            o.step(acts[k])
        sim.step()
    sums, counts, counters = reduce_stats(*sim.stats(), device=torch.device("cuda", lr))
    tb = torch.tensor([bad, episodes, alive_sum], dtype=torch.int64, device="cuda")
    if ws > 1:
        dist.all_reduce(tb)
    if rank == 0:
        line = json.dumps({"what": "config 3 check: env-sharded run, sampled envs of every rank vs the CPU oracle every tick (observations, "
                                   "rewards, flags, masks, episode_done), recorded draws injected into both sides", "n_gpus": ws,
                           "envs_total": ws * E, "envs_checked": ws * a.check, "ticks": a.ticks, "policy": a.policy,
                           "injected_ticks": min(a.ticks, a.inject_ticks), "mismatches": int(tb[0].item()),
                           "episodes_finished_in_checked_envs": int(tb[1].item()),
                           "mean_alive_fraction_checked": float(tb[2].item()) / max(1, ws * a.check * P * (a.ticks + 1)),
                           "slot_steps_all_ranks": float(counters[0]), "finished_agents_all_ranks": float(counts[S["IN_LENGTH"]]),
                           "seconds": time.perf_counter() - t0})
        print(line)
        if a.out:
            Path(a.out).write_text(line + "\n")
    sim.close()
    if ws > 1:
        dist.destroy_process_group()
    sys.exit(0 if int(tb[0].item()) == 0 else 1)


if __name__ == "__main__":
    main()

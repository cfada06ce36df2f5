"""CPU, world_size 2 over gloo: env sharding by global index + the episode-stat all-reduce."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
TOTAL_ENVS, TICKS, SEED = 4, 70, 21


def _episode_stats(env_ids):
    """Run the oracle for the given global env indices, return (sums, counts, counters)."""
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    from nmmo_b200.config import SPEC
    from nmmo_b200.dist import global_seeds
    from oracle.oracle import OracleEnv
    from util import SMALL, build_world
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=60)
    IN = SPEC["IN_N"]
    sums, counts, counters = np.zeros(IN), np.zeros(IN), np.zeros(8, np.uint64)
    for g in env_ids:
        o = OracleEnv(*world)
        o.reset(int(global_seeds(SEED, g, 1)[0]))
        for _ in range(TICKS):
            o.step(o.sample_actions(SEED + g))
            counters[0] += o.P; counters[1] += int(o.mask.sum())
            v = o.info_valid.astype(bool)
            if v.any():
                rows = o.info[v].astype(np.float64)
                ok = ~np.isnan(rows)
                sums += np.where(ok, rows, 0).sum(0); counts += ok.sum(0)
    return sums, counts, counters


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    sys.path.insert(0, str(ROOT))
    from nmmo_b200.dist import reduce_stats, shard
    base, n = shard(TOTAL_ENVS, world_size, rank)
    s, c, k = _episode_stats(range(base, base + n))
    s, c, k = reduce_stats(s, c, k)
    if rank == 0:
        np.savez(out, sums=s, counts=c, counters=k)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_every_env_once():
    sys.path.insert(0, str(ROOT))
    from nmmo_b200.dist import shard
    for total, ws in ((32768, 8), (10, 4), (4096, 1), (7, 3)):
        spans = [shard(total, ws, r) for r in range(ws)]
        ids = [i for b, n in spans for i in range(b, b + n)]
        assert ids == list(range(total))


def test_two_rank_stats_equal_single_process(tmp_path):
    out = tmp_path / "reduced.npz"
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    s, c, k = _episode_stats(range(TOTAL_ENVS))
    assert np.array_equal(got["counts"], c) and np.array_equal(got["counters"], k)
    assert np.allclose(got["sums"], s, rtol=0, atol=0)
    assert c[0] > 0

"""CPU: properties of the oracle (the checker must itself be sane before it checks anything)."""
import numpy as np
import pytest

from nmmo_b200.config import SPEC, ObsLayout
from util import SMALL, build_world

S = SPEC


def _run(world, seed, ticks, action_seed=5):
    from oracle.oracle import OracleEnv
    o = OracleEnv(*world)
    o.reset(seed)
    trace = []
    for _ in range(ticks):
        o.step(o.sample_actions(action_seed))
        trace.append((o.obs.copy(), o.rewards.copy(), o.terminated.copy(), o.truncated.copy(), o.mask.copy()))
    return o, trace


def test_deterministic_and_seed_sensitive():
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=80)
    _, a = _run(world, 3, 60)
    _, b = _run(world, 3, 60)
    _, c = _run(world, 4, 60)
    assert all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a, b))
    assert any(not np.array_equal(x[0], y[0]) for x, y in zip(a, c))


def test_state_invariants_hold_every_tick():
    from oracle.oracle import OracleEnv
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=300, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=2, NC_WEAPON_DROP_THR=1 << 30)
    cfg = world[0]
    L = ObsLayout(cfg)
    P, N, ninv = int(cfg[S["NC_N_PLAYERS"]]), int(cfg[S["NC_N_NPCS"]]), int(cfg[S["NC_N_INV"]])
    o = OracleEnv(*world)
    o.reset(9)
    seen_items = seen_listing = seen_equipped = seen_kill = 0
    for t in range(280):
        o.step(o.sample_actions(2))
        if o.episode_done:
            break
        ent, items, mp = o.snapshot()
        alive = ent[:, S["EA_STATUS"]] == 1
        # one entity per tile (ALLOW_MOVE_INTO_OCCUPIED_TILE = False)
        pos = ent[alive][:, [S["EA_ROW"], S["EA_COL"]]]
        assert len(np.unique(pos, axis=0)) == len(pos)
        # nobody stands on an impassible tile; resources stay in range
        mats = mp[pos[:, 0], pos[:, 1]]
        assert not np.isin(mats, (0, 1, 5, 14, 15)).any()
        for col in ("EA_HEALTH", "EA_FOOD", "EA_WATER"):
            v = ent[alive][:, S[col]]
            assert (v >= 0).all() and (v <= 100).all()
        assert (ent[alive][:, S["EA_HEALTH"]] > 0).all()            # the dead are culled in the same tick
        # items: owned by living players, inventories within capacity, equipment slots consistent
        used = items[:, S["IS_TYPE"]] > 0
        owners = items[used, S["IS_OWNER"]]
        assert ((owners >= 1) & (owners <= P)).all() and alive[owners - 1].all()
        assert (np.bincount(owners, minlength=P + 1) <= ninv).all()
        assert not (items[used, S["IS_EQUIPPED"]] & (items[used, S["IS_PRICE"]] > 0)).any()   # listed items are not equipped
        for p in np.flatnonzero(alive[:P]):
            for slot in ("EA_EQ_HAT", "EA_EQ_TOP", "EA_EQ_BOTTOM", "EA_EQ_HELD", "EA_EQ_AMMO"):
                r = ent[p, S[slot]]
                if r:
                    assert items[r - 1, S["IS_OWNER"]] == p + 1 and items[r - 1, S["IS_EQUIPPED"]] == 1
        seen_items += int(used.sum()); seen_listing += int((items[:, S["IS_PRICE"]] > 0).sum())
        seen_equipped += int(items[:, S["IS_EQUIPPED"]].sum())
        # observation contract: self row present, masks are 0/1, Move mask matches the map
        obs = o.obs
        for p in np.flatnonzero(alive[:P]):
            rec = obs[p]
            assert set(np.unique(rec[:L.m_end])) <= {0, 1}
            e = rec[L.o_entity:L.o_entity + L.n_ent * 62].view(np.int16).reshape(L.n_ent, 31)
            assert (e[:, 0] == p + 1).sum() == 1
            off, n = L.masks["Move.Direction"]
            r, c = ent[p, S["EA_ROW"]], ent[p, S["EA_COL"]]
            want = [not np.isin(mp[r + dr, c + dc], (0, 1, 5, 14, 15)) for dr, dc in ((-1, 0), (1, 0), (0, 1), (0, -1), (0, 0))]
            assert rec[off:off + n].tolist() == [int(x) for x in want]
            off, n = L.masks["Attack.Target"]
            tm = rec[off:off + n]
            assert tm[-1] == (0 if tm[:-1].any() else 1)             # no-op only when no valid target
        seen_kill += int((o.info[o.info_valid.astype(bool)][:, S["IN_NPC_KILLS"]] > 0).sum()) if o.info_valid.any() else 0
    assert seen_items > 0 and seen_listing > 0 and seen_equipped > 0


def test_injected_draws_override_the_counter_hash():
    from oracle.oracle import OracleEnv
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=200, NC_RES_DEPLETION=1)
    cfg = world[0]
    Sz = int(cfg[S["NC_MAP_SIZE"]])
    base = OracleEnv(*world); inj = OracleEnv(*world)
    # force every respawn draw of ticks 1..60 to 0 (< any threshold): depleted tiles come back at once
    keys = [(np.uint64(t) << np.uint64(36)) | (np.uint64(S["RS_RESPAWN"]) << np.uint64(32)) | (np.uint64(i) << np.uint64(8))
            for t in range(1, 61) for i in range(Sz * Sz)]
    inj.inject_rng(np.array(keys, np.uint64), np.zeros(len(keys), np.uint32))
    base.reset(2); inj.reset(2)
    dep = (3, 6, 8, 10, 12, 14)
    n_base = n_inj = 0
    for t in range(60):
        a = base.sample_actions(8)
        base.step(a)
        inj.step(inj.sample_actions(8))
        n_base += int(np.isin(base.snapshot()[2], dep).sum())
        n_inj += int(np.isin(inj.snapshot()[2], dep).sum())
    assert n_base > 0 and n_inj == 0


def test_horizon_truncates_and_auto_reset():
    from oracle.oracle import OracleEnv
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=12, NC_RES_DEPLETION=0)
    o = OracleEnv(*world)
    o.reset(1)
    for t in range(12):
        o.step(o.sample_actions(3))
    assert o.tick == 12 and o.episode_done
    live = o.mask.astype(bool) & ~o.terminated.astype(bool)
    assert live.any() and o.truncated[live].all() and o.info_valid[live].all()
    assert (o.info[live][:, S["IN_LENGTH"]] == 12).all()
    o.step(o.sample_actions(3))                       # the step after `done` is the reset of the next episode
    assert o.tick == 0 and not o.episode_done and o.mask.all() and not o.rewards.any()

"""GPU: the boundary (C ABI calls, VecEnv contract, per-env view, sharding, injection) and
size-independent properties at the full BASELINE.json size."""
import numpy as np
import pytest

from nmmo_b200.config import SPEC, ObsLayout
from util import SMALL, build_world, run_parity

pytestmark = pytest.mark.gpu
S = SPEC


def _sim(world, E, **kw):
    from nmmo_b200.lib import Simulator
    cfg, fcfg, maps, tab, emb = world
    return Simulator(cfg, fcfg, E, maps, tab, emb, **kw)


def _oracles(world, E):
    from oracle.oracle import OracleEnv
    return [OracleEnv(*world) for _ in range(E)]


def test_injected_rng_parity():
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=150, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=2)
    cfg = world[0]
    sim, oracles = _sim(world, 3), _oracles(world, 3)
    rng = np.random.default_rng(0)
    R = int(cfg[S["NC_N_PLAYERS"]] + cfg[S["NC_N_NPCS"]])
    Sz = int(cfg[S["NC_MAP_SIZE"]])
    for e, o in enumerate(oracles):
        keys, vals = [], []
        for t in range(0, 120):
            for row in rng.choice(np.arange(int(cfg[S["NC_N_PLAYERS"]]), R), 6, replace=False):   # NPC wander directions
                keys.append((t << 36) | (S["RS_NPC_DECIDE"] << 32) | (int(row) << 8)); vals.append(int(rng.integers(0, 2 ** 32)))
            for i in rng.choice(Sz * Sz, 40, replace=False):                                       # tile respawn draws
                keys.append((t << 36) | (S["RS_RESPAWN"] << 32) | (int(i) << 8)); vals.append(int(rng.integers(0, 2 ** 28)))
            for att in range(25):                                                                  # NPC spawn draws
                for k in range(6):
                    keys.append((t << 36) | (S["RS_NPC_SPAWN"] << 32) | (att << 8) | k); vals.append(int(rng.integers(0, 2 ** 32)))
        for i in range(1, int(cfg[S["NC_N_PLAYERS"]])):                                            # reset: spawn permutation
            keys.append((0 << 36) | (S["RS_SPAWN_PERM"] << 32) | (i << 8)); vals.append(int(rng.integers(0, 2 ** 32)))
        keys = np.array(keys, np.uint64); vals = np.array(vals, np.uint32)
        o.inject_rng(keys, vals)
        sim.inject_rng(e, keys, vals)
    run_parity(sim, oracles, seeds=np.arange(3) + 50, ticks=130)
    sim.close()


def test_explicit_maps_and_tasks():
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=90, NC_RES_DEPLETION=1)
    cfg, _, maps, tab, _ = world
    P = int(cfg[S["NC_N_PLAYERS"]])
    sim, oracles = _sim(world, 3), _oracles(world, 3)
    map_ids = np.array([3, 0, 2], np.int32)
    # give every agent of env 0 one task of each predicate family in turn
    task_ids = np.stack([(np.arange(P) * 7 + 11 * e) % len(tab) for e in range(3)]).astype(np.int32)
    run_parity(sim, oracles, seeds=np.array([5, 6, 7]), ticks=85, map_ids=map_ids, task_ids=task_ids)
    assert [o.map_id for o in oracles] == [3, 0, 2]
    sim.close()


def test_every_predicate_family_progresses():
    # all agents of an env share one task; long-lived agents so event-driven predicates move
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=260, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=2, NC_WEAPON_DROP_THR=1 << 30)
    cfg, _, _, tab, _ = world
    P = int(cfg[S["NC_N_PLAYERS"]])
    preds = sorted(set(tab[:, 0].tolist()))
    pick = [int(np.flatnonzero(tab[:, 0] == p)[0]) for p in preds]
    E = len(pick)
    sim, oracles = _sim(world, E), _oracles(world, E)
    task_ids = np.repeat(np.array(pick, np.int32)[:, None], P, axis=1)
    run_parity(sim, oracles, seeds=np.arange(E) + 300, ticks=240, task_ids=task_ids, check_state_every=0)
    sim.close()


def test_combined_tasks_product_and_weighted_sum():
    # manual_curriculum.py:119-122 `0.3 * EquipItem + 0.7 * GainExperience` and :201-202 `InventorySpaceGE * TickGE`
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=300, NC_RES_DEPLETION=0, NC_SPAWN_IMMUNITY=3, NC_WEAPON_DROP_THR=-1,
                        NC_NPC_AGGR_PCT=200, NC_NPC_NEUT_PCT=200)
    cfg, _, _, tab, _ = world
    P = int(cfg[S["NC_N_PLAYERS"]])
    pick = np.flatnonzero(tab[:, 7] != 0)
    assert (tab[pick, 7] == 2).sum() >= 3 and (tab[pick, 7] == 1).sum() >= 1
    E = len(pick)
    sim, oracles = _sim(world, E), _oracles(world, E)
    task_ids = np.repeat(pick.astype(np.int32)[:, None], P, axis=1)
    moved = []

    def watch(t, sim_, oracles_):
        moved.append(float(np.abs(sim_.rewards.cpu().numpy()).sum()))

    run_parity(sim, oracles, seeds=np.arange(E) + 500, ticks=280, task_ids=task_ids, check_state_every=0, on_tick=watch)
    assert sum(m > 0 for m in moved) > 10, "the combined tasks must make progress (rewards move)"
    sim.close()


def test_env_sharding_invariance():
    """Global env g gives the same trajectory whichever handle (env_base) hosts it."""
    import torch
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=60)
    whole = _sim(world, 4, env_base=0)
    part = _sim(world, 2, env_base=2)
    from nmmo_b200.dist import global_seeds
    whole.reset(global_seeds(9, 0, 4)); part.reset(global_seeds(9, 2, 2))
    P = whole.P
    for t in range(70):
        whole.sample_actions(9); part.sample_actions(9)
        torch.cuda.synchronize()
        assert torch.equal(whole.actions[2:], part.actions)
        assert torch.equal(whole.obs[2 * P:], part.obs) and torch.equal(whole.rewards[2 * P:], part.rewards)
        whole.step(); part.step()
    whole.close(); part.close()


def test_handles_of_different_shapes_coexist():
    # the kernels' dynamic shared-memory opt-in is process-wide: creating a small handle must not break a big one
    import torch
    big = _sim(build_world(task_dim=64, NC_N_PLAYERS=64, NC_N_NPCS=128, NC_MAP_CENTER=64, NC_HORIZON=30), 2)
    big.reset([1, 2])
    small = _sim(build_world(task_dim=64, **SMALL, NC_HORIZON=30), 2)
    small.reset([3, 4])
    for _ in range(5):
        big.sample_actions(1); big.step(); small.sample_actions(1); small.step()
    torch.cuda.synchronize()
    assert int(big.mask.sum()) > 0 and int(small.mask.sum()) > 0
    big.close(); small.close()


def test_step_host_equals_device_step():
    import torch
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=60)
    from nmmo_b200.lib import pack_actions_u8
    a, b, c, d = _sim(world, 2), _sim(world, 2), _sim(world, 2), _sim(world, 2)
    a.reset([1, 2]); b.reset([1, 2]); c.reset([1, 2]); d.reset([1, 2])
    for t in range(40):
        a.sample_actions(4)
        torch.cuda.synchronize()
        acts = a.actions.cpu().numpy()
        a.step()
        packed = pack_actions_u8(acts)
        assert packed.dtype == np.uint8 and np.array_equal(packed, pack_actions_u8(a.actions).cpu().numpy())
        for other, host_acts in ((b, acts), (c, acts.astype(np.int16)), (d, packed)):      # int32, int16 and byte-packed calls
            rew, term, trunc, mask, obs = other.step_host(host_acts, want_obs=True)
            torch.cuda.synchronize()
            assert np.array_equal(rew, a.rewards.cpu().numpy()) and np.array_equal(term, a.terminated.cpu().numpy())
            assert np.array_equal(trunc, a.truncated.cpu().numpy()) and np.array_equal(mask, a.mask.cpu().numpy())
            assert np.array_equal(obs, a.obs.cpu().numpy())
    # the wide head (Buy.MarketItem) survives the packing, negative entries stay no-ops
    wide = np.zeros((1, 1, 12), np.int32); wide[0, 0, 2] = 300; wide[0, 0, 0] = 2; wide[0, 0, 8] = -1
    pk = pack_actions_u8(wide)
    assert pk[0, 0, 0] == (2 | (1 << 2)) and pk[0, 0, 2] == 300 - 256 and pk[0, 0, 8] == 255
    a.close(); b.close(); c.close(); d.close()


def test_step_host_u8_in_two_ranges(monkeypatch):
    """nmmo_step_host_u8 on a large handle copies the actions as two env ranges and runs the step kernel of the first range
    underneath the copy of the second (NMMO_B200_H2D_SPLIT_MIN lowers the threshold for the test): same tensors as the
    device-resident step, across an episode boundary."""
    import torch
    from nmmo_b200.lib import pack_actions_u8
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=30)
    E = 13
    a = _sim(world, E)
    monkeypatch.setenv("NMMO_B200_H2D_SPLIT_MIN", "4")
    b = _sim(world, E)
    monkeypatch.delenv("NMMO_B200_H2D_SPLIT_MIN")
    seeds = list(range(3, 3 + E))
    a.reset(seeds); b.reset(seeds)
    for t in range(45):
        a.sample_actions(6)
        torch.cuda.synchronize()
        packed = pack_actions_u8(a.actions).cpu().numpy()
        a.step()
        rew, term, trunc, mask, obs = b.step_host(packed, want_obs=True)
        torch.cuda.synchronize()
        assert np.array_equal(obs, a.obs.cpu().numpy()), f"tick {t}"
        assert np.array_equal(rew, a.rewards.cpu().numpy()) and np.array_equal(mask, a.mask.cpu().numpy()), f"tick {t}"
        assert np.array_equal(term, a.terminated.cpu().numpy()) and np.array_equal(trunc, a.truncated.cpu().numpy()), f"tick {t}"
    a.close(); b.close()


def test_dense_and_incremental_writers_agree():
    import torch
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=70, NC_RES_DEPLETION=1)
    a, b = _sim(world, 3), _sim(world, 3)
    b.set_obs_full(True)
    a.reset([7, 8, 9]); b.reset([7, 8, 9])
    for t in range(90):                                     # crosses an episode boundary (auto-reset)
        a.sample_actions(2); b.sample_actions(2)
        torch.cuda.synchronize()
        assert torch.equal(a.obs, b.obs), f"tick {t}"
        a.step(); b.step()
    a.close(); b.close()


def test_mask_layouts_beyond_1024_entries_are_rejected():
    """The observation kernel keeps one 32-entry mask word per lane: nmmo_create refuses layouts whose twelve heads total
    more than 1024 entries with NM_ERR_LIMIT instead of writing wrong masks (the reference's layout has 946)."""
    from nmmo_b200.lib import NmmoError
    world = build_world(task_dim=64, **SMALL, NC_N_ENT_OBS=200)      # 3 * 201 + 385 + ... > 1024
    assert ObsLayout(world[0]).m_end > 1024
    with pytest.raises(NmmoError, match="1024"):
        _sim(world, 2)
    ok = build_world(task_dim=64, **SMALL, NC_N_ENT_OBS=60)
    assert ObsLayout(ok[0]).m_end <= 1024
    _sim(ok, 2).close()


def test_vecenv_contract():
    import torch
    from argparse import Namespace
    from nmmo_b200.emulation import unpack_batched_obs
    from nmmo_b200.vecenv import B200VecEnv
    env_ns = Namespace(num_agents=16, num_npcs=32, max_episode_length=40, maps_path="maps/", map_size=32, num_maps=4,
                       map_force_generation=False, death_fog_tick=None, task_size=64, spawn_immunity=20,
                       resilient_population=0, curriculum_file_path=None)
    wrap_ns = Namespace(eval_mode=False, early_stop_agent_num=0, use_custom_reward=True, explore_bonus_weight=0.01,
                        clip_unique_event=3, disable_give=True)
    pool = B200VecEnv(env_kwargs={"env": env_ns, "reward_wrapper": wrap_ns}, num_envs=3, envs_per_worker=1, envs_per_batch=3,
                      env_pool=False, mask_agents=True, agent="takeru")
    assert pool.agents_per_env == 16 and pool.single_action_space.shape == (12,)
    assert pool.single_observation_space.shape == (pool.driver_env.obs_sz,)
    pool.async_reset(1)
    returns = []
    for t in range(120):
        o, r, d, tr, infos, env_id, mask = pool.recv()
        B = 3 * 16
        assert o.is_cuda and o.shape == (B, pool.driver_env.obs_sz) and r.shape == (B,) and mask.shape == (B,)
        assert len(env_id) == B
        nested = unpack_batched_obs(o, pool.driver_env.unflatten_context)
        assert nested["Tile"].shape == (B, 225, 3) and nested["Entity"].shape == (B, 100, 31)
        for i in infos:
            assert "return" in i and "length" in i and "stats" in i and "curriculum" in i
            returns.append(i["return"])
        pool.sim.sample_actions(3)
        pool.send(pool.sim.actions.reshape(B, 12))
    assert len(returns) >= 3 * 16                           # at least one full episode per env (horizon 40)
    st = pool.stats()
    assert st["episodes"] >= 3 and "stats/achieved/unique_events" in st
    pool.close()


def test_task_swap_at_reset():
    """Syllabus hook (/root/reference/syllabus_wrapper.py:129-150): reset(new_task=k) gives every agent of the
    env table row k; other envs keep their tasks; the record's Task block is the row's embedding."""
    from nmmo_b200.env import EnvView
    from nmmo_b200.vecenv import B200VecEnv
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=50)
    cfg, fcfg, maps, tab, emb = world
    sim = _sim(world, 3)
    sim.reset(np.arange(3, dtype=np.uint64))
    before = [sim.task_state(e)[0].copy() for e in range(3)]
    view = EnvView(sim, 1)
    k = len(tab) - 1
    obs, _ = view.reset(seed=9, new_task=k)
    assert (sim.task_state(1)[0] == k).all()
    assert (sim.task_state(0)[0] == before[0]).all() and (sim.task_state(2)[0] == before[2]).all()
    for a, o in obs.items():
        assert (np.asarray(o["Task"]).view(np.uint16) == emb[k]).all()
    assert {t[0].spec_name for t in view.agent_task_map.values()} == {f"task_{k}"}
    view.reset(seed=10)                                  # without new_task the tasks are drawn again
    assert len(set(sim.task_state(1)[0].tolist())) > 1
    with pytest.raises(RuntimeError):
        view.reset(seed=1, new_task=len(tab))
    sim.close()


def test_env_view_matches_oracle():
    from nmmo_b200.env import EnvView, ACTION_KEYS
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=50)
    sim = _sim(world, 1)
    o = _oracles(world, 1)[0]
    env = EnvView(sim)
    obs, info = env.reset(seed=13)
    o.reset(13)
    assert env.agents == env.possible_agents
    L = ObsLayout(world[0])
    for t in range(45):
        a = o.sample_actions(6)
        acts = {ag: {} for ag in env.agents}
        for ag in env.agents:
            for k, (x, y) in enumerate(ACTION_KEYS):
                acts[ag].setdefault(x, {})[y] = int(a[ag - 1, k])
        obs, rew, term, trunc, infos = env.step(acts)
        o.step(a)
        assert sorted(obs) == [p + 1 for p in np.flatnonzero(o.mask)]
        for ag in obs:
            off, n = L.masks["Move.Direction"]
            assert np.array_equal(obs[ag]["ActionTargets"]["Move"]["Direction"], o.obs[ag - 1][off:off + n].view(np.int8))
            assert np.float32(rew[ag]) == o.rewards[ag - 1] and term[ag] == bool(o.terminated[ag - 1])
        # the secondary boundary of SURVEY.md 8(b): realm.tick / realm.players / dead_this_tick / agent_task_map / config
        realm = env.realm
        oe, _, _ = o.snapshot()
        assert realm.tick == o.tick and env.config.COMBAT_SPAWN_IMMUNITY == int(world[0][S["NC_SPAWN_IMMUNITY"]])
        assert sorted(realm.players) == [p + 1 for p in range(env.P) if oe[p, S["EA_STATUS"]] == 1]
        assert sorted(realm.players.dead_this_tick) == [p + 1 for p in range(env.P) if oe[p, S["EA_STATUS"]] == 2]
        for ag, pl in realm.players.items():
            assert pl.health.val == oe[ag - 1, S["EA_HEALTH"]] and pl.food.val == oe[ag - 1, S["EA_FOOD"]] and pl.gold.val == oe[ag - 1, S["EA_GOLD"]]
            assert pl.history.damage_inflicted == oe[ag - 1, S["EA_DMG_INFLICTED"]] and pl.fishing_exp.val == oe[ag - 1, S["EA_FISHING_EXP"]]
        tid, comp, sig, mp = o.task_state()
        tmap = env.agent_task_map
        for p in range(env.P):
            tk = tmap[p + 1][0]
            assert tk.spec_name == f"task_{tid[p]}" and tk.completed == bool(comp[p]) and tk.reward_signal_count == sig[p] and tk._max_progress == mp[p]
        if o.episode_done:
            assert env.agents == [] and any(i.get("episode_done") for i in infos.values())
            break
    sim.close()


def test_full_size_properties():
    """BASELINE.json configs[1] size: 4096 envs x 128 agents.  Size-independent checks: bit-exact
    replay determinism, a sampled subset of envs bit-exact against the oracle, conservation of the
    per-tick accounting, no event-ring overflow."""
    import torch
    world = build_world()
    E = 4096
    sim = _sim(world, E)
    L = ObsLayout(world[0])
    seeds = np.arange(E, dtype=np.uint64) + 1
    sample = [0, 1, 777, 2048, 4095]
    oracles = _oracles(world, len(sample))

    def rollout(check):
        sim.reset(seeds)
        if check:
            for o, e in zip(oracles, sample):
                o.reset(int(seeds[e]))
        digest = []
        for t in range(48):
            sim.sample_actions(5)
            if check:
                torch.cuda.synchronize()
                acts = sim.actions.cpu().numpy()
                for o, e in zip(oracles, sample):
                    assert np.array_equal(sim.obs[e * 128:(e + 1) * 128].cpu().numpy(), o.obs), f"tick {t} env {e}"
                    assert np.array_equal(acts[e], o.sample_actions(5 + e))
                    o.step(acts[e])
            sim.step()
            digest.append((int(sim.mask.sum()), float(sim.rewards.double().sum()), int(sim.obs[:, :L.m_end].sum())))
        return digest

    d1 = rollout(True)
    sums, counts, counters = sim.stats(clear=True)
    assert counters[3] == 0, "event ring overflowed"
    assert counters[0] == 48 * E * 128 and counters[1] == sum(d[0] for d in d1)
    d2 = rollout(False)
    assert d1 == d2
    sim.close()


def test_event_ring_overflow_is_an_error(monkeypatch):
    """An env that emits more events in a tick than its ring holds drops events; that must not pass silently: the
    host-buffer step call and nmmo_check / B200VecEnv.stats fail with NM_ERR_STATE (ring shrunk to 4 by a test hook)."""
    from nmmo_b200.lib import NmmoError
    monkeypatch.setenv("NMMO_B200_EV_CAP", "4")
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=50)
    sim = _sim(world, 2)
    sim.reset(np.arange(2, dtype=np.uint64) + 5)
    sim.check()
    acts = np.zeros((2, sim.P, 12), np.int32); acts[:, :, 8] = 1          # everybody moves: GO_FARTHEST / DRINK events
    with pytest.raises(NmmoError, match="event ring overflow"):
        for _ in range(6):
            sim.step_host(acts)
    with pytest.raises(NmmoError, match="event ring overflow"):
        sim.check()
    sim.close()


def test_forager_policy_matches_numpy_and_keeps_agents_alive():
    """nmmo_forage_actions (scripted survival policy of the bench's survival leg) against tests/forager_ref.py on the
    device's own records, tick by tick; and it does what it is for: most agents are still alive when uniform-random
    agents have long starved."""
    import torch
    from forager_ref import forage_actions
    from nmmo_b200.mapgen import generate_maps
    cfg, fcfg, maps, tab, emb = build_world(task_dim=64, NC_N_PLAYERS=32, NC_N_NPCS=32, NC_MAP_CENTER=48, NC_HORIZON=200)
    maps = generate_maps(cfg, 11, 3, terrain="fine")
    sim = _sim((cfg, fcfg, maps, tab, emb), 3)
    sim.reset(np.arange(3, dtype=np.uint64) + 2)
    P = sim.P
    for t in range(90):
        sim.forage_actions(7)
        torch.cuda.synchronize()
        a = sim.actions.cpu().numpy()
        obs = sim.obs.cpu().numpy().reshape(3, P, -1); mask = sim.mask.cpu().numpy().reshape(3, P)
        for e in range(3):
            tick = int(sim.snapshot(e)[3][0])
            ref = forage_actions(cfg, obs[e], mask[e], tick, 7, env_global=e)
            assert np.array_equal(a[e], ref), f"tick {t} env {e}: rows {np.flatnonzero((a[e] != ref).any(1))[:4]}"
        sim.step()
    torch.cuda.synchronize()
    alive = sim.mask.cpu().numpy().reshape(3, P).sum(1)
    assert (alive >= P // 3).all(), f"foragers alive at tick 90: {alive}"
    sim.close()


def test_vecenv_shards_reproduce_the_single_pool():
    """Two B200VecEnv shards (env_base 0 and 2) run the envs of one 4-env pool: seeds, maps, tasks and sampled actions are
    keyed by the GLOBAL env index, so the trajectories are identical."""
    import torch
    from argparse import Namespace
    from nmmo_b200.vecenv import B200VecEnv
    env_ns = Namespace(num_agents=16, num_npcs=32, max_episode_length=40, maps_path="maps/", map_size=32, num_maps=4,
                       map_force_generation=False, death_fog_tick=None, task_size=64, spawn_immunity=20,
                       resilient_population=0, curriculum_file_path=None)
    kw = {"env": env_ns}
    whole = B200VecEnv(env_kwargs=kw, num_envs=4, agent="takeru", collect_infos=False)
    parts = [B200VecEnv(env_kwargs=kw, num_envs=2, agent="takeru", collect_infos=False, env_base=b) for b in (0, 2)]
    for pool in [whole] + parts:
        pool.async_reset(7)
    for t in range(60):
        o, r, d, tr, _, _, m = whole.recv()
        got = [p.recv() for p in parts]
        assert torch.equal(o, torch.cat([g[0] for g in got])), f"tick {t}: observations differ between the pool and its shards"
        assert torch.equal(r, torch.cat([g[1] for g in got])) and torch.equal(m, torch.cat([g[6] for g in got]))
        assert torch.equal(d, torch.cat([g[2] for g in got])) and torch.equal(tr, torch.cat([g[3] for g in got]))
        whole.sim.sample_actions(3); whole.send(whole.sim.actions.reshape(-1, 12))
        for p in parts:
            p.sim.sample_actions(3); p.send(p.sim.actions.reshape(-1, 12))
    st_w = whole.stats(); st_p = [p.stats() for p in parts]
    assert st_w["episodes"] == sum(s["episodes"] for s in st_p) and st_w["episodes"] >= 4
    for pool in [whole] + parts:
        pool.close()

"""GPU parity: the CUDA simulator (through the C ABI) vs the CPU oracle, bit-exact every tick."""
import numpy as np
import pytest

from util import SMALL, build_world, run_parity

pytestmark = pytest.mark.gpu


def _make(world, E, env_base=0):
    from nmmo_b200.lib import Simulator
    from oracle.oracle import OracleEnv
    cfg, fcfg, maps, tab, emb = world
    sim = Simulator(cfg, fcfg, E, maps, tab, emb, env_base=env_base)
    oracles = [OracleEnv(cfg, fcfg, maps, tab, emb) for _ in range(E)]
    return sim, oracles


def test_small_world_episode():
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=96)
    sim, oracles = _make(world, 6)
    stats = run_parity(sim, oracles, seeds=np.arange(6) + 11, ticks=230)
    assert stats["episodes_done"] >= 6
    sim.close()


def test_small_world_rich_interactions():
    # slow starvation keeps agents alive long enough for items, market, combat and level-ups
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=400, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=3,
                        NC_WEAPON_DROP_THR=1 << 30)
    sim, oracles = _make(world, 4)
    stats = run_parity(sim, oracles, seeds=np.arange(4) + 5, ticks=420)
    assert stats["infos"] > 0
    sim.close()


def test_allow_occupied_start_kit():
    world = build_world(agent="neurips23_start_kit", task_dim=64, **SMALL, NC_HORIZON=200, NC_RES_DEPLETION=1,
                        NC_ALLOW_OCCUPIED=1, NC_SPAWN_IMMUNITY=2)
    sim, oracles = _make(world, 4)
    run_parity(sim, oracles, seeds=np.arange(4) + 100, ticks=210)
    sim.close()


def test_yaofeng_wrapper():
    # agent_zoo/yaofeng/reward_wrapper.py on the device: bonus terms and the dangerous-NPC target mask
    world = build_world(agent="yaofeng", task_dim=64, wrapper_over={"attack_bonus_weight": 0.005, "early_stop_agent_num": 2},
                        **SMALL, NC_HORIZON=200, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=3, NC_WEAPON_DROP_THR=1 << 30)
    sim, oracles = _make(world, 4)
    stats = run_parity(sim, oracles, seeds=np.arange(4) + 21, ticks=220)
    assert stats["infos"] > 0
    sim.close()


def test_team_tasks_and_named_targets():
    # reward_to="team" subjects and "left/right team (leader)" targets (nm_task_flag) next to the ordinary curriculum:
    # teams of 4, one task per team, group predicates evaluated over the alive members
    from nmmo_b200.tasks import EVENT, TF_RELATIVE_TARGET, TF_TEAM, default_curriculum, make_task_table, task_row
    T, R = TF_TEAM, TF_RELATIVE_TARGET
    rows = default_curriculum()[:20] + [
        task_row("ALL_MEMBERS_WITHIN_RANGE", 5, 0, 0, T, pred2="TICK_GE", q0=40, combine=1), task_row("ALL_DEAD", 1, 0, 0, T | R),
        task_row("ALL_DEAD", -1, 1, 0, T | R), task_row("STAY_ALIVE", 0, 0, 0, T), task_row("DISTANCE_TRAVELED", 20, 0, 0, T),
        task_row("COUNT_EVENT", EVENT["DRINK_WATER"], 8, 0, T), task_row("CAN_SEE_GROUP", 1, 0, 0, R), task_row("CAN_SEE_GROUP", -1, 0, 0, T | R),
        task_row("CAN_SEE_AGENT", 1, 0, 0, R), task_row("CAN_SEE_TILE", 4, 0, 0, T), task_row("HOARD_GOLD", 6, 0, 0, T),
        task_row("OCCUPY_TILE", 24, 24, 0, T), task_row("EARN_GOLD", 3, 0, 0, T), task_row("ATTAIN_SKILL", 1, 2, 0, T)]
    cfg, fcfg, maps, _, _ = build_world(task_dim=64, **SMALL, NC_HORIZON=150, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=3, NC_TEAM_SIZE=4)
    tab, emb = make_task_table(rows, 64, seed=2)
    sim, oracles = _make((cfg, fcfg, maps, tab, emb), 8)
    stats = run_parity(sim, oracles, seeds=np.arange(8) + 71, ticks=170)
    assert stats["infos"] > 0
    tid = sim.task_state(0)[0]
    assert all(len(set(tid[t * 4:(t + 1) * 4].tolist())) == 1 for t in range(4))
    sim.close()


def _dims(cfg):
    from nmmo_b200.config import ObsLayout
    return np.array(ObsLayout(cfg).action_dims)


def test_adversarial_actions():
    # actions ignore the masks: every head uniform over [-2, width + 2) -- invalid targets, items the agent does
    # not own, out-of-range and negative indices.  Validation must reject exactly what the oracle rejects.
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=120, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=2, NC_WEAPON_DROP_THR=1 << 30)
    sim, oracles = _make(world, 4)
    dims = _dims(world[0])
    rng = np.random.default_rng(5)

    def act(t, obs):
        return rng.integers(-2, dims[None, None, :] + 2, size=(4, sim.P, 12))

    stats = run_parity(sim, oracles, seeds=np.arange(4) + 31, ticks=125, action_fn=act)
    assert stats["episodes_done"] >= 1
    sim.close()


def test_collision_heavy_moves():
    # config 5 in miniature: a 16x16 arena packed with 24 players and 48 NPCs that (almost) always move --
    # long follow-the-leader chains, four-way contention for one tile, swaps
    world = build_world(task_dim=64, NC_N_PLAYERS=24, NC_N_NPCS=48, NC_MAP_CENTER=16, NC_HORIZON=90, NC_RES_DEPLETION=1,
                        NC_SPAWN_IMMUNITY=200)
    sim, oracles = _make(world, 6)
    L_move = 8
    rng = np.random.default_rng(9)

    def act(t, obs):
        a = np.zeros((6, sim.P, 12), np.int64)
        a[:, :, L_move] = rng.integers(0, 4, size=(6, sim.P))
        a[:, :, L_move][rng.random((6, sim.P)) < 0.1] = 4
        a[:, :, 1] = 100            # Attack.Target no-op (N_ENT_OBS)
        return a

    run_parity(sim, oracles, seeds=np.arange(6) + 77, ticks=95, action_fn=act, check_state_every=8)
    sim.close()


def test_item_heavy_with_tiny_staged_prefix(monkeypatch):
    # every harvest drops a weapon too, nobody starves: inventories fill up, items are used, sold, bought and
    # destroyed; NMMO_B200_ICAP=8 makes the observation kernel read all but 8 item rows from HBM
    monkeypatch.setenv("NMMO_B200_ICAP", "8")
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=260, NC_RES_DEPLETION=0, NC_SPAWN_IMMUNITY=3,
                        NC_WEAPON_DROP_THR=-1, NC_NPC_AGGR_PCT=200, NC_NPC_NEUT_PCT=200)     # passive NPCs only
    sim, oracles = _make(world, 3)
    peak = []

    def watch(t, sim_, oracles_):
        if t % 10 == 0:
            peak.append(int((oracles_[0].snapshot()[1][:, 0] > 0).sum()))

    stats = run_parity(sim, oracles, seeds=np.arange(3) + 41, ticks=270, on_tick=watch)
    assert max(peak) > 8, f"the item table must outgrow the staged prefix (peak live rows {max(peak)})"
    assert stats["infos"] > 0
    sim.close()


def test_default_config_full_size():
    world = build_world()
    sim, oracles = _make(world, 3)
    run_parity(sim, oracles, seeds=np.array([1, 2, 3]), ticks=160, check_state_every=32)
    sim.close()


@pytest.mark.parametrize("over,std", [(dict(), True), (dict(NC_RES_DEPLETION=3, NC_REACH=4), False), (dict(NC_N_ENT_OBS=60), False),
                                      (dict(NC_SPAWN_IMMUNITY=2, NC_HORIZON=100, NC_EARLY_STOP_N=4), True)])
def test_default_shape_std_and_generic_kernels(over, std):
    # the reference's default shape runs the compile-time-shape kernels (engine defaults folded); an engine override or another
    # record layout on the same shape must fall back to the generic instantiation -- and both must match the oracle
    world = build_world(**over)
    sim, oracles = _make(world, 3)
    assert ("_std" in sim.kernel_names()[0]) == std and ("_std" in sim.kernel_names()[1]) == std, sim.kernel_names()
    run_parity(sim, oracles, seeds=np.array([5, 6, 7]), ticks=110, check_state_every=36)
    sim.close()


@pytest.mark.parametrize("P,N,center,E,over", [
    (8, 16, 24, 7, dict(NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=1)),                      # tiny, odd env count
    (40, 120, 48, 5, dict(NC_RES_DEPLETION=2, NC_NPC_SPAWN_ATTEMPTS=32)),              # more NPCs than threads / 2
    (104, 296, 96, 3, dict(NC_RES_DEPLETION=3, NC_SPAWN_IMMUNITY=5)),                  # R = 400: second row per thread used
    (256, 256, 128, 2, dict(NC_RES_DEPLETION=4)),                                      # the player limit (one thread per player)
    (64, 64, 32, 4, dict(NC_RES_DEPLETION=1, NC_REACH=5, NC_VISION=7, NC_NPC_VISION=7, NC_LISTING_DURATION=1)),
])
def test_shape_sweep(P, N, center, E, over):
    # shapes away from the default: thread/row mappings, bitmap word counts, hash-table loads, odd E
    world = build_world(task_dim=64, NC_N_PLAYERS=P, NC_N_NPCS=N, NC_MAP_CENTER=center, NC_HORIZON=70, **over)
    sim, oracles = _make(world, E)
    stats = run_parity(sim, oracles, seeds=np.arange(E) + 1000 + P, ticks=80, check_state_every=20)
    assert stats["infos"] > 0
    sim.close()


def test_many_seeds_rich():
    # breadth over seeds: 48 small envs with long-lived agents (items, market, combat, level-ups, kills)
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=150, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=2, NC_WEAPON_DROP_THR=1 << 29)
    sim, oracles = _make(world, 48)
    stats = run_parity(sim, oracles, seeds=np.arange(48) * 7919 + 17, ticks=160, check_state_every=40)
    assert stats["episodes_done"] >= 48
    sim.close()


def test_respawn_scan_fallback(monkeypatch):
    # the depleted-tile list is distrusted every tick: respawn rebuilds it from the bit-sliced map scan
    monkeypatch.setenv("NMMO_B200_NO_DEPL_LIST", "1")
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=120, NC_RES_DEPLETION=1)
    sim, oracles = _make(world, 4)
    run_parity(sim, oracles, seeds=np.arange(4) + 61, ticks=130, check_state_every=10)
    sim.close()


@pytest.mark.parametrize("epc", ["1", "2"])
def test_envs_per_cta_variants(monkeypatch, epc):
    # the default is three environments per CTA (item table and event ring in place in HBM / L2); configurations whose
    # tables do not fit that way run two per CTA (everything in shared memory), or one
    monkeypatch.setenv("NMMO_B200_ENVS_PER_CTA", epc)
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=60, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=3)
    sim, oracles = _make(world, 5)
    stats = run_parity(sim, oracles, seeds=np.arange(5) + 3, ticks=130)
    assert stats["episodes_done"] >= 5
    sim.close()


def test_full_size_whole_episode():
    # config 1 of BASELINE.json as one lock-stepped env pair: default engine config (128 agents, 256 NPCs, 160x160
    # map, horizon 1024) with the start-kit wrapper, run until the episode ends (early stop at 8 agents) and on
    # through the automatic reset
    world = build_world(agent="neurips23_start_kit")
    sim, oracles = _make(world, 2)
    stats = run_parity(sim, oracles, seeds=np.array([1, 7]), ticks=400, check_state_every=50)
    assert stats["episodes_done"] >= 2 and stats["infos"] >= 2 * 120
    sim.close()


def test_autosample_matches_sampler_kernel():
    import torch
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=80, NC_RES_DEPLETION=1)
    sim, oracles = _make(world, 3)
    auto = torch.zeros_like(sim.actions)
    sim.set_autosample(77, auto)
    sim.reset(np.arange(3) + 9)
    for o, s in zip(oracles, np.arange(3) + 9):
        o.reset(int(s))
    for t in range(60):
        sim.sample_actions(77)
        torch.cuda.synchronize()
        a = sim.actions.cpu().numpy()
        assert np.array_equal(a, auto.cpu().numpy()), f"tick {t}: fused sampler differs from the sampler kernel"
        for e, o in enumerate(oracles):
            assert np.array_equal(a[e], o.sample_actions(77 + e)), f"tick {t} env {e}: oracle sampler differs"
            o.step(a[e])
        sim.step()
    sim.set_autosample(0, enable=False)
    sim.close()

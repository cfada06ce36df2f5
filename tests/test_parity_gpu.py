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


def test_default_config_full_size():
    world = build_world()
    sim, oracles = _make(world, 3)
    run_parity(sim, oracles, seeds=np.array([1, 2, 3]), ticks=160, check_state_every=32)
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

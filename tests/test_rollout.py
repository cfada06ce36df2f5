"""Rollout storage / sort / GAE (clean_pufferl.py:183-197, :329-348, :413-436).

CPU: oracle/rollout_oracle.py vs the golden vectors produced by executing the reference's own source lines
(tests/golden/make_golden_rollout.py).  GPU: the CUDA path (C ABI nmmo_rollout_*) vs the oracle, bit for bit:
the stored rows, the sorted order and the float32 advantages (compared as bit patterns, no tolerance).
"""
from pathlib import Path

import numpy as np
import pytest

from rollout_inputs import step_stream

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem for p in GOLDEN.glob("rollout_*.npz"))


def _fill_oracle(g):
    from oracle.rollout_oracle import RolloutOracle
    B, n_slots, stride = int(g["batch_size"]), int(g["n_slots"]), int(g["stride"])
    o = RolloutOracle(B, stride)
    env_id = np.arange(n_slots)
    n = 0
    for step, inp in step_stream(int(g["seed"]), n_slots, stride, float(g["p_alive"]), float(g["p_done"]), float(g["p_learner"])):
        if o.ptr >= B + 1:
            break
        o.store(inp["o"], inp["value"], inp["actions"], inp["logprob"], inp["r"], inp["d"], inp["mask"], inp["pool_mask"], env_id, step)
        n += 1
    return o, n


def test_fixtures_present():
    assert {"rollout_small", "rollout_pool", "rollout_nodone"} <= set(CASES)


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_lines(case):
    g = np.load(GOLDEN / f"{case}.npz")
    o, n = _fill_oracle(g)
    assert n == int(g["n_steps"]) and o.ptr == int(g["batch_size"]) + 1
    for name in ("obs", "actions", "logprobs", "rewards", "dones", "values"):
        assert np.array_equal(getattr(o, name), g[name]), name
    idxs, adv = o.gae(float(g["gamma"]), float(g["gae_lambda"]))
    assert np.array_equal(idxs, g["idxs"])
    assert np.array_equal(adv.view(np.uint32), g["advantages"].view(np.uint32)), "advantages must match the reference loop bit for bit"


def _run_device(seed, B, n_slots, stride, p_alive, p_done, p_learner, gamma, lam, use_learner=True):
    import torch
    from nmmo_b200.rollout import DeviceRollout
    from oracle.rollout_oracle import RolloutOracle
    o = RolloutOracle(B, stride)
    r = DeviceRollout(B, n_slots, stride)
    env_id = np.arange(n_slots)
    for step, inp in step_stream(seed, n_slots, stride, p_alive, p_done, p_learner):
        if o.ptr >= B + 1:
            break
        lm = inp["pool_mask"] if use_learner else None
        o.store(inp["o"], inp["value"], inp["actions"], inp["logprob"], inp["r"], inp["d"], inp["mask"], lm, env_id, step)
        r.store(torch.as_tensor(inp["o"]).cuda(), inp["value"], inp["actions"], inp["logprob"], inp["r"], inp["d"], inp["mask"],
                step, learner_mask=lm)
        assert r.ptr == o.ptr, f"step {step}: rows stored"
    n = o.ptr
    assert np.array_equal(r.obs[:n].cpu().numpy(), o.obs[:n])
    assert np.array_equal(r.actions[:n].cpu().numpy(), o.actions[:n].astype(np.int32))
    for name in ("logprobs", "rewards", "dones", "values"):
        assert np.array_equal(getattr(r, name)[:n].cpu().numpy().view(np.uint32), getattr(o, name)[:n].view(np.uint32)), name
    idxs, adv = o.gae(gamma, lam)
    d_idxs, d_adv = r.gae(gamma, lam)
    torch.cuda.synchronize()
    assert np.array_equal(d_idxs[:n].cpu().numpy(), idxs.astype(np.int32)), "sorted order"
    assert np.array_equal(d_adv[:n - 1].cpu().numpy().view(np.uint32), adv.view(np.uint32)), "advantages (bit patterns)"
    # a second rollout on the same object: reset clears the keys
    r.reset()
    assert r.ptr == 0
    r.close()
    return n


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_matches_oracle_on_golden_streams(case):
    g = np.load(GOLDEN / f"{case}.npz")
    n = _run_device(int(g["seed"]), int(g["batch_size"]), int(g["n_slots"]), int(g["stride"]), float(g["p_alive"]),
                    float(g["p_done"]), float(g["p_learner"]), float(g["gamma"]), float(g["gae_lambda"]))
    assert n == int(g["batch_size"]) + 1


@pytest.mark.gpu
def test_device_larger_and_ragged():
    # more slots than one scan block, capacity cut in the middle of a step, no learner mask
    _run_device(11, 20000, 5000, 256, 0.85, 0.03, 1.0, 0.99, 0.95, use_learner=False)
    # no dones at all: one chain runs over the whole batch
    _run_device(12, 4097, 1500, 64, 0.97, 0.0, 1.0, 0.99, 0.95)
    # every sample ends its chain
    _run_device(13, 1000, 333, 16, 0.9, 1.0, 0.5, 0.9, 0.8)


@pytest.mark.gpu
def test_rollout_from_simulator_records():
    """The append consumes the simulator's own device tensors (no host copy): records of alive agents land in
    slot order and equal the observation rows they were copied from."""
    import torch
    from nmmo_b200.lib import Simulator
    from nmmo_b200.rollout import DeviceRollout
    from util import SMALL, build_world
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=64)
    sim = Simulator(world[0], world[1], 8, *world[2:])
    sim.reset(np.arange(8) + 1)
    n = sim.E * sim.P
    roll = DeviceRollout(600, n, sim.stride)
    stored = 0
    for step in range(1, 12):
        sim.sample_actions(5)
        val = torch.rand(n, device="cuda"); lp = -torch.rand(n, device="cuda")
        before = roll.ptr
        roll.store(sim.obs, val, sim.actions, lp, sim.rewards, sim.terminated.float(), sim.mask, step)
        after = roll.ptr
        alive = sim.mask.nonzero().flatten()[: 601 - before]
        assert after - before == alive.numel()
        assert torch.equal(roll.obs[before:after], sim.obs[alive])
        assert torch.equal(roll.slot[before:after], alive.int())
        assert torch.equal(roll.values[before:after], val[alive])
        stored = after
        sim.step()
    assert stored == 601
    roll.close(); sim.close()


@pytest.mark.gpu
def test_device_evaluate_loop():
    """clean_pufferl.evaluate's loop body on the device (nmmo_b200/evaluate.py): policy reads obs and masks on the
    GPU, the batch is stored and sorted there, the reference's counters come out right."""
    import sys
    import torch
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "examples"))
    from argparse import Namespace
    from device_rollout import SmallPolicy
    from nmmo_b200.evaluate import evaluate
    from nmmo_b200.rollout import DeviceRollout
    from nmmo_b200.vecenv import B200VecEnv
    env_ns = Namespace(num_agents=16, num_npcs=32, max_episode_length=64, maps_path="maps/", map_size=32, num_maps=4,
                       map_force_generation=False, death_fog_tick=None, task_size=64, spawn_immunity=3, resilient_population=0,
                       curriculum_file_path=None)
    pool = B200VecEnv(env_kwargs={"env": env_ns}, num_envs=8, agent="takeru")
    torch.manual_seed(0)
    policy = SmallPolicy(pool.driver_env, hidden=32).cuda()
    n = 8 * 16
    roll = DeviceRollout(700, n, pool.driver_env.obs_sz)
    pool.async_reset(3)
    res = evaluate(pool, policy, roll)
    assert roll.ptr == 701 and res.global_steps == res.steps * n and 701 <= res.agent_steps <= res.global_steps
    # keys are (slot, step) with step ascending per slot; the sorted order groups slots
    idxs, adv = roll.gae(0.99, 0.95)
    torch.cuda.synchronize()
    slot = roll.slot[:701][idxs[:701].long()].cpu().numpy(); st = roll.step[:701][idxs[:701].long()].cpu().numpy()
    key = slot.astype(np.int64) * 100000 + st
    assert np.all(np.diff(key) > 0), "sorted by (slot, step), no duplicates"
    assert torch.isfinite(adv[:700]).all()
    # stored actions were valid under the stored masks
    from nmmo_b200.emulation import unpack_batched_obs
    x = unpack_batched_obs(roll.obs[:701], pool.driver_env.unflatten_context)
    for k, path in enumerate(pool.driver_env.unflatten_context.layout.masks):
        a, b = path.split(".")
        m = x["ActionTargets"][a][b]
        assert bool((m[torch.arange(701, device="cuda"), roll.actions[:701, k].long()] == 1).all()), path
    # a second call continues where the first stopped: every recv() of a call is followed by a send() of the actions
    # that were stored for it (clean_pufferl.py:289,357), so its first recv() is a new tick, and recv() refuses to
    # hand the same outputs out twice
    calls = []
    recv0, send0 = pool.recv, pool.send
    pool.recv = lambda: (calls.append("r"), recv0())[1]
    pool.send = lambda a: (calls.append("s"), send0(a))[1]
    res2 = evaluate(pool, policy, roll)
    pool.recv, pool.send = recv0, send0
    assert roll.ptr == 701 and res2.steps >= 1
    assert "".join(calls) == "rs" * res2.steps
    pool.recv()
    with pytest.raises(RuntimeError):
        pool.recv()
    pool.close(); roll.close()


@pytest.mark.gpu
def test_compact_rollout_reassembles_the_plain_one():
    """Compact storage (Market once per env-step, Task as the task id: SURVEY.md 8(f) rank 1 / VERDICT r1 weak #7) holds the
    same batch as the plain rollout: every scalar array identical, expand() == the plain rows bit for bit, in a third of the bytes."""
    import torch
    from nmmo_b200.lib import Simulator
    from nmmo_b200.rollout import DeviceRollout
    from util import SMALL, build_world
    world = build_world(task_dim=64, **SMALL, NC_HORIZON=60, NC_RES_DEPLETION=1, NC_SPAWN_IMMUNITY=2, NC_WEAPON_DROP_THR=-1)
    E = 6
    sim = Simulator(*world[:2], E, *world[2:])
    n = E * sim.P
    plain = DeviceRollout(500, n, sim.stride)
    comp = DeviceRollout(500, n, sim.stride, compact=DeviceRollout.compact_for(sim, max_steps=40))
    assert comp.row_stride == sim.stride - 384 * 32 - 64 * 2
    sim.reset(np.arange(E, dtype=np.uint64) + 3)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    step = 0
    while plain.ptr < 501 and step < 39:
        step += 1
        val = torch.randn(n, device="cuda", generator=g); lp = -torch.rand(n, device="cuda", generator=g)
        learner = (torch.rand(n, device="cuda", generator=g) < 0.7).to(torch.uint8)
        d = (sim.terminated | sim.truncated).float()
        plain.store(sim.obs, val, sim.actions, lp, sim.rewards, d, sim.mask, step, learner_mask=learner)
        comp.store(sim.obs, val, sim.actions, lp, sim.rewards, d, sim.mask, step, learner_mask=learner, task_id=sim.task_id)
        sim.sample_actions(5); sim.step()                       # listings appear: the Market block is not all zeros
    rows = plain.ptr
    assert rows == comp.ptr and rows > 200
    for name in ("actions", "logprobs", "rewards", "dones", "values", "slot", "step"):
        assert torch.equal(getattr(plain, name)[:rows], getattr(comp, name)[:rows]), name
    full = comp.expand(n=rows)
    assert torch.equal(full, plain.obs[:rows])
    assert int(plain.obs[:rows, :].to(torch.int32).sum()) > 0
    i1, a1 = plain.gae(0.99, 0.95); i2, a2 = comp.gae(0.99, 0.95)
    torch.cuda.synchronize()
    assert torch.equal(i1[:rows], i2[:rows]) and torch.equal(a1[:rows - 1].view(torch.int32), a2[:rows - 1].view(torch.int32))
    pick = i2[:rows][::3].contiguous()
    assert torch.equal(comp.expand(pick), plain.obs[pick.long()])
    with pytest.raises(Exception):
        comp.store(sim.obs, val, sim.actions, lp, sim.rewards, d, sim.mask, step)          # task ids are required
    plain.close(); comp.close(); sim.close()

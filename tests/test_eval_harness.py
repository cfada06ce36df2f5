"""PvE / PvP evaluation harness (/root/reference/evaluate.py:110-221)."""
import json

import numpy as np
import pytest
import torch

from nmmo_b200.eval_harness import EvalRunner, PolicyPool, create_kernel, eval_env_kwargs, unroll_nested_dict


def test_kernel_maps_slots_round_robin_and_shuffles_reproducibly():
    k = create_kernel(128, 3)
    assert len(k) == 128 and k[:6] == [0, 1, 2, 0, 1, 2]
    counts = np.bincount(k, minlength=3)
    assert counts.max() - counts.min() <= 1
    a, b = create_kernel(128, 3, shuffle_with_seed=7), create_kernel(128, 3, shuffle_with_seed=7)
    assert a == b and sorted(a) == sorted(k) and a != k
    assert create_kernel(128, 3, shuffle_with_seed=8) != a
    with pytest.raises(ValueError):
        create_kernel(2, 3)


def test_policy_pool_routes_rows_and_infos_by_policy():
    kernel = [0, 1, 1, 0]
    seen = {}

    def make(name, const):
        def policy(o):
            seen[name] = o[:, 0].tolist()
            B = o.shape[0]
            return torch.full((B, 12), const, dtype=torch.int64), torch.full((B,), float(const)), torch.full((B,), -float(const))
        return policy

    pp = PolicyPool({"a": make("a", 3), "b": make("b", 5)}, kernel, num_envs=2)
    obs = torch.arange(8, dtype=torch.uint8).reshape(8, 1).repeat(1, 4)
    actions, logprob, value = pp.forwards(obs)
    assert seen["a"] == [0, 3, 4, 7] and seen["b"] == [1, 2, 5, 6]
    assert actions[:, 0].tolist() == [3, 5, 5, 3, 3, 5, 5, 3] and actions.dtype == torch.int32
    assert logprob.tolist() == [3, 5, 5, 3, 3, 5, 5, 3] and value.tolist() == [-3, -5, -5, -3, -3, -5, -5, -3]
    split = pp.split_infos([{"x": 1}, {"x": 2}, {"x": 3}], [0, 5, 7])
    assert split == {"a": [{"x": 1}, {"x": 3}], "b": [{"x": 2}]}
    assert dict(unroll_nested_dict({"length": 4, "curriculum": {"task_3": (0.5, 1)}, "stats": {"cod/starved": 1.0}})) == \
        {"length": 4, "curriculum/task_3": (0.5, 1), "stats/cod/starved": 1.0}
    assert int(pp.mask.sum()) == 8


def test_eval_config_follows_the_reference():
    from nmmo_b200.config import SPEC, make_config
    for mode, n_maps in (("pve", 4), ("pvp", 256)):
        kw = eval_env_kwargs(mode)
        assert kw["env"].num_maps == n_maps and kw["env"].num_agents == 128 and kw["env"].num_npcs == 256
        assert kw["env"].max_episode_length == 1024 and kw["env"].spawn_immunity == 20 and kw["env"].resilient_population == 0
        cfg, fcfg = make_config(kw["env"], kw["reward_wrapper"], "neurips23_start_kit")
        assert cfg[SPEC["NC_EVAL_MODE"]] == 1 and cfg[SPEC["NC_EARLY_STOP_N"]] == 0 and cfg[SPEC["NC_RES_RESILIENT_N"]] == 0
        assert fcfg[SPEC["NF_HEAL_W"]] == 0.0 and fcfg[SPEC["NF_EXPLORE_W"]] == 0.0
    with pytest.raises(ValueError):
        eval_env_kwargs("duel")


@pytest.mark.gpu
def test_pvp_eval_on_device(tmp_path):
    from nmmo_b200.config import ObsLayout
    dims = None

    def random_policy(seed):
        g = torch.Generator(device="cuda"); g.manual_seed(seed)

        def policy(o):
            B = o.shape[0]
            a = torch.stack([torch.randint(0, int(n), (B,), generator=g, device=o.device) for n in dims], 1)
            return a, torch.zeros(B, device=o.device), torch.zeros(B, device=o.device)
        return policy

    def idle_policy(o):
        B = o.shape[0]
        a = torch.zeros((B, 12), dtype=torch.int64, device=o.device)
        a[:, 8] = 4                      # Move.Direction 4 = stay
        return a, torch.zeros(B, device=o.device), torch.zeros(B, device=o.device)

    runner = EvalRunner({"rand_a": random_policy(1), "idle": idle_policy, "rand_b": random_policy(2)}, save_dir=str(tmp_path),
                        num_envs=3, num_agents=16, num_npcs=32, horizon=48, map_size=32, task_size=64)
    from nmmo_b200.config import make_config
    kw = eval_env_kwargs("pvp", num_agents=16, num_npcs=32, horizon=48, map_size=32, task_size=64)
    dims = ObsLayout(make_config(kw["env"], kw["reward_wrapper"], "neurips23_start_kit")[0]).action_dims
    results, file_name = runner.perform_eval("pvp", seed=5, num_eval_episode=6, save_file_prefix="eval_pvp", steps_per_call=16)
    assert set(results) == {"rand_a", "idle", "rand_b"} and file_name == "eval_pvp_5.json"
    n_finished = sum(len(v["length"]) for v in results.values())
    assert n_finished >= 6 * 16 // 2                      # every agent of a finished episode reports once
    for pol, vals in results.items():
        assert all(1 <= x <= 48 for x in vals["length"])
        cur = [k for k in vals if k.startswith("curriculum/")]
        assert cur and all(0.0 <= x <= 1.0 for k in cur for x in vals[k])
        assert sum(len(vals[k]) for k in cur) == len(vals["length"])
    saved = json.load(open(tmp_path / file_name))
    assert saved == json.loads(json.dumps(results))
    with pytest.raises(AssertionError):
        EvalRunner({"a": idle_policy, "b": idle_policy}, num_envs=1, num_agents=16, num_npcs=32, horizon=48, map_size=32,
                   task_size=64).setup_evaluator("pve", 1)


@pytest.mark.gpu
def test_pvp_eval_on_the_heldout_tasks(tmp_path):
    """evaluate.py:20,52: evaluation runs on the 63 held-out tasks (real specs and real 2048-d embeddings, translated
    by nmmo_b200/curriculum.py); the result file is keyed by the reference's spec names."""
    from nmmo_b200.curriculum import load_table

    def idle_policy(o):
        B = o.shape[0]
        a = torch.zeros((B, 12), dtype=torch.int64, device=o.device)
        a[:, 8] = 4
        return a, torch.zeros(B, device=o.device), torch.zeros(B, device=o.device)

    runner = EvalRunner({"idle_a": idle_policy, "idle_b": idle_policy}, save_dir=str(tmp_path), num_envs=4, horizon=40)
    pool, pp = runner.setup_evaluator("pvp", seed=3)
    held = load_table("heldout")
    assert pool.task_names == held["names"] and pool.sim.P == 128
    o = pool.recv()[0]
    L = pool.driver_env.unflatten_context.layout
    tid = pool.sim.task_state(0)[0]
    task_block = o[:128, L.o_task:L.o_task + 4096].cpu().numpy().view(np.uint16)
    assert np.array_equal(task_block, held["embed"].view(np.uint16)[tid])      # the policy reads the real embeddings
    pool.send(torch.zeros((4 * 128, 12), dtype=torch.int32, device="cuda"))
    pool.close()
    results, _ = runner.perform_eval("pvp", seed=3, num_eval_episode=4, save_file_prefix="eval_pvp", steps_per_call=16)
    keys = {k for v in results.values() for k in v if k.startswith("curriculum/")}
    assert keys and all(k[len("curriculum/"):] in set(held["names"]) for k in keys)

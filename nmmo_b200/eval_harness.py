"""PvE / PvP evaluation harness on the device env.

Mirrors /root/reference/evaluate.py: ``EvalConfig`` (:33-77) becomes ``eval_env_kwargs``, the policy-pool
kernel that maps agent slots to policies (:119-122) becomes ``create_kernel`` + ``PolicyPool``, and
``EvalRunner.perform_eval`` (:165-193) keeps its loop and its result file: per policy, the list of episode
``length`` values and, per ``curriculum/<task>`` key, the list of max-progress values.

The env, the observation records and the per-agent info records stay on the GPU; each policy sees the
rows of its own agent slots (one ``index_select`` per policy and step) and the actions are merged back
with ``index_copy_``.  Policies are callables ``policy(flat_obs_uint8[B, obs_sz]) -> (actions int[B, 12],
logprob f32[B], value f32[B])`` on the env's device -- the same contract as nmmo_b200.evaluate.

``pufferlib.policy_pool`` (0.7.3) is not vendored in the reference: ``create_kernel`` restates its published
behaviour from the call site (``create_kernel(NUM_AGENTS, len(policies), shuffle_with_seed=seed)``): policy
indices repeated round-robin over the agent slots, optionally shuffled with Python's ``random`` under the
given seed; the same kernel is used for every env.  [parity unpinned for the shuffle order]
"""
from __future__ import annotations

import json
import os
import random
from collections import defaultdict
from typing import Callable, Dict, List, Optional, Sequence

from .config import default_env_args, default_wrapper_args

NUM_AGENTS = 128                 # evaluate.py:18
NUM_PVE_EVAL_EPISODE = 32        # evaluate.py:20
NUM_PVP_EVAL_EPISODE = 200       # evaluate.py:21


def eval_env_kwargs(mode: str, num_agents: int = NUM_AGENTS, num_npcs: int = 256, horizon: int = 1024,
                    map_size: int = 128, **env_over) -> Dict:
    """``EvalConfig`` (evaluate.py:33-77) + the eval-mode RewardWrapper arguments (:82-90) as the
    ``env_kwargs`` of B200VecEnv.  PvE uses 4 maps, PvP 256 (:70-75)."""
    if mode not in ("pve", "pvp"):
        raise ValueError(f"Invalid eval_mode: {mode}")
    env = default_env_args(num_agents=num_agents, num_npcs=num_npcs, max_episode_length=horizon, map_size=map_size,
                           num_maps=4 if mode == "pve" else 256, death_fog_tick=None, resilient_population=0,
                           spawn_immunity=20, **env_over)
    # only eval_mode / early_stop are passed (:84-89): the bonus weights keep their constructor defaults of 0
    wrap = default_wrapper_args("neurips23_start_kit", eval_mode=True, early_stop_agent_num=0, heal_bonus_weight=0,
                                explore_bonus_weight=0)
    return {"env": env, "reward_wrapper": wrap}


def create_kernel(agents_per_env: int, num_policies: int, shuffle_with_seed: Optional[int] = None) -> List[int]:
    """Agent slot -> policy index, one entry per agent of an env (evaluate.py:121)."""
    if num_policies <= 0 or agents_per_env < num_policies:
        raise ValueError("need 1 <= num_policies <= agents_per_env")
    kernel = [k % num_policies for k in range(agents_per_env)]
    if shuffle_with_seed is not None:
        random.Random(shuffle_with_seed).shuffle(kernel)
    return kernel


def unroll_nested_dict(d, prefix=""):
    """(flattened key, leaf) pairs with '/'-joined keys, as clean_pufferl.py:352-354 consumes them."""
    for k, v in d.items():
        if isinstance(v, dict):
            yield from unroll_nested_dict(v, f"{prefix}{k}/")
        else:
            yield f"{prefix}{k}", v


class PolicyPool:
    """The part of pufferlib's PolicyPool the evaluation uses: ``forwards`` over per-policy row sets and the
    split of the finished agents' infos by policy."""

    def __init__(self, policies: Dict[str, Callable], kernel: Sequence[int], num_envs: int, device=None):
        import torch
        if len(set(kernel)) > len(policies) or max(kernel) >= len(policies):
            raise ValueError("kernel refers to more policies than were given")
        self.names = list(policies)
        self.policies = policies
        self.kernel = list(kernel)
        self.agents_per_env = len(self.kernel)
        full = torch.tensor(self.kernel * int(num_envs), dtype=torch.int64)
        self.sample_idxs = [(full == k).nonzero().flatten().to(device) for k in range(len(self.names))]
        self.mask = torch.ones(full.numel(), dtype=torch.uint8, device=device)      # evaluate.py:168: all slots count

    def forwards(self, obs):
        import torch
        B = obs.shape[0]
        actions = torch.zeros((B, 12), dtype=torch.int32, device=obs.device)
        logprob = torch.zeros(B, dtype=torch.float32, device=obs.device)
        value = torch.zeros(B, dtype=torch.float32, device=obs.device)
        for name, idx in zip(self.names, self.sample_idxs):
            if idx.numel() == 0:
                continue
            a, lp, v = self.policies[name](obs.index_select(0, idx))
            actions.index_copy_(0, idx, a.to(torch.int32).reshape(-1, 12))
            logprob.index_copy_(0, idx, lp.to(torch.float32).reshape(-1))
            value.index_copy_(0, idx, v.to(torch.float32).reshape(-1))
        return actions, logprob, value

    def policy_of_slot(self, flat_slot: int) -> str:
        return self.names[self.kernel[int(flat_slot) % self.agents_per_env]]

    def split_infos(self, infos: List[Dict], slots) -> Dict[str, List[Dict]]:
        out: Dict[str, List[Dict]] = {}
        for info, slot in zip(infos, slots):
            out.setdefault(self.policy_of_slot(slot), []).append(info)
        return out


class EvalRunner:
    """``EvalRunner`` of evaluate.py:110-221 on B200VecEnv.  ``policies`` maps a policy name to a callable
    (the reference loads them from ``policy_store_dir``; checkpoints are outside this repo's scope)."""

    def __init__(self, policies: Dict[str, Callable], save_dir: Optional[str] = None, num_envs: int = 6,
                 device: int = 0, debug: bool = False, **env_over):
        if not policies:
            raise AssertionError("No policies found in eval_model_path")
        self.policies, self.save_dir, self.device = policies, save_dir, device
        self.num_envs = 1 if debug else int(num_envs)          # get_eval_config (evaluate.py:24-29)
        self._debug = debug
        self.env_over = env_over

    def setup_evaluator(self, mode: str, seed: int, task_rows=None, curriculum: Optional[str] = None):
        from .vecenv import B200VecEnv
        if mode == "pve":
            assert len(self.policies) == 1, "PvE mode requires only one policy"
        kw = eval_env_kwargs(mode, **self.env_over)
        if task_rows is None and curriculum is None and int(kw["env"].task_size) == 2048:
            curriculum = "heldout"       # EVAL_TASK_FILE = heldout_task_with_embedding.pkl, TASK_EMBED_DIM 2048 (evaluate.py:20,52-53)
        pool = B200VecEnv(env_kwargs=kw, num_envs=self.num_envs, agent="neurips23_start_kit", device=self.device,
                          task_rows=task_rows, curriculum=curriculum, collect_infos="async")
        kernel = create_kernel(pool.agents_per_env, len(self.policies), shuffle_with_seed=seed)
        policy_pool = PolicyPool(self.policies, kernel, self.num_envs, device=pool.sim.obs.device)
        pool.async_reset(seed)
        return pool, policy_pool

    @staticmethod
    def evaluate(pool, policy_pool: PolicyPool, steps: int):
        """``steps`` env steps of clean_pufferl.evaluate's loop (:287-357) without the storage part:
        recv, per-policy forwards, send; infos grouped by policy and flattened key (:352-355)."""
        import torch
        infos = defaultdict(lambda: defaultdict(list))
        for _ in range(steps):
            o, r, d, t, i, env_id, mask = pool.recv()
            for pol, agent_infos in policy_pool.split_infos(i, pool.last_info_slots).items():
                for agent_i in agent_infos:
                    for name, dat in unroll_nested_dict(agent_i):
                        infos[pol][name].append(dat)
            with torch.no_grad():
                actions, _, _ = policy_pool.forwards(o)
            pool.send(actions)
        # (with collect_infos="async" the records of the last step are handed out by the next call's first recv())
        return infos

    def perform_eval(self, mode: str, seed: int, num_eval_episode: int, save_file_prefix: str, task_rows=None,
                     steps_per_call: int = 64, max_calls: int = 1 << 20, curriculum: Optional[str] = None):
        pool, policy_pool = self.setup_evaluator(mode, seed, task_rows, curriculum)
        eval_results: Dict[str, Dict[str, list]] = {}
        cnt_episode, calls = 0, 0
        try:
            while cnt_episode < num_eval_episode and calls < max_calls:
                calls += 1
                infos = self.evaluate(pool, policy_pool, steps_per_call)
                for pol, vals in infos.items():
                    cnt_episode += sum(vals.get("episode_done", []))
                    if pol not in eval_results:
                        eval_results[pol] = defaultdict(list)
                    for k, v in vals.items():
                        if k == "length":
                            eval_results[pol][k] += v
                        if k.startswith("curriculum"):
                            eval_results[pol][k] += [vv[0] for vv in v]
        finally:
            pool.close()
        file_name = f"{save_file_prefix}_{seed}.json"
        if self.save_dir is not None:
            with open(os.path.join(self.save_dir, file_name), "w") as f:
                json.dump(eval_results, f)
        return eval_results, file_name

    def run(self, mode: str, seed: Optional[int] = None, num_episode: Optional[int] = None,
            save_file_prefix: Optional[str] = None, task_rows=None, curriculum: Optional[str] = None):
        assert mode in ("pve", "pvp"), f"Invalid mode: {mode}"
        num_episode = num_episode or (NUM_PVE_EVAL_EPISODE if mode == "pve" else NUM_PVP_EVAL_EPISODE)
        save_file_prefix = save_file_prefix or ("eval_pve" if mode == "pve" else "eval_pvp")
        if self._debug:
            num_episode = 4
        if seed is None:
            seed = random.randint(10000000, 99999999)
        return self.perform_eval(mode, seed, num_episode, save_file_prefix, task_rows=task_rows, curriculum=curriculum)

"""Per-environment PettingZoo-parallel view of one simulated env (parity / debugging boundary).

Mirrors what the reference's wrappers see of ``nmmo.Env``: ``reset(seed=) -> (obs, info)``,
``step({agent_id: nested action dict}) -> (obs, rewards, terminated, truncated, infos)``,
``agents`` / ``possible_agents`` with 1-based ids (/root/reference/train_helper.py:147,
/root/reference/reinforcement_learning/stat_wrapper.py:48-97).  The values already include the
RewardWrapper hooks, i.e. this is the env *as wrapped by* ``make_env_creator``
(/root/reference/reinforcement_learning/environment.py:55-74) minus the pufferlib flattening.
Everything is copied to the host: use B200VecEnv for rollouts.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .config import SPEC
from .emulation import UnflattenContext, unpack_batched_obs
from .vecenv import info_record_to_dict

ACTION_KEYS = [("Attack", "Style"), ("Attack", "Target"), ("Buy", "MarketItem"), ("Destroy", "InventoryItem"),
               ("Give", "InventoryItem"), ("Give", "Target"), ("GiveGold", "Price"), ("GiveGold", "Target"),
               ("Move", "Direction"), ("Sell", "InventoryItem"), ("Sell", "Price"), ("Use", "InventoryItem")]


class EnvView:
    def __init__(self, sim, env_index: int = 0):
        self.sim, self.k = sim, int(env_index)
        self.P = sim.P
        self.ctx = UnflattenContext(sim.cfg)
        self.possible_agents = list(range(1, self.P + 1))
        self.agents = []
        self.max_num_agents = self.P

    def _slice(self, t):
        return t[self.k * self.P:(self.k + 1) * self.P].cpu().numpy()

    def _obs(self, present):
        flat = self._slice(self.sim.obs)
        nested = unpack_batched_obs(flat, self.ctx)

        def pick(x, p):
            return {k: pick(v, p) for k, v in x.items()} if isinstance(x, dict) else x[p]

        return {p + 1: pick(nested, p) for p in range(self.P) if present[p]}

    def reset(self, seed=0, **kw):
        mask = np.zeros(self.sim.E, np.uint8); mask[self.k] = 1
        seeds = np.zeros(self.sim.E, np.uint64); seeds[self.k] = seed
        self.sim.reset(seeds, env_mask=mask)
        present = self._slice(self.sim.mask).astype(bool)
        self.agents = [p + 1 for p in range(self.P) if present[p]]
        return self._obs(present), {a: {} for a in self.agents}

    def step(self, actions: Dict[int, Dict]):
        if self.sim.E != 1:
            raise RuntimeError("EnvView.step drives the whole simulator; create it with n_envs=1")
        import torch
        flat = np.zeros((1, self.P, 12), np.int32)
        for a, d in actions.items():
            for k, (x, y) in enumerate(ACTION_KEYS):
                flat[0, a - 1, k] = int(d.get(x, {}).get(y, 0))
        self.sim.actions.copy_(torch.from_numpy(flat))
        self.sim.step()
        present = self._slice(self.sim.mask).astype(bool)
        rew, term, trunc = self._slice(self.sim.rewards), self._slice(self.sim.terminated), self._slice(self.sim.truncated)
        iv, info = self._slice(self.sim.info_valid), self._slice(self.sim.info)
        done = bool(self.sim.episode_done[self.k].item())
        ids = [p + 1 for p in range(self.P) if present[p]]
        infos = {a: (info_record_to_dict(info[a - 1]) if iv[a - 1] else {}) for a in ids}
        if done and ids:
            infos[ids[-1]]["episode_done"] = True
        self.agents = [] if done else [a for a in ids if not term[a - 1]]
        return (self._obs(present), {a: float(rew[a - 1]) for a in ids}, {a: bool(term[a - 1]) for a in ids},
                {a: bool(trunc[a - 1]) for a in ids}, infos)

    @property
    def tick(self):
        return int(self.sim.snapshot(self.k)[3][0])

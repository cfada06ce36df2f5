"""CPU restatement of the reference's rollout storage, sort and GAE loop.  TEST INFRASTRUCTURE ONLY:
imported by tests/ and bench.py's cpu_baseline leg, never by the product path (nmmo_b200/).

Follows /root/reference/reinforcement_learning/clean_pufferl.py:
  :183-197  storage arrays of batch_size + 1 rows
  :329-348  masked append of one recv() + sort_keys
  :413-414  idxs = sorted(range(len(sort_keys)), key=sort_keys.__getitem__)
  :424-436  GAE over the sorted samples (float32 tensor arithmetic, Python-float hyper-parameters)
Pinned by tests/golden/rollout_*.npz, which were produced by executing those very source lines of the
reference (tests/golden/make_golden_rollout.py).
"""
from __future__ import annotations

import numpy as np


class RolloutOracle:
    def __init__(self, batch_size: int, obs_stride: int):
        B = self.batch_size = int(batch_size)
        self.obs = np.zeros((B + 1, obs_stride), np.uint8)            # :183
        self.actions = np.zeros((B + 1, 12), np.int64)                # :184
        self.logprobs = np.zeros(B + 1, np.float32)                   # :185
        self.rewards = np.zeros(B + 1, np.float32)                    # :186
        self.dones = np.zeros(B + 1, np.float32)                      # :187
        self.values = np.zeros(B + 1, np.float32)                     # :189
        self.sort_keys = []
        self.ptr = 0

    def reset(self):
        self.sort_keys = []                                            # :414
        self.ptr = 0                                                   # :278

    def store(self, o, value, actions, logprob, r, d, mask, learner_mask, env_id, step):
        """One pass of the loop body at :329-348."""
        B, ptr = self.batch_size, self.ptr
        learner = np.asarray(mask, np.float32) * (1.0 if learner_mask is None else np.asarray(learner_mask, np.float32))   # :333
        indices = np.where(learner)[0][: B - ptr + 1]                  # :336
        end = ptr + len(indices)                                       # :337
        self.obs[ptr:end] = o[indices]                                 # :340
        self.values[ptr:end] = value[indices]                          # :341
        self.actions[ptr:end] = actions[indices]                       # :342
        self.logprobs[ptr:end] = logprob[indices]                      # :343
        self.rewards[ptr:end] = r[indices]                             # :344
        self.dones[ptr:end] = d[indices]                               # :345
        self.sort_keys.extend([(int(env_id[i]), step) for i in indices])   # :346
        self.ptr = end                                                 # :349

    def sorted_idxs(self):
        return sorted(range(len(self.sort_keys)), key=self.sort_keys.__getitem__)   # :413

    def gae(self, gamma: float, gae_lambda: float, n: int | None = None):
        """:424-436.  torch multiplies a float32 tensor by a Python float in float32 (the scalar is
        rounded to float32 first); `gamma * gae_lambda` is a double product rounded once."""
        idxs = self.sorted_idxs()
        n = len(idxs) - 1 if n is None else n
        adv = np.zeros(n, np.float32)
        g = np.float32(gamma); gl = np.float32(gamma * gae_lambda)
        one = np.float32(1.0)
        last = np.float32(0.0)
        for t in reversed(range(n)):
            i, i_nxt = idxs[t], idxs[t + 1]
            nnt = one - self.dones[i_nxt]
            nextvalues = self.values[i_nxt]
            delta = self.rewards[i_nxt] + g * nextvalues * nnt - self.values[i]
            adv[t] = last = delta + gl * nnt * last
        return np.asarray(idxs, np.int64), adv

#!/usr/bin/env python
"""Config 4 of BASELINE.json in miniature: a policy reads observations and masks on the device, samples
masked actions, the batch is stored and GAE is computed on the device -- nothing but counters reaches the host.

    python examples/device_rollout.py --envs 256 --batch 32768

The policy below is a small stand-in with the interface of the reference's policies (encode the unpacked
observation, one masked categorical head per action key, agent_zoo/takeru/policy.py:39-83,351-405).
"""
import argparse
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nmmo_b200.emulation import unpack_batched_obs  # noqa: E402
from nmmo_b200.evaluate import evaluate  # noqa: E402
from nmmo_b200.rollout import DeviceRollout  # noqa: E402
from nmmo_b200.vecenv import B200VecEnv  # noqa: E402


class SmallPolicy(torch.nn.Module):
    def __init__(self, driver_env, hidden=64):
        super().__init__()
        self.ctx = driver_env.unflatten_context
        L = self.ctx.layout
        self.heads = list(L.masks.keys())
        self.tile = torch.nn.Embedding(16, 8)
        self.me = torch.nn.Linear(31, hidden)
        self.task = torch.nn.Linear(L.task_dim, hidden)
        self.fc = torch.nn.Linear(8 + 2 * hidden, hidden)
        self.out = torch.nn.ModuleList([torch.nn.Linear(hidden, n) for (_, n) in L.masks.values()])
        self.value = torch.nn.Linear(hidden, 1)

    def forward(self, flat):
        x = unpack_batched_obs(flat, self.ctx)
        ids = x["Entity"][:, :, 0]
        mine = ((ids == x["AgentId"][:, :1]) & (ids != 0)).int().argmax(dim=1)
        me = x["Entity"][torch.arange(flat.shape[0], device=flat.device), mine].float() / 100.0
        tile = self.tile(x["Tile"][:, :, 2].long().clamp(0, 15)).mean(dim=1)
        h = torch.relu(self.fc(torch.cat([tile, torch.relu(self.me(me)), torch.relu(self.task(x["Task"].float()))], dim=-1)))
        acts, logp = [], 0.0
        for path, layer in zip(self.heads, self.out):
            a, b = path.split(".")
            logits = layer(h).masked_fill(x["ActionTargets"][a][b] == 0, -1e9)      # policy.py:332-333
            dist = torch.distributions.Categorical(logits=logits)
            pick = dist.sample()
            acts.append(pick); logp = logp + dist.log_prob(pick)
        return torch.stack(acts, dim=1).int(), logp, self.value(h).flatten()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--rollouts", type=int, default=3)
    a = ap.parse_args()
    pool = B200VecEnv(num_envs=a.envs, agent="takeru", collect_infos=False)
    policy = SmallPolicy(pool.driver_env).cuda()
    roll = DeviceRollout(a.batch, a.envs * pool.agents_per_env, pool.driver_env.obs_sz)
    pool.async_reset(1)
    for k in range(a.rollouts):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = evaluate(pool, policy, roll)
        idxs, adv = roll.gae(0.99, 0.95)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"rollout {k}: {res.steps} steps, {res.agent_steps} agent steps ({res.agent_steps / dt:,.0f}/s), "
              f"{res.global_steps} slot steps, mean advantage {adv[:a.batch].mean().item():+.4f}")
    pool.close(); roll.close()


if __name__ == "__main__":
    main()

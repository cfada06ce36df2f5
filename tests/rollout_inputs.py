"""Seeded synthetic recv() streams for the rollout-storage tests (shared by the golden generator and the tests)."""
import numpy as np


def step_stream(seed, n_slots, stride, p_alive, p_done, p_learner):
    """Yields (step, dict) forever: one dict per recv() with o, value, actions, logprob, r, d, mask, pool_mask."""
    rng = np.random.default_rng(seed)
    alive = rng.random(n_slots) < 0.9
    step = 0
    while True:
        step += 1
        alive = alive & (rng.random(n_slots) < p_alive) | (rng.random(n_slots) < 0.02)
        mask = alive.astype(np.float32)
        if step % 5 == 0:
            mask[:] = 0           # a step in which nothing is alive
        pool_mask = (rng.random(n_slots) < p_learner).astype(np.float32)
        yield step, dict(o=rng.integers(0, 256, (n_slots, stride), dtype=np.uint8), value=rng.normal(size=n_slots).astype(np.float32),
                         actions=rng.integers(0, 100, (n_slots, 12)).astype(np.int64), logprob=-rng.random(n_slots).astype(np.float32),
                         r=(rng.normal(size=n_slots) * (rng.random(n_slots) < 0.3)).astype(np.float32),
                         d=(rng.random(n_slots) < p_done).astype(np.float32), mask=mask, pool_mask=pool_mask)

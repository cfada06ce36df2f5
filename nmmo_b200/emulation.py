"""Flat observation record <-> nested observation dict (the pufferlib emulation seam).

The reference's policies call ``pufferlib.emulation.unpack_batched_obs(flat_obs, unflatten_context)``
(/root/reference/agent_zoo/takeru/policy.py:27,40-42) to turn the flat batch back into the nested
nmmo observation (Tile, Entity, Inventory, Market, Task, AgentId, CurrentTick, ActionTargets).
Here the flat record is the native-dtype byte record of include/nmmo_spec.h (``nm_obs_layout``);
``unpack_batched_obs`` returns zero-copy typed views of it, on whatever device the batch lives on.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .config import ObsLayout


class UnflattenContext:
    """What ``driver_env.unflatten_context`` carries: the record layout."""

    def __init__(self, cfg: np.ndarray):
        self.layout = ObsLayout(cfg)
        L = self.layout
        self.obs_sz = L.stride
        self.fields = {
            "AgentId": (L.o_ids, 2, "int16", (1,)),
            "CurrentTick": (L.o_ids + 2, 2, "int16", (1,)),
            "Entity": (L.o_entity, L.n_ent * 31 * 2, "int16", (L.n_ent, 31)),
            "Inventory": (L.o_inventory, L.n_inv * 16 * 2, "int16", (L.n_inv, 16)),
            "Market": (L.o_market, L.n_mkt * 16 * 2, "int16", (L.n_mkt, 16)),
            "Task": (L.o_task, L.task_dim * 2, "float16", (L.task_dim,)),
            "Tile": (L.o_tile, L.win * L.win * 3 * 2, "int16", (L.win * L.win, 3)),
        }


def unpack_batched_obs(flat, ctx: UnflattenContext) -> Dict:
    """flat: uint8 [B, obs_sz] torch tensor (any device) or numpy array -> nested dict of views."""
    L = ctx.layout
    is_np = isinstance(flat, np.ndarray)
    if not is_np:
        import torch
        dt = {"int16": torch.int16, "float16": torch.float16, "int8": torch.int8}
    B = flat.shape[0]

    def view(off, nbytes, dtype, shape):
        sl = flat[:, off:off + nbytes]
        if is_np:
            return np.ascontiguousarray(sl).view(dtype).reshape((B,) + shape) if not sl.flags.c_contiguous else sl.view(dtype).reshape((B,) + shape)
        return sl.view(dt[dtype]).reshape((B,) + shape)

    out = {name: view(*spec) for name, spec in ctx.fields.items()}
    targets: Dict = {}
    for path, (off, n) in L.masks.items():
        a, b = path.split(".")
        targets.setdefault(a, {})[b] = view(off, n, "int8", (n,))
    out["ActionTargets"] = targets
    return out

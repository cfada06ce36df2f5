"""The takeru policy (the consumer BASELINE.json configs[3] names) running on the device records.

Same network as /root/reference/agent_zoo/takeru/policy.py (``ReducedModelV2`` :21-82, encoders :85-287,
``ReducedActionDecoder`` :290-405) with the same module and parameter names, so a reference checkpoint's
``state_dict`` loads unchanged and -- tests/test_takeru_policy.py -- the reference module and this one produce the same
logits and values from the same weights.  It is written against the byte record of include/nmmo_spec.h rather than
against pufferlib's emulation layer, and for batches of hundreds of thousands of agent slots:

* the record is sliced into typed zero-copy views (nmmo_b200.emulation) -- no flatten / unflatten copies;
* the Market block is identical for every agent of an environment, so its item encoder and the Buy head's key matrix
  are computed once per environment ([E, 384, .] instead of [E * 128, 384, .]: 128 x less work and memory);
* only the occupied Entity rows (id != 0; typically 5-15 of the 100) go through the player encoder: the logits of
  the three target heads are row-wise dot products scattered into a zero [B, 101] tensor.  Empty rows are always
  masked (ActionTargets only ever offers visible entities), so the *masked* logits are identical to the reference's,
  at a tenth of the work and without the [B, 100, hidden] activation;
* the forward pass runs in chunks of environments so that activations stay bounded;
* ``forward`` samples the 12 masked categorical heads on the device and returns (actions, logprob, value), the
  contract of nmmo_b200.evaluate / eval_harness.

Reference quirks that are part of the function and therefore kept: discrete entity / item attributes are clipped to
0..255 *after* their per-attribute offsets are added (policy.py:186-187, :247-248), and the continuous scales of
``alchemy_level`` / ``alchemy_exp`` are swapped (:158-159).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import SPEC
from .emulation import UnflattenContext, unpack_batched_obs

_ENT = {k[3:].lower(): v for k, v in SPEC.items() if k.startswith("EA_")}
ENT_DISCRETE = ["id", "npc_type", "attacker_id", "message"]
ENT_CONTINUOUS = [("row", 256), ("col", 256), ("damage", 100), ("time_alive", 1024), ("freeze", 3), ("item_level", 50),
                  ("latest_combat_tick", 1024), ("gold", 100), ("health", 100), ("food", 100), ("water", 100)]
for _s in ("melee", "range", "mage", "fishing", "herbalism", "prospecting", "carving"):
    ENT_CONTINUOUS += [(f"{_s}_level", 10), (f"{_s}_exp", 100)]
ENT_CONTINUOUS += [("alchemy_level", 100), ("alchemy_exp", 10)]          # (sic) policy.py:158-159
ITEM_DISCRETE, ITEM_DISCRETE_OFFSET = [1, 14], [2.0, 0.0]
ITEM_CONTINUOUS = [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15]
ITEM_SCALE = [10, 10, 10, 100, 100, 100, 40, 40, 40, 100, 100, 100]
HEADS = ["attack_style", "attack_target", "market_buy", "inventory_destroy", "inventory_give_item", "inventory_give_player",
         "gold_quantity", "gold_target", "move", "inventory_sell", "inventory_price", "inventory_use"]      # action-space order, policy.py:293-307
HEAD_MASK = {"attack_style": ("Attack", "Style"), "attack_target": ("Attack", "Target"), "market_buy": ("Buy", "MarketItem"),
             "inventory_destroy": ("Destroy", "InventoryItem"), "inventory_give_item": ("Give", "InventoryItem"),
             "inventory_give_player": ("Give", "Target"), "gold_quantity": ("GiveGold", "Price"), "gold_target": ("GiveGold", "Target"),
             "move": ("Move", "Direction"), "inventory_sell": ("Sell", "InventoryItem"), "inventory_price": ("Sell", "Price"),
             "inventory_use": ("Use", "InventoryItem")}
HEAD_KEYS = {"attack_target": "player", "market_buy": "market", "inventory_destroy": "inventory", "inventory_give_item": "inventory",
             "inventory_give_player": "player", "gold_target": "player", "inventory_sell": "inventory", "inventory_use": "inventory"}


class _TileEncoder(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.embedding = nn.Embedding(256, 32)
        self.tile_conv_1 = nn.Conv2d(32, 16, 3)
        self.tile_conv_2 = nn.Conv2d(16, 8, 3)
        self.tile_fc = nn.Linear(8 * 11 * 11, input_size)

    def forward(self, tile):                                   # int16 [B, 225, 3]: only the material column is used
        mat = tile[:, :, 2].long().clamp_(0, 255)
        x = self.embedding(mat).view(-1, 15, 15, 32).permute(0, 3, 1, 2)       # [B, 32, 15, 15]
        x = F.relu(self.tile_conv_1(x))
        x = F.relu(self.tile_conv_2(x))
        return F.relu(self.tile_fc(x.flatten(1)))


class _PlayerEncoder(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.embedding = nn.Embedding(len(ENT_DISCRETE) * 256, 32)
        dim = len(ENT_DISCRETE) * 32 + len(ENT_CONTINUOUS)
        self.agent_fc = nn.Linear(dim, hidden_size)
        self.my_agent_fc = nn.Linear(dim, input_size)
        self.register_buffer("d_idx", torch.tensor([_ENT[k] for k in ENT_DISCRETE]), persistent=False)
        self.register_buffer("d_off", torch.tensor([256.0 * i for i in range(len(ENT_DISCRETE))]), persistent=False)
        self.register_buffer("c_idx", torch.tensor([_ENT[k] for k, _ in ENT_CONTINUOUS]), persistent=False)
        self.register_buffer("c_scale", torch.tensor([float(s) for _, s in ENT_CONTINUOUS]), persistent=False)

    def features(self, rows):                                  # int16 [..., 31] -> f32 [..., 155]
        disc = (rows.index_select(-1, self.d_idx).float() + self.d_off).long().clamp_(0, 255)    # offsets are clipped away
        return torch.cat([self.embedding(disc).flatten(-2), rows.index_select(-1, self.c_idx).float() / self.c_scale], -1)

    def my_row(self, ent, my_id):
        ids = ent[:, :, _ENT["id"]]
        mine = (ids == my_id.unsqueeze(1)) & (ids != 0)
        return torch.where(mine.any(1), mine.int().argmax(1), torch.zeros_like(my_id, dtype=torch.int64))

    def forward(self, ent, my_id):                             # int16 [B, N, 31], int16 [B]: the dense form of the reference
        feats = self.features(ent)
        me = feats[torch.arange(ent.shape[0], device=ent.device), self.my_row(ent, my_id)]
        return self.agent_fc(feats), F.relu(self.my_agent_fc(me))

    def forward_sparse(self, ent, my_id):
        """-> (keys [M, H] of the occupied rows, their (batch, row) indices, my_agent [B, input])."""
        b_idx, r_idx = (ent[:, :, _ENT["id"]] != 0).nonzero(as_tuple=True)
        keys = self.agent_fc(self.features(ent[b_idx, r_idx]))
        me = self.features(ent[torch.arange(ent.shape[0], device=ent.device), self.my_row(ent, my_id)])
        return keys, b_idx, r_idx, F.relu(self.my_agent_fc(me))


class _ItemEncoder(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.embedding = nn.Embedding(256, 32)
        self.fc = nn.Linear(2 * 32 + 12, hidden_size)
        self.register_buffer("d_idx", torch.tensor(ITEM_DISCRETE), persistent=False)
        self.register_buffer("d_off", torch.tensor(ITEM_DISCRETE_OFFSET), persistent=False)
        self.register_buffer("c_idx", torch.tensor(ITEM_CONTINUOUS), persistent=False)
        self.register_buffer("c_scale", torch.tensor([float(s) for s in ITEM_SCALE]), persistent=False)

    def forward(self, items):                                  # int16 [B, n, 16]
        disc = (items.index_select(2, self.d_idx).float() + self.d_off).long().clamp_(0, 255)
        feats = torch.cat([self.embedding(disc).flatten(2), items.index_select(2, self.c_idx).float() / self.c_scale], -1)
        return self.fc(feats)


class _Fc(nn.Module):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.fc = nn.Linear(n_in, n_out)


class _Decoder(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        out = {h: hidden_size for h in HEADS}
        out.update(attack_style=3, gold_quantity=99, move=5, inventory_price=99)
        self.layers = nn.ModuleDict({h: nn.Linear(hidden_size, out[h]) for h in HEADS})


class TakeruPolicy(nn.Module):
    """policy(flat_obs uint8 [B, obs_sz]) -> (actions int32 [B, 12], logprob f32 [B], value f32 [B])."""

    def __init__(self, ctx: UnflattenContext, input_size=256, hidden_size=256, task_size=2048, agents_per_env: int = 128,
                 envs_per_chunk: int = 64, dedup_market: bool = True, sparse_entities: bool = True):
        super().__init__()
        self.ctx = ctx
        self.P, self.envs_per_chunk, self.dedup_market = int(agents_per_env), int(envs_per_chunk), dedup_market
        self.sparse_entities = sparse_entities
        self.tile_encoder = _TileEncoder(input_size)
        self.player_encoder = _PlayerEncoder(input_size, hidden_size)
        self.item_encoder = _ItemEncoder(input_size, hidden_size)
        self.inventory_encoder = _Fc(12 * hidden_size, input_size)
        self.market_encoder = _Fc(hidden_size, input_size)
        self.task_encoder = _Fc(task_size, input_size)
        self.proj_fc = nn.Linear(5 * input_size, hidden_size)
        self.action_decoder = _Decoder(input_size, hidden_size)
        self.value_head = nn.Linear(hidden_size, 1)

    # ---- one chunk -------------------------------------------------------------------------------------------------
    def _chunk_logits(self, flat) -> (List[torch.Tensor], torch.Tensor):
        o = unpack_batched_obs(flat, self.ctx)
        B = flat.shape[0]
        tile = self.tile_encoder(o["Tile"])
        if self.sparse_entities:
            players, pb, pr, me = self.player_encoder.forward_sparse(o["Entity"], o["AgentId"][:, 0])
        else:
            players, me = self.player_encoder(o["Entity"], o["AgentId"][:, 0])
        inv_items = self.item_encoder(o["Inventory"])
        inventory = self.inventory_encoder.fc(inv_items.flatten(1))
        per_env = self.dedup_market and B % self.P == 0
        if per_env:
            # every alive agent of an env carries the same Market block (dead slots carry zeros): encode it once per env
            E = B // self.P
            alive = (o["AgentId"][:, 0] != 0).view(E, self.P)
            first = alive.int().argmax(1) + torch.arange(E, device=flat.device) * self.P
            mkt_items = self.item_encoder(o["Market"].index_select(0, first))                    # [E, n_mkt, H]
            market = self.market_encoder.fc(mkt_items).mean(-2).repeat_interleave(self.P, 0)     # [B, input]
        else:
            mkt_items = self.item_encoder(o["Market"])
            market = self.market_encoder.fc(mkt_items).mean(-2)
        task = self.task_encoder.fc(o["Task"].float())
        hidden = self.proj_fc(torch.cat([tile, me, inventory, market, task], -1))
        keys = {"player": players, "inventory": inv_items, "market": mkt_items}
        masks = o["ActionTargets"]
        logits = []
        for h in HEADS:
            a, b = HEAD_MASK[h]
            mask = masks[a][b]
            q = self.action_decoder.layers[h](hidden)                                            # [B, H] or [B, n]
            if h in HEAD_KEYS:
                k = keys[HEAD_KEYS[h]]
                if HEAD_KEYS[h] == "player" and self.sparse_entities:
                    # row-wise dot products of the occupied rows, scattered; everything else (incl. the no-op column) is 0
                    lg = torch.zeros((B, mask.shape[1]), dtype=q.dtype, device=q.device)
                    lg.index_put_((pb, pr), (k * q.index_select(0, pb)).sum(-1))
                elif h == "market_buy" and per_env:
                    lg = torch.bmm(q.view(-1, self.P, q.shape[-1]), k.transpose(1, 2)).reshape(B, -1)      # [E, P, H] x [E, H, n]
                else:
                    lg = torch.bmm(k, q.unsqueeze(-1)).squeeze(-1)
                if lg.shape[1] != mask.shape[1]:                                                 # the no-op column: a zero key
                    lg = F.pad(lg, (0, mask.shape[1] - lg.shape[1]))
            else:
                lg = q
            logits.append(lg.masked_fill(mask == 0, -1e9))
        return logits, self.value_head(hidden)

    def logits(self, flat_obs):
        """(list of 12 masked logit tensors, value [B, 1]) of the whole batch (test / small-batch use)."""
        return self._chunk_logits(flat_obs)

    @torch.no_grad()
    def forward(self, flat_obs, generator: Optional[torch.Generator] = None):
        B = flat_obs.shape[0]
        step = self.envs_per_chunk * self.P
        actions = torch.empty((B, len(HEADS)), dtype=torch.int32, device=flat_obs.device)
        logprob = torch.empty(B, dtype=torch.float32, device=flat_obs.device)
        value = torch.empty(B, dtype=torch.float32, device=flat_obs.device)
        for lo in range(0, B, step):
            hi = min(B, lo + step)
            logits, v = self._chunk_logits(flat_obs[lo:hi])
            lp_sum = torch.zeros(hi - lo, dtype=torch.float32, device=flat_obs.device)
            for k, lg in enumerate(logits):
                logp = F.log_softmax(lg.float(), -1)
                # Gumbel-max sample of the categorical (what torch.distributions.Categorical.sample does, without the host sync)
                u = torch.rand(lg.shape, device=lg.device, generator=generator).clamp_(1e-12, 1.0)
                a = (logp - torch.log(-torch.log(u))).argmax(-1)
                actions[lo:hi, k] = a.int()
                lp_sum += logp.gather(1, a.unsqueeze(1)).squeeze(1)
            logprob[lo:hi] = lp_sum
            value[lo:hi] = v.squeeze(-1)
        return actions, logprob, value

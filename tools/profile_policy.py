"""Per-section device time of the takeru policy on one chunk of real records (development aid)."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from nmmo_b200.vecenv import B200VecEnv
from nmmo_b200.takeru_policy import TakeruPolicy, HEADS
from nmmo_b200.emulation import unpack_batched_obs
E = 64
pool = B200VecEnv(num_envs=E, agent="takeru", collect_infos=False, curriculum="heldout")
pol = TakeruPolicy(pool.driver_env.unflatten_context, agents_per_env=128, envs_per_chunk=E).cuda().eval()
pool.async_reset(1)
o = pool.recv()[0]
for _ in range(8):
    a, _, _ = pol(o); pool.send(a); o = pool.recv()[0]
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    ob = unpack_batched_obs(o, pol.ctx)
    print("rows", o.shape[0])
    print("tile      ", t(lambda: pol.tile_encoder(ob["Tile"])))
    print("player sp ", t(lambda: pol.player_encoder.forward_sparse(ob["Entity"], ob["AgentId"][:, 0])))
    print("inv items ", t(lambda: pol.item_encoder(ob["Inventory"])))
    print("task      ", t(lambda: pol.task_encoder.fc(ob["Task"].float())))
    print("logits all", t(lambda: pol.logits(o)))
    print("forward   ", t(lambda: pol(o)))
    for tf in (True,):
        torch.backends.cuda.matmul.allow_tf32 = tf; torch.backends.cudnn.allow_tf32 = tf
        print("tf32 tile ", t(lambda: pol.tile_encoder(ob["Tile"])), "forward", t(lambda: pol(o)))

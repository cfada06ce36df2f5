"""Throughput of the rollout append and of sort + GAE on the device, next to the CPU restatement.

  python tools/bench_rollout.py [--batch 131072] [--envs 4096]
Prints one JSON line.  The append is an HBM -> HBM stream of the selected records (roofline: read + write of
rows x stride bytes); sort + GAE is latency-bound (a few small kernels).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nmmo_b200.rollout import DeviceRollout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=131072)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--agents", type=int, default=128)
    ap.add_argument("--stride", type=int, default=25344)
    ap.add_argument("--alive", type=float, default=0.5)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu-batch", type=int, default=32768)
    a = ap.parse_args()
    n = a.envs * a.agents
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(1)
    obs = torch.randint(0, 256, (n, a.stride), dtype=torch.uint8, device=dev, generator=g)
    actions = torch.randint(0, 100, (n, 12), dtype=torch.int32, device=dev, generator=g)
    val = torch.randn(n, device=dev, generator=g); lp = -torch.rand(n, device=dev, generator=g)
    rew = torch.randn(n, device=dev, generator=g); done = (torch.rand(n, device=dev, generator=g) < 0.02).float()
    mask = (torch.rand(n, device=dev, generator=g) < a.alive).to(torch.uint8)
    r = DeviceRollout(a.batch, n, a.stride)
    # ---- append: one recv() fills (part of) the batch
    per_call = min(int(mask.sum().item()), a.batch + 1)
    for _ in range(2):
        r.reset(); r.store(obs, val, actions, lp, rew, done, mask, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(a.reps):
        r.reset()
        e0.record(); r.store(obs, val, actions, lp, rew, done, mask, 1); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    store_ms = ms / a.reps
    store_bytes = per_call * (2 * a.stride + 2 * (12 * 4 + 4 * 4)) + n * 2
    # ---- sort + GAE over a full batch collected over several steps
    r.reset()
    step = 0
    while r.ptr < a.batch + 1 and step < 64:
        step += 1
        m = (torch.rand(n, device=dev, generator=g) < 0.2).to(torch.uint8)
        r.store(obs, val, actions, lp, rew, (torch.rand(n, device=dev, generator=g) < 0.02).float(), m, step)
    rows = r.ptr
    r.gae(0.99, 0.95)
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(a.reps):
        e0.record(); r.gae(0.99, 0.95); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    gae_ms = ms / a.reps
    # ---- CPU restatement of the same loop (Python, like the reference) on a bounded sample
    from oracle.rollout_oracle import RolloutOracle
    cb = min(a.cpu_batch, rows - 1)
    o = RolloutOracle(cb, 16)
    o.rewards[:] = r.rewards[:cb + 1].cpu().numpy(); o.values[:] = r.values[:cb + 1].cpu().numpy(); o.dones[:] = r.dones[:cb + 1].cpu().numpy()
    slot = r.slot[:cb + 1].cpu().numpy(); st = r.step[:cb + 1].cpu().numpy()
    o.sort_keys = list(zip(slot.tolist(), st.tolist())); o.ptr = cb + 1
    t0 = time.perf_counter(); o.gae(0.99, 0.95); cpu_s = time.perf_counter() - t0
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    print(json.dumps({"what": "rollout storage + GAE on the device (clean_pufferl.py:329-348, :413-436)",
                      "n_slots": n, "obs_stride": a.stride, "batch_size": a.batch,
                      "append": {"rows_per_call": per_call, "ms": store_ms, "GB/s": store_bytes / store_ms / 1e6, "peak_GB/s": peak,
                                 "frac": store_bytes / store_ms / 1e6 / peak, "alg_bytes": store_bytes},
                      "sort_gae": {"rows": rows, "ms": gae_ms, "samples_per_s": (rows - 1) / gae_ms * 1e3},
                      "cpu_port": {"rows": cb + 1, "seconds": cpu_s, "samples_per_s": cb / cpu_s, "kind": "port (Python loop, like the reference)"}}))
    r.close()


if __name__ == "__main__":
    main()

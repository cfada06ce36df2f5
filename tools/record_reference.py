#!/usr/bin/env python
"""Record one episode of the REAL reference engine for engine pinning (VERDICT r1 "next" #3).

Builds the environment exactly as /root/reference/reinforcement_learning/environment.py:52-76 does
(``nmmo.Env(Config(env_args))`` wrapped by the agent's ``RewardWrapper``), replaces the env's numpy ``Generator`` by
nmmo_b200.draw_tape.RecordingGenerator, steps it with seeded uniform-random valid actions and writes

    tests/golden/engine/<name>.npz        per tick: observations (every key of the obs dict), rewards, terminated,
                                          truncated, the realm's datastore tables (entity / item / tile), actions taken
    tests/golden/engine/<name>.tape.json  the draw log (tick, method, args, result, call site, context) and its
                                          translation into (tick, site, idx, k) keys + 32-bit values for
                                          nmmo_inject_rng, plus every call site the site rules could not map

which tests/test_engine_pinning.py replays through the oracle and the CUDA path.  The engine (`nmmo>=2.1,<2.2`,
pyproject.toml:19) is looked for in baseline/_ref (the driver's reference install) and on sys.path; when it is not
importable the script says so and exits 0 -- nothing else can be done without it.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def probe_engine():
    """-> (nmmo module or None, reason)."""
    for extra in (ROOT / "baseline" / "_ref", Path("/root/reference")):
        if extra.exists() and str(extra) not in sys.path:
            sys.path.insert(0, str(extra))
    try:
        import nmmo  # noqa: F401
        return nmmo, f"nmmo {getattr(nmmo, '__version__', '?')} from {Path(nmmo.__file__).parent}"
    except Exception as e:  # noqa: BLE001
        return None, f"engine unavailable: import nmmo failed ({type(e).__name__}: {e}); looked in baseline/_ref and sys.path"


def engine_context(chain):
    """Locals of the engine frames that the site rules need (nmmo_b200/draw_tape.py ENGINE_SITE_RULES)."""
    ctx = {}
    for f in chain:
        loc = f.f_locals
        ent = loc.get("entity") or loc.get("ent") or loc.get("self")
        for name in ("ent_id", "id"):
            v = getattr(ent, name, None)
            v = getattr(v, "val", v)
            if isinstance(v, (int, np.integer)) and "ent_id" not in ctx:
                ctx["ent_id"] = int(v)
        pos = getattr(ent, "pos", None)
        if isinstance(pos, tuple) and len(pos) == 2 and "row" not in ctx:
            ctx["row"], ctx["col"] = int(pos[0]), int(pos[1])
        for name in ("attempt", "_", "i"):
            if name in loc and isinstance(loc[name], (int, np.integer)) and f.f_code.co_name == "spawn":
                ctx.setdefault("attempt", int(loc[name]))
        if "agent_id" in loc and isinstance(loc["agent_id"], (int, np.integer)):
            ctx.setdefault("agent_id", int(loc["agent_id"]))
    return ctx


def find_generators(env):
    """Every attribute path (depth <= 3) of the env that holds a numpy Generator: they are all replaced by one proxy."""
    import numpy.random as npr
    found, seen = [], set()

    def walk(obj, path, depth):
        if id(obj) in seen or depth > 3:
            return
        seen.add(id(obj))
        for name in list(getattr(obj, "__dict__", {})):
            try:
                v = getattr(obj, name)
            except Exception:  # noqa: BLE001
                continue
            if isinstance(v, npr.Generator):
                found.append((obj, name, f"{path}.{name}"))
            elif hasattr(v, "__dict__") and not isinstance(v, type):
                walk(v, f"{path}.{name}", depth + 1)
    walk(env, "env", 0)
    return found


def dump_tables(realm):
    """Best-effort copy of the datastore tables (nmmo/datastore: EntityState / ItemState / TileState)."""
    out = {}
    ds = getattr(realm, "datastore", None)
    for name in ("Entity", "Item", "Tile"):
        try:
            tab = ds.table(name) if hasattr(ds, "table") else ds._tables[name]
            data = getattr(tab, "_data", None)
            if data is None and hasattr(tab, "get"):
                data = tab.get(list(range(tab._id_allocator.max_id)))
            out[name] = np.array(data)
        except Exception:  # noqa: BLE001
            pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agent", default="takeru")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--ticks", type=int, default=128)
    ap.add_argument("--name", default=None)
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "engine"))
    args = ap.parse_args()
    nmmo, why = probe_engine()
    if nmmo is None:
        print(why)
        return 0
    print(why)
    # ---- the env, exactly as the reference builds it (environment.py:52-76) minus the pufferlib emulation layer ----
    import importlib
    from types import SimpleNamespace
    import yaml
    from nmmo_b200.draw_tape import ENGINE_SITE_RULES, RecordingGenerator, translate
    environment = importlib.import_module("reinforcement_learning.environment")
    ref_cfg = yaml.safe_load(open("/root/reference/config.yaml")) if Path("/root/reference/config.yaml").exists() else {}
    env_args = dict(ref_cfg.get("env", {})); env_args.update(ref_cfg.get(args.agent, {}).get("env", {}))
    wrap_args = dict(ref_cfg.get("reward_wrapper", {})); wrap_args.update(ref_cfg.get(args.agent, {}).get("reward_wrapper", {}))
    agent_module = importlib.import_module(f"agent_zoo.{args.agent}")
    env = nmmo.Env(environment.Config(SimpleNamespace(**env_args)))
    env = agent_module.RewardWrapper(env, **wrap_args)
    raw = env.env if hasattr(env, "env") else env
    marker = "/" + Path(nmmo.__file__).parent.name + "/"
    obs, _ = env.reset(seed=args.seed)
    proxy = None
    for owner, name, path in find_generators(raw):
        if proxy is None:
            proxy = RecordingGenerator(getattr(owner, name), marker, lambda: int(raw.realm.tick), engine_context)
        setattr(owner, name, proxy)
        print("recording", path)
    if proxy is None:
        print("engine unavailable: no numpy Generator found on the env (API differs from nmmo 2.1)")
        return 0
    # NOTE: draws made inside reset(seed) itself (map choice, spawn order, tasks) happen before the proxy is in place;
    # a second reset with the proxy installed records them on the same seed
    obs, _ = env.reset(seed=args.seed)
    rng = np.random.default_rng(args.seed)
    frames = []
    for t in range(args.ticks):
        actions = {}
        for agent_id, ob in obs.items():
            act = {}
            for atn, sub in ob["ActionTargets"].items():
                act[atn] = {}
                for arg, mask in sub.items():
                    valid = np.flatnonzero(np.asarray(mask))
                    act[atn][arg] = int(rng.choice(valid)) if len(valid) else 0
            actions[agent_id] = act
        obs, rew, term, trunc, info = env.step(actions)
        frame = {"tick": int(raw.realm.tick), "agents": np.array(sorted(obs)), "rewards": np.array([rew[a] for a in sorted(rew)]),
                 "terminated": np.array([term[a] for a in sorted(term)]), "truncated": np.array([trunc[a] for a in sorted(trunc)])}
        for a in sorted(obs):
            for key, val in obs[a].items():
                if key != "ActionTargets":
                    frame[f"obs/{a}/{key}"] = np.asarray(val)
                else:
                    for atn, sub in val.items():
                        for arg, mask in sub.items():
                            frame[f"obs/{a}/ActionTargets/{atn}/{arg}"] = np.asarray(mask)
        for k, v in dump_tables(raw.realm).items():
            frame[f"table/{k}"] = v
        frame["actions"] = np.array([[actions[a][x][y] for x in sorted(actions[a]) for y in sorted(actions[a][x])] for a in sorted(actions)])
        frames.append(frame)
        if not env.agents:
            break
    state = {"S": int(raw.config.MAP_SIZE), "id_to_row": {}}
    keys, vals, unmapped = translate(proxy.log, ENGINE_SITE_RULES, state)
    out = Path(args.out); out.mkdir(parents=True, exist_ok=True)
    name = args.name or f"{args.agent}_seed{args.seed}"
    np.savez_compressed(out / f"{name}.npz", **{f"t{i}/{k}": v for i, f in enumerate(frames) for k, v in f.items()})
    (out / f"{name}.tape.json").write_text(json.dumps({
        "engine": why, "agent": args.agent, "seed": args.seed, "ticks": len(frames),
        "draws": [{"tick": d.tick, "method": d.method, "args": [str(a) for a in d.args], "result": d.result if np.isscalar(d.result) else list(np.ravel(d.result).tolist()),
                   "site": f"{d.module}:{d.function}:{d.lineno}", "context": d.context} for d in proxy.log],
        "keys": [int(k) for k in keys], "values": [int(v) for v in vals],
        "unmapped_call_sites": sorted({f"{m}:{f}:{l} ({meth})" for m, f, l, meth in unmapped})}, indent=0))
    print(f"wrote {out / name}.npz and .tape.json: {len(frames)} ticks, {len(proxy.log)} draws, {len(keys)} keyed, "
          f"{len(unmapped)} from unmapped call sites")
    return 0


if __name__ == "__main__":
    sys.exit(main())

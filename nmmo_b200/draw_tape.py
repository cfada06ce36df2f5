"""Draw tape: record the reference engine's numpy RNG draws and turn them into keyed injections.

north_star: "the reference's numpy RNG draws recorded and injected so integer state matches exactly".  The engine
(`nmmo>=2.1,<2.2`, /root/reference/pyproject.toml:19) draws from one sequential ``numpy.random.Generator`` per env
(seeded at /root/reference/reinforcement_learning/stat_wrapper.py:27,51 via ``env.seed`` / ``env.reset(seed=)``);
this simulator addresses every draw by ``(tick, site, idx, k)`` (include/nmmo_spec.h ``nm_site`` / ``nm_rng_key``) so
that a parallel device step and a sequential CPU step consume identical values in any order.  The bridge has three
parts, all host-side Python and all testable without the engine:

1. ``RecordingGenerator`` -- a transparent proxy around a ``numpy.random.Generator`` that logs every call
   (method, arguments, result) together with its *call site*: the innermost stack frame that belongs to the engine
   package, and the locals the site rules below need (entity id, tile position, attempt number ...).
2. value translation -- a recorded result becomes the 32-bit value that makes this repo's draw arithmetic
   (``bounded(u, n) = (u * n) >> 32``, ``u < threshold``, Fisher-Yates from the top) reproduce it.
3. ``translate`` -- site rules map a logged call to ``(site, idx, k)``; the output is the ``(keys, values)`` pair
   ``nmmo_inject_rng`` / ``OracleEnv.inject_rng`` take.

The site rules for the real engine (``ENGINE_SITE_RULES``) name upstream modules/functions from the published nmmo 2.1
source layout; they are hypotheses until the engine is importable (tools/record_reference.py reports every call site it
could not map, so the table can be completed the day it is).
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .config import SPEC

SITE = {k[3:]: v for k, v in SPEC.items() if k.startswith("RS_") and k != "RS_N"}


def rng_key(tick: int, site: int, idx: int, k: int) -> int:
    """Python mirror of nm_rng_key (include/nmmo_spec.h): tick 20b | site 4b | idx 24b | k 8b."""
    return ((tick & 0xFFFFF) << 36) | ((site & 0xF) << 32) | ((idx & 0xFFFFFF) << 8) | (k & 0xFF)


# ------------------------------------------------------------------------------------ value translation ----
def u32_for_bounded(result: int, n: int) -> int:
    """Smallest u with (u * n) >> 32 == result: makes ``bounded(u, n)`` return a recorded ``integers(0, n)``."""
    if not 0 <= result < n:
        raise ValueError(f"recorded value {result} outside [0, {n})")
    u = -((-result << 32) // n)          # ceil(result * 2^32 / n)
    assert (u * n) >> 32 == result and u < (1 << 32)
    return u


def u32_for_uniform(x: float) -> int:
    """A recorded ``random()`` in [0, 1) as a 32-bit value: ``u < thr(p)`` iff ``x < p`` for every p that is a
    multiple of 2^-32 (the thresholds of nmmo_b200/config.py ``_thr``)."""
    if not 0.0 <= x < 1.0:
        raise ValueError(f"recorded uniform {x} outside [0, 1)")
    return min(int(x * 4294967296.0), 0xFFFFFFFF)


def fisher_yates_draws(perm: Sequence[int]) -> List[Tuple[int, int]]:
    """[(i, u32)] for i = n-1 .. 1 such that the repo's shuffle (``for i = n-1..1: j = bounded(u_i, i + 1);
    swap(a[i], a[j])`` applied to the identity) ends in ``perm`` -- the inverse of a recorded ``shuffle`` /
    ``permutation`` result."""
    n = len(perm)
    if sorted(perm) != list(range(n)):
        raise ValueError("not a permutation of 0..n-1")
    work = list(range(n))
    pos = {v: i for i, v in enumerate(work)}
    out = []
    for i in range(n - 1, 0, -1):
        j = pos[perm[i]]                 # the element that must end at i sits at j <= i now
        assert j <= i
        out.append((i, u32_for_bounded(j, i + 1)))
        vi, vj = work[i], work[j]
        work[i], work[j] = vj, vi
        pos[vi], pos[vj] = j, i
    return out


# ------------------------------------------------------------------------------------------- recording ----
@dataclass
class Draw:
    tick: int
    method: str
    args: tuple
    kwargs: dict
    result: Any
    module: str            # engine module of the call site, e.g. "nmmo/entity/npc.py"
    function: str          # qualified function name of the call site
    lineno: int
    context: Dict[str, Any] = field(default_factory=dict)      # locals the site rules asked for


class RecordingGenerator:
    """Proxy of a numpy Generator: same results, every call logged with its engine call site.

    ``package_marker`` selects which stack frames count as engine code (a substring of the file path, e.g.
    "/nmmo/"); ``tick_fn`` returns the current tick (``lambda: env.realm.tick``); ``context_fn(frame_chain)`` may
    extract locals (entity ids, positions) for the site rules."""

    def __init__(self, gen, package_marker: str, tick_fn: Callable[[], int],
                 context_fn: Optional[Callable[[list], Dict[str, Any]]] = None):
        object.__setattr__(self, "_gen", gen)
        object.__setattr__(self, "_marker", package_marker)
        object.__setattr__(self, "_tick_fn", tick_fn)
        object.__setattr__(self, "_context_fn", context_fn)
        object.__setattr__(self, "log", [])

    def _site(self):
        f = sys._getframe(2)
        chain = []
        while f is not None:
            fn = f.f_code.co_filename.replace("\\", "/")
            if self._marker in fn:
                chain.append(f)
            f = f.f_back
        if not chain:
            return "<outside engine>", "?", 0, []
        top = chain[0]
        mod = top.f_code.co_filename.replace("\\", "/")
        mod = mod[mod.index(self._marker) + 1:] if self._marker in mod else mod
        qual = getattr(top.f_code, "co_qualname", top.f_code.co_name)
        return mod, qual, top.f_lineno, chain

    def __getattr__(self, name):
        attr = getattr(self._gen, name)
        if not callable(attr):
            return attr

        def call(*args, **kwargs):
            res = attr(*args, **kwargs)
            mod, qual, line, chain = self._site()
            ctx = self._context_fn(chain) if (self._context_fn and chain) else {}
            if name == "shuffle" and args:                  # in-place: the result is the shuffled argument
                logged = list(args[0])
            else:
                logged = res.tolist() if isinstance(res, np.ndarray) else res
            self.log.append(Draw(int(self._tick_fn()), name, tuple(a if np.isscalar(a) else repr(type(a)) for a in args),
                                 dict(kwargs), logged, mod, qual, int(line), ctx))
            return res
        return call


# ------------------------------------------------------------------------------------------ translation ----
@dataclass
class SiteRule:
    """How one engine call site maps to keyed draws.

    module / function: suffix / name matched against the logged call site.
    site: nm_site name ("NPC_SPAWN", ...).
    idx: callable(draw, state) -> idx of the key (entity row, tile index, attempt ...).
    ordinal: "per_idx" numbers the draws of one (tick, site, idx) 0, 1, 2 ... in call order (k);
             "fixed:<k>" uses a constant k.
    kind: "bounded" (integers(lo, hi) / choice over n), "uniform" (random()), "shuffle" (Fisher-Yates inverse:
          one key per position i, idx = i)."""
    module: str
    function: str
    site: str
    kind: str
    idx: Callable[[Draw, dict], int] = lambda d, st: 0
    ordinal: str = "per_idx"


def _n_of(d: Draw) -> Tuple[int, int]:
    """(lo, n) of an integers / choice call."""
    if d.method == "integers":
        a = [x for x in d.args if isinstance(x, (int, np.integer))]
        lo, hi = (0, a[0]) if len(a) == 1 else (a[0], a[1])
        if d.kwargs.get("endpoint"):
            hi += 1
        return int(lo), int(hi - lo)
    if d.method == "choice":
        n = d.kwargs.get("_n") or (d.args[0] if d.args and isinstance(d.args[0], (int, np.integer)) else None)
        if n is None:
            raise ValueError("choice over a sequence: the recorder must log its length as kwargs['_n'] and the index as result")
        return 0, int(n)
    raise ValueError(f"not a bounded draw: {d.method}")


def translate(log: Sequence[Draw], rules: Sequence[SiteRule], state: Optional[dict] = None):
    """-> (keys uint64[n], values uint32[n], unmapped list of (module, function, lineno, method)).

    `state` is handed to the idx callables (e.g. {"id_to_row": {...}, "S": 160})."""
    state = state or {}
    keys: List[int] = []
    vals: List[int] = []
    unmapped = []
    counters: Dict[Tuple[int, int, int], int] = {}
    for d in log:
        rule = next((r for r in rules if d.module.endswith(r.module) and d.function.split(".")[-1] == r.function.split(".")[-1]), None)
        if rule is None:
            unmapped.append((d.module, d.function, d.lineno, d.method))
            continue
        site = SITE[rule.site]
        if rule.kind == "shuffle":
            perm = list(d.result)
            if sorted(perm) != list(range(len(perm))):          # a shuffled list of objects: rank them by first appearance order
                raise ValueError("shuffle of non-index values: log the permutation of positions instead")
            for i, u in fisher_yates_draws(perm):
                keys.append(rng_key(d.tick, site, i, 0)); vals.append(u)
            continue
        idx = int(rule.idx(d, state))
        if rule.ordinal.startswith("fixed:"):
            k = int(rule.ordinal.split(":")[1])
        else:
            k = counters.get((d.tick, site, idx), 0)
            counters[(d.tick, site, idx)] = k + 1
        if rule.kind == "bounded":
            lo, n = _n_of(d)
            u = u32_for_bounded(int(d.result) - lo, n)
        elif rule.kind == "uniform":
            u = u32_for_uniform(float(d.result))
        else:
            raise ValueError(rule.kind)
        keys.append(rng_key(d.tick, site, idx, k)); vals.append(u)
    return np.asarray(keys, np.uint64), np.asarray(vals, np.uint32), unmapped


# Site rules for upstream nmmo 2.1 [UPSTREAM, hypotheses: module/function names from the published source layout;
# tools/record_reference.py lists every call site these do not cover].  idx extractors read the context the recorder
# captured (see tools/record_reference.py::engine_context).
def _row(d: Draw, st: dict) -> int:
    return int(st.get("id_to_row", {}).get(d.context.get("ent_id"), 0))


def _tile(d: Draw, st: dict) -> int:
    return int(d.context.get("row", 0)) * int(st.get("S", 0)) + int(d.context.get("col", 0))


ENGINE_SITE_RULES = [
    SiteRule("nmmo/core/env.py", "_load_map_file", "MAP", "bounded", ordinal="fixed:0"),
    SiteRule("nmmo/core/map.py", "reset", "MAP", "bounded", ordinal="fixed:0"),
    SiteRule("nmmo/lib/spawn.py", "get_spawn_locs", "SPAWN_PERM", "shuffle"),
    SiteRule("nmmo/entity/entity_manager.py", "PlayerManager.spawn", "RESILIENT", "shuffle"),
    SiteRule("nmmo/task/task_spec.py", "make_task_from_spec", "TASK", "bounded", idx=lambda d, st: int(d.context.get("agent_id", 1)) - 1, ordinal="fixed:0"),
    SiteRule("nmmo/systems/ai/move.py", "habitable", "NPC_DECIDE", "bounded", idx=_row),
    SiteRule("nmmo/systems/ai/move.py", "meander", "NPC_DECIDE", "bounded", idx=_row),
    SiteRule("nmmo/systems/ai/policy.py", "hostile", "NPC_DECIDE", "bounded", idx=_row),
    SiteRule("nmmo/core/realm.py", "step", "BUY_SHUFFLE", "shuffle"),
    SiteRule("nmmo/systems/skill.py", "process_drops", "HARVEST", "uniform", idx=_row, ordinal="per_idx"),
    SiteRule("nmmo/entity/entity_manager.py", "NPCManager.spawn", "NPC_SPAWN", "bounded", idx=lambda d, st: int(d.context.get("attempt", 0))),
    SiteRule("nmmo/entity/npc.py", "NPC.spawn", "NPC_SPAWN", "bounded", idx=lambda d, st: int(d.context.get("attempt", 0))),
    SiteRule("nmmo/core/tile.py", "Tile.step", "RESPAWN", "uniform", idx=_tile, ordinal="fixed:0"),
]

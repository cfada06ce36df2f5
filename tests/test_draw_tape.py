"""Draw tape (nmmo_b200/draw_tape.py): recording proxy, value translation and key translation, on CPU.

The real engine is not importable here, so the recorder is exercised on a stand-in "engine" package
written to a temp directory, and the translated keys are proven end to end against the oracle: a recorded map choice
and a recorded spawn permutation, injected by key, must come out of the oracle's reset exactly as recorded.
"""
import importlib
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

from nmmo_b200.config import SPEC
from nmmo_b200.draw_tape import (SITE, Draw, RecordingGenerator, SiteRule, fisher_yates_draws, rng_key, translate,
                                 u32_for_bounded, u32_for_uniform)
from util import SMALL, build_world

ROOT = Path(__file__).resolve().parent.parent


def bounded(u, n):
    return (int(u) * int(n)) >> 32


def test_rng_key_matches_the_c_header():
    # nm_rng_key: tick 20b | site 4b | idx 24b | k 8b
    assert rng_key(1, 2, 3, 4) == (1 << 36) | (2 << 32) | (3 << 8) | 4
    assert rng_key(0xFFFFF, 0xF, 0xFFFFFF, 0xFF) == (1 << 56) - 1


def test_value_translation_inverts_the_draw_arithmetic():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 4, 5, 25, 99, 128, 160 * 160, 1 << 20):
        for r in set([0, n - 1] + rng.integers(0, n, 20).tolist()):
            u = u32_for_bounded(r, n)
            assert bounded(u, n) == r and (u == 0 or bounded(u - 1, n) == r - 1)
    for p in (0.025, 0.1, 0.105, 0.02):
        thr = min(int(p * 4294967296.0), 0xFFFFFFFF)
        for x in np.concatenate([rng.random(200), [0.0, p - 1e-12, p, np.nextafter(p, 1), 0.999999999]]):
            assert (u32_for_uniform(float(x)) < thr) == (x < thr / 4294967296.0)
    with pytest.raises(ValueError):
        u32_for_bounded(5, 5)


def test_fisher_yates_inverse():
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 16, 128):
        perm = rng.permutation(n).tolist()
        a = list(range(n))
        for i, u in fisher_yates_draws(perm):
            j = bounded(u, i + 1)
            a[i], a[j] = a[j], a[i]
        assert a == perm


def _fake_engine(tmp_path):
    pkg = tmp_path / "fakeeng"
    (pkg / "entity").mkdir(parents=True)
    (pkg / "__init__.py").write_text("")
    (pkg / "entity" / "__init__.py").write_text("")
    (pkg / "entity" / "npc.py").write_text(textwrap.dedent('''
        class NPC:
            def __init__(self, ent_id, pos): self.ent_id, self.pos = ent_id, pos
            def wander(self, gen):
                return int(gen.integers(0, 4))
        def spawn(gen, attempt):
            r = int(gen.integers(16, 48)); c = int(gen.integers(16, 48))
            return r, c
    '''))
    (pkg / "tile.py").write_text(textwrap.dedent('''
        class Tile:
            def __init__(self, r, c): self.pos = (r, c)
            def step(self, gen): return gen.random() < 0.1
    '''))
    (pkg / "world.py").write_text(textwrap.dedent('''
        def order(gen, n):
            lst = list(range(n)); gen.shuffle(lst); return lst
    '''))
    sys.path.insert(0, str(tmp_path))
    for m in [k for k in sys.modules if k.startswith("fakeeng")]:
        del sys.modules[m]
    return importlib.import_module("fakeeng.entity.npc"), importlib.import_module("fakeeng.tile"), importlib.import_module("fakeeng.world")


def test_recording_proxy_and_key_translation(tmp_path):
    npc, tile, world = _fake_engine(tmp_path)
    tick = [0]

    def ctx(chain):
        out = {}
        for f in chain:
            s = f.f_locals.get("self")
            if hasattr(s, "ent_id"): out["ent_id"] = s.ent_id
            if hasattr(s, "pos"): out["row"], out["col"] = s.pos
            if "attempt" in f.f_locals: out["attempt"] = f.f_locals["attempt"]
        return out

    ref = np.random.default_rng(5)
    gen = RecordingGenerator(np.random.default_rng(5), "/fakeeng/", lambda: tick[0], ctx)
    # the proxy is transparent: same stream as an unwrapped generator
    a = npc.NPC(-7, (20, 21)).wander(gen)
    assert a == int(ref.integers(0, 4))
    tick[0] = 3
    rc0 = npc.spawn(gen, attempt=0); rc1 = npc.spawn(gen, attempt=1)
    assert rc0 == (int(ref.integers(16, 48)), int(ref.integers(16, 48)))
    stays = tile.Tile(30, 31).step(gen)
    perm = world.order(gen, 6)
    outside = gen.integers(0, 10)                       # a draw from outside the engine package
    log = gen.log
    assert [d.module for d in log[:6]] == ["fakeeng/entity/npc.py"] * 5 + ["fakeeng/tile.py"]
    assert log[0].function.endswith("wander") and log[0].context["ent_id"] == -7 and log[0].tick == 0
    assert log[1].context["attempt"] == 0 and log[3].context["attempt"] == 1 and log[1].tick == 3
    assert log[-1].module == "<outside engine>"
    rules = [SiteRule("entity/npc.py", "NPC.wander", "NPC_DECIDE", "bounded", idx=lambda d, st: st["id_to_row"][d.context["ent_id"]]),
             SiteRule("entity/npc.py", "spawn", "NPC_SPAWN", "bounded", idx=lambda d, st: d.context["attempt"]),
             SiteRule("tile.py", "Tile.step", "RESPAWN", "uniform", idx=lambda d, st: d.context["row"] * st["S"] + d.context["col"], ordinal="fixed:0"),
             SiteRule("world.py", "order", "BUY_SHUFFLE", "shuffle")]
    keys, vals, unmapped = translate(log, rules, {"id_to_row": {-7: 40}, "S": 64})
    assert unmapped == [("<outside engine>", "?", 0, "integers")]
    kv = dict(zip(keys.tolist(), vals.tolist()))
    assert bounded(kv[rng_key(0, SITE["NPC_DECIDE"], 40, 0)], 4) == a
    # the two draws of one spawn attempt are ordinals 0 and 1 of that attempt
    assert 16 + bounded(kv[rng_key(3, SITE["NPC_SPAWN"], 0, 0)], 32) == rc0[0]
    assert 16 + bounded(kv[rng_key(3, SITE["NPC_SPAWN"], 0, 1)], 32) == rc0[1]
    assert 16 + bounded(kv[rng_key(3, SITE["NPC_SPAWN"], 1, 1)], 32) == rc1[1]
    thr = int(0.1 * 4294967296.0)
    assert (kv[rng_key(3, SITE["RESPAWN"], 30 * 64 + 31, 0)] < thr) == bool(stays)
    lst = list(range(6))
    for i in range(5, 0, -1):
        j = bounded(kv[rng_key(3, SITE["BUY_SHUFFLE"], i, 0)], i + 1)
        lst[i], lst[j] = lst[j], lst[i]
    assert lst == perm


def test_translated_tape_drives_the_oracle():
    """End to end on CPU: a "recorded" map choice and spawn permutation, translated to keys and injected, come out of
    the oracle's reset exactly as recorded (nm_site RS_MAP, RS_SPAWN_PERM)."""
    from oracle.oracle import OracleEnv
    world = build_world(task_dim=64, n_maps=4, **SMALL, NC_HORIZON=32)
    cfg = world[0]
    P, ce, b = int(cfg[SPEC["NC_N_PLAYERS"]]), int(cfg[SPEC["NC_MAP_CENTER"]]), int(cfg[SPEC["NC_MAP_BORDER"]])
    rng = np.random.default_rng(3)
    perm = rng.permutation(P).tolist()
    log = [Draw(0, "integers", (0, 4), {}, 2, "eng/core/env.py", "Env._load_map_file", 10),
           Draw(0, "shuffle", ("<list>",), {}, perm, "eng/lib/spawn.py", "get_spawn_locs", 20)]
    rules = [SiteRule("core/env.py", "_load_map_file", "MAP", "bounded", ordinal="fixed:0"),
             SiteRule("lib/spawn.py", "get_spawn_locs", "SPAWN_PERM", "shuffle")]
    keys, vals, unmapped = translate(log, rules)
    assert not unmapped and len(keys) == 1 + (P - 1)
    o = OracleEnv(*world)
    o.inject_rng(keys, vals)
    o.reset(123)
    from oracle.oracle import lib
    assert lib().oracle_map_id(o.h) == 2
    ent = o.snapshot()[0]
    for p in range(P):
        k = perm[p] * (4 * ce) // P
        side, off = divmod(k, ce)
        want = [(b, b + off), (b + off, b + ce), (b + ce, b + ce - off), (b + ce - off, b)][side]
        assert (int(ent[p, SPEC["EA_ROW"]]), int(ent[p, SPEC["EA_COL"]])) == want
    # without the tape the same seed spawns elsewhere
    o2 = OracleEnv(*world); o2.reset(123)
    assert not np.array_equal(o2.snapshot()[0][:, 2:4], ent[:, 2:4])


def test_recorder_script_reports_missing_engine():
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "record_reference.py")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    assert "engine unavailable" in r.stdout or "wrote " in r.stdout

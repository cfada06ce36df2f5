"""ctypes binding of oracle/nmmo_oracle.c (test infrastructure; parity unpinned, see its header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "_build" / "liboracle.so"


def build(force: bool = False) -> Path:
    src = _DIR / "nmmo_oracle.c"
    hdr = _DIR.parent / "include" / "nmmo_spec.h"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.check_call(["make", "-C", str(_DIR), "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not _LIB.exists():
                raise
        L = C.CDLL(str(_LIB))
        vp, i32p, u8p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_inject_rng.argtypes = [vp, vp, vp, C.c_int]
        L.oracle_reset.argtypes = [vp, C.c_uint64, C.c_int, vp]
        L.oracle_step.argtypes = [vp, vp]
        L.oracle_sample_actions.argtypes = [vp, C.c_uint64, vp]
        for name, rt in (("oracle_obs", u8p), ("oracle_rewards", f32p), ("oracle_terminated", u8p),
                         ("oracle_truncated", u8p), ("oracle_mask", u8p), ("oracle_info", f32p),
                         ("oracle_info_valid", u8p)):
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [vp]
        for name in ("oracle_episode_done", "oracle_tick", "oracle_map_id", "oracle_obs_stride"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [vp]
        L.oracle_snapshot.argtypes = [vp, vp, vp, vp]
        L.oracle_env_rewards.restype = C.POINTER(C.c_double)
        L.oracle_env_rewards.argtypes = [vp]
        L.oracle_log_count.restype = C.c_int
        L.oracle_log_count.argtypes = [vp, C.c_int]
        L.oracle_log_rows.argtypes = [vp, C.c_int, vp]
        L.oracle_task_state.argtypes = [vp, vp, vp, vp, vp]
        L.oracle_step_many.argtypes = [vp, C.c_int, vp]
        L.oracle_sample_many.argtypes = [vp, C.c_int, C.c_uint64, vp]
        L.oracle_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleEnv:
    """One sequential CPU environment."""

    def __init__(self, cfg, fcfg, maps, task_table, task_embed):
        from nmmo_b200.config import SPEC, ObsLayout
        self.S = SPEC
        self.cfg = np.ascontiguousarray(cfg, np.int32)
        self.fcfg = np.ascontiguousarray(fcfg, np.float64)
        self.maps = np.ascontiguousarray(maps, np.uint8)
        self.task_table = np.ascontiguousarray(task_table, np.int32)
        self.task_embed = np.ascontiguousarray(task_embed, np.uint16)
        self.L = ObsLayout(self.cfg)
        self.P = int(cfg[SPEC["NC_N_PLAYERS"]]); self.N = int(cfg[SPEC["NC_N_NPCS"]])
        self.cap = int(cfg[SPEC["NC_ITEM_CAP"]]); self.Sz = int(cfg[SPEC["NC_MAP_SIZE"]])
        self.h = lib().oracle_create(_ptr(self.cfg), _ptr(self.fcfg), _ptr(self.maps), self.maps.shape[0],
                                     _ptr(self.task_table), _ptr(self.task_embed), self.task_table.shape[0])
        assert lib().oracle_obs_stride(self.h) == self.L.stride

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def inject_rng(self, keys, vals):
        order = np.argsort(keys)
        k = np.ascontiguousarray(np.asarray(keys, np.uint64)[order])
        v = np.ascontiguousarray(np.asarray(vals, np.uint32)[order])
        lib().oracle_inject_rng(self.h, _ptr(k), _ptr(v), len(k))

    def reset(self, seed, map_id=-1, task_ids=None):
        t = None if task_ids is None else np.ascontiguousarray(task_ids, np.int32)
        lib().oracle_reset(self.h, C.c_uint64(int(seed)), int(map_id), _ptr(t))

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (self.P, 12)
        lib().oracle_step(self.h, _ptr(a))

    def sample_actions(self, seed):
        out = np.zeros((self.P, 12), np.int32)
        lib().oracle_sample_actions(self.h, C.c_uint64(int(seed)), _ptr(out))
        return out

    def _arr(self, fn, shape, dtype):
        p = getattr(lib(), fn)(self.h)
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(p, shape=(n,)).view(dtype)[:n].reshape(shape).copy()

    @property
    def obs(self):
        return self._arr("oracle_obs", (self.P, self.L.stride), np.uint8)

    @property
    def rewards(self):
        return self._arr("oracle_rewards", (self.P,), np.float32)

    @property
    def terminated(self):
        return self._arr("oracle_terminated", (self.P,), np.uint8)

    @property
    def truncated(self):
        return self._arr("oracle_truncated", (self.P,), np.uint8)

    @property
    def mask(self):
        return self._arr("oracle_mask", (self.P,), np.uint8)

    @property
    def info(self):
        return self._arr("oracle_info", (self.P, self.S["IN_N"]), np.float32)

    @property
    def info_valid(self):
        return self._arr("oracle_info_valid", (self.P,), np.uint8)

    @property
    def episode_done(self):
        return bool(lib().oracle_episode_done(self.h))

    @property
    def tick(self):
        return lib().oracle_tick(self.h)

    @property
    def map_id(self):
        return lib().oracle_map_id(self.h)

    @property
    def env_rewards(self):
        return self._arr("oracle_env_rewards", (self.P,), np.float64)

    def log_rows(self, p):
        """Episode event log of player row p, in the reference's 9-column layout."""
        n = lib().oracle_log_count(self.h, int(p))
        out = np.zeros((n, 9), np.int32)
        if n:
            lib().oracle_log_rows(self.h, int(p), _ptr(out))
        return out

    def task_state(self):
        tid = np.zeros(self.P, np.int32); comp = np.zeros(self.P, np.int32); sig = np.zeros(self.P, np.int32)
        mp = np.zeros(self.P, np.float64)
        lib().oracle_task_state(self.h, _ptr(tid), _ptr(comp), _ptr(sig), _ptr(mp))
        return tid, comp, sig, mp

    def snapshot(self):
        ent = np.zeros((self.P + self.N, self.S["EA_N"]), np.int16)
        items = np.zeros((self.cap, self.S["IS_N"]), np.int16)
        mp = np.zeros((self.Sz, self.Sz), np.uint8)
        lib().oracle_snapshot(self.h, _ptr(ent), _ptr(items), _ptr(mp))
        return ent, items, mp


class OracleBatch:
    """Many sequential CPU envs stepped by an OpenMP loop (CPU baseline for bench.py)."""

    def __init__(self, cfg, fcfg, maps, task_table, task_embed, n_envs, threads=None):
        import os
        self.envs = [OracleEnv(cfg, fcfg, maps, task_table, task_embed) for _ in range(n_envs)]
        self.n = n_envs
        self.P = self.envs[0].P
        self.ptrs = (C.c_void_p * n_envs)(*[e.h for e in self.envs])
        self.actions = np.zeros((n_envs, self.P, 12), np.int32)
        self.threads = threads or len(os.sched_getaffinity(0))
        lib().oracle_set_threads(int(self.threads))      # not the environment: torchrun exports OMP_NUM_THREADS=1

    def reset(self, seeds):
        for e, s in zip(self.envs, seeds):
            e.reset(int(s))

    def sample(self, seed):
        lib().oracle_sample_many(self.ptrs, self.n, C.c_uint64(int(seed)), _ptr(self.actions))

    def step(self):
        lib().oracle_step_many(self.ptrs, self.n, _ptr(self.actions))

    def alive(self):
        return int(sum(int(e.mask.sum()) for e in self.envs))

"""Synthetic seeded terrain maps (host side, numpy).

The reference engine builds its maps once with ``nmmo/core/terrain.py`` (simplex noise from the
``vec_noise`` C extension, not available here) and stores them under ``PATH_MAPS``
(/root/reference/reinforcement_learning/environment.py:41); at reset an env only *loads* one
(``map_id`` drawn among ``MAP_N``).  Map generation is therefore not on the hot path: the
simulator consumes ``uint8[n_maps, S, S]`` material grids.  This module produces grids with the
same structure as upstream terrain -- thresholds water 0.30 / grass 0.70 / foliage 0.85 / stone,
a void border with a one-tile grass ring on which players spawn, profession resources scattered
on top ([UPSTREAM] nmmo/core/terrain.py, config.Terrain) -- from a multi-octave value noise that
is fully determined by (seed, map index).  BASELINE.json config 2 calls these "synthetic seeded
maps".
"""
from __future__ import annotations

import numpy as np

from .config import SPEC

VOID, WATER, GRASS, SCRUB, FOILAGE, STONE, SLAG, ORE, STUMP, TREE, FRAGMENT, CRYSTAL, WEEDS, HERB, OCEAN, FISH = range(16)


def _smooth_noise(rng: np.random.Generator, size: int, cell: int) -> np.ndarray:
    """Bicubic-ish smooth value noise: random lattice, smoothstep bilinear upsample."""
    n = size // cell + 3
    lat = rng.random((n, n))
    xs = np.arange(size) / cell
    i0 = np.floor(xs).astype(np.int64)
    t = xs - i0
    t = t * t * (3 - 2 * t)
    a = lat[i0][:, i0]
    b = lat[i0][:, i0 + 1]
    c = lat[i0 + 1][:, i0]
    d = lat[i0 + 1][:, i0 + 1]
    tr = t[:, None]
    tc = t[None, :]
    return (a * (1 - tc) + b * tc) * (1 - tr) + (c * (1 - tc) + d * tc) * tr


TERRAIN = {
    # (lattice cell, amplitude) octaves of the value noise
    "coarse": ((32, 1.0), (16, 0.5), (8, 0.25), (4, 0.125)),      # broad lakes and ranges (the default of the parity tests)
    "fine": ((16, 0.4), (8, 1.0), (4, 0.5)),                       # ponds and groves every ~10 tiles, like upstream's frequency
}


def generate_map(cfg: np.ndarray, seed: int, index: int, terrain: str = "coarse") -> np.ndarray:
    S = int(cfg[SPEC["NC_MAP_SIZE"]])
    ce = int(cfg[SPEC["NC_MAP_CENTER"]])
    rng = np.random.default_rng([int(seed) & 0xFFFFFFFF, int(index)])
    val = np.zeros((S, S))
    for cell, amp in TERRAIN[terrain]:
        val += amp * _smooth_noise(rng, S, cell)
    # rank-normalise so the material proportions follow the upstream thresholds exactly
    order = val.ravel().argsort().argsort().reshape(S, S) / float(S * S - 1)
    m = np.full((S, S), STONE, np.uint8)
    m[order <= 0.85] = FOILAGE
    m[order <= 0.70] = GRASS
    m[order <= 0.30] = WATER
    # profession resources
    hab = (m == GRASS) | (m == FOILAGE)
    near_hab = np.zeros_like(hab)
    near_hab[1:, :] |= hab[:-1, :]; near_hab[:-1, :] |= hab[1:, :]
    near_hab[:, 1:] |= hab[:, :-1]; near_hab[:, :-1] |= hab[:, 1:]
    u = rng.random((S, S))
    stone_edge = (m == STONE) & near_hab
    m[stone_edge & (u < 0.10)] = ORE
    m[stone_edge & (u >= 0.10) & (u < 0.20)] = CRYSTAL
    grass = m == GRASS
    m[grass & (u < 0.03)] = TREE
    m[grass & (u >= 0.03) & (u < 0.05)] = HERB
    water_edge = (m == WATER) & near_hab
    m[water_edge & (u < 0.08)] = FISH
    # void border, grass spawn ring, no walls right behind the ring
    half = S // 2
    rr, cc = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    l = np.maximum(np.abs(rr - half), np.abs(cc - half))
    ring = ce // 2
    inner = (l == ring - 1) & np.isin(m, (STONE, WATER, FISH, ORE, CRYSTAL))
    m[inner] = FOILAGE
    m[l == ring] = GRASS
    m[l > ring] = VOID
    kp = int(cfg[SPEC["NC_SPAWN_PATCH"]])
    if kp > 0:      # clustered spawn (BASELINE.json configs[4]): the spawn patch and a one-tile rim around it are walkable
        o = half - kp // 2
        patch = m[o - 1:o + kp + 1, o - 1:o + kp + 1]
        patch[np.isin(patch, (STONE, WATER, FISH, ORE, CRYSTAL))] = GRASS
    return m


def generate_maps(cfg: np.ndarray, seed: int, n_maps: int, terrain: str = "coarse") -> np.ndarray:
    return np.stack([generate_map(cfg, seed, i, terrain) for i in range(n_maps)]).astype(np.uint8)

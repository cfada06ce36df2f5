"""numpy restatement of the scripted survival policy (csrc/nmmo_obs.cu nmmo_forage_kernel), per agent record."""
import numpy as np

from nmmo_b200.config import SPEC, ObsLayout

MASK64 = (1 << 64) - 1
IMPASSIBLE = (1 << 0) | (1 << 1) | (1 << 5) | (1 << 14) | (1 << 15)      # Void, Water, Stone, Ocean, Fish (nmmo_spec.h)


def _mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def _hash64(seed, tick, site, idx, k):
    key = ((tick & 0xFFFFF) << 36) | ((site & 0xF) << 32) | ((idx & 0xFFFFFF) << 8) | (k & 0xFF)
    return _mix64((_mix64((seed ^ key) & MASK64) + key * 0xD6E8FEB86659FD93) & MASK64)


def _action_draw(base, head):
    x = ((base & 0xFFFFFFFF) + (base >> 32) * (2 * head + 1)) & 0xFFFFFFFF
    x ^= x >> 16; x = (x * 0x85EBCA6B) & 0xFFFFFFFF; x ^= x >> 13; x = (x * 0xC2B2AE35) & 0xFFFFFFFF; x ^= x >> 16
    return x


def forage_actions(cfg, obs, mask, tick, seed, env_global=0):
    """obs uint8 [P, stride] of one env, mask [P] -> int32 [P, 12]."""
    L = ObsLayout(cfg)
    P = obs.shape[0]
    win, vis = L.win, L.win // 2
    out = np.zeros((P, 12), np.int32)
    mv_off = L.masks["Move.Direction"][0]
    for p in range(P):
        rec = obs[p]
        my_id = int(rec[L.o_ids:L.o_ids + 2].view(np.int16)[0])
        if not mask[p] or my_id == 0:
            continue
        ent = rec[L.o_entity:L.o_entity + L.n_ent * 62].view(np.int16).reshape(L.n_ent, 31)
        rows = np.flatnonzero(ent[:, SPEC["EA_ID"]] == my_id)
        food = water = my_r = my_c = 0
        if len(rows):
            food, water = int(ent[rows[0], SPEC["EA_FOOD"]]), int(ent[rows[0], SPEC["EA_WATER"]])
            my_r, my_c = int(ent[rows[0], SPEC["EA_ROW"]]), int(ent[rows[0], SPEC["EA_COL"]])
        hungry = min(food, water) <= 60
        want = SPEC["MT_WATER"] if water <= food else SPEC["MT_FOILAGE"]
        tiles = rec[L.o_tile:L.o_tile + win * win * 6].view(np.int16).reshape(win * win, 3)[:, 2].astype(int)
        mv = rec[mv_off:mv_off + 5]
        d_out, decided = 4, not hungry
        if hungry:
            n_t = win * win
            walk = [not ((IMPASSIBLE >> m) & 1) for m in tiles]
            grid = tiles.reshape(win, win)

            def is_goal(w):
                if want == SPEC["MT_FOILAGE"]:
                    return tiles[w] == want
                r, c = divmod(w, win)
                return walk[w] and ((r > 0 and grid[r - 1, c] == want) or (r < win - 1 and grid[r + 1, c] == want) or
                                    (c > 0 and grid[r, c - 1] == want) or (c < win - 1 and grid[r, c + 1] == want))
            centre = vis * win + vis
            dist = [255] * n_t; first = [4] * n_t
            dist[centre] = 0
            hit = centre if is_goal(centre) else -1
            rnd = 1
            while rnd <= 48 and hit < 0:
                claimed = []
                for w in range(n_t):
                    if dist[w] != 255 or not walk[w]:
                        continue
                    r, c = divmod(w, win)
                    frm, step = -1, 4
                    if r > 0 and dist[w - win] == rnd - 1: frm, step = w - win, 1
                    elif r < win - 1 and dist[w + win] == rnd - 1: frm, step = w + win, 0
                    elif c > 0 and dist[w - 1] == rnd - 1: frm, step = w - 1, 2
                    elif c < win - 1 and dist[w + 1] == rnd - 1: frm, step = w + 1, 3
                    if frm < 0:
                        continue
                    claimed.append((w, step if rnd == 1 else first[frm]))
                for w, f in claimed:
                    dist[w] = rnd; first[w] = f
                goals = [w for w, _ in claimed if is_goal(w)]
                if goals:
                    hit = min(goals)
                if not claimed:
                    break
                rnd += 1
            if hit >= 0:
                d_out, decided = (4 if hit == centre else first[hit]), True
        if not decided:
            base = _hash64(seed + env_global, tick, SPEC["RS_ACTION"], p, 1)
            S = int(cfg[SPEC["NC_MAP_SIZE"]])
            dr, dc = S // 2 - my_r, S // 2 - my_c
            if (_action_draw(base, SPEC["AC_N"]) * 8) >> 32 != 0:
                vdir, hdir = (0 if dr < 0 else 1), (2 if dc > 0 else 3)
                f1, f2 = (vdir, hdir) if abs(dr) >= abs(dc) else (hdir, vdir)
                ok = lambda k: ((dr != 0) if k < 2 else (dc != 0)) and mv[k]  # noqa: E731
                if ok(f1):
                    d_out, decided = f1, True
                elif ok(f2):
                    d_out, decided = f2, True
            if not decided:
                valid = [k for k in range(4) if mv[k]]
                if valid:
                    d_out = valid[(_action_draw(base, SPEC["AC_MOVE_DIR"]) * len(valid)) >> 32]
        out[p] = [0, L.n_ent, L.n_mkt, L.n_inv, L.n_inv, L.n_ent, 0, L.n_ent, d_out, L.n_inv, 0, L.n_inv]
    return out

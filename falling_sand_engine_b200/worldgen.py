"""Synthetic worlds for the named benchmark / parity configurations.

Everything is a pure function of a counter hash h(seed, x, y), generated band by band so a
32768x32768 world never has to exist on the host at once.  The reference's world generator
(source/engine/world_generator.*) is out of scope; these are measurement fixtures.
"""
import ctypes as C

import numpy as np

from . import types as T


def _mix32(v):
    v = v.astype(np.uint32, copy=True)
    v ^= v >> np.uint32(16)
    v *= np.uint32(0x7FEB352D)
    v ^= v >> np.uint32(15)
    v *= np.uint32(0x846CA68B)
    v ^= v >> np.uint32(16)
    return v


def hash2(seed, x, y):
    """h(seed, x, y) on uint32 arrays (broadcasting)."""
    with np.errstate(over="ignore"):
        a = _mix32(np.asarray(x, dtype=np.uint32) * np.uint32(0x9E3779B1) + np.uint32(seed & 0xFFFFFFFF))
        return _mix32(a ^ (np.asarray(y, dtype=np.uint32) * np.uint32(0x85EBCA77) + np.uint32(0x165667B1)))


def cells_from_mat(table, mat, x0=0, y0=0, seed=1337, temp=None):
    """Build fse_cell records for a material-id array the way TilesCreate (game_datastruct.cpp:485-574)
    would: per-material colour policy and creation temperature, fluidAmount 2.0, nothing else set."""
    mat = np.asarray(mat)
    h, w = mat.shape
    out = np.zeros((h, w), dtype=T.CELL_DTYPE)
    out["mat"] = mat
    out["fluid"] = 2.0
    n = table.n
    base = np.array([m.color for m in table.mats], dtype=np.uint32)
    kind = np.array([m.color_kind for m in table.mats], dtype=np.uint8)
    ctemp = np.array([m.create_temp for m in table.mats], dtype=np.int16)
    jr = np.array([max(m.jitter_range, 1) for m in table.mats], dtype=np.uint32)
    js = np.array([m.jitter_shift for m in table.mats], dtype=np.uint32)
    xs = (np.arange(w, dtype=np.uint32) + np.uint32(x0))[None, :]
    ys = (np.arange(h, dtype=np.uint32) + np.uint32(y0))[:, None]
    hh = hash2(seed ^ 0xC0105, xs, ys)
    col = base[mat]
    k = kind[mat]
    col = np.where(k == T.COLOR_JITTER, col + ((hh % jr[mat]) << js[mat]), col)
    col = np.where(k == T.COLOR_POSITIONAL, col ^ (hh & np.uint32(0x0F0F0F)), col)
    out["color"] = col
    out["temp"] = ctemp[mat] if temp is None else temp
    assert n <= T.FSE_MAX_MATERIALS
    return out


def border_fill(mat, x0, y0, width, height, solid_id):
    """GENERIC_SOLID outside the tickZone (128-cell border)."""
    h, w = mat.shape
    xs = np.arange(w)[None, :] + x0
    ys = np.arange(h)[:, None] + y0
    b = T.FSE_CHUNK
    outside = (xs < b) | (xs >= width - b) | (ys < b) | (ys >= height - b)
    mat[np.broadcast_to(outside, mat.shape)] = solid_id
    return mat


def air_band(table, width, height, y0, rows, seed=1337):
    """Rows [y0, y0+rows) of an empty world: AIR inside the tickZone, GENERIC_SOLID border.  The do-nothing floor of the tick."""
    ids = _names(table)
    mat = np.full((rows, width), ids["AIR"], dtype=np.uint16)
    border_fill(mat, 0, y0, width, height, ids["GENERIC_SOLID"])
    return cells_from_mat(table, mat, 0, y0, seed)


# ---- config 1: 2048x2048 sand/water/stone column drop -------------------------------------
def column_drop_band(table, width, height, y0, rows, seed=1337, scale=None):
    """Rows [y0, y0+rows) of the column-drop world.  Coordinates of SURVEY §8d(1) are for
    2048x2048 and scale linearly with the world size (scale = width/2048)."""
    s = (width / 2048.0) if scale is None else scale
    ids = _names(table)
    mat = np.full((rows, width), ids["AIR"], dtype=np.uint16)
    xs = np.arange(width)[None, :]
    ys = (np.arange(rows) + y0)[:, None]

    def box(xa, xb, ya, yb, m):
        sel = (xs >= int(xa * s)) & (xs < int(xb * s)) & (ys >= int(ya * s)) & (ys < int(yb * s))
        mat[np.broadcast_to(sel, mat.shape)] = m

    box(0, 2048, 1856, 1920, ids["STONE"])      # floor
    box(128, 192, 128, 1920, ids["STONE"])      # walls
    box(1856, 1920, 128, 1920, ids["STONE"])
    box(384, 896, 192, 960, ids["GENERIC_SAND"])
    box(1152, 1664, 192, 960, ids["WATER"])
    # staggered 16x16 pegs every 128 cells in rows 1024..1536
    px = (xs // max(int(16 * s), 1))
    py = (ys // max(int(16 * s), 1))
    peg = (ys >= int(1024 * s)) & (ys < int(1536 * s)) & (py % 8 == 0) & (((px + 4 * ((py // 8) % 2)) % 8) == 0)
    peg = peg & (xs >= int(192 * s)) & (xs < int(1856 * s))
    mat[np.broadcast_to(peg, mat.shape)] = ids["STONE"]
    border_fill(mat, 0, y0, width, height, ids["GENERIC_SOLID"])
    return cells_from_mat(table, mat, 0, y0, seed)


# ---- config 2/3/4: mixed blobs ----------------------------------------------------------
def bench_table(base):
    """The 'Lua material table' of config 2: three registered materials (materials_register,
    game_basic.cpp:79-81) and interact=true with 2 TRANSFORM + 2 SPAWN interactions, so the
    branch world.cpp:1153-1179 is exercised.  Returns (table, {name: id})."""
    t = base.copy()
    n0 = t.n
    mats = list(t.mats)

    def reg(phys, slip, dens, iters, color):
        m = T.Material()
        m.physics, m.slipperyness, m.alpha, m.density, m.iterations = phys, slip, 255, dens, iters
        m.color, m.conduction_self, m.conduction_other, m.color_kind = color, 1.0, 1.0, T.COLOR_FIXED
        mats.append(m)
        return len(mats) - 1

    acid = reg(T.SAND, 10, 11.0, 2, 0x80FF20)   # eats stone below it
    seedm = reg(T.SAND, 12, 9.0, 2, 0x30A030)   # sprouts grass on dirt
    salt = reg(T.SAND, 16, 8.0, 2, 0xF0F0FF)    # fizzes on water
    n = len(mats)
    ids = _names(base)
    pairs = {
        (acid, ids["STONE"]): [(T.INTERACT_TRANSFORM_MATERIAL, ids["GENERIC_SAND"], 1, 0, 1)],
        (acid, ids["COBBLE_STONE"]): [(T.INTERACT_TRANSFORM_MATERIAL, ids["DIRT"], 2, 0, 2)],
        (seedm, ids["DIRT"]): [(T.INTERACT_SPAWN_MATERIAL, ids["GRASS"], 1, 0, -2)],
        (salt, ids["WATER"]): [(T.INTERACT_SPAWN_MATERIAL, ids["STEAM"], 1, 1, -2)],
    }
    # rebuild flattened arrays at the new material count (old random-material lists keep their slots)
    old = {}
    for a in range(n0):
        for b in range(n0):
            lo, hi = base.inter_offsets[a * n0 + b], base.inter_offsets[a * n0 + b + 1]
            if hi > lo:
                old[(a, b)] = [(base.inter[k].type, base.inter[k].data1, base.inter[k].data2, base.inter[k].ofs_x, base.inter[k].ofs_y)
                               for k in range(lo, hi)]
    old.update(pairs)
    flat, offs = [], [0]
    for a in range(n):
        for b in range(n):
            for (ty, prod, rad, ox, oy) in old.get((a, b), []):
                flat.append(T.Interaction(ty, prod, 0, rad, ox, oy))
            offs.append(len(flat))
    ro = list(base.react_offsets) + [base.react_offsets[n0]] * (n - n0)
    for (a, _b) in pairs:
        mats[a].interact = 1
    tbl = T.MaterialTable((T.Material * n)(*mats), t.ids, (T.Interaction * max(len(flat), 1))(*flat),
                          (C.c_int32 * len(offs))(*offs), t.react, (C.c_int32 * len(ro))(*ro))
    return tbl, {"ACID": acid, "SEED": seedm, "SALT": salt}


_DEFAULT_NAMES = ["AIR", "GENERIC_SOLID", "GENERIC_SAND", "GENERIC_LIQUID", "GENERIC_GAS", "GENERIC_PASSABLE", "GENERIC_OBJECT",
                  "STONE", "GRASS", "DIRT", "SMOOTH_STONE", "COBBLE_STONE", "SMOOTH_DIRT", "COBBLE_DIRT", "SOFT_DIRT", "WATER", "LAVA",
                  "CLOUD", "GOLD_ORE", "GOLD_MOLTEN", "GOLD_SOLID", "IRON_ORE", "OBSIDIAN", "STEAM", "SOFT_DIRT_SAND", "FIRE",
                  "FLAT_COBBLE_STONE", "FLAT_COBBLE_DIRT"]


def _names(table):
    """name -> id of the fixed materials (ids are materials_count++ order, gds.cpp:69-111)."""
    return {n: i for i, n in enumerate(_DEFAULT_NAMES)}


def mixed_band(table, width, height, y0, rows, seed=1337, air_frac=0.35, extra=None, blob=64):
    """Rows [y0, y0+rows) of the mixed-material world: `blob`x`blob`-cell blobs, per blob
    35% AIR, 20% powders, 15% liquids, 8% gases, 20% solids, 2% FIRE lining solid blobs; random temperatures
    on 10% of cells.  `extra` = ids of the registered interacting powders (bench_table)."""
    ids = _names(table)
    xs = np.arange(width, dtype=np.uint32)[None, :]
    ys = (np.arange(rows, dtype=np.uint32) + np.uint32(y0))[:, None]
    bx, by = xs // blob, ys // blob
    hb = hash2(seed, bx, by)
    u = (hb & np.uint32(0xFFFF)).astype(np.float32) / 65536.0
    pick = (hb >> np.uint32(16)) & np.uint32(0xFFFF)
    scale = (1.0 - air_frac) / 0.65
    e_pow = air_frac + 0.20 * scale
    e_liq = e_pow + 0.15 * scale
    e_gas = e_liq + 0.08 * scale
    powders = np.array([ids["GENERIC_SAND"], ids["DIRT"], ids["GOLD_ORE"], ids["SOFT_DIRT_SAND"]] + list(extra or []), dtype=np.uint16)
    liquids = np.array([ids["WATER"], ids["WATER"], ids["LAVA"], ids["GOLD_MOLTEN"]], dtype=np.uint16)
    gases = np.array([ids["STEAM"], ids["GENERIC_GAS"]], dtype=np.uint16)
    solids = np.array([ids["STONE"], ids["COBBLE_STONE"], ids["OBSIDIAN"]], dtype=np.uint16)
    mat = np.full((rows, width), ids["AIR"], dtype=np.uint16)
    u = np.broadcast_to(u, mat.shape)
    pick = np.broadcast_to(pick, mat.shape)
    is_pow = (u >= air_frac) & (u < e_pow)
    is_liq = (u >= e_pow) & (u < e_liq)
    is_gas = (u >= e_liq) & (u < e_gas)
    is_sol = u >= e_gas
    mat = np.where(is_pow, powders[pick % len(powders)], mat)
    mat = np.where(is_liq, liquids[pick % len(liquids)], mat)
    mat = np.where(is_gas, gases[pick % len(gases)], mat)
    mat = np.where(is_sol, solids[pick % len(solids)], mat)
    # FIRE lining: outer 2-cell ring of solid blobs, 40% of those cells (~2% of the world)
    lx, ly = np.broadcast_to(xs % blob, mat.shape), np.broadcast_to(ys % blob, mat.shape)
    ring = (lx < 2) | (lx >= blob - 2) | (ly < 2) | (ly >= blob - 2)
    hc = hash2(seed ^ 0xF12E, xs, ys)
    fire = is_sol & ring & ((hc % np.uint32(100)) < 40)
    mat = np.where(fire, np.uint16(ids["FIRE"]), mat).astype(np.uint16)
    border_fill(mat, 0, y0, width, height, ids["GENERIC_SOLID"])
    cells = cells_from_mat(table, mat, 0, y0, seed)
    ht = hash2(seed ^ 0x7E39, xs, ys)
    hot = (ht % np.uint32(10)) == 0
    tv = ((ht >> np.uint32(8)) % np.uint32(2048)).astype(np.int32) - 1023
    cells["temp"] = np.where(np.broadcast_to(hot, mat.shape), tv.astype(np.int16), cells["temp"])
    return cells


# ---- config 5: mostly settled, sparse activity -------------------------------------------
def sparse_band(table, width, height, y0, rows, seed=1337, pockets=256):
    """Rows [y0, y0+rows) of the sparse world: solid terrain with settled sand
    (moved=false) and settled water (moved=true) layers; `pockets` hashed 128x128 chunks hold falling sand/water."""
    ids = _names(table)
    xs = np.arange(width, dtype=np.uint32)[None, :]
    ys = (np.arange(rows, dtype=np.uint32) + np.uint32(y0))[:, None]
    mat = np.full((rows, width), ids["STONE"], dtype=np.uint16)
    cx, cy = xs // 128, ys // 128
    ncx, ncy = width // 128, height // 128
    hc = hash2(seed ^ 0x5A5A, cx, cy)
    frac = pockets / float(max((ncx - 2) * (ncy - 2), 1))
    pocket = (hc.astype(np.float64) / 4294967296.0) < frac
    lx, ly = xs % 128, ys % 128
    inner = (lx >= 8) & (lx < 120) & (ly >= 8) & (ly < 120)
    pk = np.broadcast_to(pocket & inner, mat.shape)
    mat[pk] = ids["AIR"]
    # a sand pillar and a water column standing on the pocket floor: they collapse / spread by the grid rules for
    # hundreds of ticks (no free fall into particles), which keeps exactly the pocket chunks active
    sand_blk = np.broadcast_to(pocket & (lx >= 40) & (lx < 46) & (ly >= 40) & (ly < 120), mat.shape)
    water_blk = np.broadcast_to(pocket & (lx >= 80) & (lx < 88) & (ly >= 56) & (ly < 120), mat.shape)
    mat[sand_blk] = ids["GENERIC_SAND"]
    mat[water_blk] = ids["WATER"]
    # settled strata inside the rock: sealed sand and water lenses
    lens = np.broadcast_to((~pocket) & (ly >= 40) & (ly < 88) & (lx >= 16) & (lx < 112), mat.shape)
    kind = np.broadcast_to((hc >> np.uint32(8)) % np.uint32(4), mat.shape)
    mat[lens & (kind == 0)] = ids["GENERIC_SAND"]
    mat[lens & (kind == 1)] = ids["WATER"]
    border_fill(mat, 0, y0, width, height, ids["GENERIC_SOLID"])
    cells = cells_from_mat(table, mat, 0, y0, seed)
    settled_water = lens & (kind == 1)
    cells["moved"] = np.where(settled_water, 1, 0).astype(np.uint8)
    cells["fluid"] = np.where(settled_water, np.float32(0.5), cells["fluid"])
    cells["settle"] = np.where(settled_water, 10, 0).astype(np.uint8)
    return cells


def fill_world(world, band_fn, width, height, band_rows=1024, y_lo=0, y_hi=None, **kw):
    """Upload a generated world band by band through write_rect (any object with that method)."""
    y_hi = height if y_hi is None else y_hi
    for y0 in range(y_lo, y_hi, band_rows):
        rows = min(band_rows, y_hi - y0)
        world.write_rect(0, y0, band_fn(width=width, height=height, y0=y0, rows=rows, **kw))

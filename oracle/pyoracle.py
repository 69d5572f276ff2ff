"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from falling_sand_engine_b200 import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfse_oracle.so")
_lib = None

REFERENCE, PARTITIONED, ROWS = 0, 1, 2
RNG_SLOT, RNG_LIBC = 0, 1


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("fse_oracle.cpp", "oracle_capi.cpp", "fse_oracle.hpp", "bridge_oracle.cpp")]
    srcs = [s for s in srcs if os.path.exists(s)] + [os.path.join(_HERE, "..", "include", "fse.h")]
    if not force and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.fseo_world_create.restype = C.c_void_p
        L.fseo_world_create.argtypes = [C.c_int, C.c_int]
        L.fseo_world_destroy.argtypes = [C.c_void_p]
        L.fseo_tick.restype = C.c_double
        L.fseo_tick.argtypes = [C.c_void_p, C.POINTER(T.TickArgs), C.c_int, C.c_int, C.c_int]
        L.fseo_particles_count.restype = C.c_int64
        L.fseo_particles_count.argtypes = [C.c_void_p]
        L.fseo_cell_hash.restype = C.c_uint64
        L.fseo_rng_draw.restype = C.c_uint32
        L.fseo_rng_draw.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32]
        for name in ("fseo_materials_set", "fseo_write_rect", "fseo_read_rect", "fseo_clear_dirty", "fseo_stats_rect",
                     "fseo_run_chunk", "fseo_clear_visited", "fseo_tick_temperature", "fseo_particles_add",
                     "fseo_particles_read", "fseo_particles_clear", "fseo_tick_particles"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def default_materials(seed=1337):
    L = lib()
    n, ni, nr = C.c_int(), C.c_int(), C.c_int()
    L.fseo_default_materials(C.c_uint32(seed), None, C.byref(n), None, None, C.byref(ni), None, None, C.byref(nr), None)
    mats = (T.Material * n.value)()
    ids = T.SpecialIds()
    inter = (T.Interaction * max(ni.value, 1))()
    io = (C.c_int32 * (n.value * n.value + 1))()
    react = (T.Interaction * max(nr.value, 1))()
    ro = (C.c_int32 * (n.value + 1))()
    L.fseo_default_materials(C.c_uint32(seed), mats, C.byref(n), C.byref(ids), inter, C.byref(ni), io, react, C.byref(nr), ro)
    return T.MaterialTable(mats, ids, inter, io, react, ro)


class OracleWorld:
    """Oracle world with the same method names as falling_sand_engine_b200.World."""

    def __init__(self, width, height, table=None):
        self.L = lib()
        self.width, self.height = width, height
        self.h = C.c_void_p(self.L.fseo_world_create(width, height))
        self.table = table or default_materials()
        self.L.fseo_materials_set(self.h, *self.table.args())
        self.default_schedule = ROWS  # the product's default in-row schedule (FSE_SCHEDULE_ROWS)

    def close(self):
        if self.h:
            self.L.fseo_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_materials(self, table):
        self.table = table
        self.L.fseo_materials_set(self.h, *table.args())

    def write_rect(self, x, y, cells):
        cells = np.ascontiguousarray(cells, dtype=T.CELL_DTYPE)
        h, w = cells.shape
        self.L.fseo_write_rect(self.h, x, y, w, h, cells.ctypes.data_as(C.c_void_p))

    def read_rect(self, x, y, w, h):
        out = np.zeros((h, w), dtype=T.CELL_DTYPE)
        self.L.fseo_read_rect(self.h, x, y, w, h, out.ctypes.data_as(C.c_void_p))
        return out

    def read_all(self):
        return self.read_rect(0, 0, self.width, self.height)

    def clear_dirty(self):
        self.L.fseo_clear_dirty(self.h)

    def stats(self, rect=None):
        r = rect or T.Rect(0, 0, self.width, self.height)
        s = T.Stats()
        self.L.fseo_stats_rect(self.h, r.x, r.y, r.w, r.h, C.byref(s))
        return s

    def tick(self, tick, seed=1337, cell_iter=3, zone=None, schedule=None, rng=RNG_SLOT, threads=1):
        schedule = self.default_schedule if schedule is None else schedule
        a = T.TickArgs(tick, seed, cell_iter, zone or T.zone_of(self.width, self.height))
        return self.L.fseo_tick(self.h, C.byref(a), schedule, rng, threads)

    def run_chunk(self, tick, seed, it, cx, cy, schedule=None, zone=None):
        schedule = self.default_schedule if schedule is None else schedule
        a = T.TickArgs(tick, seed, 3, zone or T.zone_of(self.width, self.height))
        self.L.fseo_run_chunk(self.h, C.byref(a), it, cx, cy, schedule)

    def clear_visited(self):
        self.L.fseo_clear_visited(self.h)

    def tick_temperature(self, zone=None):
        z = zone or T.zone_of(self.width, self.height)
        self.L.fseo_tick_temperature(self.h, C.byref(z))

    def particles_add(self, parts):
        parts = np.ascontiguousarray(parts, dtype=T.PARTICLE_DTYPE)
        self.L.fseo_particles_add(self.h, parts.ctypes.data_as(C.c_void_p), len(parts))

    def particles_count(self):
        return int(self.L.fseo_particles_count(self.h))

    def particles_read(self):
        n = self.particles_count()
        out = np.zeros(n, dtype=T.PARTICLE_DTYPE)
        if n:
            self.L.fseo_particles_read(self.h, out.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return out

    def particles_clear(self):
        self.L.fseo_particles_clear(self.h)

    def particles_tick(self, zone=None, schedule=PARTITIONED, max_rounds=16):
        """tickCells: schedule=REFERENCE walks the list in order (world.cpp:2030-2195); PARTITIONED is the GPU's
        snapshot + lowest-id-wins rounds."""
        z = zone or T.zone_of(self.width, self.height)
        if schedule == REFERENCE:
            self.L.fseo_tick_particles(self.h, C.byref(z))
        else:
            self.L.fseo_tick_particles_rounds(self.h, C.byref(z), max_rounds)


def _body_args(bodies):
    n = len(bodies)
    bw = (C.c_int * max(n, 1))(*[b.shape[1] for b in bodies])
    bh = (C.c_int * max(n, 1))(*[b.shape[0] for b in bodies])
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bodies])
    return n, bw, bh, ptrs


def bodies_raster(world, bodies, xforms, tick=0, seed=1337):
    """game.cpp:1711-1815 on an OracleWorld; `bodies` = list of (h, w) CELL_DTYPE arrays (modified in place by erase)."""
    n, bw, bh, ptrs = _body_args(bodies)
    xf = np.ascontiguousarray(xforms, dtype=np.float32).reshape(-1, 3)
    fb = np.zeros((n, 4), dtype=np.int32)
    lib().fseo_bodies_raster(world.h, n, bw, bh, ptrs, xf.ctypes.data_as(C.c_void_p), C.c_uint32(tick), C.c_uint32(seed),
                             fb.ctypes.data_as(C.c_void_p))
    return fb


# a deposit proposal as exchanged between strip ranks (oracle/fse_oracle.hpp fseo_proposal, 48 bytes)
PROPOSAL_DTYPE = np.dtype({"names": ["cell", "id", "tile", "merge", "_pad"], "formats": [np.int64, np.uint64, T.CELL_DTYPE, np.int32, np.int32],
                           "offsets": [0, 8, 16, 36, 40], "itemsize": 48})


def prt_begin(world, zone=None):
    """Stage 1 of tick_particles_rounds: integrate every particle of this world against its grid."""
    z = zone or T.zone_of(world.width, world.height)
    lib().fseo_prt_begin.argtypes = [C.c_void_p, C.c_void_p]
    lib().fseo_prt_begin(world.h, C.byref(z))


def prt_propose(world):
    """Stage 2 (one per round): every pending particle proposes a cell; returns the number of proposals."""
    lib().fseo_prt_propose.argtypes = [C.c_void_p]
    return int(lib().fseo_prt_propose(world.h))


def prt_get(world, y0, y1):
    """The proposals of this round whose cell lies in rows [y0, y1)."""
    f = lib().fseo_prt_get
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = int(f(world.h, y0, y1, None, 0))
    out = np.zeros(n, dtype=PROPOSAL_DTYPE)
    if n:
        f(world.h, y0, y1, out.ctypes.data, n)
    return out


def prt_commit(world, ext=None, hold_lo=0, hold_hi=None):
    """Stage 3: lowest id wins a cell among this world's and the other ranks' proposals (`ext`); winners' cells are written where
    this world holds them (rows [hold_lo, hold_hi))."""
    ext = np.zeros(0, dtype=PROPOSAL_DTYPE) if ext is None else np.ascontiguousarray(ext, dtype=PROPOSAL_DTYPE)
    lib().fseo_prt_commit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib().fseo_prt_commit(world.h, ext.ctypes.data if len(ext) else None, len(ext), hold_lo, world.height if hold_hi is None else hold_hi)


def prt_end(world):
    """Stage 4: drop the deposited / dead particles, keep the rest."""
    lib().fseo_prt_end.argtypes = [C.c_void_p]
    lib().fseo_prt_end(world.h)


def render_dirty(world, planes, with_flow_count=False):
    """game.cpp:1994-2066 on the oracle world; planes = (main, fire, emission[, flow]) uint8 arrays (h, w, 4), updated in place (the
    flow texture and the flowX / flowY reset need flow_enable(world) and a fourth plane).
    Returns (dirty cells, dirty FIRE cells, movingTiles[n_materials]) [+ dirty SOUP cells with with_flow_count]."""
    moving = np.zeros(256, dtype=np.int64)
    had = np.zeros(3, dtype=np.int64)
    lib().fseo_render_dirty.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    flow = planes[3].ctypes.data if len(planes) > 3 else None
    lib().fseo_render_dirty(world.h, planes[0].ctypes.data, planes[1].ctypes.data, planes[2].ctypes.data, flow, moving.ctypes.data, had.ctypes.data)
    if with_flow_count:
        return int(had[0]), int(had[1]), moving, int(had[2])
    return int(had[0]), int(had[1]), moving


def physics_check(world, x, y, cap_tiles=None):
    """world::physicsCheck (world.cpp:3330-3411).  Returns (count, action, (x, y, w, h), tiles or None).  cap_tiles: capacity of the
    tile buffer (default 65536, retried with the component's box when that is too small; nothing changes when it does not fit)."""
    res = np.zeros(6, dtype=np.int32)
    cap = cap_tiles or (1 << 16)
    tiles = np.zeros(cap, dtype=T.CELL_DTYPE)
    lib().fseo_physics_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    rc = lib().fseo_physics_check(world.h, x, y, res.ctypes.data, tiles.ctypes.data, cap)
    if rc != 0 and cap_tiles is None:
        cap = int(res[4]) * int(res[5])
        tiles = np.zeros(cap, dtype=T.CELL_DTYPE)
        rc = lib().fseo_physics_check(world.h, x, y, res.ctypes.data, tiles.ctypes.data, cap)
    if rc != 0:
        raise ValueError(f"physics_check: the {res[4]} x {res[5]} box does not fit {cap} tiles")
    box = tuple(int(v) for v in res[2:6])
    return int(res[0]), int(res[1]), box, (tiles[:box[2] * box[3]].reshape(box[3], box[2]).copy() if res[1] == 2 else None)


S_PROBE_X, S_PROBE_Y = 69, 70


def probe_position(world, tick, seed=1337, zone=None):
    """world.cpp:1930-1931 with the counter RNG: draws S_PROBE_X / S_PROBE_Y of cell (0, 0) under rng_key(seed, tick, 15)."""
    z = zone or T.Rect(T.FSE_CHUNK, T.FSE_CHUNK, world.width - 2 * T.FSE_CHUNK, world.height - 2 * T.FSE_CHUNK)
    L = lib()
    L.fseo_rng_draw.restype = C.c_uint32
    L.fseo_rng_draw.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_uint32]
    return (z.x + L.fseo_rng_draw(seed, tick, 15, 0, 0, S_PROBE_X) % max(z.w, 1), z.y + L.fseo_rng_draw(seed, tick, 15, 0, 0, S_PROBE_Y) % max(z.h, 1))


def physics_probe(world, tick, seed=1337, zone=None):
    """The probe at the end of world::tick (world.cpp:1929-1934)."""
    return physics_check(world, *probe_position(world, tick, seed, zone))


def flow_enable(world):
    """Carry flowX / flowY / prevFlowX / prevFlowY (world.hpp:116-119) from now on."""
    lib().fseo_flow_enable.argtypes = [C.c_void_p]
    lib().fseo_flow_enable(world.h)


def flow_read(world, which):
    """Whole plane: 0 flowX, 1 flowY, 2 prevFlowX, 3 prevFlowY."""
    out = np.zeros((world.height, world.width), dtype=np.float32)
    lib().fseo_flow_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    assert lib().fseo_flow_read(world.h, which, out.ctypes.data) == 0, "flow_enable first"
    return out


def layer2_write_rect(world, x, y, cells):
    c = np.ascontiguousarray(cells, dtype=T.CELL_DTYPE)
    lib().fseo_layer2_write_rect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib().fseo_layer2_write_rect(world.h, x, y, c.shape[1], c.shape[0], c.ctypes.data)


def layer2_read_rect(world, x, y, w, h):
    out = np.zeros((h, w), dtype=T.CELL_DTYPE)
    lib().fseo_layer2_read_rect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib().fseo_layer2_read_rect(world.h, x, y, w, h, out.ctypes.data)
    return out


def background_write_rect(world, x, y, colors):
    c = np.ascontiguousarray(colors, dtype=np.uint32)
    lib().fseo_background_write_rect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib().fseo_background_write_rect(world.h, x, y, c.shape[1], c.shape[0], c.ctypes.data)


def background_read_rect(world, x, y, w, h):
    out = np.zeros((h, w), dtype=np.uint32)
    lib().fseo_background_read_rect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib().fseo_background_read_rect(world.h, x, y, w, h, out.ctypes.data)
    return out


def render_layers(world, planes, draw_background_grid=False):
    """game.cpp:2068-2126 + the dirty clears of 2154-2155; planes = (layer2, background) uint8 arrays (h, w, 4), updated in place.
    Returns (dirty layer-2 cells, dirty background cells)."""
    had = np.zeros(2, dtype=np.int64)
    lib().fseo_render_layers.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib().fseo_render_layers(world.h, 1 if draw_background_grid else 0, planes[0].ctypes.data, planes[1].ctypes.data, had.ctypes.data)
    return int(had[0]), int(had[1])


def scroll(world, dx, dy):
    """world::tickChunks grid + particle shift (world.cpp:2454-2478, 2579-2582)."""
    lib().fseo_scroll.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib().fseo_scroll(world.h, dx, dy)


def entities_tick(world, ents, load_zone=(0.0, 0.0), tick=0, seed=1337):
    """world::tickEntities (world.cpp:3010-3247); returns the updated entity array."""
    e = np.ascontiguousarray(ents, dtype=T.ENTITY_DTYPE).copy()
    lib().fseo_entities_tick.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_uint32, C.c_uint32]
    lib().fseo_entities_tick(world.h, e.ctypes.data, len(e), load_zone[0], load_zone[1], tick, seed)
    return e


def entities_stamp(world, ents, load_zone=(0.0, 0.0), object_mat=6, tick=0, seed=1337):
    """WorldEntitySystem::process (game/player.cpp:173-199)."""
    e = np.ascontiguousarray(ents, dtype=T.ENTITY_DTYPE)
    lib().fseo_entities_stamp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_uint32, C.c_uint32]
    lib().fseo_entities_stamp(world.h, e.ctypes.data, len(e), load_zone[0], load_zone[1], object_mat, tick, seed)


def object_delete(world):
    """the objectDelete loop of game::tick (game.cpp:2128-2139)."""
    lib().fseo_object_delete.argtypes = [C.c_void_p]
    lib().fseo_object_delete(world.h)


def tool_erase_line(world, x0, y0, x1, y1, brush_size=5):
    """erase brush (game.cpp:593-625); returns the number of cells cleared."""
    lib().fseo_tool_erase_line.argtypes = [C.c_void_p] + [C.c_int] * 5
    return lib().fseo_tool_erase_line(world.h, x0, y0, x1, y1, brush_size)


def tool_pickaxe(world, x, y, break_size):
    size = int(break_size)
    pix = np.zeros((size, size), dtype=np.uint32)
    lib().fseo_tool_pickaxe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    n = lib().fseo_tool_pickaxe(world.h, x, y, break_size, pix.ctypes.data)
    return pix, n


def tool_hammer(world, hammer_x, hammer_y, x, y, sand_mat=2, tick=0, seed=1337):
    out = np.zeros(4, dtype=np.int32)
    lib().fseo_tool_hammer.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_uint32, C.c_uint32, C.c_void_p]
    lib().fseo_tool_hammer(world.h, hammer_x, hammer_y, x, y, sand_mat, tick, seed, out.ctypes.data)
    return tuple(int(v) for v in out)


def tool_vacuum(world, wcx, wcy, wmx, wmy, tick=0, seed=1337):
    out = np.zeros(4, dtype=np.int32)
    lib().fseo_tool_vacuum.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_uint32, C.c_uint32, C.c_void_p]
    lib().fseo_tool_vacuum(world.h, wcx, wcy, wmx, wmy, tick, seed, out.ctypes.data)
    return tuple(int(v) for v in out)


def particles_vacuum_pull(world, target_x, target_y):
    lib().fseo_particles_vacuum_pull.argtypes = [C.c_void_p, C.c_float, C.c_float]
    return lib().fseo_particles_vacuum_pull(world.h, target_x, target_y)


def explosion(world, cx, cy, radius, tick=0, seed=1337):
    """world::explosion (world.cpp:2294-2332)."""
    lib().fseo_explosion.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32]
    lib().fseo_explosion(world.h, cx, cy, radius, tick, seed)


def bodies_erase(world, bodies, xforms):
    n, bw, bh, ptrs = _body_args(bodies)
    xf = np.ascontiguousarray(xforms, dtype=np.float32).reshape(-1, 3)
    fb = np.zeros((n, 4), dtype=np.int32)
    lib().fseo_bodies_erase(world.h, n, bw, bh, ptrs, xf.ctypes.data_as(C.c_void_p), fb.ctypes.data_as(C.c_void_p))
    return fb


def outlines(mask):
    """Contours of one (h, w) uint8 mask: list of (k, 2) float32 arrays (FindPerimeter + simplify(.., 1))."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    cap_p, cap_c = mask.size * 2 + 64, mask.size // 2 + 64
    pts = np.zeros((cap_p, 2), dtype=np.float32)
    off = np.zeros(cap_c + 1, dtype=np.int32)
    n = lib().fseo_outlines(mask.ctypes.data_as(C.c_void_p), w, h, pts.ctypes.data_as(C.c_void_p), cap_p, off.ctypes.data_as(C.c_void_p), cap_c)
    assert n >= 0
    return [pts[off[k]:off[k + 1]].copy() for k in range(n)]


def body_split(tiles, angle=0.0, weld=(-1, -1), air=0):
    """Fracture hand-off of one body (tiles: (h, w) CELL_DTYPE): list of (piece record, (h', w') tile array)."""
    tiles = np.ascontiguousarray(tiles, dtype=T.CELL_DTYPE)
    h, w = tiles.shape
    pieces = np.zeros(1024, dtype=T.BODY_PIECE_DTYPE)
    out = np.zeros(4 * h * w + 64, dtype=T.CELL_DTYPE)
    lib().fseo_body_split.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_long]
    n = lib().fseo_body_split(tiles.ctypes.data, w, h, air, angle, weld[0], weld[1], pieces.ctypes.data, len(pieces), out.ctypes.data, len(out))
    assert n >= 0
    return [(pieces[k].copy(), out[pieces[k]["tile_off"]: pieces[k]["tile_off"] + pieces[k]["w"] * pieces[k]["h"]].reshape(pieces[k]["h"], pieces[k]["w"]).copy())
            for k in range(n)]


def ms_value(mask, x, y):
    """MarchingSquares::value (physics_math.cpp:1882-1890)."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    lib().fseo_ms_value.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    return lib().fseo_ms_value(mask.ctypes.data, mask.shape[1], mask.shape[0], x, y)


def p_distance(x, y, x1, y1, x2, y2):
    """pDistance (physics_math.cpp:1813-1843)."""
    lib().fseo_p_distance.restype = C.c_float
    lib().fseo_p_distance.argtypes = [C.c_float] * 6
    return lib().fseo_p_distance(x, y, x1, y1, x2, y2)


def ccl(mask):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    labels = np.zeros((h, w), dtype=np.int32)
    n = lib().fseo_ccl(mask.ctypes.data_as(C.c_void_p), w, h, labels.ctypes.data_as(C.c_void_p))
    return labels, n


def flood_component(world, x, y, cap=1000):
    bbox = np.zeros(4, dtype=np.int32)
    pix = np.zeros(cap + 1, dtype=np.int32)
    n = lib().fseo_flood_component(world.h, x, y, cap, bbox.ctypes.data_as(C.c_void_p), pix.ctypes.data_as(C.c_void_p))
    return n, bbox, pix[:n] if n <= cap else pix[:0]


def rng_draw(seed, tick, it, x, y, slot):
    return int(lib().fseo_rng_draw(seed, tick, it, x, y, slot))

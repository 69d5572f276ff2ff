"""ctypes binding of libfse_b200.so (include/fse.h) and the host-side `World` mirror.

`World` keeps the reference's method names for the tick path (source/engine/world.hpp:148-192):
tick / tickTemperature / tickCells / addCell / getTile / setTile.  Everything runs on the GPU through
the C ABI; if the shared library is missing or no CUDA device is present this module raises — there is
no CPU fallback.
"""
import ctypes as C
import time
import os

import numpy as np

from . import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfse_b200.so")
_lib = None


class FseError(RuntimeError):
    pass


EXPORTS = [
    "fse_ctx_create", "fse_ctx_destroy", "fse_last_error", "fse_version", "fse_abi_sizeof", "fse_materials_set",
    "fse_world_create", "fse_world_destroy", "fse_sync", "fse_write_rect", "fse_read_rect", "fse_clear_dirty", "fse_stats_rect",
    "fse_tick", "fse_tick_temperature", "fse_particles_add", "fse_particles_tick", "fse_particles_count", "fse_particles_read",
    "fse_particles_clear", "fse_particles_reserve", "fse_timer_start", "fse_timer_stop", "fse_launch_count",
    "fse_kernel_timing_enable", "fse_kernel_timing_read", "fse_kernel_timing_phases", "fse_strip_timeline_read",
]


def load_library(path=None):
    """dlopen the C-ABI library.  Fails loudly when it has not been built (`__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("FSE_B200_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise FseError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    L.fse_last_error.restype = C.c_char_p
    L.fse_version.restype = C.c_char_p
    L.fse_launch_count.restype = C.c_int64
    L.fse_launch_count.argtypes = [C.c_void_p]
    L.fse_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.fse_ctx_destroy.argtypes = [C.c_void_p]
    L.fse_ctx_destroy.restype = None
    L.fse_world_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.fse_world_destroy.argtypes = [C.c_void_p]
    L.fse_world_destroy.restype = None
    L.fse_materials_set.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for n in ("fse_sync", "fse_clear_dirty", "fse_particles_clear", "fse_timer_start"):
        getattr(L, n).argtypes = [C.c_void_p]
    L.fse_write_rect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.fse_read_rect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.fse_stats_rect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.fse_tick.argtypes = [C.c_void_p, C.c_void_p]
    L.fse_tick_temperature.argtypes = [C.c_void_p, C.c_void_p]
    L.fse_particles_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.fse_particles_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.fse_particles_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.fse_particles_reserve.argtypes = [C.c_void_p, C.c_int64]
    if hasattr(L, "fse_particles_tick"):
        L.fse_particles_tick.argtypes = [C.c_void_p, C.c_void_p]
    L.fse_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.fse_kernel_timing_enable.argtypes = [C.c_void_p, C.c_int]
    L.fse_kernel_timing_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise FseError(f"fse error {rc}: {load_library().fse_last_error().decode()}")


def hitbox_triangles(contours):
    """The host half of updateRigidBodyHitbox / updateChunkMesh (world.cpp:497-563; csrc/polygons.hpp): the outlines of one mask
    (list of (k, 2) arrays as fse_mask_outline returns them) -> list of (n, 3, 2) float64 triangle arrays, one per outer polygon."""
    L = load_library()
    pts = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float32).reshape(-1, 2) for c in contours]) if len(contours) else
                               np.zeros((0, 2), np.float32))
    off = np.zeros(len(contours) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(c) for c in contours])
    cap_t, cap_g = 2 * len(pts) + 16, len(contours) + 1
    tris = np.zeros((cap_t, 3, 2), dtype=np.float64)
    goff = np.zeros(cap_g + 1, dtype=np.int32)
    ng = C.c_int32(0)
    L.fse_hitbox_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    _ck(L.fse_hitbox_triangles(pts.ctypes.data, off.ctypes.data, len(contours), tris.ctypes.data, cap_t, goff.ctypes.data, cap_g, C.byref(ng)))
    return [tris[goff[g]:goff[g + 1]].copy() for g in range(ng.value)]


class Context:
    def __init__(self, device=0, table=None):
        self.L = load_library()
        self.h = C.c_void_p()
        _ck(self.L.fse_ctx_create(device, C.byref(self.h)))
        self.table = None
        if table is not None:
            self.set_materials(table)

    def set_materials(self, table):
        """materials_init / materials_register / materials_push (game_basic.cpp:79-81)."""
        _ck(self.L.fse_materials_set(self.h, table.mats, table.n, C.byref(table.ids), table.inter, table.inter_offsets, table.react,
                                     table.react_offsets))
        self.table = table

    def launch_count(self):
        return int(self.L.fse_launch_count(self.h))

    def close(self):
        if self.h:
            self.L.fse_ctx_destroy(self.h)
            self.h = None


class World:
    """Device-resident world (reference `class world`, world.hpp:63-193) for the tick path."""

    def __init__(self, ctx, width, height):
        self.L = ctx.L
        self.ctx = ctx
        self.width, self.height = width, height
        self.h = C.c_void_p()
        _ck(self.L.fse_world_create(ctx.h, width, height, C.byref(self.h)))
        self.tickZone = T.zone_of(width, height)
        self.schedule = 1

    def close(self):
        if self.h:
            self.L.fse_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- boundary -------------------------------------------------------------------------------
    def write_rect(self, x, y, cells):
        cells = np.ascontiguousarray(cells, dtype=T.CELL_DTYPE)
        h, w = cells.shape
        _ck(self.L.fse_write_rect(self.h, x, y, w, h, cells.ctypes.data_as(C.c_void_p)))

    def write_rect_ptr(self, x, y, w, h, ptr):
        _ck(self.L.fse_write_rect(self.h, x, y, w, h, C.c_void_p(ptr)))

    def read_rect(self, x, y, w, h, out=None):
        out = np.zeros((h, w), dtype=T.CELL_DTYPE) if out is None else out
        _ck(self.L.fse_read_rect(self.h, x, y, w, h, out.ctypes.data_as(C.c_void_p)))
        return out

    def read_all(self):
        return self.read_rect(0, 0, self.width, self.height)

    def getTile(self, x, y):
        return self.read_rect(x, y, 1, 1)[0, 0]

    def setTile(self, x, y, cell):
        a = np.zeros((1, 1), dtype=T.CELL_DTYPE)
        a[0, 0] = cell
        a["dirty"] = 1  # world::setTile marks dirty (world.cpp:1007)
        self.write_rect(x, y, a)

    def clear_dirty(self):
        _ck(self.L.fse_clear_dirty(self.h))

    def stats(self, rect=None):
        r = rect or T.Rect(0, 0, self.width, self.height)
        s = T.Stats()
        _ck(self.L.fse_stats_rect(self.h, r.x, r.y, r.w, r.h, C.byref(s)))
        return s

    def sync(self):
        _ck(self.L.fse_sync(self.h))

    # -- the tick ---------------------------------------------------------------------------------
    def tick(self, tick, seed=1337, cell_iter=3, zone=None):
        a = T.TickArgs(tick, seed, cell_iter, zone or self.tickZone)
        _ck(self.L.fse_tick(self.h, C.byref(a)))

    def set_schedule(self, schedule):
        """1 = rows schedule, one kernel per pass (default); 2 = the same results from the fused kernel (oracle ROWS either way)."""
        self.L.fse_set_schedule.argtypes = [C.c_void_p, C.c_int]
        _ck(self.L.fse_set_schedule(self.h, schedule))
        self.schedule = schedule

    def tickTemperature(self, zone=None):
        z = zone or self.tickZone
        _ck(self.L.fse_tick_temperature(self.h, C.byref(z)))

    tick_temperature = tickTemperature

    # -- particles ----------------------------------------------------------------------------------
    def addCell(self, parts):
        parts = np.ascontiguousarray(np.atleast_1d(parts), dtype=T.PARTICLE_DTYPE)
        _ck(self.L.fse_particles_add(self.h, parts.ctypes.data_as(C.c_void_p), len(parts)))

    particles_add = addCell

    def tickCells(self, zone=None):
        z = zone or self.tickZone
        _ck(self.L.fse_particles_tick(self.h, C.byref(z)))

    particles_tick = tickCells

    def particles_count(self):
        n = C.c_int64()
        _ck(self.L.fse_particles_count(self.h, C.byref(n)))
        return n.value

    def particles_read(self):
        n = self.particles_count()
        out = np.zeros(n, dtype=T.PARTICLE_DTYPE)
        m = C.c_int64()
        if n:
            _ck(self.L.fse_particles_read(self.h, out.ctypes.data_as(C.c_void_p), n, C.byref(m)))
        return out

    def particles_dropped(self):
        n = C.c_int64()
        self.L.fse_particles_dropped.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        _ck(self.L.fse_particles_dropped(self.h, C.byref(n)))
        return n.value

    def particles_clear(self):
        _ck(self.L.fse_particles_clear(self.h))

    def particles_reserve(self, cap):
        _ck(self.L.fse_particles_reserve(self.h, cap))

    # -- rigid-body bridge (game.cpp:1711-1815, 1896-1983) ---------------------------------------------------
    def bodies_upload(self, bodies):
        """bodies: list of (h, w) CELL_DTYPE arrays (RigidBody::tiles, index [ty, tx])."""
        self._bodies = [np.ascontiguousarray(b, dtype=T.CELL_DTYPE) for b in bodies]
        n = len(self._bodies)
        descs = (T.BodyDesc * max(n, 1))()
        for i, b in enumerate(self._bodies):
            descs[i].w, descs[i].h = b.shape[1], b.shape[0]
            descs[i].tiles = b.ctypes.data
        self.L.fse_bodies_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        _ck(self.L.fse_bodies_upload(self.h, descs, n))

    @staticmethod
    def _xf(xforms):
        xf = np.ascontiguousarray(xforms, dtype=np.float32).reshape(-1, 3)
        return xf, len(xf)

    def bodies_raster(self, xforms, tick=None, seed=1337):
        """tick keys the ids and velocities of the particles the bodies throw up; without one a per-world call counter is used,
        so that repeated calls never reissue an id."""
        if tick is None:
            self._raster_calls = getattr(self, "_raster_calls", 0) + 1
            tick = self._raster_calls
        xf, n = self._xf(xforms)
        fb = np.zeros((n, 4), dtype=np.int32)
        self.L.fse_bodies_raster.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p]
        _ck(self.L.fse_bodies_raster(self.h, xf.ctypes.data, n, tick, seed, fb.ctypes.data))
        return fb

    def bodies_erase(self, xforms):
        xf, n = self._xf(xforms)
        fb = np.zeros((n, 4), dtype=np.int32)
        need = np.zeros(n, dtype=np.uint8)
        self.L.fse_bodies_erase.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _ck(self.L.fse_bodies_erase(self.h, xf.ctypes.data, n, fb.ctypes.data, need.ctypes.data))
        return fb, need

    def bodies_split(self, i, angle=0.0, weld=(-1, -1)):
        """Fracture hand-off of uploaded body i (world::updateRigidBodyHitbox, world.cpp:288-720): one (piece record, tile array) per
        4-connected component of its non-AIR tiles, cropped, with the weld flag and the rotated position shift."""
        h, w = self._bodies[i].shape
        pieces = np.zeros(1024, dtype=T.BODY_PIECE_DTYPE)
        out = np.zeros(4 * h * w + 64, dtype=T.CELL_DTYPE)
        n = C.c_int32()
        self.L.fse_bodies_split.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                            C.c_int64]
        _ck(self.L.fse_bodies_split(self.h, i, angle, weld[0], weld[1], pieces.ctypes.data, len(pieces), C.byref(n), out.ctypes.data, len(out)))
        return [(pieces[k].copy(), out[pieces[k]["tile_off"]: pieces[k]["tile_off"] + pieces[k]["w"] * pieces[k]["h"]].reshape(pieces[k]["h"], pieces[k]["w"]).copy())
                for k in range(n.value)]

    def update_rigid_body_hitbox(self, i, angle=0.0, weld=(-1, -1)):
        """world::updateRigidBodyHitbox (world.cpp:288-720) for uploaded body i: the pieces (device), their outlines (device) and the
        triangle groups of their colliders (host: hitbox_triangles).  Returns a list of (piece record, tiles, [triangles (k, 3, 2)])."""
        out = []
        for rec, tiles in self.bodies_split(i, angle, weld):
            mask = (tiles["mat"] != 0).astype(np.uint8)[None]
            _, _, contours = self.mask_outline(mask, want_labels=False)
            out.append((rec, tiles, hitbox_triangles(contours[0])))
        return out

    def bodies_read(self, i):
        out = np.zeros(self._bodies[i].shape, dtype=T.CELL_DTYPE)
        self.L.fse_bodies_read.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _ck(self.L.fse_bodies_read(self.h, i, out.ctypes.data))
        return out

    # -- fracture outlines (world.cpp:288-720, physics_math.cpp:1766-1965) and physicsCheck flood (world.cpp:3330) ------
    def physics_check(self, x, y, cap_tiles=None):
        """world::physicsCheck (world.cpp:3330-3411): returns (count, action, (x, y, w, h), tiles or None) — action 0 nothing, 1 the
        1..10-cell crumb was deleted, 2 the 11..1000-cell component was cut out into `tiles` (h x w fse_cell array of the new body).
        cap_tiles: capacity of the tile buffer (default: 65536, retried with the component's box when that is too small — the call
        changes nothing when the buffer does not fit)."""
        class Res(C.Structure):
            _fields_ = [(n, C.c_int32) for n in ("count", "action", "x", "y", "w", "h")]
        res = Res()
        self.L.fse_physics_check.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        cap = cap_tiles or (1 << 16)
        tiles = np.zeros(cap, dtype=T.CELL_DTYPE)
        rc = self.L.fse_physics_check(self.h, x, y, C.byref(res), tiles.ctypes.data, cap)
        if rc != 0 and cap_tiles is None and 10 < res.count <= 1000 and res.w * res.h > cap:
            cap = res.w * res.h
            tiles = np.zeros(cap, dtype=T.CELL_DTYPE)
            rc = self.L.fse_physics_check(self.h, x, y, C.byref(res), tiles.ctypes.data, cap)
        _ck(rc)
        box = (res.x, res.y, res.w, res.h)
        return res.count, res.action, box, (tiles[:res.w * res.h].reshape(res.h, res.w).copy() if res.action == 2 else None)

    def probe_position(self, tick, seed=1337, zone=None):
        z = zone or self.tickZone
        x, y = C.c_int32(0), C.c_int32(0)
        self.L.fse_probe_position.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.fse_probe_position.restype = None
        self.L.fse_probe_position(seed, tick, C.byref(z), C.byref(x), C.byref(y))
        return x.value, y.value

    def physics_probe(self, tick, seed=1337, zone=None):
        """The probe at the end of world::tick (world.cpp:1929-1934): physicsCheck at a random cell of the tickZone."""
        x, y = self.probe_position(tick, seed, zone)
        return self.physics_check(x, y)

    # -- chunk files (Chunk::ChunkRead / ChunkWrite, chunk.cpp:74-330; the merge of world::frame and chunkSaveCache) ---------
    def load_chunk(self, path, x, y, layers=False):
        """Read a .pack file and merge it into the grid at (x, y), marked dirty like world::frame does (world.cpp:2374-2391): the
        object layer always, layer 2 and the background colours into their device planes with layers=True.  Cells must index into
        the context's material table (IOError otherwise).  Returns (generation_phase, layer2, background)."""
        from . import chunkfile

        n = self.ctx.table.n if getattr(self.ctx, "table", None) is not None else None
        phase, tiles, layer2, background = chunkfile.read_pack(path, n)
        tiles["dirty"] = 1
        self.write_rect(x, y, tiles)
        if layers:
            self.layer2_write_rect(x, y, layer2)
            self.background_write_rect(x, y, background)
        return phase, layer2, background

    def save_chunk(self, path, x, y, layer2=None, background=None, generation_phase=0, layers=False):
        """chunkSaveCache (world.cpp:2780-2792) + ChunkWrite: the 128 x 128 cells at (x, y) go to a .pack file (layers=True: with
        the device's layer-2 cells and background colours unless given)."""
        from . import chunkfile

        if layers and layer2 is None:
            layer2 = self.layer2_read_rect(x, y, T.FSE_CHUNK, T.FSE_CHUNK)
        if layers and background is None:
            background = self.background_read_rect(x, y, T.FSE_CHUNK, T.FSE_CHUNK)
        chunkfile.write_pack(path, self.read_rect(x, y, T.FSE_CHUNK, T.FSE_CHUNK), layer2, background, generation_phase)

    def save_world(self, world_dir, origin=(0, 0)):
        """world::saveWorld (world.cpp:3431): all chunks of the grid to <world_dir>/chunks/c_<x>_<y>.pack."""
        from . import chunkfile

        return chunkfile.save_world(self, world_dir, self.width, self.height, origin)

    def load_world(self, world_dir, origin=(0, 0)):
        """Merge every chunk file found under <world_dir>/chunks into the grid."""
        from . import chunkfile

        n = self.ctx.table.n if getattr(self.ctx, "table", None) is not None else None
        return chunkfile.load_world(self, world_dir, self.width, self.height, origin, n)

    # -- render planes / camera scroll (game.cpp:1994-2060, world.cpp:2454-2478) ---------------------------
    def pixels_enable(self, on=True):
        self.L.fse_pixels_enable.argtypes = [C.c_void_p, C.c_int]
        _ck(self.L.fse_pixels_enable(self.h, 1 if on else 0))

    def render_dirty(self, want_stats=True, with_flow_count=False):
        """Refresh the texels of all dirty cells; returns (dirty, fire, movingTiles) [+ dirty liquid cells with with_flow_count] or None
        when want_stats is False."""
        self.L.fse_render_dirty.argtypes = [C.c_void_p, C.c_void_p]
        if not want_stats:
            _ck(self.L.fse_render_dirty(self.h, None))
            return None
        st = T.RenderStats()
        _ck(self.L.fse_render_dirty(self.h, C.byref(st)))
        if with_flow_count:
            return st.dirty, st.fire, np.ctypeslib.as_array(st.moving).copy(), st.flow
        return st.dirty, st.fire, np.ctypeslib.as_array(st.moving).copy()

    # -- liquid flow accumulators (world::flowX / flowY / prevFlowX / prevFlowY) and the flow texture --------------------------
    def flow_enable(self, on=True):
        self.L.fse_flow_enable.argtypes = [C.c_void_p, C.c_int]
        _ck(self.L.fse_flow_enable(self.h, 1 if on else 0))

    def flow_read(self, which, rect=None):
        """which: 0 flowX, 1 flowY, 2 prevFlowX, 3 prevFlowY."""
        r = rect or T.Rect(0, 0, self.width, self.height)
        out = np.zeros((r.h, r.w), dtype=np.float32)
        self.L.fse_flow_read.argtypes = [C.c_void_p, C.c_int] + [C.c_int32] * 4 + [C.c_void_p]
        _ck(self.L.fse_flow_read(self.h, which, r.x, r.y, r.w, r.h, out.ctypes.data))
        return out

    # -- second cell layer and background colours (world::real_layer2 / background) ---------------------------------------------
    def layer2_write_rect(self, x, y, cells):
        c = np.ascontiguousarray(cells, dtype=T.CELL_DTYPE)
        self.L.fse_layer2_write_rect.argtypes = [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p]
        _ck(self.L.fse_layer2_write_rect(self.h, x, y, c.shape[1], c.shape[0], c.ctypes.data))

    def layer2_read_rect(self, x, y, w, h):
        out = np.zeros((h, w), dtype=T.CELL_DTYPE)
        self.L.fse_layer2_read_rect.argtypes = [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p]
        _ck(self.L.fse_layer2_read_rect(self.h, x, y, w, h, out.ctypes.data))
        return out

    def background_write_rect(self, x, y, colors):
        c = np.ascontiguousarray(colors, dtype=np.uint32)
        self.L.fse_background_write_rect.argtypes = [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p]
        _ck(self.L.fse_background_write_rect(self.h, x, y, c.shape[1], c.shape[0], c.ctypes.data))

    def background_read_rect(self, x, y, w, h):
        out = np.zeros((h, w), dtype=np.uint32)
        self.L.fse_background_read_rect.argtypes = [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p]
        _ck(self.L.fse_background_read_rect(self.h, x, y, w, h, out.ctypes.data))
        return out

    def render_layers(self, draw_background_grid=False, want_counts=True):
        """game.cpp:2068-2126: dirty layer-2 / background cells -> FSE_PIXELS_LAYER2 / FSE_PIXELS_BACKGROUND; returns the two counts."""
        self.L.fse_render_layers.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        if not want_counts:
            _ck(self.L.fse_render_layers(self.h, 1 if draw_background_grid else 0, None, None))
            return None
        a, b = C.c_int64(0), C.c_int64(0)
        _ck(self.L.fse_render_layers(self.h, 1 if draw_background_grid else 0, C.byref(a), C.byref(b)))
        return a.value, b.value

    def pixels_read(self, which, rect=None):
        r = rect or T.Rect(0, 0, self.width, self.height)
        out = np.zeros((r.h, r.w, 4), dtype=np.uint8)
        self.L.fse_pixels_read.argtypes = [C.c_void_p, C.c_int] + [C.c_int32] * 4 + [C.c_void_p]
        _ck(self.L.fse_pixels_read(self.h, which, r.x, r.y, r.w, r.h, out.ctypes.data))
        return out

    def scroll(self, dx, dy):
        """world::tickChunks: shift the grid and the loose particles by (dx, dy)."""
        self.L.fse_scroll.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        _ck(self.L.fse_scroll(self.h, dx, dy))

    def explosion(self, x, y, radius, tick=0, seed=1337):
        """world::explosion(x, y, radius) (world.cpp:2294-2332)."""
        self.L.fse_explosion.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_uint32]
        _ck(self.L.fse_explosion(self.h, x, y, radius, tick, seed))

    # -- entities <-> grid (world::tickEntities, WorldEntitySystem::process, objectDelete) ----------------------
    def entities_tick(self, ents, load_zone=(0.0, 0.0), tick=0, seed=1337):
        """world::tickEntities (world.cpp:3010-3247) on a T.ENTITY_DTYPE array; returns the updated copy (x, y, vx, vy, ground, destroy)."""
        e = np.ascontiguousarray(ents, dtype=T.ENTITY_DTYPE).copy()
        self.L.fse_entities_tick.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_uint32, C.c_uint32]
        _ck(self.L.fse_entities_tick(self.h, e.ctypes.data, len(e), load_zone[0], load_zone[1], tick, seed))
        return e

    def entities_stamp(self, ents, load_zone=(0.0, 0.0), object_mat=6, tick=0, seed=1337):
        """WorldEntitySystem::process (game/player.cpp:173-199): cells under the entities become Tiles_OBJECT until object_delete()."""
        e = np.ascontiguousarray(ents, dtype=T.ENTITY_DTYPE)
        self.L.fse_entities_stamp.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_uint32, C.c_uint32]
        _ck(self.L.fse_entities_stamp(self.h, e.ctypes.data, len(e), load_zone[0], load_zone[1], object_mat, tick, seed))

    def object_delete(self):
        self.L.fse_object_delete.argtypes = [C.c_void_p]
        _ck(self.L.fse_object_delete(self.h))

    # -- interactive tools, grid side -----------------------------------------------------------------------
    def tool_erase_line(self, x0, y0, x1, y1, brush_size=5):
        """middle-mouse erase brush (game.cpp:593-625)."""
        self.L.fse_tool_erase_line.argtypes = [C.c_void_p] + [C.c_int32] * 5
        _ck(self.L.fse_tool_erase_line(self.h, x0, y0, x1, y1, brush_size))

    def tool_pickaxe(self, x, y, break_size):
        """pickaxe (game.cpp:771-790): returns (ARGB pixels (size, size) of the broken-off body, cells taken)."""
        size = int(break_size)
        pix = np.zeros((size, size), dtype=np.uint32)
        n = C.c_int32()
        self.L.fse_tool_pickaxe.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.POINTER(C.c_int32)]
        _ck(self.L.fse_tool_pickaxe(self.h, x, y, break_size, pix.ctypes.data, C.byref(n)))
        return pix, n.value

    def tool_hammer(self, hammer_x, hammer_y, x, y, sand_mat=2, tick=0, seed=1337):
        """hammer release (game.cpp:843-890): returns (end_x, end_y, n_changed, broke)."""
        out = np.zeros(4, dtype=np.int32)
        self.L.fse_tool_hammer.argtypes = [C.c_void_p] + [C.c_int32] * 5 + [C.c_uint32, C.c_uint32, C.c_void_p]
        _ck(self.L.fse_tool_hammer(self.h, hammer_x, hammer_y, x, y, sand_mat, tick, seed, out.ctypes.data))
        return tuple(int(v) for v in out)

    def tool_vacuum(self, wcx, wcy, wmx, wmy, tick=0, seed=1337):
        """vacuum (game.cpp:2456-2585): returns (x, y, cells sucked, particles caught)."""
        out = np.zeros(4, dtype=np.int32)
        self.L.fse_tool_vacuum.argtypes = [C.c_void_p] + [C.c_int32] * 4 + [C.c_uint32, C.c_uint32, C.c_void_p]
        _ck(self.L.fse_tool_vacuum(self.h, wcx, wcy, wmx, wmy, tick, seed, out.ctypes.data))
        return tuple(int(v) for v in out)

    def particles_vacuum_pull(self, target_x, target_y):
        n = C.c_int32()
        self.L.fse_particles_vacuum_pull.argtypes = [C.c_void_p, C.c_float, C.c_float, C.POINTER(C.c_int32)]
        _ck(self.L.fse_particles_vacuum_pull(self.h, target_x, target_y, C.byref(n)))
        return n.value

    def mask_outline(self, masks, want_labels=True, as_lists=True):
        """masks: (n, h, w) uint8.  Returns (labels (n,h,w) int32 or None, n_components (n,), contours: list per mask of (k,2) float arrays —
        or, with as_lists=False, the flat (points, point offsets, mask offsets) arrays of the C ABI)."""
        masks = np.ascontiguousarray(masks, dtype=np.uint8)
        n, h, w = masks.shape
        labels = np.zeros((n, h, w), dtype=np.int32) if want_labels else None
        ncomp = np.zeros(n, dtype=np.int32)
        cap_pts, cap_c = int(masks.size) * 2 + 64, int(masks.size) // 2 + 64
        pts = np.zeros((cap_pts, 2), dtype=np.float32)
        pt_off = np.zeros(cap_c + 1, dtype=np.int32)
        mask_off = np.zeros(n + 1, dtype=np.int32)
        self.L.fse_mask_outline.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        t0 = time.perf_counter()
        _ck(self.L.fse_mask_outline(self.h, masks.ctypes.data, n, w, h, labels.ctypes.data if want_labels else None, ncomp.ctypes.data, pts.ctypes.data,
                                    cap_pts, pt_off.ctypes.data, cap_c, mask_off.ctypes.data))
        self.last_outline_s = time.perf_counter() - t0  # the C-ABI call alone
        if not as_lists:
            return labels, ncomp, (pts, pt_off, mask_off)
        contours = [[pts[pt_off[c]:pt_off[c + 1]].copy() for c in range(mask_off[m], mask_off[m + 1])] for m in range(n)]
        return labels, ncomp, contours

    def solid_mask(self, x, y, w, h):
        out = np.zeros((h, w), dtype=np.uint8)
        self.L.fse_solid_mask.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        _ck(self.L.fse_solid_mask(self.h, x, y, w, h, out.ctypes.data))
        return out

    def flood_component(self, x, y, cap=1000):
        """physicsCheck(x, y): (count, bbox, pixel indices); count == cap + 1 means 'larger than cap'."""
        cnt = C.c_int32()
        bbox = np.zeros(4, dtype=np.int32)
        pix = np.zeros(cap + 1, dtype=np.int32)
        self.L.fse_flood_component.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]
        _ck(self.L.fse_flood_component(self.h, x, y, cap, C.byref(cnt), bbox.ctypes.data, pix.ctypes.data))
        n = cnt.value
        return n, bbox, pix[:n] if n <= cap else pix[:0]

    # -- active-region tracking (world::active / lastActive, world.hpp:131-133) ---------------------------
    def active_enable(self, on=True):
        self.L.fse_active_enable.argtypes = [C.c_void_p, C.c_int]
        _ck(self.L.fse_active_enable(self.h, 1 if on else 0))

    def active_stats(self):
        self.L.fse_active_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        a, t = C.c_int64(), C.c_int64()
        _ck(self.L.fse_active_stats(self.h, C.byref(a), C.byref(t)))
        return a.value, t.value

    # -- measurement ----------------------------------------------------------------------------------
    def timer_start(self):
        _ck(self.L.fse_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        _ck(self.L.fse_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def kernel_timing(self, enable):
        _ck(self.L.fse_kernel_timing_enable(self.h, 1 if enable else 0))

    def kernel_timing_phases(self, cap=1 << 16):
        """ms of every timed colour phase since kernel_timing(True), in launch order (call before kernel_timing_read)."""
        out = np.zeros(cap, dtype=np.float32)
        n = C.c_int64()
        self.L.fse_kernel_timing_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        _ck(self.L.fse_kernel_timing_phases(self.h, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value]

    def strip_timeline_read(self, cap=1 << 14):
        out = np.zeros((cap, 3), dtype=np.float32)
        n = C.c_int64()
        self.L.fse_strip_timeline_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        _ck(self.L.fse_strip_timeline_read(self.h, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value]

    def kernel_timing_read(self):
        tot, n = C.c_double(), C.c_int64()
        _ck(self.L.fse_kernel_timing_read(self.h, C.byref(tot), C.byref(n)))
        return tot.value, n.value


from .materials import default_materials  # noqa: E402,F401  (InitMaterials, gds.cpp:117-280)

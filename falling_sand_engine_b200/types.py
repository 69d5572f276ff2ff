"""POD types of include/fse.h as ctypes structures and numpy dtypes.

Each structure mirrors the C declaration field for field; `tests/test_abi.py` checks
the sizes against the compiled library (`fse_abi_sizeof`).
"""
import ctypes as C

import numpy as np

FSE_CHUNK = 128
FSE_MAX_MATERIALS = 256
FSE_MAX_REACH = 5

AIR, SOLID, SAND, SOUP, GAS, PASSABLE, OBJECT = 0, 1, 2, 3, 4, 5, 5
INTERACT_NONE, INTERACT_TRANSFORM_MATERIAL, INTERACT_SPAWN_MATERIAL, EXPLODE = 0, 1, 2, 3
REACT_TEMPERATURE_BELOW, REACT_TEMPERATURE_ABOVE = 4, 5
COLOR_FIXED, COLOR_JITTER, COLOR_POSITIONAL = 0, 1, 2


class Material(C.Structure):
    _fields_ = [
        ("physics", C.c_int32),
        ("density", C.c_float),
        ("iterations", C.c_int32),
        ("slipperyness", C.c_int32),
        ("emit", C.c_int32),
        ("emit_color", C.c_uint32),
        ("color", C.c_uint32),
        ("add_temp", C.c_uint32),
        ("conduction_self", C.c_float),
        ("conduction_other", C.c_float),
        ("create_temp", C.c_int16),
        ("alpha", C.c_uint8),
        ("interact", C.c_uint8),
        ("react", C.c_uint8),
        ("color_kind", C.c_uint8),
        ("jitter_shift", C.c_uint8),
        ("jitter_range", C.c_uint8),
    ]


class Interaction(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("data1", C.c_int16),
        ("_pad", C.c_uint16),
        ("data2", C.c_uint32),
        ("ofs_x", C.c_int32),
        ("ofs_y", C.c_int32),
    ]


class SpecialIds(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("air", "fire", "water", "lava", "steam", "obsidian")]


class Rect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32)]


class BodyDesc(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("tiles", C.c_void_p)]


class TickArgs(C.Structure):
    _fields_ = [("tick", C.c_uint32), ("seed", C.c_uint32), ("cell_iter", C.c_int32), ("tick_zone", Rect)]


class Stats(C.Structure):
    _fields_ = [
        ("hash", C.c_uint64),
        ("count", C.c_uint64 * FSE_MAX_MATERIALS),
        ("fluid_mass", C.c_double * FSE_MAX_MATERIALS),
        ("n_dirty", C.c_uint64),
        ("n_moved", C.c_uint64),
    ]


class RenderStats(C.Structure):
    _fields_ = [("dirty", C.c_int64), ("fire", C.c_int64), ("moving", C.c_int64 * FSE_MAX_MATERIALS), ("flow", C.c_int64)]


# fse_cell (20 bytes)
CELL_DTYPE = np.dtype(
    {
        "names": ["mat", "moved", "settle", "color", "temp", "dirty", "_pad", "fluid", "fluid_diff"],
        "formats": ["<u2", "u1", "u1", "<u4", "<i2", "u1", "u1", "<f4", "<f4"],
        "offsets": [0, 2, 3, 4, 8, 10, 11, 12, 16],
        "itemsize": 20,
    }
)

# fse_particle (80 bytes)
PARTICLE_DTYPE = np.dtype(
    {
        "names": ["tile", "x", "y", "vx", "vy", "ax", "ay", "target_x", "target_y", "target_force", "lifetime", "fade_time",
                  "phase", "temporary", "in_object_state", "vacuum", "_pad2", "id"],
        "formats": [CELL_DTYPE, "<f4", "<f4", "<f4", "<f4", "<f4", "<f4", "<f4", "<f4", "<f4", "<i4", "<i4", "u1", "u1", "u1", "u1",
                    "<u4", "<u8"],
        "offsets": [0, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 60, 64, 65, 66, 67, 68, 72],
        "itemsize": 80,
    }
)


# fse_entity (32 bytes): the value fields of WorldEntity (game_datastruct.hpp:42-62)
ENTITY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("hw", "<i4"), ("hh", "<i4"), ("ground", "<i4"), ("destroy", "<i4")])


# fse_body_piece (36 bytes)
BODY_PIECE_DTYPE = np.dtype([("x0", "<i4"), ("y0", "<i4"), ("w", "<i4"), ("h", "<i4"), ("n_pixels", "<i4"), ("weld", "<i4"), ("tile_off", "<i4"),
                             ("shift_x", "<f4"), ("shift_y", "<f4")])


def zone_of(width, height):
    """tickZone of the reference: the grid minus a one-chunk border (game.cpp:1629)."""
    return Rect(FSE_CHUNK, FSE_CHUNK, width - 2 * FSE_CHUNK, height - 2 * FSE_CHUNK)


class MaterialTable:
    """A flattened material table (the argument list of fse_materials_set)."""

    def __init__(self, mats, ids, inter, inter_offsets, react, react_offsets):
        self.mats = mats  # ctypes array of Material
        self.n = len(mats)
        self.ids = ids
        self.inter = inter
        self.inter_offsets = inter_offsets
        self.react = react
        self.react_offsets = react_offsets

    def physics(self):
        return np.array([m.physics for m in self.mats], dtype=np.int32)

    def args(self):
        return (
            self.mats,
            C.c_int(self.n),
            C.byref(self.ids),
            self.inter,
            self.inter_offsets,
            self.react,
            self.react_offsets,
        )

    def dump(self, path):
        """Raw dump for a C++ host (csrc/host_demo.cpp): per array an int64 count and the structs; the special ids in between."""
        import struct

        with open(path, "wb") as f:
            def arr(a, n):
                f.write(struct.pack("<q", n))
                f.write(bytes(a)[: n * (C.sizeof(a) // max(len(a), 1))])
            arr(self.mats, self.n)
            f.write(bytes(self.ids))
            ni = self.inter_offsets[self.n * self.n] if len(self.inter_offsets) else 0
            nr = self.react_offsets[self.n] if len(self.react_offsets) else 0
            arr(self.inter, ni)
            arr(self.inter_offsets, len(self.inter_offsets))
            arr(self.react, nr)
            arr(self.react_offsets, len(self.react_offsets))

    def copy(self):
        mats = (Material * self.n)(*[Material.from_buffer_copy(bytes(m)) for m in self.mats])
        ids = SpecialIds.from_buffer_copy(bytes(self.ids))
        ni, nr = len(self.inter), len(self.react)
        inter = (Interaction * max(ni, 1))(*[Interaction.from_buffer_copy(bytes(i)) for i in self.inter])
        react = (Interaction * max(nr, 1))(*[Interaction.from_buffer_copy(bytes(i)) for i in self.react])
        io = (C.c_int32 * len(self.inter_offsets))(*self.inter_offsets)
        ro = (C.c_int32 * len(self.react_offsets))(*self.react_offsets)
        return MaterialTable(mats, ids, inter, io, react, ro)

    def with_interactions(self, pairs):
        """Return a copy whose interaction lists are replaced by `pairs`:
        {(a, b): [(type, product, radius, ofs_x, ofs_y), ...]}; sets Material.interact on every `a`."""
        t = self.copy()
        n = t.n
        flat, offs = [], [0]
        for a in range(n):
            for b in range(n):
                for (ty, prod, rad, ox, oy) in pairs.get((a, b), []):
                    flat.append(Interaction(ty, prod, 0, rad, ox, oy))
                offs.append(len(flat))
        for (a, _b) in pairs:
            t.mats[a].interact = 1
        t.inter = (Interaction * max(len(flat), 1))(*flat)
        t.inter_offsets = (C.c_int32 * len(offs))(*offs)
        return t

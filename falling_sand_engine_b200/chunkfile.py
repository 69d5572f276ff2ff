"""Chunk files of the reference (`Chunk::ChunkRead / ChunkWrite`, source/engine/chunk.cpp:74-330): the host side of world load /
save.  A `.pack` file is

    i8   generationPhase
    i32  src_size   (= 128 * 128 * 2 * 12)     i32 compressed_size
    i32  src_size2  (= 128 * 128 * 4)          i32 compressed_size2
    LZ4 block: MaterialInstanceData[2 * 128 * 128] = {u32 index, u32 color, i16 temperature} (12 bytes with padding,
               game_datastruct.hpp:106-107, chunk.hpp:26-30), first the 16384 cells of the object layer, then layer 2
    LZ4 block: u32 background[128 * 128]

LZ4 and the file format stay on the host; the cells go to / come from the device through fse_write_rect /
fse_read_rect (`World.load_chunk / save_chunk`).  The codec is the system's liblz4 (1.9.4, the version the reference vendors),
called like the reference does: LZ4_compress_fast(acceleration 10) / LZ4_decompress_safe.  A loaded cell gets the fields the
reference restores (material, colour, temperature) and MaterialInstance's defaults for the rest (fluidAmount 2.0, not moved)."""
import ctypes as C
import ctypes.util
import struct

import numpy as np

from . import types as T

CHUNK = T.FSE_CHUNK
CELLS = CHUNK * CHUNK
DISK_DTYPE = np.dtype({"names": ["index", "color", "temperature"], "formats": ["<u4", "<u4", "<i2"], "offsets": [0, 4, 8], "itemsize": 12})

_lz4 = None


def _lib():
    global _lz4
    if _lz4 is None:
        name = ctypes.util.find_library("lz4") or "liblz4.so.1"
        L = C.CDLL(name)
        L.LZ4_compressBound.argtypes = [C.c_int]
        L.LZ4_compress_fast.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.LZ4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        _lz4 = L
    return _lz4


def _compress(raw):
    L = _lib()
    cap = L.LZ4_compressBound(len(raw))
    dst = C.create_string_buffer(cap)
    n = L.LZ4_compress_fast(raw, dst, len(raw), cap, 10)  # chunk.cpp:241 / 262
    if n <= 0:
        raise IOError(f"LZ4_compress_fast failed ({n})")
    return dst.raw[:n]


def _decompress(blob, size):
    dst = C.create_string_buffer(size)
    n = _lib().LZ4_decompress_safe(blob, dst, len(blob), size)
    if n != size:
        raise IOError(f"chunk data is corrupt: decompressed {n} bytes, expected {size}")  # chunk.cpp:153-157
    return dst.raw


def write_pack(path, tiles, layer2=None, background=None, generation_phase=0):
    """ChunkWrite (chunk.cpp:219-330).  tiles: (128, 128) fse_cell array (mat, color, temp are stored); layer2: the same or None
    (AIR); background: (128, 128) u32 or None."""
    buf = np.zeros(2 * CELLS, dtype=DISK_DTYPE)
    for k, layer in enumerate((tiles, layer2)):
        if layer is None:
            continue
        layer = np.asarray(layer).reshape(CELLS)
        part = buf[k * CELLS:(k + 1) * CELLS]
        part["index"], part["color"], part["temperature"] = layer["mat"], layer["color"], layer["temp"]
    bg = np.zeros(CELLS, dtype=np.uint32) if background is None else np.ascontiguousarray(background, dtype=np.uint32).reshape(CELLS)
    raw1, raw2 = buf.tobytes(), bg.tobytes()
    c1, c2 = _compress(raw1), _compress(raw2)
    with open(path, "wb") as f:
        f.write(struct.pack("<biiii", generation_phase, len(raw1), len(c1), len(raw2), len(c2)))
        f.write(c1)
        f.write(c2)


def read_pack(path, n_materials=None):
    """ChunkRead (chunk.cpp:74-217) -> (generation_phase, tiles, layer2, background) with tiles / layer2 as (128, 128) fse_cell
    arrays and background as (128, 128) u32.  n_materials: size of the material table the cells must index into (IOError otherwise)."""
    with open(path, "rb") as f:
        head = f.read(17)
        if len(head) != 17:
            raise IOError(f"{path}: truncated chunk header ({len(head)} of 17 bytes)")
        phase, src_size, csize, src_size2, csize2 = struct.unpack("<biiii", head)
        if src_size != 2 * CELLS * DISK_DTYPE.itemsize:
            raise IOError(f"Chunk src_size was different from expected: {src_size} vs {2 * CELLS * DISK_DTYPE.itemsize}")  # chunk.cpp:126
        if src_size2 != CELLS * 4:
            raise IOError(f"Chunk src_size2 was different from expected: {src_size2} vs {CELLS * 4}")  # chunk.cpp:139
        L = _lib()
        if not (0 < csize <= L.LZ4_compressBound(src_size)) or not (0 < csize2 <= L.LZ4_compressBound(src_size2)):
            raise IOError(f"{path}: compressed sizes {csize}, {csize2} are not those of a chunk file")
        c1, c2 = f.read(csize), f.read(csize2)
        if len(c1) != csize or len(c2) != csize2:
            raise IOError(f"{path}: truncated chunk data")
    disk = np.frombuffer(_decompress(c1, src_size), dtype=DISK_DTYPE)
    if n_materials is not None and int(disk["index"].max()) >= n_materials:
        raise IOError(f"{path}: material index {int(disk['index'].max())} is outside the material table ({n_materials} entries)")
    if int(disk["index"].max()) > 0xffff:
        raise IOError(f"{path}: material index {int(disk['index'].max())} does not fit fse_cell.mat")
    bg = np.frombuffer(_decompress(c2, src_size2), dtype=np.uint32).reshape(CHUNK, CHUNK).copy()
    layers = []
    for k in range(2):
        part = disk[k * CELLS:(k + 1) * CELLS]
        cells = np.zeros(CELLS, dtype=T.CELL_DTYPE)
        cells["mat"], cells["color"], cells["temp"] = part["index"], part["color"], part["temperature"]
        cells["fluid"] = 2.0  # MaterialInstance's default (game_datastruct.hpp:215); moved, settle and fluid_diff stay 0
        layers.append(cells.reshape(CHUNK, CHUNK))
    return phase, layers[0], layers[1], bg


def pack_path(world_dir, cx, cy):
    """Chunk::ChunkInit (chunk.cpp:20): <world>/chunks/c_<x>_<y>.pack, (x, y) in chunk units."""
    import os

    return os.path.join(world_dir, "chunks", f"c_{cx}_{cy}.pack")


def _whole_chunks(width, height, what):
    if width % CHUNK or height % CHUNK:
        raise ValueError(f"{what}: the grid must be whole chunks ({width} x {height} is not a multiple of {CHUNK})")


def save_world(world, world_dir, width, height, origin=(0, 0)):
    """world::saveWorld (world.cpp:3431-3452): every chunk of the grid goes to its .pack file (chunkSaveCache + ChunkWrite).
    `world` is anything with read_rect(x, y, w, h) (+ layer2_read_rect / background_read_rect when it keeps those planes); `origin` is
    the chunk coordinate of the grid's top-left chunk.  Returns the number of files written."""
    import os

    _whole_chunks(width, height, "save_world")
    os.makedirs(os.path.join(world_dir, "chunks"), exist_ok=True)
    layers = hasattr(world, "layer2_read_rect") and hasattr(world, "background_read_rect")
    n = 0
    for j in range(height // CHUNK):
        for i in range(width // CHUNK):
            l2 = world.layer2_read_rect(i * CHUNK, j * CHUNK, CHUNK, CHUNK) if layers else None
            bg = world.background_read_rect(i * CHUNK, j * CHUNK, CHUNK, CHUNK) if layers else None
            write_pack(pack_path(world_dir, origin[0] + i, origin[1] + j), world.read_rect(i * CHUNK, j * CHUNK, CHUNK, CHUNK), l2, bg)
            n += 1
    return n


def load_world(world, world_dir, width, height, origin=(0, 0), n_materials=None):
    """The reverse: every chunk file that exists is merged into the grid (ChunkRead + the frame() merge, world.cpp:2374-2391: cells,
    layer 2 and background, all marked dirty).  `world` is anything with write_rect(x, y, cells) (+ layer2_write_rect /
    background_write_rect).  Returns the number of chunks loaded."""
    import os

    _whole_chunks(width, height, "load_world")
    layers = hasattr(world, "layer2_write_rect") and hasattr(world, "background_write_rect")
    n = 0
    for j in range(height // CHUNK):
        for i in range(width // CHUNK):
            path = pack_path(world_dir, origin[0] + i, origin[1] + j)
            if not os.path.exists(path):
                continue
            _, tiles, layer2, bg = read_pack(path, n_materials)
            tiles["dirty"] = 1
            world.write_rect(i * CHUNK, j * CHUNK, tiles)
            if layers:
                world.layer2_write_rect(i * CHUNK, j * CHUNK, layer2)
                world.background_write_rect(i * CHUNK, j * CHUNK, bg)
            n += 1
    return n

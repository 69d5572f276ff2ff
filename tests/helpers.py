"""Shared helpers of the parity tests: scenario builders and exact comparison of oracle vs CUDA worlds."""
import functools

import numpy as np

from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G

FIELDS = ["mat", "moved", "settle", "color", "temp", "dirty", "fluid", "fluid_diff"]


def assert_cells_equal(a, b, what=""):
    for f in FIELDS:
        if f in ("fluid", "fluid_diff"):
            same = a[f].view(np.uint32) == b[f].view(np.uint32)  # bit-exact floats
        else:
            same = a[f] == b[f]
        if not same.all():
            ys, xs = np.nonzero(~same)
            k = 0
            raise AssertionError(f"{what}: field {f!r} differs at {len(ys)} cells; first (x={xs[k]}, y={ys[k]}): "
                                 f"{a[ys[k], xs[k]]} vs {b[ys[k], xs[k]]}")


def sort_particles(p):
    return p[np.argsort(p["id"], kind="stable")]


def assert_particles_equal(a, b, what=""):
    a, b = sort_particles(a), sort_particles(b)
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)} particles"
    assert a.tobytes() == b.tobytes(), f"{what}: particle records differ"


def build_mixed(world, table, W, H, seed=1337, extra=None, blob=32):
    G.fill_world(world, functools.partial(G.mixed_band, table, seed=seed, extra=extra, blob=blob), W, H, band_rows=512)


def build_column(world, table, W, H, seed=1337):
    G.fill_world(world, functools.partial(G.column_drop_band, table, seed=seed), W, H, band_rows=512)


def empty_world_cells(table, W, H, seed=1):
    """AIR interior, GENERIC_SOLID border."""
    mat = np.zeros((H, W), dtype=np.uint16)
    G.border_fill(mat, 0, 0, W, H, 1)
    return G.cells_from_mat(table, mat, 0, 0, seed)

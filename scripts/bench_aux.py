"""HBM-bound helper kernels around the tick on the 8192^2 mixed world (BASELINE configs[1] generator), device-timed with CUDA
events on the library's stream: algorithmic bytes per cell, achieved GB/s, fraction of the measured copy peak.
usage: python scripts/bench_aux.py [size] [reps]"""
import functools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import types as T  # noqa: E402
from falling_sand_engine_b200 import worldgen as G  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
peak, peak_src = bench.peaks()
table, extra = bench.make_table()
ctx = fse.Context(0, table)
w = fse.World(ctx, N, N)
w.particles_reserve(1 << 25)
G.fill_world(w, functools.partial(G.mixed_band, table, seed=1337, extra=list(extra.values())), N, N, band_rows=1024)
for t in range(3):
    w.tick(t)
w.pixels_enable(True)
w.sync()
cells = N * N
zone = (N - 256) * (N - 256)


def timed(fn, reps=REPS, before=None):
    ms = []
    for _ in range(reps):
        if before:
            before()
        w.sync()
        w.timer_start()
        fn()
        ms.append(w.timer_stop())
    return float(np.median(ms))


rows = []


def row(name, ms, bytes_moved, note):
    gbs = bytes_moved / (ms * 1e-3) / 1e9
    rows.append({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(bytes_moved / 1e6, 1), "GB/s": round(gbs, 1),
                 "frac_of_peak": round(gbs / peak, 3), "note": note})


# every cell dirty (the frame after a full-world load): 1 + 1 + 4 B read, 2 or 3 texels written; AIR cells skip the colour read
ms = timed(lambda: w.render_dirty(want_stats=False))
d, f, moving = w.render_dirty()
row("render_dirty_kernel", ms, cells * 1 + d * (1 + 4 + 8), f"{d} of {cells} cells dirty after a tick: flag plane + (mat + colour in, 2 texels out) per dirty cell")
w.clear_dirty()
ms = timed(lambda: w.render_dirty(want_stats=False))
row("render_dirty_kernel (clean world)", ms, cells * 1, "flag plane only")
ms = timed(lambda: w.clear_dirty())
row("clear_dirty_kernel", ms, cells * 2, "flag plane read + written")
ms = timed(lambda: w.tick_temperature())
row("temperature_kernel", ms, zone * (2 + 1 + 2) + (cells - zone) * 4, "temperature + material in, temperature out; the cells around the zone are copied, then the planes trade places")
ms = timed(lambda: w.scroll(128, 0))
row("fse_scroll (7 planes)", ms, cells * 17 * 2, "one pass: every plane read once and written once into the second plane set, which then becomes the world")
w.flow_enable(True)
w.tick(90)
ms = timed(lambda: w.render_dirty(want_stats=False), before=lambda: None)
row("render_dirty_kernel (+ flow texture)", ms, cells * 1 + d * (1 + 4 + 8), "same accounting as above; liquid cells add 16 B of accumulators in and 12 B out")
w.flow_enable(False)
l2 = np.zeros((1024, N), dtype=T.CELL_DTYPE)
l2["mat"], l2["color"] = 7, 0x808080
bg = np.full((1024, N), 0xFF102030, dtype=np.uint32)
def dirty_layers():
    for y0 in range(0, N, 1024):
        w.layer2_write_rect(0, y0, l2)
        w.background_write_rect(0, y0, bg)
ms = timed(lambda: w.render_layers(want_counts=False), reps=3, before=dirty_layers)
row("render_layers_kernel (all dirty)", ms, cells * (1 + 1 + 4 + 4 + 8 + 1), "dirty byte + layer-2 material and colour + background colour in, two texels + the cleared dirty byte out")
ms = timed(lambda: w.render_layers(want_counts=False))
row("render_layers_kernel (clean)", ms, cells * 1, "dirty plane only")
w.particles_clear()
w.tick(3)
w.particles_tick()  # first call allocates the scratch pools
tk = 4
for label in ("1 tick", "3 ticks"):
    for _ in range(1 if label == "1 tick" else 3):
        w.tick(tk)
        tk += 1
    n0 = w.particles_count()
    ms = timed(lambda: w.particles_tick(), reps=1)
    row(f"fse_particles_tick ({label} of spawns)", ms, n0 * 80 * 2,
        f"{n0} loose particles in, {w.particles_count()} out (80-byte records in and out; grid accesses are random; deposit rounds with a host sync each)")
print(json.dumps({"world": f"{N}x{N}", "peak_GB/s": peak, "peak_source": peak_src, "rows": rows}, indent=1))

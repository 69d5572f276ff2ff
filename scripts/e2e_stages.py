"""Where a whole game tick (bench.py's e2e leg) spends its wall-clock time: every stage followed by a stream sync, 8192^2 mixed world.
usage: python scripts/e2e_stages.py [size] [ticks]"""
import functools
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import types as T  # noqa: E402
from falling_sand_engine_b200 import worldgen as G  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
TICKS = int(sys.argv[2]) if len(sys.argv) > 2 else 12
table, extra = bench.make_table()
ctx = fse.Context(0, table)
w = fse.World(ctx, N, N)
w.particles_reserve(1 << 25)
G.fill_world(w, functools.partial(G.mixed_band, table, seed=1337, extra=list(extra.values())), N, N, band_rows=1024)
w.pixels_enable(True)
n_chunks = 16
pinned = torch.empty((n_chunks, 128, 128, T.CELL_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
src = G.mixed_band(table, N, N, 0, 128, seed=1337, extra=list(extra.values()))[:, :128]
pn = pinned.numpy()
for i in range(n_chunks):
    pn[i] = np.ascontiguousarray(src).view(np.uint8).reshape(128, 128, -1)
acc = {}


def stage(name, fn):
    t0 = time.perf_counter()
    fn()
    t1 = time.perf_counter()
    w.sync()
    t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0])
    a[0] += t1 - t0
    a[1] += t2 - t0


counts = []
for t in range(TICKS + 4):
    if t == 4:
        acc.clear()
    stage("chunk merges", lambda: [w.write_rect_ptr(0, 128 + 128 * i, 128, 128, pn[i].ctypes.data) for i in range(n_chunks)])
    stage("fse_tick", lambda: w.tick(t))
    counts.append(w.particles_count())
    stage("fse_particles_tick", lambda: w.particles_tick())
    if t % 4 == 2:
        stage("fse_tick_temperature", lambda: w.tick_temperature())
    stage("fse_render_dirty (+ stats)", lambda: w.render_dirty(want_stats=True))
    stage("fse_clear_dirty", lambda: w.clear_dirty())
out = {k: {"host_call_ms_per_tick": round(1e3 * v[0] / TICKS, 3), "with_sync_ms_per_tick": round(1e3 * v[1] / TICKS, 3)} for k, v in acc.items()}
out["total_ms_per_tick"] = round(sum(v[1] for v in acc.values()) * 1e3 / TICKS, 3)
out["particles_before_tickCells"] = counts
print(json.dumps(out, indent=1))

"""BASELINE configs[3]: 8192x8192 mixed world with 2000 rigid bodies — raster, tick, erase, CCL + outline of every body
each tick (wall clock around the C-ABI calls, results read back; B200 only)."""
import functools
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import worldgen as G  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
TICKS = int(sys.argv[3]) if len(sys.argv) > 3 else 10
table, extra = bench.make_table()
ctx = fse.Context(0, table)
w = fse.World(ctx, N, N)
w.particles_reserve(1 << 25)
G.fill_world(w, functools.partial(G.mixed_band, table, seed=1337, extra=list(extra.values())), N, N, band_rows=1024)
rng = np.random.default_rng(7)
bodies, masks = [], []
for b in range(NB):
    bw, bh = int(rng.integers(16, 33)), int(rng.integers(16, 33))
    hh = G.hash2(b + 1, np.arange(bw, dtype=np.uint32)[None, :], np.arange(bh, dtype=np.uint32)[:, None])
    m = np.where((hh % np.uint32(100)) < 80, 22, 0).astype(np.uint16)  # OBSIDIAN with hashed holes
    bodies.append(G.cells_from_mat(table, np.broadcast_to(m, (bh, bw)).copy(), 0, 0, b))
    mk = np.zeros((32, 32), dtype=np.uint8)
    mk[:bh, :bw] = m != 0
    masks.append(mk)
masks = np.stack(masks)
w.bodies_upload(bodies)
xf = np.stack([rng.uniform(200, N - 200, NB), rng.uniform(200, N - 200, NB), rng.uniform(-3.1, 3.1, NB)], axis=1).astype(np.float32)
for t in range(3):
    w.bodies_raster(xf, tick=t); w.tick(t); w.bodies_erase(xf); w.mask_outline(masks)
w.particles_clear()
w.sync()
acc = dict(raster=0.0, tick=0.0, erase=0.0, outline=0.0, outline_abi=0.0)
for t in range(3, 3 + TICKS):
    xf[:, 1] += 1.0
    xf[:, 2] += 0.02
    t0 = time.perf_counter(); w.bodies_raster(xf, tick=t); w.sync(); t1 = time.perf_counter()
    w.tick(t); w.sync(); t2 = time.perf_counter()
    w.bodies_erase(xf); w.sync(); t3 = time.perf_counter()
    labels, ncomp, contours = w.mask_outline(masks); t4 = time.perf_counter()
    acc["raster"] += t1 - t0; acc["tick"] += t2 - t1; acc["erase"] += t3 - t2; acc["outline"] += t4 - t3; acc["outline_abi"] += w.last_outline_s
px = sum(int((b["mat"] != 0).sum()) for b in bodies)
print(f"{N}x{N} mixed world, {NB} bodies ({px} pixels), {TICKS} ticks: ms per tick " + "  ".join(f"{k} {1e3 * v / TICKS:.2f}" for k, v in acc.items())
      + f"  total {1e3 * (sum(acc.values()) - acc['outline_abi']) / TICKS:.2f}  ({len(contours)} contours)")

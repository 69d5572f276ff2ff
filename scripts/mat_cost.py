"""Cost of one tick on uniform worlds: which material family dominates the tick kernels (B200 only)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import worldgen as G  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sched = int(sys.argv[2]) if len(sys.argv) > 2 else 1
table, extra = bench.make_table()
ids = G._names(table)
ctx = fse.Context(0, table)


def world_of(fn):
    w = fse.World(ctx, N, N)
    w.set_schedule(sched)
    w.particles_reserve(1 << 24)
    ys, xs = np.mgrid[0:N, 0:N].astype(np.uint32)
    mat = fn(xs, ys).astype(np.uint16)
    G.border_fill(mat, 0, 0, N, N, ids["GENERIC_SOLID"])
    for y0 in range(0, N, 1024):
        w.write_rect(0, y0, G.cells_from_mat(table, mat[y0:y0 + 1024], 0, y0, 7))
    return w


A, S, WAT, GAS, ST, F = ids["AIR"], ids["GENERIC_SAND"], ids["WATER"], ids["GENERIC_GAS"], ids["STONE"], ids["FIRE"]
cases = {
    "air": lambda x, y: np.full(x.shape, A),
    "stone": lambda x, y: np.full(x.shape, ST),
    "sand solid": lambda x, y: np.full(x.shape, S),
    "sand falling (1 in 4 rows)": lambda x, y: np.where(y % 4 == 0, S, A),
    "sand sparse (1 in 16 cells)": lambda x, y: np.where((x * 7 + y * 13) % 16 == 0, S, A),
    "water full": lambda x, y: np.full(x.shape, WAT),
    "water columns (every other col)": lambda x, y: np.where(x % 2 == 0, WAT, A),
    "water rain (1 in 16)": lambda x, y: np.where((x * 7 + y * 13) % 16 == 0, WAT, A),
    "gas sparse (1 in 16)": lambda x, y: np.where((x * 7 + y * 13) % 16 == 0, GAS, A),
    "fire on stone rows (every 8th row)": lambda x, y: np.where(y % 8 == 0, np.where(x % 2 == 0, F, ST), ST),
}
print(f"{N}x{N}, schedule {sched}")
for name, fn in cases.items():
    w = world_of(fn)
    w.tick(0)
    w.particles_clear()
    w.sync()
    w.kernel_timing(True)
    w.timer_start()
    for t in range(1, 4):
        w.tick(t)
    ms = w.timer_stop() / 3
    kt = w.kernel_timing_read()
    w.kernel_timing(False)
    cells = (N - 256) ** 2 * 3
    print(f"{name:36s} {ms:8.2f} ms/tick  {cells / ms / 1e6:7.2f} Gcell-updates/s   {kt}")
    w.particles_clear()
    w.close()

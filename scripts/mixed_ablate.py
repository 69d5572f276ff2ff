"""Tick cost of the mixed bench world with one material family replaced by AIR (B200 only): what the rule time is spent on."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import types as T  # noqa: E402
from falling_sand_engine_b200 import worldgen as G  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sched = int(sys.argv[2]) if len(sys.argv) > 2 else 1
table, extra = bench.make_table()
ids = G._names(table)
phys = np.array([m.physics for m in table.mats], dtype=np.int32)
ctx = fse.Context(0, table)
fam = {"none": [], "interacting powders": [-2], "plain powders": [-3], "gases": [T.GAS], "liquids": [T.SOUP], "fire": [-1], "everything": [T.SAND, T.SOUP, T.GAS, -1]}
xids = np.array(sorted(extra.values()))
print(f"{N}x{N} mixed, schedule {sched}; family replaced by AIR -> ms/tick (ticks 5..8)")
for name, kill in fam.items():
    w = fse.World(ctx, N, N)
    w.set_schedule(sched)
    w.particles_reserve(1 << 25)
    for y0 in range(0, N, 1024):
        cells = G.mixed_band(table, N, N, y0, 1024, seed=1337, extra=list(extra.values()))
        mat = cells["mat"].astype(np.uint16).copy()
        for k in kill:
            if k == -1:
                sel = mat == ids["FIRE"]
            elif k == -2:
                sel = np.isin(mat, xids)
            elif k == -3:
                sel = (phys[mat] == T.SAND) & ~np.isin(mat, xids)
            else:
                sel = (phys[mat] == k) & (mat != ids["FIRE"])
            mat[sel] = ids["AIR"]
        if kill:
            cells = G.cells_from_mat(table, mat, 0, y0, 1337)
        w.write_rect(0, y0, cells)
    for t in range(5):
        w.tick(t)
    w.particles_clear()
    w.sync()
    w.timer_start()
    for t in range(5, 9):
        w.tick(t)
    ms = w.timer_stop() / 4
    print(f"{name:20s} {ms:8.2f} ms/tick   {(N - 256) ** 2 * 3 / ms / 1e6:6.2f} Gcell-updates/s")
    w.close()

"""Small driver for ncu captures: N ticks of the mixed world (BASELINE configs[1] generator) at --size."""
import argparse
import functools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import worldgen as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=2048)
ap.add_argument("--ticks", type=int, default=4)
ap.add_argument("--workload", default="mixed")
a = ap.parse_args()
table, extra = bench.make_table()
ctx = fse.Context(0, table)
w = fse.World(ctx, a.size, a.size)
w.particles_reserve(1 << 22)
fn = {"mixed": functools.partial(G.mixed_band, table, seed=1337, extra=list(extra.values())),
      "column": functools.partial(G.column_drop_band, table, seed=1337),
      "sparse": functools.partial(G.sparse_band, table, seed=1337)}[a.workload]
G.fill_world(w, fn, a.size, a.size, band_rows=1024)
for t in range(a.ticks):
    w.tick(t)
w.sync()
print("done", w.stats().hash)

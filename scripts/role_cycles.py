"""Pass-1 phase accounting of the per-pass tick kernel on the mixed world (profiling aid; needs a build with
-DFSE_ROLE_CYCLES: `make -C falling_sand_engine_b200/csrc clean all EXTRA_NVCCFLAGS=-DFSE_ROLE_CYCLES`)."""
import ctypes as C, functools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import worldgen as G
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
workload = sys.argv[2] if len(sys.argv) > 2 else "mixed"
table, extra = bench.make_table()
ctx = fse.Context(0, table); w = fse.World(ctx, size, size); w.particles_reserve(1 << 25)
import numpy as np
ids = G._names(table)
if workload == "mixed":
    G.fill_world(w, functools.partial(G.mixed_band, table, seed=1337, extra=list(extra.values())), size, size, band_rows=1024)
elif workload != "air":
    ys, xs = np.mgrid[0:size, 0:size].astype(np.uint32)
    A, S, WAT, GAS = ids["AIR"], ids["GENERIC_SAND"], ids["WATER"], ids["GENERIC_GAS"]
    mat = {"water": lambda: np.full(xs.shape, WAT), "sand3": lambda: np.where(ys % 3 == 0, S, A), "waterhalf": lambda: np.where(xs % 2 == 0, WAT, A),
           "gas16": lambda: np.where((xs * 7 + ys * 13) % 16 == 0, GAS, A)}[workload]().astype(np.uint16)
    G.border_fill(mat, 0, 0, size, size, ids["GENERIC_SOLID"])
    for y0 in range(0, size, 1024):
        w.write_rect(0, y0, G.cells_from_mat(table, mat[y0:y0 + 1024], 0, y0, 7))
for t in range(2 if workload != 'mixed' else 5): w.tick(t)
w.L.fse_debug_role_cycles.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
out = (C.c_uint64 * 64)()
w.L.fse_debug_role_cycles(w.h, 1, None)
for t in range(5, 8): w.tick(t)
w.L.fse_debug_role_cycles(w.h, 1, out)
n = max(out[0], 1)
print("chunk passes", out[0])
print("per chunk: D+wait %.0f (warp 0 own D %.0f)  C1+bar %.0f  C2+bar %.0f  area %.0f cycles" % (out[1]/n, out[6]/n, out[2]/n, out[3]/n, out[4]/n))
print("rows per chunk: seen %.1f  active %.1f  with gather %.1f  with area %.1f" % (out[7]/n, out[8]/n, out[9]/n, out[10]/n))
for ps in (0, 1):
    for who, off in (("compute thread 0", 0), ("IO lane 0", 4)):
        o = out[16 + ps * 8 + off: 16 + ps * 8 + off + 4]
        m = max(o[3], 1)
        print("pass %d %-16s per chunk: mbarrier wait %.0f  step barrier %.0f  work %.0f cycles" % (ps + 1, who, o[0] / m, o[1] / m, o[2] / m))
for ps in (0, 1):
    m = max(out[16 + ps * 8 + 4 + 3], 1)
    o = out[32 + ps * 4: 32 + ps * 4 + 3]
    print("pass %d IO lane 0 per chunk: store side %.0f  wait for the old store to leave the slot %.0f  load issue %.0f cycles" % (ps + 1, o[0] / m, o[1] / m, o[2] / m))

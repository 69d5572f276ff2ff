"""Per-role cycle accounting of tick_rows_kernel on the mixed world (profiling aid)."""
import ctypes as C, functools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import worldgen as G
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
table, extra = bench.make_table()
ctx = fse.Context(0, table); w = fse.World(ctx, size, size); w.particles_reserve(1 << 23)
G.fill_world(w, functools.partial(G.mixed_band, table, seed=1337, extra=list(extra.values())), size, size, band_rows=1024)
for t in range(3): w.tick(t)
w.L.fse_debug_role_cycles.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
out = (C.c_uint64 * 12)()
w.L.fse_debug_role_cycles(w.h, 1, None)
for t in range(3, 6): w.tick(t)
w.L.fse_debug_role_cycles(w.h, 1, out)
n = out[4]
print("pass1 phases per chunk: D+bar %.0f (D own %.0f)  C1+bar %.0f  C2+bar %.0f  active rows %.1f" % (out[5]/n, out[10]/n, out[6]/n, out[7]/n, out[9]/n))
print("chunks", n, "avg cycles per chunk per role: pass1 %.0f pass2 %.0f pass3 %.0f io %.0f" % tuple(out[i] / max(n, 1) for i in range(4)))

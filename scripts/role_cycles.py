"""Pass-1 phase accounting of the per-pass tick kernel (profiling aid; needs a build with -DFSE_ROLE_CYCLES:
`scripts/build_variant.sh role -DFSE_ROLE_CYCLES`, then FSE_B200_LIB=_variants/libfse_role.so python scripts/role_cycles.py [size] [workload]).
Thread 0 of every pass-1 CTA adds up clock64() differences per phase of the row step (decide incl. the wait at the barrier, own-column
commit, gather + refunds, area effects) and counts the rows; run_pass adds them to a device buffer at the end of the pass."""
import ctypes as C, functools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import worldgen as G
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
workload = sys.argv[2] if len(sys.argv) > 2 else "mixed"
table, extra = bench.make_table()
ctx = fse.Context(0, table); w = fse.World(ctx, size, size); w.particles_reserve(1 << 25)
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
for t in range(2 if workload != "mixed" else 5): w.tick(t)
w.L.fse_debug_role_cycles.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
out = (C.c_uint64 * 64)()
w.L.fse_debug_role_cycles(w.h, 1, None)
w.sync(); w.timer_start()
for t in range(5, 8): w.tick(t)
ms = w.timer_stop() / 3
w.L.fse_debug_role_cycles(w.h, 1, out)
n = max(out[30], 1)
ph = [out[32 + q] / n for q in range(10)]
print(f"{size}^2 {workload}: {ms:.2f} ms / tick (instrumented build); pass-1 chunk passes {out[30]}; cycles per chunk pass {out[31] / n:.0f}")
print("per chunk pass, thread 0: decide + wait at the barrier %.0f (own decide %.0f), own-column commit + barrier %.0f, gather + refunds %.0f, area effects %.0f; outside the row step %.0f"
      % (ph[0], ph[5], ph[1], ph[2], ph[3], out[31] / n - ph[0] - ph[1] - ph[2] - ph[3]))
print("rows per chunk pass: stepped %.1f, with an active cell %.1f, with a gather phase %.1f, with area effects %.1f" % (ph[6], ph[7], ph[8], ph[9]))
act = max(ph[7], 1e-9)
print("per ACTIVE row: decide + wait %.0f, commit %.0f, gather (rows that have one: %.0f) , area (rows that have one: %.0f)"
      % (ph[0] / max(ph[6], 1e-9), ph[1] / act, ph[2] / max(ph[8], 1e-9), ph[3] / max(ph[9], 1e-9)))

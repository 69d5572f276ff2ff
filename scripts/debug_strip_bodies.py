"""Debug aid: run tests/strip_bodies_gpu_worker.py on N ranks and print where feedback / tiles / grid differ from the oracle."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O
from falling_sand_engine_b200 import strips, worldgen as G
from tests.strip_bodies_scene import scene
nranks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 6
W, H = 1024, 1536
out = "/tmp/sbdbg"
cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1", "--master-port", "29677",
       os.path.join(ROOT, "tests", "strip_bodies_gpu_worker.py"), str(W), str(H), str(ticks), out]
r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, FSE_TICK_MIN_CHUNKS="1", FSE_FUSED_MAX_CHUNKS="0"))
print("worker rc", r.returncode, r.stderr[-1500:] if r.returncode else "")
O.build()
table = O.default_materials(1337)
ow = O.OracleWorld(W, H, table)
ow.write_rect(0, 0, G.mixed_band(table, W, H, 0, H, seed=21, air_frac=0.6, blob=48))
bodies, xf = scene(table, W, H, nranks)
ob = [b.copy() for b in bodies]
fbs = []
for t in range(ticks):
    fbs.append(O.bodies_raster(ow, ob, xf, tick=t)); ow.tick(t, seed=1337); ow.particles_tick(); fbs.append(O.bodies_erase(ow, ob, xf))
    xf[:, 1] += 1.5; xf[:, 2] += 0.05
fbs = np.stack(fbs)
ref = ow.read_all()
for k in range(nranks):
    g = np.load(f"{out}.fb{k}.npy")
    bad = np.argwhere((g != fbs).any(axis=2))
    print(f"rank {k}: feedback mismatches (call, body):", bad[:20].tolist(), "of", len(bad))
    for c, b in bad[:6]:
        print("   call", c, "raster" if c % 2 == 0 else "erase", "body", b, "oracle", fbs[c, b].tolist(), "gpu", g[c, b].tolist())
    lo, hi = strips.strip_layout(H, k, nranks)[:2]
    cells = np.load(f"{out}.rank{k}.npy")
    diff = np.argwhere(cells["mat"] != ref[lo:hi]["mat"])
    print(f"rank {k}: rows {lo}..{hi}: {len(diff)} cells differ in material; first", (diff[:5] + [lo, 0]).tolist())
    tiles = np.concatenate([b.reshape(-1) for b in ob])
    gt = np.load(f"{out}.tiles{k}.npy")
    print(f"rank {k}: tiles equal:", tiles.tobytes() == gt.tobytes(), "differing pixels", int((tiles["mat"] != gt["mat"]).sum()))

"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck); no oracle in the loop.

  compute-sanitizer --tool racecheck python scripts/sanitize_small.py [per_pass|split|fused|aux]

Worlds are tiny on purpose (the tools slow kernels down 10-100x).  The summaries are kept under profiles/.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

what = sys.argv[1] if len(sys.argv) > 1 else "per_pass"
if what in ("per_pass", "split"):
    os.environ["FSE_FUSED_MAX_CHUNKS"] = "0"  # force tick_pass_kernel<1>/<2> + tick_pass3_kernel on a small world
    os.environ["FSE_TICK_MIN_CHUNKS"] = "1"
if what == "split":  # pass 2 split in every phase: pass-1 row masks + tick_pass2_apply_kernel + the row-skipping pass 2, no settled-row skipping
    os.environ["FSE_P2_SPLIT"] = "2"
    os.environ["FSE_ROW_SKIP"] = "0"

import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import materials as M  # noqa: E402
from falling_sand_engine_b200 import worldgen as G  # noqa: E402
from tests import helpers as Hh  # noqa: E402
from tests.test_bridge_cpu import make_body  # noqa: E402

base = M.default_materials(1337)
table, extra = G.bench_table(base)
ctx = fse.Context(0, table)
W, H = 640, 512
w = fse.World(ctx, W, H)
Hh.build_mixed(w, table, W, H, seed=1337, extra=list(extra.values()), blob=16)
if what in ("per_pass", "fused", "split"):
    w.set_schedule({"per_pass": 1, "fused": 2, "split": 1}[what])
    if what != "split":
        w.flow_enable(True)  # flowX / flowY reductions from pass 1
    for t in range(2):
        w.tick(t)
    if what == "per_pass":
        w.active_enable(True)
        w.tick(2)
else:
    w.tick(0)
    w.tick_temperature()
    w.particles_tick()
    w.pixels_enable(True)
    w.render_dirty(want_stats=True)
    w.clear_dirty()
    w.explosion(300, 260, 12, tick=1)
    w.scroll(-128, 0)
    w.scroll(5, -3)
    w.flow_enable(True)
    w.tick(1)
    w.render_dirty(want_stats=True)
    l2 = np.zeros((64, 96), dtype=fse.types.CELL_DTYPE)
    l2["mat"], l2["color"] = 7, 0x808080
    w.layer2_write_rect(100, 200, l2)
    w.background_write_rect(90, 190, np.full((32, 48), 0xFF112233, dtype=np.uint32))
    w.render_layers(draw_background_grid=True)
    w.scroll(-128, 0)
    for k in range(6):
        w.physics_probe(k)
    bodies = [make_body(table, 20, 24, seed=1, fill=0.8), make_body(table, 16, 16, seed=4, fill=1.0), make_body(table, 70, 60, seed=2, fill=0.7)]
    xf = np.array([(300.0, 300.0, 0.3), (310.0, 290.0, -0.7), (200.0, 200.0, 1.0)], dtype=np.float32)
    w.bodies_upload(bodies)
    w.bodies_raster(xf, tick=1)
    w.bodies_erase(xf)
    masks = (np.arange(2 * 24 * 20).reshape(2, 24, 20) % 7 != 0).astype(np.uint8)
    w.mask_outline(masks)
    w.update_rigid_body_hitbox(0, angle=0.3)
    w.flood_component(300, 300)
    ents = np.array([(250.0, 220.0, 1.5, 2.0, 8, 14, 0, 0), (256.0, 228.0, -1.0, -2.5, 8, 14, 0, 0), (400.0, 300.0, 0.0, 3.0, 6, 10, 0, 0)], dtype=fse.types.ENTITY_DTYPE)
    ents = w.entities_tick(ents, tick=2)
    w.entities_stamp(ents, tick=2)
    w.tool_erase_line(120, 150, 380, 330, 7)
    w.tool_pickaxe(280, 250, 19.0)
    w.tool_hammer(300, 280, 290, 268, tick=2)
    w.tool_vacuum(320, 300, 350, 330, tick=2)
    w.particles_vacuum_pull(320.0, 300.0)
    w.object_delete()
    w.particles_tick()
s = w.stats()
w.sync()
print(f"sanitize_small {what}: hash={s.hash:016x} particles={w.particles_count()} launches={ctx.launch_count()}")
w.close()
ctx.close()

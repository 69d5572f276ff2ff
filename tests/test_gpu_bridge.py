"""Rigid-body bridge, fracture outlines and flood fill on the GPU vs the CPU oracle (bit-exact).  `pytest -m gpu`."""
import numpy as np
import pytest

import falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh
from tests.test_bridge_cpu import make_body

pytestmark = pytest.mark.gpu


def _bodies(table, n, rng, size=(12, 33)):
    bodies, xf = [], []
    for i in range(n):
        w, h = int(rng.integers(*size)), int(rng.integers(*size))
        bodies.append(make_body(table, w, h, seed=100 + i, fill=0.7 if i % 3 else 1.0))
        xf.append((float(rng.uniform(160, 850)), float(rng.uniform(160, 600)), float(rng.uniform(-3.1, 3.1))))
    return bodies, np.array(xf, dtype=np.float32)


def test_raster_erase_exact_and_sequential_order(oracle, gpu_ctx, table):
    """Overlapping bodies over sand, water and air: the dependency rounds must reproduce the sequential loops
    (game.cpp:1711-1815, 1896-1983) exactly — grid, body tiles, feedback counters and spawned particles."""
    W, H = 1024, 768
    gpu_ctx.set_materials(table)
    gw, ow = fse.World(gpu_ctx, W, H), oracle.OracleWorld(W, H, table)
    cells = G.mixed_band(table, W, H, 0, H, seed=21, air_frac=0.6, blob=48)
    rng = np.random.default_rng(7)
    bodies, xf = _bodies(table, 60, rng)
    xf[1, :2] = xf[0, :2] + 5       # two pairs of overlapping bodies: cross-body order matters
    xf[3, :2] = xf[2, :2] + (3, -2)
    for w in (gw, ow):
        w.write_rect(0, 0, cells)
    ob = [b.copy() for b in bodies]
    gw.bodies_upload(bodies)
    for tick in range(3):
        fb_o = oracle.bodies_raster(ow, ob, xf, tick=tick)
        fb_g = gw.bodies_raster(xf, tick=tick)
        assert np.array_equal(fb_o, fb_g), tick
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"raster {tick}")
        Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"raster {tick}")
        for w in (gw, ow):
            w.tick(tick)
        fe_o = oracle.bodies_erase(ow, ob, xf)
        fe_g, need = gw.bodies_erase(xf)
        assert np.array_equal(fe_o, fe_g) and need.all()
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"erase {tick}")
        for i in range(len(bodies)):
            assert gw.bodies_read(i).tobytes() == ob[i].tobytes(), (tick, i)
        xf[:, 1] += 1.5       # bodies drift down a little between ticks (the host's Box2D step)
        xf[:, 2] += 0.05


def test_large_bodies_use_the_global_claim_plane(oracle, gpu_ctx, table):
    """Bodies whose footprint box exceeds the 64x64 shared-memory claim map (here 70x60 and 90x40) run their rounds on the claim
    plane in global memory; they overlap each other and two small bodies, so body order and both claim paths are exercised."""
    W, H = 768, 640
    gpu_ctx.set_materials(table)
    gw, ow = fse.World(gpu_ctx, W, H), oracle.OracleWorld(W, H, table)
    cells = G.mixed_band(table, W, H, 0, H, seed=5, air_frac=0.6, blob=48)
    bodies = [make_body(table, 20, 24, seed=1, fill=0.8), make_body(table, 70, 60, seed=2, fill=0.7), make_body(table, 90, 40, seed=3, fill=1.0),
              make_body(table, 16, 16, seed=4, fill=1.0), make_body(table, 260, 2, seed=5, fill=0.9)]  # the last: a 267 x 11 box fits the shared-memory map but not 8-bit coordinates
    xf = np.array([(300.0, 300.0, 0.3), (310.0, 290.0, -0.7), (330.0, 320.0, 1.9), (335.0, 300.0, 0.0), (200.0, 450.0, 0.02)], dtype=np.float32)
    for w in (gw, ow):
        w.write_rect(0, 0, cells)
    ob = [b.copy() for b in bodies]
    gw.bodies_upload(bodies)
    for tick in range(2):
        assert np.array_equal(oracle.bodies_raster(ow, ob, xf, tick=tick), gw.bodies_raster(xf, tick=tick))
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"raster {tick}")
        Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"raster {tick}")
        fe_o = oracle.bodies_erase(ow, ob, xf)
        fe_g, _ = gw.bodies_erase(xf)
        assert np.array_equal(fe_o, fe_g)
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"erase {tick}")
        for i in range(len(bodies)):
            assert gw.bodies_read(i).tobytes() == ob[i].tobytes(), (tick, i)
        xf[:, 1] += 2.0
        xf[:, 2] += 0.11


def test_round_trip_identity_on_empty_grid(gpu_ctx, table):
    W = H = 384
    gpu_ctx.set_materials(table)
    gw = fse.World(gpu_ctx, W, H)
    gw.write_rect(0, 0, Hh.empty_world_cells(table, W, H))
    before = gw.read_all()
    body = make_body(table, 24, 20, fill=1.0)
    gw.bodies_upload([body])
    xf = [(180.0, 170.0, 0.0)]
    fb = gw.bodies_raster(xf)
    assert fb[0, 2] == 480
    fe, _ = gw.bodies_erase(xf)
    assert fe[0, 2] == 480 and fe[0, 3] == 0
    after = gw.read_all()
    for f in ("mat", "color", "temp", "fluid", "moved"):
        assert np.array_equal(after[f], before[f]), f
    assert gw.bodies_read(0).tobytes() == body.tobytes()


def test_outlines_and_labels_match_oracle(oracle, gpu_ctx, table):
    gpu_ctx.set_materials(table)
    gw = fse.World(gpu_ctx, 384, 384)
    rng = np.random.default_rng(11)
    masks = []
    for k in range(24):
        m = (rng.random((48, 64)) < (0.35 + 0.02 * k)).astype(np.uint8)
        if k % 4 == 0:
            m[10:30, 8:50] = 1
            m[15:20, 20:30] = 0  # a hole
        if k % 5 == 0:
            m[:, :] = 0
            m[5:40, 5:9] = 1     # a crack: two separate pieces
            m[5:40, 11:60] = 1
        masks.append(m)
    masks = np.stack(masks)
    labels, ncomp, contours = gw.mask_outline(masks)
    for k in range(len(masks)):
        lo, no = oracle.ccl(masks[k])
        assert no == ncomp[k], k
        assert np.array_equal(lo, labels[k]), k
        co = oracle.outlines(masks[k])
        assert len(co) == len(contours[k]), (k, len(co), len(contours[k]))
        for a, b in zip(co, contours[k]):
            assert a.tobytes() == b.tobytes(), k
    assert ncomp[5] == 2  # the cracked plate fell into two pieces


def test_chunk_solid_mask_outline(oracle, gpu_ctx, table):
    """updateChunkMesh (world.cpp:722-959): SOLID mask of a 128x128 chunk -> contours."""
    W = H = 512
    gpu_ctx.set_materials(table)
    gw = fse.World(gpu_ctx, W, H)
    cells = G.mixed_band(table, W, H, 0, H, seed=5, blob=24)
    gw.write_rect(0, 0, cells)
    phys = table.physics()
    m = gw.solid_mask(128, 128, 128, 128)
    assert np.array_equal(m, (phys[cells["mat"][128:256, 128:256]] == T.SOLID).astype(np.uint8))
    labels, ncomp, contours = gw.mask_outline(m[None])
    co = oracle.outlines(m)
    assert len(co) == len(contours[0]) > 0
    for a, b in zip(co, contours[0]):
        assert a.tobytes() == b.tobytes()


def test_flood_component_matches_oracle(oracle, gpu_ctx, table):
    W = H = 512
    gpu_ctx.set_materials(table)
    gw, ow = fse.World(gpu_ctx, W, H), oracle.OracleWorld(W, H, table)
    cells = Hh.empty_world_cells(table, W, H)
    rng = np.random.default_rng(2)
    blob = (rng.random((60, 60)) < 0.62)
    cells["mat"][200:260, 200:260] = np.where(blob, 7, 0)
    cells["mat"][300, 150:400] = 7          # a 250-long line: many BFS wavefronts
    cells["mat"][320:360, 150:190] = 7      # 1600 cells: over the cap
    for w in (gw, ow):
        w.write_rect(0, 0, cells)
    seeds = [(x, y) for y in range(200, 260, 7) for x in range(200, 260, 7)] + [(200, 300), (399, 300), (160, 330), (10, 10), (140, 140)]
    for (x, y) in seeds:
        no, bo, po = oracle.flood_component(ow, x, y)
        ng, bg, pg = gw.flood_component(x, y)
        assert no == ng, (x, y, no, ng)
        if 0 < no <= 1000:
            assert list(bo) == list(bg) and np.array_equal(po, pg), (x, y)


def test_explosion_exact(oracle, gpu_ctx, table):
    """fse_explosion vs the oracle restatement of world::explosion: identical grid, dirty flags and particle set."""
    from oracle import pyoracle as O
    W, H = 640, 512
    tbl, extra = G.bench_table(table)
    gpu_ctx.set_materials(tbl)
    ow, gw = oracle.OracleWorld(W, H, tbl), fse.World(gpu_ctx, W, H)
    Hh.build_mixed(ow, tbl, W, H, seed=21, extra=list(extra.values()), blob=16)
    Hh.build_mixed(gw, tbl, W, H, seed=21, extra=list(extra.values()), blob=16)
    for (cx, cy, r) in ((300, 250, 24), (10, 10, 12), (630, 500, 30)):  # the last two reach over the world border
        O.explosion(ow, cx, cy, r, tick=3, seed=77)
        gw.explosion(cx, cy, r, tick=3, seed=77)
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "explosion")
    Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), "explosion particles")
    for t in range(4):  # and the thrown particles land identically
        ow.tick(t); gw.tick(t)
        ow.particles_tick(); gw.particles_tick()
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "after explosion")
    Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), "particles after explosion")


def test_render_dirty_exact(oracle, gpu_ctx, table):
    """fse_render_dirty vs the oracle restatement of the dirty -> texture loop (game.cpp:1994-2060): the three RGBA planes, the
    movingTiles histogram and the dirty / fire counts after ticks of a mixed world; clean cells keep their old texels."""
    from oracle import pyoracle as O
    W, H = 640, 512
    tbl, extra = G.bench_table(table)
    gpu_ctx.set_materials(tbl)
    ow, gw = oracle.OracleWorld(W, H, tbl), fse.World(gpu_ctx, W, H)
    ow.default_schedule = 2
    Hh.build_mixed(ow, tbl, W, H, seed=4, extra=list(extra.values()))
    Hh.build_mixed(gw, tbl, W, H, seed=4, extra=list(extra.values()))
    gw.pixels_enable(True)
    planes = [np.zeros((H, W, 4), dtype=np.uint8) for _ in range(3)]
    for t in range(3):
        ow.clear_dirty()
        gw.clear_dirty()
        ow.tick(t, seed=2)
        gw.tick(t, seed=2)
        d_o, f_o, m_o = O.render_dirty(ow, planes)
        d_g, f_g, m_g = gw.render_dirty()
        assert (d_o, f_o) == (d_g, f_g) and d_o > 0 and np.array_equal(m_o, m_g), t
        for which in range(3):
            assert np.array_equal(planes[which], gw.pixels_read(which)), (t, which)
    assert f_o > 0
    sub = gw.pixels_read(0, T.Rect(100, 60, 33, 17))
    assert np.array_equal(sub, planes[0][60:77, 100:133])


@pytest.mark.parametrize("dx,dy", [(128, 0), (-128, 256), (7, -5), (0, 1), (-1000, 3)])
def test_scroll_exact(oracle, gpu_ctx, table, dx, dy):
    """fse_scroll vs the oracle's restatement of the tickChunks shift: cells, dirty flags (not shifted) and particles."""
    from oracle import pyoracle as O
    W, H = 640, 512
    gpu_ctx.set_materials(table)
    ow, gw = oracle.OracleWorld(W, H, table), fse.World(gpu_ctx, W, H)
    Hh.build_mixed(ow, table, W, H, seed=12)
    Hh.build_mixed(gw, table, W, H, seed=12)
    parts = np.zeros(3, dtype=T.PARTICLE_DTYPE)
    parts["x"], parts["y"], parts["id"] = [5.5, 300.0, 639.0], [7.25, 200.0, 500.0], [1, 2, 3]
    parts["tile"]["mat"] = 2
    for w in (ow, gw):
        w.tick(0)
        w.particles_add(parts)
    O.scroll(ow, dx, dy)
    gw.scroll(dx, dy)
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"scroll {dx},{dy}")
    Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), "scroll particles")
    ow.tick(1)
    gw.tick(1)
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "tick after scroll")


def test_bodies_split_matches_oracle(oracle, gpu_ctx, table):
    """fse_bodies_split on uploaded bodies: a cracked plate, a body with hashed holes (dozens of crumbs) and an intact one; pieces,
    crop boxes, weld flags, shifts and tile arrays equal the oracle's; the pieces of the plate union to the plate."""
    gpu_ctx.set_materials(table)
    gw = fse.World(gpu_ctx, 512, 384)
    plate = make_body(table, 40, 24, fill=1.0)
    plate["mat"][:, 17] = 0
    plate["mat"][3:6, 25:28] = 0
    plate["mat"][4, 26] = 22
    bodies = [plate, make_body(table, 64, 48, seed=9, fill=0.55), make_body(table, 16, 16, seed=2, fill=1.0), make_body(table, 128, 128, seed=4, fill=0.62)]
    gw.bodies_upload(bodies)
    for i, b in enumerate(bodies):
        want = oracle.body_split(b, angle=0.3 * i, weld=(5, 5))
        got = gw.bodies_split(i, angle=0.3 * i, weld=(5, 5))
        assert len(want) == len(got), i
        for (rw, tw), (rg, tg) in zip(want, got):
            for f in ("x0", "y0", "w", "h", "n_pixels", "weld", "tile_off"):
                assert rw[f] == rg[f], (i, f)
            assert abs(rw["shift_x"] - rg["shift_x"]) < 1e-4 and abs(rw["shift_y"] - rg["shift_y"]) < 1e-4
            assert tw.tobytes() == tg.tobytes(), i
    got = gw.bodies_split(0)
    assert len(got) == 3 and sum(int((t["mat"] != 0).sum()) for _, t in got) == int((plate["mat"] != 0).sum())
    gw.close()


def test_physics_check_cut_out_exact(oracle, gpu_ctx, table):
    """fse_physics_check vs the oracle's world::physicsCheck: AIR seed, abandoned flood, deleted crumb, two cut-outs (block with a hole,
    sprawling L); result records, body tiles and the grid afterwards are identical; a tile buffer that is too small changes nothing."""
    from oracle import pyoracle as O
    from tests.test_bridge_cpu import _physcheck_world
    W, H = 768, 640
    gpu_ctx.set_materials(table)
    ow, gw = oracle.OracleWorld(W, H, table), fse.World(gpu_ctx, W, H)
    cells = _physcheck_world(table, W, H)
    for w in (ow, gw):
        w.write_rect(0, 0, cells)
        w.clear_dirty()
    for (x, y) in [(150, 150), (420, 315), (5, 5), (161, 250), (305, 208), (200, 360), (200, 360), (-3, 10), (100, 9999)]:
        a, b = O.physics_check(ow, x, y) if 0 <= x < W and 0 <= y < H else (0, 0, (0, 0, 0, 0), None), gw.physics_check(x, y)
        assert a[:3] == b[:3], ((x, y), a[:3], b[:3])
        assert (a[3] is None) == (b[3] is None)
        if a[3] is not None:
            Hh.assert_cells_equal(a[3], b[3], f"body tiles of the cut-out at {(x, y)}")
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"grid after physicsCheck{(x, y)}")
    gw.write_rect(0, 0, cells)
    ow.write_rect(0, 0, cells)
    with pytest.raises(fse.FseError):
        gw.physics_check(200, 360, cap_tiles=100)
    assert np.array_equal(gw.read_all()["mat"], cells["mat"])
    assert gw.physics_check(305, 208)[:2] == O.physics_check(ow, 305, 208)[:2] == (105, 2)
    ow.tick(0); gw.tick(0)
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "tick after the cut-outs")

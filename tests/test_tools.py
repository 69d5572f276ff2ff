"""Interactive tools, grid side: erase brush (game.cpp:593-625), pickaxe (771-790), hammer (843-890), vacuum
(2456-2585, 2640-2664) over world::forLine / forLineCornered (world.cpp:3250-3313).

CPU part: pins of the oracle restatement derived from the reference source.  GPU part (`-m gpu`): the kernels against the oracle,
bit-exact on the grid and the particle pool."""
import numpy as np
import pytest

from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh

SAND, STONE, WATER = 2, 7, 15


def _scene(table, W=640, H=512):
    cells = Hh.empty_world_cells(table, W, H)
    mat = cells["mat"].copy()
    mat[260:330, 160:420] = STONE
    mat[250:260, 200:260] = SAND
    mat[240:260, 300:360] = WATER
    return G.cells_from_mat(table, mat, 0, 0, 5)


def test_erase_brush_clears_a_band(oracle, table):
    """forLine + the brush square minus its |dx| + |dy| == size corners: a diagonal stroke through stone leaves an AIR band about
    brush_size wide, marked dirty; AIR cells are not touched (no dirty flag)."""
    ow = oracle.OracleWorld(640, 512, table)
    ow.write_rect(0, 0, _scene(table))
    ow.clear_dirty()
    n = oracle.tool_erase_line(ow, 180, 270, 300, 320, 6)
    c = ow.read_all()
    assert n > 500 and int((c["mat"][260:330, 160:420] == 0).sum()) == n
    assert int(c["dirty"].sum()) == n
    for k in range(0, 121, 10):  # every point of the stroke is clear
        assert c["mat"][270 + (50 * k) // 120, 180 + k] == 0
    assert c["mat"][325, 185] == STONE  # far from the stroke


def test_pickaxe_takes_a_disc_of_solid_and_returns_its_colours(oracle, table):
    ow = oracle.OracleWorld(640, 512, table)
    cells = _scene(table)
    ow.write_rect(0, 0, cells)
    pix, n = oracle.tool_pickaxe(ow, 200, 255, 16.0)
    c = ow.read_all()
    taken = (cells["mat"][255:271, 200:216] == STONE) & (c["mat"][255:271, 200:216] == 0)
    assert n == int(taken.sum()) and 60 < n < 16 * 16           # stone part of the disc only (rows 260..), sand above stays
    assert (c["mat"][255:260, 200:216] == SAND).all()
    assert np.array_equal(pix != 0, taken)                      # pixel (yy, xx) = colour of the cell that was taken
    assert np.array_equal(pix[taken], cells["color"][255:271, 200:216][taken])


def test_hammer_cracks_solid_into_darkened_sand_until_it_breaks_out(oracle, table):
    """Start inside the stone slab, release point 40 cells below-left: the crack runs up-right in jittered segments, turns
    STONE into GENERIC_SAND at half brightness and stops one cell after it leaves the slab (`broke`)."""
    ow = oracle.OracleWorld(640, 512, table)
    cells = _scene(table)
    ow.write_rect(0, 0, cells)
    ex, ey, n, broke = oracle.tool_hammer(ow, 300, 290, 270, 330, SAND, tick=4)
    c = ow.read_all()
    changed = (cells["mat"] == STONE) & (c["mat"] == SAND)
    assert n == int(changed.sum()) and n > 25 and broke == 1
    assert changed[ey, ex] and ey < 290 and ex > 300 and not changed[:260].any()
    old, new = cells["color"][changed], c["color"][changed]
    assert np.array_equal((new >> 16) & 0xff, ((old >> 16) & 0xff) // 2) and np.array_equal(new & 0xff, (old & 0xff) // 2)
    assert (c["dirty"][changed] == 1).all() and (c["temp"][changed] == 0).all()


def test_vacuum_sucks_the_disc_and_pulls_the_particles_in(oracle, table):
    """The walk from the screen centre stops at the first SAND cell; the matter of the 11 x 11 disc becomes `phase` particles with
    lifetime 6 held by the vacuum; once their lifetime is up they fly to the player and are collected within 10 cells."""
    ow = oracle.OracleWorld(640, 512, table)
    cells = _scene(table)
    ow.write_rect(0, 0, cells)
    x, y, n, caught = oracle.tool_vacuum(ow, 230, 200, 230, 300, tick=2)
    assert (x, y) == (230, 250) and n > 30 and caught == 0
    p = ow.particles_read()
    assert len(p) == n and p["phase"].all() and p["vacuum"].all() and (p["lifetime"] == 6).all()
    c = ow.read_all()
    assert int((cells["mat"] != 0).sum()) - int((c["mat"] != 0).sum()) == n
    assert oracle.tool_vacuum(ow, 230, 200, 230, 200 + 300, tick=3)[0] == -1  # out of reach
    collected = 0
    for t in range(60):
        ow.particles_tick(schedule=oracle.REFERENCE)
        collected += oracle.particles_vacuum_pull(ow, 230.0, 200.0)
    assert collected == n and ow.particles_count() == 0


@pytest.mark.gpu
def test_tools_match_oracle_on_gpu(oracle, gpu_ctx, table):
    import falling_sand_engine_b200 as fse

    W, H = 640, 512
    gpu_ctx.set_materials(table)
    gw, ow = fse.World(gpu_ctx, W, H), oracle.OracleWorld(W, H, table)
    cells = G.mixed_band(table, W, H, 0, H, seed=9, air_frac=0.3, blob=24)
    for w in (gw, ow):
        w.write_rect(0, 0, cells)
        w.tick(0)
        w.particles_tick()
    rng = np.random.default_rng(12)
    for k in range(10):
        a = [int(v) for v in rng.integers(140, 370, 8)]
        assert oracle.tool_erase_line(ow, a[0], a[1], a[2] + 120, a[3], 3 + k) >= 0
        gw.tool_erase_line(a[0], a[1], a[2] + 120, a[3], 3 + k)
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"erase {k}")
        po, no = oracle.tool_pickaxe(ow, a[4], a[5], 9.5 + k)
        pg, ng = gw.tool_pickaxe(a[4], a[5], 9.5 + k)
        assert no == ng and np.array_equal(po, pg), k
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"pickaxe {k}")
        hx, hy, rx, ry = a[6] + 60, a[7], a[6] + 60 + int(rng.integers(-60, 60)), a[7] + int(rng.integers(-60, 60))
        assert oracle.tool_hammer(ow, hx, hy, rx, ry, 2, tick=k) == gw.tool_hammer(hx, hy, rx, ry, 2, tick=k), k
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"hammer {k}")
        mx, my = a[0] + int(rng.integers(-90, 90)), a[5] + int(rng.integers(-90, 90))
        assert oracle.tool_vacuum(ow, a[0], a[5], mx, my, tick=k) == gw.tool_vacuum(a[0], a[5], mx, my, tick=k), k
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"vacuum {k}")
        Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"vacuum {k}")
        assert oracle.particles_vacuum_pull(ow, float(a[0]), float(a[5])) == gw.particles_vacuum_pull(float(a[0]), float(a[5]))
        for w in (gw, ow):
            w.tick(k + 1)
            w.particles_tick()
        Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"tick {k}")
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "after the tool sequence")

"""Edge cases of the C ABI against the oracle (bit-exact): ragged world sizes, tick zones that are offset and not aligned to the
world's chunk grid, the smallest world and zone, every cell_iter the ABI takes, empty inputs, and the argument errors.  `pytest -m gpu`."""
import numpy as np
import pytest

import falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _fill(table, W, H, seed, blob=16):
    cells = G.mixed_band(table, W, H, 0, H, seed=seed, blob=blob)
    return cells


@pytest.mark.parametrize("W,H,zone", [
    (416, 390, (16, 23, 384, 256)),      # ragged world, zone offset by (16, 23): chunks do not sit on the world's 128-grid
    (384, 384, (128, 128, 128, 128)),    # the smallest world and the smallest zone: one chunk, three idle colours
    (528, 701, (64, 311, 384, 128)),     # a strip of three chunks low in a tall ragged world
    (1056, 416, (16, 16, 1024, 384)),    # zone as large as the 16-cell margin allows
])
@pytest.mark.parametrize("sched", ["rows", "rows_fused"])
def test_offset_and_ragged_zones_exact(oracle, gpu_ctx, table, W, H, zone, sched, monkeypatch):
    if sched == "rows_fused":
        monkeypatch.setenv("FSE_FUSED_MAX_CHUNKS", "296")
    gpu_ctx.set_materials(table)
    ow, gw = oracle.OracleWorld(W, H, table), fse.World(gpu_ctx, W, H)
    ow.default_schedule = 2
    cells = _fill(table, W, H, seed=W + H)
    ow.write_rect(0, 0, cells)
    gw.write_rect(0, 0, cells)
    z = T.Rect(*zone)
    for t in range(6):
        ow.tick(t, seed=9, zone=z)
        gw.tick(t, seed=9, zone=z)
        ow.particles_tick(zone=z)
        gw.particles_tick(zone=z)
        if t == 2:
            ow.tick_temperature(zone=z)
            gw.tick_temperature(zone=z)
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"{W}x{H} zone {zone}")
    Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), "particles")
    # cells outside the zone (beyond the reach of its border chunks) never change
    after = gw.read_all()
    far = np.ones((H, W), dtype=bool)
    far[max(0, z.y - 16):z.y + z.h + 16, max(0, z.x - 16):z.x + z.w + 16] = False
    assert np.array_equal(after["mat"][far], cells["mat"][far])


@pytest.mark.parametrize("cell_iter", [0, 1, 2, 4])
def test_every_cell_iter(oracle, gpu_ctx, table, cell_iter):
    W, H = 640, 512
    gpu_ctx.set_materials(table)
    ow, gw = oracle.OracleWorld(W, H, table), fse.World(gpu_ctx, W, H)
    ow.default_schedule = 2
    cells = _fill(table, W, H, seed=3)
    ow.write_rect(0, 0, cells)
    gw.write_rect(0, 0, cells)
    for t in range(3):
        ow.tick(t, cell_iter=cell_iter)
        gw.tick(t, cell_iter=cell_iter)
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"cell_iter {cell_iter}")
    if cell_iter == 0:
        Hh.assert_cells_equal(cells, gw.read_all(), "cell_iter 0 is a no-op")


def test_uniform_worlds_and_empty_inputs(oracle, gpu_ctx, table):
    """All AIR, all SOLID, all WATER inside the border; ticks with an empty particle pool; particles outside the zone stay put; a
    particle below the world is dropped; reads and writes of single cells at the corners."""
    W, H = 512, 384
    gpu_ctx.set_materials(table)
    for mat in (0, 7, 15):
        ow, gw = oracle.OracleWorld(W, H, table), fse.World(gpu_ctx, W, H)
        ow.default_schedule = 2
        m = np.full((H, W), mat, dtype=np.uint16)
        G.border_fill(m, 0, 0, W, H, 1)
        cells = G.cells_from_mat(table, m, 0, 0, 5)
        ow.write_rect(0, 0, cells)
        gw.write_rect(0, 0, cells)
        assert gw.particles_count() == 0
        gw.particles_tick()  # empty pool
        for t in range(3):
            ow.tick(t); gw.tick(t)
            ow.particles_tick(); gw.particles_tick()
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"uniform material {mat}")
        Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"uniform material {mat}")
        parts = np.zeros(3, dtype=T.PARTICLE_DTYPE)
        parts["x"], parts["y"], parts["vy"], parts["id"] = [5.0, 300.0, 200.0], [5.0, 200.0, H + 3.0], [1.0, 1.0, 0.0], [1, 2, 3]
        parts["tile"]["mat"], parts["tile"]["fluid"] = 2, 2.0
        ow.particles_add(parts); gw.particles_add(parts)
        ow.particles_tick(); gw.particles_tick()
        Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), "edge particles")
        one = gw.read_rect(W - 1, H - 1, 1, 1)
        assert one.shape == (1, 1) and one["mat"][0, 0] == 1
        one["mat"], one["color"] = 2, 0x123456
        gw.write_rect(0, 0, one)
        assert gw.read_rect(0, 0, 1, 1)["color"][0, 0] == 0x123456
        gw.close()
        ow.close()


def test_argument_errors(gpu_ctx, table):
    """The C ABI refuses what it cannot do, with a message, and leaves the world usable."""
    gpu_ctx.set_materials(table)
    with pytest.raises(fse.FseError):
        fse.World(gpu_ctx, 200, 384)       # narrower than three chunks
    with pytest.raises(fse.FseError):
        fse.World(gpu_ctx, 390, 384)       # width not a multiple of 16
    gw = fse.World(gpu_ctx, 512, 384)
    for zone in [(128, 128, 100, 128), (128, 128, 256, 0), (8, 128, 256, 128), (128, 128, 512, 128), (128, 300, 256, 128)]:
        with pytest.raises(fse.FseError):
            gw.tick(0, zone=T.Rect(*zone))
    with pytest.raises(fse.FseError):
        gw.tick(0, cell_iter=5)
    with pytest.raises(fse.FseError):
        gw.read_rect(500, 0, 20, 1)
    with pytest.raises(fse.FseError):
        gw.write_rect(0, 380, np.zeros((8, 8), dtype=T.CELL_DTYPE))
    bad = np.zeros((2, 2), dtype=T.CELL_DTYPE)
    bad["mat"] = 250
    with pytest.raises(fse.FseError):
        gw.write_rect(10, 10, bad)         # material outside the table
    with pytest.raises(fse.FseError):
        gw.explosion(100, 100, 0)
    with pytest.raises(fse.FseError):
        gw.pixels_read(0)                  # planes not enabled
    gw.tick(0)                             # still fine
    gw.close()

"""CPU pins of the render-only planes next to the grid: the liquid flow accumulators flowX / flowY
(world.cpp:1334, 1374, 1402, 1432), the flow texture (game.cpp:2040-2062) and the layer-2 / background loops
(game.cpp:2068-2126).  Known answers are worked out here from the reference's formulas, not taken from the oracle."""
import numpy as np
import pytest

from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh

WATER, STONE, SAND = 15, 7, 2
f32 = np.float32


def _one_water_cell(oracle, table, schedule):
    W, H = 384, 384
    ow = oracle.OracleWorld(W, H, table)
    cells = Hh.empty_world_cells(table, W, H)
    cells["mat"][200, 128:256] = STONE          # floor under the water cell
    cells["mat"][199, 190], cells["fluid"][199, 190] = WATER, 2.0
    cells["color"][199, 190] = 0x204060
    ow.write_rect(0, 0, cells)
    from oracle import pyoracle as O
    O.flow_enable(ow)
    ow.tick(0, seed=3, cell_iter=1, schedule=schedule)
    return ow


@pytest.mark.parametrize("schedule", [0, 2])
def test_flow_accumulators_of_one_cell(oracle, table, schedule):
    """A 2.0 water cell on a stone floor with AIR left, right and above, one iteration: left flow = 2/3 (three-way split,
    world.cpp:1358), right flow = half of what is left, up flow = rem - CalculateVerticalFlowValue(rem, 0); nothing goes down.
    flowX = -left + right, flowY = -up at the source cell (1374, 1402, 1432), zero elsewhere — under either schedule."""
    from oracle import pyoracle as O
    ow = _one_water_cell(oracle, table, schedule)
    fx, fy = O.flow_read(ow, 0), O.flow_read(ow, 1)
    rem = f32(2.0)
    fl = f32(rem / f32(3.0))
    rem = f32(rem - fl)
    fr = f32(rem / f32(2.0))
    rem = f32(rem - fr)
    s = rem  # remaining + 0 above; sum <= 2 * MaxValue + MaxCompression -> second arm of CalculateVerticalFlowValue (world.cpp:1026)
    assert f32(0.5) < s < f32(1.1)
    val = f32(f32(f32(0.5) * f32(0.5) + f32(s * f32(0.1))) / f32(f32(0.5) + f32(0.1)))
    fu = f32(rem - val)
    assert fu > 0
    want_x = f32(f32(f32(0.0) - fl) + fr)
    want_y = f32(f32(0.0) - fu)
    assert fx[199, 190] == want_x and fy[199, 190] == want_y
    fx[199, 190] = fy[199, 190] = 0
    assert not fx.any() and not fy.any()
    cells = ow.read_rect(189, 198, 3, 2)
    assert cells["mat"][1, 0] == WATER and cells["mat"][1, 2] == WATER and cells["mat"][0, 1] == WATER  # the three targets became water


def test_flow_texture_known_answer(oracle, table):
    """game.cpp:2040-2062 on the cell above: newFlow = prev + (flow - prev) * 0.25, negative y halved, byte = clamp(newFlow *
    (3 / iterations + 0.5) / 4 + 0.5) * 255 truncated; r = x, g = y, b = 0, a = 255; accumulators reset, prevFlow kept; a clean or
    non-liquid cell leaves the texture alone and an AIR or non-liquid dirty cell only resets its accumulators."""
    from oracle import pyoracle as O
    ow = _one_water_cell(oracle, table, 2)
    W, H = ow.width, ow.height
    fx, fy = O.flow_read(ow, 0)[199, 190], O.flow_read(ow, 1)[199, 190]
    planes = [np.full((H, W, 4), 9, dtype=np.uint8) for _ in range(4)]
    d, f, moving, nflow = O.render_dirty(ow, planes, with_flow_count=True)
    cells = ow.read_all()
    dirty_soup = (cells["dirty"] != 0) & (cells["mat"] == WATER)
    assert nflow == int(dirty_soup.sum()) and dirty_soup[199, 190]
    iters = int(table.mats[WATER].iterations)
    nx = f32(np.float64(0.0) + np.float64(f32(fx - f32(0.0))) * 0.25)
    ny = f32(np.float64(0.0) + np.float64(f32(fy - f32(0.0))) * 0.25)
    if ny < 0:
        ny = f32(np.float64(ny) * 0.5)
    k = 3.0 / iters + 0.5
    by = int(min(max(np.float64(ny) * k / 4.0 + 0.5, 0.0), 1.0) * 255)
    bx = int(min(max(np.float64(nx) * k / 4.0 + 0.5, 0.0), 1.0) * 255)
    assert planes[3][199, 190].tolist() == [bx, by, 0, 255]
    assert by < 127  # the cell pushed water upwards
    assert planes[3][10, 10].tolist() == [9, 9, 9, 9] and planes[3][200, 190].tolist() == [9, 9, 9, 9]  # AIR / STONE: untouched
    assert O.flow_read(ow, 2)[199, 190] == nx and O.flow_read(ow, 3)[199, 190] == ny
    assert not O.flow_read(ow, 0).any() and not O.flow_read(ow, 1).any()


def test_flow_sums_agree_between_schedules(oracle, table):
    """The product's ROWS schedule against the reference order on a slab of water spreading over a floor: the total left / right
    transport (sum |flowX|) and the net vertical transport agree within 5 % after the first tick (the in-row scan order only
    decides who of two neighbours moves first)."""
    from oracle import pyoracle as O
    W, H = 640, 384
    tot = {}
    for schedule in (0, 2):
        ow = oracle.OracleWorld(W, H, table)
        cells = Hh.empty_world_cells(table, W, H)
        cells["mat"][250, 128:512] = STONE
        cells["mat"][200:250, 256:384] = WATER
        ow.write_rect(0, 0, cells)
        O.flow_enable(ow)
        ow.tick(0, seed=5, schedule=schedule)
        fx, fy = O.flow_read(ow, 0).astype(np.float64), O.flow_read(ow, 1).astype(np.float64)
        tot[schedule] = (np.abs(fx).sum(), fy.sum(), np.abs(fy).sum())
    for a, b in zip(tot[0], tot[2]):
        assert abs(a - b) <= 0.05 * max(abs(a), abs(b), 1.0), tot


def test_layer2_and_background_known_answers(oracle, table):
    """game.cpp:2068-2126: a dirty layer-2 cell is its colour + material alpha, AIR is transparent (or the 0x888888 / 0x444444 checker
    by cell index with draw_background_grid); a dirty background cell is its ARGB colour; clean cells are not touched; the dirty marks
    are cleared afterwards (2154-2155), so a second call changes nothing."""
    from oracle import pyoracle as O
    W, H = 16, 8
    ow = oracle.OracleWorld(W, H, table)
    l2 = np.zeros((2, 3), dtype=T.CELL_DTYPE)
    l2["mat"][0] = [STONE, 0, WATER]
    l2["color"][0] = [0x102030, 0x999999, 0x0A0B0C]
    l2["mat"][1] = [0, 0, SAND]
    l2["color"][1, 2] = 0xFFEEDD
    l2["temp"][1, 2] = -7
    O.layer2_write_rect(ow, 4, 2, l2)
    O.background_write_rect(ow, 1, 1, np.array([[0x80112233, 0x00FFFFFF]], dtype=np.uint32))
    planes = [np.full((H, W, 4), 5, dtype=np.uint8) for _ in range(2)]
    n2, nb = O.render_layers(ow, planes, draw_background_grid=True)
    assert (n2, nb) == (6, 2)
    a_stone, a_water, a_sand = int(table.mats[STONE].alpha), int(table.mats[WATER].alpha), int(table.mats[SAND].alpha)
    assert planes[0][2, 4].tolist() == [0x10, 0x20, 0x30, a_stone] and planes[0][2, 6].tolist() == [0x0A, 0x0B, 0x0C, a_water]
    assert planes[0][3, 6].tolist() == [0xFF, 0xEE, 0xDD, a_sand]
    i = 5 + 2 * W  # AIR at (5, 2): checker by cell index
    assert planes[0][2, 5].tolist() == ([0x88] * 3 if i % 2 == 0 else [0x44] * 3) + [255]
    i = 4 + 3 * W
    assert planes[0][3, 4].tolist() == ([0x88] * 3 if i % 2 == 0 else [0x44] * 3) + [255]
    assert planes[0][0, 0].tolist() == [5, 5, 5, 5]
    assert planes[1][1, 1].tolist() == [0x11, 0x22, 0x33, 0x80] and planes[1][1, 2].tolist() == [0xFF, 0xFF, 0xFF, 0x00]
    assert planes[1][2, 4].tolist() == [5, 5, 5, 5]
    back = O.layer2_read_rect(ow, 4, 2, 3, 2)
    assert np.array_equal(back["mat"], l2["mat"]) and np.array_equal(back["color"], l2["color"]) and back["temp"][1, 2] == -7
    assert not back["dirty"].any()
    assert O.background_read_rect(ow, 1, 1, 2, 1).tolist() == [[0x80112233, 0x00FFFFFF]]
    before = [p.copy() for p in planes]
    assert O.render_layers(ow, planes) == (0, 0)
    assert all(np.array_equal(a, b) for a, b in zip(before, planes))
    # without the grid a dirty AIR cell is transparent black
    O.layer2_write_rect(ow, 5, 2, np.zeros((1, 1), dtype=T.CELL_DTYPE))
    assert O.render_layers(ow, planes, draw_background_grid=False) == (1, 0)
    assert planes[0][2, 5].tolist() == [0, 0, 0, 0]


def test_scroll_moves_layer2_and_background(oracle, table):
    """world.cpp:2474-2476: real_layer2 and background move with the grid; their dirty planes do not."""
    from oracle import pyoracle as O
    W, H = 24, 16
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, Hh.empty_world_cells(table, W, H))
    l2 = np.zeros((H, W), dtype=T.CELL_DTYPE)
    l2["mat"] = (np.arange(W * H).reshape(H, W) % 5 == 0) * STONE
    l2["color"] = np.arange(W * H).reshape(H, W)
    bg = (np.arange(W * H, dtype=np.uint32).reshape(H, W) * np.uint32(2654435761)) | np.uint32(0xFF000000)
    O.layer2_write_rect(ow, 0, 0, l2)
    O.background_write_rect(ow, 0, 0, bg)
    O.scroll(ow, 3, -2)
    ys, xs = np.mgrid[0:H, 0:W]
    sy, sx = ys + 2, xs - 3
    ok = (sx >= 0) & (sx < W) & (sy >= 0) & (sy < H)
    want_col, want_bg = l2["color"].copy(), bg.copy()
    want_col[ok], want_bg[ok] = l2["color"][sy[ok], sx[ok]], bg[sy[ok], sx[ok]]
    got = O.layer2_read_rect(ow, 0, 0, W, H)
    assert np.array_equal(got["color"], want_col) and np.array_equal(O.background_read_rect(ow, 0, 0, W, H), want_bg)
    assert got["dirty"].all()

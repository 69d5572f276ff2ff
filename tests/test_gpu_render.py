"""Render-only planes on the GPU vs their CPU restatements (bit-exact): liquid flow accumulators + flow texture, layer 2 and background
planes, the camera scroll with those planes.  `pytest -m gpu`."""
import numpy as np
import pytest

import falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [False, True])
def test_flow_accumulators_and_texture_exact(oracle, gpu_ctx, table, fused, monkeypatch):
    """fse_flow_enable: flowX / flowY after every tick, then the flow texture, prevFlow planes, the reset and the hadFlow count after
    fse_render_dirty, against the oracle (ROWS schedule) on a mixed world with liquids — per-pass kernels and the fused kernel.  The
    cell state must not depend on whether the accumulators are kept."""
    from oracle import pyoracle as O
    if fused:
        monkeypatch.setenv("FSE_FUSED_MAX_CHUNKS", "296")
    W, H = 640, 512
    tbl, extra = G.bench_table(table)
    gpu_ctx.set_materials(tbl)
    ow, gw, gplain = oracle.OracleWorld(W, H, tbl), fse.World(gpu_ctx, W, H), fse.World(gpu_ctx, W, H)
    ow.default_schedule = 2
    for w in (ow, gw, gplain):
        Hh.build_mixed(w, tbl, W, H, seed=9, extra=list(extra.values()))
    O.flow_enable(ow)
    gw.flow_enable(True)
    gw.pixels_enable(True)
    planes = [np.zeros((H, W, 4), dtype=np.uint8) for _ in range(4)]
    total = 0
    for t in range(4):
        for w in (ow, gw, gplain):
            w.clear_dirty()
            w.tick(t, seed=6)
        for which in (0, 1):
            a, b = O.flow_read(ow, which), gw.flow_read(which)
            assert (a == b).all(), (t, which, int((a != b).sum()))  # == : the sign of a zero does not matter
        assert np.abs(O.flow_read(ow, 0)).sum() > 0 and np.abs(O.flow_read(ow, 1)).sum() > 0
        if t == 1:
            continue  # accumulators carry over a tick without render (the reference only resets dirty cells)
        d_o, f_o, m_o, fl_o = O.render_dirty(ow, planes, with_flow_count=True)
        d_g, f_g, m_g, fl_g = gw.render_dirty(with_flow_count=True)
        assert (d_o, f_o, fl_o) == (d_g, f_g, fl_g) and np.array_equal(m_o, m_g), t
        total += fl_o
        for which in range(4):
            assert np.array_equal(planes[which], gw.pixels_read(which)), (t, which)
        for which in range(4):
            a, b = O.flow_read(ow, which), gw.flow_read(which)
            assert (a == b).all(), (t, "after render", which)
    assert total > 1000
    Hh.assert_cells_equal(gplain.read_all(), gw.read_all(), "flow accumulators must not change the cells")
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "cells")
    sub = gw.flow_read(2, T.Rect(130, 140, 31, 9))
    assert (sub == O.flow_read(ow, 2)[140:149, 130:161]).all()
    gw.flow_enable(False)
    with pytest.raises(fse.FseError):
        gw.flow_read(0)


def test_layers_render_and_scroll_exact(oracle, gpu_ctx, table):
    """fse_layer2_* / fse_background_* / fse_render_layers / fse_scroll against the oracle: random layer-2 cells and background colours
    written in rects, rendered with and without the background grid, scrolled with the grid, read back."""
    from oracle import pyoracle as O
    W, H = 512, 384
    gpu_ctx.set_materials(table)
    ow, gw = oracle.OracleWorld(W, H, table), fse.World(gpu_ctx, W, H)
    Hh.build_mixed(ow, table, W, H, seed=3)
    Hh.build_mixed(gw, table, W, H, seed=3)
    rng = np.random.default_rng(8)
    planes = [np.zeros((H, W, 4), dtype=np.uint8) for _ in range(2)]
    assert gw.render_layers() == (0, 0)  # nothing written yet: nothing allocated, nothing dirty
    for step in range(4):
        for _ in range(3):
            rw, rh = int(rng.integers(1, 200)), int(rng.integers(1, 120))
            x, y = int(rng.integers(0, W - rw + 1)), int(rng.integers(0, H - rh + 1))
            l2 = np.zeros((rh, rw), dtype=T.CELL_DTYPE)
            l2["mat"] = rng.choice([0, 0, 7, 9, 15, 22], size=(rh, rw))
            l2["color"] = rng.integers(0, 1 << 24, size=(rh, rw))
            l2["temp"] = rng.integers(-500, 500, size=(rh, rw))
            O.layer2_write_rect(ow, x, y, l2)
            gw.layer2_write_rect(x, y, l2)
            bg = rng.integers(0, 1 << 32, size=(rh // 2 + 1, rw // 2 + 1), dtype=np.uint64).astype(np.uint32)
            O.background_write_rect(ow, x, y, bg)
            gw.background_write_rect(x, y, bg)
        if step == 2:
            dx, dy = 37, -21
            O.scroll(ow, dx, dy)
            gw.scroll(dx, dy)
            Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "scroll with layers")
        a, b = O.layer2_read_rect(ow, 0, 0, W, H), gw.layer2_read_rect(0, 0, W, H)
        for f in ("mat", "color", "temp", "dirty", "fluid"):
            assert np.array_equal(a[f], b[f]), (step, f)
        assert np.array_equal(O.background_read_rect(ow, 0, 0, W, H), gw.background_read_rect(0, 0, W, H)), step
        grid = step % 2 == 0
        assert O.render_layers(ow, planes, draw_background_grid=grid) == gw.render_layers(draw_background_grid=grid), step
        assert np.array_equal(planes[0], gw.pixels_read(fse.PIXELS_LAYER2)), step
        assert np.array_equal(planes[1], gw.pixels_read(fse.PIXELS_BACKGROUND)), step
    assert gw.render_layers() == (0, 0)
    ow.tick(0)
    gw.tick(0)
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "tick after scroll (the plane sets were swapped)")
    cells = np.zeros((1, 1), dtype=T.CELL_DTYPE)
    cells["mat"] = 200
    with pytest.raises(fse.FseError):
        gw.layer2_write_rect(0, 0, cells)

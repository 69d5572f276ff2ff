"""CUDA path vs the CPU oracle, through the C ABI (include/fse.h).  Needs a B200: `pytest -m gpu`.

Bar (BASELINE.json north_star): bit-exact, cell for cell and field for field, against the oracle run under the
GPU's partitioned visiting order with the same counter RNG; the oracle's reference-order run is compared in
tests/test_oracle_pins.py (exact for order-independent rules, conservation/histograms otherwise).
"""
import numpy as np
import pytest

import falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


# name -> (FSE_SCHEDULE_*, oracle Schedule).  "rows" = one launch per pass and colour phase (default)
SCHEDULES = {"rows": (1, 2), "rows_fused": (2, 2)}
ALL_SCHEDULES = ["rows", "rows_fused"]


def _pair(oracle, gpu_ctx, table, W, H, sched="rows"):
    ow = oracle.OracleWorld(W, H, table)
    gpu_ctx.set_materials(table)
    gw = fse.World(gpu_ctx, W, H)
    gw.set_schedule(SCHEDULES[sched][0])
    ow.default_schedule = SCHEDULES[sched][1]
    return ow, gw


def _run_and_compare(ow, gw, ticks, seed=1337, every=1, cell_iter=3, what=""):
    for t in range(ticks):
        ow.tick(t, seed=seed, cell_iter=cell_iter)
        gw.tick(t, seed=seed, cell_iter=cell_iter)
        if (t + 1) % every == 0 or t == ticks - 1:
            Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"{what} tick {t}")
            Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"{what} tick {t}")


def test_roundtrip_rect(oracle, gpu_ctx, table):
    W = H = 384
    gpu_ctx.set_materials(table)
    gw = fse.World(gpu_ctx, W, H)
    cells = G.mixed_band(table, W, H, 0, H, seed=5, blob=16)
    cells["moved"] = (np.arange(W * H).reshape(H, W) % 3 == 0)
    cells["settle"] = (np.arange(W * H).reshape(H, W) % 11)
    cells["fluid_diff"] = np.linspace(-1, 1, W * H, dtype=np.float32).reshape(H, W)
    gw.write_rect(0, 0, cells)
    back = gw.read_all()
    Hh.assert_cells_equal(cells, back, "roundtrip")
    sub = gw.read_rect(17, 33, 100, 50)
    Hh.assert_cells_equal(cells[33:83, 17:117], sub, "sub-rect")


@pytest.mark.parametrize("sched", ALL_SCHEDULES)
def test_column_drop_exact(oracle, gpu_ctx, table, sched):
    W = H = 512
    ow, gw = _pair(oracle, gpu_ctx, table, W, H, sched)
    Hh.build_column(ow, table, W, H)
    Hh.build_column(gw, table, W, H)
    _run_and_compare(ow, gw, 40, every=4, what="column")


@pytest.mark.parametrize("sched", ALL_SCHEDULES)
def test_mixed_exact(oracle, gpu_ctx, table, sched):
    W = H = 512
    ow, gw = _pair(oracle, gpu_ctx, table, W, H, sched)
    Hh.build_mixed(ow, table, W, H, seed=99)
    Hh.build_mixed(gw, table, W, H, seed=99)
    _run_and_compare(ow, gw, 30, seed=7, every=3, what="mixed")


@pytest.mark.parametrize("sched", ALL_SCHEDULES)
def test_mixed_interactions_exact(oracle, gpu_ctx, table, sched):
    W, H = 640, 512
    tbl, extra = G.bench_table(table)
    ow, gw = _pair(oracle, gpu_ctx, tbl, W, H, sched)
    Hh.build_mixed(ow, tbl, W, H, seed=3, extra=list(extra.values()), blob=16)
    Hh.build_mixed(gw, tbl, W, H, seed=3, extra=list(extra.values()), blob=16)
    _run_and_compare(ow, gw, 24, seed=11, every=3, what="interactions")


def test_full_game_loop_exact(oracle, gpu_ctx, table):
    """world::tick + tickCells (+ tickTemperature on tick % 4 == 2, game.cpp:2157) for 40 ticks: grid, dirty flags and
    the particle pool stay bit-identical to the oracle under the partitioned schedule."""
    W = H = 512
    ow, gw = _pair(oracle, gpu_ctx, table, W, H)
    Hh.build_column(ow, table, W, H)
    Hh.build_column(gw, table, W, H)
    for t in range(40):
        for w in (ow, gw):
            w.tick(t)
            w.particles_tick()
            if t % 4 == 2:
                w.tick_temperature()
        if t % 5 == 4:
            Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"loop tick {t}")
            Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"loop tick {t}")
    assert gw.particles_count() > 0


def test_particles_deposit_conflicts(oracle, gpu_ctx, table):
    """Many particles aimed at the same few cells: exercises the claim rounds, the spiral search, liquid merging
    and the bounce; conservation of cells + particles is checked as well."""
    W = H = 384
    ow, gw = _pair(oracle, gpu_ctx, table, W, H)
    cells = Hh.empty_world_cells(table, W, H)
    floor = G.cells_from_mat(table, np.full((8, W - 256), 7, dtype=np.uint16), 128, 250)
    rng = np.random.default_rng(5)
    n = 4000
    parts = np.zeros(n, dtype=T.PARTICLE_DTYPE)
    parts["x"] = 180 + rng.integers(0, 24, n)
    parts["y"] = 200 + rng.integers(0, 30, n)
    parts["vx"] = (rng.integers(0, 10, n) - 5) / 20.0
    parts["vy"] = 1.0
    parts["ay"] = 0.1
    parts["fade_time"] = 60
    parts["id"] = np.arange(1, n + 1)
    is_water = rng.integers(0, 2, n) == 1
    parts["tile"]["mat"] = np.where(is_water, 15, 2)
    parts["tile"]["fluid"] = np.where(is_water, 0.25, 2.0)
    parts["tile"]["color"] = 0x123456
    for w in (ow, gw):
        w.write_rect(0, 0, cells)
        w.write_rect(128, 250, floor)
        w.particles_add(parts)
    for t in range(30):
        ow.particles_tick()
        gw.particles_tick()
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"deposit tick {t}")
        Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"deposit tick {t}")
    s = gw.stats()
    live = gw.particles_read()
    assert s.count[2] + int((live["tile"]["mat"] == 2).sum()) == int((~is_water).sum())


STATE_FIELDS = ["mat", "moved", "settle", "color", "temp", "fluid", "fluid_diff"]


@pytest.mark.parametrize("fused", ["0", "1"])  # per-pass kernels + state kernel / single fused kernel
def test_active_tracking_equals_full_sweep(oracle, gpu_ctx, table, monkeypatch, fused):
    """chunk sleeping has no reference behaviour to match beyond 'same cells as a full sweep'.  A sparse
    world (sealed lenses of settled sand / water in rock + pockets of falling sand and water) is ticked with tracking
    on, with tracking off, and by the oracle: identical cell state every few ticks (dirty flags of sleeping chunks are
    not refreshed), and most chunks are asleep at the end."""
    monkeypatch.setenv("FSE_ACTIVE_FUSED", fused)
    W, H = 1536, 1024
    cells = G.sparse_band(table, W, H, 0, H, seed=11, pockets=6)
    gpu_ctx.set_materials(table)
    ga, gf = fse.World(gpu_ctx, W, H), fse.World(gpu_ctx, W, H)
    ow = oracle.OracleWorld(W, H, table)
    for w in (ga, gf, ow):
        w.write_rect(0, 0, cells)
    ga.active_enable(True)
    awake0, total = ga.active_stats()
    assert awake0 == total == (W // 128) * (H // 128)
    for t in range(48):
        for w in (ga, gf, ow):
            w.tick(t)
            w.particles_tick()
            if t % 4 == 2:
                w.tick_temperature()
        if t % 8 == 7:
            a, f, o = ga.read_all(), gf.read_all(), ow.read_all()
            for fld in STATE_FIELDS:
                assert a[fld].tobytes() == f[fld].tobytes() == o[fld].tobytes(), (t, fld)
            Hh.assert_particles_equal(ga.particles_read(), gf.particles_read(), f"active tick {t}")
    awake, total = ga.active_stats()
    zone_chunks = (W // 128 - 2) * (H // 128 - 2)
    awake_in_zone = awake - (total - zone_chunks)  # border chunks are never ticked, so they never fall asleep
    assert 6 <= awake_in_zone <= 6 * 9, (awake_in_zone, zone_chunks)  # the 6 pockets (+ at most their 3x3 neighbourhoods)
    # waking: drop a sand block into a sleeping region through the public write path and check it falls
    blk = G.cells_from_mat(table, np.full((8, 8), 2, dtype=np.uint16), 700, 300)
    air = G.cells_from_mat(table, np.zeros((60, 40), dtype=np.uint16), 690, 290)
    for w in (ga, gf):
        w.write_rect(690, 290, air)
        w.write_rect(700, 300, blk)
    for t in range(48, 60):
        for w in (ga, gf):
            w.tick(t)
            w.particles_tick()
    a, f = ga.read_all(), gf.read_all()
    for fld in STATE_FIELDS:
        assert a[fld].tobytes() == f[fld].tobytes(), fld
    assert (a["mat"][300:308, 700:708] == 2).sum() < 64


def test_stats_match(oracle, gpu_ctx, table):
    W = H = 512
    ow, gw = _pair(oracle, gpu_ctx, table, W, H)
    Hh.build_mixed(ow, table, W, H, seed=21)
    Hh.build_mixed(gw, table, W, H, seed=21)
    for t in range(5):
        ow.tick(t)
        gw.tick(t)
    so, sg = ow.stats(), gw.stats()
    assert so.hash == sg.hash
    assert list(so.count) == list(sg.count)
    assert so.n_dirty == sg.n_dirty and so.n_moved == sg.n_moved
    for m in range(table.n):
        assert abs(so.fluid_mass[m] - sg.fluid_mass[m]) <= 1e-6 * max(1.0, abs(so.fluid_mass[m]))


def test_temperature_exact(oracle, gpu_ctx, table):
    W, H = 512, 384
    ow, gw = _pair(oracle, gpu_ctx, table, W, H)
    Hh.build_mixed(ow, table, W, H, seed=4, blob=8)
    Hh.build_mixed(gw, table, W, H, seed=4, blob=8)
    for _ in range(6):
        ow.tick_temperature()
        gw.tick_temperature()
    Hh.assert_cells_equal(ow.read_all(), gw.read_all(), "temperature")


def test_determinism(gpu_ctx, table):
    W = H = 512
    hashes = []
    for _ in range(2):
        gpu_ctx.set_materials(table)
        gw = fse.World(gpu_ctx, W, H)
        Hh.build_mixed(gw, table, W, H, seed=8)
        for t in range(10):
            gw.tick(t)
        hashes.append(gw.stats().hash)
        gw.close()
    assert hashes[0] == hashes[1]


def test_phase_parts_and_longest_first_order_do_not_change_results(oracle, gpu_ctx, table, monkeypatch):
    """Large worlds cut a colour phase into parts on separate streams and launch the chunks longest-first (last tick's
    pass-1 cycles).  Chunks of a phase are independent, so neither may change a single bit: force both on a small world."""
    monkeypatch.setenv("FSE_TICK_MIN_CHUNKS", "1")
    W = H = 1280
    tbl, extra = G.bench_table(table)
    ow, gw = _pair(oracle, gpu_ctx, tbl, W, H, "rows")
    Hh.build_mixed(ow, tbl, W, H, seed=5, extra=list(extra.values()))
    Hh.build_mixed(gw, tbl, W, H, seed=5, extra=list(extra.values()))
    _run_and_compare(ow, gw, 6, seed=3, what="parts + longest-first")


def test_benched_kernel_path_at_more_than_one_wave(oracle, gpu_ctx, table, monkeypatch):
    """The path bench.py times, with no override: a 4096 x 2304 world has 480 zone chunks, 120 per colour phase... too few for
    the default threshold, so the world is 5376 x 4352 (40 x 32 zone chunks, 320 per phase > 2 x 148): every phase runs
    tick_pass_kernel<1> / <2> / tick_pass3_kernel cut into 3 parts on 3 streams, chunks in longest-first order from the second
    tick on, more than one wave of CTA slots on pass 1.  Whole-grid state hash, per-material counts, dirty / moved totals and the
    emitted-particle count must equal the oracle's ROWS schedule; the full planes are compared on a 1024-row band as well."""
    for k in ("FSE_FUSED_MAX_CHUNKS", "FSE_TICK_MIN_CHUNKS", "FSE_TICK_PARTS", "FSE_TICK_LPT"):
        monkeypatch.delenv(k, raising=False)
    W, H = 5376, 4352
    tbl, extra = G.bench_table(table)
    ow, gw = _pair(oracle, gpu_ctx, tbl, W, H, "rows")
    gw.particles_reserve(1 << 23)  # a fresh mixed world drops millions of grains in its first tick (the reference's vector just grows)
    Hh.build_mixed(ow, tbl, W, H, seed=1337, extra=list(extra.values()), blob=64)
    Hh.build_mixed(gw, tbl, W, H, seed=1337, extra=list(extra.values()), blob=64)
    for t in range(3):
        n0 = gpu_ctx.launch_count()
        ow.tick(t, seed=1337)
        gw.tick(t, seed=1337)
        assert gpu_ctx.launch_count() - n0 >= 12 * 9  # 12 phases x 3 parts x 3 pass kernels (+ the ordering kernels)
        so, sg = ow.stats(), gw.stats()
        assert so.hash == sg.hash, t
        assert list(so.count) == list(sg.count) and so.n_dirty == sg.n_dirty and so.n_moved == sg.n_moved
        assert ow.particles_count() == gw.particles_count()
    Hh.assert_cells_equal(ow.read_rect(0, 1500, W, 1024), gw.read_rect(0, 1500, W, 1024), "band of the large world")


@pytest.mark.parametrize("split", ["0", "2"])
def test_pass2_split_modes_exact(oracle, gpu_ctx, table, monkeypatch, split):
    """Pass 2 as one chain of row steps (FSE_P2_SPLIT=0) or split into a row-parallel liquid apply + the rows that still hold powder or
    gas (FSE_P2_SPLIT=2: in every phase; the default uses it only where few rows hold live powder): bit-identical to the oracle either
    way, on a world with liquids over gas pockets, steam made over lava in mid-tick and sand sliding into liquid."""
    monkeypatch.setenv("FSE_P2_SPLIT", split)
    monkeypatch.setenv("FSE_ROW_SKIP", "0")  # no settled-row skipping: the phases the split is for
    W, H = 768, 640
    tbl, extra = G.bench_table(table)
    ow, gw = _pair(oracle, gpu_ctx, tbl, W, H, "rows")
    Hh.build_mixed(ow, tbl, W, H, seed=19, extra=list(extra.values()), blob=16)
    Hh.build_mixed(gw, tbl, W, H, seed=19, extra=list(extra.values()), blob=16)
    for t in range(12):
        for w in (ow, gw):
            w.tick(t, seed=5)
            w.particles_tick()
        if t % 3 == 2:
            Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"split {split} tick {t}")
            Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"split {split} tick {t}")


def test_small_phases_pick_the_fused_kernel_with_identical_results(oracle, gpu_ctx, table, monkeypatch):
    """Default kernel selection (no override): a 640x512 world has 2-3 chunks per colour phase, far below one wave of the fused
    kernel, so the rows schedule runs there; results must not depend on the choice."""
    monkeypatch.delenv("FSE_FUSED_MAX_CHUNKS", raising=False)
    W, H = 640, 512
    ow, gw = _pair(oracle, gpu_ctx, table, W, H, "rows")
    Hh.build_mixed(ow, table, W, H, seed=17)
    Hh.build_mixed(gw, table, W, H, seed=17)
    _run_and_compare(ow, gw, 3, seed=5, what="automatic kernel choice")
    n0 = gpu_ctx.launch_count()
    gw.tick(3, seed=5)
    assert gpu_ctx.launch_count() - n0 == 12  # one (fused) kernel per colour phase


@pytest.mark.parametrize("seed", [11, 23, 37])
@pytest.mark.parametrize("sched", ["rows", "rows_fused"])
def test_random_worlds_long_runs(oracle, gpu_ctx, table, sched, seed):
    """Longer runs of the whole game loop (tick + tickCells + tickTemperature every 4th tick + camera scroll + render planes) on
    small-blob worlds of different seeds and shapes: every plane, the particle pool and the textures stay bit-identical."""
    from oracle import pyoracle as O
    W, H = 512 + 128 * (seed % 3), 384 + 128 * (seed % 2)
    tbl, extra = G.bench_table(table)
    ow, gw = _pair(oracle, gpu_ctx, tbl, W, H, sched)
    Hh.build_mixed(ow, tbl, W, H, seed=seed, extra=list(extra.values()), blob=16)
    Hh.build_mixed(gw, tbl, W, H, seed=seed, extra=list(extra.values()), blob=16)
    gw.pixels_enable(True)
    planes = [np.zeros((H, W, 4), dtype=np.uint8) for _ in range(3)]
    probes = []
    for t in range(30):
        for w in (ow, gw):
            w.tick(t, seed=seed)
        # the physicsCheck probe at the end of world::tick (world.cpp:1929-1934), several per tick here so that some hit SOLID blobs
        for k in range(6):
            a, b = O.physics_probe(ow, 100 * t + k, seed), gw.physics_probe(100 * t + k, seed)
            assert a[:3] == b[:3], (t, k, a[:3], b[:3])
            if a[3] is not None:
                Hh.assert_cells_equal(a[3], b[3], f"cut-out tiles, tick {t}")
            probes.append(a[1])
        for w in (ow, gw):
            w.particles_tick()
            if t % 4 == 2:
                w.tick_temperature()
        if t == 13:
            O.scroll(ow, -128, 0)
            gw.scroll(-128, 0)
        O.render_dirty(ow, planes)
        gw.render_dirty(want_stats=False)
        ow.clear_dirty()
        gw.clear_dirty()
        if t % 6 == 5:
            Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"{sched} seed {seed} tick {t}")
            Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"{sched} seed {seed} tick {t}")
            for which in range(3):
                assert np.array_equal(planes[which], gw.pixels_read(which)), (t, which)
    assert 1 in probes or 2 in probes, "no probe of the run hit a loose SOLID component"


def test_chunk_save_and_load_through_pack_files(gpu_ctx, table, tmp_path):
    """World.save_chunk / load_chunk: chunkSaveCache + ChunkWrite and ChunkRead + the frame() merge (world.cpp:2374-2391, 2780-2792;
    chunk.cpp:74-330).  A ticked chunk written to a .pack file and merged into another world carries material, colour and
    temperature; the merged cells are dirty and have the reference's defaults for the per-tick fields."""
    W = H = 512
    gpu_ctx.set_materials(table)
    a, b = fse.World(gpu_ctx, W, H), fse.World(gpu_ctx, W, H)
    Hh.build_mixed(a, table, W, H, seed=31)
    for t in range(3):
        a.tick(t)
    path = str(tmp_path / "c.pack")
    a.save_chunk(path, 128, 256, generation_phase=2)
    phase, layer2, background = b.load_chunk(path, 256, 128)
    assert phase == 2 and not layer2["mat"].any() and not background.any()
    src, dst = a.read_rect(128, 256, 128, 128), b.read_rect(256, 128, 128, 128)
    for f in ("mat", "color", "temp"):
        assert np.array_equal(src[f], dst[f]), f
    assert dst["dirty"].all() and (dst["fluid"] == 2.0).all() and not dst["moved"].any()
    # with the layer planes: layer-2 cells and background colours travel through the file into the other world's planes
    l2 = np.zeros((128, 128), dtype=T.CELL_DTYPE)
    l2["mat"][10:20, 5:50], l2["color"][10:20, 5:50], l2["temp"][10:20, 5:50] = 7, 0x334455, 12
    bg = (np.arange(128 * 128, dtype=np.uint32).reshape(128, 128) * np.uint32(40503)) | np.uint32(0xFF000000)
    a.layer2_write_rect(128, 256, l2)
    a.background_write_rect(128, 256, bg)
    a.save_chunk(path, 128, 256, layers=True)
    b.load_chunk(path, 256, 128, layers=True)
    got = b.layer2_read_rect(256, 128, 128, 128)
    for f in ("mat", "color", "temp"):
        assert np.array_equal(got[f], l2[f]), f
    assert got["dirty"].all() and np.array_equal(b.background_read_rect(256, 128, 128, 128), bg)
    assert b.render_layers() == (128 * 128, 128 * 128)
    d = str(tmp_path / "world")
    assert a.save_world(d) == 16 and b.load_world(d) == 16
    assert np.array_equal(b.layer2_read_rect(128, 256, 128, 128)["color"], l2["color"])
    a.close()
    b.close()

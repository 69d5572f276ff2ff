"""North-star tolerance for the order-dependent rule families, ROWS (the product's in-row schedule) against REFERENCE (the
reference's left-to-right in-place scan), both run by the CPU oracle with the same counter RNG.  CPU only.

BASELINE.json north_star: "for rules whose outcome depends on the reference's in-chunk scan order, results must be exact in
per-material cell counts and mass conservation and match settled-state histograms; that tolerance is stated explicitly".
Every test below states its tolerance next to the assertion.  tests/test_oracle_pins.py holds the sand-pile pin
(world.cpp:1602-1727); this file adds liquids (1269-1537, 1728-1745), gas (1569-1585, 1799-1819, 1862-1890), fire (1101-1146)
and pair interactions (1153-1179).

Loose particles are integrated with the reference's list-order loop in every arm, so that only the in-row schedule of
world::tick differs between the arms (the particle deposit schedule has its own pins in tests/test_oracle_pins.py).
"""
import numpy as np
import pytest

from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh

AIR, SAND, GAS, STONE, GRASS, DIRT, WATER, LAVA, STEAM, FIRE = 0, 2, 4, 7, 8, 9, 15, 16, 23, 25
SEEDS = (1, 2, 3)


def _world(oracle, table, W, H):
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, Hh.empty_world_cells(table, W, H))
    return ow


def _put(ow, table, x, y, mat, w=1, h=1, **fields):
    c = G.cells_from_mat(table, np.full((h, w), mat, dtype=np.uint16), x, y)
    for k, v in fields.items():
        c[k] = v
    ow.write_rect(x, y, c)


def _run(oracle, ow, ticks, seed, sched, each=None):
    for t in range(ticks):
        ow.tick(t, seed=seed, schedule=sched)
        ow.particles_tick(schedule=oracle.REFERENCE)
        if each:
            each(t)


def _count(ow, mat):
    return int((ow.read_all()["mat"] == mat).sum()) + int((ow.particles_read()["tile"]["mat"] == mat).sum())


def _mass(ow, mat):
    c, p = ow.read_all(), ow.particles_read()
    sel = c["mat"] == mat
    grid = float((c["fluid"][sel].astype(np.float64) + c["fluid_diff"][sel].astype(np.float64)).sum())
    return grid + float(p["tile"]["fluid"][p["tile"]["mat"] == mat].astype(np.float64).sum())


@pytest.mark.parametrize("sched", ["PARTITIONED", "ROWS"])
@pytest.mark.parametrize("seed", [1, 2])
def test_conservation_under_the_gpu_schedules(oracle, table, sched, seed):
    """SURVEY §8c pin 3 under the product's schedules (tests/test_oracle_pins.py runs it under REFERENCE): a column-drop world of
    sand, water and stone.  Exact: the sand count (grid + loose particles) never changes.  Liquid: the water mass
    sum(fluidAmount + fluidAmountDiff) over grid cells + particles never grows by more than float round-off (1e-3 of ~1e4)
    and only shrinks through the FLUID_MinValue sinks (world.cpp:1277-1281, 1343-1347, 1732-1735)."""
    W = H = 512
    ow = oracle.OracleWorld(W, H, table)
    Hh.build_column(ow, table, W, H, seed=seed)
    s = getattr(oracle, sched)
    sand0, prev = _count(ow, SAND), _mass(ow, WATER)
    m0 = prev
    for t in range(25):
        ow.tick(t, seed=seed, schedule=s)
        ow.particles_tick(schedule=oracle.PARTITIONED)
        assert _count(ow, SAND) == sand0, t
        m = _mass(ow, WATER)
        assert m <= prev + 1e-3, (t, m, prev)
        prev = m
    assert prev > 0.5 * m0


def _basin(oracle, table, sched, seed, ticks=300):
    W, H = 768, 384
    ow = _world(oracle, table, W, H)
    _put(ow, table, 128, 300, STONE, w=512, h=4)
    _put(ow, table, 128, 180, STONE, w=4, h=120)
    _put(ow, table, 636, 180, STONE, w=4, h=120)
    _put(ow, table, 150, 200, WATER, w=100, h=80, fluid=0.5)
    masses = []
    _run(oracle, ow, ticks, seed, sched, each=lambda t: masses.append(_mass(ow, WATER)) if t % 20 == 0 else None)
    reg = ow.read_all()[180:300, 132:636]
    wat = reg["mat"] == WATER
    return wat.sum(axis=0), int(wat.sum()), masses


def test_liquid_basin_surface_profile(oracle, table):
    """SURVEY §8c pin 5, liquids: a 100 x 80 block of water released at the left end of a 504-wide stone basin; after 300 ticks
    the per-column water height profile (cells of WATER per column) under ROWS is compared with REFERENCE.
    Tolerance: total WATER cells within 2 %; mean, standard deviation and maximum of the column heights within 0.5 / 0.5 / 2 cells;
    the profile itself within 3 cells in every column after a 9-column box filter (local scan-order jitter); the mass is
    non-increasing under both schedules."""
    prof = {}
    for sched in (oracle.REFERENCE, oracle.ROWS):
        h, n, masses = _basin(oracle, table, sched, seed=1)
        assert all(b <= a + 1e-3 for a, b in zip(masses, masses[1:])), masses
        prof[sched] = (h.astype(np.float64), n)
    (h0, n0), (h1, n1) = prof[oracle.REFERENCE], prof[oracle.ROWS]
    assert n0 > 1000 and abs(n0 - n1) <= 0.02 * n0, (n0, n1)
    assert abs(h0.mean() - h1.mean()) <= 0.5 and abs(h0.std() - h1.std()) <= 0.5 and abs(h0.max() - h1.max()) <= 2
    k = np.ones(9) / 9
    assert np.abs(np.convolve(h0, k, "same") - np.convolve(h1, k, "same")).max() <= 3.0


def _plume(oracle, table, sched, seed, mat, ticks=60):
    W, H = 384, 512
    ow = _world(oracle, table, W, H)
    _put(ow, table, 140, 128, STONE, w=104, h=4)
    _put(ow, table, 140, 380, STONE, w=104, h=4)
    _put(ow, table, 140, 128, STONE, w=4, h=256)
    _put(ow, table, 240, 128, STONE, w=4, h=256)
    _put(ow, table, 170, 350, mat, w=40, h=20)
    _run(oracle, ow, ticks, seed, sched)
    c = ow.read_all()["mat"][132:380, 144:240]
    ys, xs = np.nonzero(c == mat)
    return len(ys), ys.mean(), ys.std(), xs.std(), np.histogram(ys, bins=8, range=(0, 248))[0] / max(len(ys), 1)


@pytest.mark.parametrize("mat", [GAS, STEAM])
def test_gas_plume_vertical_distribution(oracle, table, mat):
    """Gas rules (pass 1 rise, pass 2 diagonal, pass 3 sideways + STEAM condensation 1/10): an 800-cell block released at the
    bottom of a sealed 96 x 248 chamber, 60 ticks, 3 seeds.  GENERIC_GAS never transforms: its cell count is exact.  Tolerance on
    3-seed means, ROWS vs REFERENCE: mean height of the plume within 6 rows (it has risen ~60), vertical and horizontal spread
    (standard deviations) within 4 cells, every bin of the 8-bin vertical histogram within 0.10 of the plume; for STEAM the
    surviving cell count within 20 % (condensation needs a boxed-in cell, which depends on the plume's density)."""
    res = {}
    for sched in (oracle.REFERENCE, oracle.ROWS):
        rs = [_plume(oracle, table, sched, seed, mat) for seed in SEEDS]
        if mat == GAS:
            assert all(r[0] == 800 for r in rs)
        res[sched] = [np.mean([r[i] for r in rs], axis=0) for i in range(5)]
    a, b = res[oracle.REFERENCE], res[oracle.ROWS]
    assert abs(a[0] - b[0]) <= 0.2 * a[0], (a[0], b[0])
    assert abs(a[1] - b[1]) <= 6.0, (a[1], b[1])
    assert abs(a[2] - b[2]) <= 4.0 and abs(a[3] - b[3]) <= 4.0, (a[2:4], b[2:4])
    assert np.abs(a[4] - b[4]).max() <= 0.10, (a[4], b[4])
    assert a[1] < 200 and b[1] < 200  # the plume did rise from y ~ 228 of the chamber


def _fire(oracle, table, sched, seed, ticks=80):
    W = H = 384
    ow = _world(oracle, table, W, H)
    rng = np.random.default_rng(100 + seed)
    block = np.full((60, 80), STONE, dtype=np.uint16)
    block[rng.random(block.shape) < 0.04] = FIRE  # embers inside a porous block: overlapping 5 x 5 neighbourhoods
    ow.write_rect(150, 180, G.cells_from_mat(table, block, 150, 180))
    _put(ow, table, 150, 178, FIRE, w=80, h=2)
    _run(oracle, ow, ticks, seed, sched)
    m = ow.read_all()["mat"][170:250, 140:240]
    return int((block == STONE).sum()) - int((m == STONE).sum()), int((m == FIRE).sum())


def test_fire_spread_burnt_cell_count(oracle, table):
    """FIRE (world.cpp:1101-1146): ignition of SOLID cells in the 5 x 5 neighbourhood w.p. 1/500 each, burn-out 1/150 (1/120
    alone).  A 80 x 60 stone block with a burning top edge and 4 % embers inside, 80 ticks, 3 seeds.  Tolerance on 3-seed means,
    ROWS vs REFERENCE: burnt SOLID cells within 10 %, live FIRE cells within 10 % (the only order dependence is which of two fire
    cells wins a contested neighbour and tickVisited on freshly ignited cells)."""
    res = {}
    for sched in (oracle.REFERENCE, oracle.ROWS):
        rs = [_fire(oracle, table, sched, seed) for seed in SEEDS]
        res[sched] = np.mean(rs, axis=0)
    a, b = res[oracle.REFERENCE], res[oracle.ROWS]
    assert a[0] > 300, a  # it did burn
    assert abs(a[0] - b[0]) <= 0.10 * a[0] and abs(a[1] - b[1]) <= 0.10 * a[1], (a, b)


def _interactions(oracle, table, sched, seed, ticks=40):
    tbl, extra = G.bench_table(table)
    W, H = 640, 384
    ow = oracle.OracleWorld(W, H, tbl)
    ow.write_rect(0, 0, Hh.empty_world_cells(tbl, W, H))
    _put(ow, tbl, 140, 220, STONE, w=330, h=40)
    _put(ow, tbl, 150, 214, extra["ACID"], w=60, h=6)   # eats the stone below it: TRANSFORM STONE -> GENERIC_SAND, box radius 1 at (0, +1)
    _put(ow, tbl, 260, 220, DIRT, w=100, h=8)
    _put(ow, tbl, 270, 218, extra["SEED"], w=60, h=2)   # sprouts on dirt: SPAWN GRASS into AIR, box radius 1 at (0, -2)
    _put(ow, tbl, 380, 214, WATER, w=80, h=6, fluid=0.5)
    _put(ow, tbl, 380, 214, STONE, w=1, h=6)
    _put(ow, tbl, 459, 214, STONE, w=1, h=6)
    _put(ow, tbl, 390, 212, extra["SALT"], w=60, h=1)   # fizzes on water: SPAWN STEAM, box radius 1 at (+1, -2)
    _run(oracle, ow, ticks, seed, sched)
    return [_count(ow, m) for m in (SAND, GRASS, STEAM, extra["ACID"], extra["SEED"], extra["SALT"], STONE)]


def test_pair_interaction_product_counts(oracle, table):
    """Pair interactions (world.cpp:1153-1179) with the bench table's registered powders: ACID on STONE (TRANSFORM), SEED on
    DIRT and SALT on WATER (SPAWN).  40 ticks, 3 seeds.  Exact under both schedules: the interacting powders themselves are
    never consumed (their counts stay 360 / 120 / 60).  Tolerance on 3-seed means, ROWS vs REFERENCE: GENERIC_SAND produced
    from stone within 5 %, remaining STONE within 1 %, GRASS sprouted within 25 %, STEAM cells alive within 35 % (steam
    keeps moving and condensing, its count at a given tick is the noisiest of the four)."""
    res = {}
    for sched in (oracle.REFERENCE, oracle.ROWS):
        rs = np.array([_interactions(oracle, table, sched, seed) for seed in SEEDS], dtype=np.float64)
        assert (rs[:, 3] == 360).all() and (rs[:, 4] == 120).all() and (rs[:, 5] == 60).all(), rs
        res[sched] = rs.mean(axis=0)
    a, b = res[oracle.REFERENCE], res[oracle.ROWS]
    assert a[0] > 1000 and a[1] > 5 and a[2] > 5, a  # every branch produced something
    assert abs(a[0] - b[0]) <= 0.05 * a[0], (a, b)
    assert abs(a[6] - b[6]) <= 0.01 * a[6], (a, b)
    assert abs(a[1] - b[1]) <= 0.25 * a[1] + 2, (a, b)
    assert abs(a[2] - b[2]) <= 0.35 * a[2] + 3, (a, b)

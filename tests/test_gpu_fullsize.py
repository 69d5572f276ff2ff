"""Size-independent properties of the CUDA path at BASELINE.json's full size (configs[1]: the 8192 x 8192 mixed world bench.py times).

The oracle finishes a 512^2 tick in seconds, not an 8192^2 one, so at full size the CUDA path is checked through properties that do
not need it (`pytest -m gpu`, through the C ABI):
  * every kernel path that is bit-exact against the oracle on small worlds (tests/test_gpu_parity.py) gives the SAME grid, particle
    pool and per-material counts at full size — per-pass kernels with the default gate, settled-row skipping forced on, the pass-2
    split forced on, the fused kernel — so a size-dependent defect (32-bit index, more than one wave of CTAs, phase parts on streams,
    longest-first order) cannot hide in one of them;
  * the same run twice gives the same state (determinism: counter RNG, no order-dependent atomics in the results);
  * a checksum of checksums: the state hash and the counts of the whole grid equal the sums over disjoint bands;
  * conservation on the column-drop world (configs[0] scaled to 8192^2): the powder is conserved exactly between grid and particle
    pool, liquid mass only shrinks (FLUID_MinValue sinks), as tests/test_oracle_pins.py pins for the reference order;
  * idempotence: a settled world (SOLID and AIR only) is a fixed point of the tick, with and without row skipping.
"""
import functools

import numpy as np
import pytest

import falling_sand_engine_b200 as fse
from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G

pytestmark = pytest.mark.gpu

N = 8192
SEED = 1337
ENV_KEYS = ("FSE_FUSED_MAX_CHUNKS", "FSE_TICK_MIN_CHUNKS", "FSE_TICK_PARTS", "FSE_TICK_LPT", "FSE_ROW_SKIP", "FSE_P2_SPLIT")


@pytest.fixture(scope="module")
def mixed_bands(table):
    """The bench world, generated once on the host (1 GB) and written into every world of this module."""
    tbl, extra = G.bench_table(table)
    bands = [G.mixed_band(tbl, N, N, y0, 1024, seed=SEED, extra=list(extra.values())) for y0 in range(0, N, 1024)]
    return tbl, bands


def _world(gpu_ctx, tbl, bands, monkeypatch, env, schedule=1):
    for k in ENV_KEYS:
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    gpu_ctx.set_materials(tbl)
    w = fse.World(gpu_ctx, N, N)
    w.set_schedule(schedule)
    w.particles_reserve(1 << 24)
    for i, b in enumerate(bands):
        w.write_rect(0, 1024 * i, b)
    return w


def _state(w):
    s = w.stats()
    return (s.hash, tuple(s.count), s.n_dirty, s.n_moved, w.particles_count())


def test_full_size_kernel_paths_agree(gpu_ctx, mixed_bands, monkeypatch):
    tbl, bands = mixed_bands
    variants = [
        ("per-pass kernels, default gate", {}, 1),
        ("the same again (determinism)", {}, 1),
        ("settled-row skipping forced on", {"FSE_ROW_SKIP": "1"}, 1),
        ("pass-2 split forced on", {"FSE_ROW_SKIP": "0", "FSE_P2_SPLIT": "2"}, 1),
        ("fused kernel", {}, 2),
    ]
    ref = None
    for name, env, sched in variants:
        w = _world(gpu_ctx, tbl, bands, monkeypatch, env, sched)
        states = []
        for t in range(3):
            w.tick(t, seed=SEED)
            w.particles_tick()
            states.append(_state(w))
        if ref is None:
            ref = states
            # checksum of checksums on the first variant: disjoint bands add up to the whole grid
            whole = w.stats()
            h, cnt, nd = 0, np.zeros(len(whole.count), dtype=np.int64), 0
            for y0 in range(0, N, 1024):
                s = w.stats(T.Rect(0, y0, N, 1024))
                h = (h + s.hash) & 0xFFFFFFFFFFFFFFFF
                cnt += np.array(s.count, dtype=np.int64)
                nd += s.n_dirty
            assert h == whole.hash and list(cnt) == list(whole.count) and nd == whole.n_dirty
            assert int(cnt.sum()) == N * N
            assert states[0][0] != states[1][0]  # the world does move
        else:
            for t, (a, b) in enumerate(zip(ref, states)):
                assert a[0] == b[0], f"{name}: state hash differs after tick {t}"
                assert a == b, f"{name}: counts / dirty / moved / particles differ after tick {t}"
        w.close()


def test_full_size_column_world_conserves(gpu_ctx, table, monkeypatch):
    for k in ENV_KEYS:
        monkeypatch.delenv(k, raising=False)
    ids = G._names(table)
    SAND, WATER = ids["GENERIC_SAND"], ids["WATER"]
    gpu_ctx.set_materials(table)
    w = fse.World(gpu_ctx, N, N)
    w.particles_reserve(1 << 24)
    G.fill_world(w, functools.partial(G.column_drop_band, table, seed=SEED), N, N, band_rows=1024)
    s0 = w.stats()
    assert s0.count[SAND] > 0 and s0.fluid_mass[WATER] > 0
    prev = None
    for t in range(8):
        w.tick(t, seed=SEED)
        w.particles_tick()
        s = w.stats()
        p = w.particles_read()
        assert s.count[SAND] + int((p["tile"]["mat"] == SAND).sum()) == s0.count[SAND], t
        mass = s.fluid_mass[WATER] + float(p["tile"]["fluid"][p["tile"]["mat"] == WATER].astype(np.float64).sum())
        if prev is not None:
            assert mass <= prev * (1 + 1e-6) + 1e-3, t
        prev = mass
    assert prev > 0.5 * s0.fluid_mass[WATER]
    w.close()


@pytest.mark.parametrize("skip", ["0", "1"])
def test_full_size_settled_world_is_a_fixed_point(gpu_ctx, mixed_bands, monkeypatch, skip):
    tbl, bands = mixed_bands
    phys = np.array([m.physics for m in tbl.mats], dtype=np.int32)
    ids = G._names(tbl)
    for k in ENV_KEYS:
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("FSE_ROW_SKIP", skip)
    gpu_ctx.set_materials(tbl)
    w = fse.World(gpu_ctx, N, N)
    for i, b in enumerate(bands):
        mat = b["mat"].astype(np.uint16).copy()
        mat[(phys[mat] != T.SOLID) | (mat == ids["FIRE"])] = ids["AIR"]
        w.write_rect(0, 1024 * i, G.cells_from_mat(tbl, mat, 0, 1024 * i, SEED))
    w.clear_dirty()
    h0 = _state(w)
    for t in range(2):
        w.tick(t, seed=SEED)
    assert _state(w) == h0
    assert w.particles_count() == 0 and w.stats().n_dirty == 0
    w.close()

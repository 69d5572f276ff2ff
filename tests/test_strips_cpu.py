"""Host-side logic of the multi-GPU path on CPU: strip layout, and the per-phase halo protocol run by two / three gloo
ranks on oracle worlds must reproduce the single-process tick bit for bit (SURVEY.md §8c pin 8)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from falling_sand_engine_b200 import strips, types as T, worldgen as G
from tests import helpers as Hh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_strip_layout_covers_world():
    for H, n in ((1024, 2), (2048, 3), (8192, 8), (32768, 8), (640, 3)):
        prev = 0
        for r in range(n):
            lo, hi, hlo, hhi, j0, j1 = strips.strip_layout(H, r, n)
            assert lo == prev and hi > lo
            assert hlo <= lo and hhi >= hi and 0 <= hlo and hhi <= H
            if r > 0:
                assert (lo - 128) % 128 == 0 and lo - hlo == strips.GHOST
            assert j1 > j0
            prev = hi
        assert prev == H
    with pytest.raises(ValueError):
        strips.strip_layout(512, 0, 3)


def test_phase_messages_pair_up():
    H, n = 2048, 4
    lay = [strips.strip_layout(H, r, n) for r in range(n)]
    for ofy in (0, 1):
        msgs = {r: strips.phase_messages(r, n, lay[r][4], lay[r][5], ofy) for r in range(n)}
        for r in range(n):
            for peer, kind, ylo, yhi in msgs[r]:
                other = "recv" if kind == "send" else "send"
                assert (r, other, ylo, yhi) in msgs[peer], (r, peer, kind, ylo, yhi)


@pytest.mark.parametrize("nranks", [2, 3])
def test_strip_protocol_matches_single_world(oracle, table, tmp_path, nranks):
    W, H, ticks = 384, 896, 6
    out = str(tmp_path / "strip")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + nranks), WORLD_SIZE=str(nranks), OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "strip_cpu_worker.py"), str(W), str(H), str(ticks), out],
                              env=dict(env, RANK=str(r))) for r in range(nranks)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, G.mixed_band(table, W, H, 0, H, seed=77, blob=32))
    for t in range(ticks):
        ow.tick(t, seed=1337)
    ref = ow.read_all()
    rows = 0
    parts = []
    for r in range(nranks):
        lo, hi = strips.strip_layout(H, r, nranks)[:2]
        got = np.load(f"{out}.rank{r}.npy")
        Hh.assert_cells_equal(ref[lo:hi], got, f"strip {r}/{nranks}")
        rows += hi - lo
        parts.append(np.load(f"{out}.parts{r}.npy"))
    assert rows == H
    Hh.assert_particles_equal(ow.particles_read(), np.concatenate(parts), "strip particles")


@pytest.mark.parametrize("nranks", [2, 3])
def test_strip_particle_protocol_matches_single_world(oracle, table, tmp_path, nranks):
    """tick + tickCells over strips: particles live on the rank that owns their row, the deposit proposals that target the band
    around a cut are exchanged every round and committed by both sides (lowest id wins).  Grid and particle pool must equal the
    single-world tick + tick_particles_rounds (the schedule the CUDA path implements) bit for bit.  This is the host-side
    protocol for particle migration between GPUs; the CUDA path still refuses particles on multi-rank strips."""
    W, H, ticks = 384, 896, 8
    out = str(tmp_path / "pstrip")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29520 + nranks), WORLD_SIZE=str(nranks), OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "strip_particles_cpu_worker.py"), str(W), str(H), str(ticks), out],
                              env=dict(env, RANK=str(r))) for r in range(nranks)]
    for p in procs:
        assert p.wait(timeout=900) == 0
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, G.mixed_band(table, W, H, 0, H, seed=77, blob=32))
    deposited = 0
    for t in range(ticks):
        ow.tick(t, seed=1337)
        before = ow.particles_count()
        ow.particles_tick()
        deposited += before - ow.particles_count()
    assert deposited > 0  # the run actually settles particles, also next to the cuts
    ref = ow.read_all()
    parts = []
    for r in range(nranks):
        lo, hi = strips.strip_layout(H, r, nranks)[:2]
        Hh.assert_cells_equal(ref[lo:hi], np.load(f"{out}.rank{r}.npy"), f"strip {r}/{nranks}")
        parts.append(np.load(f"{out}.parts{r}.npy"))
    Hh.assert_particles_equal(ow.particles_read(), np.concatenate(parts), "strip particles")
    counts = sum(np.load(f"{out}.counts{r}.npy") for r in range(nranks))
    assert counts[0] > 0 and counts[1] > 0, counts  # proposals crossed a cut and particles changed owner during the run


@pytest.mark.parametrize("nranks", [2, 3])
def test_strip_box_edit_protocol_matches_single_world(oracle, table, tmp_path, nranks):
    """The protocol of the calls that edit the grid from one place on strips (rigid-body raster / erase, tickEntities), over gloo with
    oracle worlds: the product's own plan (`fse_strip_plan`, host code of libfse_b200.so) says who runs which body / entity and which
    rectangles travel; every rank holds only its window (junk elsewhere).  Grid, particles, feedback, body tiles and entity records
    must equal the single-world oracle — with bodies lying across the cuts, overlapping across them and drifting from strip to strip."""
    from oracle import pyoracle as O
    from tests.strip_bodies_scene import scene, entities
    W, H, ticks = 1024, 1024, 4
    out = str(tmp_path / "sbox")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29540 + nranks), WORLD_SIZE=str(nranks), OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "strip_bodies_cpu_worker.py"), str(W), str(H), str(ticks), out],
                              env=dict(env, RANK=str(r))) for r in range(nranks)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, G.mixed_band(table, W, H, 0, H, seed=21, air_frac=0.6, blob=48))
    bodies, xf = scene(table, W, H, nranks)
    ents = entities(H, nranks)
    fbs = []
    for t in range(ticks):
        ents = O.entities_tick(ow, ents, tick=t)
        fbs.append(O.bodies_raster(ow, bodies, xf, tick=t))
        fbs.append(O.bodies_erase(ow, bodies, xf))
        xf[:, 1] += 1.5
        xf[:, 2] += 0.05
    fbs = np.stack(fbs)
    assert fbs[:, :, 2].sum() > 0 and (fbs[0::2, :, 0] + fbs[0::2, :, 1]).sum() > 0
    ref = ow.read_all()
    tiles = np.concatenate([b.reshape(-1) for b in bodies])
    parts = []
    for r in range(nranks):
        lo, hi = strips.strip_layout(H, r, nranks)[:2]
        assert np.array_equal(fbs, np.load(f"{out}.fb{r}.npy")), f"feedback on rank {r}"
        assert tiles.tobytes() == np.load(f"{out}.tiles{r}.npy").tobytes(), f"body tiles on rank {r}"
        assert ents.tobytes() == np.load(f"{out}.ents{r}.npy").tobytes(), f"entities on rank {r}"
        Hh.assert_cells_equal(ref[lo:hi], np.load(f"{out}.rank{r}.npy"), f"strip {r}/{nranks}")
        parts.append(np.load(f"{out}.parts{r}.npy"))
    Hh.assert_particles_equal(ow.particles_read(), np.concatenate(parts), "strip particles")

"""Multi-GPU strips over NCCL vs the single-world oracle (bit-exact for any strip count; SURVEY.md §8c pin 8): world tick, loose
particles (migration between ranks + deposit rounds with the band proposals exchanged), temperature, an explosion and an eraser
stroke across the cuts, camera scrolls (horizontal, vertical — rows change ranks — and both at once), and the rigid-body bridge.
Needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_strips_gpu.py -m gpu`."""
import os
import subprocess
import sys

import numpy as np
import pytest

from falling_sand_engine_b200 import strips, worldgen as G
from tests import helpers as Hh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("nranks", [2, 4])
def test_strips_match_oracle(oracle, table, tmp_path, nranks):
    if _ngpu() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    W, H, ticks = 1024, 1536, 8
    out = str(tmp_path / "strip")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + nranks), os.path.join(ROOT, "tests", "strip_gpu_worker.py"), str(W), str(H), str(ticks), out]
    # small strips: force the large-world machinery (per-pass kernels in parts on streams, longest-first order of the interior chunks)
    env = dict(os.environ, FSE_TICK_MIN_CHUNKS="1", FSE_FUSED_MAX_CHUNKS="0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, G.mixed_band(table, W, H, 0, H, seed=77, blob=32))
    deposited = 0
    from oracle import pyoracle as O
    for t in range(ticks):
        ow.tick(t, seed=1337)
        if t == 2:
            O.explosion(ow, W // 2, H // 2 + 5, 40, tick=t, seed=1337)
            O.tool_erase_line(ow, 200, H // 4 - 30, 700, 3 * H // 4 + 20, 9)
        if t == 4:
            O.scroll(ow, -128, 0)
        if t == 5:
            O.scroll(ow, 0, 128)
        if t == 6:
            O.scroll(ow, 36, -97)
        before = ow.particles_count()
        ow.particles_tick()
        deposited += before - ow.particles_count()
        if t % 4 == 2:
            ow.tick_temperature()
    assert deposited > 0
    ref = ow.read_all()
    parts = []
    for k in range(nranks):
        lo, hi = strips.strip_layout(H, k, nranks)[:2]
        Hh.assert_cells_equal(ref[lo:hi], np.load(f"{out}.rank{k}.npy"), f"strip {k}/{nranks}")
        parts.append(np.load(f"{out}.parts{k}.npy"))
    Hh.assert_particles_equal(ow.particles_read(), np.concatenate(parts), "strip particles")


@pytest.mark.parametrize("nranks", [2, 4])
def test_strip_bodies_match_oracle(oracle, table, tmp_path, nranks):
    """The rigid-body bridge on multi-rank strips (game.cpp:1711-1815 raster, 1896-1983 erase, around the tick and tickCells): every rank
    makes every call; a body — with every body whose footprint box overlaps its own — is run by the rank that holds the box, the part of
    the box in a neighbour's rows travels there afterwards, feedback and the rewritten body tiles are shared over the ranks.  Grid,
    particle pool, feedback of every call and the tiles of every body must equal the single-world oracle on every rank."""
    if _ngpu() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    from oracle import pyoracle as O
    from tests.strip_bodies_scene import scene, stone_blocks, tool_calls, run_tool, entities
    W, H, ticks = 1024, 1536, 6
    out = str(tmp_path / "sb")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + nranks), os.path.join(ROOT, "tests", "strip_bodies_gpu_worker.py"), str(W), str(H), str(ticks), out]
    env = dict(os.environ, FSE_TICK_MIN_CHUNKS="1", FSE_FUSED_MAX_CHUNKS="0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, G.mixed_band(table, W, H, 0, H, seed=21, air_frac=0.6, blob=48))
    bodies, xf = scene(table, W, H, nranks)
    for (x0, y0, cells) in stone_blocks(table, H, nranks):
        ow.write_rect(x0, y0, cells)
    ob = [b.copy() for b in bodies]
    fbs, tools = [], []
    ents = entities(H, nranks)
    for t in range(ticks):
        ents = O.entities_tick(ow, ents, tick=t)
        O.entities_stamp(ow, ents, tick=t)
        fbs.append(O.bodies_raster(ow, ob, xf, tick=t))
        ow.tick(t, seed=1337)
        if t == 1:
            for call in tool_calls(H, nranks, t):
                tools.append(np.asarray(run_tool(ow, call, O), dtype=np.int64).reshape(-1))
        ow.particles_tick()
        if t == 3:
            tools.append(np.array([O.particles_vacuum_pull(ow, 500.0, 700.0)], dtype=np.int64))
        fbs.append(O.bodies_erase(ow, ob, xf))
        O.object_delete(ow)
        xf[:, 1] += 1.5
        xf[:, 2] += 0.05
    fbs, tools = np.stack(fbs), np.concatenate(tools)
    assert (tools[:-1] != 0).sum() > 50  # the pickaxe took pixels, the hammers cracked stone
    assert fbs[:, :, 2].sum() > 0 and (fbs[0::2, :, 0] + fbs[0::2, :, 1]).sum() > 0  # pixels were placed and grains / liquid were displaced
    ref = ow.read_all()
    tiles = np.concatenate([b.reshape(-1) for b in ob])
    parts = []
    for k in range(nranks):
        lo, hi = strips.strip_layout(H, k, nranks)[:2]
        assert np.array_equal(fbs, np.load(f"{out}.fb{k}.npy")), f"feedback on rank {k}"
        assert np.array_equal(tools, np.load(f"{out}.tools{k}.npy")), f"tool results on rank {k}"
        assert ents.tobytes() == np.load(f"{out}.ents{k}.npy").tobytes(), f"entities on rank {k}"
        assert tiles.tobytes() == np.load(f"{out}.tiles{k}.npy").tobytes(), f"body tiles on rank {k}"
        Hh.assert_cells_equal(ref[lo:hi], np.load(f"{out}.rank{k}.npy"), f"strip {k}/{nranks}")
        parts.append(np.load(f"{out}.parts{k}.npy"))
    Hh.assert_particles_equal(ow.particles_read(), np.concatenate(parts), "strip particles")

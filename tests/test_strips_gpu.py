"""Multi-GPU strips over NCCL vs the single-world oracle (bit-exact for any strip count; SURVEY.md §8c pin 8): world tick, loose
particles (migration between ranks + deposit rounds with the band proposals exchanged), temperature, an explosion and an eraser
stroke across the cuts, and a horizontal camera scroll.
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

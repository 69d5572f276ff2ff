"""The CUDA path against the committed regression vectors (tests/golden/world_hashes.json) without the oracle in the loop:
state hash, per-material counts, particle count and dirty / moved totals after the whole game loop.  `pytest -m gpu`."""
import json
import os

import numpy as np
import pytest

import falling_sand_engine_b200 as fse
from scripts import make_golden as MG

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "world_hashes.json")))


@pytest.mark.parametrize("case", MG.CASES, ids=[c[0] for c in MG.CASES])
@pytest.mark.parametrize("sname,sched", [("rows", 1), ("rows_fused", 2)])
def test_gpu_reproduces_golden(gpu_ctx, table, case, sname, sched):
    gpu_ctx.set_materials(table)
    gw = fse.World(gpu_ctx, case[1], case[2])
    if case[3] == "mixed_bench":  # the table has to be on the device before the world is filled
        from falling_sand_engine_b200 import worldgen as G

        tbl, _ = G.bench_table(table)
        gpu_ctx.set_materials(tbl)
    gw.set_schedule(sched)

    class _W:  # make_golden.build calls world.set_materials for the bench table; the context already has it
        def __getattr__(self, n):
            return getattr(gw, n)

        def set_materials(self, t):
            pass

    MG.build(case, _W(), table)
    got = MG.run(case, gw)
    assert got == GOLD["cases"][case[0]]["rows"]  # both kernel choices must reproduce the vectors of the rows schedule
    gw.close()

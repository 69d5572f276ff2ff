"""Writes tests/golden/world_hashes.json: state hashes, per-material counts and particle counts of small seeded worlds after
N ticks of the whole game loop under each schedule of the oracle (counter RNG).

These are REGRESSION vectors of this repository's own oracle, not outputs of the reference: the reference cannot be built here
(SDL2 / FMOD / xmake absent) and ships no tests or golden data for the tick, so the oracle stays "parity
unpinned" (DESIGN.md §6).  What the vectors pin: the oracle does not drift between rounds (tests/test_golden.py, CPU), and the
CUDA path reproduces them without the oracle in the loop (tests/test_gpu_golden.py).
usage: python scripts/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from falling_sand_engine_b200 import worldgen as G  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests import helpers as Hh  # noqa: E402

CASES = [  # name, W, H, generator, seed, ticks
    ("mixed_512x384", 512, 384, "mixed", 21, 12),
    ("mixed_interactions_640x512", 640, 512, "mixed_bench", 5, 10),
    ("column_512x512", 512, 512, "column", 0, 16),
]
SCHEDULES = {"reference": O.REFERENCE, "classes": O.PARTITIONED, "rows": O.ROWS}


def build(case, world, table):
    name, W, H, gen, seed, ticks = case
    if gen == "column":
        Hh.build_column(world, table, W, H)
        return table
    if gen == "mixed_bench":
        tbl, extra = G.bench_table(table)
        world.set_materials(tbl)
        Hh.build_mixed(world, tbl, W, H, seed=seed, extra=list(extra.values()), blob=16)
        return tbl
    Hh.build_mixed(world, table, W, H, seed=seed, blob=24)
    return table


def run(case, world, schedule=None):
    """tick + tickCells + tickTemperature on tick % 4 == 2 (game.cpp:2157); returns the record stored in the fixture."""
    name, W, H, gen, seed, ticks = case
    for t in range(ticks):
        if schedule is None:
            world.tick(t, seed=1337 + seed)
        else:
            world.tick(t, seed=1337 + seed, schedule=schedule)
        if schedule == O.REFERENCE:
            world.particles_tick(schedule=O.REFERENCE)
        else:
            world.particles_tick()
        if t % 4 == 2:
            world.tick_temperature()
    s = world.stats()
    counts = np.ctypeslib.as_array(s.count)
    return {"hash": f"{s.hash:016x}", "counts": {str(i): int(c) for i, c in enumerate(counts) if c}, "particles": int(world.particles_count()),
            "dirty": int(s.n_dirty), "moved": int(s.n_moved)}


def main():
    O.build()
    table = O.default_materials(1337)
    out = {"note": "regression vectors of this repository's oracle (not reference outputs); see scripts/make_golden.py", "cases": {}}
    for case in CASES:
        rec = {}
        for sname, sched in SCHEDULES.items():
            ow = O.OracleWorld(case[1], case[2], table)
            build(case, ow, table)
            rec[sname] = run(case, ow, sched)
            ow.close()
        out["cases"][case[0]] = rec
    with open(os.path.join(ROOT, "tests", "golden", "world_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps({k: {s: v[s]["hash"] for s in v} for k, v in out["cases"].items()}, indent=1))


if __name__ == "__main__":
    main()

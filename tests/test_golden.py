"""The oracle against the committed regression vectors (tests/golden/world_hashes.json, written by scripts/make_golden.py).
They are outputs of this repository's oracle, not of the reference (which cannot be built here): they pin the oracle against
drift between rounds; the CUDA path is checked against the same vectors in tests/test_gpu_golden.py."""
import json
import os

import pytest

from scripts import make_golden as MG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "world_hashes.json")))


@pytest.mark.parametrize("case", MG.CASES, ids=[c[0] for c in MG.CASES])
@pytest.mark.parametrize("sname", list(MG.SCHEDULES))
def test_oracle_reproduces_golden(oracle, table, case, sname):
    ow = oracle.OracleWorld(case[1], case[2], table)
    MG.build(case, ow, table)
    got = MG.run(case, ow, MG.SCHEDULES[sname])
    assert got == GOLD["cases"][case[0]][sname]


def test_golden_counts_cover_the_world():
    """Every schedule accounts for every cell of the grid (the per-material counts of a fixture add up to W x H)."""
    size = {c[0]: c[1] * c[2] for c in MG.CASES}
    for name, rec in GOLD["cases"].items():
        for sname, r in rec.items():
            assert sum(r["counts"].values()) == size[name], (name, sname)

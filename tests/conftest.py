import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The library runs colour phases that fit one wave of the fused kernel in that kernel (small worlds).  The test worlds are all
# that small, so force the per-pass kernels here ("rows" = per-pass, "rows_fused" = fused); one test removes the override.
os.environ.setdefault("FSE_FUSED_MAX_CHUNKS", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def table(oracle):
    return oracle.default_materials(1337)


@pytest.fixture(scope="session")
def gpu_ctx(table):
    """CUDA context of the product library; gpu tests fail (not skip) when the extension or the GPU is missing."""
    import falling_sand_engine_b200 as fse

    ctx = fse.Context(0, table)
    yield ctx
    ctx.close()

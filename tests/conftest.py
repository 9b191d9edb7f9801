import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_factory():
    from oracle.oracle import OracleEngine
    return OracleEngine


@pytest.fixture(scope="session")
def cuda_factory():
    """Engine factory for the CUDA engine; fails loudly when the library or a GPU is missing."""
    from examl_b200.capi import Engine

    def make(mxtips, nparts, nbranches=1, **kw):
        return Engine(mxtips, nparts, nbranches, **kw)
    return make

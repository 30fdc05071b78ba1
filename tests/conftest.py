import json
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_sessionstart(session):
    """The shared libraries are build artefacts (git-ignored). If a fresh checkout runs the tests before
    `__graft_entry__.build()`, build them here (nvcc cross-compiles without a GPU; ~3 min on 8 cores)."""
    from constantine_b200 import _lib
    from oracle import oracle
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    if not os.path.exists(oracle.LIB_PATH):
        oracle.build()


@pytest.fixture(scope="session")
def kat():
    """Golden vectors extracted from the reference's own tests (tests/golden/make_golden.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "msm_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    oracle.load()
    return oracle


@pytest.fixture()
def rng():
    return random.Random(0xC77)

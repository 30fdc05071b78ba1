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

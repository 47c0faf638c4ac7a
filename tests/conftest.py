import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle_py import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libref.so not built (no /root/reference at build time)")
    return Ref()


@pytest.fixture(scope="session")
def ctx():
    from roborts_edu_slam_b200 import matcher
    c = matcher.Context(0)   # raises loudly when the extension or the GPU is missing
    yield c
    c.close()


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(20261018)

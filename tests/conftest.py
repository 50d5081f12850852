import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref); skipped when it has not been built."""
    import oracle
    if not oracle.Ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle.Ref()


@pytest.fixture(scope="session")
def refz():
    import oracle
    if not oracle.Ref.available(zero_malloc=True):
        pytest.skip("oracle/_ref not built")
    return oracle.Ref(zero_malloc=True)


@pytest.fixture(scope="session")
def checker():
    """Strongest CPU checker available: the compiled reference if present, else the C port."""
    import oracle
    return oracle.Ref() if oracle.Ref.available() else oracle.Port()


@pytest.fixture(scope="session")
def ctx():
    from lssp_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def test_vector(n, k=0):
    i = np.arange(n, dtype=np.float64)
    return np.sin(i * (0.37 + 0.11 * k)) + 0.25 * np.cos(i * 1.3 + k)

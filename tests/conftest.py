import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_native_once():
    # the driver's CPU round check imports the package: make sure the .so files exist
    import __graft_entry__ as g

    g.build()


@pytest.fixture(scope="session", autouse=True)
def _built():
    _build_native_once()


@pytest.fixture(scope="session")
def port():
    from oracle import oracle

    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle

    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle.ref()


@pytest.fixture(scope="session")
def sa_golden():
    d = json.load(open(os.path.join(GOLDEN, "sa_golden.json")))
    return [(c["name"], bytes.fromhex(c["text_hex"]), np.array(c["sa"], dtype=np.int32)) for c in d["cases"]]


@pytest.fixture(scope="session")
def search_golden():
    return json.load(open(os.path.join(GOLDEN, "search_golden.json")))


def random_cases(seed=0, sizes=(3, 7, 8, 9, 15, 16, 17, 31, 33, 64, 65, 100, 257, 1000, 4097), sigmas=(1, 2, 3, 4, 5, 17, 256)):
    """Small random texts over assorted alphabets, NUL included half of the time."""
    rng = np.random.default_rng(seed)
    out = []
    for sig in sigmas:
        for n in sizes:
            alpha = rng.choice(256, size=sig, replace=False).astype(np.uint8)
            if rng.random() < 0.5:
                alpha[0] = 0
            out.append(alpha[rng.integers(0, sig, n)].tobytes())
    return out

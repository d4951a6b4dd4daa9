import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ora():
    import oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        oracle.build(ref=False)
    return oracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference compiled from /root/reference (only where it was built)."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle.Reference()


@pytest.fixture(scope="session")
def golden_cases():
    z = np.load(os.path.join(GOLDEN, "cases.npz"))
    names = sorted({k.rsplit(".", 1)[0] for k in z.files if not k.startswith("combine")})
    return z, names


@pytest.fixture(scope="session")
def golden_real():
    return np.load(os.path.join(GOLDEN, "real_pair.npz"))


@pytest.fixture(scope="session")
def golden_triple():
    """The three images the reference ships through its own SIFT and ExhaustiveMatching
    (tests/golden/make_golden.py::real_triple): BASELINE config 1 restated."""
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_triple.npz"))

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build (oracle/_ref) — the checker."""
    import flam3_b200 as fb

    p = os.path.join(ROOT, "oracle", "_ref", "libflam3_ref.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)")
    return fb.Flam3Lib(p)


@pytest.fixture(scope="session")
def core():
    import flam3_b200 as fb

    return fb.load_core()


@pytest.fixture(scope="session")
def gpu():
    """The drop-in libflam3 (flam3_render on the B200)."""
    import flam3_b200 as fb

    return fb.load_dropin()


@pytest.fixture(scope="session")
def flame0():
    return open(os.path.join(ROOT, "tests", "golden", "test_flame0.flam3")).read()


@pytest.fixture(scope="session")
def flame1():
    return open(os.path.join(ROOT, "tests", "golden", "test_flame1.flam3")).read()

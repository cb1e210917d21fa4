import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no binaries (they are git-ignored): build what is missing, exactly like
    ``__graft_entry__.build()`` does.  The product itself never builds or falls back on its own."""
    from focusflow_official_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    if not os.path.exists(os.path.join(ROOT, "oracle", "libpwc_ref.so")):
        import subprocess

        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN

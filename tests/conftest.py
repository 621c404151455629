import os
import sys

# Several GPU tests run a few emulated ranks (one context each) on ONE device, whose exchange kernels spin on each
# other.  With CUDA's default lazy module loading the first launch of a not-yet-loaded kernel synchronises the
# device - inside that window it would wait for a spinning peer and run into the exchange time-out (seen when
# a test subset is run with -k, so that earlier tests have not warmed the kernels).  Load everything up front.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    """One device context for the GPU tests (fails loudly without the CUDA library)."""
    import myzkp_b200

    c = myzkp_b200.Context(0)
    yield c
    c.close()

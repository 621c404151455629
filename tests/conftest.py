import os
import sys

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

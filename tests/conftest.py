import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/): the checker, never the thing under test."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def doa():
    import gr_doa_b200
    return gr_doa_b200


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False

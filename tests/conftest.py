import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/otb_oracle.c), compiled on first use.  Test infrastructure only."""
    from oracle import otb_oracle
    otb_oracle.build()
    return otb_oracle


@pytest.fixture(scope="session")
def golden():
    """Fixtures recorded from the unmodified reference by tools/make_golden.py."""
    def load(name):
        return np.load(GOLDEN / f"{name}.npz")
    return load


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


import torch

# the oracle's per-tile tensors are tiny: torch's intra-op pool thrashes on many-core hosts (a 128-core box ran the
# GPU suite 17x slower than an 8-thread one), so cap it
torch.set_num_threads(min(8, torch.get_num_threads()))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")

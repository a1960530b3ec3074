import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden(dict):
    def t(self, key):
        return torch.from_numpy(np.ascontiguousarray(self[key]))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return Golden({k: z[k] for k in z.files})


@pytest.fixture
def golden():
    return load_golden


def normwise(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| — the parity metric for fp32 bilinear sampling / scatter-mean."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

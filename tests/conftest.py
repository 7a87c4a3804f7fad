import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_layers(name, dtype=None):
    import torch
    w = load_golden("weights_" + name)
    dt = dtype or torch.float32
    layers = [(torch.tensor(w[f"W{i}"]).to(dt), torch.tensor(w[f"b{i}"]).to(dt)) for i in range(4)]
    emb = torch.tensor(w["emb"]).to(dt) if "emb" in w.files else None
    return layers, emb


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")

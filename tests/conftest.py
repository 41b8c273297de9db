import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


@pytest.fixture(scope="session")
def load_golden():
    return golden


@pytest.fixture(autouse=True)
def _inference_mode_like_the_reference():
    """The reference evaluates under @torch.no_grad() (evaluate.py:22,111; out.py); the kernels are
    inference-only and refuse tensors that require grad while autograd is enabled."""
    import torch
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    yield
    torch.set_grad_enabled(prev)

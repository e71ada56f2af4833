"""Shared pytest plumbing.

* registers the ``gpu`` marker (``-m gpu`` = parity tests proper, run on a B200;
  ``-m "not gpu"`` = oracle vs golden vectors, host logic, C-ABI export check, gloo sharding);
* puts the product package directory and the repo root on sys.path.  The product package lives in
  ``a-watermark-for-diffusion-models_b200/`` (not an importable name), so its importable
  child ``gswm`` is reached by path.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "a-watermark-for-diffusion-models_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_arrays():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_arrays.npz"))


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")

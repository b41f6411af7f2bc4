"""pytest configuration: markers, import paths, and shared fixtures.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol surface (runs anywhere).
`-m gpu`: parity of the CUDA path against the oracle and the golden vectors (needs a B200).
"""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_state_dict(g, prefix="sd."):
    return {k[len(prefix):]: g[k] for k in g if k.startswith(prefix)}


@pytest.fixture(scope="session")
def dvae():
    """The product package (directory name has a hyphen, so import it by string)."""
    return importlib.import_module("disentanglement-vae_b200")

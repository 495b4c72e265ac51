import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DIR = "/root/reference"  # exists only in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ckpt_prefix():
    from oracle.tf_bundle import default_checkpoint_prefix
    return default_checkpoint_prefix()


@pytest.fixture(scope="session")
def weights(ckpt_prefix):
    from oracle.tf_bundle import load_checkpoint
    return load_checkpoint(ckpt_prefix)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "suite64.npz"))


@pytest.fixture(scope="session")
def suite64():
    from oracle.roomnet_oracle import synthetic_suite
    return synthetic_suite(64)


@pytest.fixture(scope="session")
def capi():
    from roomnet_b200 import _capi
    return _capi


needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE_DIR), reason="/root/reference not present")

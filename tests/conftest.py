import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def cuda_backend():
    """The product path: bdm_b200.backend over libbdm_b200.so on cuda:0. No fallback."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bdm_b200 import backend
    return backend


@pytest.fixture(scope="session")
def ref_backend():
    """The reference's own CUDA extension, prebuilt into oracle/_ref by oracle/build_ref.py (optional)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import build_ref
    mod = build_ref.load_ref()
    if mod is None:
        pytest.skip("oracle/_ref/_pvcnn_backend.so not present")
    return mod

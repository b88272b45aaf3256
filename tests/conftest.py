import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _gpu_run(config):
    expr = config.getoption("-m") or ""
    return "gpu" in expr and "not gpu" not in expr


@pytest.fixture(scope="session", autouse=True)
def _backend(request):
    """CPU runs (-m "not gpu") exercise the host logic on the numpy checker backend from oracle/;
    GPU runs (-m gpu) use the product's CUDA backend and fail loudly if it is missing."""
    from tnsp_b200 import backend
    if _gpu_run(request.config):
        backend.set_backend(None)
        backend.get()  # raises without CUDA / the built library
    else:
        from oracle import numpy_backend
        numpy_backend.install()
    yield


@pytest.fixture(scope="session")
def ref_tat():
    from oracle.ref import load_reference_tat
    m = load_reference_tat()
    if m is None:
        pytest.skip("oracle/_ref reference build not available (make -C oracle ref)")
    return m

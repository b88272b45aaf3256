"""child process of test_reference_seam.py: the reference's OWN PyTAT test-suite (/root/reference/PyTAT/tests, unmodified) run
against this repository's TAT module installed under the name `TAT` (CPU checker backend for the float64 kernels)."""
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import numpy_backend  # noqa: E402

numpy_backend.install()
import tnsp_b200.TAT as T  # noqa: E402

T.install_as_TAT()
import pytest  # noqa: E402

sys.exit(pytest.main(["/root/reference/PyTAT/tests", "-q", "-p", "no:cacheprovider", "--no-header", "--tb=line"]))

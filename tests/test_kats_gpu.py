"""GPU run (-m gpu) of the reference's own known-answer vectors (test_reference_kats.py) through the
C-ABI / sm_100a kernels."""
import pytest

from test_reference_kats import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu

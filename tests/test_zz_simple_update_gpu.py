"""GPU run (-m gpu) of the simple-update parity test: fixtures of the unmodified reference, now through the
C-ABI / sm_100a kernels.  Added after the last GPU session of round 1 (no GPU minutes were left to run it), so the module sorts
last: under `-x` it cannot hide any other test."""
import pytest

from test_simple_update import test_simple_update_matches_the_reference as _case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["heis_3x3_D2_Dc4", "j1j2U1_4x4_d1_Dc9", "tJ_4x4_D1_Dc8"])
def test_simple_update_matches_the_reference(case):
    _case(case)

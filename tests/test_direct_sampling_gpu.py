"""GPU run (-m gpu) of the direct-sampling parity tests: the fixtures written by the unmodified reference
(tests/golden/direct_sampling.npz), now through the C-ABI / sm_100a kernels.  All five cases have been through a B200 run
(profiles/r01_final_direct_sampling_gpu_all.txt, command: scripts/gpu_runs/gpu_direct_all.sh)."""
import pytest

from test_direct_sampling import test_direct_sampling_matches_the_reference as _case

pytestmark = pytest.mark.gpu

GPU_CASES = ["heis_3x3_D2_Dc4", "heis_4x4_D3_Dc5_truncating", "heisU1_4x4_d1_Dc6", "tJ_4x4_D1_Dc8", "hubbardFF_4x4_D1_Dc8"]


@pytest.mark.parametrize("case", GPU_CASES)
def test_direct_sampling_matches_the_reference(case):
    _case(case)

"""-m gpu: the SURVEY 8f rows whose parity tests ran on the CPU checker only in round 1, now through the C-ABI / sm_100a kernels:
long-range observables through the configuration pool, gauge fixing / bond expansion, the Born-probability and lock-step direct-sampling
cases and the direct-sampling driver case (fixtures written by the unmodified reference).  14 passed on the B200 (round 2, call N).  The
pseudo-inverse SR formula test reads the SR rows as numpy arrays (a checker-side test; its two-rank form runs on hardware in
tests/test_nccl_gpu.py)."""
import pytest

from golden_loader import DRIVER_CASES
from test_direct_sampling import (test_direct_sampling_probability_is_the_born_probability,  # noqa: F401
                                  test_lockstep_direct_sampling_equals_independent_chains)
from test_gradient_driver import test_driver_matches_reference_loop as _driver_case
from test_gradient_driver import (test_bond_expansion_grows_the_bonds_and_perturbs_by_epsilon, test_driver_with_lockstep_direct_sampling,  # noqa: F401
                                  test_gauge_fixing_matches_the_reference)
from test_long_range import test_long_range_needs_the_configuration_cache, test_long_range_observables_match_the_reference  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [c for c in DRIVER_CASES if "direct" in c])
def test_direct_sampling_driver_matches_reference_loop(case):
    _driver_case(case)

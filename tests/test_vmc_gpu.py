"""GPU run (-m gpu) of the end-to-end VMC parity tests: golden vectors of the unmodified reference
and lock-step batching, now through the C-ABI / sm_100a kernels."""
import pytest

from test_vmc_batched import test_chain_rng_matches_reference_seed_recipe, test_lockstep_equals_independent_chains  # noqa: F401
from golden_loader import DRIVER_CASES
from test_gradient_driver import test_driver_matches_reference_loop as _driver_case
from test_gradient_driver import test_ergodic_driver_energy_is_exact_expectation, test_lockstep_driver_lowers_the_energy  # noqa: F401
from test_checkpoint import test_reference_checkpoint_loads_and_reproduces_amplitudes  # noqa: F401
from test_vmc_golden import test_amplitude_energy_holes, test_sweep_trajectory_gradient  # noqa: F401

pytestmark = pytest.mark.gpu


# the direct-sampling driver case was added after the last GPU session of round 1: it runs on the CPU checker only until it has
# been through a GPU run (a crash under `-x` would hide every later test)
@pytest.mark.parametrize("case", [c for c in DRIVER_CASES if "direct" not in c])
def test_driver_matches_reference_loop(case):
    _driver_case(case)

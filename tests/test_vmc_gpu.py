"""GPU run (-m gpu) of the end-to-end VMC parity tests: golden vectors of the unmodified reference
and lock-step batching, now through the C-ABI / sm_100a kernels."""
import pytest

from test_vmc_batched import test_chain_rng_matches_reference_seed_recipe, test_lockstep_equals_independent_chains  # noqa: F401
from test_gradient_driver import (test_driver_matches_reference_loop, test_ergodic_driver_energy_is_exact_expectation,  # noqa: F401
                                  test_lockstep_driver_lowers_the_energy)
from test_checkpoint import test_reference_checkpoint_loads_and_reproduces_amplitudes  # noqa: F401
from test_vmc_golden import test_amplitude_energy_holes, test_sweep_trajectory_gradient  # noqa: F401

pytestmark = pytest.mark.gpu

"""GPU run (-m gpu) of the charge-dense embedding parity tests (sector discovery on the device)."""
import pytest

from test_dense_embedding import (test_dense_reproduces_reference_fixture, test_lockstep_dense_equals_sector_evaluation,  # noqa: F401
                                  test_sweep_trajectory_dense_equals_reference_trajectory)

pytestmark = pytest.mark.gpu

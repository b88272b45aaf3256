"""-m gpu: fermionic sector-compact tensors and fermionic lock-step chains through the sm_100a kernels (signed regrouping,
csrc/ragged.cu rt_repack_signed_kernel): the same cases as tests/test_sector_fermi.py."""
import pytest

from test_sector_fermi import (test_fermionic_fixture_single_chain_sector_engine, test_fermionic_lockstep_batch_equals_one_by_one,  # noqa: F401
                               test_signed_operations_equal_block_symmetric_tensors)

pytestmark = pytest.mark.gpu

"""GPU parity run (-m gpu): the same differential tests as test_tat_vs_reference.py, but every
operation now goes planner -> C-ABI (libtnsp_b200.so) -> sm_100a kernels, compared with the
unmodified reference PyTAT (oracle/_ref) on identical inputs."""
import pytest

from test_tat_vs_reference import (test_block_layout_and_transpose, test_contract, test_merge_split_reverse, test_qr,  # noqa: F401
                                   test_svd)

pytestmark = pytest.mark.gpu

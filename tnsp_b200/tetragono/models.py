"""Synthetic lattices of the BASELINE configurations, built through the same
API a tetraku plugin uses (reference: tetraku/tetraku/models/heisenberg/__init__.py:22-58,
tetragono/tetragono/common_tensor/*)."""
from __future__ import annotations

import numpy as np

from .. import TAT
from .state import AbstractLattice, AbstractState, SamplingLattice


def spin_half_SS_array():
    """S.S of two spin-1/2 as an array [i0, i1, o0, o1]"""
    sx = np.array([[0, 0.5], [0.5, 0]])
    sz = np.array([[0.5, 0], [0, -0.5]])
    isy = np.array([[0, 0.5], [-0.5, 0]])  # i*Sy (real)
    return np.einsum("ac,bd->abcd", sx, sx) - np.einsum("ac,bd->abcd", isy, isy) + np.einsum("ac,bd->abcd", sz, sz)


def spin_half_SS(Tensor):
    """S.S as a tensor with names I0 I1 O0 O1 (reference: tetragono/common_tensor/No.py)"""
    t = Tensor(["I0", "I1", "O0", "O1"], [2, 2, 2, 2]).zero_()
    t.storage = spin_half_SS_array().reshape(-1)
    return t


def heisenberg_state(L1, L2, J=1.0):
    state = AbstractState(TAT.No.D.Tensor, L1, L2)
    state.physics_edges[...] = 2
    SS = spin_half_SS(TAT.No.D.Tensor)
    H = SS * (-J)
    state.hamiltonians["vertical_bond"] = H
    state.hamiltonians["horizontal_bond"] = H
    return state


def heisenberg_lattice(L1, L2, D, J=1.0):
    """cfg1 / dense stand-ins: NoSymmetry Heisenberg PEPS with bond dimension D"""
    state = AbstractLattice(heisenberg_state(L1, L2, J))
    state.virtual_bond["R"] = D
    state.virtual_bond["D"] = D
    return state


def neel_configuration(L1, L2):
    return np.array([[[(l1 + l2) % 2] for l2 in range(L2)] for l1 in range(L1)], dtype=np.int64)


def random_sampling_lattice(abstract, seed=2333):
    TAT.random.seed(seed)
    return SamplingLattice(abstract)

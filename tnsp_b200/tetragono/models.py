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


# ---------------------------------------------------------------------------------------------------
# generic (TAT-module independent) builders: the same functions build the model on this repository's
# device tensors and on the reference's PyTAT classes (reference arm of bench.py, golden generator)
# ---------------------------------------------------------------------------------------------------
def _is_u1(Tensor):
    return Tensor.model.__name__.rsplit(".", 1)[-1] in ("BoseU1", "U1")


def spin_half_SS_of(Tensor):
    """S.S for `Tensor`'s symmetry: NoSymmetry (dim-2 edges) or BoseU1 with physical edge
    [(+1,1),(-1,1)] (charge = 2 Sz); same matrix elements, same basis order (up, down)."""
    arr = spin_half_SS_array()
    if not _is_u1(Tensor):
        t = Tensor(["I0", "I1", "O0", "O1"], [2, 2, 2, 2]).zero_()
        t.storage = arr.reshape(-1)
        return t
    pe, cpe = [(+1, 1), (-1, 1)], [(-1, 1), (+1, 1)]
    t = Tensor(["I0", "I1", "O0", "O1"], [cpe, cpe, pe, pe]).zero_()
    q = (+1, -1)
    for i0 in range(2):
        for i1 in range(2):
            for o0 in range(2):
                for o1 in range(2):
                    if arr[i0, i1, o0, o1] != 0:
                        t[{"I0": (-q[i0], 0), "I1": (-q[i1], 0), "O0": (q[o0], 0), "O1": (q[o1], 0)}] = arr[i0, i1, o0, o1]
    return t


def j1j2_abstract_lattice(Tensor, L1, L2, D, J1=1.0, J2=0.5, state_classes=None):
    """J1-J2 Heisenberg model on the square lattice (cfg2; reference model tetraku/models/J1J2/__init__.py:22-68,
    which ships NoSymmetry only -- the BoseU1 variant follows SURVEY.md 8d: physical edge [(+1,1),(-1,1)],
    virtual edges [(-1,d),(0,d),(+1,d)] with D = 3 d, total symmetry 0).
    Returns (abstract lattice, sweep hopping hamiltonians = nearest-neighbour terms only, because the
    sweep-order builder rejects diagonal terms, sampling.py:156-190)."""
    AS, AL = state_classes if state_classes is not None else (AbstractState, AbstractLattice)
    u1 = _is_u1(Tensor)
    state = AS(Tensor, L1, L2)
    state.physics_edges[...] = [(+1, 1), (-1, 1)] if u1 else 2
    SS = spin_half_SS_of(Tensor)
    H1, H2 = SS * (-J1), SS * (-J2)
    state.hamiltonians["vertical_bond"] = H1
    state.hamiltonians["horizontal_bond"] = H1
    if J2 != 0:
        for l1 in range(L1 - 1):
            for l2 in range(L2 - 1):
                state.hamiltonians[(l1, l2, 0), (l1 + 1, l2 + 1, 0)] = H2
                state.hamiltonians[(l1, l2 + 1, 0), (l1 + 1, l2, 0)] = H2
    lat = AL(state)
    if u1:
        if D % 3:
            raise ValueError("U(1) J1-J2 lattice: D must be a multiple of 3 (segments -1, 0, +1)")
        ve = [(-1, D // 3), (0, D // 3), (+1, D // 3)]
    else:
        ve = D
    lat.virtual_bond["R"] = ve
    lat.virtual_bond["D"] = ve
    return lat


def nearest_neighbour_terms(lattice):
    return {k: v for k, v in lattice._hamiltonians.items() if len(k) == 1 or k[0][0] == k[1][0] or k[0][1] == k[1][1]}


def neel_points(lattice):
    """Neel configuration as edge points [l1][l2] -> {orbit: (symmetry, index)} for No / BoseU1 spin-1/2 lattices"""
    S = lattice.Tensor.model.Symmetry
    if _is_u1(lattice.Tensor):
        return [[{0: (S(+1) if (l1 + l2) % 2 == 0 else S(-1), 0)} for l2 in range(lattice.L2)] for l1 in range(lattice.L1)]
    return [[{0: (S(), (l1 + l2) % 2)} for l2 in range(lattice.L2)] for l1 in range(lattice.L1)]


# ---------------------------------------------------------------------------------------------------
# fermionic models of BASELINE cfg3 / cfg4
# ---------------------------------------------------------------------------------------------------
def tJ_abstract_state(L1, L2, T, t, J):
    """t-J model, symmetry (particle number: FermiU1, 2 Sz: BoseU1), T up and T down particles
    (tetraku/tetraku/models/tJ/__init__.py:22-45)"""
    from . import common_tensor
    op = common_tensor.FermiU1_tJ
    state = AbstractState(TAT.FermiU1BoseU1.D.Tensor, L1, L2)
    state.physics_edges[...] = op.EF
    H = (-t) * op.CC + (J / 2) * (op.SS - op.nn / 4)
    state.hamiltonians["vertical_bond"] = H
    state.hamiltonians["horizontal_bond"] = H
    state.total_symmetry = (T * 2, 0)
    return state


def tJ_abstract_lattice(L1, L2, D, T, t, J):
    """cfg4 family (tetraku/tetraku/models/tJ/__init__.py:48-80): the particles are fed in through column 0 (its vertical bonds
    carry the charge still to be distributed over the rows below, +- 2 particles of either spin) and along the rows; all other
    vertical bonds are trivial.  `D` is the dimension PER charge sector (9 sectors per charged bond)."""
    state = AbstractLattice(tJ_abstract_state(L1, L2, T, t, J))
    # `D`: dimension per charge sector, or the 9 per-sector dimensions in the order of the list below (BASELINE cfg4: D = 10 as
    # 1,1,1,1,2,1,1,1,1); an uncharged bond then carries their sum in its single sector
    profile = [D] * 9 if isinstance(D, int) else list(D)
    D = D if isinstance(D, int) else sum(profile)

    def charged(Q):
        sectors = [(dn, ds) for dn, spins in ((-2, (0,)), (-1, (-1, 1)), (0, (-2, 0, 2)), (1, (-1, 1)), (2, (0,))) for ds in spins]
        return [((2 * Q + dn, ds), d) for (dn, ds), d in zip(sectors, profile) if d > 0]

    per_row = T / L1
    for l1 in range(L1 - 1):
        state.virtual_bond[l1, 0, "D"] = charged(int(T * (L1 - l1 - 1) / L1))
        for l2 in range(1, L2):
            state.virtual_bond[l1, l2, "D"] = [((0, 0), D)]
    for l1 in range(L1):
        for l2 in range(L2 - 1):
            state.virtual_bond[l1, l2, "R"] = charged(int(per_row * (L2 - l2 - 1) / L2))
    return state


# BASELINE cfg3 (D = 8 on a bond with 9 possible charge fluctuations): two states without fluctuation, one for every single-particle
# fluctuation of either spin, one for each spin flip (n_up +- 1, n_down -+ 1); the double fluctuations (+-1, +-1) are left out
HUBBARD_D8 = {(0, 0): 2, (1, 0): 1, (-1, 0): 1, (0, 1): 1, (0, -1): 1, (1, -1): 1, (-1, 1): 1}
TJ_D10 = (1, 1, 1, 1, 2, 1, 1, 1, 1)


def staggered_fermion_configuration(lattice, per_row):
    """total physical indices [L1, L2, 1] of a start configuration with `per_row` = (pattern of length L2) cycled and shifted by
    one site from row to row; pattern entries are indices of the physical edge"""
    L1, L2 = lattice.L1, lattice.L2
    import numpy as np
    out = np.zeros((L1, L2, 1), dtype=np.int64)
    for l1 in range(L1):
        for l2 in range(L2):
            out[l1, l2, 0] = per_row[l1 % len(per_row)][l2]
    return out


def hubbard_fermi_fermi_abstract_state(L1, L2, T, t, U):
    """Hubbard model with symmetry FermiU1 (n_up) x FermiU1 (n_down), T particles in total, half of each spin
    (tetraku/tetraku/models/hubbard/fermi_fermi.py:22-47; cfg3 family)"""
    from . import common_tensor
    op = common_tensor.FermiFermi_Hubbard
    if T % 2:
        raise RuntimeError("T must be even number")
    state = AbstractState(TAT.FermiU1FermiU1.D.Tensor, L1, L2)
    state.total_symmetry = (T // 2, T // 2)
    state.physics_edges[...] = [((0, 0), 1), ((0, 1), 1), ((1, 0), 1), ((1, 1), 1)]
    state.hamiltonians["vertical_bond"] = -t * op.CSCS
    state.hamiltonians["horizontal_bond"] = -t * op.CSCS
    state.hamiltonians["single_site"] = U * op.NN
    return state


def hubbard_fermi_fermi_abstract_lattice(L1, L2, D, T, t, U):
    """cfg3 family.  The reference ships no lattice for this symmetry; the bonds follow the charge-flow pattern of its FermiU1
    Hubbard lattice (tetraku/tetraku/models/hubbard/__init__.py:47-75), as SURVEY.md 8d specifies: column-0 vertical bonds carry
    the (up, down) charge still to be distributed below, +- 1 of either spin, row bonds the row's share, the rest is trivial;
    `D` per charge sector (9 sectors per charged bond)."""
    state = AbstractLattice(hubbard_fermi_fermi_abstract_state(L1, L2, T, t, U))
    half = T // 2
    # `D`: dimension per charge sector, or {(a, b): dimension} for the fluctuations (a, b) in {-1, 0, 1}^2 of (n_up, n_down) around
    # the mean charge of the bond (BASELINE cfg3: D = 8, see HUBBARD_D8); an uncharged bond carries the sum in its single sector
    profile = {(a, b): D for a in (-1, 0, 1) for b in (-1, 0, 1)} if isinstance(D, int) else dict(D)
    D = D if isinstance(D, int) else sum(profile.values())

    def charged(Q):
        return [((Q + a, Q + b), profile[a, b]) for a in (-1, 0, 1) for b in (-1, 0, 1) if profile.get((a, b), 0) > 0]

    per_row = half / L1
    for l1 in range(L1 - 1):
        state.virtual_bond[l1, 0, "D"] = charged(int(half * (L1 - l1 - 1) / L1))
        for l2 in range(1, L2):
            state.virtual_bond[l1, l2, "D"] = [((0, 0), D)]
    for l1 in range(L1):
        for l2 in range(L2 - 1):
            state.virtual_bond[l1, l2, "R"] = charged(int(per_row * (L2 - l2 - 1) / L2))
    return state


def hubbard_abstract_state(L1, L2, T, t, U):
    """the Hubbard model the reference ships: symmetry FermiU1 (total particle number T), physical edge empty / singly occupied
    (dimension 2) / doubly occupied (tetraku/tetraku/models/hubbard/__init__.py:22-44)"""
    from . import common_tensor
    op = common_tensor.Fermi_Hubbard
    state = AbstractState(TAT.FermiU1.D.Tensor, L1, L2)
    state.total_symmetry = T
    state.physics_edges[...] = [(0, 1), (1, 2), (2, 1)]
    state.hamiltonians["vertical_bond"] = -t * op.CSCS
    state.hamiltonians["horizontal_bond"] = -t * op.CSCS
    state.hamiltonians["single_site"] = U * op.NN
    return state


def hubbard_abstract_lattice(L1, L2, D, T, t, U):
    """its lattice (hubbard/__init__.py:47-75): column-0 vertical bonds carry the particles still to be distributed below (+- 1),
    row bonds the row's share, other vertical bonds are trivial; `D` per charge sector"""
    state = AbstractLattice(hubbard_abstract_state(L1, L2, T, t, U))
    per_row = T / L1
    for l1 in range(L1 - 1):
        Q = int(T * (L1 - l1 - 1) / L1)
        state.virtual_bond[l1, 0, "D"] = [(Q - 1, D), (Q, D), (Q + 1, D)]
        for l2 in range(1, L2):
            state.virtual_bond[l1, l2, "D"] = [(0, D)]
    for l1 in range(L1):
        for l2 in range(L2 - 1):
            Q = int(per_row * (L2 - l2 - 1) / L2)
            state.virtual_bond[l1, l2, "R"] = [(Q - 1, D), (Q, D), (Q + 1, D)]
    return state


def j1j2_reference_abstract_state(L1, L2, J1, J2):
    """the J1-J2 model exactly as the reference ships it (tetraku/tetraku/models/J1J2/__init__.py:22-50): NoSymmetry, and the
    anti-diagonal terms keyed (lower-left site, upper-right site) -- `j1j2_abstract_lattice` above keys them the other way round,
    which is the same operator (SS is symmetric) but another dictionary key and iteration order"""
    from . import common_tensor
    state = AbstractState(TAT.No.D.Tensor, L1, L2)
    state.physics_edges[...] = 2
    J1SS, J2SS = -J1 * common_tensor.No.SS, -J2 * common_tensor.No.SS
    for l1, l2 in state.sites():
        if l1 != L1 - 1:
            state.hamiltonians[(l1, l2, 0), (l1 + 1, l2, 0)] = J1SS
        if l2 != L2 - 1:
            state.hamiltonians[(l1, l2, 0), (l1, l2 + 1, 0)] = J1SS
            if l1 != L1 - 1:
                state.hamiltonians[(l1, l2, 0), (l1 + 1, l2 + 1, 0)] = J2SS
            if l1 != 0:
                state.hamiltonians[(l1, l2, 0), (l1 - 1, l2 + 1, 0)] = J2SS
    return state


def j1j2_reference_abstract_lattice(L1, L2, D, J1, J2):
    state = AbstractLattice(j1j2_reference_abstract_state(L1, L2, J1, J2))
    state.virtual_bond["R"] = D
    state.virtual_bond["D"] = D
    return state

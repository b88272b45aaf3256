"""Operator tensors the shipped models are built from, in float64 (reference: tetragono/tetragono/common_tensor/: `No.py`, `Fermi.py`, `Parity.py`, `Fermi_Hubbard.py`, `Parity_Hubbard.py`,
`FermiU1_Hubbard.py`, `FermiFermi_Hubbard.py`, `FermiU1_tJ.py` and `tensor_toolkit.py`; the reference defines them as complex128 and its models take
`.to(float)`, tetraku/models/*/).  Same attribute names: `common_tensor.No.SS`, `common_tensor.FermiFermi_Hubbard.NN / CSCS / Up.CC ...`,
`common_tensor.FermiU1_tJ.CC / SS / nn / EF`.  Operators that are not real (`pauli_y`, `Sy`) are absent -- only float64 tensors are
device-backed; their real products (`SySy`, `pauli_y_pauli_y`) are there.  Built on first access (they need a backend), set-up work.

A creation operator is the rank-3 tensor ("O0", "I0", "T") with every symmetry-allowed element equal to one, whose third edge "T"
carries the charge it adds; the matching annihilation operator carries the opposite charge on an edge of the opposite arrow, and a
hopping term contracts the two "T" edges -- the fermionic signs then come out of `Tensor.contract`, not out of a table.
"""
from .. import TAT as _TAT


def rename_io(t, m):
    """relabel the sites an operator acts on: I{i}, O{i} -> I{m[i]}, O{m[i]} (tensor_toolkit.py:32-39)"""
    if not isinstance(m, dict):
        m = dict(enumerate(m))
    return t.edge_rename({f"{d}{i}": f"{d}{j}" for i, j in m.items() for d in "IO"})


def kronecker_product(res, *others):
    for t in others:
        res = res.contract(t, set())
    return res


def half_reverse(tensor):
    return tensor.reverse_edge(set(tensor.names), False, {name for name in tensor.names if name.startswith("O")})


class _Namespace:
    def __init__(self, **items):
        self.__dict__.update(items)


def _ladder(Tensor, EF, ET, charge):
    """(creation, annihilation) operator pair for the mode that carries `charge`"""
    Edge = Tensor.Edge
    minus = tuple(-c for c in charge)
    creation = Tensor(["O0", "I0", "T"], [EF, ET, Edge([minus], False)]).range_(1, 0)
    annihilation = Tensor(["O0", "I0", "T"], [EF, ET, Edge([charge], True)]).range_(1, 0)
    return creation, annihilation


def _hop(creation, annihilation, i, j):
    """c^dagger_i c_j"""
    return rename_io(creation, [i]).contract(rename_io(annihilation, [j]), {("T", "T")})


def _number(creation, annihilation):
    return creation.contract(annihilation, {("I0", "O0"), ("T", "T")})


def _build_No():
    Tensor = _TAT.No.D.Tensor

    def matrix(elements):
        t = Tensor(["I0", "O0"], [2, 2]).zero_()
        for (i, o), v in elements.items():
            t[{"I0": i, "O0": o}] = v
        return t

    identity = matrix({(0, 0): 1, (1, 1): 1})
    pauli_x = matrix({(0, 1): 1, (1, 0): 1})
    pauli_z = matrix({(0, 0): 1, (1, 1): -1})
    i_pauli_y = matrix({(0, 1): 1, (1, 0): -1})     # i * sigma_y is real: sigma_y (x) sigma_y = -(i sigma_y) (x) (i sigma_y)
    pauli_x_pauli_x = kronecker_product(rename_io(pauli_x, [0]), rename_io(pauli_x, [1]))
    pauli_y_pauli_y = -1.0 * kronecker_product(rename_io(i_pauli_y, [0]), rename_io(i_pauli_y, [1]))
    pauli_z_pauli_z = kronecker_product(rename_io(pauli_z, [0]), rename_io(pauli_z, [1]))
    SxSx, SySy, SzSz = pauli_x_pauli_x / 4, pauli_y_pauli_y / 4, pauli_z_pauli_z / 4
    return _Namespace(Tensor=Tensor, identity=identity, pauli_x=pauli_x, pauli_z=pauli_z, Sx=pauli_x / 2, Sz=pauli_z / 2,
                      pauli_x_pauli_x=pauli_x_pauli_x, pauli_y_pauli_y=pauli_y_pauli_y, pauli_z_pauli_z=pauli_z_pauli_z,
                      SxSx=SxSx, SySy=SySy, SzSz=SzSz, SS=SxSx + SySy + SzSz)


def _species(Tensor, EF, ET, charge):
    """the operators of one fermion species on sites with physical edge EF (and its conjugate ET)"""
    CP, CM = _ladder(Tensor, EF, ET, charge)
    C0C1, C1C0 = _hop(CP, CM, 0, 1), _hop(CP, CM, 1, 0)
    return dict(CP=CP, CM=CM, C0C1=C0C1, C1C0=C1C0, CC=C0C1 + C1C0, I=Tensor(["O0", "I0"], [EF, ET]).identity_({("I0", "O0")}),
                N=rename_io(CP, [0]).contract(rename_io(CM, [0]), {("T", "T"), ("I0", "O0")}))


def _two_species(Tensor, EF, ET, up, down):
    Up, Down = _Namespace(**_species(Tensor, EF, ET, up)), _Namespace(**_species(Tensor, EF, ET, down))
    return _Namespace(Tensor=Tensor, EF=EF, ET=ET, Up=Up, Down=Down, NN=Up.N.contract(Down.N, {("I0", "O0")}), CSCS=Up.CC + Down.CC)


def _build_Fermi():
    """spinless fermions, symmetry FermiU1 (Fermi.py:21-39)"""
    Tensor = _TAT.FermiU1.D.Tensor
    EF, ET = Tensor.Edge([0, 1], False), Tensor.Edge([0, -1], True)
    return _Namespace(Tensor=Tensor, EF=EF, ET=ET, **_species(Tensor, EF, ET, (1,)))


def _build_Parity():
    """spinless fermions with only the parity conserved, symmetry FermiZ2 (Parity.py:21-44): also the pair creation /
    annihilation terms CP2 = c^dagger_0 c^dagger_1 and CM2 = c_1 c_0"""
    Tensor = _TAT.FermiZ2.D.Tensor
    EF, ET = Tensor.Edge([False, True], False), Tensor.Edge([False, True], True)
    CP = Tensor(["O0", "I0", "T"], [EF, ET, Tensor.Edge([True], False)]).zero_()
    CP[{"O0": (True, 0), "I0": (False, 0), "T": (True, 0)}] = 1
    CM = Tensor(["O0", "I0", "T"], [EF, ET, Tensor.Edge([True], True)]).zero_()
    CM[{"O0": (False, 0), "I0": (True, 0), "T": (True, 0)}] = 1
    C0C1, C1C0 = _hop(CP, CM, 0, 1), _hop(CP, CM, 1, 0)
    return _Namespace(Tensor=Tensor, EF=EF, ET=ET, CP=CP, CM=CM, C0C1=C0C1, C1C0=C1C0, CC=C0C1 + C1C0,
                      I=Tensor(["O0", "I0"], [EF, ET]).identity_({("I0", "O0")}),
                      N=rename_io(CP, [0]).contract(rename_io(CM, [0]), {("T", "T"), ("I0", "O0")}),
                      CP2=rename_io(CP, [0]).contract(rename_io(CP.reverse_edge({"T"}), [1]), {("T", "T")}),
                      CM2=rename_io(CM, [1]).contract(rename_io(CM.reverse_edge({"T"}), [0]), {("T", "T")}))


def _merge_spin(t, sites):
    """modes 0 .. sites-1 are the up modes of the sites, sites .. 2*sites-1 their down modes: merge (up, down) of every site into
    one physical edge; the merge sign goes to the input side only (the reference's `put_sign_in_H`)"""
    groups = {f"{d}{i}": [f"{d}{i}", f"{d}{i + sites}"] for i in range(sites) for d in "IO"}
    return t.merge_edge(groups, True, {f"O{i}" for i in range(sites)})


def _with_spectators(t, acting_on, modes=4):
    """`t` on the modes `acting_on`, the identity on the other ones"""
    f_identity = _with_spectators.identity
    return kronecker_product(rename_io(t, acting_on), *(rename_io(f_identity, [m]) for m in range(modes) if m not in acting_on))


def _spinful(f):
    """the operators of a spinful site built from two spinless modes `f` (Fermi_Hubbard.py:18-87, Parity_Hubbard.py:18-133)"""
    _with_spectators.identity = f.I
    one = lambda t, i: rename_io(t, [i])  # noqa: E731
    N0 = _merge_spin(kronecker_product(one(f.N, 0), one(f.I, 1)), 1)
    N1 = _merge_spin(kronecker_product(one(f.I, 0), one(f.N, 1)), 1)
    return dict(Tensor=f.Tensor, CC=f.CC, I=f.I, N=f.N, C0C1=f.C0C1, C1C0=f.C1C0,
                CSCS=_merge_spin(_with_spectators(f.CC, [0, 1]) + _with_spectators(f.CC, [2, 3]), 2),
                NN=_merge_spin(kronecker_product(one(f.N, 0), one(f.N, 1)), 1), N0=N0, N1=N1,
                CUCD=_merge_spin(f.C0C1, 1), CDCU=_merge_spin(f.C1C0, 1), CUCU=N0, CDCD=N1)


def _build_Fermi_Hubbard():
    """spinful site = (up mode, down mode) of spinless fermions, symmetry FermiU1 (total particle number)"""
    return _Namespace(**_spinful(__getattr__("Fermi")))


def _build_Parity_Hubbard():
    """the same with only the parity conserved (FermiZ2), plus the singlet / triplet pair creation + annihilation terms between
    two sites (Parity_Hubbard.py:38-70)"""
    f = __getattr__("Parity")
    items = _spinful(f)
    a, b = _with_spectators(f.CP2, [0, 3]) + _with_spectators(f.CM2, [3, 0]), _with_spectators(f.CP2, [1, 2]) + _with_spectators(f.CM2, [2, 1])
    return _Namespace(CP2=f.CP2, CM2=f.CM2, singlet=_merge_spin(a + b, 2), triplet=_merge_spin(a - b, 2), **items)


def _build_FermiU1_Hubbard():
    """one site = empty, up, down, double; symmetry (particle number: FermiU1, 2 Sz: BoseU1) (FermiU1_Hubbard.py:21-50)"""
    Tensor = _TAT.FermiU1BoseU1.D.Tensor
    EF = Tensor.Edge([(0, 0), (1, 1), (1, -1), (2, 0)], False)
    ET = Tensor.Edge([(0, 0), (-1, -1), (-1, 1), (-2, 0)], True)
    return _two_species(Tensor, EF, ET, (1, 1), (1, -1))


def _build_FermiFermi_Hubbard():
    """one site = (n_up, n_down) in {0,1}^2, symmetry FermiU1 x FermiU1 (FermiFermi_Hubbard.py:22-60)"""
    Tensor = _TAT.FermiU1FermiU1.D.Tensor
    EF = Tensor.Edge([(0, 0), (0, 1), (1, 0), (1, 1)], False)
    ET = Tensor.Edge([(0, 0), (0, -1), (-1, 0), (-1, -1)], True)
    return _two_species(Tensor, EF, ET, (1, 0), (0, 1))


def _build_FermiU1_tJ():
    """one site = empty, down, up; symmetry (particle number: FermiU1, 2 Sz: BoseU1) (FermiU1_tJ.py:22-74)"""
    Tensor = _TAT.FermiU1BoseU1.D.Tensor
    EF = Tensor.Edge([(0, 0), (1, -1), (1, 1)], False)
    ET = Tensor.Edge([(0, 0), (-1, 1), (-1, -1)], True)
    CPU, CMU = _ladder(Tensor, EF, ET, (1, 1))
    CPD, CMD = _ladder(Tensor, EF, ET, (1, -1))
    C0UC1U, C1UC0U = _hop(CPU, CMU, 0, 1), _hop(CPU, CMU, 1, 0)
    C0DC1D, C1DC0D = _hop(CPD, CMD, 0, 1), _hop(CPD, CMD, 1, 0)
    CC = C0UC1U + C0DC1D + C1UC0U + C1DC0D
    Sz2 = _number(CPU, CMU) - _number(CPD, CMD)
    SzSz4 = rename_io(Sz2, [0]).contract(rename_io(Sz2, [1]), set())
    # 2 (SxSx + SySy) = S+_0 S-_1 + S-_0 S+_1, written with the hopping terms (two exchanges of fermion operators: one sign)
    exchange = {("I0", "O0"), ("O1", "I1")}
    SxSxSySy2 = -1 * (C0DC1D.contract(C1UC0U, exchange) + C0UC1U.contract(C1DC0D, exchange))
    n = _number(CPU, CMU) + _number(CPD, CMD)
    return _Namespace(Tensor=Tensor, EF=EF, ET=ET, CPU=CPU, CPD=CPD, CMU=CMU, CMD=CMD, C0UC1U=C0UC1U, C0DC1D=C0DC1D, C1UC0U=C1UC0U,
                      C1DC0D=C1DC0D, CC=CC, Sz2=Sz2, SzSz4=SzSz4, SxSxSySy2=SxSxSySy2, SS=SzSz4 / 4 + SxSxSySy2 / 2, n=n,
                      nn=rename_io(n, [0]).contract(rename_io(n, [1]), set()))


_BUILDERS = {"No": _build_No, "Fermi": _build_Fermi, "Parity": _build_Parity, "Fermi_Hubbard": _build_Fermi_Hubbard, "Parity_Hubbard": _build_Parity_Hubbard, "FermiU1_Hubbard": _build_FermiU1_Hubbard, "FermiFermi_Hubbard": _build_FermiFermi_Hubbard,
             "FermiU1_tJ": _build_FermiU1_tJ}
_BUILT = {}


def __getattr__(name):
    if name in _BUILDERS:
        if name not in _BUILT:
            _BUILT[name] = _BUILDERS[name]()
        return _BUILT[name]
    raise AttributeError(name)

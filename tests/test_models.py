"""Operator tensors and fermionic model builders (`tetragono/common_tensor.py`, `models.tJ_* / hubbard_fermi_fermi_*`) against what the
UNMODIFIED reference built from its own `common_tensor` + tetraku models: the Hamiltonian terms, physical edges, total symmetry and
site-tensor structure stored in the t-J and Hubbard fixtures (tests/golden/tJ_4x4_D1_Dc8.npz: t = 1, J = 0.4;
hubbardFF_4x4_D1_Dc8.npz: t = 1, U = 4), and the spin-1/2 SS of the Heisenberg fixture."""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import build_lattice, load
from tnsp_b200.tetragono import common_tensor, models
from tnsp_b200.tetragono.state import SamplingLattice


def _same_model(mine, fixture):
    assert mine.total_symmetry == fixture.total_symmetry
    assert set(mine._hamiltonians) == set(fixture._hamiltonians)
    for positions, want in fixture._hamiltonians.items():
        got = mine._hamiltonians[positions]
        assert set(got.names) == set(want.names)
        got = got.transpose(want.names)
        assert got._edges == want._edges
        a, b = np.asarray(got.storage), np.asarray(want.storage)
        assert np.abs(a - b).max() <= 1e-14 * max(1.0, np.abs(b).max())
    for l1, l2 in fixture.sites():
        assert dict(mine.physics_edges[l1, l2].items()) == dict(fixture.physics_edges[l1, l2].items())
        assert mine[l1, l2].names == fixture[l1, l2].names and mine[l1, l2]._edges == fixture[l1, l2]._edges


@pytest.mark.parametrize("case", ["tJ_4x4_D1_Dc8", "hubbardFF_4x4_D1_Dc8", "model_hubbard_4x4_D1"])
def test_fermionic_models_equal_the_reference_models(case):
    meta, z = load(case)
    fixture = build_lattice(meta, z)
    if case.startswith("tJ"):
        abstract = models.tJ_abstract_lattice(4, 4, 1, 2, 1.0, 0.4)
    elif case.startswith("model_hubbard"):
        abstract = models.hubbard_abstract_lattice(4, 4, 1, 8, 1.0, 4.0)       # the Hubbard lattice the reference ships (FermiU1)
    else:
        abstract = models.hubbard_fermi_fermi_abstract_lattice(4, 4, 1, 8, 1.0, 4.0)
    TAT.random.seed(2333)
    _same_model(SamplingLattice(abstract), fixture)


def test_shipped_j1j2_model_equals_the_reference():
    """tests/golden/model_j1j2_3x4_D2.npz (`make_golden.py j1j2model`): same Hamiltonian keys IN THE SAME ORDER, tensors, site structure"""
    meta, z = load("model_j1j2_3x4_D2")
    fixture = build_lattice(meta, z)
    TAT.random.seed(2333)
    mine = SamplingLattice(models.j1j2_reference_abstract_lattice(3, 4, 2, 1.0, 0.5))
    _same_model(mine, fixture)
    assert list(mine._hamiltonians) == [tuple(tuple(p) for p in h["positions"]) for h in meta["hamiltonians"]]
    # randn_ from the same seed fills the same values: the whole state is the reference's
    for l1, l2 in fixture.sites():
        assert np.array_equal(np.asarray(mine[l1, l2].storage), np.asarray(fixture[l1, l2].storage))


def test_spin_half_operators():
    op = common_tensor.No
    meta, z = load("heis_3x3_D2_Dc4")
    fixture = build_lattice(meta, z)
    want = fixture._hamiltonians[((0, 0, 0), (0, 1, 0))]          # the reference's -J * common_tensor.No.SS.to(float), J = 1
    got = (-1.0 * op.SS).transpose(want.names)
    assert np.abs(np.asarray(got.storage) - np.asarray(want.storage)).max() <= 1e-15
    sz = np.asarray(op.Sz.storage).reshape(2, 2)
    assert np.array_equal(sz, np.diag([0.5, -0.5]))
    total = np.asarray((op.SxSx + op.SySy + op.SzSz).transpose(["O0", "O1", "I0", "I1"]).storage).reshape(4, 4)
    assert np.allclose(np.linalg.eigvalsh(total), [-0.75, 0.25, 0.25, 0.25])


def test_fermionic_operator_algebra():
    """n = c^dagger c is a projector; the hopping term is Hermitian; N_up N_down counts double occupancy"""
    ff = common_tensor.FermiFermi_Hubbard
    for species in (ff.Up, ff.Down):
        n = species.N
        nn = n.contract(n.edge_rename({"I0": "X"}), {("I0", "O0")}).edge_rename({"X": "I0"})
        assert float((nn.transpose(n.names) - n).norm_max()) <= 1e-15
    swap = {"I0": "O0", "O0": "I0", "I1": "O1", "O1": "I1"}
    assert float((ff.CSCS - ff.CSCS.conjugate().edge_rename(swap)).norm_max()) <= 1e-15
    assert float(ff.NN.norm_sum()) == 1.0
    tj = common_tensor.FermiU1_tJ
    assert float((tj.CC - tj.CC.conjugate().edge_rename(swap)).norm_max()) <= 1e-15
    assert float((tj.SS - tj.SS.conjugate().edge_rename(swap)).norm_max()) <= 1e-15
    with pytest.raises(AttributeError):
        common_tensor.Nothing


def test_every_real_common_tensor_equals_the_reference():
    """tests/golden/common_tensor.npz (`make_golden.py common`): every real operator tensor of the reference's common_tensor modules
    No, Fermi, Parity, Fermi_Hubbard, Parity_Hubbard, FermiU1_Hubbard, FermiFermi_Hubbard, FermiU1_tJ (incl. Up.* / Down.*), as `.to(float)`: same names in the same order,
    same edges, same values"""
    import json
    import os
    from golden_loader import HERE, tensor_from
    z = np.load(os.path.join(HERE, "common_tensor.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    checked = 0
    for module, info in meta.items():
        mine = getattr(common_tensor, module)
        mod = getattr(TAT, info["symmetry"])
        for path, desc in info["tensors"].items():
            got = mine
            for part in path.split("."):
                got = getattr(got, part)
            want = tensor_from(mod, desc, z)
            assert got.names == want.names, (module, path)
            assert got._edges == want._edges, (module, path)
            a, b = np.asarray(got.storage), np.asarray(want.storage)
            assert a.shape == b.shape and np.abs(a - b).max() <= 1e-15, (module, path)
            checked += 1
    assert checked == 105


def test_shipped_hubbard_model_exact_energy_and_gradient():
    """2x2 Hubbard lattice of the reference's shipped model (FermiU1; the physical edge has a segment of dimension 2), 2 particles,
    measured exactly by ergodic enumeration: energy (value, deviation) and gradient against the unmodified reference
    (tests/golden/model_hubbard_2x2_D2.npz, `make_golden.py hubbard`), <= 1e-10"""
    from golden_loader import tensor_from
    from tnsp_b200.tetragono.observer import Observer
    from tnsp_b200.tetragono.sampling import ErgodicSampling
    meta, z = load("model_hubbard_2x2_D2")
    fixture = build_lattice(meta, z)
    TAT.random.seed(2333)
    mine = SamplingLattice(models.hubbard_abstract_lattice(2, 2, 2, 2, 1.0, 4.0))
    _same_model(mine, fixture)
    sampling = ErgodicSampling(fixture, 16, None)
    assert sampling.total_step == int(z["count"][0])
    obs = Observer(fixture, enable_energy=True, enable_gradient=True)
    with obs:
        for _ in range(sampling.total_step):
            p, c = sampling()
            obs(p, c)
    want = z["energy"]
    assert np.abs(np.array(obs.total_energy) - want).max() <= 1e-10 * np.abs(want).max()
    grad = obs.gradient
    for l1, l2 in fixture.sites():
        w = tensor_from(TAT.FermiU1, meta["gradient"][l1][l2], z)
        g = grad[l1][l2].transpose(w.names)
        assert g._edges == w._edges
        scale = max(1.0, np.abs(np.asarray(w.storage)).max())
        assert np.abs(np.asarray(g.storage) - np.asarray(w.storage)).max() <= 1e-10 * scale

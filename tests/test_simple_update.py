"""Simple update against the UNMODIFIED reference (tests/golden/simple_update.npz, written by
`make_golden.py simple`): from the fixture PEPS, `SimpleUpdateLattice.update` must give the same bond dimensions (integers: exact),
the same bond environments (normalised singular values) and the same site tensors up to the sign gauge of the svd (|elements|),
and the state converted back for sampling must have the same exact amplitude on the fixture configuration (gauge invariant).
Cases: no symmetry (absolute and relative cut), a truncating one, U(1), U(1) with diagonal J2 terms (the long-range update that
carries the operator leg across a bond), fermionic t-J and Hubbard."""
import os

import numpy as np
import pytest

from golden_loader import HERE, build_lattice, config_points, load
from tnsp_b200.tetragono.configuration import Configuration
from tnsp_b200.tetragono.simple_update import (SimpleUpdateLattice, sampling_lattice_to_simple_update_lattice,
                                               simple_update_lattice_to_sampling_lattice)

CASES = ["heis_3x3_D2_Dc4", "heis_3x3_D2_Dc4:relative", "heis_4x4_D3_Dc5_truncating", "heisU1_4x4_d1_Dc6", "j1j2U1_4x4_d1_Dc9",
         "tJ_4x4_D1_Dc8", "hubbardFF_4x4_D1_Dc8"]
TOL = 1e-8  # relative; the reference factorises with LAPACK, this build with Householder / one-sided Jacobi kernels


def _close(got, want, what):
    got, want = np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)
    assert got.shape == want.shape, what
    assert np.abs(got - want).max() <= TOL * max(1e-300, np.abs(want).max()), what


@pytest.mark.parametrize("case", CASES)
def test_simple_update_matches_the_reference(case):
    gold = np.load(os.path.join(HERE, "simple_update.npz"))
    fixture = case.split(":")[0]
    meta, z = load(fixture)
    lat = build_lattice(meta, z)
    steps, tau, dim = gold[case + "_par"]
    su = sampling_lattice_to_simple_update_lattice(lat)
    su.update(int(steps), float(tau), int(dim) if dim >= 1 else float(dim))
    for l1, l2 in su.sites():
        t = su[l1, l2]
        assert [t.edge_by_name(n).dimension for n in t.names] == list(gold[f"{case}_dims_{l1}_{l2}"])
        assert su.virtual_bond[l1, l2] == {n: t.edge_by_name(n) for n in t.names if not n.startswith("P")}
        for d in "RD":
            env = su.environment[l1, l2, d]
            key = f"{case}_env_{l1}_{l2}_{d}"
            assert (env is None) == (key not in gold.files)
            if env is not None:
                _close(env.storage, gold[key], key)
        _close(np.abs(np.asarray(t.storage)), gold[f"{case}_site_{l1}_{l2}"], f"site {l1} {l2}")
    back = simple_update_lattice_to_sampling_lattice(su)
    conf = Configuration(back, 256)
    if case.startswith(("tJ", "hubbard")):
        conf.import_configuration(np.load(os.path.join(HERE, "gauge_fixing.npz"))[fixture + "_conf"])
    else:
        pts = config_points(meta)
        for l1, l2 in back.sites():
            for o, p in pts[l1][l2].items():
                conf[l1, l2, o] = p
    _close([float(conf.hole(()))], gold[case + "_ws"], "amplitude of the converted state")


def test_environment_handler_and_errors():
    meta, z = load("heis_3x3_D2_Dc4")
    su = SimpleUpdateLattice(build_lattice(meta, z))
    assert su.environment[0, 0, "L"] is None and su.environment[0, 0, "R"] is None and su.environment[2, 2, "D"] is None
    with pytest.raises(ValueError):
        su.environment[0, 0, "X"]
    with pytest.raises(ValueError):
        su.environment[0, 0, "U"] = None
    marker = object()
    su.environment[1, 1, "L"] = marker
    assert su.environment[1, 0, "R"] is marker
    su.environment[1, 1, "U"] = marker
    assert su.environment[0, 1, "D"] is marker
    with pytest.raises(ValueError):
        simple_update_lattice_to_sampling_lattice(build_lattice(meta, z))
    with pytest.raises(ValueError):
        sampling_lattice_to_simple_update_lattice(su)
    with pytest.raises(NotImplementedError):
        su.observe_energy()


def test_imaginary_time_evolution_lowers_the_energy():
    """physics check without the reference: simple update of the 3x3 Heisenberg fixture lowers the exact energy expectation
    (all 2^9 configurations through the ergodic sampler and the sampling Observer)"""
    from tnsp_b200.tetragono.observer import Observer
    from tnsp_b200.tetragono.sampling import ErgodicSampling
    meta, z = load("heis_3x3_D2_Dc4")
    lat = build_lattice(meta, z)

    def energy(state):
        obs = Observer(state, enable_energy=True)
        sampling = ErgodicSampling(state, 64, nb=64)
        with obs:
            for _ in range(sampling.calls):
                p, c = sampling()
                obs(p, c)
        return obs.energy[0]

    before = energy(lat)
    su = sampling_lattice_to_simple_update_lattice(lat)
    su.update(20, 0.05, 2)
    after = energy(simple_update_lattice_to_sampling_lattice(su))
    assert after < before - 0.05


def test_reference_simple_update_checkpoint_loads():
    """a `pickle.dump(SimpleUpdateLattice)` of the unmodified reference (state_su_heisU1_4x4_d1.pkl, the heisU1 case above after
    its update) loads bit-exactly -- tensors and bond environments -- and converts to the sampling state with the same amplitude"""
    from tnsp_b200.tetragono.checkpoint import load_reference_state
    gold = np.load(os.path.join(HERE, "simple_update.npz"))
    case = "heisU1_4x4_d1_Dc6"
    su = load_reference_state(os.path.join(HERE, "state_su_heisU1_4x4_d1.pkl"))
    assert isinstance(su, SimpleUpdateLattice)
    for l1, l2 in su.sites():
        assert np.array_equal(np.abs(np.asarray(su[l1, l2].storage)), gold[f"{case}_site_{l1}_{l2}"])
        for d in "RD":
            env = su.environment[l1, l2, d]
            key = f"{case}_env_{l1}_{l2}_{d}"
            assert (env is None) == (key not in gold.files)
            if env is not None:
                assert np.array_equal(np.asarray(env.storage), gold[key])
    meta, _ = load(case)
    back = simple_update_lattice_to_sampling_lattice(su)
    conf = Configuration(back, 256)
    pts = config_points(meta)
    for l1, l2 in back.sites():
        for o, p in pts[l1][l2].items():
            conf[l1, l2, o] = p
    assert abs(float(conf.hole(())) - gold[case + "_ws"][0]) <= 1e-10 * abs(gold[case + "_ws"][0])

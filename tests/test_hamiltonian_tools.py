"""Hamiltonian bookkeeping of AbstractState (`hamiltonians.trace_repeated() / sort_points() / check_hermite()`,
abstract_state.py:200-240 with utility.py:421-480) and `SamplingLattice.lattice_dot` (lattice.py:921-934) against the UNMODIFIED
reference on the fermionic t-J model, where the traces and renames carry signs (tests/golden/hamiltonian_tools.npz, written by
`make_golden.py hamiltonian`).  Pure data movement: bit-exact; the trace and the dot product <= 1e-12."""
import json
import os

import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import HERE, build_lattice, load, tensor_from
from tnsp_b200.tetragono.state import sort_points, trace_repeated


def _gold():
    z = np.load(os.path.join(HERE, "hamiltonian_tools.npz"))
    return json.loads(bytes(z["meta"]).decode()), z


def test_trace_repeated_and_sort_points_match_the_reference():
    meta, z = _gold()
    mod = TAT.FermiU1BoseU1
    tensors = {name: tensor_from(mod, meta[name], z) for name in ("three", "four")}
    for case in meta["cases"]:
        fn = trace_repeated if case["kind"] == "trace" else sort_points
        points = tuple(tuple(p) for p in case["points"])
        got, new_points = fn(tensors[case["tensor"]], points)
        assert [list(p) for p in new_points] == case["new_points"]
        want = tensor_from(mod, case["result"], z)
        assert set(got.names) == set(want.names)
        got = got.transpose(want.names)
        assert got._edges == want._edges
        a, b = np.asarray(got.storage).reshape(-1), np.asarray(want.storage).reshape(-1)
        if case["kind"] == "sort":
            assert got.names == want.names and np.array_equal(a, b)
        else:
            assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
        again, _ = fn(tensors[case["tensor"]], points)
        assert again is fn(tensors[case["tensor"]], points)[0]    # kept per tensor identity, like the reference's pools


def test_hamiltonians_regroup_and_check_hermite():
    meta, z = _gold()
    lat_meta, lat_z = load("tJ_4x4_D1_Dc8")
    lat = build_lattice(lat_meta, lat_z)
    mod = TAT.FermiU1BoseU1
    lat.hamiltonians.check_hermite(1e-12)
    before = len(lat.hamiltonians)
    # a two-site term given with its points in descending order joins the term on the same bond after sort_points
    pair = ((0, 0, 0), (0, 1, 0))
    term = lat.hamiltonians[pair]
    flipped = term.edge_rename({"I0": "I1", "I1": "I0", "O0": "O1", "O1": "O0"})
    lat.hamiltonians[(0, 1, 0), (0, 0, 0)] = flipped
    assert len(lat.hamiltonians) == before + 1
    lat.hamiltonians.sort_points()
    assert len(lat.hamiltonians) == before
    doubled = lat.hamiltonians[pair].transpose(term.names)
    assert np.abs(np.asarray(doubled.storage) - 2 * np.asarray(term.storage)).max() <= 1e-14
    # check_hermite: the symmetrised random tensor of the fixture passes, the raw one does not
    hermitian, raw = tensor_from(mod, meta["hermitian"], z), tensor_from(mod, meta["three"], z)
    points = ((0, 0, 0), (0, 1, 0), (0, 2, 0))
    lat._hamiltonians = {points: hermitian}
    lat.hamiltonians.check_hermite(1e-12)
    lat._hamiltonians = {points: raw}
    with pytest.raises(ValueError):
        lat.hamiltonians.check_hermite(1e-12)


def test_lattice_dot_matches_the_reference():
    _, z = _gold()
    lat_meta, lat_z = load("tJ_4x4_D1_Dc8")
    lat = build_lattice(lat_meta, lat_z)
    assert abs(lat.lattice_dot() - z["lattice_dot"][0]) <= 1e-12 * z["lattice_dot"][0]
    lattice = [[lat[l1, l2] for l2 in range(lat.L2)] for l1 in range(lat.L1)]
    doubled = [[lat[l1, l2] * 2.0 for l2 in range(lat.L2)] for l1 in range(lat.L1)]
    assert abs(lat.lattice_dot(lattice, doubled) - 2 * z["lattice_dot"][0]) <= 1e-12 * z["lattice_dot"][0]


def test_observer_setters():
    from tnsp_b200.tetragono.observer import Observer
    lat_meta, lat_z = load("heis_3x3_D2_Dc4")
    lat = build_lattice(lat_meta, lat_z)
    obs = Observer(lat, enable_energy=True)
    obs.set_classical_energy(lambda configuration: 1.0)
    obs.restrict_subspace(None)
    obs.cache_configuration("drop")
    with pytest.raises(ValueError):
        obs.cache_configuration("never")
    with obs:
        with pytest.raises(RuntimeError):
            obs.cache_configuration(True)
        with pytest.raises(RuntimeError):
            obs.restrict_subspace(None)

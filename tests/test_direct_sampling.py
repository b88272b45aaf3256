"""Direct sampling against the UNMODIFIED reference (tests/golden/direct_sampling.npz, written by
`make_golden.py direct`): from the same seed the same configurations are drawn (integers: exact) with the same probabilities
(<= 1e-9; finite cuts on both the sampled boundary and the double-layer boundary), for a lattice without symmetry, a
truncating one, a U(1) one and two fermionic ones (t-J, Hubbard).  Plus a physics check: the drawn frequencies follow |psi|^2."""
import os

import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import HERE, build_lattice, load
from tnsp_b200.tetragono.direct_sampling import DirectSampling, double_layer_rows_from_below


@pytest.mark.parametrize("case", ["heis_3x3_D2_Dc4", "heis_4x4_D3_Dc5_truncating", "heisU1_4x4_d1_Dc6", "tJ_4x4_D1_Dc8", "hubbardFF_4x4_D1_Dc8"])
def test_direct_sampling_matches_the_reference(case):
    gold = np.load(os.path.join(HERE, "direct_sampling.npz"))
    Dc, dl_cut, seed = (int(x) for x in gold[case + "_par"])
    meta, z = load(case)
    lat = build_lattice(meta, z)
    TAT.random.seed(seed)
    sampling = DirectSampling(lat, Dc, None, dl_cut)
    for want_conf, want_p in zip(gold[case + "_conf"], gold[case + "_poss"]):
        p, c = sampling()
        assert np.array_equal(c.export_configuration(), want_conf)
        assert abs(p - want_p) <= 1e-9 * want_p


def test_double_layer_boundary_is_the_norm():
    """rows from below, untruncated: closing the top row gives <psi|psi> = sum over configurations of the squared amplitude"""
    from tnsp_b200.tetragono.sampling import ErgodicSampling
    meta, z = load("heis_3x3_D2_Dc4")
    lat = build_lattice(meta, z)
    rows = double_layer_rows_from_below(lat, 64, normalize=False)
    t = rows[0][0]
    for l2 in range(1, lat.L2):
        t = t.contract(rows[0][l2], {("R", "L")})
    norm = float(t)
    s = ErgodicSampling(lat, 64, nb=64)
    total = 0.0
    for _ in range(s.calls):
        p, c = s()
        ws = np.asarray(c.hole(()).storage).reshape(-1)
        total += float((ws**2 * np.isfinite(p)).sum())
    assert abs(norm - total) <= 1e-10 * total


def test_direct_sampling_probability_is_the_born_probability():
    """untruncated: the probability returned with a configuration is |psi(s)|^2 / <psi|psi>"""
    meta, z = load("heis_3x3_D2_Dc4")
    lat = build_lattice(meta, z)
    rows = double_layer_rows_from_below(lat, 64, normalize=False)
    t = rows[0][0]
    for l2 in range(1, lat.L2):
        t = t.contract(rows[0][l2], {("R", "L")})
    norm = float(t)
    TAT.random.seed(8)
    sampling = DirectSampling(lat, 64, None, 64)
    for _ in range(4):
        p, c = sampling()
        ws = float(c.hole(()))
        assert abs(p - ws**2 / norm) <= 1e-9 * p


def test_lockstep_direct_sampling_equals_independent_chains():
    """nb chains per call (no symmetry): chain c draws what a single-chain sampler with the same engine seed draws"""
    from tnsp_b200.tetragono.sampling import ChainRng
    meta, z = load("heis_3x3_D2_Dc4")
    lat = build_lattice(meta, z)
    seeds = [5, 6, 7, 8]
    rng = ChainRng(len(seeds))
    rng.seed(seeds)
    batch = DirectSampling(lat, 4, None, 4, nb=len(seeds), rng=rng)
    got = [batch() for _ in range(3)]
    for c, seed in enumerate(seeds):
        TAT.random.seed(seed)
        single = DirectSampling(lat, 4, None, 4)
        for step in range(3):
            p, conf = single()
            assert np.array_equal(got[step][1].export_configuration()[c], conf.export_configuration())
            assert abs(got[step][0][c] - p) <= 1e-9 * p

"""Observables beyond the 2x2 window of `Configuration.replace` (`ConfigurationPool.wss`, lattice.py:562-614) against
the UNMODIFIED reference (tests/golden/long_range.npz, written by `make_golden.py longrange`): the first two-site Hamiltonian tensor
placed on distant site pairs, measured by `Observer(cache_configuration=True | "drop")` over sweep samples from a fixed seed.  Same
trajectory (exact) and the same expectation values <= 1e-9 (boundary cut 64: every contraction route is exact)."""
import os

import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import HERE, build_lattice, config_points, load
from tnsp_b200.tetragono.observer import Observer
from tnsp_b200.tetragono.sampling import SweepSampling


@pytest.mark.parametrize("case", ["heis_3x3_D2_Dc4", "heisU1_4x4_d1_Dc6", "tJ_4x4_D1_Dc8"])
def test_long_range_observables_match_the_reference(case):
    gold = np.load(os.path.join(HERE, "long_range.npz"))
    meta, z = load(case)
    lat = build_lattice(meta, z)
    term = [h for p, h in lat.hamiltonians if len(p) == 2][0]
    pairs = [tuple(tuple(int(x) for x in site) for site in pair) for pair in gold[case + "_pairs"]]
    TAT.random.seed(int(gold[case + "_seed"][0]))
    sampling = SweepSampling(lat, 64, None, None)
    if case.startswith("tJ"):
        sampling.configuration.import_configuration(np.load(os.path.join(HERE, "gauge_fixing.npz"))[case + "_conf"])
    else:
        pts = config_points(meta)
        for l1, l2 in lat.sites():
            for o, p in pts[l1][l2].items():
                sampling.configuration[l1, l2, o] = p
    for mode, tag in ((True, case), ("drop", case + "_drop")):
        obs = Observer(lat, cache_configuration=mode)
        obs.add_observer("far", {pair: term for pair in pairs})
        with obs:
            for want in gold[tag + "_conf"]:
                p, c = sampling()
                assert np.array_equal(c.export_configuration(), want)
                obs(p, c)
        got = np.array([obs._result_reweight["far"][pair] / obs._total_weight for pair in pairs])
        assert np.abs(got - gold[tag + "_far"]).max() <= 1e-9 * max(1.0, np.abs(gold[tag + "_far"]).max())


def test_long_range_needs_the_configuration_cache():
    """without the cache the reference raises NotImplementedError for such a term (observer.py:371-374); so does this build"""
    meta, z = load("heis_3x3_D2_Dc4")
    lat = build_lattice(meta, z)
    term = [h for p, h in lat.hamiltonians if len(p) == 2][0]
    sampling = SweepSampling(lat, 64, None, None)
    pts = config_points(meta)
    for l1, l2 in lat.sites():
        for o, p in pts[l1][l2].items():
            sampling.configuration[l1, l2, o] = p
    obs = Observer(lat)
    obs.add_observer("far", {((0, 0, 0), (2, 1, 0)): term})
    with obs, pytest.raises(NotImplementedError):
        p, c = sampling()
        obs(p, c)
    with pytest.raises(ValueError):
        Observer(lat, cache_configuration="sometimes")


def test_split_replacement_rule():
    """the last cluster of changed sites that fits a 2x2 window stays a replacement, the rest goes to the half configuration
    (lattice.py:648-684)"""
    from tnsp_b200.tetragono.configuration import ConfigurationPool
    meta, z = load("heis_4x4_D4_Dc16")
    pool = ConfigurationPool(build_lattice(meta, z))
    first, second = pool._split_replacement({(0, 0, 0): 1, (3, 3, 0): 0})
    assert first == {(3, 3, 0): 0} and second == {(0, 0, 0): 1}
    first, second = pool._split_replacement({(0, 0, 0): 1, (1, 1, 0): 0, (1, 2, 0): 1})
    assert first == {(1, 2, 0): 1} and second == {(0, 0, 0): 1, (1, 1, 0): 0}
    first, second = pool._split_replacement({(2, 1, 0): 1, (2, 2, 0): 0})
    assert first == {} and second == {(2, 1, 0): 1, (2, 2, 0): 0}

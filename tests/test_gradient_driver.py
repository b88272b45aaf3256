"""The optimisation driver (`gradient_descent`) against the UNMODIFIED reference's own loop
(sampling_lattice/gradient.py:93-445; fixtures `tests/golden/driver_*.npz` written by make_golden.py driver): same seed ->
same Markov chains (configurations are integers: exact), energy of every step and the PEPS tensors after the last
update within 1e-8 (three optimisation steps amplify the 1e-10 per-sample agreement)."""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import DRIVER_CASES, build_lattice, load, tensor_from
from tnsp_b200.tetragono.gradient import gradient_descent, lattice_dot


@pytest.mark.parametrize("case", DRIVER_CASES)
def test_driver_matches_reference_loop(case):
    meta, z = load(case)
    lat = build_lattice(meta, z)
    conf = np.array(z["start_configuration"])
    TAT.random.seed(meta["seed"])
    energies = []
    for whole, _ in gradient_descent(lat, sampling_method=meta.get("method", "sweep"), configuration_cut_dimension=meta["Dc"],
                                     sampling_configurations=conf, **meta["kwargs"]):
        energies.append(whole["energy"])
    want = z["step_energy"]
    assert np.abs(np.array(energies) - want).max() <= 1e-8 * np.abs(want).max()
    assert np.array_equal(conf, z["last_configuration"])
    mod = getattr(TAT, meta["symmetry"])
    for l1 in range(meta["L1"]):
        for l2 in range(meta["L2"]):
            t = tensor_from(mod, meta["final_sites"][l1][l2], z)
            got, ref = np.asarray(lat[l1, l2].storage), np.asarray(t.storage)
            assert np.abs(got - ref).max() <= 1e-8 * np.abs(ref).max()


def test_driver_rejects_out_of_scope_options():
    meta, z = load(DRIVER_CASES[0])
    lat = build_lattice(meta, z)
    for kw in ({"use_check_difference": True},):
        with pytest.raises(NotImplementedError):
            next(gradient_descent(lat, 1, 1, **{"sampling_configurations": np.array(z["start_configuration"]), **kw}))
    with pytest.raises(ValueError):
        next(gradient_descent(lat, 1, 1, sampling_method="nonsense"))


def test_lockstep_driver_lowers_the_energy():
    """8 chains in lock step, SR natural gradient, relative step: the energy estimated from 8 x 6 samples per step goes down"""
    meta, z = load("driver_heis_3x3_D2_Dc4_plain")
    lat = build_lattice(meta, z)
    conf = np.array(z["start_configuration"])
    TAT.random.seed(3)
    energies = [whole["energy"][0] for whole, _ in gradient_descent(
        lat, 48, 6, 0.1, chains=8, sampling_method="sweep", configuration_cut_dimension=4, sampling_configurations=conf,
        use_natural_gradient=True, conjugate_gradient_method_step=4, use_fix_relative_step_size=True)]
    assert len(energies) == 6 and np.all(np.isfinite(energies))
    assert np.mean(energies[-2:]) < np.mean(energies[:2])
    assert lattice_dot([[lat[0, 0]]], [[lat[0, 0]]]) > 0


def test_ergodic_driver_energy_is_exact_expectation():
    """ergodic enumeration through the driver = sum_s |psi(s)|^2 E_s / sum_s |psi(s)|^2 computed by brute force"""
    meta, z = load("driver_heis_3x3_D2_Dc4_plain")
    lat = build_lattice(meta, z)
    (whole, _), = list(gradient_descent(lat, sampling_method="ergodic", configuration_cut_dimension=4))
    from tnsp_b200.tetragono.observer import Observer
    from tnsp_b200.tetragono.sampling import ErgodicSampling
    s = ErgodicSampling(lat, 4)
    num = den = 0.0
    for _ in range(s.total_step):
        p, c = s()
        obs = Observer(lat, enable_energy=True)
        with obs:
            obs(p, c)
        w = float(c.hole(()))**2
        num += w * obs.total_energy[0]
        den += w
    assert abs(whole["energy"][0] - num / den) <= 1e-9 * abs(num / den)


def test_driver_writes_reference_checkpoints(tmp_path):
    from tnsp_b200.tetragono.checkpoint import load_reference_state, read_configurations
    meta, z = load("driver_heis_3x3_D2_Dc4_plain")
    lat = build_lattice(meta, z)
    conf = np.array(z["start_configuration"])
    TAT.random.seed(9)
    steps = list(gradient_descent(lat, 4, 2, 0.01, sampling_method="sweep", configuration_cut_dimension=4, sampling_configurations=conf,
                                  save_state_file=str(tmp_path / "state_%s.dat"), save_configuration_file=str(tmp_path / "conf_%s.dat")))
    assert len(steps) == 2
    back = load_reference_state(str(tmp_path / "state_1.dat"))
    for l1 in range(lat.L1):
        for l2 in range(lat.L2):
            assert np.array_equal(np.asarray(back[l1, l2].storage), np.asarray(lat[l1, l2].storage))
    assert np.array_equal(read_configurations(str(tmp_path / "conf_1.dat")), conf)
    raw = np.fromfile(str(tmp_path / "conf_1.dat"), dtype=np.int64)
    assert list(raw[:2]) == [1, conf.ndim] and list(raw[2:2 + conf.ndim]) == list(conf.shape)      # the reference's header


def test_pseudo_inverse_natural_gradient_matches_the_reference_formula():
    """observer.py:697-899 restated with numpy (the reference needs ScaLAPACK, absent here): same Gram matrix, same regularised
    inverse of its eigenvalues; and, unregularised on a full-rank sample set, the result solves (Delta - <Delta>) NG = conj(E - <E>)"""
    from tnsp_b200.tetragono.observer import Observer
    from tnsp_b200.tetragono.sampling import ChainRng, SweepSampling
    meta, z = load("driver_heis_3x3_D2_Dc4_plain")
    lat = build_lattice(meta, z)
    nb = 6
    rng = ChainRng(nb)
    rng.seed([11, 12, 13, 14, 15, 16])
    s = SweepSampling(lat, 4, nb=nb, rng=rng)
    s.configuration.import_configuration(np.broadcast_to(np.array(z["start_configuration"]), (nb,) + z["start_configuration"].shape))
    results = {}
    for which, (r_pinv, a_pinv) in {"regularised": (1e-3, 1e-6), "plain": (0.0, 0.0)}.items():
        obs = Observer(lat, enable_energy=True, enable_gradient=True, enable_natural_gradient=True)
        with obs:
            for _ in range(3):
                p, c = s()
                obs(p, c)
        rows = np.concatenate([np.asarray(d[2]) for d in obs._Deltas], axis=0)
        es = np.concatenate([d[1] for d in obs._Deltas])
        delta = np.concatenate([np.asarray(obs._Delta[l1][l2].storage).reshape(-1) for l1, l2 in lat.sites()]) / obs._total_weight
        energy = obs._total_energy_value()
        D, e = rows - delta, es - energy
        L, U = np.linalg.eigh(D @ D.T)
        num = r_pinv * L[-1] + a_pinv
        l_inv = np.array([0.0 if l <= 0 else 1 / (l * (1 + (num / l)**6)) for l in L])
        want = 2 * (D.T @ (U @ (l_inv * (U.T @ e))))
        got = obs.natural_gradient_by_direct_pseudo_inverse(r_pinv, a_pinv, ["unused"])
        flat = np.concatenate([np.asarray(got[l1][l2].transpose(obs._Delta[l1][l2].names).storage).reshape(-1) for l1, l2 in lat.sites()])
        assert np.abs(flat - want).max() <= 1e-9 * np.abs(want).max()
        results[which] = (D, e, flat)
    D, e, flat = results["plain"]
    keep = np.linalg.eigvalsh(D @ D.T) > 1e-9 * np.linalg.eigvalsh(D @ D.T)[-1]
    if keep.sum() >= len(e) - 1:                     # centred rows: rank Ns - 1 at most
        resid = D @ (flat / 2) - e
        assert np.abs(resid - resid.mean()).max() <= 1e-6 * max(1.0, np.abs(e).max())


@pytest.mark.parametrize("case", ["heis_3x3_D2_Dc4", "heis_4x4_D3_Dc5_truncating", "heisU1_4x4_d1_Dc6", "tJ_4x4_D1_Dc8", "hubbardFF_4x4_D1_Dc8"])
def test_gauge_fixing_matches_the_reference(case):
    """SamplingLattice.expand_dimension(1.0, 0) (lattice.py:821-919): the state is unchanged (amplitude at a truncation-free cut)
    and the TRUNCATED amplitude -- which depends on the gauge that was fixed -- equals the reference's after ITS gauge fixing
    (tests/golden/gauge_fixing.npz, written by make_golden.py gauge)"""
    import os
    from golden_loader import HERE, config_points
    from tnsp_b200.tetragono.configuration import Configuration
    meta, z = load(case)
    lat = build_lattice(meta, z)
    gold = np.load(os.path.join(HERE, "gauge_fixing.npz"))
    want_before, want_after, want_exact = gold[case]

    def amplitude(cut):
        conf = Configuration(lat, cut)
        if case + "_conf" in gold.files:          # fermionic cases: a configuration drawn by the reference's direct sampler
            conf.import_configuration(gold[case + "_conf"])
        else:
            for l1, row in enumerate(config_points(meta)):
                for l2, site in enumerate(row):
                    for o, p in site.items():
                        conf[l1, l2, o] = p
        return float(conf.hole(()))

    assert abs(amplitude(meta["Dc"]) - want_before) <= 1e-10 * abs(want_before)
    exact = amplitude(64)
    lat.expand_dimension(1.0, 0)
    assert abs(amplitude(64) - exact) <= 1e-12 * abs(exact)
    assert abs(amplitude(64) - want_exact) <= 1e-10 * abs(want_exact)
    assert abs(amplitude(meta["Dc"]) - want_after) <= 1e-9 * abs(want_after)


def test_bond_expansion_grows_the_bonds_and_perturbs_by_epsilon():
    meta, z = load("heis_3x3_D2_Dc4")
    lat = build_lattice(meta, z)
    from golden_loader import config_points
    from tnsp_b200.tetragono.configuration import Configuration

    def amplitude():
        conf = Configuration(lat, 64)
        for l1, row in enumerate(config_points(meta)):
            for l2, site in enumerate(row):
                for o, p in site.items():
                    conf[l1, l2, o] = p
        return float(conf.hole(()))

    before = amplitude()
    TAT.random.seed(77)
    lat.expand_dimension(3, 1e-6)
    assert all(lat[l1, l2].edge_by_name("R").dimension == 3 for l1 in range(3) for l2 in range(2))
    assert all(lat[l1, l2].edge_by_name("D").dimension == 3 for l1 in range(2) for l2 in range(3))
    assert abs(amplitude() - before) <= 1e-3 * abs(before)


def test_driver_fix_gauge_option_runs():
    meta, z = load("driver_heis_3x3_D2_Dc4_plain")
    lat = build_lattice(meta, z)
    conf = np.array(z["start_configuration"])
    TAT.random.seed(5)
    energies = [w["energy"][0] for w, _ in gradient_descent(lat, 4, 2, 0.01, sampling_method="sweep", configuration_cut_dimension=4,
                                                              sampling_configurations=conf, fix_gauge=True)]
    assert len(energies) == 2 and np.all(np.isfinite(energies))


def test_batched_ergodic_enumeration_equals_one_by_one():
    """nb configurations per call (a9): the exact energy of the 3 x 3 lattice (512 configurations) is the same whether they are
    enumerated one by one or in lock-step batches, also when the batch size does not divide the count (surplus chains weigh zero)"""
    meta, z = load("driver_heis_3x3_D2_Dc4_plain")
    lat = build_lattice(meta, z)
    (one, _), = list(gradient_descent(lat, sampling_method="ergodic", configuration_cut_dimension=4))
    for chains in (64, 100):
        (many, _), = list(gradient_descent(lat, sampling_method="ergodic", configuration_cut_dimension=4, chains=chains))
        assert abs(many["energy"][0] - one["energy"][0]) <= 1e-11 * abs(one["energy"][0])


def test_batched_ergodic_sequence_is_the_single_chain_sequence():
    from tnsp_b200.tetragono.sampling import ErgodicSampling
    meta, z = load("driver_heis_3x3_D2_Dc4_plain")
    lat = build_lattice(meta, z)
    for rank, size in ((0, 1), (1, 3)):
        single = ErgodicSampling(lat, 4, rank=rank, size=size)
        batch = ErgodicSampling(lat, 4, rank=rank, size=size, nb=5)
        want = np.stack([single()[1].export_configuration() for _ in range(10)])
        got = np.concatenate([batch()[1].export_configuration() for _ in range(2)])
        assert np.array_equal(got, want)


def test_driver_with_lockstep_direct_sampling():
    meta, z = load("driver_heis_3x3_D2_Dc4_plain")
    lat = build_lattice(meta, z)
    TAT.random.seed(4)
    energies = [w["energy"][0] for w, _ in gradient_descent(lat, 24, 3, 0.02, chains=8, sampling_method="direct", configuration_cut_dimension=4,
                                                              direct_sampling_cut_dimension=4, use_natural_gradient=True,
                                                              conjugate_gradient_method_step=3, use_fix_relative_step_size=True)]
    assert len(energies) == 3 and np.all(np.isfinite(energies))


def test_driver_embeds_symmetric_lattices_for_lockstep_chains():
    """a U(1) lattice with chains > 1: the driver samples its charge-dense embedding and projects the gradient back; one chain
    through the embedded path must reproduce the symmetric single-chain path step by step (same seeds -> same configurations)"""
    meta, z = load("driver_heisU1_4x4_d1_Dc6_line_search")
    conf0 = np.array(z["start_configuration"])
    results = {}
    for which in ("symmetric", "embedded"):
        lat = build_lattice(meta, z)
        conf = conf0.copy()
        TAT.random.seed(12)
        kw = dict(sampling_method="sweep", configuration_cut_dimension=meta["Dc"], use_natural_gradient=True, conjugate_gradient_method_step=2)
        if which == "symmetric":
            energies = [w["energy"] for w, _ in gradient_descent(lat, 5, 2, 0.01, chain_seeds=[77], sampling_configurations=conf, **kw)]
        else:
            # two identical chains (same seed): the batch statistics equal the single chain's
            energies = [w["energy"] for w, _ in gradient_descent(lat, 10, 2, 0.01, chains=2, chain_seeds=[77, 77],
                                                                 sampling_configurations=conf, **kw)]
        results[which] = (np.array(energies), [np.asarray(lat[l1, l2].storage).copy() for l1 in range(lat.L1) for l2 in range(lat.L2)])
    e_s, t_s = results["symmetric"]
    e_e, t_e = results["embedded"]
    assert np.abs(e_s[:, 0] - e_e[:, 0]).max() <= 1e-8 * np.abs(e_s[:, 0]).max()
    scale = max(np.abs(a).max() for a in t_s)
    for a, b in zip(t_s, t_e):
        assert np.abs(a - b).max() <= 1e-7 * scale

"""Charge-dense embedding of a U(1) PEPS (tnsp_b200/tetragono/dense_embedding.py): the lock-step batch
engine evaluates symmetric models as dense tensors with exact zeros; amplitudes, local energies,
holes and sweep trajectories must equal the symmetric (sector) evaluation -- which test_vmc_golden.py
pins to the unmodified reference (fixture j1j2U1_4x4_d1_Dc9) -- within 1e-10."""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import build_lattice, config_points, load
from tnsp_b200.tetragono import dense_embedding as de
from tnsp_b200.tetragono import models
from tnsp_b200.tetragono.configuration import Configuration
from tnsp_b200.tetragono.observer import Observer
from tnsp_b200.tetragono.sampling import ChainRng, SweepSampling

RTOL = 1e-10


def _u1_lattice():
    meta, z = load("j1j2U1_4x4_d1_Dc9")
    return meta, z, build_lattice(meta, z)


def _sz0_configurations(n, L1, L2, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        flat = np.array([0, 1] * (L1 * L2 // 2))
        rng.shuffle(flat)
        out.append(flat.reshape(L1, L2, 1))
    return np.stack(out)


def test_dense_reproduces_reference_fixture():
    """dense embedding vs the reference's own numbers (ws, E_s) on the fixture configuration"""
    meta, z, lat = _u1_lattice()
    dl = de.embed_lattice(lat)
    conf = Configuration(dl, meta["Dc"])
    conf.import_configuration(de.embed_configuration(lat, config_points(meta)))
    ws = conf.hole(())
    assert abs(float(ws) - z["ws"][0]) <= RTOL * abs(z["ws"][0])
    obs = Observer(dl, enable_energy=True)
    with obs:
        obs(float(ws)**2, conf)
    e = obs._whole_result_reweight["energy"] / obs._total_weight
    assert abs(e - z["energy_s"][0]) <= RTOL * abs(z["energy_s"][0])


def test_lockstep_dense_equals_sector_evaluation():
    """nb different Sz=0 configurations: dense lock-step batch vs one-by-one U(1) sector evaluation"""
    meta, z, lat = _u1_lattice()
    L1, L2, Dc = meta["L1"], meta["L2"], meta["Dc"]
    dl = de.embed_lattice(lat)
    nb = 6
    confs = _sz0_configurations(nb, L1, L2, 3)
    batch = Configuration(dl, Dc, nb)
    batch.import_configuration(confs)
    ws_b = np.asarray(batch.hole(()).storage).reshape(-1)
    obs_b = Observer(dl, enable_energy=True, enable_gradient=True)
    with obs_b:
        obs_b(ws_b**2, batch)
    holes_b = batch.holes()
    S = lat.Symmetry
    e_sum = 0.0
    for c in range(nb):
        one = Configuration(lat, Dc)
        for l1 in range(L1):
            for l2 in range(L2):
                one[l1, l2, 0] = (S(+1) if confs[c, l1, l2, 0] == 0 else S(-1), 0)
        ws = float(one.hole(()))
        assert abs(ws - ws_b[c]) <= RTOL * max(abs(ws), 1e-300)
        if ws == 0:   # charge flow forbids this configuration: zero weight, skipped by the observers
            assert ws_b[c] == 0
            continue
        obs = Observer(lat, enable_energy=True, enable_gradient=True)
        with obs:
            obs(ws**2, one)
        e_sum += obs._whole_result_reweight["energy"]
        holes = one.holes()
        for l1 in range(L1):
            for l2 in range(L2):
                want = holes[l1][l2].clear_symmetry()
                names = [n for n in want.names if n != "T"]
                w = np.asarray(want.transpose(names + (["T"] if "T" in want.names else [])).storage).reshape(-1)
                g = np.atleast_2d(np.asarray(holes_b[l1][l2].transpose(names).storage))[c]
                assert np.abs(g - w).max() <= RTOL * max(np.abs(w).max(), 1e-300)
    assert abs(obs_b._whole_result_reweight["energy"] - e_sum) <= 1e-9 * abs(e_sum)


def test_sweep_trajectory_dense_equals_reference_trajectory():
    """same seed -> the dense lock-step chain visits the configurations the unmodified reference visited"""
    meta, z, lat = _u1_lattice()
    dl = de.embed_lattice(lat)
    TAT.random.seed(meta["seed"])
    s = SweepSampling(dl, meta["Dc"], None, models.nearest_neighbour_terms(dl))
    s.configuration.import_configuration(de.embed_configuration(lat, config_points(meta)))
    obs = Observer(dl, enable_energy=True, enable_gradient=True)
    with obs:
        for i in range(meta["n_samples"]):
            p, c = s()
            # fixture stores total edge indices; index 0 = charge +1 (up), 1 = charge -1 (down) in both pictures
            assert np.array_equal(c.export_configuration(), z["traj_config"][i]), f"trajectory diverged at sample {i}"
            assert abs(p - z["traj_possibility"][i]) <= RTOL * z["traj_possibility"][i]
            obs(p, c)
    assert np.abs(np.array(obs.total_energy) - z["traj_energy"]).max() <= 1e-9 * np.abs(z["traj_energy"]).max()
    grad = de.project_gradient(lat, obs.gradient)
    mod = TAT.BoseU1
    from golden_loader import tensor_from
    gs = max(np.abs(z[meta["gradient"][l1][l2]["storage"]]).max() for l1 in range(4) for l2 in range(4))
    for l1 in range(4):
        for l2 in range(4):
            want = tensor_from(mod, meta["gradient"][l1][l2], z)
            got = grad[l1][l2]
            if got.names != want.names:
                got = got.transpose(want.names)
            assert np.abs(np.asarray(got.storage) - np.asarray(want.storage)).max() <= 1e-9 * gs

"""The benched configuration at its REAL size (BASELINE cfg2: 6x6 J1-J2 U(1), D = 2+2+2, Dc = 36) against the fixture written by the
UNMODIFIED reference (tests/golden/make_golden.py cfg2size): cache-cold amplitude, local energy and holes of a lock-step batch on
the sector-compact engine, and the sweep trajectory + gradient of a single sector-compact chain from the reference's seed.
(Lock-step batches with more than one chain warm their environment caches in a different order than a single chain, so at a
truncating Dc only cache-cold quantities are comparable chain by chain -- SURVEY.md section 7.)"""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import build_lattice, config_points, load, tensor_from
from tnsp_b200.tetragono import models
from tnsp_b200.tetragono.configuration import Configuration
from tnsp_b200.tetragono.observer import Observer, _blocks_of
from tnsp_b200.tetragono.sampling import SweepSampling

NAME = "j1j2U1_6x6_d2_Dc36"


def at_size_cold_check(nb, engine=None, holes=True):
    meta, z = load(NAME)
    lat = build_lattice(meta, z)
    L1, L2, Dc = meta["L1"], meta["L2"], meta["Dc"]
    conf = Configuration(lat, Dc, nb, engine=engine)
    for l1, row in enumerate(config_points(meta)):
        for l2, site in enumerate(row):
            for o, pt in site.items():
                conf[l1, l2, o] = pt
    ws = np.asarray(conf.hole(()).storage).reshape(-1)
    assert ws.shape == (nb,)
    assert np.abs(ws - z["ws"][0]).max() <= 1e-10 * abs(z["ws"][0])
    obs = Observer(lat, enable_energy=True, enable_gradient=holes)
    with obs:
        obs(ws**2, conf)
    e = obs._whole_result_reweight["energy"] / obs._total_weight
    assert abs(e - z["energy_s"][0]) <= 1e-10 * abs(z["energy_s"][0])
    if holes:
        B = TAT.tensor._bk.get()
        got = conf.holes()
        for l1 in range(L1):
            for l2 in range(L2):
                want = tensor_from(TAT.BoseU1, meta["holes"][l1][l2], z)
                target = obs._Delta[l1][l2]
                w = np.asarray(want.transpose(target.names).storage).reshape(-1)
                h = got[l1][l2]
                if getattr(h, "is_ragged", False):
                    g = np.atleast_2d(B.to_numpy(_blocks_of(h.transpose(target.names), target)))
                else:
                    g = np.atleast_2d(np.asarray(h.transpose(target.names).storage))
                assert np.abs(g - w[None, :]).max() <= 1e-9 * np.abs(w).max(), (l1, l2)
    return float(np.abs(ws - z["ws"][0]).max() / abs(z["ws"][0])), float(abs(e - z["energy_s"][0]) / abs(z["energy_s"][0]))


def at_size_trajectory_check():
    meta, z = load(NAME)
    lat = build_lattice(meta, z)
    TAT.random.seed(meta["seed"])
    s = SweepSampling(lat, meta["Dc"], None, models.nearest_neighbour_terms(lat))
    s.configuration = Configuration(lat, meta["Dc"], 1, engine="sector")
    for l1, row in enumerate(config_points(meta)):
        for l2, site in enumerate(row):
            for o, pt in site.items():
                s.configuration[l1, l2, o] = pt
    obs = Observer(lat, enable_energy=True, enable_gradient=True, enable_natural_gradient=True)
    with obs:
        for i in range(meta["n_samples"]):
            p, c = s()
            assert np.array_equal(c.export_configuration(), z["traj_config"][i]), f"trajectory diverged at sample {i}"
            assert abs(p - z["traj_possibility"][i]) <= 1e-9 * z["traj_possibility"][i]
            obs(p, c)
    assert np.abs(np.array(obs.total_energy) - z["traj_energy"]).max() <= 1e-9 * np.abs(z["traj_energy"]).max()
    L1, L2 = meta["L1"], meta["L2"]
    for key, grad, tol in (("gradient", obs.gradient, 1e-9), ("natural_gradient", obs.natural_gradient_by_conjugate_gradient(meta["cg_step"], 0.0), 1e-8)):
        gs = max(np.abs(z[meta[key][l1][l2]["storage"]]).max() for l1 in range(L1) for l2 in range(L2))
        for l1 in range(L1):
            for l2 in range(L2):
                want = tensor_from(TAT.BoseU1, meta[key][l1][l2], z)
                got = grad[l1][l2]
                if got.names != want.names:
                    got = got.transpose(want.names)
                assert np.abs(np.asarray(got.storage) - np.asarray(want.storage)).max() <= tol * gs, (key, l1, l2)


def test_cfg2_at_size_cold_two_chains():
    at_size_cold_check(2, holes=False)

"""End-to-end parity of the sampling-VMC path against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py): amplitude, local energy, holes (per configuration, cache
cold), then a whole sweep-sampling trajectory from a fixed seed with its energy, gradient and SR
natural gradient.  Tolerance 1e-10 relative (float64) as BASELINE.json's north_star states; the
configurations of the trajectory are integers and must match exactly.
"""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import CASES, build_lattice, config_points, load, tensor_from
from tnsp_b200.tetragono.configuration import Configuration
from tnsp_b200.tetragono.observer import Observer
from tnsp_b200.tetragono.sampling import SweepSampling

RTOL = 1e-10


def _close(a, b, rtol=RTOL, scale=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    if scale is None:
        scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    assert np.abs(a - b).max() <= rtol * scale, f"max rel err {np.abs(a - b).max() / scale:.3e}"


def _set_config(conf, points):
    for l1, row in enumerate(points):
        for l2, site in enumerate(row):
            for o, p in site.items():
                conf[l1, l2, o] = p


@pytest.mark.parametrize("case", CASES)
def test_amplitude_energy_holes(case):
    meta, z = load(case)
    lat = build_lattice(meta, z)
    conf = Configuration(lat, meta["Dc"])
    _set_config(conf, config_points(meta))
    ws = conf.hole(())
    assert ws.names == meta["ws_names"]
    _close(float(ws), z["ws"][0])
    obs = Observer(lat, enable_energy=True, enable_gradient=True)
    with obs:
        obs(float(ws)**2, conf)
    _close(obs._whole_result_reweight["energy"] / obs._total_weight, z["energy_s"][0])
    holes = conf.holes()
    mod = getattr(TAT, meta["symmetry"])
    for l1 in range(meta["L1"]):
        for l2 in range(meta["L2"]):
            want = tensor_from(mod, meta["holes"][l1][l2], z)
            got = holes[l1][l2]
            assert got.names == want.names and got._edges == want._edges
            _close(np.asarray(got.storage), np.asarray(want.storage))


@pytest.mark.parametrize("case", CASES)
def test_sweep_trajectory_gradient(case):
    meta, z = load(case)
    lat = build_lattice(meta, z)
    TAT.random.seed(meta["seed"])
    hopping = None
    if meta.get("sweep_nearest_neighbour_only"):
        from tnsp_b200.tetragono.models import nearest_neighbour_terms
        hopping = nearest_neighbour_terms(lat)
    sampling = SweepSampling(lat, meta["Dc"], None, hopping)
    _set_config(sampling.configuration, config_points(meta))
    obs = Observer(lat, enable_energy=True, enable_gradient=True, enable_natural_gradient=True)
    with obs:
        for i in range(meta["n_samples"]):
            p, c = sampling()
            assert np.array_equal(c.export_configuration(), z["traj_config"][i]), f"trajectory diverged at sample {i}"
            _close(p, z["traj_possibility"][i])
            obs(p, c)
    _close(np.array(obs.total_energy), z["traj_energy"], 1e-9)
    mod = getattr(TAT, meta["symmetry"])
    grad = obs.gradient
    ng = obs.natural_gradient_by_conjugate_gradient(meta["cg_step"], 0.0)
    # tolerance relative to the largest gradient entry of the lattice (some tensors are exactly 0 in the reference)
    gs = max(np.abs(z[meta["gradient"][l1][l2]["storage"]]).max() for l1 in range(meta["L1"]) for l2 in range(meta["L2"]))
    ns = max(np.abs(z[meta["natural_gradient"][l1][l2]["storage"]]).max() for l1 in range(meta["L1"]) for l2 in range(meta["L2"]))
    for l1 in range(meta["L1"]):
        for l2 in range(meta["L2"]):
            want = tensor_from(mod, meta["gradient"][l1][l2], z)
            assert grad[l1][l2].names == want.names and grad[l1][l2]._edges == want._edges
            _close(np.asarray(grad[l1][l2].storage), np.asarray(want.storage), 1e-9, gs)
            want = tensor_from(mod, meta["natural_gradient"][l1][l2], z)
            _close(np.asarray(ng[l1][l2].storage), np.asarray(want.storage), 1e-8, ns)

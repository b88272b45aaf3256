"""Sector-compact lock-step engine (tnsp_b200/TAT/ragged.py): chains with DIFFERENT symmetry sectors evaluated in one batch must
equal the one-by-one block-symmetric evaluation -- which test_vmc_golden.py pins to the unmodified reference (fixture
j1j2U1_4x4_d1_Dc9) -- within 1e-10: amplitudes, local energies, holes, gradients and sweep trajectories."""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import build_lattice, config_points, load, tensor_from
from tnsp_b200.TAT import ragged
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


def test_tensor_ops_equal_block_symmetric_ops():
    """contract / qr / svd (with a truncating cut) / norms on sector-compact tensors vs the block-symmetric device tensors"""
    U1 = TAT.BoseU1
    rng = np.random.default_rng(0)
    e_l = U1.Edge([(-1, 3), (0, 4), (1, 3)])
    e_p = U1.Edge([(-1, 1), (1, 1)])
    nb = 3
    a = U1.D.Tensor(["L", "P", "R"], [e_l, e_p, e_l.conjugate()])
    b = U1.D.Tensor(["L", "Q", "R"], [e_l, e_p.conjugate(), e_l.conjugate()])
    a = U1.D.Tensor.from_batch(a.names, a._edges, rng.standard_normal((nb, a.storage.size)))
    b = U1.D.Tensor.from_batch(b.names, b._edges, rng.standard_normal((nb, b.storage.size)))
    ra, rb = ragged.RTensor.from_symmetric(a), ragged.RTensor.from_symmetric(b)

    def dense(t):
        return np.asarray(t.clear_symmetry()._host()).reshape(nb, -1)

    def rd(t):
        return np.asarray(TAT.tensor._bk.get().to_numpy(t.to_dense()))

    assert np.array_equal(rd(ra), dense(a))
    c, rc = a.contract(b, {("R", "L")}), ra.contract(rb, {("R", "L")})
    assert c.names == rc.names and np.abs(rd(rc) - dense(c)).max() < 1e-13
    q, r = rc.qr("r", {"R"}, "X", "Y")
    assert np.abs(rd(q.contract(r, {("X", "Y")})) - rd(rc)).max() < 1e-12
    qq = q.conjugate().edge_rename({"X": "X2"}).contract(q, {(n, n) for n in ("L", "P", "Q")})
    eye = rd(qq).reshape(nb, q.core.edges[-1].dim, -1)
    assert np.abs(eye - np.round(eye)).max() < 1e-12          # orthonormal columns (identity on the live bond indices)
    u, s, v = rc.svd({"L", "P"}, "U", "V", "SU", "SV", 6)
    u0, s0, v0 = c.svd({"L", "P"}, "U", "V", "SU", "SV", 6)
    rec = u.contract(s, {("U", "SU")}).contract(v, {("SV", "V")})
    rec0 = u0.contract(s0, {("U", "SU")}).contract(v0, {("SV", "V")})
    assert np.abs(rd(rec) - dense(rec0.transpose(rec.names))).max() < 1e-12
    n2 = rc.conjugate().contract(rc, {(n, n) for n in rc.names})
    assert np.allclose(n2.storage, np.asarray(rc.norm_2().numpy())**2, rtol=1e-13)
    assert np.allclose(np.asarray(rc.norm_2().numpy()), np.asarray(c.norm_2().numpy()), rtol=1e-13)
    assert np.allclose(np.asarray((rc / rc.norm_max()).norm_max().numpy()), 1.0)


def test_sector_engine_reproduces_reference_fixture():
    meta, z, lat = _u1_lattice()
    conf = Configuration(lat, meta["Dc"], 1, engine="sector")
    for l1, row in enumerate(config_points(meta)):
        for l2, site in enumerate(row):
            for o, pt in site.items():
                conf[l1, l2, o] = pt
    ws = conf.hole(())
    assert abs(float(ws) - z["ws"][0]) <= RTOL * abs(z["ws"][0])
    obs = Observer(lat, enable_energy=True)
    with obs:
        obs(float(ws)**2, conf)
    e = obs._whole_result_reweight["energy"] / obs._total_weight
    assert abs(e - z["energy_s"][0]) <= RTOL * abs(z["energy_s"][0])


def test_lockstep_sector_equals_one_by_one_evaluation():
    """nb different Sz=0 configurations (different sector structures) in one batch vs one-by-one block-symmetric evaluation"""
    meta, z, lat = _u1_lattice()
    L1, L2, Dc = meta["L1"], meta["L2"], meta["Dc"]
    nb = 5
    confs = _sz0_configurations(nb, L1, L2, 3)
    batch = Configuration(lat, Dc, nb)
    assert batch._ragged
    batch.import_configuration(confs)
    ws_b = np.asarray(batch.hole(()).storage).reshape(-1)
    obs_b = Observer(lat, enable_energy=True, enable_gradient=True)
    with obs_b:
        obs_b(ws_b**2, batch)
    holes_b = batch.holes()
    S = lat.Symmetry
    e_sum, w_sum = 0.0, 0.0
    delta = None
    for c in range(nb):
        one = Configuration(lat, Dc)
        for l1 in range(L1):
            for l2 in range(L2):
                one[l1, l2, 0] = (S(+1) if confs[c, l1, l2, 0] == 0 else S(-1), 0)
        ws = float(one.hole(()))
        assert abs(ws - ws_b[c]) <= RTOL * max(abs(ws), 1e-300)
        if ws == 0:
            assert ws_b[c] == 0
            continue
        obs = Observer(lat, enable_energy=True, enable_gradient=True)
        with obs:
            obs(ws**2, one)
        e_sum += obs._whole_result_reweight["energy"]
        holes = one.holes()
        from tnsp_b200.tetragono.observer import _blocks_of
        for l1 in range(L1):
            for l2 in range(L2):
                want = holes[l1][l2]
                target = obs._Delta[l1][l2]
                w = np.asarray(want.transpose(target.names).storage).reshape(-1)
                g = np.atleast_2d(TAT.tensor._bk.get().to_numpy(_blocks_of(holes_b[l1][l2].transpose(target.names), target)))[c]
                assert np.abs(g - w).max() <= RTOL * max(np.abs(w).max(), 1e-300)
    assert abs(obs_b._whole_result_reweight["energy"] - e_sum) <= 1e-9 * abs(e_sum)


def test_sweep_trajectory_sector_equals_reference_trajectory():
    """same seed -> the sector-compact chain visits the configurations the unmodified reference visited; energy and gradient"""
    meta, z, lat = _u1_lattice()
    TAT.random.seed(meta["seed"])
    s = SweepSampling(lat, meta["Dc"], None, models.nearest_neighbour_terms(lat))
    s.configuration = Configuration(lat, meta["Dc"], 1, engine="sector")
    pts = config_points(meta)
    for l1, row in enumerate(pts):
        for l2, site in enumerate(row):
            for o, pt in site.items():
                s.configuration[l1, l2, o] = pt
    obs = Observer(lat, enable_energy=True, enable_gradient=True)
    with obs:
        for i in range(meta["n_samples"]):
            p, c = s()
            assert np.array_equal(c.export_configuration(), z["traj_config"][i]), f"trajectory diverged at sample {i}"
            assert abs(p - z["traj_possibility"][i]) <= RTOL * z["traj_possibility"][i]
            obs(p, c)
    assert np.abs(np.array(obs.total_energy) - z["traj_energy"]).max() <= 1e-9 * np.abs(z["traj_energy"]).max()
    grad = obs.gradient
    gs = max(np.abs(z[meta["gradient"][l1][l2]["storage"]]).max() for l1 in range(4) for l2 in range(4))
    for l1 in range(4):
        for l2 in range(4):
            want = tensor_from(TAT.BoseU1, meta["gradient"][l1][l2], z)
            got = grad[l1][l2]
            if got.names != want.names:
                got = got.transpose(want.names)
            assert np.abs(np.asarray(got.storage) - np.asarray(want.storage)).max() <= 1e-9 * gs


def test_learnt_capacities_and_overflow_correction():
    """buffer capacities: a calibration batch learns the largest stored size per operation, later batches allocate CAP_FACTOR x that and
    still give the same amplitudes; a capacity that is too small stores the chain EMPTY (amplitude zero), is counted, reported by the
    sampler as a warning and corrected by a larger factor -- never silent, never out of bounds"""
    import warnings
    from tnsp_b200.tetragono import sampling
    from tnsp_b200.tetragono.sampling import calibrate_sector_engine
    meta, z, lat = _u1_lattice()
    L1, L2, Dc = meta["L1"], meta["L2"], meta["Dc"]
    nb = 4
    confs = _sz0_configurations(nb, L1, L2, 5)
    saved = (dict(ragged._CAPS), dict(ragged._LEARN), ragged.CAP_FACTOR, ragged.CAP_MIN_CHAINS, dict(sampling._CAPACITY_BUMPS))
    try:
        ragged._CAPS.clear()
        ragged.CAP_MIN_CHAINS = 1
        ref = Configuration(lat, Dc, nb)                                   # learning phase: dense bounds
        ragged._LEARN.update(all=True, cycles=0)
        ref.import_configuration(confs)
        ws_ref = np.asarray(ref.hole(()).storage).reshape(-1).copy()
        calibrate_sector_engine(lat, Dc, confs[0], models.nearest_neighbour_terms(lat), chains=nb, sweeps=1)
        assert not ragged._LEARN["all"] and len(ragged._CAPS) > 0
        capped = Configuration(lat, Dc, nb)
        capped.import_configuration(confs)
        ws = np.asarray(capped.hole(()).storage).reshape(-1)
        assert np.abs(ws - ws_ref).max() <= RTOL * np.abs(ws_ref).max()
        assert TAT.tensor._bk.get().rt_overflow() == 0
        # now far too small (this lattice's tensors store <= 42 elements, below the floor of a learnt capacity, so the allocation
        # itself is shrunk): tensors beyond 8 elements are dropped -- stored EMPTY, counted, nothing written out of bounds
        B = TAT.tensor._bk.get()
        import torch
        ragged.CAP_FACTOR = 1.0
        sampling._CAPACITY_BUMPS.update(n=0, dropped=0)
        real_cap, real_alloc = ragged._cap, B.rt_alloc
        ragged._cap = lambda key, dense, nb=None: (min(dense, 8), False)
        B.rt_alloc = lambda nb_, size: torch.zeros((nb_, max(int(size), 2)), dtype=torch.float64)
        try:
            small = Configuration(lat, Dc, nb)
            small.import_configuration(confs)
            ws0 = np.asarray(small.hole(()).storage).reshape(-1)
        finally:
            ragged._cap, B.rt_alloc = real_cap, real_alloc
        assert np.all(ws0 == 0.0)
        assert B.rt_overflow(clear=False) > 0
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter("always")
            sampling._check_capacity()
        assert any("exceeded their learnt capacity" in str(w.message) for w in rec)
        assert ragged.CAP_FACTOR == 1.5 and sampling._CAPACITY_BUMPS["n"] == 1
        assert TAT.tensor._bk.get().rt_overflow() == 0                       # the counter was cleared
    finally:
        ragged._CAPS.clear()
        ragged._CAPS.update(saved[0])
        ragged._LEARN.update(saved[1])
        ragged.CAP_FACTOR, ragged.CAP_MIN_CHAINS = saved[2], saved[3]
        sampling._CAPACITY_BUMPS.update(saved[4])


def test_table_cache_does_not_grow_with_the_number_of_sweeps():
    """group tables are cached on their first label array; entries keyed on label arrays of environments that no longer exist must be
    purged -- the bond labels of the PEPS site tensors live as long as the lattice, and their caches once grew by ~2.4 GB per step of
    the cfg2 bench (one sorted table per regrouping of every environment of every sweep)"""
    import gc
    import torch
    meta, z, lat = _u1_lattice()
    L1, L2, Dc = meta["L1"], meta["L2"], meta["Dc"]
    nb = 3
    rng = ChainRng(nb)
    rng.seed([11, 12, 13])
    s = SweepSampling(lat, Dc, None, models.nearest_neighbour_terms(lat), nb=nb, rng=rng)
    s.configuration.import_configuration(_sz0_configurations(nb, L1, L2, 9))
    obs = Observer(lat, enable_energy=True, enable_gradient=True)

    def cached():
        gc.collect()
        n = 0
        for o in gc.get_objects():
            if isinstance(o, torch.Tensor):
                n += len(getattr(o, "__dict__", {}).get("_rt_tables", ()))
        return n

    counts = []
    for _ in range(5):
        with obs:
            p, c = s()
            obs(p, c)
        del p, c
        counts.append(cached())
    # steady state after the first sweeps: no systematic growth (accept / reject patterns make the count fluctuate a little)
    assert counts[4] <= 1.25 * counts[1] + 16, counts

"""Fermionic sector-compact tensors (TAT/ragged_fermi.py): every signed operation against the block-symmetric device tensors,
whose planner is pinned to the unmodified reference (tests/test_tat_vs_reference.py).  Random tensors, every fermionic integer
symmetry, random arrows; chains of a batch carry different data (and, through unit edges, different parities)."""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from tnsp_b200.TAT import ragged

SYMS = ["FermiU1", "FermiU1BoseU1", "FermiU1FermiU1"]


def _edge(mod, rng, n_seg, max_dim, arrow):
    S = mod.Symmetry
    seen, segs = set(), []
    while len(segs) < n_seg:
        s = tuple(int(rng.integers(-1, 2)) for _ in range(S.length))
        if s in seen:
            continue
        seen.add(s)
        segs.append((S(*s), int(rng.integers(1, max_dim + 1))))
    return mod.Edge(segs, arrow)


def _rand(mod, rng, names, edges, nb):
    t = mod.D.Tensor(names, edges)
    return mod.D.Tensor.from_batch(t.names, t._edges, rng.standard_normal((nb, t.storage.size)))


def _dense(t, nb):
    """dense expansion of a block-symmetric tensor (plain placement of the blocks, no signs)"""
    dims = [e.dimension for e in t._edges]
    h = np.atleast_2d(t._host())
    out = np.zeros([h.shape[0]] + dims)
    starts = t._segment_starts()
    for b, pos in enumerate(t._table.positions):
        bd = [int(d) for d in t._table.dims[b]]
        off, size = int(t._table.offsets[b]), int(t._table.sizes[b])
        sl = (slice(None),) + tuple(slice(int(starts[i][int(p)]), int(starts[i][int(p)]) + bd[i]) for i, p in enumerate(pos))
        out[sl] = h[:, off:off + size].reshape([h.shape[0]] + bd)
    return out.reshape(h.shape[0], -1)


def _rd(t):
    return np.asarray(TAT.tensor._bk.get().to_numpy(t.to_dense()))


@pytest.mark.parametrize("sym", SYMS)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_signed_operations_equal_block_symmetric_tensors(sym, seed):
    mod = getattr(TAT, sym)
    rng = np.random.default_rng(100 * seed + len(sym))
    nb = 3
    ar = [bool(x) for x in rng.integers(0, 2, size=6)]
    eA, eB, eC, eK, eL = (_edge(mod, rng, 3, 2, ar[i]) for i in range(5))
    a = _rand(mod, rng, ["A", "K", "B", "L"], [eA, eK, eB, eL], nb)
    b = _rand(mod, rng, ["L2", "C", "K2"], [eL.conjugate(), eC, eK.conjugate()], nb)
    ra, rb = ragged.RTensor.from_symmetric(a), ragged.RTensor.from_symmetric(b)
    assert np.array_equal(_rd(ra), _dense(a, nb))
    # transpose (edge_operator.hpp:521-555)
    order = ["B", "L", "A", "K"]
    assert np.abs(_rd(ra.transpose(order)) - _dense(a.transpose(order), nb)).max() < 1e-14
    # conjugate with and without the trivial metric (conjugate.hpp:48-97)
    for tm in (False, True):
        assert np.abs(_rd(ra.conjugate(tm)) - _dense(a.conjugate(tm), nb)).max() < 1e-14
    # contract over two edges (contract.hpp:306-620)
    c = a.contract(b, {("K", "K2"), ("L", "L2")})
    rc = ra.contract(rb, {("K", "K2"), ("L", "L2")})
    assert c.names == rc.names
    scale = max(np.abs(_dense(c, nb)).max(), 1e-300)
    assert np.abs(_rd(rc) - _dense(c, nb)).max() <= 1e-12 * scale
    # contract of a conjugated operand, all edges: <a|a>
    full = a.conjugate().contract(a, {(n, n) for n in a.names})
    rfull = ra.conjugate().contract(ra, {(n, n) for n in ra.names})
    assert np.allclose(np.asarray(rfull.storage).reshape(-1), np.asarray(full.storage).reshape(-1), rtol=1e-12)
    # qr / svd: the factors contract back to the tensor under the fermionic contraction rules
    q, r = rc.qr("r", {"C"}, "X", "Y")
    assert np.abs(_rd(q.contract(r, {("X", "Y")}).transpose(rc.names)) - _rd(rc)).max() <= 1e-11 * scale
    u, s, v = rc.svd({"A"}, "U", "V", "SU", "SV", -1)
    usv = u.contract(s, {("U", "SU")}).contract(v, {("SV", "V")}).transpose(rc.names)
    assert np.abs(_rd(usv) - _rd(rc)).max() <= 1e-11 * scale
    # a truncating cut keeps the same singular values as the block-symmetric svd
    cut = 3
    u1, s1, v1 = rc.svd({"A", "B"}, "U", "V", "SU", "SV", cut)
    u0, s0, v0 = c.svd({"A", "B"}, "U", "V", "SU", "SV", cut)
    rec1 = u1.contract(s1, {("U", "SU")}).contract(v1, {("SV", "V")}).transpose(rc.names)
    rec0 = u0.contract(s0, {("U", "SU")}).contract(v0, {("SV", "V")}).transpose(c.names)
    assert np.abs(_rd(rec1) - _dense(rec0, nb)).max() <= 1e-10 * scale


# ---- VMC level: fermionic lock-step batches on the sector-compact engine ---------------------------------------------------
from golden_loader import build_lattice, config_points, load, tensor_from  # noqa: E402
from tnsp_b200.tetragono.configuration import Configuration  # noqa: E402
from tnsp_b200.tetragono.observer import Observer, _blocks_of  # noqa: E402
from tnsp_b200.tetragono.sampling import SweepSampling  # noqa: E402


@pytest.mark.parametrize("name", ["tJ_4x4_D1_Dc8", "hubbardFF_4x4_D1_Dc8"])
def test_fermionic_fixture_single_chain_sector_engine(name):
    """one sector-compact chain reproduces the unmodified reference: cache-cold ws / E_s, then the sweep trajectory + gradient"""
    meta, z = load(name)
    lat = build_lattice(meta, z)
    conf = Configuration(lat, meta["Dc"], 1, engine="sector")
    for l1, row in enumerate(config_points(meta)):
        for l2, site in enumerate(row):
            for o, pt in site.items():
                conf[l1, l2, o] = pt
    ws = float(np.asarray(conf.hole(()).storage).reshape(-1)[0])
    assert abs(ws - z["ws"][0]) <= 1e-10 * abs(z["ws"][0])
    obs = Observer(lat, enable_energy=True)
    with obs:
        obs(ws**2, conf)
    e = obs._whole_result_reweight["energy"] / obs._total_weight
    assert abs(e - z["energy_s"][0]) <= 1e-10 * abs(z["energy_s"][0])
    TAT.random.seed(meta["seed"])
    s = SweepSampling(lat, meta["Dc"])
    s.configuration = Configuration(lat, meta["Dc"], 1, engine="sector")
    for l1, row in enumerate(config_points(meta)):
        for l2, site in enumerate(row):
            for o, pt in site.items():
                s.configuration[l1, l2, o] = pt
    obs = Observer(lat, enable_energy=True, enable_gradient=True)
    with obs:
        for i in range(meta["n_samples"]):
            p, c = s()
            assert np.array_equal(c.export_configuration(), z["traj_config"][i]), f"trajectory diverged at sample {i}"
            assert abs(p - z["traj_possibility"][i]) <= 1e-9 * z["traj_possibility"][i]
            obs(p, c)
    assert np.abs(np.array(obs.total_energy) - z["traj_energy"]).max() <= 1e-9 * np.abs(z["traj_energy"]).max()
    mod = getattr(TAT, meta["symmetry"])
    grad = obs.gradient
    L1, L2 = meta["L1"], meta["L2"]
    gs = max(np.abs(z[meta["gradient"][l1][l2]["storage"]]).max() for l1 in range(L1) for l2 in range(L2))
    for l1 in range(L1):
        for l2 in range(L2):
            want = tensor_from(mod, meta["gradient"][l1][l2], z)
            got = grad[l1][l2]
            if got.names != want.names:
                got = got.transpose(want.names)
            assert np.abs(np.asarray(got.storage) - np.asarray(want.storage)).max() <= 1e-9 * gs


@pytest.mark.parametrize("name", ["tJ_4x4_D1_Dc8", "hubbardFF_4x4_D1_Dc8"])
def test_fermionic_lockstep_batch_equals_one_by_one(name):
    """the configurations the reference's chain visited, evaluated cache-cold as ONE lock-step batch (different charges, hence
    different parities and sectors per chain) vs one by one on the block-symmetric tensors: ws, E_s, holes"""
    meta, z = load(name)
    lat = build_lattice(meta, z)
    L1, L2, Dc = meta["L1"], meta["L2"], meta["Dc"]
    confs = np.array(z["traj_config"])
    nb = confs.shape[0]
    batch = Configuration(lat, Dc, nb)
    assert batch._ragged
    batch.import_configuration(confs)
    ws_b = np.asarray(batch.hole(()).storage).reshape(-1)
    obs_b = Observer(lat, enable_energy=True, enable_gradient=True)
    with obs_b:
        obs_b(ws_b**2, batch)
    holes_b = batch.holes()
    B = TAT.tensor._bk.get()
    e_sum = 0.0
    for c in range(nb):
        one = Configuration(lat, Dc)
        one.import_configuration(confs[c])
        ws = float(one.hole(()))
        assert abs(ws - ws_b[c]) <= 1e-10 * max(abs(ws), 1e-300)
        obs = Observer(lat, enable_energy=True, enable_gradient=True)
        with obs:
            obs(ws**2, one)
        e_sum += obs._whole_result_reweight["energy"]
        holes = one.holes()
        for l1 in range(L1):
            for l2 in range(L2):
                target = obs._Delta[l1][l2]
                w = np.asarray(holes[l1][l2].transpose(target.names).storage).reshape(-1)
                g = np.atleast_2d(B.to_numpy(_blocks_of(holes_b[l1][l2].transpose(target.names), target)))[c]
                assert np.abs(g - w).max() <= 1e-9 * max(np.abs(w).max(), 1e-300), (c, l1, l2)
    assert abs(obs_b._whole_result_reweight["energy"] - e_sum) <= 1e-9 * abs(e_sum)

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

"""Differential tests: tnsp_b200.TAT (product planner; CPU checker backend or CUDA backend) against the
UNMODIFIED reference PyTAT built into oracle/_ref, on identical random block-symmetric tensors for
all eight symmetry types.  Integer structure (names, edges, block order) must be identical;
values bit-exact for pure data movement, <= 1e-12 relative for GEMM, and gauge-invariant
quantities (U*S*V, Q*R, singular values, kept dimensions) for the factorizations.
Runs on CPU (default) and, with -m gpu, through the C-ABI on the B200.
"""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from helpers import FERMI, SYMS, conj_edge, describe, make_tensor, rand_edge, storage

pytestmark = []


def _pair(ref, sym, names, edges, rng):
    a = make_tensor(TAT, sym, names, edges)
    b = make_tensor(ref, sym, names, edges)
    n = len(storage(b))
    vals = rng.standard_normal(n)
    if n:
        a.storage = vals
        b.storage = vals
    da, db, na = describe(a, sym), describe(b, sym), len(storage(a))
    assert da == db
    assert na == n
    return a, b


def _same(a, b, sym, tol=0.0):
    da, db = describe(a, sym), describe(b, sym)
    assert da == db
    x, y = storage(a), storage(b)
    assert x.shape == y.shape
    if tol == 0.0:
        assert np.array_equal(x, y)
    else:
        scale = max(1.0, np.abs(y).max() if y.size else 1.0)
        assert np.abs(x - y).max() <= tol * scale if x.size else True


def _cases(n):
    return [(s, i) for s in SYMS for i in range(n)]


@pytest.mark.parametrize("sym,seed", _cases(6))
def test_block_layout_and_transpose(ref_tat, sym, seed):
    rng = np.random.default_rng(1000 + seed)
    rank = int(rng.integers(1, 5))
    names = [f"n{i}" for i in range(rank)]
    edges = [rand_edge(rng, sym) for _ in range(rank)]
    a, b = _pair(ref_tat, sym, names, edges, rng)
    perm = list(rng.permutation(rank))
    target = [names[i] for i in perm]
    _same(a.transpose(target), b.transpose(target), sym)


@pytest.mark.parametrize("sym,seed", _cases(6))
def test_merge_split_reverse(ref_tat, sym, seed):
    rng = np.random.default_rng(2000 + seed)
    rank = int(rng.integers(2, 5))
    names = [f"n{i}" for i in range(rank)]
    edges = [rand_edge(rng, sym) for _ in range(rank)]
    a, b = _pair(ref_tat, sym, names, edges, rng)
    # merge a random contiguous-free group of names
    k = int(rng.integers(1, rank + 1))
    group = [names[i] for i in rng.permutation(rank)[:k]]
    apply_parity = bool(rng.integers(0, 2))
    ma = a.merge_edge({"M": group}, apply_parity)
    mb = b.merge_edge({"M": group}, apply_parity)
    _same(ma, mb, sym)
    # split back
    sa = ma.split_edge({"M": [(n, a.edge_by_name(n).segments) for n in group]}, apply_parity)
    sb = mb.split_edge({"M": [(n, b.edge_by_name(n).segments) for n in group]}, apply_parity)
    _same(sa, sb, sym)
    if FERMI[sym]:
        rev = {names[i] for i in range(rank) if rng.integers(0, 2)}
        excl = {n for n in rev if rng.integers(0, 2)}
        _same(a.reverse_edge(rev, apply_parity, excl), b.reverse_edge(rev, apply_parity, excl), sym)
    _same(a.conjugate(), b.conjugate(), sym)
    _same(a.conjugate(True), b.conjugate(True), sym)


@pytest.mark.parametrize("sym,seed", _cases(10))
def test_contract(ref_tat, sym, seed):
    rng = np.random.default_rng(3000 + seed)
    r1, r2 = int(rng.integers(1, 5)), int(rng.integers(1, 5))
    nc = int(rng.integers(0, min(r1, r2) + 1))
    e1 = [rand_edge(rng, sym) for _ in range(r1)]
    e2 = [rand_edge(rng, sym) for _ in range(r2)]
    n1 = [f"a{i}" for i in range(r1)]
    n2 = [f"b{i}" for i in range(r2)]
    i1 = list(rng.permutation(r1)[:nc])
    i2 = list(rng.permutation(r2)[:nc])
    pairs = set()
    for x, y in zip(i1, i2):
        e2[y] = conj_edge(sym, e1[x])
        pairs.add((n1[x], n2[y]))
    a1, b1 = _pair(ref_tat, sym, n1, e1, rng)
    a2, b2 = _pair(ref_tat, sym, n2, e2, rng)
    _same(a1.contract(a2, pairs), b1.contract(b2, pairs), sym, tol=1e-12)


def _reconstruct_ok(orig, parts, pairs_chain, tol=1e-10):
    r = parts[0]
    for t, pr in zip(parts[1:], pairs_chain):
        r = r.contract(t, pr)
    r = r.transpose(list(orig.names))
    x, y = storage(r), storage(orig)
    assert np.abs(x - y).max() <= tol * max(1.0, np.abs(y).max()) if y.size else True


@pytest.mark.parametrize("sym,seed", _cases(8))
def test_qr(ref_tat, sym, seed):
    rng = np.random.default_rng(4000 + seed)
    rank = int(rng.integers(2, 5))
    names = [f"n{i}" for i in range(rank)]
    edges = [rand_edge(rng, sym) for _ in range(rank)]
    a, b = _pair(ref_tat, sym, names, edges, rng)
    k = int(rng.integers(1, rank))
    free = {names[i] for i in rng.permutation(rank)[:k]}
    direction = "r" if rng.integers(0, 2) else "q"
    qa, ra = a.qr(direction, free, "Q", "R")
    qb, rb = b.qr(direction, free, "Q", "R")
    d1, d2 = describe(qa, sym), describe(qb, sym)
    assert d1 == d2
    d1, d2 = describe(ra, sym), describe(rb, sym)
    assert d1 == d2
    _reconstruct_ok(a, [qa, ra], [{("Q", "R")}])
    # isometry of Q:  Q^dagger Q == identity on the common edge (compare with the reference's own value)
    ga = qa.conjugate().edge_rename({"Q": "Q2"}).contract(qa, {(n, n) for n in qa.names if n != "Q"})
    gb = qb.conjugate().edge_rename({"Q": "Q2"}).contract(qb, {(n, n) for n in qb.names if n != "Q"})
    _same(ga, gb, sym, tol=1e-10)


@pytest.mark.parametrize("sym,seed", _cases(8))
def test_svd(ref_tat, sym, seed):
    rng = np.random.default_rng(5000 + seed)
    rank = int(rng.integers(2, 5))
    names = [f"n{i}" for i in range(rank)]
    edges = [rand_edge(rng, sym, max_dim=4) for _ in range(rank)]
    a, b = _pair(ref_tat, sym, names, edges, rng)
    k = int(rng.integers(1, rank))
    free = {names[i] for i in rng.permutation(rank)[:k]}
    for cut in (-1, int(rng.integers(1, 6))):
        ua, sa, va = a.svd(free, "U", "V", "SU", "SV", cut)
        ub, sb, vb = b.svd(free, "U", "V", "SU", "SV", cut)
        d1, d2 = describe(ua, sym), describe(ub, sym)
        assert d1 == d2
        d1, d2 = describe(sa, sym), describe(sb, sym)
        assert d1 == d2
        d1, d2 = describe(va, sym), describe(vb, sym)
        assert d1 == d2
        x, y = storage(sa), storage(sb)
        assert np.abs(x - y).max() <= 1e-10 * max(1.0, np.abs(y).max()) if y.size else True
        ra = ua.contract(sa, {("U", "SU")}).contract(va, {("SV", "V")}).transpose(names)
        rb = ub.contract(sb, {("U", "SU")}).contract(vb, {("SV", "V")}).transpose(names)
        p, q = storage(ra), storage(rb)
        assert np.abs(p - q).max() <= 1e-9 * max(1.0, np.abs(q).max()) if q.size else True


@pytest.mark.parametrize("sym,seed", _cases(6))
def test_trace(ref_tat, sym, seed):
    """partial trace (bosonic symmetries) against the reference's trace.hpp"""
    rng = np.random.default_rng(7000 + seed)
    n_pairs = int(rng.integers(1, 3))
    n_free = int(rng.integers(0, 3))
    names, edges, pairs = [], [], set()
    for i in range(n_pairs):
        e = rand_edge(rng, sym)
        names += [f"a{i}", f"b{i}"]
        edges += [e, conj_edge(sym, e)]
        pairs.add((f"a{i}", f"b{i}"))
    for i in range(n_free):
        names.append(f"f{i}")
        edges.append(rand_edge(rng, sym))
    perm = list(rng.permutation(len(names)))
    names, edges = [names[i] for i in perm], [edges[i] for i in perm]
    a, b = _pair(ref_tat, sym, names, edges, rng)
    got = a.trace(pairs)
    if np.asarray(got.storage).size == 0:
        # no block of the result satisfies the symmetry: the reference's trace.hpp fails on this (bad optional access /
        # bad_alloc); it is never driven there
        pytest.skip("empty result: outside what the reference's trace supports")
    want = b.trace(pairs)
    _same(got, want, sym, tol=1e-12)


@pytest.mark.parametrize("sym,seed", _cases(5))
def test_identity(ref_tat, sym, seed):
    """identity_ between paired edges for all symmetry types, fermionic signs included (identity.hpp)"""
    rng = np.random.default_rng(8000 + seed)
    n_pairs = int(rng.integers(1, 3))
    names, edges, pairs = [], [], set()
    for i in range(n_pairs):
        e = rand_edge(rng, sym)
        names += [f"a{i}", f"b{i}"]
        edges += [e, conj_edge(sym, e)]
        pairs.add((f"a{i}", f"b{i}") if rng.integers(0, 2) else (f"b{i}", f"a{i}"))
    perm = list(rng.permutation(len(names)))
    names, edges = [names[i] for i in perm], [edges[i] for i in perm]
    a, b = _pair(ref_tat, sym, names, edges, rng)
    _same(a.identity_(pairs), b.identity_(pairs), sym)


@pytest.mark.parametrize("sym,seed", _cases(6))
def test_exponential(ref_tat, sym, seed):
    """tensor exponential over paired edges for all symmetry types (exponential.hpp: merge with the fermionic reverse / merge
    signs, per-sector Pade scaling-and-squaring, split back); values <= 1e-12 relative (the reference solves with LAPACK gesv)"""
    rng = np.random.default_rng(9000 + seed)
    n_pairs = int(rng.integers(1, 3))
    names, edges, pairs = [], [], set()
    for i in range(n_pairs):
        e = rand_edge(rng, sym, max_seg=3, max_dim=2)
        names += [f"a{i}", f"b{i}"]
        edges += [e, conj_edge(sym, e)]
        pairs.add((f"a{i}", f"b{i}") if rng.integers(0, 2) else (f"b{i}", f"a{i}"))
    perm = list(rng.permutation(len(names)))
    names, edges = [names[i] for i in perm], [edges[i] for i in perm]
    a, b = _pair(ref_tat, sym, names, edges, rng)
    for step in (8, 2):
        _same(a.exponential(pairs, step), b.exponential(pairs, step), sym, tol=1e-12)
    _same(a.exponential(pairs), b.exponential(pairs), sym, tol=1e-12)

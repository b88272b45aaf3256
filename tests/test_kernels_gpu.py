"""Kernel-level parity (-m gpu): each C-ABI entry point on the B200 against the numpy checker
(oracle/numpy_backend.py) on the same descriptor tables, including batched (nb > 1) launches,
broadcast operands, ragged sector shapes and sizes beyond one tile / shared memory."""
import numpy as np
import pytest
import torch

import tnsp_b200.TAT as TAT
from tnsp_b200 import backend
from tnsp_b200.TAT import plan as P

pytestmark = pytest.mark.gpu


def _both():
    from oracle.numpy_backend import NumpyBackend
    return backend.get(), NumpyBackend()


def _u1_edge(dims, arrow=False):
    E = TAT.BoseU1.Edge
    return E([(q, d) for q, d in zip(range(-(len(dims) // 2), len(dims)), dims)], arrow)


@pytest.mark.parametrize("nb", [1, 7, 300])
def test_pack_bit_exact(nb):
    cu, ck = _both()
    rng = np.random.default_rng(nb)
    E = TAT.BoseU1.Edge
    edges = (_u1_edge([3, 4, 2]), _u1_edge([2, 5, 3]), _u1_edge([4, 1, 3]), _u1_edge([2, 2, 2]))
    names = ("a", "b", "c", "d")
    p = P.edge_operator_plan(E, names, edges, None, None, {"M": ["c", "a"], "N": ["d", "b"]}, ["N", "M"])
    src = rng.standard_normal((nb, p.src_size))
    d_ck = ck.zeros(nb, p.dst_size)
    ck.pack(p, torch.from_numpy(src), d_ck)
    p._dev = None
    d_cu = cu.zeros(nb, p.dst_size)
    cu.pack(p, cu.from_numpy(src), d_cu)
    assert np.array_equal(cu.to_numpy(d_cu), d_ck.numpy())


@pytest.mark.parametrize("shape", [(5, 7, 3), (16, 16, 16), (64, 16, 4), (33, 65, 17), (130, 70, 200), (300, 257, 129), (1, 1, 46656), (2, 3, 5000),
                                   (40, 1, 9000), (700, 1, 40), (1, 300, 64)])
@pytest.mark.parametrize("flags", [0, 1, 2, 3])
def test_grouped_gemm(shape, flags):
    cu, ck = _both()
    m, n, k = shape
    rng = np.random.default_rng(m * n + flags)
    nb = 5

    class Plan:
        pass
    p = Plan()
    # two sectors of different shape in one launch
    m2, n2, k2 = max(1, m // 2), n + 3, max(1, k - 1)
    p.gemm = np.array([[m, n, k, 0, 0, 0, flags, 1], [m2, n2, k2, m * k, k * n, m * n, flags, -1]], dtype=np.int64)
    p._dev = None
    a = rng.standard_normal((nb, m * k + m2 * k2))
    b = rng.standard_normal((1, k * n + k2 * n2))  # broadcast operand
    c_ck = ck.zeros(nb, m * n + m2 * n2)
    ck.gemm(p, torch.from_numpy(a), torch.from_numpy(b), c_ck)
    c_cu = cu.zeros(nb, m * n + m2 * n2)
    cu.gemm(p, cu.from_numpy(a), cu.from_numpy(b), c_cu)
    ref = c_ck.numpy()
    # tolerance: float64 GEMM, different summation order -> 1e-13 relative to |A||B| scale
    assert np.abs(cu.to_numpy(c_cu) - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()) * k


def _factor_plan(shapes, flag):
    class Plan:
        pass
    p = Plan()
    rows, ao, o1, o2, so = [], 0, 0, 0, 0
    for m, n in shapes:
        k = min(m, n)
        rows.append((m, n, k, ao, o1, o2, so, 0))
        ao += m * n
        o1 += m * k
        o2 += k * n
        so += k
    p.sectors = np.array(rows, dtype=np.int64).reshape(-1, 8)
    p.s_total = so
    p.flag = flag
    p._dev = None
    return p, ao, o1, o2, so


@pytest.mark.parametrize("shapes", [[(16, 16), (4, 64), (64, 4)], [(64, 16), (16, 64), (1, 1), (5, 3)], [(216, 36)], [(40, 200), (180, 170)],
                                    [(900, 120), (120, 900), (300, 300), (70, 50), (9, 7)], [(2100, 260), (33, 1500)]])
@pytest.mark.parametrize("use_qr", [True, False])
def test_batched_qr(shapes, use_qr):
    cu, _ = _both()
    p, ao, o1, o2, _ = _factor_plan(shapes, use_qr)
    nb = 3
    rng = np.random.default_rng(len(shapes))
    a = rng.standard_normal((nb, ao))
    t1, t2 = cu.zeros(nb, o1), cu.zeros(nb, o2)
    cu.qr(p, cu.from_numpy(a), t1, t2)
    T1, T2 = cu.to_numpy(t1), cu.to_numpy(t2)
    for (m, n, k, a_off, x1, x2, _, _) in p.sectors:
        for b in range(nb):
            M = a[b, a_off:a_off + m * n].reshape(m, n)
            F1 = T1[b, x1:x1 + m * k].reshape(m, k)
            F2 = T2[b, x2:x2 + k * n].reshape(k, n)
            assert np.abs(F1 @ F2 - M).max() <= 1e-12 * max(1.0, np.abs(M).max()) * max(m, n)
            if use_qr:
                assert np.abs(F1.T @ F1 - np.eye(k)).max() <= 1e-12
                assert np.abs(np.tril(F2, -1)).max() == 0.0
            else:
                assert np.abs(F2 @ F2.T - np.eye(k)).max() <= 1e-12
                assert np.abs(np.triu(F1, 1)).max() == 0.0


@pytest.mark.parametrize("shapes", [[(16, 16), (4, 64), (64, 4)], [(64, 16), (16, 64), (1, 1), (5, 3)], [(36, 216)], [(150, 120), (60, 60)],
                                    [(900, 100), (100, 900), (130, 130), (70, 50), (9, 7)], [(1500, 210), (40, 1200)]])
def test_batched_svd_and_cut(shapes):
    cu, ck = _both()
    p, ao, o1, o2, so = _factor_plan(shapes, True)
    nb = 3
    rng = np.random.default_rng(7 + len(shapes))
    a = rng.standard_normal((nb, ao))
    t1, t2, s = cu.zeros(nb, o1), cu.zeros(nb, o2), cu.zeros(nb, so)
    cu.svd(p, cu.from_numpy(a), t1, s, t2)
    T1, T2, S = cu.to_numpy(t1), cu.to_numpy(t2), cu.to_numpy(s)
    for (m, n, k, a_off, x1, x2, xs, _) in p.sectors:
        for b in range(nb):
            M = a[b, a_off:a_off + m * n].reshape(m, n)
            U = T1[b, x1:x1 + m * k].reshape(m, k)
            Vt = T2[b, x2:x2 + k * n].reshape(k, n)
            sv = S[b, xs:xs + k]
            ref = np.linalg.svd(M, compute_uv=False)
            assert np.abs(sv - ref).max() <= 1e-12 * ref.max()          # singular values vs LAPACK
            assert np.abs((U * sv) @ Vt - M).max() <= 1e-12 * ref.max() * max(m, n)
            assert np.abs(U.T @ U - np.eye(k)).max() <= 1e-11
            assert np.abs(Vt @ Vt.T - np.eye(k)).max() <= 1e-11
    # greedy cut: device ranking == literal restatement of svd.hpp:429-481 (bit-exact integers)
    for cut, rel in ((5, 0.0), (1 << 40, 0.3), (17, 0.05)):
        c_cu = cu.to_numpy(cu.svd_cut(p, s, cut, rel))
        c_ck = ck.svd_cut(p, torch.from_numpy(S), cut, rel).numpy()
        assert np.array_equal(c_cu, c_ck)


def test_streaming_kernels():
    cu, ck = _both()
    rng = np.random.default_rng(3)
    nb, size = 9, 1000
    x = rng.standard_normal((nb, size))
    y = rng.standard_normal((nb, size))
    X, Y = cu.from_numpy(x), cu.from_numpy(y)
    for kind in (-1, 1, 2):
        assert np.allclose(cu.to_numpy(cu.norm(X, kind)), ck.norm(torch.from_numpy(x), kind).numpy(), rtol=1e-14, atol=0)
    al = rng.standard_normal(nb)
    for op in (0, 1):
        assert np.array_equal(cu.to_numpy(cu.scale(X, cu.from_numpy(al), op)), ck.scale(torch.from_numpy(x), torch.from_numpy(al), op).numpy())
    for op in range(4):
        assert np.array_equal(cu.to_numpy(cu.binary(X, Y, op)), ck.binary(torch.from_numpy(x), torch.from_numpy(y), op).numpy())
        assert np.array_equal(cu.to_numpy(cu.unary(X, op)), ck.unary(torch.from_numpy(x), op).numpy())
    idx = rng.integers(0, 4, size=nb).astype(np.int32)
    src = rng.standard_normal((1, 4 * 50))
    assert np.array_equal(cu.to_numpy(cu.gather_rows(cu.from_numpy(src), 50, cu.from_numpy(idx))),
                          ck.gather_rows(torch.from_numpy(src), 50, torch.from_numpy(idx)).numpy())
    mask = (rng.random(nb) < 0.5).astype(np.uint8)
    assert np.array_equal(cu.to_numpy(cu.select(cu.from_numpy(mask), X, Y)), ck.select(torch.from_numpy(mask), torch.from_numpy(x), torch.from_numpy(y)).numpy())
    w, e = rng.random(nb), rng.standard_normal(nb)
    d0, e0 = np.zeros((1, size)), np.zeros((1, size))
    D, ED = cu.from_numpy(d0), cu.from_numpy(e0)
    cu.grad_accumulate(X, cu.from_numpy(w), cu.from_numpy(e), D, ED)
    td, te = torch.from_numpy(d0.copy()), torch.from_numpy(e0.copy())
    ck.grad_accumulate(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(e), td, te)
    assert np.allclose(cu.to_numpy(D), td.numpy(), rtol=1e-13, atol=1e-13)
    assert np.allclose(cu.to_numpy(ED), te.numpy(), rtol=1e-13, atol=1e-13)


def _block_matrix(rng, m, n, n_sec, zero_rows=0, zero_cols=0):
    """m x n matrix that is block diagonal after chain-specific row / column permutations (what the
    charge-dense embedding of a symmetric tensor looks like), plus all-zero rows / columns"""
    M = np.zeros((m, n))
    rows = rng.permutation(m)[:m - zero_rows]
    cols = rng.permutation(n)[:n - zero_cols]
    rcut = np.sort(rng.choice(np.arange(1, len(rows)), n_sec - 1, replace=False)) if n_sec > 1 else []
    ccut = np.sort(rng.choice(np.arange(1, len(cols)), n_sec - 1, replace=False)) if n_sec > 1 else []
    ksum = 0
    for t, (rs, cs) in enumerate(zip(np.split(rows, rcut), np.split(cols, ccut))):
        blk = rng.standard_normal((len(rs), len(cs)))
        if t % 2 == 1:
            # staircase block (what triangular R factors / rank-limited bonds leave inside a sector): still connected
            blk = np.triu(blk)
            blk[0, :] = rng.standard_normal(len(cs))
        elif t % 3 == 2:
            # rank-1 sector (bond larger than the rank it can carry, as at the lattice boundary)
            blk = np.outer(rng.standard_normal(len(rs)), rng.standard_normal(len(cs)))
        M[np.ix_(rs, cs)] = blk
        ksum += min(len(rs), len(cs))
    return M, ksum


@pytest.fixture
def discovery():
    cu, _ = _both()
    cu.sector_discovery = True
    yield cu
    cu.sector_discovery = False


@pytest.mark.parametrize("shape", [(300, 120, 5, 3, 2), (120, 300, 4, 0, 5), (216, 216, 7, 0, 0), (700, 216, 1, 0, 0), (90, 90, 90, 0, 0),
                                   (1296, 216, 6, 0, 0), (6, 36, 3, 0, 4), (36, 6, 3, 2, 0), (6, 6, 2, 2, 2), (1, 5, 1, 0, 0)])
@pytest.mark.parametrize("use_qr", [True, False])
def test_sector_discovery_qr(shape, use_qr, discovery):
    """dense-embedded (block structured, per-chain permuted) matrices: sectors are found on the device"""
    m, n, n_sec, zr, zc = shape
    cu = discovery
    p, ao, o1, o2, _ = _factor_plan([(m, n)], use_qr)
    nb = 3
    rng = np.random.default_rng(m + n)
    mats = [_block_matrix(rng, m, n, n_sec, zr, zc) for _ in range(nb)]
    a = np.stack([M.reshape(-1) for M, _ in mats])
    t1, t2 = cu.zeros(nb, o1), cu.zeros(nb, o2)
    cu.qr(p, cu.from_numpy(a), t1, t2)
    T1, T2 = cu.to_numpy(t1), cu.to_numpy(t2)
    k = min(m, n)
    for b in range(nb):
        M, ksum = mats[b]
        F1, F2 = T1[b].reshape(m, k), T2[b].reshape(k, n)
        assert np.abs(F1 @ F2 - M).max() <= 1e-12 * max(m, n)
        G = F1.T @ F1 if use_qr else F2 @ F2.T          # isometry on the used bond indices, zero elsewhere
        d = np.diag(G)
        assert np.abs(G - np.diag(d)).max() <= 1e-12
        assert np.all((np.abs(d - 1) <= 1e-12) | (d == 0)) and int(round(d.sum())) <= ksum
        # the factors keep the sector structure: factor entries are non-zero only on rows / columns of one sector
        rows_used = (M != 0).any(axis=1)
        cols_used = (M != 0).any(axis=0)
        assert not F1[~rows_used].any() and not F2[:, ~cols_used].any()


@pytest.mark.parametrize("shape", [(216, 216, 7, 0, 0), (150, 260, 4, 3, 2), (260, 150, 5, 2, 0), (200, 180, 1, 0, 0), (64, 64, 64, 0, 0),
                                   (300, 300, 1, 0, 0), (6, 36, 3, 0, 4), (36, 6, 3, 2, 0), (6, 6, 2, 2, 2), (1, 5, 1, 0, 0)])
def test_sector_discovery_svd(shape, discovery):
    m, n, n_sec, zr, zc = shape
    cu = discovery
    p, ao, o1, o2, so = _factor_plan([(m, n)], True)
    nb = 3
    rng = np.random.default_rng(m * 3 + n)
    mats = [_block_matrix(rng, m, n, n_sec, zr, zc) for _ in range(nb)]
    a = np.stack([M.reshape(-1) for M, _ in mats])
    t1, t2, s = cu.zeros(nb, o1), cu.zeros(nb, o2), cu.zeros(nb, so)
    cu.svd(p, cu.from_numpy(a), t1, s, t2)
    T1, T2, S = cu.to_numpy(t1), cu.to_numpy(t2), cu.to_numpy(s)
    k = min(m, n)
    for b in range(nb):
        M, ksum = mats[b]
        U, Vt, sv = T1[b].reshape(m, k), T2[b].reshape(k, n), S[b]
        ref = np.linalg.svd(M, compute_uv=False)
        assert np.all(np.diff(sv) <= 0)
        assert np.abs(sv - ref).max() <= 1e-12 * ref.max()
        assert np.abs((U * sv) @ Vt - M).max() <= 1e-12 * ref.max() * max(m, n)
        rank = int((ref > 1e-10 * ref.max()).sum())
        G = U.T @ U
        assert np.abs(G[:rank, :rank] - np.eye(rank)).max() <= 1e-11
        G = Vt @ Vt.T
        assert np.abs(G[:rank, :rank] - np.eye(rank)).max() <= 1e-11


@pytest.mark.parametrize("queue_min", [0, 1 << 60])
@pytest.mark.parametrize("shape", [(216, 216, 7), (216, 1296, 7), (1296, 216, 6), (36, 216, 5), (512, 512, 3)])
def test_sector_paths_agree(shape, queue_min, discovery):
    """the per-sector work-queue path and the one-CTA-per-chain kernels factorise a whole batch of chains with
    different sector structures identically (same sector order, same bond indices): Q*R, U*S*V and the singular
    values against numpy, chain by chain"""
    m, n, n_sec = shape
    cu = discovery
    nb = 37
    rng = np.random.default_rng(m * 7 + n)
    mats = [_block_matrix(rng, m, n, n_sec, 0, 0)[0] for _ in range(nb)]
    a = np.stack([M.reshape(-1) for M in mats])
    k = min(m, n)
    old = cu.lib.tnsp_sector_queue_min(queue_min)
    try:
        for use_qr in (True, False):
            p, ao, o1, o2, so = _factor_plan([(m, n)], use_qr)
            t1, t2 = cu.zeros(nb, o1), cu.zeros(nb, o2)
            cu.qr(p, cu.from_numpy(a), t1, t2)
            F1, F2 = cu.to_numpy(t1).reshape(nb, m, k), cu.to_numpy(t2).reshape(nb, k, n)
            for b in range(nb):
                assert np.abs(F1[b] @ F2[b] - mats[b]).max() <= 1e-12 * max(m, n)
        p, ao, o1, o2, so = _factor_plan([(m, n)], True)
        t1, t2, s = cu.zeros(nb, o1), cu.zeros(nb, o2), cu.zeros(nb, so)
        cu.svd(p, cu.from_numpy(a), t1, s, t2)
        U, Vt, S = cu.to_numpy(t1).reshape(nb, m, k), cu.to_numpy(t2).reshape(nb, k, n), cu.to_numpy(s)
        for b in range(nb):
            ref = np.linalg.svd(mats[b], compute_uv=False)
            assert np.all(np.diff(S[b]) <= 0)
            assert np.abs(S[b] - ref).max() <= 1e-12 * ref.max()
            assert np.abs((U[b] * S[b]) @ Vt[b] - mats[b]).max() <= 1e-12 * ref.max() * max(m, n)
    finally:
        cu.lib.tnsp_sector_queue_min(old)


@pytest.mark.parametrize("dims", [
    # (dims of tensor 1, axes of tensor 1 that are contracted, dims of tensor 2, contracted axes of tensor 2 in pairing order)
    ((36, 6, 36), (1,), (6, 6, 6), (0,)),            # 1296 x 36 x 6: middle index contracted, k short of a slab
    ((6, 36, 6, 36), (1, 3), (36, 36, 5), (1, 0)),   # 36 x 5 x 1296 rows small -> still gathered if m >= 48? (m = 36: packed path)
    ((36, 36, 6), (2,), (6, 216), (0,)),             # 1296 x 216 x 6
    ((216, 6, 36), (0,), (216, 36), (0,)),           # A stored k-major: 216 x 36 x 216
    ((8, 27, 37), (1,), (3, 27, 11), (1,)),          # odd sizes, n = 33 not a multiple of 8, k = 27
    ((1300,), (), (7,), ()),                          # outer product, k = 1
    ((216, 36, 6), (1, 2), (6, 36, 216), (1, 0)),    # 216 x 216 x 216 with permuted common order
])
@pytest.mark.parametrize("nb", [1, 5, 300])
def test_gemm_gather_in_place_operands(dims, nb):
    """contract of dense tensors through the offset-table GEMM (no packed operands) against numpy tensordot;
    one operand broadcast to all chains in a second pass"""
    import tnsp_b200.TAT as TAT
    from tnsp_b200.TAT import tensor as tt
    cu, ck = _both()
    d1, c1, d2, c2 = dims
    rng = np.random.default_rng(sum(d1) + 7 * sum(d2) + nb)
    T = TAT.No.D.Tensor
    n1 = [f"a{i}" for i in range(len(d1))]
    n2 = [f"b{i}" for i in range(len(d2))]
    x1 = rng.standard_normal((nb, int(np.prod(d1))))
    for nb2 in (nb, 1):
        x2 = rng.standard_normal((nb2, int(np.prod(d2))))
        t1 = T.from_batch(n1, [TAT.No.Edge(d) for d in d1], x1)
        t2 = T.from_batch(n2, [TAT.No.Edge(d) for d in d2], x2)
        pairs = {(n1[i], n2[j]) for i, j in zip(c1, c2)}
        got = t1.contract(t2, pairs)
        want = np.stack([np.tensordot(x1[b].reshape(d1), x2[b % nb2].reshape(d2), axes=(list(c1), list(c2))).reshape(-1) for b in range(nb)])
        res = np.atleast_2d(np.asarray(got.storage))
        kk = int(np.prod([d1[i] for i in c1])) if c1 else 1
        assert res.shape == want.shape
        assert np.abs(res - want).max() <= 1e-13 * max(1.0, np.abs(want).max()) * max(kk, 1)
        # the packed path gives the same numbers
        cu.gather_gemm = False
        try:
            tt._PLAN_CACHE.clear()
            res2 = np.atleast_2d(np.asarray(t1.contract(t2, pairs).storage))
        finally:
            cu.gather_gemm = True
            tt._PLAN_CACHE.clear()
        assert np.abs(res2 - want).max() <= 1e-13 * max(1.0, np.abs(want).max()) * max(kk, 1)


@pytest.mark.parametrize("shape", [((12, 18, 9), (1,), 3), ((6, 36, 6, 6), (0, 2), 4), ((36, 36, 6, 6), (1, 3), 5)])
def test_factor_operand_read_in_place(shape, discovery):
    """qr / svd of a dense(-embedded) tensor whose row group is NOT leading: the sector kernels read the operand through
    the planner's offset table (no merged copy) and must give the factors of the merged matrix -- checked against numpy
    on the explicitly transposed data, and against the packed path"""
    import tnsp_b200.TAT as TAT
    from tnsp_b200.TAT import tensor as tt
    dims, row_axes, n_sec = shape
    cu = discovery
    nb = 5
    names = [f"x{i}" for i in range(len(dims))]
    col_axes = [i for i in range(len(dims)) if i not in row_axes]
    m = int(np.prod([dims[i] for i in row_axes])); n = int(np.prod([dims[i] for i in col_axes]))
    rng = np.random.default_rng(m + 3 * n)
    T = TAT.No.D.Tensor
    mats = [_block_matrix(rng, m, n, n_sec)[0] for _ in range(nb)]
    # store the m x n matrix as a tensor with the original index order
    perm = list(row_axes) + col_axes
    inv = np.argsort(perm)
    data = np.stack([M.reshape([dims[i] for i in perm]).transpose(inv).reshape(-1) for M in mats])
    t = T.from_batch(names, [TAT.No.Edge(d) for d in dims], data)
    free = {names[i] for i in row_axes}
    results = {}
    for mode in (True, False):
        cu.gather_gemm = mode          # in-place operands on / off
        tt._PLAN_CACHE.clear()
        try:
            q, r = t.qr("r", {names[i] for i in col_axes}, "Q", "R")
            u, s_, v = t.svd(free, "U", "V", "SU", "SV")
        finally:
            cu.gather_gemm = True
        rec_qr = np.atleast_2d(np.asarray(q.contract(r, {("Q", "R")}).transpose(names).storage))
        rec_svd = np.atleast_2d(np.asarray(u.contract(s_, {("U", "SU")}).contract(v, {("SV", "V")}).transpose(names).storage))
        sv = np.atleast_2d(np.asarray(s_.storage)).reshape(nb, -1)
        results[mode] = (rec_qr, rec_svd, sv)
        assert np.abs(rec_qr - data).max() <= 1e-12 * max(m, n)
        assert np.abs(rec_svd - data).max() <= 1e-12 * max(m, n)
        k = min(m, n)
        for b in range(nb):
            ref = np.linalg.svd(mats[b], compute_uv=False)
            got = np.sort(np.diag(sv[b].reshape(k, k)))[::-1]
            assert np.abs(got - ref).max() <= 1e-12 * ref.max()
    tt._PLAN_CACHE.clear()


@pytest.mark.parametrize("shapes", [[(400, 90), (90, 400), (64, 64)], [(150, 120), (60, 60)]])
def test_descriptor_kernels_agree_with_first_generation(shapes):
    """blocked Householder / QR-preconditioned Jacobi against the column-by-column kernels they replace, plus a rank-deficient
    sector (zero columns): same singular values, both reconstruct"""
    cu, _ = _both()
    p, ao, o1, o2, so = _factor_plan(shapes, True)
    nb = 2
    rng = np.random.default_rng(11)
    a = rng.standard_normal((nb, ao))
    m0, n0 = shapes[0]
    a[:, :m0 * n0].reshape(nb, m0, n0)[:, :, : n0 // 3] = 0.0          # a third of the first sector's columns vanish
    out = {}
    old = cu.lib.tnsp_factor_desc_kernels(-1)
    try:
        for gen in (0, 1):
            cu.lib.tnsp_factor_desc_kernels(gen)
            t1, t2, s = cu.zeros(nb, o1), cu.zeros(nb, o2), cu.zeros(nb, so)
            cu.svd(p, cu.from_numpy(a), t1, s, t2)
            q1, q2 = cu.zeros(nb, o1), cu.zeros(nb, o2)
            cu.qr(p, cu.from_numpy(a.copy()), q1, q2)
            out[gen] = [cu.to_numpy(x) for x in (t1, s, t2, q1, q2)]
    finally:
        cu.lib.tnsp_factor_desc_kernels(old)
    assert np.abs(out[0][1] - out[1][1]).max() <= 1e-12 * np.abs(out[0][1]).max()
    for gen in (0, 1):
        T1, S, T2, Q1, Q2 = out[gen]
        for (m, n, k, a_off, x1, x2, xs, _) in p.sectors:
            for b in range(nb):
                M = a[b, a_off:a_off + m * n].reshape(m, n)
                U, Vt, sv = T1[b, x1:x1 + m * k].reshape(m, k), T2[b, x2:x2 + k * n].reshape(k, n), S[b, xs:xs + k]
                assert np.abs((U * sv) @ Vt - M).max() <= 1e-12 * max(1.0, sv.max()) * max(m, n)
                assert np.all(np.diff(sv) <= 0)
                Q, R = Q1[b, x1:x1 + m * k].reshape(m, k), Q2[b, x2:x2 + k * n].reshape(k, n)
                assert np.abs(Q @ R - M).max() <= 1e-12 * max(1.0, np.abs(M).max()) * max(m, n)


@pytest.mark.parametrize("shape", [(216, 216), (36, 216), (150, 120)])
def test_jacobi_cached_norms_agree_with_three_dot_products(shape):
    """the sweeps that carry the column norms and the ones that recompute them per pair (default) give the same singular values
    and both reconstruct, on block-diagonal (discovered sectors) and dense matrices, including a graded spectrum"""
    cu, _ = _both()
    m, n = shape
    rng = np.random.default_rng(m + n)
    nb = 4
    mats = []
    for b in range(nb):
        M = rng.standard_normal((m, n))
        if b % 2 == 0:                       # block diagonal: three sectors
            mask = (np.arange(m)[:, None] % 3) == (np.arange(n)[None, :] % 3)
            M = M * mask
        if b >= 2:                           # graded columns: norms from 1 down to 1e-9
            M = M * np.logspace(0, -9, n)[None, :]
        mats.append(M.reshape(-1))
    a = np.stack(mats)
    p, ao, o1, o2, so = _factor_plan([shape], True)
    out = {}
    old = cu.lib.tnsp_jacobi_cached_norms(-1)
    saved = cu.sector_discovery
    try:
        cu.sector_discovery = True
        for mode in (0, 1):
            cu.lib.tnsp_jacobi_cached_norms(mode)
            t1, t2, s = cu.zeros(nb, o1), cu.zeros(nb, o2), cu.zeros(nb, so)
            cu.svd(p, cu.from_numpy(a), t1, s, t2)
            out[mode] = [cu.to_numpy(x) for x in (t1, s, t2)]
    finally:
        cu.lib.tnsp_jacobi_cached_norms(old)
        cu.sector_discovery = saved
    k = min(m, n)
    for b in range(nb):
        ref = np.linalg.svd(a[b].reshape(m, n), compute_uv=False)
        for mode in (0, 1):
            U, S, Vt = out[mode][0][b].reshape(m, k), out[mode][1][b], out[mode][2][b].reshape(k, n)
            sv = np.sort(S)[::-1]
            assert np.abs(sv - ref).max() <= 1e-12 * ref.max()
            assert np.abs((U * S) @ Vt - a[b].reshape(m, n)).max() <= 1e-12 * ref.max() * max(m, n)


@pytest.mark.parametrize("dims", [((1296, 216), (216, 216)), ((216, 216), (216, 36)), ((300, 40), (40, 36)), ((64, 7), (7, 100))])
def test_gemm_zero_fragment_skipping_is_exact(dims):
    """block-sparse operands (the zero pattern of charge conservation) through the row-stream GEMM with and without the
    zero-fragment test: identical results, and both equal to the dense product"""
    cu, _ = _both()
    (m, k), (_, n) = dims
    rng = np.random.default_rng(m + n + k)
    nb = 5
    qa, qk, qn = rng.integers(0, 5, m), rng.integers(0, 5, k), rng.integers(0, 5, n)
    a = rng.standard_normal((nb, m, k)) * (qa[:, None] == qk[None, :])
    b = rng.standard_normal((nb, k, n)) * (qk[:, None] == qn[None, :])
    a[0] = rng.standard_normal((m, k))          # one fully dense chain, one fully zero
    b[0] = rng.standard_normal((k, n))
    a[1] = 0.0
    T = TAT.No.D.Tensor
    t1 = T.from_batch(["i", "x"], [TAT.No.Edge(m), TAT.No.Edge(k)], cu.from_numpy(a.reshape(nb, -1)))
    t2 = T.from_batch(["x", "j"], [TAT.No.Edge(k), TAT.No.Edge(n)], cu.from_numpy(b.reshape(nb, -1)))
    out = {}
    old = cu.lib.tnsp_gemm_skip_zero_fragments(-1)
    try:
        for mode in range(5):                    # 0 dense, 1 .. 4 the skipping variants
            cu.lib.tnsp_gemm_skip_zero_fragments(mode)
            out[mode] = cu.to_numpy(t1.contract(t2, {("x", "x")}).data).reshape(nb, m, n)
    finally:
        cu.lib.tnsp_gemm_skip_zero_fragments(old)
    for mode in range(1, 5):
        assert np.array_equal(out[0], out[mode]), mode
    ref = np.matmul(a, b)
    assert np.abs(out[1] - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max()) * k

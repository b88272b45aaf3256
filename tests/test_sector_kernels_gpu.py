"""-m gpu: the sector-compact kernels (csrc/ragged.cu, rt_* part of csrc/factor_sector.cu) against their numpy specification
(oracle/numpy_ragged.py) on the same seeded inputs, and against brute-force dense arithmetic.  Integer tables bit-exact; data
<= 1e-12; factorisations through gauge-invariant quantities (Q R, U S V, singular values)."""
import numpy as np
import pytest

from tnsp_b200 import backend
from tnsp_b200.TAT import ragged
from test_sector_engine import (test_lockstep_sector_equals_one_by_one_evaluation, test_sector_engine_reproduces_reference_fixture,  # noqa: F401
                                test_sweep_trajectory_sector_equals_reference_trajectory, test_tensor_ops_equal_block_symmetric_ops)

pytestmark = pytest.mark.gpu


def _labels(rng, nb, dim, per_chain, lo=-1, hi=1, dead=0):
    lab = rng.integers(lo, hi + 1, size=(nb if per_chain else 1, dim)).astype(np.int32)
    for _ in range(dead):
        lab[rng.integers(0, lab.shape[0]), rng.integers(0, dim)] = ragged.DEAD
    return lab


def _case(seed, nb, da, db, dk, dead=0):
    """a[L.., K], b[K', N..] with random per-chain labels; returns plain numpy inputs"""
    rng = np.random.default_rng(seed)
    la = [_labels(rng, nb, d, i % 2 == 0) for i, d in enumerate(da)]
    lk = _labels(rng, nb, dk, True, dead=dead)
    lb = [_labels(rng, nb, d, i % 2 == 1) for i, d in enumerate(db)]
    ta = rng.integers(-1, 2, size=nb).astype(np.int32)
    tb = rng.integers(-1, 2, size=nb).astype(np.int32)
    a = rng.standard_normal((nb, int(np.prod(da)) * dk))
    b = rng.standard_normal((nb, dk * int(np.prod(db))))
    return dict(la=la, lk=lk, lb=lb, ta=ta, tb=tb, a=a, b=b, da=da, db=db, dk=dk, nb=nb)


def _run(B, case, cut):
    backend.set_backend(B)
    ragged._PLANS.clear()
    E = ragged.Edge
    na = [f"A{i}" for i in range(len(case["da"]))]
    nbn = [f"B{i}" for i in range(len(case["db"]))]
    ea = [E(d, B.from_numpy(l), 1) for d, l in zip(case["da"], case["la"])] + [E(case["dk"], B.from_numpy(case["lk"]), 1)]
    eb = [E(case["dk"], B.from_numpy(case["lk"]), -1)] + [E(d, B.from_numpy(l), 1) for d, l in zip(case["db"], case["lb"])]
    a = ragged.RTensor.from_dense(na + ["K"], ea, case["a"], case["ta"])
    b = ragged.RTensor.from_dense(["K2"] + nbn, eb, case["b"], case["tb"])
    out = {}
    num = lambda t: np.asarray(B.to_numpy(t))  # noqa: E731
    out["a"], out["b"] = num(a.to_dense()), num(b.to_dense())
    pa = a._primary()
    out["tab"] = num(pa.rt)
    out["match"] = num(pa.match)[:, :3 + 2 * ragged.SMAX]
    c = a.contract(b, {("K", "K2")})
    out["c"] = num(c.to_dense())
    q, r = c.qr("r", set(nbn), "X", "Y")
    out["qr"] = num(q.contract(r, {("X", "Y")}).to_dense())
    qq = q.conjugate().edge_rename({"X": "X2"}).contract(q, {(n, n) for n in na})
    out["qq"] = num(qq.to_dense())
    u, s, v = c.svd(set(na[:1]), "U", "V", "SU", "SV", cut)
    out["usv"] = num(u.contract(s, {("U", "SU")}).contract(v, {("SV", "V")}).transpose(c.names).to_dense())
    sd = num(s.to_dense()).reshape(case["nb"], s.core.edges[0].dim, -1)
    out["sv"] = np.sort(np.diagonal(sd, axis1=1, axis2=2), axis=1)[:, ::-1]
    out["dot"] = np.asarray(c.conjugate().contract(c, {(n, n) for n in c.names}).storage)      # full contraction (rt_dot)
    out["norm"] = np.asarray(c.norm_2().numpy())
    out["nmax"] = np.asarray(c.norm_max().numpy())
    out["scaled"] = num((c / c.norm_max()).to_dense())
    out["sum"] = num((c + c * 0.5).to_dense())
    out["sizes"] = (case["da"], case["db"])
    return out


@pytest.mark.parametrize("shape", [((5, 3), (4,), 6, 0, 3), ((7, 2, 3), (2, 5), 6, 2, 4), ((40, 3), (90,), 70, 3, 20), ((130,), (9, 11), 37, 0, 50),
                                   ((60, 40), (4,), 5, 1, 3)])      # merged group of 2400 indices: the CTA-wide rt_sort (smaller groups: one warp per chain)
def test_kernels_equal_specification(shape):
    from oracle.numpy_backend import NumpyBackend
    da, db, dk, dead, cut = shape
    case = _case(hash(shape) % 1000, 6, da, db, dk, dead)
    cu = backend.get()
    try:
        got = _run(cu, case, cut)
        want = _run(NumpyBackend(), case, cut)
    finally:
        backend.set_backend(cu)
        ragged._PLANS.clear()
    S, H = ragged.SMAX, ragged.HDR
    for g, w, gm, wm in zip(got["tab"], want["tab"], got["match"], want["match"]):     # only the defined entries of the tables
        n = int(w[0])
        assert g[0] == w[0] and g[1] == w[1]
        assert np.array_equal(g[2:2 + n], w[2:2 + n]) and np.array_equal(g[2 + S:3 + S + n], w[2 + S:3 + S + n])
        assert np.array_equal(g[H:], w[H:])
        assert gm[0] == wm[0] and np.array_equal(gm[2:3 + n], wm[2:3 + n]) and np.array_equal(gm[3 + S:3 + S + n], wm[3 + S:3 + S + n])
    for k in ("a", "b"):
        assert np.array_equal(got[k], want[k]), k
    nb = case["nb"]
    # brute force: dense contraction of the projected operands
    A = want["a"].reshape(nb, -1, dk)
    Bm = want["b"].reshape(nb, dk, -1)
    brute = np.einsum("cik,ckj->cij", A, Bm).reshape(nb, -1)
    scale = max(np.abs(brute).max(), 1.0)
    assert np.abs(got["c"] - brute).max() <= 1e-12 * scale
    assert np.abs(want["c"] - brute).max() <= 1e-12 * scale
    assert np.abs(got["qr"] - got["c"]).max() <= 1e-11 * scale
    assert np.abs(got["qq"] - np.round(got["qq"])).max() <= 1e-11          # Q^T Q = 1 on the live bond indices
    assert np.abs(got["qq"] - want["qq"]).max() <= 1e-11
    assert np.abs(got["sv"] - want["sv"]).max() <= 1e-11 * scale
    assert np.abs(got["usv"] - want["usv"]).max() <= 1e-9 * scale
    assert np.allclose(got["dot"], got["norm"]**2, rtol=1e-12, atol=0)
    for k in ("norm", "nmax", "dot"):
        assert np.allclose(got[k], want[k], rtol=1e-13, atol=0)
    for k in ("scaled", "sum"):
        assert np.abs(got[k] - want[k]).max() <= 1e-12 * scale


@pytest.mark.parametrize("env", [{"TNSP_RT_GEMM": "24"}, {"TNSP_RT_GEMM": "38"}, {"TNSP_RT_GEMM": "25", "TNSP_RT_REPACK_THREADS": "256"},
                                 {"TNSP_RT_GEMM": "12", "TNSP_RT_JACOBI": "3", "TNSP_RT_REPACK_THREADS": "64"}])
def test_kernel_variants_equal_specification(env):
    """the alternative kernels kept beside the defaults (TMA-fed GEMM with / without the cross-item prefetch, 2-warp GEMM CTAs, the
    pair-parallel Jacobi rotation phase, other regrouping CTA sizes) are selected per process by environment variables: the
    specification test above, re-run in a child process with them set"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_sector_kernels_gpu.py"), "-m", "gpu", "-q", "-x", "-k",
                        "test_kernels_equal_specification"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]

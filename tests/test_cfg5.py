"""cfg5 (BASELINE.json): block-symmetric contract / qr / svd on random U(1) rank-5 tensors, a batch of samples per call.
Small bond dimension against the UNMODIFIED reference PyTAT sample by sample (structure identical, contraction <= 1e-12,
gauge-invariant quantities of the factorisations <= 1e-10); the GPU run adds the full microbenchmark shape (Dc = 64) through
size-independent properties: Q R = T, Q^T Q = 1, U S V = T without truncation, singular values sorted per sector and the
truncated reconstruction error equal to the discarded weight."""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from helpers import describe, storage
from tnsp_b200 import cfg5


def _ref_tensor(ref, names, edges_, values):
    t = ref.BoseU1.D.Tensor(names, edges_)
    t.storage = values
    return t


@pytest.mark.parametrize("Dc,nb", [(8, 1), (8, 3), (16, 4)])
def test_cfg5_matches_reference_per_sample(ref_tat, Dc, nb):
    a1, a2, v1, v2 = cfg5.random_batch(TAT, Dc, nb)
    T, (Q, R), (U, S, V) = cfg5.steps(a1, a2, Dc)
    (n1, e1), (n2, e2) = cfg5.structures(ref_tat, Dc)
    got = {k: np.atleast_2d(np.asarray(t.storage)) for k, t in dict(T=T, S=S).items()}
    QR = np.atleast_2d(np.asarray(Q.contract(R, {("R", "L")}).transpose(T.names).storage))
    USV = np.atleast_2d(np.asarray(U.contract(S, {("R", "L")}).contract(V, {("R", "L")}).transpose(T.names).storage))
    for b in range(nb):
        rT, (rQ, rR), (rU, rS, rV) = cfg5.steps(_ref_tensor(ref_tat, n1, e1, v1[b]), _ref_tensor(ref_tat, n2, e2, v2[b]), Dc)
        if b == 0:
            # a batch shares ONE block structure: after a truncating svd the common edge holds, per sector, the largest
            # kept count of any sample and the other samples' surplus values are exact zeros, so the
            # U / S / V structures equal the reference's only for a single sample
            parts = [("T", T, rT), ("Q", Q, rQ), ("R", R, rR)] + ([("U", U, rU), ("S", S, rS), ("V", V, rV)] if nb == 1 else [])
            for what, mine, theirs in parts:
                d_mine, d_ref = describe(mine, "BoseU1"), describe(theirs, "BoseU1")   # never let pytest repr() a reference tensor
                assert d_mine == d_ref, what
        t = storage(rT)
        scale = np.abs(t).max()
        assert np.abs(got["T"][b] - t).max() <= 1e-12 * scale
        assert np.abs(QR[b] - t).max() <= 1e-10 * scale
        sv_ref = np.sort(np.abs(storage(rS)[storage(rS) != 0]))[::-1]
        sv = np.sort(np.abs(got["S"][b][got["S"][b] != 0]))[::-1]
        assert sv.shape == sv_ref.shape and np.abs(sv - sv_ref).max() <= 1e-10 * sv_ref.max()   # same Dc values kept
        ref_usv = storage(rU.contract(rS, {("R", "L")}).contract(rV, {("R", "L")}).transpose(rT.names))
        assert np.abs(USV[b] - ref_usv).max() <= 1e-9 * scale


def _properties(Dc, nb):
    a1, a2, _, _ = cfg5.random_batch(TAT, Dc, nb)
    T, (Q, R), (U, S, V) = cfg5.steps(a1, a2, Dc)
    t = np.atleast_2d(np.asarray(T.storage))
    scale = np.abs(t).max()
    qr = np.atleast_2d(np.asarray(Q.contract(R, {("R", "L")}).transpose(T.names).storage))
    assert np.abs(qr - t).max() <= 1e-10 * scale
    # Q^T Q = identity on the common edge
    g = Q.conjugate().edge_rename({"R": "R_"}).contract(Q, {(n, n) for n in Q.names if n != "R"})
    eye = g.same_shape().identity_({("R_", "R")}) if hasattr(g, "identity_") else None
    if eye is not None:
        assert np.abs(np.atleast_2d(np.asarray(g.storage)) - np.atleast_2d(np.asarray(eye.storage))).max() <= 1e-10
    # untruncated SVD reconstructs, truncated one loses exactly the discarded singular values (Frobenius norm)
    Uf, Sf, Vf = T.svd({"L1", "L2"}, "R", "L", "L", "R", -1)
    full = np.atleast_2d(np.asarray(Uf.contract(Sf, {("R", "L")}).contract(Vf, {("R", "L")}).transpose(T.names).storage))
    assert np.abs(full - t).max() <= 1e-10 * scale
    cutv = np.atleast_2d(np.asarray(U.contract(S, {("R", "L")}).contract(V, {("R", "L")}).transpose(T.names).storage))
    s_all = np.atleast_2d(np.asarray(Sf.storage))
    s_kept = np.atleast_2d(np.asarray(S.storage))
    kept = Dc
    assert S.edge_by_name("L").dimension >= Dc   # union over the batch of the per-sample kept sectors
    for b in range(nb):
        sv_all = np.sort(np.abs(s_all[b][s_all[b] != 0]))[::-1]
        sv_kept = np.sort(np.abs(s_kept[b][s_kept[b] != 0]))[::-1]
        assert np.allclose(sv_kept, sv_all[:kept], rtol=1e-10)          # the greedy cut keeps the globally largest values
        lost = np.sqrt((sv_all[kept:]**2).sum())
        err = np.sqrt(((cutv[b] - t[b])**2).sum())
        assert abs(err - lost) <= 1e-8 * max(lost, 1.0)


def test_cfg5_properties_small():
    _properties(16, 3)


@pytest.mark.gpu
def test_cfg5_properties_microbench_shape():
    _properties(64, 16)


@pytest.mark.gpu
@pytest.mark.parametrize("Dc,nb", [(8, 1), (8, 3), (16, 4)])
def test_cfg5_matches_reference_per_sample_gpu(ref_tat, Dc, nb):
    test_cfg5_matches_reference_per_sample(ref_tat, Dc, nb)

"""Row-stream GEMM on block-sparse operands shaped like the charge-dense embedding of cfg2 (zero pattern of U(1) charge
conservation, natural index order, a different charge assignment per chain): times the zero-fragment skipping variants
(tnsp_gemm_skip_zero_fragments 0 .. 4) against each other, CUDA-event timed, results compared bit for bit.
    python scripts/mb_gemm_sparse.py [nb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnsp_b200 import backend
import tnsp_b200.TAT as TAT
from tnsp_b200.TAT import tensor as tt

QD = np.array([-1, -1, 0, 0, 1, 1])


def operands(rng, nb, X, n_cols, variety=8):
    """A [(x, d1, d2), (dc, d)] (36*X/36 ... ) and B [(dc, d), j]; `variety` distinct charge assignments cycled over the chains"""
    As, Bs = [], []
    for _ in range(variety):
        q_dc = rng.choice([-3, -2, -1, 0, 1, 2, 3], size=36)
        q_x = rng.choice([-3, -2, -1, 0, 1, 2, 3], size=X)
        qk = (q_dc[:, None] + QD[None, :]).reshape(-1)                          # 216
        qr = (q_x[:, None, None] + QD[None, :, None] + QD[None, None, :]).reshape(-1)
        qn = np.sort(rng.choice(qk, size=n_cols))                               # sector-sorted bond out of a QR
        A = rng.standard_normal((qr.size, qk.size)) * (qr[:, None] == qk[None, :])
        B = rng.standard_normal((qk.size, n_cols)) * (qk[:, None] == qn[None, :])
        # B is the R factor of a QR: upper triangular inside every sector (row t of the factor only reaches columns >= t)
        for q in np.unique(qn):
            ks, js = np.nonzero(qk == q)[0], np.nonzero(qn == q)[0]
            tri = np.arange(len(ks))[:, None] >= np.arange(len(js))[None, :]
            B[np.ix_(ks, js)] *= tri
        As.append(A); Bs.append(B)
    idx = np.arange(nb) % variety
    return np.stack(As)[idx], np.stack(Bs)[idx], As[0], Bs[0]


def density(A, B):
    m, k = A.shape; n = B.shape[1]
    kp, npad, mp = (k + 3) // 4 * 4, (n + 7) // 8 * 8, (m + 7) // 8 * 8
    Bp = np.zeros((kp, npad)); Bp[:k, :n] = B
    Ap = np.zeros((mp, kp)); Ap[:m, :k] = A
    fb = (Bp.reshape(kp // 4, 4, npad // 8, 8) != 0).any(axis=(1, 3))
    fa = (Ap.reshape(mp // 8, 8, kp // 4, 4) != 0).any(axis=(1, 3))
    return fb.mean(), fa.mean(), (fa[:, :, None] & fb[None, :, :]).mean()


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = backend.get()
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    rng = np.random.default_rng(0)
    T = TAT.No.D.Tensor
    for X, n_cols in ((36, 216), (6, 216), (36, 36)):
        a, b, A0, B0 = operands(rng, nb, X, n_cols)
        m, k = A0.shape
        print(f"{m}x{n_cols}x{k} nb={nb}: fragment density B %.2f A %.2f both %.2f" % density(A0, B0), flush=True)
        t1 = T.from_batch(["i", "x"], [TAT.No.Edge(m), TAT.No.Edge(k)], B.from_numpy(a.reshape(nb, -1)))
        t2 = T.from_batch(["x", "j"], [TAT.No.Edge(k), TAT.No.Edge(n_cols)], B.from_numpy(b.reshape(nb, -1)))
        ref = None
        old = B.lib.tnsp_gemm_skip_zero_fragments(-1)
        for mode in range(5):
            B.lib.tnsp_gemm_skip_zero_fragments(mode)
            ms = timeit(lambda: t1.contract(t2, {("x", "x")}))
            out = t1.contract(t2, {("x", "x")}).data
            if ref is None: ref = out.clone()
            same = bool(torch.equal(out, ref))
            print(f"   mode {mode}: {ms:8.3f} ms   identical to mode 0: {same}", flush=True)
        B.lib.tnsp_gemm_skip_zero_fragments(old)


if __name__ == "__main__":
    main()

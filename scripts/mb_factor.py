"""Micro-benchmark of the discovered-sector QR / SVD kernels and the grouped GEMM on cfg2-shaped inputs
(block-structured, per-chain permuted matrices).  CUDA-event timing on the launching stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnsp_b200 import backend

from mb_factor_lib import block_matrix, plan_of

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    B = backend.get(); B.sector_discovery = True
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    rng = np.random.default_rng(0)
    base = {}
    for (m, n, ns) in [(216, 1296, 7), (216, 216, 6), (36, 216, 5), (1296, 216, 7)]:
        mats = np.stack([block_matrix(rng, m, n, ns, even=True).reshape(-1) for _ in range(8)])
        a = B.from_numpy(mats[np.arange(nb) % 8])
        for flag in (False, True):
            p, k = plan_of(m, n, flag)
            t1, t2 = B.zeros(nb, m * k), B.zeros(nb, k * n)
            ms = timeit(lambda: B.qr(p, a.clone(), t1, t2))
            msc = timeit(lambda: a.clone())
            print(f"qr  {m}x{n} sectors {ns} use_qr={flag} nb={nb}: {ms - msc:.3f} ms  ({8*(2*m*n+m*k+k*n)*nb/(ms-msc)/1e6:.0f} GB/s)")
        p, k = plan_of(m, n, True)
        t1, t2, s = B.zeros(nb, m * k), B.zeros(nb, k * n), B.zeros(nb, k)
        ms = timeit(lambda: B.svd(p, a, t1, s, t2))
        print(f"svd {m}x{n} sectors {ns} nb={nb}: {ms:.3f} ms  ({8*(m*n+m*k+k*n+k)*nb/ms/1e6:.0f} GB/s)")
    class G: pass
    for (m, n, k) in [(1296, 36, 36), (216, 36, 36), (1296, 36, 216), (7776, 36, 36), (1296, 216, 216), (1296, 216, 6), (216, 216, 6), (46656, 6, 6)]:
        g = G(); g.gemm = np.array([[m, n, k, 0, 0, 0, 0, 1]], dtype=np.int64); g._dev = None
        a = B.from_numpy(rng.standard_normal((nb, m * k))); b = B.from_numpy(rng.standard_normal((nb, k * n))); c = B.zeros(nb, m * n)
        ms = timeit(lambda: B.gemm(g, a, b, c))
        by = 8 * (m * k + k * n + m * n) * nb
        print(f"gemm {m}x{n}x{k} nb={nb}: {ms:.3f} ms  {2*m*n*k*nb/ms/1e9:.2f} TF/s  {by/ms/1e6:.0f} GB/s")

main()

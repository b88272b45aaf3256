"""Micro-benchmark of the contract kernels on the cfg2 shapes: packed operands + gemm_stream (old) against the
in-place offset-table GEMM (gemm_gather), CUDA-event timed.  Cases with a permuted operand include the pack.
    python scripts/mb_gemm.py [nb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnsp_b200 import backend
import tnsp_b200.TAT as TAT
from tnsp_b200.TAT import tensor as tt

CASES = [
    # name, dims1, contracted axes 1, dims2, contracted axes 2
    ("1296x36x36 plain", (1296, 36), (1,), (36, 36), (0,)),
    ("1296x36x36 mid-index", (36, 36, 36), (1,), (36, 36), (0,)),
    ("7776x36x36 plain", (7776, 36), (1,), (36, 36), (0,)),
    ("1296x36x216 plain", (1296, 216), (1,), (216, 36), (0,)),
    ("1296x216x216 plain", (1296, 216), (1,), (216, 216), (0,)),
    ("1296x216x216 permuted", (36, 6, 36, 36), (1, 3), (36, 6, 216), (1, 0)),
    ("1296x216x6 plain", (1296, 6), (1,), (6, 216), (0,)),
    ("216x36x36 plain", (216, 36), (1,), (36, 36), (0,)),
    ("216x216x216 plain", (216, 216), (1,), (216, 216), (0,)),
    ("216x216x36 plain", (216, 36), (1,), (36, 216), (0,)),
    ("216x36x216 plain", (216, 216), (1,), (216, 36), (0,)),
    ("1296x2x1 outer", (1296, 1), (1,), (1, 2), (0,)),
    ("1296x36x6 plain", (1296, 6), (1,), (6, 36), (0,)),
    ("36x36x216 plain", (36, 216), (1,), (216, 36), (0,)),
    ("36x216x6 plain", (36, 6), (1,), (6, 216), (0,)),
    ("36x36x36 plain", (36, 36), (1,), (36, 36), (0,)),
    ("36x6x216 k-major", (216, 36), (0,), (216, 6), (0,)),
    ("1x1x7776 dot", (7776,), (0,), (7776,), (0,)),
]


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
    for name, d1, c1, d2, c2 in CASES:
        n1 = [f"a{i}" for i in range(len(d1))]; n2 = [f"b{i}" for i in range(len(d2))]
        t1 = T.from_batch(n1, [TAT.No.Edge(d) for d in d1], B.from_numpy(rng.standard_normal((nb, int(np.prod(d1))))))
        t2 = T.from_batch(n2, [TAT.No.Edge(d) for d in d2], B.from_numpy(rng.standard_normal((nb, int(np.prod(d2))))))
        pairs = {(n1[i], n2[j]) for i, j in zip(c1, c2)}
        k = int(np.prod([d1[i] for i in c1])); m = int(np.prod(d1)) // k; n = int(np.prod(d2)) // k
        out = {}
        for mode in (False, True):
            B.gather_gemm = mode
            tt._PLAN_CACHE.clear()
            ms = timeit(lambda: t1.contract(t2, pairs))
            out[mode] = (ms, t1.contract(t2, pairs).data.clone())
        B.gather_gemm = True
        by = 8 * (m * k + k * n + m * n) * nb
        fl = 2 * m * n * k * nb
        err = float((out[True][1] - out[False][1]).abs().max())
        print(f"{name:24s} nb={nb}: packed {out[False][0]:.3f} ms ({fl/out[False][0]/1e9:.1f} TF, {by/out[False][0]/1e6:.0f} GB/s) | "
              f"gather {out[True][0]:.3f} ms ({fl/out[True][0]/1e9:.1f} TF, {by/out[True][0]/1e6:.0f} GB/s)  diff {err:.1e}", flush=True)

if __name__ == "__main__":
    main()

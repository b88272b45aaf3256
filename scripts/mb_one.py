"""one kernel launch of a chosen kind for ncu:  python scripts/mb_one.py qr|svd|gemm m n [k|nsec] nb"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnsp_b200 import backend
from mb_factor_lib import block_matrix, plan_of
kind, m, n, x, nb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
B = backend.get(); B.sector_discovery = True
rng = np.random.default_rng(0)
if kind in ("qr", "lq", "svd"):
    mats = np.stack([block_matrix(rng, m, n, x, even=True).reshape(-1) for _ in range(8)])
    a = B.from_numpy(mats[np.arange(nb) % 8])
    p, k = plan_of(m, n, kind != "lq")
    t1, t2, s = B.zeros(nb, m * k), B.zeros(nb, k * n), B.zeros(nb, k)
    for _ in range(3):
        if kind == "svd": B.svd(p, a, t1, s, t2)
        else: B.qr(p, a.clone(), t1, t2)
else:
    class G: pass
    g = G(); g.gemm = np.array([[m, n, x, 0, 0, 0, 0, 1]], dtype=np.int64); g._dev = None
    a = B.from_numpy(rng.standard_normal((nb, m * x))); b = B.from_numpy(rng.standard_normal((nb, x * n))); c = B.zeros(nb, m * n)
    for _ in range(3): B.gemm(g, a, b, c)
torch.cuda.synchronize()

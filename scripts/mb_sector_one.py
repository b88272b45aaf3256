"""a few launches of one discovered-sector factorisation for ncu:  python scripts/mb_sector_one.py svd|lq|qr m n nb"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnsp_b200 import backend
from mb_factor_lib import plan_of
from mb_sector import PATTERNS, matrix
kind, m, n, nb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
B = backend.get(); B.sector_discovery = True
rng = np.random.default_rng(0)
pats = PATTERNS[(kind, m, n)]
mats = np.stack([matrix(rng, m, n, pats[i % len(pats)]).reshape(-1) for i in range(8)])
a = B.from_numpy(mats[np.arange(nb) % 8])
p, k = plan_of(m, n, kind == "qr")
t1, t2, s = B.zeros(nb, m * k), B.zeros(nb, k * n), B.zeros(nb, k)
for _ in range(3):
    if kind == "svd": B.svd(p, a, t1, s, t2)
    else: B.qr(p, a, t1, t2)
torch.cuda.synchronize()

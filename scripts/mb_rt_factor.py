"""Microbenchmark of the sector-compact factorisations at cfg2's hot shapes (rt_factor: plan + qr / svd work kernels):
    python scripts/mb_rt_factor.py [nb] [reps]
  svd 216 x 216  (Dc*D x Dc*D, ~6 sectors up to 56 x 56 per chain)   -- the R-factor product of two_line_to_one_line
  qr 1296 x 216  (Dc*D*D x Dc*D, ~7 sectors up to 268 x 56)           -- the QR sweep of two_line_to_one_line
Per-chain labels: the bond of dimension 36 carries the charges -2..2 with multiplicities 4, 8, 12, 8, 4 in a per-chain random order
(what a greedy cut leaves), the PEPS bonds -1, 0, 1 twice each.  CUDA-event times per call, work kernels timed through the backend."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tnsp_b200 import backend
from tnsp_b200.TAT import ragged

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = backend.get()
ragged.CAPS_ENABLED = False
rng = np.random.default_rng(1)
base = np.repeat(np.array([-2, -1, 0, 1, 2]), [4, 8, 12, 8, 4]).astype(np.int32)


def bond_labels():
    return np.stack([rng.permutation(base) for _ in range(nb)])


def peps_labels():
    return np.array([[-1, -1, 0, 0, 1, 1]], dtype=np.int32)


def tensor(dims, per_chain):
    E = ragged.Edge
    edges = [E(d, B.from_numpy(bond_labels() if pc else peps_labels()), 1) for d, pc in zip(dims, per_chain)]
    size = int(np.prod(dims))
    dense = torch.randn((nb, size), dtype=torch.float64, device="cuda")
    return ragged.RTensor.from_dense([f"E{i}" for i in range(len(dims))], edges, dense)


def timed(fn, label):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / reps:.3f} ms per call ({nb} chains)")


t = tensor((36, 6, 36, 6), (True, False, True, False))
timed(lambda: t.svd({"E0", "E1"}, "U", "V", "SU", "SV", 36), "svd 216 x 216 cut 36")
del t
t = tensor((36, 6, 6, 36, 6), (True, False, False, True, False))
timed(lambda: t.qr("r", {"E3", "E4"}, "Q", "R"), "qr 1296 x 216")
del t
t = tensor((36, 6, 36, 6), (True, False, True, False))
timed(lambda: t.qr("r", {"E2", "E3"}, "Q", "R"), "qr 216 x 216")

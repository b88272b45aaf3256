"""Micro-benchmark of the discovered-sector QR / SVD entry points on cfg2-shaped inputs: per-chain permuted block
matrices with the sector sizes traced from the 6x6 U(1) J1-J2 workload.  Times the one-CTA-per-chain kernels against
the per-sector work-queue path (tnsp_sector_queue_min) with CUDA events on the launching stream.
    python scripts/mb_sector.py [nb]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnsp_b200 import backend
from mb_factor_lib import plan_of

PATTERNS = {
    ("svd", 216, 216): [[(40, 56), (48, 48), (24, 24), (8, 8), (14, 48), (56, 24), (26, 8)], [(24, 24), (40, 48), (40, 56), (24, 48), (8, 24), (8, 8), (56, 8), (16, 0)],
                        [(52, 56), (48, 48), (24, 24), (26, 48), (8, 8), (8, 24), (50, 8)]],
    ("svd", 36, 216): [[(8, 48), (12, 56), (8, 48), (4, 24), (4, 24), (0, 16)], [(8, 34), (12, 62), (8, 64), (4, 38), (4, 8), (0, 10)]],
    ("svd", 216, 36): [[(48, 8), (56, 12), (48, 8), (24, 4), (24, 4), (16, 0)]],
    ("lq", 216, 1296): [[(24, 28), (48, 108), (56, 204), (48, 224), (24, 160), (8, 64), (8, 508)], [(56, 208), (48, 144), (24, 64), (48, 208), (24, 144), (8, 16), (8, 60), (0, 452)],
                        [(48, 172), (56, 252), (48, 248), (24, 68), (24, 160), (8, 16), (8, 64), (0, 316)]],
    ("lq", 216, 216): [[(40, 56), (66, 48), (58, 24), (32, 8), (14, 48), (6, 32)], [(52, 56), (60, 48), (46, 24), (26, 48), (20, 8), (8, 24), (4, 8)]],
    ("lq", 36, 1296): [[(8, 144), (12, 256), (8, 300), (4, 256), (4, 48), (0, 292)]],
    ("lq", 216, 36): [[(48, 8), (56, 12), (48, 8), (24, 4), (24, 4), (16, 0)]],
    ("qr", 1296, 216): [[(108, 28), (204, 48), (224, 56), (160, 48), (64, 24), (536, 12)]],
}


def matrix(rng, m, n, pattern):
    M = np.zeros((m, n))
    rows, cols = rng.permutation(m), rng.permutation(n)
    r0 = c0 = 0
    for ms, ns in pattern:
        assert r0 + ms <= m and c0 + ns <= n, (m, n, pattern)
        if ms and ns:
            k = min(ms, ns)
            u, _ = np.linalg.qr(rng.standard_normal((ms, k)))
            v, _ = np.linalg.qr(rng.standard_normal((ns, k)))
            sv = np.exp(-np.linspace(0, 9, k))      # condition number ~ 1e4 as in the traced boundary bonds
            M[np.ix_(rows[r0:r0 + ms], cols[c0:c0 + ns])] = (u * sv) @ v.T
        r0 += ms; c0 += ns
    return M


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = backend.get(); B.sector_discovery = True
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    rng = np.random.default_rng(0)
    for (kind, m, n), pats in PATTERNS.items():
        mats = np.stack([matrix(rng, m, n, pats[i % len(pats)]).reshape(-1) for i in range(8)])
        a = B.from_numpy(mats[np.arange(nb) % 8])
        p, k = plan_of(m, n, kind == "qr")
        res = {}
        for name, thr in (("chain", 1 << 60), ("queue", 0)):
            B.lib.tnsp_sector_queue_min(thr)
            t1, t2, s = B.zeros(nb, m * k), B.zeros(nb, k * n), B.zeros(nb, k)
            if kind == "svd":
                ms = timeit(lambda: B.svd(p, a, t1, s, t2))
                by = 8 * (m * n + m * k + k * n + k) * nb
            else:
                ms = timeit(lambda: B.qr(p, a, t1, t2))
                by = 8 * (2 * m * n + m * k + k * n) * nb
            res[name] = (ms, [x.clone() for x in (t1, t2, s)])
            print(f"{kind} {m}x{n} nb={nb} {name}: {ms:.3f} ms  ({by / ms / 1e6:.0f} GB/s)", flush=True)
        A = a.reshape(nb, m, n)
        for name in res:
            t1, t2, s = res[name][1]
            rec = (t1.reshape(nb, m, k) * s.reshape(nb, 1, k)) @ t2.reshape(nb, k, n) if kind == "svd" else t1.reshape(nb, m, k) @ t2.reshape(nb, k, n)
            print(f"    {name}: max |factors - A| = {float((rec - A).abs().max()):.2e}", flush=True)
        if kind == "svd":
            print(f"    max |sigma_queue - sigma_chain| = {float((res['queue'][1][2] - res['chain'][1][2]).abs().max()):.2e}")
    B.lib.tnsp_sector_queue_min(2048)

if __name__ == "__main__":
    main()

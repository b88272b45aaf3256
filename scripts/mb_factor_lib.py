import numpy as np

def block_matrix(rng, m, n, n_sec, even=False):
    """block diagonal after per-chain row / column permutations; `even`: sectors of (nearly) equal size"""
    M = np.zeros((m, n))
    rows = rng.permutation(m); cols = rng.permutation(n)
    if even:
        rc = [m * i // n_sec for i in range(1, n_sec)]; cc = [n * i // n_sec for i in range(1, n_sec)]
    else:
        rc = np.sort(rng.choice(np.arange(1, m), n_sec - 1, replace=False)); cc = np.sort(rng.choice(np.arange(1, n), n_sec - 1, replace=False))
    for rs, cs in zip(np.split(rows, rc), np.split(cols, cc)):
        M[np.ix_(rs, cs)] = rng.standard_normal((len(rs), len(cs)))
    return M

def plan_of(m, n, flag):
    class P: pass
    p = P(); k = min(m, n)
    p.sectors = np.array([[m, n, k, 0, 0, 0, 0, 0]], dtype=np.int64); p.s_total = k; p.flag = flag; p._dev = None
    return p, k

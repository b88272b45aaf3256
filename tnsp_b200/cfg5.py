"""cfg5 of BASELINE.json: the TAT block-symmetric contract / qr / svd microbenchmark on random U(1) tensors shaped like the
boundary-MPS compression steps (single_layer_auxiliaries.py:462-498), a batch of independent samples
sharing one block structure.

    line_1 (L1, R1, D)      L1, R1: total dimension Dc over charges -2..2 (weights 1:2:2:2:1); D: 8 over -1..1 (1:2:1)
    line_2 (L2, R2, U, D)   all of total dimension 8 over -1..1 (1:2:1)
    step A  T = contract(line_1, line_2, {(D, U)})                         -> rank-5 (L1, R1, L2, R2, D)
    step B  Q, R = T.qr('r', {R1, R2}, "R", "L")
    step C  U, S, V = T.svd({L1, L2}, "R", "L", "L", "R", cut = Dc)

Used by tests/test_cfg5.py (parity) and scripts/mb_cfg5.py (timing)."""
import numpy as np


def _split(total, weights):
    unit = total // sum(weights)
    dims = [w * unit for w in weights]
    dims[len(dims) // 2] += total - sum(dims)
    return dims


def edges(mod, Dc, d=8):
    """(big, small) U(1) edges of module `mod` (tnsp_b200.TAT or the reference TAT)"""
    S, E = mod.BoseU1.Symmetry, mod.BoseU1.Edge
    big = E([(S(q), n) for q, n in zip((-2, -1, 0, 1, 2), _split(Dc, (1, 2, 2, 2, 1))) if n > 0])
    small = E([(S(q), n) for q, n in zip((-1, 0, 1), _split(d, (1, 2, 1))) if n > 0])
    return big, small


def structures(mod, Dc, d=8):
    """names and edges of line_1 and line_2 (total charge zero, like boundary tensors of a charge-neutral row)"""
    big, small = edges(mod, Dc, d)
    line_1 = (["L1", "R1", "D"], [big, big.conjugate(), small])
    line_2 = (["L2", "R2", "U", "D"], [small, small.conjugate(), small.conjugate(), small])
    return line_1, line_2


def steps(line_1, line_2, Dc):
    """the three operations on tensors of either module; returns (T, (Q, R), (U, S, V))"""
    T = line_1.contract(line_2, {("D", "U")})
    Q, R = T.qr("r", {"R1", "R2"}, "R", "L")
    U, S, V = T.svd({"L1", "L2"}, "R", "L", "L", "R", Dc)
    return T, (Q, R), (U, S, V)


def random_batch(mod, Dc, nb, seed=7, d=8):
    """batched device (or checker) tensors of tnsp_b200.TAT: one structure, nb independent samples"""
    rng = np.random.default_rng(seed)
    (n1, e1), (n2, e2) = structures(mod, Dc, d)
    T = mod.BoseU1.D.Tensor
    s1 = T(n1, e1).storage.size
    s2 = T(n2, e2).storage.size
    v1, v2 = rng.standard_normal((nb, s1)), rng.standard_normal((nb, s2))
    return T.from_batch(n1, e1, v1), T.from_batch(n2, e2, v2), v1, v2

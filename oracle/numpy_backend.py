"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of every C-ABI entry point of
include/tnsp_b200.h, used as the checker for the CUDA kernels and to exercise the host-side logic
on machines without a GPU.  The product (tnsp_b200/) never imports this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may.

Each function follows the reference's per-sector algorithm:
  pack            TAT/include/TAT/utility/multidimension_span.hpp:250-383 (strided copy, optional negate)
                  driven by edge_operator.hpp:651-688
  gemm            contract.hpp:194-250 (loop of ?gemm_ over the descriptor list of :539-616)
  qr              qr.hpp:178-304  (?geqrf/?orgqr or ?gelqf/?orglq -> numpy.linalg.qr, LAPACK underneath)
  svd             svd.hpp:104-211 (?gesvd 'S','S' -> numpy.linalg.svd)
  svd_cut         svd.hpp:429-481 (greedy cross-sector cut, literal restatement)
  diag_scatter    svd.hpp:213-254
  norm/scale/...  tensor.hpp:631-660, scalar.hpp:46-118, conjugate.hpp:99-116
Parity is pinned through tests/test_tat_vs_reference.py and tests/test_reference_kats.py (the real reference in oracle/_ref and
the golden vectors of PyTAT/tests).
"""
from __future__ import annotations

import numpy as np
import torch

from .numpy_ragged import RaggedMixin

R = 8


class NumpyBackend(RaggedMixin):
    name = "numpy-checker"

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    # buffers
    def empty(self, nb, size):
        return torch.zeros((nb, size), dtype=torch.float64)

    def zeros(self, nb, size):
        return torch.zeros((nb, size), dtype=torch.float64)

    def from_numpy(self, array):
        return torch.from_numpy(np.array(array, dtype=None, copy=True))

    def to_numpy(self, t):
        return t.detach().numpy()

    def upload(self, array):
        return torch.from_numpy(np.ascontiguousarray(array))

    def launch_count(self):
        return self.launches

    def synchronize(self):
        pass

    @staticmethod
    def _v(t, nb):
        a = t.detach().numpy()
        return np.broadcast_to(a, (nb, a.shape[1])) if a.shape[0] == 1 and nb != 1 else a

    # K1
    def pack(self, plan, src, dst):
        self.launches += 1
        nb = dst.shape[0]
        s = self._v(src, nb)
        d = dst.numpy()
        for row in plan.desc:
            so, do, sr = int(row[0]), int(row[1]), int(row[2])
            rank, neg = sr >> 1, sr & 1
            dims = [int(x) for x in row[3:3 + rank]]
            ss = [int(x) for x in row[3 + R:3 + R + rank]]
            ds = [int(x) for x in row[3 + 2 * R:3 + 2 * R + rank]]
            sv = np.lib.stride_tricks.as_strided(s[:, so:], shape=[nb] + dims, strides=[s.strides[0]] + [x * 8 for x in ss], writeable=False)
            dv = np.lib.stride_tricks.as_strided(d[:, do:], shape=[nb] + dims, strides=[d.strides[0]] + [x * 8 for x in ds])
            dv[...] = -sv if neg else sv

    # K2
    def gemm(self, plan, a, b, c):
        self.launches += 1
        nb = c.shape[0]
        A, B, C = self._v(a, nb), self._v(b, nb), c.numpy()
        for m, n, k, ao, bo, co, flags, alpha in plan.gemm:
            m, n, k = int(m), int(n), int(k)
            am = A[:, ao:ao + m * k].reshape(nb, k, m).transpose(0, 2, 1) if flags & 1 else A[:, ao:ao + m * k].reshape(nb, m, k)
            bm = B[:, bo:bo + k * n].reshape(nb, n, k).transpose(0, 2, 1) if flags & 2 else B[:, bo:bo + k * n].reshape(nb, k, n)
            C[:, co:co + m * n] = (float(alpha) * np.matmul(am, bm)).reshape(nb, m * n)

    gather_gemm = True

    def gemm_gather(self, plan, a, b, c):
        """contract.hpp:622-857 restated without the intermediate merged copies: operands indexed in place"""
        self.launches += 1
        tab, flags, m, n, k = plan.gather
        aro, aco, bro, bco = tab[:m], tab[m:m + k], tab[m + k:m + 2 * k], tab[m + 2 * k:]
        nb = c.shape[0]
        A, B = self._v(a, nb), self._v(b, nb)
        am = A[:, aro[:, None].astype(np.int64) + aco[None, :]]
        bm = B[:, bro[:, None].astype(np.int64) + bco[None, :]]
        c.numpy()[:, :m * n] = np.matmul(am, bm).reshape(nb, m * n)

    def qr_destroys_input(self, plan):
        return False

    def factor_in_place(self, plan):
        return len(plan.sectors) == 1 and int(plan.sectors[0][0]) * int(plan.sectors[0][1]) >= 2048

    @staticmethod
    def _gathered(plan, a):
        """the merged matrix of every chain, read through the plan's offset table"""
        m, n = int(plan.sectors[0][0]), int(plan.sectors[0][1])
        ro, co = plan.rc_tab[:m].astype(np.int64), plan.rc_tab[m:].astype(np.int64)
        return torch.from_numpy(np.ascontiguousarray(a.numpy()[:, (ro[:, None] + co[None, :]).reshape(-1)]))

    # K3
    def qr(self, plan, a, out1, out2, in_place=False):
        self.launches += 1
        if in_place:
            a = self._gathered(plan, a)
        nb = a.shape[0]
        A, O1, O2 = a.numpy(), out1.numpy(), out2.numpy()
        for m, n, k, ao, o1, o2, _so, _ in plan.sectors:
            m, n, k = int(m), int(n), int(k)
            if m * n == 0:
                continue
            for b in range(nb):
                M = A[b, ao:ao + m * n].reshape(m, n)
                F1 = np.zeros((m, k))
                F2 = np.zeros((k, n))
                k0 = 0
                # one descriptor = dense(-embedded) matrix: factorise per connected block of the zero pattern,
                # null vectors are zeros (what the discovered-sector CUDA kernels do, csrc/factor_sector.cu)
                for rows, cols in (_zero_pattern_blocks(M) if len(plan.sectors) == 1 else [(np.arange(m), np.arange(n))]):
                    Ms = M[np.ix_(rows, cols)]
                    ks = min(Ms.shape)
                    if plan.flag:
                        q, r = np.linalg.qr(Ms, mode="reduced")
                        F1[rows, k0:k0 + ks] = q
                        F2[np.ix_(np.arange(k0, k0 + ks), cols)] = r
                    else:
                        q, r = np.linalg.qr(Ms.T, mode="reduced")
                        F1[rows, k0:k0 + ks] = r.T
                        F2[np.ix_(np.arange(k0, k0 + ks), cols)] = q.T
                    k0 += ks
                O1[b, o1:o1 + m * k] = F1.reshape(-1)
                O2[b, o2:o2 + k * n] = F2.reshape(-1)

    # K4
    def svd(self, plan, a, out1, s, out2, in_place=False):
        self.launches += 1
        if in_place:
            a = self._gathered(plan, a)
        nb = a.shape[0]
        A, O1, S, O2 = a.numpy(), out1.numpy(), s.numpy(), out2.numpy()
        for m, n, k, ao, o1, o2, so, _ in plan.sectors:
            m, n, k = int(m), int(n), int(k)
            if m * n == 0:
                continue
            for b in range(nb):
                M = A[b, ao:ao + m * n].reshape(m, n)
                if len(plan.sectors) != 1:
                    u, sv, vt = np.linalg.svd(M, full_matrices=False)
                else:
                    # per connected block of the zero pattern, then merged in descending order (ties: block order)
                    us, ss, vs = [], [], []
                    for rows, cols in _zero_pattern_blocks(M):
                        bu, bs, bv = np.linalg.svd(M[np.ix_(rows, cols)], full_matrices=False)
                        for t in range(len(bs)):
                            cu_ = np.zeros(m)
                            cu_[rows] = bu[:, t]
                            cv_ = np.zeros(n)
                            cv_[cols] = bv[t]
                            us.append(cu_)
                            ss.append(bs[t])
                            vs.append(cv_)
                    order = sorted(range(len(ss)), key=lambda t: (-ss[t], t))
                    u, sv, vt = np.zeros((m, k)), np.zeros(k), np.zeros((k, n))
                    for r_, t in enumerate(order):
                        u[:, r_], sv[r_], vt[r_] = us[t], ss[t], vs[t]
                O1[b, o1:o1 + m * k] = u.reshape(-1)
                S[b, so:so + k] = sv
                O2[b, o2:o2 + k * n] = vt.reshape(-1)

    def svd_cut(self, plan, s, remain_cut, relative_cut):
        """literal restatement of svd.hpp:429-470"""
        self.launches += 1
        nb = s.shape[0]
        S = s.numpy()
        ns = len(plan.sectors)
        counts = np.zeros((nb, ns), dtype=np.int32)
        for b in range(nb):
            vecs = [S[b, int(r[6]):int(r[6]) + int(r[2])] for r in plan.sectors]
            total = sum(len(v) for v in vecs)
            mx = max([float(v.max()) for v in vecs if len(v)] + [0.0])
            thr = relative_cut * mx
            rc = min(int(remain_cut), total)
            for _ in range(rc):
                best, best_v = -1, 0.0
                for i, v in enumerate(vecs):
                    if counts[b, i] != len(v) and v[counts[b, i]] > best_v:
                        best_v = float(v[counts[b, i]])
                        best = i
                if best_v > thr:
                    counts[b, best] += 1
                else:
                    break
        return torch.from_numpy(counts)

    def svd_mask(self, plan, counts, out1, s, out2):
        self.launches += 1
        O1, S, O2 = out1.numpy(), s.numpy(), out2.numpy()
        cn = counts.numpy()
        for i, (m, n, k, _ao, o1, o2, so, _) in enumerate(plan.sectors):
            m, n, k = int(m), int(n), int(k)
            for b in range(S.shape[0]):
                keep = int(cn[b, i])
                if keep >= k:
                    continue
                O1[b, o1:o1 + m * k].reshape(m, k)[:, keep:] = 0
                O2[b, o2:o2 + k * n].reshape(k, n)[keep:, :] = 0
                S[b, so + keep:so + k] = 0

    def diag_scatter(self, blk, s, dst):
        self.launches += 1
        S, D = s.numpy(), dst.numpy()
        for so, do, r, sign in blk:
            so, do, r = int(so), int(do), int(r)
            idx = do + np.arange(r) * (r + 1)
            D[:, idx] = -S[:, so:so + r] if sign else S[:, so:so + r]

    # streaming
    def norm(self, x, kind):
        self.launches += 1
        a = x.numpy()
        if kind == -1:
            r = np.abs(a).max(axis=1) if a.shape[1] else np.zeros(a.shape[0])
        elif kind == 1:
            r = np.abs(a).sum(axis=1)
        else:
            r = np.sqrt((a * a).sum(axis=1))
        return torch.from_numpy(np.ascontiguousarray(r))

    def scale(self, x, alpha, op, nb=None):
        self.launches += 1
        nb = max(x.shape[0], alpha.shape[0]) if nb is None else nb
        a = alpha.numpy().reshape(-1, 1)
        X = self._v(x, nb)
        return torch.from_numpy(np.ascontiguousarray(X / a if op else X * a))

    def binary(self, a, b, op):
        self.launches += 1
        nb = max(a.shape[0], b.shape[0])
        A, B = self._v(a, nb), self._v(b, nb)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = [A + B, A - B, A * B, A / B][op]
        return torch.from_numpy(np.ascontiguousarray(r))

    def unary(self, a, op):
        self.launches += 1
        A = a.numpy()
        with np.errstate(divide="ignore"):
            r = [np.sqrt(np.abs(A)), np.where(A == 0, 0.0, 1.0 / np.where(A == 0, 1.0, A)), -A, np.abs(A)][op]
        return torch.from_numpy(np.ascontiguousarray(r))

    def block_sign(self, blk, x):
        self.launches += 1
        r = x.numpy().copy()
        for off, size, sign in blk:
            if sign:
                r[:, int(off):int(off) + int(size)] *= -1
        return torch.from_numpy(r)

    def gather_rows(self, src, row_size, index):
        self.launches += 1
        return torch.from_numpy(np.ascontiguousarray(src.numpy().reshape(-1, row_size)[index.numpy()]))

    def select(self, mask, a, b):
        self.launches += 1
        nb = mask.shape[0]
        m = mask.numpy().astype(bool).reshape(-1, 1)
        return torch.from_numpy(np.ascontiguousarray(np.where(m, self._v(a, nb), self._v(b, nb))))

    def grad_accumulate(self, holes, weight, energy, delta, edelta):
        self.launches += 1
        w = weight.numpy().reshape(-1, 1)
        h = np.where(w != 0, holes.numpy(), 0.0) * w   # chains of zero weight (zero amplitude) are skipped
        delta.numpy()[...] += h.sum(axis=0)
        edelta.numpy()[...] += (h * energy.numpy().reshape(-1, 1)).sum(axis=0)


def _zero_pattern_blocks(M):
    """connected components of the bipartite row/column graph of the non-zeros of M, ordered by their first
    row (the sector order of csrc/factor_sector.cu); all-zero rows / columns belong to no block"""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    m, n = M.shape
    nz = M != 0
    g = csr_matrix((np.ones(int(nz.sum()), dtype=np.int8), (np.nonzero(nz)[0], np.nonzero(nz)[1] + m)), shape=(m + n, m + n))
    _, lab = connected_components(g, directed=False)
    used_r, used_c = nz.any(axis=1), nz.any(axis=0)
    out = []
    seen = set()
    for i in range(m):
        if used_r[i] and lab[i] not in seen:
            seen.add(lab[i])
            out.append((np.nonzero((lab[:m] == lab[i]) & used_r)[0], np.nonzero((lab[m:] == lab[i]) & used_c)[0]))
    return out


def install():
    """Install the checker backend into tnsp_b200 (tests only)."""
    from tnsp_b200 import backend
    b = NumpyBackend()
    backend.set_backend(b)
    return b

"""Host-side arithmetic for the scalar types that are NOT on the hot path: ``TAT.<symmetry>.{S, C, Z}.Tensor`` (float32, complex64,
complex128; PyTAT/PyTAT.cpp:32-46 instantiates all four for every symmetry).

The sampling-VMC path is float64 and lives on the B200 (tnsp_b200.backend: no CPU fallback).  The other scalar types only appear
in model definitions -- the reference's own ``tetragono.common_tensor`` builds every operator as a complex tensor at import time
(common_tensor/No.py:22-24) and models convert them with ``.to(float)`` -- so they are plain numpy here: a few d^2 x d^2 arrays,
never a sample.  Same descriptor-driven operations as the device backend (the host planner of plan.py is shared), any dtype.
"""
from __future__ import annotations

import numpy as np
import torch

R = 8
_TORCH = {"float32": torch.float32, "float64": torch.float64, "complex64": torch.complex64, "complex128": torch.complex128}
_BACKENDS: dict = {}


def backend(dtype):
    b = _BACKENDS.get(dtype)
    if b is None:
        b = _BACKENDS[dtype] = HostScalarBackend(dtype)
    return b


class HostScalarBackend:
    name = "host-scalars"
    gather_gemm = False
    sector_discovery = False

    def __init__(self, dtype):
        self.dtype = dtype
        self.np = np.dtype(dtype)
        self.real = np.dtype({"float32": "float32", "complex64": "float32"}.get(dtype, "float64"))
        self.torch = _TORCH[dtype]
        self.device = torch.device("cpu")

    # ---- buffers ----
    def zeros(self, nb, size):
        return torch.zeros((nb, size), dtype=self.torch)

    empty = zeros

    def from_numpy(self, array):
        a = np.array(array, copy=True)
        if a.dtype.kind in "fc":
            a = a.astype(self.np if a.dtype.kind == "c" or self.np.kind == "c" else self.np)
        return torch.from_numpy(np.ascontiguousarray(a))

    def to_numpy(self, t):
        return t.detach().numpy()

    def upload(self, array):
        return torch.from_numpy(np.ascontiguousarray(array))

    def launch_count(self):
        return 0

    def synchronize(self):
        pass

    @staticmethod
    def _v(t, nb):
        a = t.detach().numpy()
        return np.broadcast_to(a, (nb, a.shape[1])) if a.shape[0] == 1 and nb != 1 else a

    def factor_in_place(self, plan):
        return False

    def qr_destroys_input(self, plan):
        return False

    # ---- edge operator copy loop (edge_operator.hpp:651-688) ----
    def pack(self, plan, src, dst):
        nb = dst.shape[0]
        s, d = self._v(src, nb), dst.numpy()
        item = d.itemsize
        for row in plan.desc:
            so, do, sr = int(row[0]), int(row[1]), int(row[2])
            rank, neg = sr >> 1, sr & 1
            dims = [int(x) for x in row[3:3 + rank]]
            ss = [int(x) for x in row[3 + R:3 + R + rank]]
            ds = [int(x) for x in row[3 + 2 * R:3 + 2 * R + rank]]
            sv = np.lib.stride_tricks.as_strided(s[:, so:], shape=[nb] + dims, strides=[s.strides[0]] + [x * item for x in ss], writeable=False)
            dv = np.lib.stride_tricks.as_strided(d[:, do:], shape=[nb] + dims, strides=[d.strides[0]] + [x * item for x in ds])
            dv[...] = -sv if neg else sv

    # ---- per-sector gemm (contract.hpp:194-250) ----
    def gemm(self, plan, a, b, c):
        nb = c.shape[0]
        A, B, C = self._v(a, nb), self._v(b, nb), c.numpy()
        for m, n, k, ao, bo, co, flags, alpha in plan.gemm:
            m, n, k = int(m), int(n), int(k)
            am = A[:, ao:ao + m * k].reshape(nb, k, m).transpose(0, 2, 1) if flags & 1 else A[:, ao:ao + m * k].reshape(nb, m, k)
            bm = B[:, bo:bo + k * n].reshape(nb, n, k).transpose(0, 2, 1) if flags & 2 else B[:, bo:bo + k * n].reshape(nb, k, n)
            C[:, co:co + m * n] = (float(alpha) * np.matmul(am, bm)).reshape(nb, m * n)

    # ---- per-sector factorisations (qr.hpp:178-304, svd.hpp:104-211) ----
    def qr(self, plan, a, out1, out2, in_place=False):
        A, O1, O2 = a.numpy(), out1.numpy(), out2.numpy()
        for m, n, k, ao, o1, o2, _so, _ in plan.sectors:
            m, n, k = int(m), int(n), int(k)
            if m * n == 0:
                continue
            for b in range(A.shape[0]):
                M = A[b, ao:ao + m * n].reshape(m, n)
                if plan.flag:
                    q, r = np.linalg.qr(M, mode="reduced")
                    O1[b, o1:o1 + m * k], O2[b, o2:o2 + k * n] = q.reshape(-1), r.reshape(-1)
                else:
                    q, r = np.linalg.qr(M.T, mode="reduced")
                    O1[b, o1:o1 + m * k], O2[b, o2:o2 + k * n] = r.T.reshape(-1), q.T.reshape(-1)

    def svd(self, plan, a, out1, s, out2, in_place=False):
        A, O1, S, O2 = a.numpy(), out1.numpy(), s.numpy(), out2.numpy()
        for m, n, k, ao, o1, o2, so, _ in plan.sectors:
            m, n, k = int(m), int(n), int(k)
            if m * n == 0:
                continue
            for b in range(A.shape[0]):
                u, sv, vt = np.linalg.svd(A[b, ao:ao + m * n].reshape(m, n), full_matrices=False)
                O1[b, o1:o1 + m * k], S[b, so:so + k], O2[b, o2:o2 + k * n] = u.reshape(-1), sv, vt.reshape(-1)

    def svd_cut(self, plan, s, remain_cut, relative_cut):
        """greedy cross-sector cut (svd.hpp:429-470)"""
        S = np.abs(s.numpy())
        nb, ns = S.shape[0], len(plan.sectors)
        counts = np.zeros((nb, ns), dtype=np.int32)
        for b in range(nb):
            vecs = [S[b, int(r[6]):int(r[6]) + int(r[2])] for r in plan.sectors]
            top = max([float(v.max()) for v in vecs if len(v)] + [0.0])
            for _ in range(min(int(remain_cut), sum(len(v) for v in vecs))):
                best, best_v = -1, 0.0
                for i, v in enumerate(vecs):
                    if counts[b, i] != len(v) and v[counts[b, i]] > best_v:
                        best, best_v = i, float(v[counts[b, i]])
                if best_v > relative_cut * top:
                    counts[b, best] += 1
                else:
                    break
        return torch.from_numpy(counts)

    def svd_mask(self, plan, counts, out1, s, out2):
        O1, S, O2 = out1.numpy(), s.numpy(), out2.numpy()
        cn = counts.numpy()
        for i, (m, n, k, _ao, o1, o2, so, _) in enumerate(plan.sectors):
            m, n, k = int(m), int(n), int(k)
            for b in range(S.shape[0]):
                keep = int(cn[b, i])
                if keep < k:
                    O1[b, o1:o1 + m * k].reshape(m, k)[:, keep:] = 0
                    O2[b, o2:o2 + k * n].reshape(k, n)[keep:, :] = 0
                    S[b, so + keep:so + k] = 0

    def diag_scatter(self, blk, s, dst):
        S, D = s.numpy(), dst.numpy()
        for so, do, r, sign in blk:
            so, do, r = int(so), int(do), int(r)
            idx = do + np.arange(r) * (r + 1)
            D[:, idx] = -S[:, so:so + r] if sign else S[:, so:so + r]

    def _div(self, x, y):
        """x / y as the reference's C++ computes it: std::complex division is the textbook formula (libgcc __divdc3 without its
        overflow rescue), numpy's loop uses Smith's algorithm and differs in the last bit"""
        if self.np.kind != "c":
            return x / y
        x, y = np.asarray(x, dtype=self.np), np.asarray(y, dtype=self.np)
        a, b, c, d = x.real, x.imag, y.real, y.imag
        den = c * c + d * d
        with np.errstate(divide="ignore", invalid="ignore"):
            return ((a * c + b * d) / den + 1j * ((b * c - a * d) / den)).astype(self.np)

    # ---- elementwise (scalar.hpp:46-118, tensor.hpp:631-660, conjugate.hpp:99-116) ----
    def norm(self, x, kind):
        a = np.abs(x.numpy()).astype(np.float64)
        r = (a.max(axis=1) if a.shape[1] else np.zeros(a.shape[0])) if kind == -1 else (a.sum(axis=1) if kind == 1 else np.sqrt((a * a).sum(axis=1)))
        return torch.from_numpy(np.ascontiguousarray(r))

    def scale(self, x, alpha, op, nb=None):
        nb = max(x.shape[0], alpha.shape[0]) if nb is None else nb
        a = alpha.numpy().reshape(-1, 1)
        X = self._v(x, nb)
        return torch.from_numpy(np.ascontiguousarray((self._div(X, a) if op else X * a).astype(self.np)))

    def binary(self, a, b, op):
        nb = max(a.shape[0], b.shape[0])
        A, B = self._v(a, nb), self._v(b, nb)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = [A + B, A - B, A * B, None][op] if op != 3 else self._div(A, B)
        return torch.from_numpy(np.ascontiguousarray(r.astype(self.np)))

    def unary(self, a, op):
        A = a.numpy()
        with np.errstate(divide="ignore"):
            r = [np.sqrt(np.abs(A)), np.where(A == 0, 0.0, 1.0 / np.where(A == 0, 1.0, A)), -A, np.abs(A), np.conj(A)][op]
        return torch.from_numpy(np.ascontiguousarray(r.astype(self.np)))

    def block_sign(self, blk, x):
        r = x.numpy().copy()
        for off, size, sign in blk:
            if sign:
                r[:, int(off):int(off) + int(size)] *= -1
        return torch.from_numpy(r)

    def gather_rows(self, src, row_size, index):
        return torch.from_numpy(np.ascontiguousarray(src.numpy().reshape(-1, row_size)[index.numpy()]))

    def select(self, mask, a, b):
        nb = mask.shape[0]
        m = mask.numpy().astype(bool).reshape(-1, 1)
        return torch.from_numpy(np.ascontiguousarray(np.where(m, self._v(a, nb), self._v(b, nb))))

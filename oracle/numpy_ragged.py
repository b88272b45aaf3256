"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, one Python loop per chain) of the sector-compact kernels
(`backend.rt_*`, csrc/ragged.cu and the rt_* part of csrc/factor_sector.cu).  It is the specification the CUDA kernels are
tested against and lets the host logic of tnsp_b200/TAT/ragged.py run on machines without a GPU.  The product never imports it.

Per chain this is literally the reference's block-symmetric algorithm:
  rt_sort    merged edge of a group of edges: indices grouped by summed symmetry (edge_operator.hpp:321-404; sectors are kept
             in ascending charge order here, the order is not observable)
  rt_match   pairing of row and column sectors whose symmetries sum to the tensor's total (core.hpp:162-190)
  rt_repack  transpose / regroup (edge_operator.hpp:651-688)
  rt_gemm    one ?gemm per sector (contract.hpp:539-616)
  rt_factor  per-sector QR (qr.hpp:178-304, common edge qr.hpp:419-429) and SVD with the global greedy cut (svd.hpp:104-211,
             429-481), numpy.linalg (LAPACK) underneath
"""
from __future__ import annotations

import numpy as np
import torch

SMAX = 64
HDR = 3 + 2 * SMAX
MSTRIDE = 4 + 2 * SMAX
DEAD = 1 << 30
DEAD_MIN = 1 << 29


def _np(t):
    return t.detach().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _row(a, c):
    return a[c if a.shape[0] > 1 else 0]


class Table:
    """decoded view of one chain's group table"""

    def __init__(self, row, M):
        self.nsec, self.nvalid = int(row[0]), int(row[1])
        self.skey = row[2:2 + SMAX]
        self.sstart = row[2 + SMAX:3 + 2 * SMAX]
        self.perm = row[HDR:HDR + M]
        self.inv = row[HDR + M:HDR + 2 * M]

    def count(self, i):
        return int(self.sstart[i + 1] - self.sstart[i])

    def sector_of(self, pos):
        return int(np.searchsorted(self.sstart[:self.nsec + 1], pos, side="right") - 1)

    def find(self, key):
        for i in range(self.nsec):
            if int(self.skey[i]) == key:
                return i
        return -1


class Match:
    def __init__(self, row):
        self.size = int(row[0])
        self.moff = row[2:3 + SMAX]
        self.mcol = row[3 + SMAX:3 + 2 * SMAX]


def table_M(tab):
    return (tab.shape[1] - HDR) // 2


def _put(row, off, mat):
    """store one sector matrix; the alignment pad behind an odd-sized sector is an explicit zero"""
    flat = np.asarray(mat, dtype=np.float64).reshape(-1)
    row[off:off + flat.size] = flat
    if flat.size & 1:
        row[off + flat.size] = 0.0


class RaggedMixin:
    def rt_alloc(self, nb, size):
        return torch.full((nb, max(int(size), 1) + SMAX), float("nan"), dtype=torch.float64)   # NaN: reading an unwritten element shows

    # ---- merged group: indices sorted by summed label --------------------------------------------------------------
    def rt_sort(self, edges):
        self.launches += 1
        nbT = max([int(a.shape[0]) for a, _, _ in edges] + [1])
        M = 1
        for _, _, d in edges:
            M *= int(d)
        out = np.zeros((nbT, HDR + 2 * M), dtype=np.int32)
        for c in range(nbT):
            key = np.zeros(1, dtype=np.int64)
            dead = np.zeros(1, dtype=bool)
            for a, s, d in edges:
                lab = _row(_np(a), c).astype(np.int64)
                dd = np.abs(lab) >= DEAD_MIN
                key = (key[:, None] + int(s) * lab[None, :]).reshape(-1)
                dead = (dead[:, None] | dd[None, :]).reshape(-1)
            idx = np.arange(M)
            valid = idx[~dead]
            order = valid[np.argsort(key[valid], kind="stable")]
            perm = np.concatenate([order, idx[dead]])
            keys = key[order]
            row = out[c]
            distinct = sorted(set(int(k) for k in keys))
            if len(distinct) > SMAX:
                raise RuntimeError("rt_sort: more than SMAX sectors in one group")
            row[0], row[1] = len(distinct), len(order)
            for i, k in enumerate(distinct):
                row[2 + i] = k
                row[2 + SMAX + i] = int(np.searchsorted(keys, k, side="left"))
            row[2 + SMAX + len(distinct)] = len(order)
            row[HDR:HDR + M] = perm
            inv = np.empty(M, dtype=np.int32)
            inv[perm] = np.arange(M, dtype=np.int32)
            row[HDR + M:HDR + 2 * M] = inv
        return torch.from_numpy(out)

    # ---- sector pairing -------------------------------------------------------------------------------------------
    def rt_match(self, rt, rs, ct, cs, t1, s1, t2, s2, nbm, cap=0):
        self.launches += 1
        R, C = _np(rt), _np(ct)
        M, N = table_M(R), table_M(C)
        nbm = max(int(nbm), R.shape[0], C.shape[0], 1 if t1 is None else int(t1.shape[0]), 1 if t2 is None else int(t2.shape[0]))
        out = np.zeros((nbm, MSTRIDE), dtype=np.int32)
        tsum = np.zeros(nbm, dtype=np.int32)
        for c in range(nbm):
            tr, tc = Table(_row(R, c), M), Table(_row(C, c), N)
            t = 0
            if t1 is not None:
                t += int(s1) * int(_row(_np(t1).reshape(-1, 1), c)[0])
            if t2 is not None:
                t += int(s2) * int(_row(_np(t2).reshape(-1, 1), c)[0])
            tsum[c] = t
            row = out[c]
            off = 0
            for i in range(tr.nsec):
                row[2 + i] = off
                want = t - int(rs) * int(tr.skey[i])
                j = -1
                for jj in range(tc.nsec):
                    if int(cs) * int(tc.skey[jj]) == want:
                        j = jj
                        break
                row[3 + SMAX + i] = j
                if j >= 0:
                    off += tr.count(i) * tc.count(j)
                    off += off & 1          # sector matrices start at even element offsets (16-byte aligned bulk copies)
            row[2 + tr.nsec] = off
            row[0] = off
            if cap and off > cap:          # learnt capacity exceeded: the chain is stored empty and counted
                self.rt_overflows = getattr(self, "rt_overflows", 0) + 1
                row[:] = 0
                row[1] = 1
                row[3 + SMAX:3 + 2 * SMAX] = -1
        return torch.from_numpy(out), (torch.from_numpy(tsum) if (t1 is not None or t2 is not None) else None)

    # ---- regroup --------------------------------------------------------------------------------------------------
    @staticmethod
    def _decode(plan):
        p = _np(plan)
        nr, nc = int(p[0]), int(p[1])
        ent = p[2:].reshape(-1, 5)[:, :3]          # (dim, source group, source stride); the magic numbers are for the device
        return [tuple(int(x) for x in e) for e in ent[:nr]], [tuple(int(x) for x in e) for e in ent[nr:nr + nc]]

    @staticmethod
    def _source_index(entries, idx):
        """merged destination group indices -> (source row part, source column part)"""
        r = np.zeros_like(idx)
        c = np.zeros_like(idx)
        rest = idx.copy()
        for dim, in_col, stride in reversed(entries):
            i = rest % dim
            rest = rest // dim
            if in_col:
                c += i * stride
            else:
                r += i * stride
        return r, c

    def rt_repack_pair(self, plan0, src0, dst0, spec0, plan1, src1, dst1, spec1):
        self.rt_repack(plan0, src0, dst0, spec0)
        self.rt_repack(plan1, src1, dst1, spec1)

    @staticmethod
    def _parity(labels, mask):
        labels = np.asarray(labels, dtype=np.int64)
        c0 = ((labels + 32768) % 65536) - 32768
        c1 = (labels - c0) // 65536
        p = np.zeros(labels.shape, dtype=np.int64)
        if mask & 1:
            p ^= c0 & 1
        if mask & 2:
            p ^= c1 & 1
        return p

    def _entry_bits(self, entries, first, idx, labels, c, mask):
        """parity bit k (k = first + position in `entries`) of every merged destination index in `idx`"""
        bits = np.zeros_like(idx)
        rest = idx.copy()
        for k in range(len(entries) - 1, -1, -1):
            dim = entries[k][0]
            i = rest % dim
            rest = rest // dim
            lab = _row(_np(labels[first + k][0]), c)
            bits |= self._parity(lab[i], mask) << (first + k)
        return bits

    def rt_repack(self, plan, src, dst, match_spec=None, sign=None):
        self.launches += 1
        if match_spec is not None:
            rs, cs, t1, s1, t2, s2 = match_spec
            dst.match, _ = self.rt_match(dst.rt, rs, dst.ct, cs, t1, s1, t2, s2, 1, dst.data.shape[1])
        rows, cols = self._decode(plan)
        src_dense = isinstance(src, torch.Tensor)
        dst_dense = isinstance(dst, torch.Tensor)
        S = _np(src if src_dense else src.data)
        D = _np(dst if dst_dense else dst.data)
        dst_dense_size = D.shape[1] if dst_dense else 0
        nb = D.shape[0]
        if not src_dense:
            nb = max(nb, src.match.shape[0])
        for c in range(nb):
            if dst_dense:
                n_el = dst_dense_size
                e = np.arange(n_el, dtype=np.int64)
                r_idx, c_idx = e, np.zeros_like(e)
                dst_pos = e
            else:
                tr, tc = Table(_row(_np(dst.rt), c), dst.M), Table(_row(_np(dst.ct), c), dst.N)
                mt = Match(_row(_np(dst.match), c))
                r_list, c_list, p_list = [], [], []
                for i in range(tr.nsec):
                    j = int(mt.mcol[i])
                    if j < 0:
                        continue
                    m, n = tr.count(i), tc.count(j)
                    rr = tr.perm[tr.sstart[i]:tr.sstart[i] + m].astype(np.int64)
                    cc = tc.perm[tc.sstart[j]:tc.sstart[j] + n].astype(np.int64)
                    r_list.append(np.repeat(rr, n))
                    c_list.append(np.tile(cc, m))
                    p_list.append(int(mt.moff[i]) + np.arange(m * n, dtype=np.int64))
                    if (m * n) & 1:
                        D[c, int(mt.moff[i]) + m * n] = 0.0
                if not r_list:
                    continue
                r_idx, c_idx, dst_pos = np.concatenate(r_list), np.concatenate(c_list), np.concatenate(p_list)
            sr1, sc1 = self._source_index(rows, r_idx)
            sr2, sc2 = self._source_index(cols, c_idx)
            sr, sc = sr1 + sr2, sc1 + sc2
            srow = _row(S, c)
            if src_dense:
                val = srow[sr + sc]
            else:
                tr, tc = Table(_row(_np(src.rt), c), src.M), Table(_row(_np(src.ct), c), src.N)
                mt = Match(_row(_np(src.match), c))
                val = np.zeros(len(sr))
                p = tr.inv[sr].astype(np.int64)
                q = tc.inv[sc].astype(np.int64)
                for k in range(len(sr)):
                    if p[k] >= tr.nvalid or q[k] >= tc.nvalid:
                        continue
                    i = tr.sector_of(p[k])
                    j = int(mt.mcol[i])
                    if j < 0 or not (tc.sstart[j] <= q[k] < tc.sstart[j + 1]):
                        continue
                    val[k] = srow[int(mt.moff[i]) + (p[k] - tr.sstart[i]) * tc.count(j) + (q[k] - tc.sstart[j])]
            if sign is not None:
                # sign(element) = s0 + lin . x + sum_{k<j} Q_kj x_k x_j over the parity bits x of the destination's indexed edges
                quad, per_chain, labels, mask = sign
                quad, pc = _np(quad).astype(np.int64), int(_row(_np(per_chain).reshape(-1, 1), c)[0])
                x = self._entry_bits(rows, 0, r_idx, labels, c, mask) | self._entry_bits(cols, len(rows), c_idx, labels, c, mask)
                sg = np.full(len(x), (pc >> 31) & 1, dtype=np.int64)
                y = x & (pc & 0x7FFFFFFF)
                for k in range(len(rows) + len(cols)):
                    sg ^= (y >> k) & 1
                    z = x & int(quad[k]) if k < len(quad) else np.zeros_like(x)
                    par = np.zeros_like(x)
                    for j in range(len(rows) + len(cols)):
                        par ^= (z >> j) & 1
                    sg ^= ((x >> k) & 1) & par
                val = np.where(sg == 1, -val, val)
            D[c, dst_pos] = val

    def rt_dot(self, plan, src, dst, t1, s1, t2, s2, nb):
        """full contraction: sum over the stored elements of dst of dst[e] * src[same multi-index] (specification: regroup src into
        dst's layout, multiply elementwise, add)"""
        self.launches += 1

        class _F:
            pass
        tmp = _F()
        for k in ("rows", "cols", "rt", "rs", "ct", "cs", "match", "M", "N"):
            setattr(tmp, k, getattr(dst, k))
        tmp.data = self.rt_alloc(max(nb, dst.data.shape[0]), dst.M * dst.N)
        self.rt_repack(plan, src, tmp)
        D, T = _np(dst.data), _np(tmp.data)
        out = self.rt_alloc(nb, 2)
        o = out.numpy()
        match = np.zeros((nb, MSTRIDE), dtype=np.int32)
        tsum = np.zeros(nb, dtype=np.int32)
        for c in range(nb):
            t = 0
            if t1 is not None:
                t += int(s1) * int(_row(_np(t1).reshape(-1, 1), c)[0])
            if t2 is not None:
                t += int(s2) * int(_row(_np(t2).reshape(-1, 1), c)[0])
            tsum[c] = t
            size = int(_row(_np(dst.match), c)[0])
            o[c, 0] = float(np.dot(_row(D, c)[:size], _row(T, c)[:size])) if t == 0 else 0.0
            o[c, 1] = 0.0
            if t == 0:
                match[c, 0], match[c, 3] = 2, 2
            else:
                match[c, 3 + SMAX] = -1
        return out, torch.from_numpy(match), (torch.from_numpy(tsum) if (t1 is not None or t2 is not None) else None)

    # ---- grouped GEMM over (chain, sector) --------------------------------------------------------------------------
    def rt_gemm(self, A, B, C, ksign, nb, match_spec=None):
        self.launches += 1
        tsum = None
        if match_spec is not None:
            rs, cs, t1, s1, t2, s2 = match_spec
            C.match, tsum = self.rt_match(C.rt, rs, C.ct, cs, t1, s1, t2, s2, 1, C.data.shape[1])
        Ad, Bd, Cd = _np(A.data), _np(B.data), _np(C.data)
        flops = 0
        for c in range(nb):
            ar, ak, am = Table(_row(_np(A.rt), c), A.M), Table(_row(_np(A.ct), c), A.N), Match(_row(_np(A.match), c))
            bk, bn, bm = Table(_row(_np(B.rt), c), B.M), Table(_row(_np(B.ct), c), B.N), Match(_row(_np(B.match), c))
            cm = Match(_row(_np(C.match), c))
            a_row, b_row = _row(Ad, c), _row(Bd, c)
            for i in range(ar.nsec):
                jc = int(cm.mcol[i])
                if jc < 0:
                    continue
                m, n = ar.count(i), bn.count(jc)
                out = np.zeros((m, n))
                jk = int(am.mcol[i])
                if jk >= 0:
                    ib = bk.find(int(ksign) * int(ak.skey[jk]))
                    if ib >= 0 and int(bm.mcol[ib]) == jc:
                        k = ak.count(jk)
                        assert k == bk.count(ib), "rt_gemm: common sector sizes differ"
                        a = a_row[int(am.moff[i]):int(am.moff[i]) + m * k].reshape(m, k)
                        b = b_row[int(bm.moff[ib]):int(bm.moff[ib]) + k * n].reshape(k, n)
                        out = a @ b
                        flops += 2 * m * n * k
                _put(Cd[c], int(cm.moff[i]), out)
        self.rt_flops = getattr(self, "rt_flops", 0) + flops
        return tsum

    # ---- per-sector factorisations ----------------------------------------------------------------------------------
    def rt_overflow(self, clear=True):
        n = getattr(self, "rt_overflows", 0)
        if clear:
            self.rt_overflows = 0
        return n

    def rt_factor(self, kind, F, fsign, tt, tts, t1, t1s, kdim, remain_cut, relative_cut, nb, caps=None):
        """F = rows | cols storage of the tensor (effective labels = fsign * stored), (tt, tts) its target, (t1, t1s) the target
        of the first factor.  Bond label of row sector i on the first factor: lam = t1 - (row charge of i)."""
        self.launches += 1
        Fd = _np(F.data)
        kd = max(int(kdim), 1)
        labels = np.full((nb, kd), DEAD, dtype=np.int32)
        per_chain = []
        for c in range(nb):
            tr, tc = Table(_row(_np(F.rt), c), F.M), Table(_row(_np(F.ct), c), F.N)
            mt = Match(_row(_np(F.match), c))
            tq = 0 if t1 is None else int(t1s) * int(_row(_np(t1).reshape(-1, 1), c)[0])
            row = _row(Fd, c)
            secs = []
            for i in range(tr.nsec):
                j = int(mt.mcol[i])
                if j < 0:
                    continue
                m, n = tr.count(i), tc.count(j)
                if m == 0 or n == 0:
                    continue
                a = row[int(mt.moff[i]):int(mt.moff[i]) + m * n].reshape(m, n)
                lam = tq - int(fsign) * int(F.rs) * int(tr.skey[i])
                if kind == "qr":
                    q, r = np.linalg.qr(a)
                    secs.append(dict(i=i, j=j, lam=lam, first=q, second=r, keep=min(m, n)))
                else:
                    u, s, vt = np.linalg.svd(a, full_matrices=False)
                    secs.append(dict(i=i, j=j, lam=lam, first=u, s=s, second=vt, keep=len(s)))
            if kind == "svd":
                # global greedy cut (svd.hpp:429-481): repeatedly take the sector whose next singular value is largest
                # (strict >, first sector wins ties), stop at remain_cut values or at relative_cut * sigma_max
                taken = [0] * len(secs)
                top = max([float(x["s"][0]) for x in secs if len(x["s"])] + [0.0])
                total = 0
                while total < remain_cut:
                    best, which = 0.0, -1
                    for k, x in enumerate(secs):
                        if taken[k] < len(x["s"]) and float(x["s"][taken[k]]) > best:
                            best, which = float(x["s"][taken[k]]), k
                    if which < 0 or not best > relative_cut * top:
                        break
                    taken[which] += 1
                    total += 1
                for k, x in enumerate(secs):
                    x["keep"] = taken[k]
            pos = 0
            for x in secs:
                labels[c, pos:pos + x["keep"]] = x["lam"]
                pos += x["keep"]
            assert pos <= kd
            per_chain.append(secs)
        lab_t = torch.from_numpy(labels)
        tab = self.rt_sort([(lab_t, 1, kd)])
        out = {"labels": lab_t, "bond_col": (tab, 1), "bond_row": (tab, -1)}
        first = self.rt_alloc(nb, caps[0] if caps else F.M * kd)
        second = self.rt_alloc(nb, caps[1] if caps else kd * F.N)
        s_data = self.rt_alloc(nb, kd * kd)
        m_first, _ = self.rt_match(F.rt, F.rs * fsign, tab, 1, t1, t1s, None, 0, nb, first.shape[1])
        m_second, _ = self.rt_match(tab, -1, F.ct, F.cs * fsign, tt, tts, t1, -t1s if t1 is not None else 0, nb, second.shape[1])
        m_s, _ = self.rt_match(tab, -1, tab, 1, None, 0, None, 0, nb, s_data.shape[1])
        fd, sd, ssd = first.numpy(), second.numpy(), s_data.numpy()
        for c in range(nb):
            tr, tc = Table(_row(_np(F.rt), c), F.M), Table(_row(_np(F.ct), c), F.N)
            tb = Table(_row(tab.numpy(), c), kd)
            m1, m2, m3 = Match(_row(m_first.numpy(), c)), Match(_row(m_second.numpy(), c)), Match(_row(m_s.numpy(), c))
            for x in per_chain[c]:
                k = x["keep"]
                if k == 0:
                    continue
                i, j = x["i"], x["j"]
                ib = tb.find(x["lam"])
                if int(m1.mcol[i]) < 0 or int(m2.mcol[ib]) < 0:
                    continue                     # (dropped by a capacity check)
                assert ib >= 0 and tb.count(ib) == k and int(m1.mcol[i]) == ib, "rt_factor: first factor layout"
                _put(fd[c], int(m1.moff[i]), x["first"][:, :k])
                assert int(m2.mcol[ib]) == j, "rt_factor: second factor layout"
                _put(sd[c], int(m2.moff[ib]), x["second"][:k, :])
                if kind == "svd":
                    assert int(m3.mcol[ib]) == ib
                    _put(ssd[c], int(m3.moff[ib]), np.diag(x["s"][:k]))
        out["first"] = (m_first, first)
        out["second"] = (m_second, second)
        if kind == "svd":
            out["s"] = (m_s, s_data)
        return out

    # ---- elementwise on the stored sectors ---------------------------------------------------------------------------
    def rt_scale(self, data, match, vec, op):
        self.launches += 1
        D, Mt, v = _np(data), _np(match), _np(vec).reshape(-1)
        nb = max(D.shape[0], Mt.shape[0], v.shape[0])
        out = self.rt_alloc(nb, D.shape[1])
        o = out.numpy()
        for c in range(nb):
            size = int(_row(Mt, c)[0])
            a = float(v[c if v.shape[0] > 1 else 0])
            o[c, :size] = _row(D, c)[:size] * a if op == 0 else _row(D, c)[:size] / a
        return out

    def rt_binary(self, a, b, match, op):
        self.launches += 1
        A, Bm, Mt = _np(a), _np(b), _np(match)
        nb = max(A.shape[0], Bm.shape[0], Mt.shape[0])
        out = self.rt_alloc(nb, A.shape[1])
        o = out.numpy()
        for c in range(nb):
            size = int(_row(Mt, c)[0])
            x, y = _row(A, c)[:size], _row(Bm, c)[:size]
            o[c, :size] = [x + y, x - y, x * y, x / np.where(y == 0, 1, y)][op]
        return out

    def rt_norm(self, data, match, kind):
        self.launches += 1
        D, Mt = _np(data), _np(match)
        nb = max(D.shape[0], Mt.shape[0])
        out = np.zeros(nb)
        for c in range(nb):
            size = int(_row(Mt, c)[0])
            x = _row(D, c)[:size]
            if np.isnan(x).any():
                raise RuntimeError("rt_norm: an element of a stored sector was never written")
            out[c] = (np.abs(x).max() if len(x) else 0.0) if kind == -1 else (np.abs(x).sum() if kind == 1 else np.sqrt((x * x).sum()))
        return torch.from_numpy(out)

    def rt_scalar(self, data, match):
        self.launches += 1
        D, Mt = _np(data), _np(match)
        nb = max(D.shape[0], Mt.shape[0])
        out = np.zeros(nb)
        for c in range(nb):
            out[c] = _row(D, c)[0] if int(_row(Mt, c)[0]) > 0 else 0.0
        return torch.from_numpy(out)

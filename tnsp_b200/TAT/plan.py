"""Host planners: turn an edge operation on block-symmetric tensors into descriptor lists.

Everything here is integer bookkeeping (sector pairing, offsets, fermi signs) and is cached by
(structure, arguments); the descriptor arrays are consumed by the CUDA kernels behind the C-ABI
(``include/tnsp_b200.h``).  No tensor data is touched.

Reference behaviour reproduced (bit-exact results of the integer rules, own data structures):

* ``edge_operator_plan``  -- TAT/include/TAT/implement/edge_operator.hpp:34-692
  (cut -> split -> reverse -> transpose -> reverse -> merge in one pass, with fermi signs)
* ``contract_plan``       -- TAT/include/TAT/implement/contract.hpp:306-620 (symmetric) and
  :622-857 (no symmetry, optional fused batch edge)
* ``svd_plan`` / ``qr_plan`` -- implement/svd.hpp:259-538, implement/qr.hpp:309-508
* ``conjugate_signs``     -- implement/conjugate.hpp:31-119
"""
from __future__ import annotations

import itertools

import numpy as np

from .structure import BlockTable, block_table

MAX_RANK = 8  # rank limit of one pack descriptor after coalescing (device struct size)

# internal names (never visible to users)
C0, C1, C2 = "\x00Contract_0", "\x00Contract_1", "\x00Contract_2"
SVD_U, SVD_V = "\x00SVD_U", "\x00SVD_V"
QR_1, QR_2 = "\x00QR_1", "\x00QR_2"


class PackPlan:
    """Descriptor list of one edge operation.

    desc : int64 [n, 3 + 3*MAX_RANK]  rows = (src_off, dst_off, sign|rank<<1, dims[R], sstr[R], dstr[R])
           dims are listed slowest-first in destination order and padded with 1 / 0 strides.
    estart : int64 [n+1]  prefix sum of elements per descriptor
    identity : the operation is a plain copy of the whole storage (can alias)
    """

    __slots__ = ("names", "edges", "table", "desc", "estart", "total", "identity", "src_size", "dst_size", "covers_all", "_dev")

    def __init__(self, names, edges, table, rows, src_size):
        self.names = names
        self.edges = edges
        self.table = table
        self.src_size = src_size
        self.dst_size = table.size
        n = len(rows)
        desc = np.zeros((n, 3 + 3 * MAX_RANK), dtype=np.int64)
        counts = np.zeros(n + 1, dtype=np.int64)
        for j, (so, do, sign, dims, sstr, dstr) in enumerate(rows):
            r = len(dims)
            if r > MAX_RANK:
                raise RuntimeError(f"pack descriptor rank {r} exceeds {MAX_RANK} after coalescing")
            desc[j, 0] = so
            desc[j, 1] = do
            desc[j, 2] = (r << 1) | int(sign)
            desc[j, 3:3 + MAX_RANK] = 1
            desc[j, 3:3 + r] = dims
            desc[j, 3 + MAX_RANK:3 + MAX_RANK + r] = sstr
            desc[j, 3 + 2 * MAX_RANK:3 + 2 * MAX_RANK + r] = dstr
            counts[j + 1] = int(np.prod(dims)) if r else 1
        self.desc = desc
        self.estart = np.cumsum(counts)
        self.total = int(self.estart[-1])
        self.covers_all = self.total == self.dst_size
        self.identity = False
        if self.src_size == self.dst_size == self.total:
            ok = True
            for so, do, sign, dims, sstr, dstr in rows:
                if sign or so != do or len(dims) > 1 or (len(dims) == 1 and (sstr[0] != 1 or dstr[0] != 1)):
                    ok = False
                    break
            self.identity = ok
        self._dev = None


def _coalesce(dims, sstr, dstr):
    """Drop unit dims and fuse neighbouring axes that are contiguous in both source and destination."""
    d, s, t = [], [], []
    for a, b, c in zip(dims, sstr, dstr):
        if a == 1:
            continue
        if d and s[-1] == b * a and t[-1] == c * a:
            d[-1] *= a
            s[-1] = b
            t[-1] = c
        else:
            d.append(a)
            s.append(b)
            t.append(c)
    return d, s, t


def _group_table(S, group_edges):
    """Enumerate constituent position tuples of a merge/split group row-major.

    Returns list of (total_symmetry, size, n_odd) in iteration order and the shape.
    (edge_operator.hpp:150-172 / :361-379)
    """
    shape = [e.segments_size for e in group_edges]
    out = []
    for pos in itertools.product(*[range(n) for n in shape]):
        sym = S()
        size = 1
        odd = 0
        for e, p in zip(group_edges, pos):
            s, d = e.segments[p]
            sym = sym + s
            size *= d
            odd += 1 if s.parity else 0
        out.append((sym, size, odd))
    return out, shape


def edge_operator_plan(
    EdgeT,
    names,
    edges,
    split_map,
    reversed_names,
    merge_map,
    new_names,
    apply_parity=False,
    excl_split=frozenset(),
    excl_rev_before=frozenset(),
    excl_rev_after=frozenset(),
    excl_merge=frozenset(),
    cut_map=None,
):
    """Plan one generalised edge operation. All arguments are hashable-free python containers.

    split_map : {name: [(new_name, segments-tuple), ...]}
    merge_map : {new_name: [old names...]}
    cut_map   : {name: {symmetry: new_dim}}
    """
    S = EdgeT.Symmetry
    is_fermi = S.is_fermi_symmetry
    rank0 = len(names)
    src_table = block_table(edges)

    # step 1: cut (edge_operator.hpp:92-125)
    if cut_map:
        edges_bs = []
        for n, e in zip(names, edges):
            cm = cut_map.get(n)
            if cm is None:
                edges_bs.append(e)
            else:
                segs = []
                for s, d in e.segments:
                    if s in cm:
                        nd = cm[s]
                        if nd != 0:
                            segs.append((s, min(nd, d)))
                    else:
                        segs.append((s, d))
                edges_bs.append(EdgeT(tuple(segs), e.arrow))
    else:
        edges_bs = list(edges)

    # step 2: split (edge_operator.hpp:127-205)
    names_as, edges_as, split_flag = [], [], []
    split_tables = {}  # index_before_split -> (dict pos tuple -> (position, offset, n_odd), count)
    for i, (n, e) in enumerate(zip(names, edges_bs)):
        sp = split_map.get(n) if split_map else None
        if sp is not None:
            group = []
            for new_name, segs in sp:
                names_as.append(new_name)
                split_flag.append(i)
                ne = EdgeT(segs, e.arrow)
                edges_as.append(ne)
                group.append(ne)
            table, shape = _group_table(S, group)
            bank = [0] * e.segments_size
            tab = {}
            for pos, (sym, size, odd) in zip(itertools.product(*[range(k) for k in shape]), table):
                p = e.find_by_symmetry(sym)
                if p is None:
                    # a constituent combination that does not exist in the edge being split: never addressed
                    # by a zero-total-symmetry block unless the user supplied an inconsistent split
                    tab[pos] = None
                    continue
                tab[pos] = (p, bank[p], odd)
                bank[p] += size
            split_tables[i] = (tab, len(group))
        else:
            names_as.append(n)
            split_flag.append(i)
            edges_as.append(e)
    rank_t = len(names_as)

    # step 3: reverse before transpose (:211-232)
    rev_before = [False] * rank_t
    edges_bt = list(edges_as)
    if is_fermi and reversed_names:
        for j, n in enumerate(names_as):
            if n in reversed_names:
                rev_before[j] = True
                edges_bt[j] = edges_as[j].reversed()

    # names before merge (:234-262)
    rank_m = len(new_names)
    names_bm, merge_flag = [], []
    for k, n in enumerate(new_names):
        if merge_map and n in merge_map:
            for mn in merge_map[n]:
                names_bm.append(mn)
                merge_flag.append(k)
        else:
            names_bm.append(n)
            merge_flag.append(k)
    if len(names_bm) != rank_t:
        raise RuntimeError("Tensor to transpose with Different Rank")

    # step 4: transpose plan (:270-292)
    where = {n: j for j, n in enumerate(names_as)}
    if len(where) != rank_t:
        raise RuntimeError("Duplicated names in edge operation")
    d2s = []
    for n in names_bm:
        if n not in where:
            raise RuntimeError("Tensor to transpose with incompatible name list")
        d2s.append(where[n])
    edges_at = [edges_bt[j] for j in d2s]

    # step 5+6: reverse before merge and merge (:294-417)
    edges_bm = list(edges_at)
    rev_after = [False] * rank_t
    result_edges = []
    merge_tables = {}
    start = 0
    for k in range(rank_m):
        end = start
        while end < rank_t and merge_flag[end] == k:
            end += 1
        if is_fermi:
            arrow = False if start == end else edges_bm[start].arrow
            for j in range(start, end):
                if edges_bm[j].arrow != arrow:
                    edges_bm[j] = edges_bm[j].reversed()
                    rev_after[j] = True
        else:
            arrow = False
        if end != start + 1:
            group = edges_bm[start:end]
            table, shape = _group_table(S, group)
            merged = []  # [sym, dim]
            tab = {}
            for pos, (sym, size, odd) in zip(itertools.product(*[range(q) for q in shape]), table):
                p = None
                for q, (ms, _) in enumerate(merged):
                    if tuple.__eq__(ms, sym):
                        p = q
                        break
                if p is None:
                    merged.append([sym, 0])
                    p = len(merged) - 1
                tab[pos] = (p, merged[p][1], odd)
                merged[p][1] += size
            merge_tables[k] = tab
            result_edges.append(EdgeT(tuple((s, d) for s, d in merged), arrow))
        else:
            result_edges.append(edges_bm[start])
        start = end
    result_edges = tuple(result_edges)
    dst_table = block_table(result_edges)

    # marks (:430-484)
    if is_fermi:
        if apply_parity:
            split_mark = [n not in excl_split for n in names]
            rb_mark = [n not in excl_rev_before for n in names_as]
            ra_mark = [n not in excl_rev_after for n in names_bm]
            merge_mark = [n not in excl_merge for n in new_names]
        else:
            split_mark = [n in excl_split for n in names]
            rb_mark = [n in excl_rev_before for n in names_as]
            ra_mark = [n in excl_rev_after for n in names_bm]
            merge_mark = [n in excl_merge for n in new_names]
        rev_flag = [(rev_after[j] and ra_mark[j]) != (rev_before[d2s[j]] and rb_mark[d2s[j]]) for j in range(rank_t)]
        swapped_pairs = [(a, b) for a in range(rank_t) for b in range(a + 1, rank_t) if d2s[a] > d2s[b]]

    # main loop over fine blocks (:486-690)
    fine_table = BlockTable(tuple(edges_bm)) if rank_t else None
    rows = []
    if rank_t == 0:
        fine_positions = [()]
    else:
        fine_positions = [tuple(int(x) for x in p) for p in fine_table.positions]
    src_edges = edges  # original (uncut) edges give source leadings
    bm_dims = [e.dims for e in edges_bm]
    bm_par = [e.parities for e in edges_bm]
    res_dims = [e.dims for e in result_edges]
    src_dims = [e.dims for e in src_edges]
    for pos_bm in fine_positions:
        dims_bm = [bm_dims[j][pos_bm[j]] for j in range(rank_t)]
        parity = False
        if is_fermi:
            par = [bm_par[j][pos_bm[j]] for j in range(rank_t)]
            for j in range(rank_t):
                if rev_flag[j] and par[j]:
                    parity = not parity
            for a, b in swapped_pairs:
                if par[a] and par[b]:
                    parity = not parity
        # after merge
        pos_am, off_am = [0] * rank_m, [0] * rank_m
        j = 0
        for k in range(rank_m):
            b = j
            while j < rank_t and merge_flag[j] == k:
                j += 1
            if j != b + 1:
                p, o, odd = merge_tables[k][tuple(pos_bm[b:j])]
                pos_am[k], off_am[k] = p, o
                if is_fermi and (odd & 2) and merge_mark[k]:
                    parity = not parity
            else:
                pos_am[k] = pos_bm[b]
        # before split
        pos_as = [0] * rank_t
        dims_as = [0] * rank_t
        for j in range(rank_t):
            pos_as[d2s[j]] = pos_bm[j]
            dims_as[d2s[j]] = dims_bm[j]
        pos_bs, off_bs = [0] * rank0, [0] * rank0
        j = 0
        bad = False
        for i in range(rank0):
            b = j
            while j < rank_t and split_flag[j] == i:
                j += 1
            if i in split_tables:
                ent = split_tables[i][0][tuple(pos_as[b:j])]
                if ent is None:
                    bad = True
                    break
                p, o, odd = ent
                pos_bs[i], off_bs[i] = p, o
                if is_fermi and (odd & 2) and split_mark[i]:
                    parity = not parity
            else:
                pos_bs[i] = pos_as[b]
        if bad:
            raise RuntimeError("Inconsistent split plan: split segments do not exist in the original edge")
        if cut_map:
            pos_bc = []
            for i in range(rank0):
                sym = edges_bs[i].segments[pos_bs[i]][0]
                pos_bc.append(src_edges[i].find_by_symmetry(sym))
        else:
            pos_bc = pos_bs
        # leadings
        lead_am = [1] * rank_m
        for k in range(rank_m - 2, -1, -1):
            lead_am[k] = lead_am[k + 1] * res_dims[k + 1][pos_am[k + 1]]
        lead_bm = [0] * rank_t
        for j in range(rank_t - 1, -1, -1):
            if j != rank_t - 1 and merge_flag[j] == merge_flag[j + 1]:
                lead_bm[j] = lead_bm[j + 1] * dims_bm[j + 1]
            else:
                lead_bm[j] = lead_am[merge_flag[j]]
        lead_bs = [1] * rank0
        for i in range(rank0 - 2, -1, -1):
            lead_bs[i] = lead_bs[i + 1] * src_dims[i + 1][pos_bc[i + 1]]
        lead_as = [0] * rank_t
        for j in range(rank_t - 1, -1, -1):
            if j != rank_t - 1 and split_flag[j] == split_flag[j + 1]:
                lead_as[j] = lead_as[j + 1] * dims_as[j + 1]
            else:
                lead_as[j] = lead_bs[split_flag[j]]
        sb = src_table.block_by_positions(pos_bc)
        db = dst_table.block_by_positions(pos_am)
        if sb is None or db is None:
            raise RuntimeError("edge operation addresses a block that does not exist")
        so = int(src_table.offsets[sb]) + sum(o * l for o, l in zip(off_bs, lead_bs))
        do = int(dst_table.offsets[db]) + sum(o * l for o, l in zip(off_am, lead_am))
        if 0 in dims_bm:
            continue
        sstr = [lead_as[d2s[j]] for j in range(rank_t)]
        d, s, t = _coalesce(dims_bm, sstr, lead_bm)
        rows.append((so, do, parity, d, s, t))
    return PackPlan(tuple(new_names), result_edges, dst_table, rows, src_table.size)


# ------------------------------------------------------------------------------------------------
# contract
# ------------------------------------------------------------------------------------------------
class ContractPlan:
    __slots__ = ("pack1", "pack2", "unpack", "gemm", "prod_size", "zero_fill", "names", "edges", "table", "flops", "m1_size", "m2_size", "fuse_l", "_dev",
                 "gather", "_gdev")


def _common_order(names_1, names_2, map12, map21, free_1, free_2, size_1, size_2):
    """Operand layout heuristic (contract.hpp:414-474). Returns (common_1, common_2, right_1, right_2)."""
    common_1, common_2 = [], []

    def fit_1():
        for n in names_1:
            if n in map12:
                common_1.append(n)
                common_2.append(map12[n])

    def fit_2():
        for n in names_2:
            if n in map21:
                common_2.append(n)
                common_1.append(map21[n])

    last_1, last_2 = (names_1[-1] if names_1 else None), (names_2[-1] if names_2 else None)
    if len(free_1) == 0:
        r1 = True
        fit_2()
        r2 = (not common_2) or common_2[-1] == last_2
    elif len(free_2) == 0:
        r2 = True
        fit_1()
        r1 = (not common_1) or common_1[-1] == last_1
    elif size_1 > size_2:
        if free_1[-1] != last_1:
            r1 = True
            fit_1()
            r2 = (not common_2) or common_2[-1] == last_2
        elif free_2[-1] != last_2:
            r2 = True
            fit_2()
            r1 = (not common_1) or common_1[-1] == last_1
        else:
            r1 = r2 = False
            fit_1()
    else:
        if free_2[-1] != last_2:
            r2 = True
            fit_2()
            r1 = (not common_1) or common_1[-1] == last_1
        elif free_1[-1] != last_1:
            r1 = True
            fit_1()
            r2 = (not common_2) or common_2[-1] == last_2
        else:
            r1 = r2 = False
            fit_2()
    return common_1, common_2, r1, r2


GATHER_MIN_M = 17   # up to 16 rows the one-warp-per-matrix / small-tile kernels (on packed operands) are used


def _rowstream_fits(n, k):
    """mirror of rowstream_shape() in csrc/gemm_gather.cu: one n-pass (<= 64 columns) of the zero-padded B operand of a
    chain must fit in 104 KiB of shared memory (wider B operands are sliced over the grid)"""
    if n <= 64:
        nt = (n + 7) // 8
    else:
        nt = min(range(8, 4, -1), key=lambda t: (((n + 8 * t - 1) // (8 * t)) * 8 * t - n, -t))
    return ((k + 3) // 4 * 4) * (8 * nt + 4) * 8 <= 104 * 1024


def _gather_tables(names_1, edges_1, names_2, edges_2, free_1, common_1, free_2, common_2, m, n, k):
    """Offset tables that let the GEMM read both operands of a dense contraction in place (csrc/gemm_gather.cu):
    element (r, kk) of the merged matrix of tensor 1 is data[row_off[r] + col_off[kk]], where r / kk enumerate the
    free / common edges row-major in the reference's merge order (edge_operator.hpp:321-404 with one segment per
    edge; contract.hpp:622-857).  Returns (int32 table aro|aco|bro|bco, flags) or None when the gather path does
    not apply (tiny or empty problems, offsets beyond int32)."""
    if m < GATHER_MIN_M or n == 0 or k == 0 or not _rowstream_fits(n, k):
        return None

    def strides(edges):
        st, acc = [], 1
        for e in reversed(edges):
            st.append(acc)
            acc *= e.dimension
        return list(reversed(st)), acc

    def offsets(names, edges, stride, group):
        o = np.zeros(1, dtype=np.int64)
        for nm in group:
            i = names.index(nm)
            o = (o[:, None] + np.arange(edges[i].dimension, dtype=np.int64)[None, :] * stride[i]).reshape(-1)
        return o

    s1, size1 = strides(edges_1)
    s2, size2 = strides(edges_2)
    if max(size1, size2) >= 2**31:
        return None
    aro = offsets(names_1, edges_1, s1, free_1)
    aco = offsets(names_1, edges_1, s1, common_1)
    bro = offsets(names_2, edges_2, s2, common_2)
    bco = offsets(names_2, edges_2, s2, free_2)
    assert len(aro) == m and len(aco) == k and len(bro) == k and len(bco) == n
    # lanes run along the direction in which the operand is contiguous in memory: the group holding the last edge
    flags = (1 if names_1 and names_1[-1] in common_1 else 0) | (2 if names_2 and names_2[-1] in free_2 else 0)
    return np.concatenate([aro, aco, bro, bco]).astype(np.int32), flags, m, n, k


# gemm descriptor row: (m, n, k, a_off, b_off, c_off, flags, alpha_sign)
#   flags bit0: A stored [k x m] (else [m x k]);  bit1: B stored [n x k] (else [k x n]); row-major, dense
GEMM_COLS = 8


def contract_plan(EdgeT, names_1, edges_1, names_2, edges_2, pairs, fuse_names=()):
    S = EdgeT.Symmetry
    is_fermi = S.is_fermi_symmetry
    map12 = {a: b for a, b in pairs}
    map21 = {b: a for a, b in pairs}
    if len(map12) != len(pairs) or len(map21) != len(pairs):
        raise RuntimeError("Duplicated names in contract pairs")
    for a, b in pairs:
        if a not in names_1 or b not in names_2:
            raise RuntimeError("Missing name in contract")
    t1, t2 = block_table(edges_1), block_table(edges_2)
    plan = ContractPlan()
    plan._dev = None
    plan._gdev = None
    plan.gather = None
    fuse = [n for n in names_1 if n in fuse_names] if fuse_names else []
    if fuse and S.length != 0:
        raise RuntimeError("fuse_names is only supported for tensors without symmetry")
    free_1 = [n for n in names_1 if n not in map12 and n not in fuse]
    free_2 = [n for n in names_2 if n not in map21 and n not in fuse]
    e1 = dict(zip(names_1, edges_1))
    e2 = dict(zip(names_2, edges_2))
    for a, b in pairs:
        if S.length == 0:
            if e1[a].dimension != e2[b].dimension:
                raise RuntimeError("Contracting two edge with different dimension")
        elif e1[a].conjugate() != e2[b]:
            raise RuntimeError("Incompatible edge segments in contract")
    common_1, common_2, r1, r2 = _common_order(names_1, names_2, map12, map21, free_1, free_2, t1.size, t2.size)

    if S.length == 0:
        # contract.hpp:622-857
        names_res = list(fuse) + free_1 + free_2
        edges_res = [e1[n] for n in fuse] + [e1[n] for n in free_1] + [e2[n] for n in free_2]
        for n in fuse:
            if e1[n] != e2[n]:
                raise RuntimeError("Cannot fuse two edge with different shape")
        order1 = [C0, C1, C2] if r1 else [C0, C2, C1]
        order2 = [C0, C2, C1] if r2 else [C0, C1, C2]
        plan.pack1 = edge_operator_plan(EdgeT, names_1, edges_1, None, None, {C1: free_1, C2: common_1, C0: fuse}, order1)
        plan.pack2 = edge_operator_plan(EdgeT, names_2, edges_2, None, None, {C2: free_2, C1: common_2, C0: fuse}, order2)
        l = plan.pack1.edges[0].dimension
        m = plan.pack1.edges[1 if r1 else 2].dimension
        k = plan.pack1.edges[2 if r1 else 1].dimension
        n = plan.pack2.edges[1 if r2 else 2].dimension
        plan.names = tuple(names_res)
        plan.edges = tuple(edges_res)
        plan.table = block_table(plan.edges)
        rows = []
        flags = (0 if r1 else 1) | (2 if r2 else 0)
        if m and n and k:
            for i in range(l):
                rows.append((m, n, k, i * m * k, i * k * n, i * m * n, flags, 1))
        plan.gemm = np.array(rows, dtype=np.int64).reshape(-1, GEMM_COLS)
        plan.prod_size = plan.table.size
        plan.zero_fill = bool(m and n and not k)
        plan.unpack = None
        plan.fuse_l = l
        plan.gather = _gather_tables(names_1, edges_1, names_2, edges_2, free_1, common_1, free_2, common_2, m, n, k) if l == 1 and not fuse else None
    else:
        # contract.hpp:306-620
        rev_1, rev_2, rev_res, common_rev_1 = set(), set(), set(), set()
        split_1, split_2 = [], []
        names_res = []
        for n in names_1:
            e = e1[n]
            if n not in map12:
                split_1.append((n, e.segments))
                names_res.append(n)
                if is_fermi and e.arrow:
                    rev_1.add(n)
                    rev_res.add(n)
            elif is_fermi and not e.arrow:
                rev_1.add(n)
                common_rev_1.add(n)
        for n in names_2:
            e = e2[n]
            if n not in map21:
                split_2.append((n, e.segments))
                names_res.append(n)
                if is_fermi and e.arrow:
                    rev_2.add(n)
                    rev_res.add(n)
            elif is_fermi and e.arrow:
                rev_2.add(n)
        plan.pack1 = edge_operator_plan(
            EdgeT, names_1, edges_1, None, rev_1, {C1: free_1, C2: common_1}, [C1, C2] if r1 else [C2, C1],
            False, (), common_rev_1, (), {C2})
        plan.pack2 = edge_operator_plan(
            EdgeT, names_2, edges_2, None, rev_2, {C2: free_2, C1: common_2}, [C2, C1] if r2 else [C1, C2])
        m1, m2 = plan.pack1, plan.pack2
        edge_0 = m1.edges[0 if r1 else 1]
        edge_1 = m2.edges[0 if r2 else 1]
        edge_c2 = m2.edges[1 if r2 else 0]
        prod_edges = (edge_0, edge_1)
        prod_table = block_table(prod_edges)
        rows = []
        covered = 0
        for p0, (sym, m) in enumerate(edge_0.segments):
            if edge_0.find_by_symmetry(sym) != p0:
                continue  # duplicated symmetry: handled by its first appearance (contract.hpp:541-545)
            p1 = edge_1.find_by_symmetry(-sym)
            if p1 is None:
                continue
            n = edge_1.segments[p1][1]
            pc = edge_c2.find_by_symmetry(sym)
            if pc is None:
                continue
            k = edge_c2.segments[pc][1]
            if not (m and n and k):
                continue
            b1 = m1.table.block_by_positions((p0, pc) if r1 else (pc, p0))
            b2 = m2.table.block_by_positions((p1, pc) if r2 else (pc, p1))
            bc = prod_table.block_by_positions((p0, p1))
            alpha = -1 if (is_fermi and (r1 != (not r2)) and sym.parity) else 1
            flags = (0 if r1 else 1) | (2 if r2 else 0)
            rows.append((m, n, k, int(m1.table.offsets[b1]), int(m2.table.offsets[b2]), int(prod_table.offsets[bc]), flags, alpha))
            covered += m * n
        plan.gemm = np.array(rows, dtype=np.int64).reshape(-1, GEMM_COLS)
        plan.prod_size = prod_table.size
        plan.zero_fill = covered != prod_table.size
        plan.unpack = edge_operator_plan(
            EdgeT, (C1, C2), prod_edges, {C1: split_1, C2: split_2}, rev_res, None, names_res)
        plan.names = plan.unpack.names
        plan.edges = plan.unpack.edges
        plan.table = plan.unpack.table
        plan.fuse_l = 1
    plan.m1_size = plan.pack1.dst_size
    plan.m2_size = plan.pack2.dst_size
    plan.flops = int(sum(2 * r[0] * r[1] * r[2] for r in plan.gemm))
    return plan


# ------------------------------------------------------------------------------------------------
# svd / qr
# ------------------------------------------------------------------------------------------------
class FactorPlan:
    """Shared plan of the two matrix factorizations.

    sectors : int64 [ns, 8] rows = (m, n, k, a_off, out1_off, out2_off, s_off, spare) on the merged
              matrix (row-major m x n), out1 = m x k, out2 = k x n.
    """
    __slots__ = ("merge", "sectors", "t1_names", "t1_edges", "t1_table", "t2_names", "t2_edges", "t2_table",
                 "s_syms", "s_total", "flag", "extra", "_dev", "rc_tab", "_rcdev")


def _merge_offsets(EdgeT, names, edges, rows_group, cols_group, merge):
    """int32 table ro[m] | co[n]: element (i, j) of the merged matrix of a DENSE tensor is data[ro[i] + co[j]] (rows /
    columns enumerate the two name groups row-major, the merge order of edge_operator.hpp:321-404 with one segment
    per edge), so that qr / svd can read their operand in place (csrc/factor_sector.cu).  None when not applicable."""
    if EdgeT.Symmetry.length != 0 or merge.identity:
        return None
    st, acc = [], 1
    for e in reversed(edges):
        st.append(acc)
        acc *= e.dimension
    st.reverse()
    if acc >= 2**31 or acc == 0:
        return None

    def offsets(group):
        o = np.zeros(1, dtype=np.int64)
        for nm in group:
            i = names.index(nm)
            o = (o[:, None] + np.arange(edges[i].dimension, dtype=np.int64)[None, :] * st[i]).reshape(-1)
        return o

    return np.concatenate([offsets(rows_group), offsets(cols_group)]).astype(np.int32)


def _factor_common(EdgeT, merged):
    """Common edge of the two factors (svd.hpp:380-393, qr.hpp:415-429) and the per-sector list."""
    edge_0, edge_1 = merged.edges
    seg_1, seg_2 = [], []
    for s0, d0 in edge_0.segments:
        p = edge_1.find_by_symmetry(-s0)
        if p is not None:
            k = min(d0, edge_1.segments[p][1])
            seg_1.append((-s0, k))
            seg_2.append((s0, k))
    return EdgeT(tuple(seg_1), False), EdgeT(tuple(seg_2), False)


def _factor_sectors(merged, t1_table, t2_table, common_2):
    edge_0, edge_1 = merged.edges
    rows = []
    s_syms = []
    s_off = 0
    for sym, _ in edge_0.segments:
        p0 = edge_0.find_by_symmetry(sym)
        p1 = edge_1.find_by_symmetry(-sym)
        if p1 is None:
            continue
        pc = common_2.find_by_symmetry(sym)
        m = edge_0.segments[p0][1]
        n = edge_1.segments[p1][1]
        k = common_2.segments[pc][1]
        ba = merged.table.block_by_positions((p0, p1))
        b1 = t1_table.block_by_positions((p0, pc))
        b2 = t2_table.block_by_positions((pc, p1))
        rows.append((m, n, k, int(merged.table.offsets[ba]), int(t1_table.offsets[b1]), int(t2_table.offsets[b2]), s_off, 0))
        s_syms.append(sym)
        s_off += k
    return np.array(rows, dtype=np.int64).reshape(-1, 8), s_syms, s_off


def svd_plan(EdgeT, names, edges, free_names_u, common_name_u, common_name_v):
    S = EdgeT.Symmetry
    is_fermi = S.is_fermi_symmetry
    for n in free_names_u:
        if n not in names:
            raise RuntimeError("Missing name in svd")
    put_v_right = (not names) or (names[-1] not in free_names_u)
    list_u, list_v = [], []
    rev_in, rev_u, rev_v = set(), set(), set()
    res_u, res_v = [], []
    fe_u, fe_v = [], []
    if put_v_right:
        res_v.append(common_name_v)
    else:
        res_u.append(common_name_u)
    for n, e in zip(names, edges):
        if n in free_names_u:
            list_u.append(n)
            res_u.append(n)
            fe_u.append((n, e.segments))
            if is_fermi and e.arrow:
                rev_u.add(n)
                rev_in.add(n)
        else:
            list_v.append(n)
            res_v.append(n)
            fe_v.append((n, e.segments))
            if is_fermi and e.arrow:
                rev_v.add(n)
                rev_in.add(n)
    if put_v_right:
        res_u.append(common_name_u)
    else:
        res_v.append(common_name_v)
    p = FactorPlan()
    p._dev = None
    p.merge = edge_operator_plan(EdgeT, names, edges, None, rev_in, {SVD_U: list_u, SVD_V: list_v},
                                 [SVD_U, SVD_V] if put_v_right else [SVD_V, SVD_U])
    p._rcdev = None
    p.rc_tab = _merge_offsets(EdgeT, names, edges, list_u if put_v_right else list_v, list_v if put_v_right else list_u, p.merge)
    c1, c2 = _factor_common(EdgeT, p.merge)
    p.t1_names = (SVD_U, common_name_u) if put_v_right else (SVD_V, common_name_v)
    p.t1_edges = (p.merge.edges[0], c1)
    p.t2_names = (common_name_v, SVD_V) if put_v_right else (common_name_u, SVD_U)
    p.t2_edges = (c2, p.merge.edges[1])
    p.t1_table = block_table(p.t1_edges)
    p.t2_table = block_table(p.t2_edges)
    p.sectors, syms, p.s_total = _factor_sectors(p.merge, p.t1_table, p.t2_table, c2)
    # symmetry of s is always the one of the common edge of tensor U (svd.hpp:426)
    p.s_syms = [(-s if put_v_right else s) for s in syms]
    p.flag = put_v_right
    if is_fermi:
        rev_u = rev_u | {common_name_u}
    p.extra = dict(res_u=tuple(res_u), res_v=tuple(res_v), fe_u=fe_u, fe_v=fe_v, rev_u=frozenset(rev_u), rev_v=frozenset(rev_v))
    return p


def svd_split_plans(EdgeT, p, common_name_u, common_name_v, remain):
    """Second half of svd once the per-sector kept counts are known (svd.hpp:488-531).

    remain : list of kept dimension per sector (aligned with p.s_syms)
    returns (plan_u, plan_v, s_edges, s_blocks) where s_blocks = [(sector index, kept, sign)]
    """
    put_v_right = p.flag
    rem_u = {s: r for s, r in zip(p.s_syms, remain)}
    rem_v = {-s: r for s, r in zip(p.s_syms, remain)}
    if put_v_right:
        tu = (p.t1_names, p.t1_edges)
        tv = (p.t2_names, p.t2_edges)
    else:
        tu = (p.t2_names, p.t2_edges)
        tv = (p.t1_names, p.t1_edges)
    ex = p.extra
    plan_u = edge_operator_plan(EdgeT, tu[0], tu[1], {SVD_U: ex["fe_u"]}, ex["rev_u"], None, ex["res_u"],
                                False, (), (), (), (), {common_name_u: rem_u})
    plan_v = edge_operator_plan(EdgeT, tv[0], tv[1], {SVD_V: ex["fe_v"]}, ex["rev_v"], None, ex["res_v"],
                                False, (), (), (), (), {common_name_v: rem_v})
    S = EdgeT.Symmetry
    seg_u, seg_v, blocks = [], [], []
    for i, (s, r) in enumerate(zip(p.s_syms, remain)):
        if r == 0:
            continue
        seg_u.append((-s, r))
        seg_v.append((s, r))
        sign = bool(S.is_fermi_symmetry and (not put_v_right) and s.parity)
        blocks.append((i, r, sign))
    s_edges = (EdgeT(tuple(seg_u), False), EdgeT(tuple(seg_v), True))
    return plan_u, plan_v, s_edges, blocks


def qr_plan(EdgeT, names, edges, direction, free_names, common_name_q, common_name_r):
    S = EdgeT.Symmetry
    is_fermi = S.is_fermi_symmetry
    for n in free_names:
        if n not in names:
            raise RuntimeError("Missing name in qr")
    if direction in ("r", "R"):
        use_r_name = True
    elif direction in ("q", "Q"):
        use_r_name = False
    else:
        raise RuntimeError("Invalid direction in QR")
    use_qr = (not names) or ((names[-1] in free_names) == use_r_name)
    list_1, list_2 = [], []
    rev_in, rev_1, rev_2 = set(), set(), set()
    res_1, res_2 = [], [common_name_r if use_qr else common_name_q]
    fe_1, fe_2 = [], []
    for n, e in zip(names, edges):
        if ((n in free_names) == use_r_name) == use_qr:
            list_2.append(n)
            res_2.append(n)
            fe_2.append((n, e.segments))
            if is_fermi and e.arrow:
                rev_2.add(n)
                rev_in.add(n)
        else:
            list_1.append(n)
            res_1.append(n)
            fe_1.append((n, e.segments))
            if is_fermi and e.arrow:
                rev_1.add(n)
                rev_in.add(n)
    res_1.append(common_name_q if use_qr else common_name_r)
    p = FactorPlan()
    p._dev = None
    p.merge = edge_operator_plan(EdgeT, names, edges, None, rev_in, {QR_1: list_1, QR_2: list_2}, [QR_1, QR_2])
    p._rcdev = None
    p.rc_tab = _merge_offsets(EdgeT, names, edges, list_1, list_2, p.merge)
    c1, c2 = _factor_common(EdgeT, p.merge)
    p.t1_names = (QR_1, common_name_q if use_qr else common_name_r)
    p.t1_edges = (p.merge.edges[0], c1)
    p.t2_names = (common_name_r if use_qr else common_name_q, QR_2)
    p.t2_edges = (c2, p.merge.edges[1])
    p.t1_table = block_table(p.t1_edges)
    p.t2_table = block_table(p.t2_edges)
    p.sectors, p.s_syms, p.s_total = _factor_sectors(p.merge, p.t1_table, p.t2_table, c2)
    p.flag = use_qr
    if is_fermi:
        (rev_1 if use_qr else rev_2).add(common_name_q)
    plan_1 = edge_operator_plan(EdgeT, p.t1_names, p.t1_edges, {QR_1: fe_1}, rev_1, None, res_1)
    plan_2 = edge_operator_plan(EdgeT, p.t2_names, p.t2_edges, {QR_2: fe_2}, rev_2, None, res_2,
                                False, (), (() if use_qr else {common_name_q}), (), ())
    p.extra = (plan_1, plan_2)
    return p


def conjugate_signs(edges, table, trivial_metric):
    """Per-block sign of conjugate() (conjugate.hpp:48-97)."""
    signs = np.zeros(len(table.positions), dtype=np.int64)
    S = edges[0].Symmetry if edges else None
    if S is None or not S.is_fermi_symmetry:
        return signs
    for b, pos in enumerate(table.positions):
        n_odd = 0
        tot = False
        for e, p in zip(edges, pos):
            if e.parities[int(p)]:
                n_odd += 1
                if e.arrow and trivial_metric:
                    tot = not tot
        signs[b] = int(tot != bool(n_odd & 2))
    return signs

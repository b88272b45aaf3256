"""Sector-compact lock-step tensors: block-symmetric tensors whose symmetry sectors differ from chain to chain.

This is the batched analogue of the reference's block-symmetric ``Tensor`` (TAT/include/TAT/structure/core.hpp:152-191)
for a lock-step batch of Monte-Carlo chains.  In a symmetric PEPS the sampled physical charges sit in the dim-1
``P_l1_l2_o`` edges of every environment tensor (tetragono/tetragono/sampling_lattice/lattice.py:319-339) and the greedy
cross-sector cut of svd.hpp:429-481 keeps a different number of singular values per sector in every chain, so the chains of a
batch never share one block structure.  Round 1 therefore ran symmetric models on zero-padded dense tensors (27x the
reference's GEMM flops).  Here every chain keeps exactly the reference's sectors:

* every index of every edge carries a per-chain integer charge label (``Edge`` below; dim-1 edges with host-known charges --
  the physical ``P`` edges and the total-symmetry edge ``T`` -- only contribute to a per-chain target charge);
* a tensor is stored as a *matrix of sectors* for one grouping (rows | cols) of its edges (``Form``): per chain, the merged row
  indices sorted by charge (``rt_sort``: the per-chain analogue of the merged edge of edge_operator.hpp:321-404), the sector
  pairing row charge + column charge = target (``rt_match``: contract.hpp:539-616 with the sample axis added), and the sector
  matrices back to back, row-major, nothing else -- no zeros are stored, moved or multiplied;
* ``contract`` = regroup both operands (``rt_repack``: the transpose / merge of edge_operator.hpp:651-688) + ONE grouped GEMM
  over (chain x sector) (``rt_gemm``: contract.hpp:582-616); ``qr`` / ``svd`` factorise every (chain, sector) matrix from a
  device-side work queue and emit the new bond with its per-chain labels (qr.hpp:419-429, svd.hpp:405-481: the greedy cut
  ranks the singular values of all sectors of a chain together);
* everything is planned ON THE DEVICE from the labels: the host never reads a size back, so a lock-step sweep issues
  kernels without synchronising.

Integer (U1-type) symmetries: BoseU1, FermiU1, FermiU1BoseU1, FermiU1FermiU1.  The host side here is bookkeeping of names
and handles only; every array operation is a kernel behind ``backend.rt_*`` (csrc/ragged.cu, csrc/factor_sector.cu).
"""
from __future__ import annotations

import weakref

import numpy as np

from .. import backend as _bk
from .tensor import BatchScalar

SMAX = 64                    # sectors per edge group and chain (device tables are sized for it)
HDR = 3 + 2 * SMAX           # group table header: nsec, nvalid, skey[SMAX], sstart[SMAX + 1]
MSTRIDE = 4 + 2 * SMAX       # match table: size, flag, moff[SMAX + 1], mcol[SMAX] (+1 pad)
DEAD = 1 << 30               # label of a bond index that carries no state (|label| >= 2^29)
STATS = {"contract": 0, "qr": 0, "svd": 0, "repack": 0, "sort": 0, "match": 0, "signed_hit": 0}

_PLANS: dict = {}

# Learnt capacities.  The stored size of a sector-compact tensor is only known on the device; allocating the dense bound
# prod(dims) for every result would keep the zero-padded footprint of round 1 (165 GB at 592 chains of cfg2).  Every operation
# signature therefore measures its largest per-chain size during a learning phase (a small calibration batch, or the first two sweeps:
# the only device -> host reads of the engine) and later allocates CAP_FACTOR x that.  A chain that still does not fit is stored EMPTY and
# counted on the device (backend.rt_overflow, checked by the samplers once per sweep): never silent, never out of bounds.
_CAPS: dict = {}
CAP_FACTOR = 2.0
CAP_MIN_CHAINS = 64          # smaller batches always allocate the dense bound: their maxima fluctuate too much and memory is no issue
CAPS_ENABLED = True
TABLE_CACHE_MAX = 4096       # merged dimension up to which a group table is shared through its label arrays
_LEARN = {"all": True, "cycles": 0}     # learning phase: every operation allocates the dense bound and records its largest size


def freeze_capacities():
    """end of the learning phase: from now on an operation signature seen before allocates CAP_FACTOR x its largest recorded size"""
    _LEARN["all"] = False


def learning_cycle_done(cycles=2):
    """called by the drivers after every sweep + observation while learning: the phase ends after `cycles` of them"""
    if _LEARN["all"]:
        _LEARN["cycles"] += 1
        if _LEARN["cycles"] >= cycles:
            freeze_capacities()


def _cap(key, dense, nb=None):
    """(elements to allocate, learning?)"""
    if not CAPS_ENABLED:
        return dense, False
    if nb is not None and nb < CAP_MIN_CHAINS and not _LEARN["all"]:
        return dense, False
    c = _CAPS.get(key)
    if _LEARN["all"] or c is None:
        return dense, True
    return min(dense, int(c * CAP_FACTOR) + 64), False


def _learn(key, match):
    B = _bk.get()
    mx = int(B.to_numpy(match[:, 0]).max()) if match.shape[0] else 0
    _CAPS[key] = max(_CAPS.get(key, 0), mx)


def pack_symmetry(sym):
    """integer label of a symmetry value: component i weighs 65536^i (sums of labels = labels of sums)"""
    if any(k != "U1" for k in type(sym).kinds):
        raise NotImplementedError("the sector-compact lock-step engine supports integer (U1-type) symmetries only")
    out = 0
    for i, v in enumerate(sym):
        out += int(v) * (65536**i)
    return out


def fermi_mask(S):
    """bit i set: component i of a packed label is fermionic (its oddness enters the parity)"""
    return sum(1 << i for i, f in enumerate(S.fermi_flags) if f)


def label_parity(labels, mask):
    """parity bit of packed labels (numpy int array) for the fermionic components in `mask`"""
    labels = np.asarray(labels, dtype=np.int64)
    par = np.zeros(labels.shape, dtype=np.int64)
    rest = labels.copy()
    i = 0
    while mask >> i:
        c = ((rest + 32768) % 65536) - 32768
        rest = (rest - c) // 65536
        if (mask >> i) & 1:
            par ^= c & 1
        i += 1
    return par


class Edge:
    """one edge of a sector-compact tensor: dimension, charge labels (device int32 [nbL, dim] with nbL in {1, nb}) and a sign
    (effective label = sign * array, so that conjugation and contraction results never copy label arrays).  A *unit* edge has
    dimension 1 and host-known labels `harr` [nbL] (physical P edges, total-symmetry edge): it never reaches a kernel."""
    __slots__ = ("dim", "arr", "sign", "arrow", "harr", "par")

    def __init__(self, dim, arr, sign=1, arrow=False, harr=None, par=None):
        self.dim, self.arr, self.sign, self.arrow, self.harr = int(dim), arr, int(sign), bool(arrow), harr
        self.par = par          # (fermi mask, parity bits of harr): cached by unit_parity, shared by flipped copies

    @property
    def unit(self):
        return self.harr is not None

    def flipped(self, s):
        return self if s == 1 else Edge(self.dim, self.arr, -self.sign, not self.arrow, self.harr, self.par)

    def unit_parity(self, mask):
        """parity bits (uint32 [nbL]) of a unit edge's charges (the parity of a label does not depend on its sign)"""
        p = self.par
        if p is None or p[0] != mask:
            p = self.par = (mask, label_parity(np.asarray(self.harr, dtype=np.int64).reshape(-1), mask).astype(np.uint32))
        return p[1]

    def host_labels(self):
        if self.harr is not None:
            return self.sign * np.asarray(self.harr, dtype=np.int64).reshape(-1, 1)
        return self.sign * _bk.get().to_numpy(self.arr).astype(np.int64)


class Form:
    """matrix-of-sectors storage for the grouping rows | cols (tuples of non-unit edge positions)"""
    __slots__ = ("rows", "cols", "rt", "rs", "ct", "cs", "match", "data", "M", "N", "_c")

    def __init__(self, rows, cols, rt, rs, ct, cs, match, data, M, N):
        self.rows, self.cols, self.rt, self.rs, self.ct, self.cs, self.match, self.data, self.M, self.N = \
            rows, cols, rt, rs, ct, cs, match, data, M, N
        self._c = None          # ctypes view, built by the backend on first use (backend._form)


class Core:
    """edges + data of a tensor, shared by renamed / conjugated views (the reference's refcounted Core, tensor.hpp:129-135)"""
    __slots__ = ("edges", "nb", "target", "tsign", "forms", "primary", "tables", "fermi", "dims", "sig", "_gd", "sforms")

    def __init__(self, edges, nb, target, tsign, fermi=0):
        self.edges = tuple(edges)
        self.dims = tuple([e.dim for e in self.edges])
        self.sig = (self.dims, tuple([e.harr is not None for e in self.edges]))      # what a plan depends on: dimensions, unit flags
        self._gd = {}
        self.sforms = None                           # signed regroupings of a fermionic tensor (ragged_fermi.signed_form), a few most recent
        self.nb = nb
        self.fermi = fermi                           # mask of the fermionic label components (0: bosonic symmetry)
        self.target, self.tsign = target, tsign      # device int32 [nbt] or None: sum of non-unit labels of a stored element
        self.forms = {}
        self.primary = None
        self.tables = {}

    # ---- group tables ---------------------------------------------------------------------------
    def group_dim(self, ids):
        m = self._gd.get(ids)
        if m is None:
            m = 1
            for i in ids:
                m *= self.dims[i]
            self._gd[ids] = m
        return m

    def table(self, ids):
        """(device table, sign) of the merged group `ids`: indices sorted by charge, per chain.  Tables depend on the label arrays
        only, so they are kept ON the first label array (shared by every tensor that carries these edges: site tensors, the
        boundary tensors derived from them, both sides of a bond)."""
        got = self.tables.get(ids)
        if got is None:
            B = _bk.get()
            es = [self.edges[i] for i in ids]
            if not es:
                got = (_empty_table(), 1)
            else:
                g = es[0].sign
                M = self.group_dim(ids)
                if M > TABLE_CACHE_MAX:
                    # a group spanning (most of) a big tensor: its table is as large as the tensor, keep it with the tensor only
                    STATS["sort"] += 1
                    hit = (B.rt_sort([(e.arr, e.sign * g, e.dim) for e in es]),)
                else:
                    key = tuple((id(e.arr), e.sign * g) for e in es)
                    store = es[0].arr.__dict__.setdefault("_rt_tables", {})
                    hit = store.get(key)
                    if hit is not None and not all(r() is e.arr for r, e in zip(hit[1], es[1:])):
                        hit = None          # an id was recycled by another array
                    if hit is None:
                        STATS["sort"] += 1
                        if len(store) >= 8:
                            # entries whose other label arrays are gone (the bonds of environments of earlier sweeps) can never hit
                            # again: without this a persistent first array -- the bond labels of a PEPS site tensor -- collected
                            # ~2.4 GB of dead tables per cfg2 step at 2368 chains
                            for k in [k for k, h in store.items() if any(r() is None for r in h[1])]:
                                del store[k]
                        # weak references to the other arrays: no reference cycles through the cache (device memory is freed by
                        # reference counting, not by the cycle collector)
                        hit = store[key] = (B.rt_sort([(e.arr, e.sign * g, e.dim) for e in es]), tuple(weakref.ref(e.arr) for e in es[1:]))
                got = (hit[0], g)
            self.tables[ids] = got
        return got

    def form(self, rows, cols):
        """the storage regrouped as rows | cols (cached; built from the primary form by one launch that also pairs the sectors)"""
        f = self.forms.get((rows, cols))
        if f is not None:
            return f
        f, job, ckey = self.form_job(rows, cols)
        _bk.get().rt_repack(*job)
        if ckey is not None:
            _learn(ckey, f.match)
        return f

    def form_job(self, rows, cols):
        """(form, arguments of the rt_repack launch that fills it): lets a contraction regroup both operands in ONE launch"""
        B = _bk.get()
        rt, rs = self.table(rows)
        ct, cs = self.table(cols)
        M, N = self.group_dim(rows), self.group_dim(cols)
        src = self.forms[self.primary]
        nbd = max(src.data.shape[0], src.match.shape[0], rt.shape[0], ct.shape[0], 1 if self.target is None else self.target.shape[0])
        ckey = ("form", self.dims, src.rows, src.cols, rows, cols)
        cap, learning = _cap(ckey, M * N, nbd)
        f = Form(rows, cols, rt, rs, ct, cs, None, B.rt_alloc(nbd, cap), M, N)
        STATS["repack"] += 1
        if len(self.forms) >= 4:          # keep the primary and the most recent regroupings only
            for k in list(self.forms):
                if k != self.primary:
                    del self.forms[k]
                    break
        self.forms[(rows, cols)] = f
        return f, (_repack_plan(self, src, f), src, f, (rs, cs, self.target, self.tsign, None, 0)), (ckey if learning else None)

    def set_primary(self, f):
        self.forms[(f.rows, f.cols)] = f
        self.primary = (f.rows, f.cols)
        self.tables.setdefault(f.rows, (f.rt, f.rs))
        self.tables.setdefault(f.cols, (f.ct, f.cs))


_EMPTY = {}


def _empty_table():
    """table of the empty edge group (one merged index of charge 0)"""
    B = _bk.get()
    t = _EMPTY.get(id(B))
    if t is None:
        t = _EMPTY[id(B)] = B.rt_sort([])
    return t


def _plan_array(rows, cols):
    """int32 descriptor [nr, nc, then per destination edge: dim, source group, source stride, magic lo, magic hi] with
    magic = floor(2^64 / dim) + 1 (exact division of a 32-bit index by one multiply-high on the device)"""
    out = [len(rows), len(cols)]
    for dim, grp, stride in rows + cols:
        m = ((1 << 64) // dim + 1) if dim > 1 else 0        # dimension 1 (kept for fermionic parities): the device skips the division
        lo, hi = m & 0xFFFFFFFF, (m >> 32) & 0xFFFFFFFF
        out += [dim, grp, stride, lo - (1 << 32) if lo >= (1 << 31) else lo, hi - (1 << 32) if hi >= (1 << 31) else hi]
    return np.array(out, dtype=np.int32)


def _repack_plan(core, src, dst, keep_dim1=False):
    """int32 descriptor of a regrouping: for every edge of the destination row group, then column group (slowest first):
    dimension, 1 if the edge sits in the source's column group, its stride inside that source group.  Edges of dimension 1 are
    dropped unless `keep_dim1` (signed regroupings of fermionic tensors need the parity of every device-labelled edge)"""
    key = ("rp", core.dims, src.rows, src.cols, dst.rows, dst.cols, keep_dim1)
    p = _PLANS.get(key)
    if p is None:
        where = {}
        for grp, ids in ((0, src.rows), (1, src.cols)):
            stride = 1
            for i in reversed(ids):
                where[i] = (grp, stride)
                stride *= core.edges[i].dim
        rows = [(core.edges[i].dim,) + where[i] for i in dst.rows if core.edges[i].dim != 1 or keep_dim1]
        cols = [(core.edges[i].dim,) + where[i] for i in dst.cols if core.edges[i].dim != 1 or keep_dim1]
        p = _PLANS[key] = _bk.get().upload(_plan_array(rows, cols))
    return p


def _dense_plan(dims, order_rows, order_cols, to_dense):
    """descriptor of dense <-> form conversions: the dense side is one group (all edges, row-major in tensor order)"""
    key = ("dp", tuple(dims), order_rows, order_cols, to_dense)
    p = _PLANS.get(key)
    if p is None:
        strides, acc = {}, 1
        for i in reversed(range(len(dims))):
            strides[i] = acc
            acc *= dims[i]
        if to_dense:
            # destination = dense (rows = all edges); source = form
            where = {}
            for grp, ids in ((0, order_rows), (1, order_cols)):
                stride = 1
                for i in reversed(ids):
                    where[i] = (grp, stride)
                    stride *= dims[i]
            rows = [(dims[i],) + where[i] for i in range(len(dims)) if dims[i] != 1]
            cols = []
        else:
            rows = [(dims[i], 0, strides[i]) for i in order_rows if dims[i] != 1]
            cols = [(dims[i], 0, strides[i]) for i in order_cols if dims[i] != 1]
        p = _PLANS[key] = _bk.get().upload(_plan_array(rows, cols))
    return p


class RTensor:
    """PyTAT-shaped view (names + core + conjugation sign) used by the lock-step drivers of tnsp_b200.tetragono"""
    __slots__ = ("names", "core", "sign")
    is_ragged = True

    def __init__(self, names, core, sign=1):
        self.names = list(names)
        self.core = core
        self.sign = sign

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def from_dense(cls, names, edges, dense, target=None, rows=None, fermi=0):
        """tensor from a dense device / host array [nb, prod dims] (entries outside the sectors are dropped)"""
        B = _bk.get()
        if isinstance(dense, np.ndarray):
            dense = B.from_numpy(np.ascontiguousarray(dense, dtype=np.float64))
        nb = dense.shape[0]
        if target is not None and not hasattr(target, "shape"):
            target = None if target == 0 else target
        if isinstance(target, np.ndarray):
            nb = max(nb, target.shape[0])
            target = B.from_numpy(np.ascontiguousarray(target, dtype=np.int32))
        core = Core(edges, nb, target, 1, fermi)
        nonunit = tuple(i for i, e in enumerate(core.edges) if not e.unit)
        if rows is None:
            rows = nonunit[:1]
        rows = tuple(rows)
        cols = tuple(i for i in nonunit if i not in rows)
        rt, rs = core.table(rows)
        ct, cs = core.table(cols)
        nbm = nb if target is not None else max(rt.shape[0], ct.shape[0])
        match, _ = B.rt_match(rt, rs, ct, cs, target, 1, None, 0, nbm)
        M, N = core.group_dim(rows), core.group_dim(cols)
        data = B.rt_alloc(dense.shape[0], M * N)
        f = Form(rows, cols, rt, rs, ct, cs, match, data, M, N)
        B.rt_repack(_dense_plan([e.dim for e in core.edges], rows, cols, False), dense, f)
        core.set_primary(f)
        return cls(names, core, 1)

    @classmethod
    def from_symmetric(cls, tensor, unit_names=()):
        """sector-compact copy of a block-symmetric device tensor (uniform structure; nb chains of data allowed).  Edges named
        in `unit_names` must have dimension 1: they become unit edges whose charge goes into the target."""
        S = tensor.Symmetry
        dims = [e.dimension for e in tensor._edges]
        h = np.atleast_2d(tensor._host())
        nb = h.shape[0]
        dense = np.zeros([nb] + dims, dtype=np.float64)
        starts = tensor._segment_starts()
        for b, pos in enumerate(tensor._table.positions):
            bd = [int(d) for d in tensor._table.dims[b]]
            off, size = int(tensor._table.offsets[b]), int(tensor._table.sizes[b])
            sl = (slice(None),) + tuple(slice(int(starts[i][int(p)]), int(starts[i][int(p)]) + bd[i]) for i, p in enumerate(pos))
            dense[sl] = h[:, off:off + size].reshape([nb] + bd)
        B = _bk.get()
        edges, target = [], 0
        for n, e in zip(tensor.names, tensor._edges):
            lab = np.concatenate([np.full(d, pack_symmetry(s), dtype=np.int32) for s, d in e.segments]) if e.segments else np.zeros(0, np.int32)
            if n in unit_names:
                if e.dimension != 1:
                    raise RuntimeError("a unit edge must have dimension 1")
                edges.append(Edge(1, None, 1, e.arrow, lab.reshape(1).copy()))
                target -= int(lab[0])
            else:
                edges.append(Edge(e.dimension, B.from_numpy(lab.reshape(1, -1)), 1, e.arrow))
        return cls.from_dense(list(tensor.names), edges, dense.reshape(nb, -1), np.array([target], dtype=np.int32) if target else None,
                              fermi=fermi_mask(S) if S.is_fermi_symmetry else 0)

    @classmethod
    def scalar_one(cls, value=1.0, fermi=0):
        return cls.from_dense([], [], np.array([[float(value)]]), fermi=fermi)

    def to_dense(self):
        """dense device array [nb, prod dims] in this tensor's edge order (zeros outside the sectors)"""
        B = _bk.get()
        core = self.core
        f = core.forms[core.primary]
        dims = [e.dim for e in core.edges]
        size = int(np.prod(dims)) if dims else 1
        nb = max(f.data.shape[0], f.match.shape[0])
        out = B.empty(nb, size)
        B.rt_repack(_dense_plan(dims, f.rows, f.cols, True), f, out)
        return out

    # ---- simple accessors -------------------------------------------------------------------------
    @property
    def rank(self):
        return len(self.names)

    @property
    def nb(self):
        return self.core.nb

    def edge_rename(self, dictionary):
        for n in dictionary:
            if n not in self.names:
                raise RuntimeError("Name missing in edge_rename")
        return RTensor([dictionary.get(n, n) for n in self.names], self.core, self.sign)

    def conjugate(self, trivial_metric=False):
        if _is_fermi(self):
            return _fermi_conjugate(self, trivial_metric)
        return RTensor(self.names, self.core, -self.sign)

    def copy(self):
        return RTensor(self.names, self.core, self.sign)      # immutable storage: a view is a copy

    __copy__ = copy

    def __deepcopy__(self, memo):
        return self.copy()

    def effective_edge(self, i):
        return self.core.edges[i].flipped(self.sign)

    def edge_labels(self, name):
        """host int64 [nbL, dim] of the effective labels of an edge (tests / set-up only: reads the device)"""
        return self.sign * self.core.edges[self.names.index(name)].host_labels()

    def _primary(self):
        return self.core.forms[self.core.primary]

    @property
    def storage(self):
        """host values of a tensor whose edges are all of dimension 1 (amplitudes): float array [nb]"""
        return _bk.get().to_numpy(self.scalar().t)

    def scalar(self):
        core = self.core
        if any(e.dim != 1 for e in core.edges):
            raise RuntimeError("Try to get the only element of the tensor which contains more than one element")
        f = core.form((), tuple(i for i, e in enumerate(core.edges) if not e.unit))
        return BatchScalar(_bk.get().rt_scalar(f.data, f.match))

    def __float__(self):
        return float(self.scalar())

    # ---- elementwise ------------------------------------------------------------------------------
    def _with_data(self, data):
        core, f = self.core, self._primary()
        nb = max(core.nb, data.shape[0])
        new = Core(core.edges, nb, core.target, core.tsign, core.fermi)
        new.tables = dict(core.tables)
        new.set_primary(Form(f.rows, f.cols, f.rt, f.rs, f.ct, f.cs, f.match, data, f.M, f.N))
        return RTensor(self.names, new, self.sign)

    def _scale(self, value, op):
        B = _bk.get()
        f = self._primary()
        if isinstance(value, BatchScalar):
            vec = value.t.contiguous()
        else:
            vec = B.from_numpy(np.array([float(value)], dtype=np.float64))
        return self._with_data(B.rt_scale(f.data, f.match, vec, op))

    def __mul__(self, o):
        if isinstance(o, RTensor):
            return self._binary(o, 2)
        return self._scale(o, 0)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, RTensor):
            return self._binary(o, 3)
        return self._scale(o, 1)

    def __imul__(self, o):
        r = self * o
        self.core, self.sign = r.core, r.sign
        return self

    def __itruediv__(self, o):
        r = self / o
        self.core, self.sign = r.core, r.sign
        return self

    def __neg__(self):
        return self._scale(-1.0, 0)

    def _binary(self, other, op):
        B = _bk.get()
        f = self._primary()
        if other.rank == 0 and self.rank != 0:
            return self._scale(other.scalar(), {2: 0, 3: 1}[op])
        if _is_fermi(self) and other.names != self.names:
            other = other.transpose(self.names)
        order = [other.names.index(n) for n in self.names]
        g = other.core.form(tuple(order[i] for i in f.rows), tuple(order[i] for i in f.cols))
        return self._with_data(B.rt_binary(f.data, g.data, f.match, op))

    def __add__(self, o):
        return self._binary(o, 0)

    def __sub__(self, o):
        return self._binary(o, 1)

    def _norm(self, kind):
        f = self._primary()
        r = _bk.get().rt_norm(f.data, f.match, kind)
        if r.shape[0] == 1:
            return float(_bk.get().to_numpy(r)[0])
        return BatchScalar(r)

    def norm_max(self):
        return self._norm(-1)

    def norm_2(self):
        return self._norm(2)

    def norm_sum(self):
        return self._norm(1)

    def transpose(self, target_names):
        """edge order is bookkeeping only for bosonic symmetries (storage is addressed by groups)"""
        target_names = list(target_names)
        if target_names == self.names:
            return self
        if _is_fermi(self):
            return _fermi_transpose(self, target_names)
        return _reordered(self, target_names)

    # ---- contract ---------------------------------------------------------------------------------
    def contract(self, other, contract_pairs, fuse_names=frozenset()):
        if fuse_names:
            raise NotImplementedError("fuse_names exists for tensors without symmetry only")
        if not isinstance(other, RTensor):
            raise TypeError("contract needs two tensors of the same type")
        if _is_fermi(self) or _is_fermi(other):
            return _fermi_contract(self, other, contract_pairs)
        return _contract(self, other, contract_pairs)

    # ---- qr / svd ---------------------------------------------------------------------------------
    def qr(self, free_names_direction, free_names, common_name_q, common_name_r):
        if free_names_direction in ("r", "R"):
            r_names = set(free_names)
            q_names = [n for n in self.names if n not in r_names]
        elif free_names_direction in ("q", "Q"):
            q_names = [n for n in self.names if n in set(free_names)]
        else:
            raise RuntimeError("Invalid direction in QR")
        for n in free_names:
            if n not in self.names:
                raise RuntimeError("Missing name in qr")
        return _factor(self, q_names, "qr", common_name_q, common_name_r, None, None, -1)

    def svd(self, free_names_u, common_name_u, common_name_v, singular_name_u, singular_name_v, cut=-1):
        for n in free_names_u:
            if n not in self.names:
                raise RuntimeError("Missing name in svd")
        u_names = [n for n in self.names if n in set(free_names_u)]
        return _factor(self, u_names, "svd", common_name_u, common_name_v, singular_name_u, singular_name_v, cut)


# -------------------------------------------------------------------------------------------------
def _is_fermi(t):
    return t.core.fermi != 0


def _reordered(t, target_names):
    order = [t.names.index(n) for n in target_names]
    if sorted(order) != list(range(len(t.names))):
        raise RuntimeError("Tensor to transpose with incompatible name list")
    core = t.core
    inv = {old: new for new, old in enumerate(order)}
    new = Core([core.edges[i] for i in order], core.nb, core.target, core.tsign, core.fermi)
    f = core.forms[core.primary]
    rows, cols = tuple(inv[i] for i in f.rows), tuple(inv[i] for i in f.cols)
    new.set_primary(Form(rows, cols, f.rt, f.rs, f.ct, f.cs, f.match, f.data, f.M, f.N))
    return RTensor(target_names, new, t.sign)


def _contract(a, b, pairs):
    B = _bk.get()
    STATS["contract"] += 1
    ca, cb = a.core, b.core
    key = ("ct", tuple(a.names), ca.sig, tuple(b.names), cb.sig, frozenset(pairs))
    p = _PLANS.get(key)
    if p is None:
        pairs = list(pairs)
        map12 = dict(pairs)
        if len(map12) != len(pairs) or len({y for _, y in pairs}) != len(pairs):
            raise RuntimeError("Duplicated names in contract pairs")
        for x, y in pairs:
            if x not in a.names or y not in b.names:
                raise RuntimeError("Missing name in contract")
        ea, eb = ca.edges, cb.edges
        ka, kb = [], []
        for i, n in enumerate(a.names):
            if n in map12:
                j = b.names.index(map12[n])
                if ea[i].dim != eb[j].dim:
                    raise RuntimeError("Contracting two edge with different dimension")
                if ea[i].unit != eb[j].unit:
                    if ea[i].dim != 1:
                        raise RuntimeError("unit / non-unit edge mismatch in contract")
                    continue        # a dim-1 pair of which one side is device-labelled: handled through the targets
                if not ea[i].unit:
                    ka.append(i)
                    kb.append(j)
        used_b = set(map12.values())
        fa = [i for i, n in enumerate(a.names) if n not in map12]
        fb = [j for j, n in enumerate(b.names) if n not in used_b]
        fa_n, fb_n = tuple(i for i in fa if not ea[i].unit), tuple(j for j in fb if not eb[j].unit)
        # positions of the result's indexed row / column edges (free edges of a, then of b)
        rows = tuple(fa.index(i) for i in fa_n)
        cols = tuple(len(fa) + fb.index(j) for j in fb_n)
        p = _PLANS[key] = (fa_n, tuple(ka), tuple(kb), fb_n, tuple(fa), tuple(fb), [a.names[i] for i in fa] + [b.names[j] for j in fb], rows, cols)
    fa_n, ka, kb, fb_n, fa, fb, names, rows, cols = p
    if not fa_n and not fb_n and ka:
        return _dot(a, b, ka, kb, fa, fb, names)
    asg, bsg = a.sign, b.sign
    A, Bf = ca.forms.get((fa_n, ka)), cb.forms.get((kb, fb_n))
    if A is None and Bf is None and ca is not cb:
        A, job_a, key_a = ca.form_job(fa_n, ka)
        Bf, job_b, key_b = cb.form_job(kb, fb_n)
        B.rt_repack_pair(*job_a, *job_b)
        if key_a is not None:
            _learn(key_a, A.match)
        if key_b is not None:
            _learn(key_b, Bf.match)
    else:
        A = ca.form(fa_n, ka) if A is None else A
        Bf = cb.form(kb, fb_n) if Bf is None else Bf
    # result core: free edges of a, then of b, with the operands' conjugation signs folded in
    ea, eb = ca.edges, cb.edges
    edges = ([ea[i] for i in fa] if asg == 1 else [ea[i].flipped(asg) for i in fa]) + \
            ([eb[j] for j in fb] if bsg == 1 else [eb[j].flipped(bsg) for j in fb])
    rs, cs = A.rs * asg, Bf.cs * bsg
    nb = max(ca.nb, cb.nb, A.match.shape[0], Bf.match.shape[0])
    cap, learning = _cap(key, A.M * Bf.N, nb)
    data = B.rt_alloc(nb, cap)
    C = Form(rows, cols, A.rt, rs, Bf.ct, cs, None, data, A.M, Bf.N)
    ksign = -(asg * A.cs) * (bsg * Bf.rs)
    # ONE launch: sector pairing of the result (rows of a, columns of b, summed targets) + every sector GEMM of every chain
    target = B.rt_gemm(A, Bf, C, ksign, nb, (rs, cs, ca.target, ca.tsign * asg, cb.target, cb.tsign * bsg))
    if learning:
        _learn(key, C.match)
    core = Core(edges, nb, target, 1, ca.fermi | cb.fermi)
    core.set_primary(C)
    return RTensor(names, core, 1)


def _dot(a, b, ka, kb, fa, fb, names):
    """contraction over every indexed edge of both tensors: one pass over the stored elements of `a` (the larger data is walked in
    its own layout, `b` is read through its tables) -- no regrouping, no table over the whole tensor"""
    B = _bk.get()
    if b.core.forms[b.core.primary].M * b.core.forms[b.core.primary].N > a.core.forms[a.core.primary].M * a.core.forms[a.core.primary].N:
        a, b, ka, kb, fa, fb = b, a, kb, ka, fb, fa       # walk the operand with the larger index space
        swapped = True
    else:
        swapped = False
    A, S = a.core.forms[a.core.primary], b.core.forms[b.core.primary]
    key = ("dot", tuple(e.dim for e in a.core.edges), A.rows, A.cols, tuple(e.dim for e in b.core.edges), S.rows, S.cols, ka, kb)
    plan = _PLANS.get(key)
    if plan is None:
        pair = dict(zip(ka, kb))
        where = {}
        for grp, ids in ((0, S.rows), (1, S.cols)):
            stride = 1
            for j in reversed(ids):
                where[j] = (grp, stride)
                stride *= b.core.edges[j].dim
        rows = [(a.core.edges[i].dim,) + where[pair[i]] for i in A.rows if a.core.edges[i].dim != 1]
        cols = [(a.core.edges[i].dim,) + where[pair[i]] for i in A.cols if a.core.edges[i].dim != 1]
        plan = _PLANS[key] = B.upload(_plan_array(rows, cols))
    nb = max(a.core.nb, b.core.nb, A.match.shape[0], S.match.shape[0])
    data, match, target = B.rt_dot(plan, S, A, a.core.target, a.core.tsign * a.sign, b.core.target, b.core.tsign * b.sign, nb)
    if swapped:
        a, b, fa, fb = b, a, fb, fa
    edges = [a.core.edges[i].flipped(a.sign) for i in fa] + [b.core.edges[j].flipped(b.sign) for j in fb]
    core = Core(edges, nb, target, 1, a.core.fermi | b.core.fermi)
    empty = _empty_table()
    core.set_primary(Form((), (), empty, 1, empty, 1, match, data, 1, 1))
    return RTensor(names, core, 1)


def _unit_target(t, ids):
    """host int32 [nbL] = -(sum of the effective labels of the unit edges at positions `ids`); None when there is none"""
    edges = t.core.edges
    total = None
    for i in ids:
        e = edges[i]
        if e.harr is not None:
            v = -(e.sign * t.sign) * np.asarray(e.harr, dtype=np.int64).reshape(-1)
            total = v if total is None else total + v
    if total is None or not total.any():
        return None
    return total.astype(np.int32)


def _factor(t, first_names, kind, name_1, name_2, sing_1, sing_2, cut):
    """shared front end of qr (first = Q side) and svd (first = U side); svd.hpp:259-538, qr.hpp:309-508"""
    B = _bk.get()
    STATS[kind] += 1
    core = t.core
    first = [t.names.index(n) for n in first_names]
    second = [i for i in range(len(t.names)) if i not in first]
    rows = tuple(i for i in first if not core.edges[i].unit)
    cols = tuple(i for i in second if not core.edges[i].unit)
    fermi = core.fermi != 0
    # fermionic tensors: the transposition to (first..., second...) carries its sign; the new bond points from the second factor
    # to the first (arrow true on Q / U, false on R / V; S: false, true) as in the reference's use_qr / put_v_right branch
    F = _signed_form(t, rows, cols, _fermi_factor_form(t, first, second)) if fermi else core.form(rows, cols)
    nb = max(core.nb, F.match.shape[0])
    kdim = min(F.M, F.N)
    remain_cut, relative_cut = (1 << 30), 0.0
    if kind == "svd" and cut > 0:
        if cut >= 1:
            remain_cut = int(cut)
        else:
            relative_cut = float(cut)
    kdim = min(kdim, remain_cut)
    # target of the first factor: charge carried by the unit edges on its side (zero on the hot path: the R / U side of
    # two_line_to_one_line has no physical edges)
    u1 = _unit_target(t, first)
    u2 = _unit_target(t, second)
    if u1 is not None and u2 is not None:
        t1 = B.from_numpy(np.broadcast_to(u1, (nb,)).copy() if u1.shape[0] != nb else u1)
        t1s = 1
    elif u2 is None:
        t1, t1s = core.target, core.tsign * t.sign      # everything charged sits on the first side
    else:
        t1, t1s = None, 0
    # bond labels + per-sector factorisation, all planned on the device
    fkey = ("fac", kind, core.dims, rows, cols, kdim)
    c1, l1 = _cap(fkey + (1,), F.M * max(kdim, 1), nb)
    c2, l2 = _cap(fkey + (2,), max(kdim, 1) * F.N, nb)
    out = B.rt_factor(kind, F, t.sign, core.target, core.tsign * t.sign, t1, t1s, kdim, remain_cut, relative_cut, nb, (c1, c2))
    if l1:
        _learn(fkey + (1,), out["first"][0])
        _learn(fkey + (2,), out["second"][0])
    lab = out["labels"]                      # device int32 [nb, kdim]: effective label of the bond on the first factor
    e1 = [core.edges[i].flipped(t.sign) for i in first]
    e2 = [core.edges[i].flipped(t.sign) for i in second]
    bond_1 = Edge(kdim, lab, 1, fermi)
    bond_2 = Edge(kdim, lab, -1, False)

    def make(names, edges, rows_ids, cols_ids, form_data, target, tsign, tables):
        c = Core(edges, nb, target, tsign, core.fermi)
        for ids, tb in tables.items():
            c.tables[ids] = tb
        rt, rs = c.table(rows_ids)
        ct, cs = c.table(cols_ids)
        c.set_primary(Form(rows_ids, cols_ids, rt, rs, ct, cs, form_data[0], form_data[1], c.group_dim(rows_ids), c.group_dim(cols_ids)))
        return RTensor(names, c, 1)

    n1 = len(first)
    rows_1 = tuple(i for i in range(n1) if not e1[i].unit)
    cols_2 = tuple(1 + i for i in range(len(second)) if not e2[i].unit)
    tb_bond_c = {(n1,): out["bond_col"]}
    tb_bond_r = {(0,): out["bond_row"]}
    tab_rows = {rows_1: (F.rt, F.rs * t.sign)}
    tab_cols = {cols_2: (F.ct, F.cs * t.sign)}
    if t1 is None:
        tgt1, tgt2 = (None, 0), (core.target, core.tsign * t.sign)
    elif u2 is None:
        tgt1, tgt2 = (t1, t1s), (None, 0)
    else:
        tgt1 = (t1, 1)
        tgt2 = (B.from_numpy(np.broadcast_to(u2, (nb,)).copy() if u2.shape[0] != nb else u2), 1)
    first_t = make([t.names[i] for i in first] + [name_1], e1 + [bond_1], rows_1, (n1,), out["first"], tgt1[0], tgt1[1], {**tab_rows, **tb_bond_c})
    second_t = make([name_2] + [t.names[i] for i in second], [bond_2] + e2, (0,), cols_2, out["second"], tgt2[0], tgt2[1], {**tab_cols, **tb_bond_r})
    if kind == "qr":
        return first_t, second_t
    s_t = make([sing_1, sing_2], [bond_2, bond_1], (0,), (1,), out["s"], None, 0, {(0,): out["bond_row"], (1,): out["bond_col"]})
    return first_t, s_t, second_t


# ---- fermionic variants are installed by ragged_fermi (signs of edge_operator.hpp:497-555, contract.hpp:570-580) ----
def _fermi_contract(a, b, pairs):
    raise NotImplementedError("fermionic sector-compact tensors: signs not installed")


def _fermi_transpose(t, names):
    raise NotImplementedError("fermionic sector-compact tensors: signs not installed")


def _fermi_conjugate(t, trivial_metric):
    raise NotImplementedError("fermionic sector-compact tensors: signs not installed")


def _fermi_factor_form(t, first, second):
    raise NotImplementedError("fermionic sector-compact tensors: signs not installed")


def _signed_form(t, rows, cols, form):
    raise NotImplementedError("fermionic sector-compact tensors: signs not installed")


from . import ragged_fermi as _ragged_fermi  # noqa: E402,F401  (installs the fermionic variants above)

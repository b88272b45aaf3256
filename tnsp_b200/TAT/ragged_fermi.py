"""Fermionic signs of sector-compact tensors (TAT/ragged.py).

The reference attaches a sign to every block an edge operation moves (TAT/include/TAT/implement/edge_operator.hpp:497-555, 591,
612): the XOR of (i) the parities of reversed edges that are flagged to carry a sign, (ii) one factor per pair of odd edges whose
order the transpose swaps, (iii) the "reverse the order of n_odd fermions" factor (n_odd & 2) of every flagged merge / split group.
All three are polynomials of degree <= 2 over GF(2) in the PARITIES of the element's indices, so for a lock-step batch -- where
the parity of an index differs from chain to chain -- a sign is described once per operation by a quadratic form

    sign(element) = s0 + sum_i L_i p_i + sum_{i<j} Q_ij p_i p_j        (mod 2; p_i = parity of the element's index on edge i)

(`SignForm`), and evaluated per element on the device from the per-chain labels.  Dimension-1 edges with host-known charges (the
physical P edges, the total-symmetry edge) have one parity per chain: they are folded on the host into a per-chain constant and
per-chain linear coefficients of the remaining edges.

The forms below follow the host planner of tnsp_b200/TAT/plan.py rule for rule (contract.hpp:365-412, 488-517, 570-580; svd.hpp:
312-350; qr.hpp:339-394; conjugate.hpp:48-97); tests/test_sector_fermi.py checks every one of them against that planner (itself
pinned to the unmodified reference) on random tensors of all fermionic integer symmetries.
"""
from __future__ import annotations

import numpy as np

from .. import backend as _bk
from . import ragged
from .ragged import Core, Form, RTensor, STATS, _PLANS


SIGNED_CACHE_MAX = 16384      # index-space size (M x N) up to which a signed regrouping is kept with its core


class SignForm:
    """quadratic form over GF(2) on the edge positions of one tensor: lin = set of positions, quad = set of frozenset pairs.
    A form is built once (add_*), then only read; `key()` and the folds of `fold_units` are remembered on it (the forms of a
    contraction live in its cached plan, so a program point pays for them once)"""
    __slots__ = ("lin", "quad", "_key", "_folds")

    def __init__(self):
        self.lin = set()
        self.quad = set()
        self._key = None
        self._folds = None

    def add_lin(self, i):
        self.lin ^= {i}
        self._key = self._folds = None

    def add_pair(self, i, j):
        if i != j:
            self.quad ^= {frozenset((i, j))}
            self._key = self._folds = None

    def add_clique(self, group):
        """(n_odd & 2) of a group of edges = sum over its pairs"""
        g = list(group)
        for a in range(len(g)):
            for b in range(a + 1, len(g)):
                self.add_pair(g[a], g[b])

    def add_transposition(self, order):
        """order[k] = original position of the edge that ends at place k: one factor per inverted pair"""
        for a in range(len(order)):
            for b in range(a + 1, len(order)):
                if order[a] > order[b]:
                    self.add_pair(order[a], order[b])

    def empty(self):
        return not self.lin and not self.quad

    def key(self):
        k = self._key
        if k is None:
            k = self._key = (tuple(sorted(self.lin)), tuple(sorted(tuple(sorted(p)) for p in self.quad)))
        return k


def unit_parities(t, mask):
    """{position: int array [nbL] of parities} of the unit edges of tensor view `t` (effective labels)"""
    edges = t.core.edges
    return {i: edges[i].unit_parity(mask) for i, u in enumerate(t.core.sig[1]) if u}


def _compile_fold(form, entries, units):
    """the chain-independent part of folding `form` onto the indexed edges `entries` when the edges at `units` are unit edges:
    (quad int32 [n] | None when it vanishes, constant linear bits, [(unit position, shift)] linear terms carried by a unit parity,
    [unit position] constant terms, [(unit, unit)] constant products)"""
    slot = {pos: k for k, pos in enumerate(entries)}
    n = len(entries)
    if n > 30:
        raise NotImplementedError("sign form over more than 30 indexed edges")
    quad = [0] * max(n, 1)
    lin0 = 0
    lin_terms, const_terms, const_pairs = [], [], []
    for i in form.lin:
        k = slot.get(i)
        if k is not None:
            lin0 ^= 1 << k
        elif i in units:
            const_terms.append(i)
    for pr in form.quad:
        i, j = tuple(pr)
        ki, kj = slot.get(i), slot.get(j)
        if ki is not None and kj is not None:
            if ki > kj:
                ki, kj = kj, ki
            quad[ki] ^= 1 << kj
        elif ki is not None and j in units:
            lin_terms.append((j, ki))
        elif kj is not None and i in units:
            lin_terms.append((i, kj))
        elif i in units and j in units:
            const_pairs.append((i, j))
    return (np.array(quad, dtype=np.int32), any(quad), lin0, tuple(lin_terms), tuple(const_terms), tuple(const_pairs))


def fold_units(form, t, entries):
    """Reduce a sign form to the indexed edges `entries` (positions of the non-unit edges in the order the kernel sees them).

    Returns (quad masks int32 [n_ent]: bit j of entry k set iff Q_kj = 1 and j > k; per-chain int32 [nb or 1]: bits 0..n_ent-1
    the linear coefficients, bit 31 the constant).  None, None when the form vanishes identically."""
    up = unit_parities(t, t.core.fermi)
    ckey = (tuple(entries), tuple(up))
    folds = form._folds
    if folds is None:
        folds = form._folds = {}
    comp = folds.get(ckey)
    if comp is None:
        comp = folds[ckey] = _compile_fold(form, entries, set(up))
    quad, any_quad, lin0, lin_terms, const_terms, const_pairs = comp
    if not lin_terms and not const_terms and not const_pairs:
        if not any_quad and lin0 == 0:
            return None, None
        return quad, np.array([lin0], dtype=np.int32)
    # unit parities are uint32 arrays [nbL]: bits 0..30 the linear coefficients, bit 31 the constant, reinterpreted as int32 at the end
    acc = None
    for j, k in lin_terms:
        v = up[j] << np.uint32(k)
        acc = v if acc is None else acc ^ v
    const_a = None
    for i in const_terms:
        const_a = up[i] if const_a is None else const_a ^ up[i]
    for i, j in const_pairs:
        v = up[i] & up[j]
        const_a = v if const_a is None else const_a ^ v
    if const_a is not None:
        v = (const_a & np.uint32(1)) << np.uint32(31)
        acc = v if acc is None else acc ^ v
    if lin0:
        acc = acc ^ np.uint32(lin0)
    acc = np.atleast_1d(acc)
    if not any_quad and not acc.any():
        return None, None
    return quad, acc.view(np.int32)


# -------------------------------------------------------------------------------------------------
# signed regrouping
# -------------------------------------------------------------------------------------------------
def signed_form(t, rows, cols, form):
    """storage of tensor view `t` regrouped as rows | cols with the sign form applied.  Falls back to the cached unsigned regrouping
    when the form vanishes.  The result depends on the (immutable) core, the grouping and the form only, so the few most recent ones
    are kept with the core: a site tensor or an environment enters several contractions of a sweep with the same signs."""
    core = t.core
    if form.empty():
        return core.form(rows, cols)
    skey = (rows, cols, form.key())
    if core.sforms is not None:
        hit = core.sforms.get(skey)
        if hit is not None:
            STATS["signed_hit"] += 1
            return hit
    f = _build_signed_form(t, rows, cols, form)
    if f.M * f.N <= SIGNED_CACHE_MAX:          # small tensors only (site tensors, strip pieces): the big environments would double the footprint
        if core.sforms is None:
            core.sforms = {}
        elif len(core.sforms) >= 2:
            core.sforms.pop(next(iter(core.sforms)))
        core.sforms[skey] = f
    return f


_DEVICE_CONSTANTS: dict = {}


def _device_constant(B, array):
    """device copy of a small chain-independent int32 array (the quadratic part of a sign form, a chain-independent linear part): the same
    few hundred arrays recur at every program point of every sweep, so they are uploaded once instead of once per regrouping"""
    key = (id(B), array.shape, array.tobytes())
    got = _DEVICE_CONSTANTS.get(key)
    if got is None:
        if len(_DEVICE_CONSTANTS) >= 8192:
            _DEVICE_CONSTANTS.clear()
        got = _DEVICE_CONSTANTS[key] = B.upload(array)
    return got


def _build_signed_form(t, rows, cols, form):
    core = t.core
    entries = list(rows) + list(cols)          # every device-labelled edge, dimension 1 included: its parity is only known there
    quad, per_chain = fold_units(form, t, entries)
    if quad is None:
        return core.form(rows, cols)
    B = _bk.get()
    src = core.forms[core.primary]
    rt, rs = core.table(rows)
    ct, cs = core.table(cols)
    M, N = core.group_dim(rows), core.group_dim(cols)
    ckey = ("form", core.dims, src.rows, src.cols, rows, cols)
    nbd = max(src.data.shape[0], src.match.shape[0], rt.shape[0], ct.shape[0], per_chain.shape[0], 1 if core.target is None else core.target.shape[0])
    cap, learning = ragged._cap(ckey, M * N, nbd)
    f = Form(rows, cols, rt, rs, ct, cs, None, B.rt_alloc(nbd, cap), M, N)
    STATS["repack"] += 1
    labels = [(core.edges[i].arr, core.edges[i].dim) for i in entries]
    B.rt_repack(ragged._repack_plan(core, src, f, True), src, f, (rs, cs, core.target, core.tsign, None, 0),
                sign=(_device_constant(B, quad), _device_constant(B, per_chain) if per_chain.shape[0] == 1 else B.upload(per_chain), labels, core.fermi))
    if learning:
        ragged._learn(ckey, f.match)
    return f


def _with_primary(t, f, edges=None, names=None, sign=None):
    core = t.core
    new = Core(core.edges if edges is None else edges, max(core.nb, f.data.shape[0]), core.target, core.tsign, core.fermi)
    new.set_primary(f)
    return RTensor(t.names if names is None else names, new, t.sign if sign is None else sign)


def fermi_transpose(t, target_names):
    """transpose = relabelling of the edge order + one factor per swapped pair of odd edges (edge_operator.hpp:521-555)"""
    order = [t.names.index(n) for n in target_names]
    if sorted(order) != list(range(len(t.names))):
        raise RuntimeError("Tensor to transpose with incompatible name list")
    form = SignForm()
    form.add_transposition(order)
    core = t.core
    p = core.forms[core.primary]
    f = signed_form(t, p.rows, p.cols, form)
    inv = {old: new for new, old in enumerate(order)}
    new = Core([core.edges[i] for i in order], max(core.nb, f.data.shape[0]), core.target, core.tsign, core.fermi)
    new.set_primary(Form(tuple(inv[i] for i in f.rows), tuple(inv[i] for i in f.cols), f.rt, f.rs, f.ct, f.cs, f.match, f.data, f.M, f.N))
    return RTensor(list(target_names), new, t.sign)


def fermi_conjugate(t, trivial_metric):
    """conjugate.hpp:48-97: labels and arrows flip; sign = (n_odd & 2) + (trivial_metric: parities of the edges with arrow true)"""
    form = SignForm()
    n = len(t.names)
    form.add_clique(range(n))
    if trivial_metric:
        for i in range(n):
            if t.effective_edge(i).arrow:
                form.add_lin(i)
    core = t.core
    p = core.forms[core.primary]
    f = signed_form(t, p.rows, p.cols, form)
    new = Core(core.edges, max(core.nb, f.data.shape[0]), core.target, core.tsign, core.fermi)
    new.tables = dict(core.tables)
    new.set_primary(f)
    return RTensor(t.names, new, -t.sign)


def fermi_contract(a, b, pairs):
    """contract.hpp:306-620 with the operand layout fixed to (free | common) x (common | free), for which the sector GEMMs need no
    extra factor (contract.hpp:570-580: alpha = -1 only when exactly one operand has its common edges first):
      operand 1: common edges are brought to arrow true WITH sign (the odd ones among those that had arrow false), free edges to
                 arrow false without; transposition to (free..., common...); merge sign (n_odd & 2) on the common group;
      operand 2: reversals without sign; transposition to (common in the order of operand 1..., free...)."""
    ea, eb = a.core.edges, b.core.edges
    pkey = ("fplan", tuple(a.names), a.core.sig, a.sign, tuple([e.arrow for e in ea]), tuple(b.names), b.core.sig, b.sign,
            tuple([e.arrow for e in eb]), frozenset(pairs))
    plan = _PLANS.get(pkey)
    if plan is None:
        pairs = list(pairs)
        map12 = dict(pairs)
        for x, y in pairs:
            if x not in a.names or y not in b.names:
                raise RuntimeError("Missing name in contract")
        used_b = set(map12.values())
        ka_all = [i for i, n in enumerate(a.names) if n in map12]
        kb_all = [b.names.index(map12[a.names[i]]) for i in ka_all]
        fa = [i for i, n in enumerate(a.names) if n not in map12]
        fb = [j for j, n in enumerate(b.names) if n not in used_b]
        form_a, form_b = SignForm(), SignForm()
        for i in ka_all:
            if not a.effective_edge(i).arrow:
                form_a.add_lin(i)
        form_a.add_transposition(fa + ka_all)
        form_a.add_clique(ka_all)
        form_b.add_transposition(kb_all + fb)
        for i, j in zip(ka_all, kb_all):
            if ea[i].dim != eb[j].dim:
                raise RuntimeError("Contracting two edge with different dimension")
            if ea[i].unit != eb[j].unit:
                raise NotImplementedError("contract of a host-labelled dimension-1 edge with a device-labelled one")
        fa_n = tuple(i for i in fa if not ea[i].unit)
        fb_n = tuple(j for j in fb if not eb[j].unit)
        ka = tuple(i for i in ka_all if not ea[i].unit)
        kb = tuple(j for i, j in zip(ka_all, kb_all) if not ea[i].unit)
        names = [a.names[i] for i in fa] + [b.names[j] for j in fb]
        rows = tuple(k for k, i in enumerate(fa) if not ea[i].unit)
        cols = tuple(len(fa) + k for k, j in enumerate(fb) if not eb[j].unit)
        plan = _PLANS[pkey] = (fa, fb, fa_n, fb_n, ka, kb, names, rows, cols, form_a, form_b)
    fa, fb, fa_n, fb_n, ka, kb, names, rows, cols, form_a, form_b = plan
    B = _bk.get()
    STATS["contract"] += 1
    A = signed_form(a, fa_n, ka, form_a)
    Bf = signed_form(b, kb, fb_n, form_b)
    nb = max(a.core.nb, b.core.nb, A.match.shape[0], Bf.match.shape[0], A.data.shape[0], Bf.data.shape[0])
    edges = [ea[i].flipped(a.sign) for i in fa] + [eb[j].flipped(b.sign) for j in fb]
    rs, cs = A.rs * a.sign, Bf.cs * b.sign
    key = ("fct", a.core.dims, fa_n, ka, b.core.dims, kb, fb_n)
    cap, learning = ragged._cap(key, A.M * Bf.N, nb)
    C = Form(rows, cols, A.rt, rs, Bf.ct, cs, None, B.rt_alloc(nb, cap), A.M, Bf.N)
    ksign = -(a.sign * A.cs) * (b.sign * Bf.rs)
    target = B.rt_gemm(A, Bf, C, ksign, nb, (rs, cs, a.core.target, a.core.tsign * a.sign, b.core.target, b.core.tsign * b.sign))
    if learning:
        ragged._learn(key, C.match)
    core = Core(edges, nb, target, 1, a.core.fermi | b.core.fermi)
    core.set_primary(C)
    return RTensor(names, core, 1)


def fermi_factor_input(t, first, second):
    """svd.hpp:312-350 / qr.hpp:339-394 with the matrix laid out as (first | second) (the reference's put_v_right / use_qr branch):
    edges with arrow true are reversed WITHOUT sign, the transposition to (first..., second...) carries its sign"""
    form = SignForm()
    form.add_transposition(list(first) + list(second))
    return form


def install():
    ragged._fermi_contract = fermi_contract
    ragged._fermi_transpose = fermi_transpose
    ragged._fermi_conjugate = fermi_conjugate
    ragged._fermi_factor_form = fermi_factor_input
    ragged._signed_form = signed_form


install()

"""Shared generators for the differential tests (same random structures on both implementations)."""
import numpy as np

SYMS = ["No", "BoseZ2", "BoseU1", "FermiU1", "FermiZ2", "FermiU1BoseZ2", "FermiU1BoseU1", "FermiU1FermiU1"]
KINDS = {
    "No": [], "BoseZ2": ["Z2"], "BoseU1": ["U1"], "FermiU1": ["U1"], "FermiZ2": ["Z2"],
    "FermiU1BoseZ2": ["U1", "Z2"], "FermiU1BoseU1": ["U1", "U1"], "FermiU1FermiU1": ["U1", "U1"],
}
FERMI = {"No": False, "BoseZ2": False, "BoseU1": False, "FermiU1": True, "FermiZ2": True, "FermiU1BoseZ2": True,
         "FermiU1BoseU1": True, "FermiU1FermiU1": True}


def rand_sym(rng, sym):
    out = []
    for k in KINDS[sym]:
        out.append(bool(rng.integers(0, 2)) if k == "Z2" else int(rng.integers(-1, 2)))
    return tuple(out)


def neg_sym(sym, s):
    return tuple(v if k == "Z2" else -v for v, k in zip(s, KINDS[sym]))


def rand_edge(rng, sym, max_seg=3, max_dim=3):
    """python description: (segments [(sym tuple, dim)], arrow)"""
    if sym == "No":
        return ([((), int(rng.integers(1, max_dim + 2)))], False)
    n = int(rng.integers(1, max_seg + 1))
    seen, segs = set(), []
    for _ in range(n):
        s = rand_sym(rng, sym)
        if s in seen:
            continue
        seen.add(s)
        segs.append((s, int(rng.integers(1, max_dim + 1))))
    arrow = bool(rng.integers(0, 2)) if FERMI[sym] else False
    return (segs, arrow)


def conj_edge(sym, e):
    segs, arrow = e
    return ([(neg_sym(sym, s), d) for s, d in segs], (not arrow) if FERMI[sym] else False)


def make_edge(mod, sym, e):
    """build an Edge object in module `mod` (reference TAT or tnsp_b200.TAT)"""
    m = getattr(mod, sym)
    segs, arrow = e
    if sym == "No":
        return m.Edge(segs[0][1])
    S = m.Symmetry
    return m.Edge([(S(*s), d) for s, d in segs], arrow)


def make_tensor(mod, sym, names, edges, values=None):
    m = getattr(mod, sym)
    t = m.D.Tensor(list(names), [make_edge(mod, sym, e) for e in edges])
    if values is not None:
        t.storage = values
    return t


def storage(t):
    return np.array(t.storage, dtype=np.float64).reshape(-1).copy()


FIELDS = {"No": (), "BoseZ2": ("z2",), "BoseU1": ("u1",), "FermiU1": ("fermi",), "FermiZ2": ("parity",),
          "FermiU1BoseZ2": ("fermi", "z2"), "FermiU1BoseU1": ("fermi", "u1"), "FermiU1FermiU1": ("fermi_0", "fermi_1")}


def describe(t, sym):
    """plain-python structure description comparable across both modules (never repr() a reference
    tensor: its text output is not safe in a process that also loaded torch)"""
    out = []
    for n in t.names:
        e = t.edge_by_name(n)
        segs = tuple((tuple(int(getattr(s, f)) for f in FIELDS[sym]), int(d)) for s, d in e.segments)
        out.append((str(n), segs, bool(e.arrow)))
    return tuple(out)

"""Host-side integer metadata of block-symmetric tensors: Symmetry, Edge, block tables.

This is the part of the path that must be bit-exact with the reference:

* Symmetry algebra         -- reference TAT/include/TAT/structure/symmetry.hpp:94-239
* Edge (segments + arrow)  -- reference TAT/include/TAT/structure/edge.hpp:32-292
* Block table / layout     -- reference TAT/include/TAT/structure/core.hpp:152-191

Nothing here touches tensor data; block tables are plain numpy int64 arrays that the planners in
``plan.py`` turn into descriptor lists for the CUDA kernels.
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_symmetry_class", "Edge", "BlockTable", "block_table"]


class SymmetryBase(tuple):
    """Tuple of integers; ``bool`` (Z2) components add by XOR, ``int`` (U1) components add normally.

    Ordering, equality and hashing are those of the underlying tuple, as in the reference
    (symmetry.hpp: comparison of std::tuple; PyTAT.hpp:147-153 hashes the base tuple).
    """

    __slots__ = ()
    kinds: tuple = ()  # "Z2" or "U1" per component
    fermi_flags: tuple = ()  # bool per component
    field_names: tuple = ()
    short_name = "No"
    is_fermi_symmetry = False
    length = 0

    def __new__(cls, *args):
        if len(args) == 1:
            a = args[0]
            if type(a) is cls:
                return a
            if isinstance(a, (tuple, list)):
                args = tuple(a)
        n = cls.length
        if len(args) > n:
            raise TypeError(f"{cls.__name__} takes at most {n} components")
        vals = []
        for i in range(n):
            v = args[i] if i < len(args) else 0
            if cls.kinds[i] == "Z2":
                vals.append(bool(v))
            else:
                if isinstance(v, bool) or not isinstance(v, (int, np.integer)):
                    if isinstance(v, bool):
                        v = int(v)
                    else:
                        raise TypeError(f"invalid symmetry component {v!r}")
                vals.append(int(v))
        return tuple.__new__(cls, vals)

    def __add__(self, other):
        other = type(self)(other)
        return tuple.__new__(type(self), [(a ^ b) if k == "Z2" else (a + b) for a, b, k in zip(self, other, self.kinds)])

    __radd__ = __add__

    def __sub__(self, other):
        other = type(self)(other)
        return tuple.__new__(type(self), [(a ^ b) if k == "Z2" else (a - b) for a, b, k in zip(self, other, self.kinds)])

    def __neg__(self):
        return tuple.__new__(type(self), [a if k == "Z2" else -a for a, k in zip(self, self.kinds)])

    def __eq__(self, other):
        if type(other) is not type(self):
            try:
                other = type(self)(other)
            except TypeError:
                return NotImplemented
        return tuple.__eq__(self, other)

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = tuple.__hash__

    @property
    def parity(self) -> bool:
        r = False
        for a, f in zip(self, self.fermi_flags):
            if f:
                r ^= bool(a % 2) if not isinstance(a, bool) else a
        return r

    def __repr__(self):
        return f"{self.short_name}Symmetry[{self}]"

    def __str__(self):
        if self.length == 0:
            return ""
        if self.length == 1:
            return str(int(self[0]))
        return "(" + ",".join(str(int(a)) for a in self) + ")"

    def __reduce__(self):
        return (type(self), tuple(self))

    def __getnewargs__(self):
        return tuple(self)


def make_symmetry_class(short_name, components):
    """components: list of (field_name, kind, is_fermi)."""
    ns = {
        "__slots__": (),
        "kinds": tuple(k for _, k, _ in components),
        "fermi_flags": tuple(f for _, _, f in components),
        "field_names": tuple(n for n, _, _ in components),
        "short_name": short_name,
        "is_fermi_symmetry": any(f for _, _, f in components),
        "length": len(components),
    }
    for i, (field, _, _) in enumerate(components):
        ns[field] = property(lambda self, i=i: tuple.__getitem__(self, i))
    cls = type(short_name + "Symmetry", (SymmetryBase,), ns)
    cls.__module__ = __name__
    globals()[cls.__name__] = cls          # picklable by reference (PyTAT.hpp:160-175 pickles symmetries)
    return cls


class _Segments(tuple):
    """segment list of an edge: an immutable tuple inside (edges are hashable plan-cache keys) that compares equal to the Python
    list PyTAT returns (PyTAT.hpp:257-262)"""
    __slots__ = ()

    def __eq__(self, other):
        return tuple.__eq__(self, tuple(other)) if isinstance(other, (list, tuple)) else NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = tuple.__hash__


class Edge:
    """Ordered list of (symmetry, dimension) segments plus a fermi arrow (edge.hpp:32-292).

    Immutable; equality compares segments and (for fermionic symmetries) the arrow.
    Subclasses created per symmetry type carry ``Symmetry``.
    """

    __slots__ = ("segments", "arrow", "_hash", "_dims", "_syms", "_parity", "_comp")
    Symmetry = None

    def __init__(self, *args):
        S = self.Symmetry
        arrow = False
        if len(args) == 0:
            segs = []
        elif len(args) == 1:
            a = args[0]
            if isinstance(a, Edge):
                segs, arrow = a.segments, a.arrow
            elif isinstance(a, tuple) and len(a) == 2 and isinstance(a[0], (list, tuple)) and isinstance(a[1], (bool, np.bool_)) \
                    and not (S.length == 2 and not isinstance(a[0], list) and self._looks_like_symmetry(a)):
                # (segments, arrow) pair
                segs, arrow = a[0], bool(a[1])
            else:
                segs = a
        elif len(args) == 2:
            segs, arrow = args[0], bool(args[1])
        else:
            raise TypeError("Edge(segments[, arrow])")
        self.segments = self._normalize(segs)
        self.arrow = bool(arrow) if S.is_fermi_symmetry else False
        self._hash = None
        self._dims = None
        self._syms = None
        self._parity = None
        self._comp = None

    @classmethod
    def _looks_like_symmetry(cls, a):
        return False

    @classmethod
    def _normalize(cls, segs):
        S = cls.Symmetry
        if isinstance(segs, (int, np.integer)) and not isinstance(segs, bool):
            return _Segments(((S(), int(segs)),))
        out = []
        for item in segs:
            if isinstance(item, S):
                out.append((item, 1))
            elif isinstance(item, (tuple, list)) and len(item) == 2 and cls._is_pair(item):
                out.append((S(item[0]), int(item[1])))
            else:
                out.append((S(item), 1))
        return _Segments(out)

    @classmethod
    def _is_pair(cls, item):
        """Distinguish (symmetry, dim) from a bare 2-component symmetry tuple."""
        S = cls.Symmetry
        a, b = item
        if isinstance(a, S) or isinstance(a, (tuple, list)):
            return True
        # a is a scalar element: pair iff symmetry has a single component (or none)
        return S.length <= 1

    # -- reference accessors -------------------------------------------------------------------
    @property
    def segments_size(self):
        return len(self.segments)

    @property
    def dimension(self):
        return sum(d for _, d in self.segments)

    total_dimension = dimension

    @property
    def dims(self):
        if self._dims is None:
            self._dims = tuple(d for _, d in self.segments)
        return self._dims

    @property
    def syms(self):
        if self._syms is None:
            self._syms = tuple(s for s, _ in self.segments)
        return self._syms

    @property
    def parities(self):
        if self._parity is None:
            self._parity = tuple(s.parity for s, _ in self.segments)
        return self._parity

    def components(self):
        """int64 array [ncomp, nseg] of the symmetry components (Z2 as 0/1)."""
        if self._comp is None:
            n = self.Symmetry.length
            self._comp = np.array([[int(s[c]) for s, _ in self.segments] for c in range(n)], dtype=np.int64).reshape(n, len(self.segments))
        return self._comp

    def position_by_symmetry(self, symmetry):
        symmetry = self.Symmetry(symmetry)
        for i, (s, _) in enumerate(self.segments):
            if s == symmetry:
                return i
        raise RuntimeError("No such symmetry in this edge")

    def find_by_symmetry(self, symmetry):
        """Position or None (first match, as edge.hpp:156-169)."""
        for i, (s, _) in enumerate(self.segments):
            if tuple.__eq__(s, symmetry):
                return i
        return None

    def dimension_by_symmetry(self, symmetry):
        return self.segments[self.position_by_symmetry(symmetry)][1]

    def coord_by_point(self, point):
        s, o = point
        return (self.position_by_symmetry(s), o)

    def point_by_coord(self, coord):
        p, o = coord
        return (self.segments[p][0], o)

    def coord_by_index(self, index):
        off = int(index)
        for p, (_, d) in enumerate(self.segments):
            if off < d:
                return (p, off)
            off -= d
        raise RuntimeError("Index is more than edge total dimension")

    def index_by_coord(self, coord):
        return self.index_by_point(self.point_by_coord(coord))

    def point_by_index(self, index):
        return self.point_by_coord(self.coord_by_index(index))

    def index_by_point(self, point):
        s, o = point
        s = self.Symmetry(s)
        r = int(o)
        for sym, d in self.segments:
            if sym == s:
                return r
            r += d
        raise RuntimeError("The symmetry not found in this edge")

    def conjugate(self):
        return type(self)(tuple((-s, d) for s, d in self.segments), (not self.arrow) if self.Symmetry.is_fermi_symmetry else False)

    conjugated = conjugate

    def reversed(self):
        """Same segments, flipped arrow."""
        return type(self)(self.segments, not self.arrow)

    def with_arrow(self, arrow):
        if bool(arrow) == self.arrow:
            return self
        return type(self)(self.segments, arrow)

    def __eq__(self, other):
        if not isinstance(other, Edge):
            try:
                other = type(self)(other)
            except Exception:
                return NotImplemented
        return type(self) is type(other) and self.segments == other.segments and self.arrow == other.arrow

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __hash__(self):
        if self._hash is None:
            self._hash = hash((self.Symmetry.short_name, self.segments, self.arrow))
        return self._hash

    def __repr__(self):
        S = self.Symmetry
        if S.length == 0:
            return f"{S.short_name}Edge[{self.dimension}]"
        return f"{S.short_name}Edge{self}"

    def __str__(self):
        S = self.Symmetry
        if S.length == 0:
            return str(self.dimension)
        body = "{" + ",".join(f"{s}:{d}" for s, d in self.segments) + "}"
        if S.is_fermi_symmetry:
            return "{arrow:" + ("1" if self.arrow else "0") + ",segment:" + body + "}"
        return body

    def __reduce__(self):
        return (type(self), (tuple((tuple(s), d) for s, d in self.segments), self.arrow))


class BlockTable:
    """Block layout of a tensor with the given edges (core.hpp:152-191).

    positions : int64 [nblock, rank]   segment position of each existing block, row-major order
    dims      : int64 [nblock, rank]
    offsets   : int64 [nblock]         element offset of each block in the contiguous storage
    sizes     : int64 [nblock]
    size      : total storage elements
    """

    __slots__ = ("edges", "rank", "positions", "dims", "offsets", "sizes", "size", "_index", "uid")
    _next_uid = 0

    def __init__(self, edges):
        self.edges = edges
        rank = self.rank = len(edges)
        BlockTable._next_uid += 1
        self.uid = BlockTable._next_uid
        if rank == 0:
            self.positions = np.zeros((1, 0), dtype=np.int64)
            self.dims = np.zeros((1, 0), dtype=np.int64)
        else:
            S = edges[0].Symmetry
            shape = tuple(e.segments_size for e in edges)
            if 0 in shape:
                self.positions = np.zeros((0, rank), dtype=np.int64)
            elif S.length == 0:
                self.positions = np.zeros((1, rank), dtype=np.int64)
            else:
                mask = None
                for c in range(S.length):
                    total = None
                    for i, e in enumerate(edges):
                        comp = e.components()[c].reshape([-1 if j == i else 1 for j in range(rank)])
                        total = comp if total is None else total + comp
                    if S.kinds[c] == "Z2":
                        total = total & 1
                    m = np.broadcast_to(total == 0, shape)
                    mask = m if mask is None else (mask & m)
                self.positions = np.argwhere(mask).astype(np.int64).reshape(-1, rank)
            dims = np.empty_like(self.positions)
            for i, e in enumerate(edges):
                d = np.asarray(e.dims, dtype=np.int64)
                dims[:, i] = d[self.positions[:, i]] if len(d) else 0
            self.dims = dims
        self.sizes = np.prod(self.dims, axis=1).astype(np.int64) if self.dims.shape[1] else np.ones(len(self.dims), dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)[:-1]]).astype(np.int64) if len(self.sizes) else np.zeros(0, dtype=np.int64)
        self.size = int(self.sizes.sum())
        self._index = None

    @property
    def index(self):
        if self._index is None:
            self._index = {tuple(int(x) for x in p): i for i, p in enumerate(self.positions)}
        return self._index

    def block_by_positions(self, positions):
        return self.index.get(tuple(int(p) for p in positions))


_TABLES: dict = {}


def block_table(edges) -> BlockTable:
    edges = tuple(edges)
    t = _TABLES.get(edges)
    if t is None:
        t = _TABLES[edges] = BlockTable(edges)
    return t

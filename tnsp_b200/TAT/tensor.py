"""Device-backed block-symmetric tensor with the PyTAT ``Tensor`` interface.

Mirrors the class bound by the reference at PyTAT/PyTAT.hpp:569-1147 (same method names, argument
meaning and error behaviour).  Storage is one float64 device buffer of shape [nb, size] where
``nb`` is the number of Monte-Carlo chains that share this tensor's block structure (nb == 1 is an
ordinary tensor; every operation then is the reference operation applied to all chains at once by
a single kernel launch).  All heavy work goes through ``tnsp_b200.backend`` -> C-ABI -> CUDA.
"""
from __future__ import annotations

import numpy as np

from .. import backend as _bk
from . import plan as _plan
from . import structure as _structure
from .structure import block_table

_PLAN_CACHE: dict = {}
STATS = {"contract": 0, "svd": 0, "qr": 0, "pack": 0, "flops": 0, "pack_elems": 0}


PLAN_CACHE_MAX = 20000      # plans (host tables + their device copies); a lock-step model needs a few hundred


def _cached(key, builder):
    p = _PLAN_CACHE.get(key)
    if p is None:
        if len(_PLAN_CACHE) >= PLAN_CACHE_MAX:
            # single chains of a symmetric model meet new block structures for ever (per-sample sectors, per-sample cuts): the oldest half
            # goes -- a plan can always be rebuilt
            for k in list(_PLAN_CACHE)[:PLAN_CACHE_MAX // 2]:
                del _PLAN_CACHE[k]
        p = _PLAN_CACHE[key] = builder()
    return p


def _fs(x):
    return frozenset(x) if x else frozenset()


class BatchScalar:
    """A per-chain scalar living on the device (result of norm_* on a batched tensor)."""

    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __float__(self):
        v = self.numpy()
        if v.size != 1:
            raise TypeError("batched scalar with more than one chain cannot be converted to float")
        return float(v[0])

    def _bin(self, other, f):
        o = other.t if isinstance(other, BatchScalar) else other
        return BatchScalar(f(self.t, o))

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __rtruediv__(self, o):
        return self._bin(o, lambda a, b: b / a)

    def __pow__(self, e):
        return BatchScalar(self.t**e)

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)


class _EdgesProxy:
    def __init__(self, tensor):
        self._t = tensor

    def __getitem__(self, i):
        return self._t._edges[i]

    def __len__(self):
        return len(self._t._edges)

    def __iter__(self):
        return iter(self._t._edges)

    def __call__(self, key):
        if isinstance(key, str):
            return self._t.edge_by_name(key)
        return self._t._edges[key]


class _StorageView(np.ndarray):
    """What `tensor.storage` returns: a numpy array of the tensor's elements (PyTAT.hpp:461-506 hands out an ndarray that aliases the
    tensor's memory).  Here the elements live on the device: the array is a host snapshot, and every write through it
    (`storage[...] = x`, in-place arithmetic) is sent back to the tensor."""

    def __new__(cls, tensor):
        host = np.asarray(tensor._host(), dtype=tensor._np)      # (aliases the tensor's memory when the buffers live on the host)
        obj = host.view(cls)
        obj._t = tensor
        return obj

    def __array_finalize__(self, obj):
        self._t = getattr(obj, "_t", None) if obj is not None and getattr(obj, "base", None) is None else None

    def _push(self):
        owner = self
        while isinstance(owner.base, np.ndarray) and getattr(owner.base, "_t", None) is not None:
            owner = owner.base
        if getattr(owner, "_t", None) is not None:
            owner._t._set_host(np.asarray(owner))

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self, key, value)
        if self._t is not None:
            self._push()

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        plain = tuple(np.asarray(x) if isinstance(x, _StorageView) else x for x in inputs)
        if out is not None:
            targets = tuple(np.asarray(o) if isinstance(o, _StorageView) else o for o in out)
            result = getattr(ufunc, method)(*plain, out=targets, **kwargs)
            for o in out:
                if isinstance(o, _StorageView) and o._t is not None:
                    o._push()
            return out[0] if len(out) == 1 else out
        return getattr(ufunc, method)(*plain, **kwargs)


class _BlocksProxy:
    def __init__(self, tensor, readonly):
        self._t = tensor
        self._ro = readonly

    def _locate(self, position):
        t = self._t
        if len(position) and isinstance(position[0], (tuple, list)):
            names = [n for n, _ in position]
            syms = {n: t.Symmetry(s) for n, s in position}
        else:
            names = list(position)
            syms = {n: t.Symmetry() for n in names}
        pos = []
        for n, e in zip(t.names, t._edges):
            p = e.find_by_symmetry(syms[n])
            if p is None:
                raise RuntimeError("No such symmetry in this edge")
            pos.append(p)
        b = t._table.block_by_positions(pos)
        if b is None:
            raise RuntimeError("Try to get a block which does not exist")
        perm = [t.names.index(n) for n in names]
        return b, perm

    def __getitem__(self, position):
        b, perm = self._locate(position)
        t = self._t
        off, size = int(t._table.offsets[b]), int(t._table.sizes[b])
        dims = [int(d) for d in t._table.dims[b]]
        return t._host()[off:off + size].reshape(dims).transpose(perm).copy()

    def __setitem__(self, position, value):
        b, perm = self._locate(position)
        t = self._t
        off, size = int(t._table.offsets[b]), int(t._table.sizes[b])
        dims = [int(d) for d in t._table.dims[b]]
        a = t._host().copy()
        view = a[off:off + size].reshape(dims).transpose(perm)
        view[...] = value
        t._set_host(a)


def _matrix_exponential(a, q):
    """exp(a) as the reference computes it (exponential.hpp:92-150): scale by 2^-j with j = max(0, 1 + int(log2(max|a|))), the
    diagonal (q, q) Pade approximant D^-1 N, then j squarings."""
    n = a.shape[0]
    if n == 0:
        return a.copy()
    biggest = float(np.abs(a).max())
    j = max(0, 1 + int(np.log2(biggest))) if biggest > 0 else 0
    a = a / float(1 << j)
    d, num, x = np.eye(n), np.eye(n), np.eye(n)
    c = 1.0
    for k in range(1, q + 1):
        c = (c * (q - k + 1)) / ((2 * q - k + 1) * k)
        x = a @ x
        num = num + c * x
        d = d + (c if k % 2 == 0 else -c) * x
    f = np.linalg.solve(d, num)
    for _ in range(j):
        f = f @ f
    return f


class Tensor:
    """Created per (symmetry, scalar) by ``TAT/__init__.py``; class attributes: Symmetry, Edge, model, dtype ..."""

    __slots__ = ("names", "_edges", "_table", "_data")
    Symmetry = None
    Edge = None
    model = None
    is_real = True
    is_complex = False
    dtype = "float64"
    btype = "D"
    _np = np.dtype("float64")
    _host_only = False        # S / C / Z scalar types: model-definition constants, host arithmetic (TAT/host_scalars.py)

    @classmethod
    def _B(cls):
        if cls._host_only:
            from . import host_scalars
            return host_scalars.backend(cls.dtype)
        return _bk.get()

    # -- construction --------------------------------------------------------------------------
    def __init__(self, *args, **kwargs):
        if len(args) == 0 and not kwargs:
            self._init([], [], None)
            return
        if len(args) == 1 and not kwargs and isinstance(args[0], Tensor):
            o = args[0]
            self.names, self._edges, self._table, self._data = list(o.names), o._edges, o._table, o._data
            return
        if len(args) >= 1 and isinstance(args[0], str) and not kwargs:
            raise NotImplementedError("text constructor is outside the hot path")
        if (len(args) >= 1 and isinstance(args[0], (int, float, np.number)) and not isinstance(args[0], bool)) or "number" in kwargs:
            number = kwargs.get("number", args[0] if args else 0)
            names = list(kwargs.get("names", args[1] if len(args) > 1 else []))
            syms = list(kwargs.get("edge_symmetry", args[2] if len(args) > 2 else []))
            arrows = list(kwargs.get("edge_arrow", args[3] if len(args) > 3 else []))
            edges = []
            for i in range(len(names)):
                s = self.Symmetry(syms[i]) if i < len(syms) else self.Symmetry()
                ar = bool(arrows[i]) if i < len(arrows) else False
                edges.append(self.Edge(((s, 1),), ar))
            self._init(names, edges, None)
            if self._table.size != 1:
                raise RuntimeError("Invalid symmetries for a rank-0-like tensor")
            self._set_host(np.array([number]))
            return
        names = kwargs.get("names", args[0] if args else [])
        edges = kwargs.get("edges", args[1] if len(args) > 1 else [])
        self._init(list(names), [e if type(e) is self.Edge else self.Edge(e) for e in edges], None)

    def _init(self, names, edges, data):
        if len(names) != len(edges):
            raise RuntimeError("Different Rank in Tensor Construction")
        if len(set(names)) != len(names):
            raise RuntimeError("Duplicated names in Tensor Construction")
        self.names = list(names)
        self._edges = tuple(edges)
        self._table = block_table(self._edges)
        self._data = data

    @classmethod
    def _make(cls, names, edges, table, data):
        t = cls.__new__(cls)
        t.names = list(names)
        t._edges = tuple(edges)
        t._table = table
        t._data = data
        return t

    # -- data helpers --------------------------------------------------------------------------
    @property
    def data(self):
        """device buffer [nb, size] (allocated uninitialised on first use, like the reference)"""
        if self._data is None:
            self._data = self._B().zeros(1, self._table.size)
        return self._data

    @property
    def nb(self):
        return self.data.shape[0]

    def _host(self):
        a = self._B().to_numpy(self.data)
        return a[0] if a.shape[0] == 1 else a

    def _set_host(self, array):
        a = np.asarray(array)
        if a.dtype.kind == "c" and self._np.kind != "c":
            a = a.real
        a = a.astype(self._np)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        if a.shape[1] != self._table.size:
            raise ValueError("storage size mismatch")
        self._data = self._B().from_numpy(a)

    @classmethod
    def from_batch(cls, names, edges, array):
        """Build a batched tensor from a host array [nb, size] (or a device buffer)."""
        edges = tuple(e if type(e) is cls.Edge else cls.Edge(e) for e in edges)
        t = cls._make(names, edges, block_table(edges), None)
        if isinstance(array, np.ndarray):
            t._set_host(array)
        else:
            t._data = array
        return t

    # -- simple accessors ----------------------------------------------------------------------
    @property
    def edges(self):
        return _EdgesProxy(self)

    @property
    def rank(self):
        return len(self.names)

    def edge_by_name(self, name):
        try:
            return self._edges[self.names.index(name)]
        except ValueError:
            raise RuntimeError("No such name in tensor") from None

    @property
    def storage(self):
        return _StorageView(self)

    @storage.setter
    def storage(self, value):
        if isinstance(value, _StorageView):
            value = np.asarray(value)
        a = np.empty(self._host().shape, dtype=self._np)
        a[...] = value
        self._set_host(a)

    @property
    def blocks(self):
        return _BlocksProxy(self, False)

    @property
    def const_blocks(self):
        return _BlocksProxy(self, True)

    def _flat_index(self, position):
        pos, offs = [], []
        for n, e in zip(self.names, self._edges):
            if n not in position:
                raise RuntimeError("Name not found in position map")
            v = position[n]
            if isinstance(v, (tuple, list)):
                s, o = v
                p = e.find_by_symmetry(self.Symmetry(s))
                if p is None:
                    raise RuntimeError("No such symmetry in this edge")
                o = int(o)
            else:
                p, o = e.coord_by_index(int(v))
            pos.append(p)
            offs.append(o)
        b = self._table.block_by_positions(pos)
        if b is None:
            raise RuntimeError("Try to get an element in a block which does not exist")
        dims = self._table.dims[b]
        idx = 0
        for o, d in zip(offs, dims):
            if o >= d:
                raise RuntimeError("Index out of range")
            idx = idx * int(d) + o
        return int(self._table.offsets[b]) + idx

    def __getitem__(self, position):
        v = self._host()[..., self._flat_index(position)]
        if np.ndim(v) == 0:
            return complex(v) if self.is_complex else float(v)
        return v

    def __setitem__(self, position, value):
        a = self._host().copy()
        a[..., self._flat_index(position)] = value
        self._set_host(a)

    def __float__(self):
        if self._table.size == 0:
            return 0.0      # no block satisfies the symmetry: the reference converts such a tensor to zero (tensor.hpp:296-310)
        if self._table.size != 1:
            raise RuntimeError("Try to get the only element of the tensor which contains more than one element")
        v = self._host().reshape(-1)[0]
        if self.is_complex:
            raise TypeError("can't convert complex to float")
        return float(v)

    def __complex__(self):
        if self._table.size == 0:
            return 0j
        if self._table.size != 1:
            raise RuntimeError("Try to get the only element of the tensor which contains more than one element")
        return complex(self._host().reshape(-1)[0])

    def scalar(self):
        """the single element per chain as a device vector [nb] (batched analogue of float(tensor))"""
        if self._table.size != 1:
            raise RuntimeError("Try to get the only element of the tensor which contains more than one element")
        return BatchScalar(self.data[:, 0])

    def __repr__(self):
        return f"{self.btype}{self.Symmetry.short_name}Tensor" + self._shape_str()

    def _shape_str(self):
        return "{names:[" + ",".join(self.names) + "],edges:[" + ",".join(str(e) for e in self._edges) + "]}"

    def __str__(self):
        blocks = []
        h = np.atleast_2d(self._host())[0]
        for b, pos in enumerate(self._table.positions):
            syms = ",".join(str(e.segments[int(p)][0]) for e, p in zip(self._edges, pos))
            off, size = int(self._table.offsets[b]), int(self._table.sizes[b])
            blocks.append("[" + syms + "]:[" + ",".join(repr(float(x)).rstrip("0").rstrip(".") if float(x) == int(x) else repr(float(x))
                                                       for x in h[off:off + size]) + "]")
        if self.Symmetry.length == 0:
            return "{names:[" + ",".join(self.names) + "],edges:[" + ",".join(str(e) for e in self._edges) + "],blocks:" + \
                (blocks[0].split(":", 1)[1] if blocks else "[]") + "}"
        return "{names:[" + ",".join(self.names) + "],edges:[" + ",".join(str(e) for e in self._edges) + "],blocks:{" + ",".join(blocks) + "}}"

    # -- copies / fills ------------------------------------------------------------------------
    def copy(self):
        return self._make(self.names, self._edges, self._table, self.data.clone())

    __copy__ = copy

    def __deepcopy__(self, memo):
        return self.copy()

    def same_shape(self):
        return self._make(self.names, self._edges, self._table, None)

    def zero_(self):
        self._data = self._B().zeros(self.data.shape[0], self._table.size)
        return self

    zero = zero_

    def range_(self, first=0, step=1):
        # the reference fills by repeated addition (now += step), not first + i*step
        inc = np.full(self._table.size, step, dtype=self._np if self.is_complex else np.float64)
        if inc.size:
            inc[0] = first
        self._set_host(np.cumsum(inc))
        return self

    range = range_

    def set_(self, function):
        self._set_host(np.array([function() for _ in range(self._table.size)], dtype=self._np))
        return self

    set = set_

    def map(self, function):
        h = np.atleast_2d(self._host())
        out = np.array([[function(complex(x) if self.is_complex else float(x)) for x in row] for row in h], dtype=self._np)
        r = self.same_shape()
        r._set_host(out)
        return r

    def transform_(self, function):
        self._data = self.map(function)._data
        return self

    transform = transform_

    def randn_(self, mean=0.0, stddev=1.0):
        from . import random as _random
        self._set_host(_random._normal_fill(self._table.size, mean, stddev))
        return self

    randn = randn_

    def rand_(self, min=0.0, max=1.0):
        from . import random as _random
        self._set_host(_random._uniform_fill(self._table.size, min, max))
        return self

    rand = rand_

    def to(self, new_type):
        """scalar type conversion (PyTAT.hpp:611-640): a type object (float / complex), a numpy-style name or the one-letter code"""
        s = new_type if isinstance(new_type, str) else getattr(new_type, "__name__", str(new_type))
        code = {"float": "D", "float64": "D", "D": "D", "complex": "Z", "complex128": "Z", "Z": "Z", "float32": "S", "S": "S",
                "complex64": "C", "C": "C"}.get(s)
        if code is None:
            raise RuntimeError(f"Invalid scalar type {new_type!r} in Tensor.to")
        if code == self.btype:
            return self
        target = getattr(self.model, code).Tensor
        r = target._make(self.names, self._edges, self._table, None)
        r._set_host(np.atleast_2d(self._host()))
        return r

    def sqrt(self):
        return self._make(self.names, self._edges, self._table, self._B().unary(self.data, 0))

    def reciprocal(self):
        return self._make(self.names, self._edges, self._table, self._B().unary(self.data, 1))

    # -- norms ---------------------------------------------------------------------------------
    def _norm(self, kind):
        r = self._B().norm(self.data, kind)
        if r.shape[0] == 1:
            return float(self._B().to_numpy(r)[0])
        return BatchScalar(r)

    def norm_max(self):
        return self._norm(-1)

    def norm_num(self):
        return float(self._table.size)

    def norm_sum(self):
        return self._norm(1)

    def norm_2(self):
        return self._norm(2)

    # -- arithmetic ----------------------------------------------------------------------------
    def _scalar_vec(self, value):
        B = self._B()
        if isinstance(value, BatchScalar):
            return value.t.contiguous()
        return B.from_numpy(np.array([value], dtype=self._np if (self.is_complex or not isinstance(value, complex)) else np.complex128))

    def _aligned(self, other):
        if other.names != self.names:
            other = other.transpose(self.names)
        if other._edges != self._edges:
            raise RuntimeError("Scalar Operator in different Tensor Shape")
        return other

    def _binary(self, other, op, reverse=False):
        B = self._B()
        if isinstance(other, Tensor):
            if other.rank == 0 or self.rank == 0:
                # rank-0 operand acts as a scalar (scalar.hpp:60-78)
                if other.rank == 0 and self.rank != 0:
                    return self._binary(BatchScalar(other.data[:, 0]), op)
                if self.rank == 0 and other.rank != 0:
                    return other._binary(BatchScalar(self.data[:, 0]), op, reverse=True)
            o = self._aligned(other)
            a, b = (o.data, self.data) if reverse else (self.data, o.data)
            return self._make(self.names, self._edges, self._table, B.binary(a, b, op))
        if op in (2, 3) and not reverse:
            return self._make(self.names, self._edges, self._table, B.scale(self.data, self._scalar_vec(other), op - 2))
        if op == 2:
            return self._make(self.names, self._edges, self._table, B.scale(self.data, self._scalar_vec(other), 0))
        # scalar (+,-) tensor or scalar / tensor: broadcast the scalar to a tensor
        vec = self._scalar_vec(other)
        nb = max(vec.shape[0], self.data.shape[0])
        ones = B.from_numpy(np.ones((1, self._table.size), dtype=self._np))
        full = B.scale(ones, vec, 0, nb)
        a, b = (full, self.data) if reverse else (self.data, full)
        return self._make(self.names, self._edges, self._table, B.binary(a, b, op))

    def __add__(self, o):
        return self._binary(o, 0)

    def __radd__(self, o):
        return self._binary(o, 0, True)

    def __sub__(self, o):
        return self._binary(o, 1)

    def __rsub__(self, o):
        return self._binary(o, 1, True)

    def __mul__(self, o):
        return self._binary(o, 2)

    def __rmul__(self, o):
        return self._binary(o, 2, True)

    def __truediv__(self, o):
        return self._binary(o, 3)

    def __rtruediv__(self, o):
        return self._binary(o, 3, True)

    def _inplace(self, o, op):
        self._data = self._binary(o, op)._data
        return self

    def __iadd__(self, o):
        return self._inplace(o, 0)

    def __isub__(self, o):
        return self._inplace(o, 1)

    def __imul__(self, o):
        return self._inplace(o, 2)

    def __itruediv__(self, o):
        return self._inplace(o, 3)

    def __neg__(self):
        return self._make(self.names, self._edges, self._table, self._B().unary(self.data, 2))

    # -- edge operations -----------------------------------------------------------------------
    def edge_rename(self, dictionary):
        for n in dictionary:
            if n not in self.names:
                raise RuntimeError("Name missing in edge_rename")
        return self._make([dictionary.get(n, n) for n in self.names], self._edges, self._table, self.data)

    def _run_pack(self, p, src=None):
        """Apply a PackPlan to this tensor's data -> new device buffer."""
        B = self._B()
        src = self.data if src is None else src
        STATS["pack"] += 1
        if p.identity:
            return src
        STATS["pack_elems"] += p.total * src.shape[0]
        dst = B.empty(src.shape[0], p.dst_size) if p.covers_all else B.zeros(src.shape[0], p.dst_size)
        B.pack(p, src, dst)
        return dst

    def _edge_operator(self, split_map, reversed_names, merge_map, new_names, apply_parity=False, excl_split=(), excl_rb=(),
                       excl_ra=(), excl_merge=()):
        sm = None
        if split_map:
            sm = {}
            for k, v in split_map.items():
                sm[k] = [(n, (e.segments if isinstance(e, _structure.Edge) else self.Edge(e).segments)) for n, e in v]
        key = ("eo", type(self), tuple(self.names), self._edges,
               tuple(sorted((k, tuple(v)) for k, v in sm.items())) if sm else None,
               _fs(reversed_names), tuple(sorted((k, tuple(v)) for k, v in merge_map.items())) if merge_map else None,
               tuple(new_names), bool(apply_parity), _fs(excl_split), _fs(excl_rb), _fs(excl_ra), _fs(excl_merge))
        p = _cached(key, lambda: _plan.edge_operator_plan(self.Edge, tuple(self.names), self._edges, sm, _fs(reversed_names), merge_map,
                                                           list(new_names), apply_parity, _fs(excl_split), _fs(excl_rb), _fs(excl_ra),
                                                           _fs(excl_merge)))
        return self._make(p.names, p.edges, p.table, self._run_pack(p))

    def edge_operator(self, split_map, reversed_names, merge_map, new_names, apply_parity=False, parity_exclude_names_split=(),
                      parity_exclude_names_reverse_before_transpose=(), parity_exclude_names_reverse_after_transpose=(),
                      parity_exclude_names_merge=()):
        return self._edge_operator(split_map, reversed_names, merge_map, new_names, apply_parity, parity_exclude_names_split,
                                   parity_exclude_names_reverse_before_transpose, parity_exclude_names_reverse_after_transpose,
                                   parity_exclude_names_merge)

    def transpose(self, target_names):
        target_names = list(target_names)
        if target_names == self.names:
            return self._make(self.names, self._edges, self._table, self.data)
        return self._edge_operator(None, None, None, target_names)

    def reverse_edge(self, reversed_names, apply_parity=False, parity_exclude_names=()):
        return self._edge_operator(None, reversed_names, None, self.names, apply_parity, (), parity_exclude_names)

    def merge_edge(self, merge, apply_parity=False, parity_exclude_names_merge=(), parity_exclude_names_reverse=()):
        for new_name, olds in merge.items():
            for o in olds:
                if o not in self.names:
                    raise RuntimeError("No such edge in merge map")
        target = []
        for name in reversed(self.names):
            found = False
            for after, before in merge.items():
                if name in before:
                    if name == before[-1]:
                        target.append(after)
                    found = True
                    break
            if not found:
                target.append(name)
        for after, before in merge.items():
            if len(before) == 0:
                target.append(after)
        target.reverse()
        return self._edge_operator(None, None, {k: list(v) for k, v in merge.items()}, target, apply_parity, (), (),
                                   parity_exclude_names_reverse, parity_exclude_names_merge)

    def split_edge(self, split, apply_parity=False, parity_exclude_names_split=()):
        for old in split:
            if old not in self.names:
                raise RuntimeError("No such edge in split map")
        target = []
        for name in self.names:
            if name in split:
                target.extend(n for n, _ in split[name])
            else:
                target.append(name)
        return self._edge_operator(split, None, None, target, apply_parity, parity_exclude_names_split)

    # -- contract ------------------------------------------------------------------------------
    def contract(self, another_tensor, contract_pairs, fuse_names=frozenset()):
        other = another_tensor
        if type(other) is not type(self):
            raise TypeError("contract needs two tensors of the same type")
        B = self._B()
        pairs = frozenset((a, b) for a, b in contract_pairs)
        fuse = _fs(fuse_names)
        key = ("ct", type(self), tuple(self.names), self._edges, tuple(other.names), other._edges, pairs, fuse)
        p = _cached(key, lambda: _plan.contract_plan(self.Edge, tuple(self.names), self._edges, tuple(other.names), other._edges,
                                                      sorted(pairs), fuse))
        STATS["contract"] += 1
        d1, d2 = self.data, other.data
        nb = max(d1.shape[0], d2.shape[0])
        if d1.shape[0] != d2.shape[0] and min(d1.shape[0], d2.shape[0]) != 1:
            raise RuntimeError("contract of two batched tensors with different chain counts")
        STATS["flops"] += p.flops * nb
        if p.gather is not None and B.gather_gemm:
            # dense operands are read in place through offset tables: no merged / transposed copies (csrc/gemm_gather.cu)
            prod = B.empty(nb, p.prod_size)
            B.gemm_gather(p, d1, d2, prod)
            return self._make(p.names, p.edges, p.table, prod)
        m1 = self._run_pack(p.pack1, d1)
        m2 = other._run_pack(p.pack2, d2)
        prod = B.zeros(nb, p.prod_size) if p.zero_fill else B.empty(nb, p.prod_size)
        if len(p.gemm):
            B.gemm(p, m1, m2, prod)
        if p.unpack is not None:
            prod = self._run_pack(p.unpack, prod)
        return self._make(p.names, p.edges, p.table, prod)

    # -- conjugate -----------------------------------------------------------------------------
    def conjugate(self, trivial_metric=False):
        if self.is_complex:
            values = self._B().unary(self.data, 4)
            if self.Symmetry.length == 0:
                return self._make(self.names, self._edges, self._table, values)
            return self._make(self.names, self._edges, self._table, values)._conjugate_structure(trivial_metric)
        if self.Symmetry.length == 0:
            return self._make(self.names, self._edges, self._table, self.data)
        return self._conjugate_structure(trivial_metric)

    def _conjugate_structure(self, trivial_metric):
        key = ("cj", type(self), self._edges, bool(trivial_metric))

        def build():
            edges = tuple(e.conjugate() for e in self._edges)
            signs = _plan.conjugate_signs(self._edges, self._table, trivial_metric)
            blk = np.stack([self._table.offsets, self._table.sizes, signs], axis=1).astype(np.int64) if len(signs) else np.zeros((0, 3), np.int64)
            blk = blk[blk[:, 1] > 0]
            return edges, block_table(edges), blk, bool(signs.any())

        edges, table, blk, any_sign = _cached(key, build)
        data = self._B().block_sign(blk, self.data) if any_sign else self.data
        return self._make(self.names, edges, table, data)

    # -- svd / qr ------------------------------------------------------------------------------
    def svd(self, free_names_u, common_name_u, common_name_v, singular_name_u, singular_name_v, cut=-1):
        B = self._B()
        free_u = _fs(free_names_u)
        remain_cut, relative_cut = (1 << 62), 0.0
        if cut > 0:
            if cut >= 1:
                remain_cut = int(cut)
            else:
                relative_cut = float(cut)
        key = ("svd", type(self), tuple(self.names), self._edges, free_u, common_name_u, common_name_v)
        p = _cached(key, lambda: _plan.svd_plan(self.Edge, tuple(self.names), self._edges, free_u, common_name_u, common_name_v))
        STATS["svd"] += 1
        nb = self.data.shape[0]
        in_place = p.rc_tab is not None and B.factor_in_place(p)   # operand read through offset tables: no merged copy
        merged = self.data if in_place else self._run_pack(p.merge)
        t1 = B.zeros(nb, p.t1_table.size)
        t2 = B.zeros(nb, p.t2_table.size)
        s = B.zeros(nb, max(p.s_total, 1))
        B.svd(p, merged, t1, s, t2, in_place)
        ks = [int(r[2]) for r in p.sectors]
        ns = len(ks)
        if self.Symmetry.length == 0 and relative_cut == 0.0 and nb > 1:
            # one sector, integer cut: structure known without reading the device (exact zeros excepted)
            remain = [min(ks[0], remain_cut)] if ns else []
        elif ns == 0:
            remain = []
        else:
            counts = B.svd_cut(p, s, remain_cut, relative_cut)
            ch = B.to_numpy(counts)
            remain = [int(x) for x in ch.max(axis=0)]
            if nb > 1 and (ch != ch.max(axis=0, keepdims=True)).any():
                B.svd_mask(p, counts, t1, s, t2)
        skey = ("svd2", key, tuple(remain), singular_name_u, singular_name_v)

        def build2():
            pu, pv, s_edges, blocks = _plan.svd_split_plans(self.Edge, p, common_name_u, common_name_v, remain)
            s_table = block_table(s_edges)
            blk = []
            for i, r, sign in blocks:
                sym = p.s_syms[i]
                bidx = s_table.block_by_positions((s_edges[0].find_by_symmetry(-sym), s_edges[1].find_by_symmetry(sym)))
                blk.append((int(p.sectors[i][6]), int(s_table.offsets[bidx]), r, int(sign)))
            return pu, pv, s_edges, s_table, np.array(blk, dtype=np.int64).reshape(-1, 4)

        pu, pv, s_edges, s_table, blk = _cached(skey, build2)
        tu, tv = (t1, t2) if p.flag else (t2, t1)
        u = self._make(pu.names, pu.edges, pu.table, self._run_pack(pu, tu))
        v = self._make(pv.names, pv.edges, pv.table, self._run_pack(pv, tv))
        sd = B.zeros(nb, s_table.size)
        B.diag_scatter(blk, s, sd)
        st = self._make([singular_name_u, singular_name_v], s_edges, s_table, sd)
        return u, st, v

    def qr(self, free_names_direction, free_names, common_name_q, common_name_r):
        B = self._B()
        free = _fs(free_names)
        key = ("qr", type(self), tuple(self.names), self._edges, free_names_direction, free, common_name_q, common_name_r)
        p = _cached(key, lambda: _plan.qr_plan(self.Edge, tuple(self.names), self._edges, free_names_direction, free, common_name_q,
                                                common_name_r))
        STATS["qr"] += 1
        nb = self.data.shape[0]
        in_place = p.rc_tab is not None and B.factor_in_place(p)
        merged = self.data if in_place else self._run_pack(p.merge)
        if merged is self.data and not in_place and B.qr_destroys_input(p):
            merged = merged.clone()
        t1 = B.zeros(nb, p.t1_table.size)
        t2 = B.zeros(nb, p.t2_table.size)
        B.qr(p, merged, t1, t2, in_place)
        p1, p2 = p.extra
        r1 = self._make(p1.names, p1.edges, p1.table, self._run_pack(p1, t1))
        r2 = self._make(p2.names, p2.edges, p2.table, self._run_pack(p2, t2))
        return (r1, r2) if p.flag else (r2, r1)

    # -- identity ------------------------------------------------------------------------------
    def identity_(self, pairs):
        """Set to the identity between the paired edges (identity.hpp:30-136); host-side fill, not on the hot path.  A block is
        touched when every pair carries opposite symmetries; its "diagonal" (equal indices inside every pair) is set to +1, or
        to -1 when bringing every pair into the order (arrow false, arrow true) is an odd permutation of odd-parity edges."""
        pairs = [tuple(p) for p in pairs]
        rank = len(self.names)
        partner = {}
        for a, b in pairs:
            partner[a], partner[b] = b, a
        if len(partner) != rank or any(n not in partner for n in self.names):
            raise RuntimeError("identity_ needs every edge in exactly one pair")
        S = self.Symmetry
        ordered, destination, seen, nxt = [], [0] * rank, set(), 0
        for i, n in enumerate(self.names):
            if i in seen:
                continue
            j = self.names.index(partner[n])
            seen |= {i, j}
            ordered.append((i, j))
            first, second = (i, j) if not self._edges[i].arrow else (j, i)
            destination[first], destination[second] = nxt, nxt + 1
            nxt += 2
        t = self._table
        a = np.zeros(t.size)
        for b, pos in enumerate(t.positions):
            syms = [self._edges[i].segments[int(pos[i])][0] for i in range(rank)]
            dims = [int(d) for d in t.dims[b]]
            if any(not tuple.__eq__(-syms[i], syms[j]) or dims[i] != dims[j] for i, j in ordered):
                continue
            sign = 1.0
            if S.is_fermi_symmetry:
                odd = False
                for i in range(rank):
                    for j in range(i + 1, rank):
                        if destination[i] > destination[j]:
                            odd ^= bool(syms[i].parity) and bool(syms[j].parity)
                sign = -1.0 if odd else 1.0
            if 0 in dims:
                continue
            blockv = np.zeros(dims)
            if not ordered:
                blockv[...] = sign          # rank 0: the identity is the number one
            else:
                grids = np.indices([dims[i] for i, _ in ordered]).reshape(len(ordered), -1)
                index = [None] * rank
                for k, (i, j) in enumerate(ordered):
                    index[i] = index[j] = grids[k]
                blockv[tuple(index)] = sign
            off = int(t.offsets[b])
            a[off:off + blockv.size] = blockv.reshape(-1)
        self._set_host(a)
        return self

    identity = identity_

    # -- out of the hot path -------------------------------------------------------------------
    def _not_on_path(self, *a, **k):
        raise NotImplementedError("this Tensor method is outside the sampling-VMC hot path (SURVEY.md section 8)")

    shrink = expand = _not_on_path

    def exponential(self, pairs, step=8):
        """Tensor exponential over paired edges (reference: TAT/include/TAT/implement/exponential.hpp:155-273, PyTAT default
        step = 8): merge the first / second names of the pairs into a square matrix of sectors (reversing arrow-true first edges,
        parity applied on the second edge only, merge sign on the first group only), exponentiate every sector by scaling and
        squaring of the (step, step) Pade approximant (exponential.hpp:92-150), split back.  Set-up work of simple update (one
        d^2 x d^2 matrix per Hamiltonian term and call of `update`): the sector matrices are exponentiated on the host, the two
        edge operators are the device pack kernel."""
        pairs = [tuple(p) for p in pairs]
        rank = len(self.names)
        if 2 * len(pairs) != rank:
            raise RuntimeError("Invalid pairs in exponential")
        valid = [True] * rank
        merge_1, merge_2, split_1, split_2 = [], [], [], []
        reverse_names, parity_names = set(), set()
        for i in range(rank - 1, -1, -1):
            if not valid[i]:
                continue
            name = self.names[i]
            other = former = None
            for n1, n2 in pairs:
                if n1 == name:
                    other, former = n2, True
                    break
                if n2 == name:
                    other, former = n1, False
                    break
            if other is None or other not in self.names:
                raise RuntimeError("Invalid pairs in exponential")
            j = self.names.index(other)
            valid[j] = False
            (name_1, index_1), (name_2, index_2) = ((name, i), (other, j)) if former else ((other, j), (name, i))
            merge_1.append(name_1)
            merge_2.append(name_2)
            split_1.append((name_1, self._edges[index_1]))
            split_2.append((name_2, self._edges[index_2]))
            if self._edges[index_1].conjugate() != self._edges[index_2]:
                raise RuntimeError("Incompatible edges in exponential")
            if self.Symmetry.is_fermi_symmetry and self._edges[index_1].arrow:
                reverse_names.update((name_1, name_2))
                parity_names.add(name_2)
        for l in (merge_1, merge_2, split_1, split_2):
            l.reverse()
        e1, e2 = "Exponential_1", "Exponential_2"
        merged = self._edge_operator(None, reverse_names, {e1: merge_1, e2: merge_2}, [e1, e2], False, (), parity_names, (), {e2})
        src = np.atleast_2d(merged._host())
        out = np.zeros_like(src)
        t = merged._table
        for b in range(len(t.positions)):
            n, m = (int(d) for d in t.dims[b])
            if n != m:
                raise RuntimeError("Incompatible edges in exponential")
            off = int(t.offsets[b])
            for c in range(src.shape[0]):
                out[c, off:off + n * n] = _matrix_exponential(src[c, off:off + n * n].reshape(n, n), step).reshape(-1)
        result = merged.same_shape()
        result._set_host(out)
        return result._edge_operator({e1: split_1, e2: split_2}, reverse_names, None, merge_1 + merge_2, False, {e2}, parity_names)

    def trace(self, trace_pairs, fuse_names=None):
        """Partial trace over pairs of edges (trace.hpp; the reference's DirectSampling uses it): a contraction
        with the identity between the conjugates of the paired edges -- no kernel of its own.  Bosonic symmetries only (the
        fermionic version carries the parity signs of trace.hpp:120-160 and is not needed on the sweep / ergodic path)."""
        if fuse_names:
            # fuse (trace.hpp:60-118, tensors without symmetry): out[.., x, ..] = in[.., a = x, b = x, ..]; a contraction with the
            # three-index delta for every fused pair, then the ordinary trace of the remaining pairs
            if self.Symmetry.length != 0:
                raise RuntimeError("fuse_names is only supported for tensors without symmetry")
            result = self
            for new_name, (n1, n2) in dict(fuse_names).items():
                d = result.edge_by_name(n1).dimension
                if result.edge_by_name(n2).dimension != d:
                    raise RuntimeError("Cannot fuse two edge with different shape")
                delta = np.zeros((d, d, d))
                idx = np.arange(d)
                delta[idx, idx, idx] = 1.0
                tee = type(self).from_batch(["__fuse_a", "__fuse_b", new_name], [d, d, d], delta.reshape(1, -1))
                result = tee.contract(result, {("__fuse_a", n1), ("__fuse_b", n2)})
            return result.trace(trace_pairs) if trace_pairs else result
        pairs = [tuple(p) for p in trace_pairs]
        if not pairs:
            return self.copy()
        used = [n for p in pairs for n in p]
        if len(set(used)) != len(used) or any(n not in self.names for n in used):
            raise RuntimeError("Invalid trace pairs")
        names, edges, identity_pairs, contract_pairs = [], [], [], set()
        for i, (a, b) in enumerate(pairs):
            ea, eb = self.edge_by_name(a), self.edge_by_name(b)
            if ea.conjugate() != eb:
                raise RuntimeError("Incompatible edge segments in trace")
            na, nb_ = f"__trace_a{i}", f"__trace_b{i}"
            names += [na, nb_]
            edges += [ea.conjugate(), eb.conjugate()]
            identity_pairs.append((na, nb_))
            contract_pairs |= {(a, na), (b, nb_)}
        eye = type(self)(names, edges).identity_(identity_pairs)
        return self.contract(eye, contract_pairs)

    # -- binary wire format (io.hpp:686-760, version 1) -------------------------------------------
    # "TAT" | u16 version = 1 | names: u64 count, (u64 length, bytes)* | edges: u64 count, per edge [u8 arrow if the symmetry is
    # fermionic] u64 segments, (symmetry padded to 8 bytes, u64 dimension)* | storage: u64 count, float64*.
    # A symmetry is the raw libstdc++ std::tuple of the reference: components in REVERSE order, int32 / bool each aligned to
    # its size (checked against dumps of the unmodified reference for all eight symmetry types, tests/test_wire_format.py).
    # The reference pickles a tensor as exactly these bytes (PyTAT.hpp:768-771), so states pickled by either side load on
    # the other once `TAT.install_as_TAT()` has aliased the module tree (SamplingLattice.__getstate__, lattice.py:706-744).
    @classmethod
    def _pack_symmetry(cls, sym):
        out = bytearray(8)
        off = 0
        for v, kind in zip(reversed(tuple(sym)), reversed(cls.Symmetry.kinds)):
            if kind == "Z2":
                out[off] = 1 if v else 0
                off += 1
            else:
                off = (off + 3) & ~3
                out[off:off + 4] = int(v).to_bytes(4, "little", signed=True)
                off += 4
        return bytes(out)

    @classmethod
    def _unpack_symmetry(cls, raw):
        vals, off = [], 0
        for kind in reversed(cls.Symmetry.kinds):
            if kind == "Z2":
                vals.append(raw[off] != 0)
                off += 1
            else:
                off = (off + 3) & ~3
                vals.append(int.from_bytes(raw[off:off + 4], "little", signed=True))
                off += 4
        return cls.Symmetry(*reversed(vals))

    def dump(self):
        """bytes of the reference's binary tensor format (a batched tensor must hold a single chain)"""
        import struct
        if self.nb != 1:
            raise RuntimeError("dump: a lock-step batch has no single-tensor wire format; dump the chains one by one")
        out = [b"TAT", struct.pack("<H", 1), struct.pack("<Q", len(self.names))]
        for n in self.names:
            raw = str(n).encode()
            out += [struct.pack("<Q", len(raw)), raw]
        out.append(struct.pack("<Q", len(self._edges)))
        for e in self._edges:
            if self.Symmetry.is_fermi_symmetry:
                out.append(b"\x01" if e.arrow else b"\x00")
            out.append(struct.pack("<Q", len(e.segments)))
            for sym, dim in e.segments:
                out += [self._pack_symmetry(sym), struct.pack("<Q", int(dim))]
        data = np.ascontiguousarray(self._host(), dtype="<f8").reshape(-1)
        out += [struct.pack("<Q", data.size), data.tobytes()]
        return b"".join(out)

    def load(self, raw):
        """replace this tensor by the one stored in `raw` (io.hpp:718-760); returns self"""
        import struct
        raw = bytes(raw)
        if raw[:3] != b"TAT":
            raise RuntimeError("load: version-0 dumps (TAT < 0.2, old block order) are not supported")
        version, = struct.unpack_from("<H", raw, 3)
        if version != 1:
            raise RuntimeError(f"load: unknown tensor dump version {version}")
        pos = 5

        def u64():
            nonlocal pos
            v, = struct.unpack_from("<Q", raw, pos)
            pos += 8
            return v

        names = []
        for _ in range(u64()):
            n = u64()
            names.append(raw[pos:pos + n].decode())
            pos += n
        edges = []
        for _ in range(u64()):
            arrow = False
            if self.Symmetry.is_fermi_symmetry:
                arrow = raw[pos] != 0
                pos += 1
            segments = []
            for _ in range(u64()):
                sym = self._unpack_symmetry(raw[pos:pos + 8])
                dim, = struct.unpack_from("<Q", raw, pos + 8)
                pos += 16
                segments.append((sym, int(dim)))
            edges.append(self.Edge(segments, arrow))
        self._init(names, edges, None)
        count = u64()
        if count != self._table.size:
            raise RuntimeError("load: storage size does not match the block structure")
        self._set_host(np.frombuffer(raw, dtype="<f8", count=count, offset=pos).copy())
        return self

    def __getstate__(self):
        return self.dump()

    def __setstate__(self, state):
        self.load(state)

    # -- dense embedding (clear_symmetry.hpp: bosonic symmetries only) ----------------------------------
    def _segment_starts(self):
        return [np.concatenate([[0], np.cumsum(e.dims)[:-1]]).astype(np.int64) if len(e.dims) else np.zeros(0, np.int64) for e in self._edges]

    def clear_symmetry(self):
        """NoSymmetry tensor with the same names and total dimensions, exact zeros in the forbidden blocks; a FERMIONIC tensor is
        converted to the parity-symmetry tensor instead (reference: TAT/include/TAT/implement/clear_symmetry.hpp:30-134, PyTAT
        `clear_symmetry`).  Host-side (set-up only)."""
        if self.Symmetry.is_fermi_symmetry:
            return self.clear_fermi_symmetry()
        return self.clear_bose_symmetry()

    def clear_fermi_symmetry(self):
        """FermiZ2 tensor: every edge keeps its arrow and gets the segments (even, total even dimension), (odd, total odd dimension);
        a block lands at the offsets its segments have among the segments of equal parity (clear_symmetry.hpp:69-134)"""
        if not self.Symmetry.is_fermi_symmetry:
            raise RuntimeError("It is invalid to call clear fermi symmetry on a bose symmetry tensor")
        from .. import TAT as _tat
        Z2 = getattr(_tat.FermiZ2, self.btype).Tensor
        edges, offsets = [], []
        for e in self._edges:
            dims, offs = [0, 0], []
            for sym, d in e.segments:
                p = int(bool(sym.parity))
                offs.append((p, dims[p]))
                dims[p] += d
            edges.append(Z2.Edge([(bool(p), dims[p]) for p in (0, 1) if dims[p]], e.arrow))
            offsets.append(offs)
        result = Z2(list(self.names), edges).zero_()
        h = np.atleast_2d(self._host())
        nb = h.shape[0]
        out = np.zeros((nb, result._table.size), dtype=self._np)
        rt = result._table
        for b, pos in enumerate(self._table.positions):
            bd = [int(d) for d in self._table.dims[b]]
            if 0 in bd:
                continue
            par = [offsets[i][int(p)][0] for i, p in enumerate(pos)]
            start = [offsets[i][int(p)][1] for i, p in enumerate(pos)]
            rpos = [result._edges[i].find_by_symmetry(result.Symmetry(bool(par[i]))) for i in range(len(pos))]
            rb = rt.block_by_positions(rpos)
            rd = [int(d) for d in rt.dims[rb]]
            view = out[:, int(rt.offsets[rb]):int(rt.offsets[rb]) + int(rt.sizes[rb])].reshape([nb] + rd)
            sl = (slice(None),) + tuple(slice(start[i], start[i] + bd[i]) for i in range(len(pos)))
            off, size = int(self._table.offsets[b]), int(self._table.sizes[b])
            view[sl] = h[:, off:off + size].reshape([nb] + bd)
        result._set_host(out)
        return result

    def clear_bose_symmetry(self):
        from .. import TAT as _tat
        No = getattr(_tat.No, self.btype).Tensor
        dims = [e.dimension for e in self._edges]
        h = np.atleast_2d(self._host())
        nb = h.shape[0]
        dense = np.zeros([nb] + dims, dtype=self._np)
        starts = self._segment_starts()
        for b, pos in enumerate(self._table.positions):
            bd = [int(d) for d in self._table.dims[b]]
            off, size = int(self._table.offsets[b]), int(self._table.sizes[b])
            sl = (slice(None),) + tuple(slice(int(starts[i][int(p)]), int(starts[i][int(p)]) + bd[i]) for i, p in enumerate(pos))
            dense[sl] = h[:, off:off + size].reshape([nb] + bd)
        return No.from_batch(list(self.names), [No.Edge(d) for d in dims], dense.reshape(nb, -1))


    def fill_from_dense(self, dense):
        """inverse of clear_symmetry: read this tensor's blocks out of a dense array / NoSymmetry tensor
        with the same names (entries outside the blocks are dropped = projection onto the symmetric sector)"""
        if isinstance(dense, Tensor):
            if dense.names != self.names:
                dense = dense.transpose(self.names)
            dense = np.atleast_2d(dense._host())
        dims = [e.dimension for e in self._edges]
        dense = np.asarray(dense, dtype=np.float64)
        nb = dense.size // max(1, int(np.prod(dims))) if dims else dense.size
        dense = dense.reshape([nb] + dims)
        out = np.zeros((nb, self._table.size))
        starts = self._segment_starts()
        for b, pos in enumerate(self._table.positions):
            bd = [int(d) for d in self._table.dims[b]]
            off, size = int(self._table.offsets[b]), int(self._table.sizes[b])
            sl = (slice(None),) + tuple(slice(int(starts[i][int(p)]), int(starts[i][int(p)]) + bd[i]) for i, p in enumerate(pos))
            out[:, off:off + size] = dense[sl].reshape(nb, size)
        self._set_host(out)
        return self

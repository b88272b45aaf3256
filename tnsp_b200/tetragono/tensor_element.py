"""Sparse table <s'|H|s> of a Hamiltonian / observable term.

Same content and ordering as the reference's ``tensor_element`` (tetragono/tetragono/tensor_element.py:20-108):
``element_pool[in_points][out_points] = one-element tensor``, where points are ``(symmetry, index)``
and entries appear in the order the reference enumerates them (blocks row-major over the tensor's
own name order, then elements row-major) -- this order decides which hop ``uniform_int`` selects.

``ElementTable`` is the vectorised form used by the lock-step batch: for every flattened input
configuration the list of connected output configurations and matrix elements.
"""
from __future__ import annotations

import itertools

import numpy as np

_POOL: dict = {}


def tensor_element(tensor):
    key = id(tensor)
    got = _POOL.get(key)
    if got is None or got[0] is not tensor:
        got = _POOL[key] = (tensor, _calculate(tensor))
    return got[1]


def _calculate(tensor):
    result = {}
    names = list(tensor.names)
    where = {n: i for i, n in enumerate(names)}
    rank = len(names)
    body = rank // 2
    Edge = tensor.model.Edge
    Symmetry = tensor.model.Symmetry
    edges = [tensor.edge_by_name(n) for n in names]
    for pos in itertools.product(*[range(len(e.segments)) for e in edges]):
        syms = [e.segments[p][0] for e, p in zip(edges, pos)]
        total = Symmetry()
        for s in syms:
            total = total + s
        if total != Symmetry():
            continue
        block = tensor.const_blocks[[(n, s) for n, s in zip(names, syms)]]
        for idx in itertools.product(*[range(d) for d in block.shape]):
            value = block[idx]
            if value == 0:
                continue
            template = type(tensor)(names, [Edge([s], e.arrow) for s, e in zip(syms, edges)])
            template.storage = [value]
            point = [(syms[i], idx[i]) for i in range(rank)]
            edge_in = tuple((-point[where[f"I{i}"]][0], point[where[f"I{i}"]][1]) for i in range(body))
            edge_out = tuple(point[where[f"O{i}"]] for i in range(body))
            result.setdefault(edge_in, {})[edge_out] = template
    return result


class ElementTable:
    """Vectorised <s'|H|s>: configurations are flattened TOTAL edge indices of the physical edges.

    count[i]        number of connected s' of input configuration i
    targets[i, k]   flattened index of the k-th connected s' (reference enumeration order)
    values[i, k]    matrix element
    """

    def __init__(self, tensor, physics_edges):
        self.tensor = tensor
        self.edges = list(physics_edges)
        self.body = len(self.edges)
        self.dims = [e.dimension for e in self.edges]
        pool = tensor_element(tensor)
        n_in = int(np.prod(self.dims))
        kmax = max([len(v) for v in pool.values()] + [1])
        self.count = np.zeros(n_in, dtype=np.int32)
        self.targets = np.zeros((n_in, kmax), dtype=np.int64)
        self.values = np.zeros((n_in, kmax), dtype=np.float64)
        self.kmax = kmax
        for edge_in, outs in pool.items():
            i = self.flatten([e.index_by_point(p) for e, p in zip(self.edges, edge_in)])
            for k, (edge_out, template) in enumerate(outs.items()):
                self.targets[i, k] = self.flatten([e.index_by_point(p) for e, p in zip(self.edges, edge_out)])
                self.values[i, k] = float(template)
            self.count[i] = len(outs)

    def flatten(self, indices):
        r = 0
        for d, i in zip(self.dims, indices):
            r = r * d + np.asarray(i, dtype=np.int64)
        return r

    def unflatten(self, flat):
        out = []
        flat = np.asarray(flat, dtype=np.int64)
        for d in reversed(self.dims):
            out.append(flat % d)
            flat = flat // d
        return out[::-1]


_TABLES: dict = {}


def element_table(tensor, physics_edges):
    key = id(tensor)
    got = _TABLES.get(key)
    if got is None or got.tensor is not tensor:
        got = _TABLES[key] = ElementTable(tensor, physics_edges)
    return got

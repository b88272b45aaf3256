"""Memoising dependency graph used by the boundary-MPS environments.

Same observable semantics as the reference's ``lazy`` package (lazy_graph/lazy/__init__.py:23-311):
``Root(value)`` holds a value, ``Node(func, *args)`` computes ``func`` of its (possibly lazy)
arguments on demand and caches the result, ``reset`` invalidates everything downstream,
``bool(node)`` tells whether the cache is warm and ``Copy`` clones a graph keeping warm caches.
Own implementation (iterative evaluation, strong downstream lists pruned lazily).
"""
from __future__ import annotations

import weakref


class Node:
    __slots__ = ("_value", "_func", "_args", "_down", "__weakref__")

    def __init__(self, func, *args):
        self._value = None
        self._func = func
        self._args = args
        self._down = []
        me = weakref.ref(self)
        for a in args:
            if isinstance(a, Node):
                a._down.append(me)

    def upstream(self):
        return [a for a in self._args if isinstance(a, Node)]

    def reset(self, value=None):
        """Set (or clear) the cached value and invalidate every node that depends on it."""
        if self._value is value:
            return
        self._value = value
        todo = [self]
        while todo:
            node = todo.pop()
            alive = []
            for ref in node._down:
                d = ref()
                if d is None:
                    continue
                alive.append(ref)
                if d._value is not None:
                    d._value = None
                    todo.append(d)
                else:
                    # still walk: a cold node may have warm descendants only if they were set by hand; not used
                    pass
            node._down = alive

    def __bool__(self):
        return self._value is not None

    def __call__(self):
        if self._value is not None:
            return self._value
        stack = [self]
        while stack:
            node = stack[-1]
            if node._value is not None:
                stack.pop()
                continue
            missing = [a for a in node._args if isinstance(a, Node) and a._value is None]
            if missing:
                stack.extend(missing)
                continue
            node._value = node._func(*[a._value if isinstance(a, Node) else a for a in node._args])
            if node._value is None and len(stack) > 1:
                # an unset Root (or a function returning None) UPSTREAM of the node asked for: its parent would ask for it again for ever
                raise RuntimeError("lazy graph: an upstream node evaluated to None (a Root is not set)")
            stack.pop()
        return self._value


def Root(value=None):
    node = Node(_nothing)
    node._value = value
    return node


def _nothing():
    return None


class Copy:
    """Clone a graph node by node (upstream first); a clone keeps the cached value iff all of its
    upstream clones carry the same values as the originals (lazy/__init__.py:39-85)."""

    __slots__ = ("_map",)

    def __init__(self):
        self._map = {}

    def __call__(self, node):
        got = self._map.get(id(node))
        if got is not None:
            return got[1]
        # clone upstream nodes first, whatever order the caller walks the graph in: a clone must never keep an edge to
        # a node of the ORIGINAL graph (it would silently follow the original's later updates)
        stack = [node]
        while stack:
            cur = stack[-1]
            if id(cur) in self._map:
                stack.pop()
                continue
            missing = [a for a in cur._args if isinstance(a, Node) and id(a) not in self._map]
            if missing:
                stack.extend(missing)
                continue
            self._clone(cur)
            stack.pop()
        return self._map[id(node)][1]

    def _clone(self, node):
        args = tuple(self._map[id(a)][1] if isinstance(a, Node) else a for a in node._args)
        new = Node(node._func, *args)
        valid = True
        for n, o in zip(new.upstream(), node.upstream()):
            if o._value is not n._value:
                valid = False
                break
        if valid:
            new._value = node._value
        self._map[id(node)] = (node, new)  # keep the original alive so ids stay unique
        return new

"""Loading states written by the reference.

A reference checkpoint is `pickle.dump(SamplingLattice)` (utility.py:365-388 `write_to_file`): a dict of the slots of
AbstractState / AbstractLattice / SamplingLattice (abstract_state.py:246-249, abstract_lattice.py:87, lattice.py:704) at
`data_version` 6, whose leaves are pybind objects pickled as their binary dumps: `TAT.<Sym>.Symmetry` (raw tuple bytes,
PyTAT.hpp:175-189), `TAT.<Sym>.Edge` ([arrow] + segments, PyTAT.hpp:295-309) and `TAT.<Sym>.D.Tensor` (io.hpp:686-760).
`load_reference_state` maps those class paths onto this repository's classes with a restricted Unpickler -- nothing else is
importable from the stream -- and rebuilds a device-backed `SamplingLattice`.
"""
from __future__ import annotations

import io
import pickle

import numpy as np

from .. import TAT as _TAT
from .state import SamplingLattice

_SYMS = ("No", "BoseZ2", "BoseU1", "FermiU1", "FermiU1BoseZ2", "FermiU1BoseU1", "FermiZ2", "FermiU1FermiU1")
_RENAMED = {"Fermi": "FermiU1", "Parity": "FermiZ2", "FermiZ2": "FermiU1BoseZ2", "FermiU1": "FermiU1BoseU1", "Z2": "BoseZ2", "U1": "BoseU1"}


def _symmetry_from_bytes(model, raw):
    T = model.D.Tensor
    return T._unpack_symmetry(bytes(raw) + bytes(8))


class _SymmetryShim:
    """pybind's pickle protocol builds the object empty and then calls __setstate__(bytes); a Symmetry of this repository is
    an immutable tuple, so the stream is first read into this shim and converted afterwards"""
    model = None

    def __setstate__(self, state):
        self.value = _symmetry_from_bytes(self.model, state)


class _EdgeShim:
    model = None

    def __setstate__(self, state):
        raw = bytes(state)
        S, pos, arrow = self.model.Symmetry, 0, False
        if S.is_fermi_symmetry:
            arrow, pos = raw[0] != 0, 1
        count = int.from_bytes(raw[pos:pos + 8], "little")
        pos += 8
        segments = []
        for _ in range(count):
            sym = _symmetry_from_bytes(self.model, raw[pos:pos + 8])
            segments.append((sym, int.from_bytes(raw[pos + 8:pos + 16], "little")))
            pos += 16
        self.value = self.model.Edge(segments, arrow)


class _StateShim:
    def __setstate__(self, state):
        self.state = state[1] if isinstance(state, tuple) else state


def _shim(base, model):
    return type(base.__name__, (base,), {"model": model})


class _ReferenceUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        parts = module.split(".")
        if parts[0] == "TAT" and len(parts) >= 2:
            sym = parts[1]
            if sym not in _SYMS:
                raise pickle.UnpicklingError(f"unknown symmetry module {module}")
            model = getattr(_TAT, sym)
            if len(parts) == 2 and name == "Symmetry":
                return _shim(_SymmetryShim, model)
            if len(parts) == 2 and name in ("Edge", "EdgeSegment"):
                return _shim(_EdgeShim, model)
            if len(parts) == 3 and name == "Tensor":
                if parts[2] not in ("D", "float", "float64"):
                    raise pickle.UnpicklingError("only float64 states are device-backed (SURVEY.md section 8: scalar types out of scope)")
                return model.D.Tensor
        if module.startswith("tetragono") and name == "SamplingLattice":
            return _StateShim
        if module.startswith("numpy") and name in ("_reconstruct", "ndarray", "dtype"):
            return getattr(__import__(module, fromlist=[name]), name)
        if (module, name) in (("builtins", "set"), ("builtins", "frozenset"), ("collections", "OrderedDict")):
            return getattr(__import__(module, fromlist=[name]), name)
        raise pickle.UnpicklingError(f"{module}.{name} is not part of a sampling-lattice checkpoint")


def _plain(x):
    if isinstance(x, (_SymmetryShim, _EdgeShim)):
        return x.value
    if isinstance(x, dict):
        return {_plain(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_plain(v) for v in x)
    if isinstance(x, np.ndarray) and x.dtype == object:
        return [_plain(v) for v in x.tolist()]
    return x


def load_reference_state(source):
    """`source`: bytes, a path or a binary file object holding the reference's pickle of a SamplingLattice (data_version 6)."""
    if isinstance(source, (bytes, bytearray)):
        stream = io.BytesIO(source)
    elif isinstance(source, str):
        stream = open(source, "rb")
    else:
        stream = source
    try:
        shim = _ReferenceUnpickler(stream).load()
    finally:
        if isinstance(source, str):
            stream.close()
    if not isinstance(shim, _StateShim):
        raise RuntimeError("not a sampling-lattice checkpoint")
    state = shim.state
    version = state.get("data_version", 0)
    if version != 6:
        raise RuntimeError(f"checkpoint data_version {version}: only version 6 (current reference) is supported; "
                           "re-save the state with the reference first")
    state = {k: _plain(v) for k, v in state.items()}
    lat = SamplingLattice.__new__(SamplingLattice)
    lat.Tensor = state["Tensor"]
    lat.L1, lat.L2 = int(state["L1"]), int(state["L2"])
    lat._physics_edges = [[dict(state["_physics_edges"][l1][l2]) for l2 in range(lat.L2)] for l1 in range(lat.L1)]
    lat._hamiltonians = dict(state["_hamiltonians"])
    lat._total_symmetry = state["_total_symmetry"]
    lat.attribute = dict(state.get("attribute") or {})
    lat._virtual_bond = [[dict(state["_virtual_bond"][l1][l2]) for l2 in range(lat.L2)] for l1 in range(lat.L1)]
    lattice = state["_lattice"]
    lat._lattice = [[lattice[l1][l2] for l2 in range(lat.L2)] for l1 in range(lat.L1)]
    return lat

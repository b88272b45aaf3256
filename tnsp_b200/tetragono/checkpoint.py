"""Reading and writing states in the reference's checkpoint format.

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


class _SimpleUpdateShim(_StateShim):
    pass


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
        if module.startswith("tetragono") and name == "SimpleUpdateLattice":
            return _SimpleUpdateShim
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
    """`source`: bytes, a path or a binary file object holding the reference's pickle of a SamplingLattice (data_version 6), or of a
    SimpleUpdateLattice (simple_update_lattice.py:148-196, data_version 6: site tensors that include their bond environments +
    `_environment_h/_v`), which comes back as this repository's `SimpleUpdateLattice` -- convert it with
    `simple_update_lattice_to_sampling_lattice` to sample it."""
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
    if isinstance(shim, _SimpleUpdateShim):
        from .simple_update import SimpleUpdateLattice
        lat = SimpleUpdateLattice.__new__(SimpleUpdateLattice)
        lat._environment_h = [[state["_environment_h"][l1][l2] for l2 in range(int(state["L2"]) - 1)] for l1 in range(int(state["L1"]))]
        lat._environment_v = [[state["_environment_v"][l1][l2] for l2 in range(int(state["L2"]))] for l1 in range(int(state["L1"]) - 1)]
    else:
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


# ---------------------------------------------------------------------------------------------------
# writing: a pickle the UNMODIFIED reference loads with plain `pickle.load`
# ---------------------------------------------------------------------------------------------------
def _edge_bytes(edge):
    S = type(edge).Symmetry
    T = getattr(_TAT, S.short_name).D.Tensor
    out = [b"\x01" if edge.arrow else b"\x00"] if S.is_fermi_symmetry else []
    out.append(len(edge.segments).to_bytes(8, "little"))
    for sym, dim in edge.segments:
        out += [T._pack_symmetry(sym), int(dim).to_bytes(8, "little")]
    return b"".join(out)


def _symmetry_bytes(sym):
    """sizeof(Symmetry) raw bytes (PyTAT.hpp:175-189): the packed tuple without the padding that follows it inside a segment"""
    T = getattr(_TAT, type(sym).short_name).D.Tensor
    raw = T._pack_symmetry(sym)
    kinds = type(sym).kinds
    if not kinds:
        return raw[:1]
    size, align = 0, 1
    for kind in reversed(kinds):
        w = 1 if kind == "Z2" else 4
        size = (size + w - 1) // w * w + w
        align = max(align, w)
    size = (size + align - 1) // align * align
    return raw[:size]


class _Stub:
    """stands in for a pybind class of the reference while pickling: written as GLOBAL <module> <name>, rebuilt on the other
    side by `copyreg.__newobj__(cls)` + `__setstate__(bytes)` exactly like pybind11's py::pickle"""


def save_reference_state(state, target):
    """Write `state` (a device-backed SamplingLattice) as the reference's `write_to_file` would (utility.py:340-362):
    `pickle.load` in an environment with the reference's TAT + tetragono returns its SamplingLattice, data_version 6."""
    import copyreg
    import sys
    import types

    sym = state.Tensor.Symmetry.short_name
    names = {}
    installed = {}

    def stub(module, name):
        key = (module, name)
        if key not in names:
            parts = module.split(".")
            for depth in range(1, len(parts) + 1):      # the whole dotted chain: pickle imports the module by name
                dotted = ".".join(parts[:depth])
                if dotted not in installed:
                    installed[dotted] = sys.modules.get(dotted)
                    pkg = types.ModuleType(dotted)
                    pkg.__path__ = []
                    sys.modules[dotted] = pkg
                    if depth > 1:
                        setattr(sys.modules[".".join(parts[:depth - 1])], parts[depth - 1], pkg)
            cls = type(name, (_Stub,), {"__module__": module, "__qualname__": name})
            setattr(sys.modules[module], name, cls)
            names[key] = cls
        return names[key]

    tensor_cls = stub(f"TAT.{sym}.D", "Tensor")
    edge_cls = stub(f"TAT.{sym}", "Edge")
    symmetry_cls = stub(f"TAT.{sym}", "Symmetry")
    lattice_cls = stub("tetragono.sampling_lattice.lattice", "SamplingLattice")

    def Carrier(cls, payload):
        inst = object.__new__(cls)          # an instance of the stand-in class: pickle's NEWOBJ wants args[0] is type(obj)
        inst.payload = payload
        return inst

    def carry(x):
        T = state.Tensor
        if isinstance(x, T):
            return Carrier(tensor_cls, x.dump())
        if isinstance(x, T.Edge):
            return Carrier(edge_cls, _edge_bytes(x))
        if isinstance(x, T.Symmetry):
            return Carrier(symmetry_cls, _symmetry_bytes(x))
        if isinstance(x, dict):
            return {carry(k): carry(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(carry(v) for v in x)
        return x

    lattice = np.empty((state.L1, state.L2), dtype=object)
    for l1 in range(state.L1):
        for l2 in range(state.L2):
            lattice[l1, l2] = carry(state[l1, l2])
    payload = {
        "Tensor": tensor_cls, "L1": state.L1, "L2": state.L2,
        "_physics_edges": carry([[dict(state._physics_edges[l1][l2]) for l2 in range(state.L2)] for l1 in range(state.L1)]),
        "_hamiltonians": carry(dict(state._hamiltonians)),
        "_total_symmetry": carry(state._total_symmetry),
        "_site_number": None, "data_version": 6, "attribute": dict(state.attribute),
        "_virtual_bond": carry([[dict(state._virtual_bond[l1][l2]) for l2 in range(state.L2)] for l1 in range(state.L1)]),
        "_lattice": lattice,
    }

    class Writer(pickle.Pickler):
        def reducer_override(self, obj):
            if isinstance(obj, _Stub):
                return copyreg.__newobj__, (type(obj),), obj.payload
            return NotImplemented

    stream = open(target, "wb") if isinstance(target, str) else target
    try:
        Writer(stream, protocol=4).dump(Carrier(lattice_cls, payload))
    finally:
        if isinstance(target, str):
            stream.close()
        for module, old in installed.items():
            if old is None:
                sys.modules.pop(module, None)
            else:
                sys.modules[module] = old


# ---------------------------------------------------------------------------------------------------
# configuration files (utility.py:390-418): int64 header (ranks, ndim, shape...) + one int64 block per rank
# ---------------------------------------------------------------------------------------------------
def write_configurations(config, file_name):
    """every rank writes its block at its offset, rank 0 the header; same bytes as the reference's MPI-IO version"""
    from .. import dist as _dist
    config = np.ascontiguousarray(config, dtype=np.int64)
    rank, size = _dist.rank(), _dist.world_size()
    head = np.array([size, config.ndim, *config.shape], dtype=np.int64)
    if rank == 0:
        with open(file_name, "wb") as f:
            f.write(head.tobytes())
            f.truncate(head.nbytes + size * config.nbytes)
    _dist.barrier()
    with open(file_name, "r+b") as f:
        f.seek(head.nbytes + rank * config.nbytes)
        f.write(config.tobytes())
    _dist.barrier()


def read_configurations(file_name):
    """the block of this rank (or, when the file holds fewer blocks than there are ranks, a uniformly drawn one)"""
    from .. import dist as _dist
    from ..TAT import random as _random
    rank, size = _dist.rank(), _dist.world_size()
    with open(file_name, "rb") as f:
        stored, ndim = np.frombuffer(f.read(16), dtype=np.int64)
        shape = np.frombuffer(f.read(8 * int(ndim)), dtype=np.int64)
        count = int(np.prod(shape))
        if stored >= size:
            choose = rank
        else:
            # fewer blocks than ranks: every rank draws its own block (the reference draws under `seed_differ`, utility.py:398-412:
            # a common base from the shared engine, then a rank-dependent stream)
            base = _random.uniform_int(0, 2**31 - 1)()
            choose = int(np.random.RandomState((base + rank) % 2**31).randint(0, int(stored)))
        f.seek(16 + 8 * int(ndim) + choose * count * 8)
        return np.frombuffer(f.read(count * 8), dtype=np.int64).reshape(tuple(int(x) for x in shape)).copy()

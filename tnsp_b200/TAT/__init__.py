"""Drop-in replacement of the reference Python module ``TAT`` (PyTAT/PyTAT.cpp:58-166) for the
sampling-VMC hot path: same submodule / class tree, tensors live on the B200.

    TAT.{No,BoseZ2,BoseU1,FermiU1,FermiU1BoseZ2,FermiU1BoseU1,FermiZ2,FermiU1FermiU1}
        .Symmetry  .Edge  .EdgeSegment  .{S,D,C,Z}.Tensor   (+ float/complex/float64/... aliases)
    TAT.{Normal,Z2,U1} aliases, TAT.random, TAT.version

Only float64 ("D") tensors are device-backed (the hot path); S / C / Z tensors are host-side constants for model definitions
(TAT/host_scalars.py).
"""
import sys
import types

from . import random  # noqa: F401
from .structure import Edge as _EdgeBase
from .structure import make_symmetry_class
from .tensor import STATS, BatchScalar, Tensor as _TensorBase  # noqa: F401

version = "0.3.17+b200"
__version__ = version
information = "tnsp_b200: B200-native TAT (sampling-VMC hot path)"

_SPEC = {
    "No": [],
    "BoseZ2": [("z2", "Z2", False)],
    "BoseU1": [("u1", "U1", False)],
    "FermiU1": [("fermi", "U1", True)],
    "FermiU1BoseZ2": [("fermi", "U1", True), ("z2", "Z2", False)],
    "FermiU1BoseU1": [("fermi", "U1", True), ("u1", "U1", False)],
    "FermiZ2": [("parity", "Z2", True)],
    "FermiU1FermiU1": [("fermi_0", "U1", True), ("fermi_1", "U1", True)],
}


def _unsupported_scalar(short, sym):
    class Tensor:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"TAT.{sym}.{short}.Tensor: only float64 (D) tensors are device-backed in this build")
    return Tensor


def _build():
    me = sys.modules[__name__]
    for sym_name, comps in _SPEC.items():
        m = types.ModuleType(f"{__name__}.{sym_name}")
        S = make_symmetry_class(sym_name, comps)
        E = type(sym_name + "Edge", (_EdgeBase,), {"__slots__": (), "Symmetry": S})
        m.Symmetry = S
        m.Edge = E
        m.EdgeSegment = E
        for short, dtype, real in (("S", "float32", True), ("D", "float64", True), ("C", "complex64", False), ("Z", "complex128", False)):
            sm = types.ModuleType(f"{__name__}.{sym_name}.{short}")
            import numpy as _np
            T = type("Tensor", (_TensorBase,), {"__slots__": (), "Symmetry": S, "Edge": E, "model": m, "dtype": dtype, "btype": short,
                                                "is_real": real, "is_complex": not real, "_np": _np.dtype(dtype),
                                                # float64 lives on the B200; the other scalar types are host-side constants
                                                "_host_only": short != "D"})
            T.__module__ = sm.__name__
            T.__qualname__ = "Tensor"
            sm.Tensor = T
            setattr(m, short, sm)
            sys.modules[sm.__name__] = sm
        m.float = m.float64 = m.D
        m.float32 = m.S
        m.complex = m.complex128 = m.Z
        m.complex64 = m.C
        setattr(me, sym_name, m)
        sys.modules[m.__name__] = m
    me.Normal = me.No
    me.Z2 = me.BoseZ2
    me.U1 = me.BoseU1


_build()


class _CallableModule(types.ModuleType):
    """`TAT()` returns the build information, like the reference module (PyTAT.hpp:77-84)"""

    def __call__(self):
        return self.information


sys.modules[__name__].__class__ = _CallableModule


def parity(p):
    if p == +1:
        return False
    if p == -1:
        return True
    raise RuntimeError("The parity should be either +1 or -1.")


arrow = parity


def install_as_TAT():
    """Make ``import TAT`` resolve to this module (drop-in for code written against PyTAT)."""
    sys.modules["TAT"] = sys.modules[__name__]
    for k, v in list(sys.modules.items()):
        if k.startswith(__name__ + "."):
            sys.modules["TAT" + k[len(__name__):]] = v
    return sys.modules[__name__]

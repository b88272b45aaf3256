"""``TAT.random``: the reference's global ``std::mt19937_64`` with libstdc++ distributions
(PyTAT/PyTAT.hpp:87-126), reproduced by the host part of the C-ABI so that seeds give the same
streams (and the same ``randn_`` PEPS) as the reference."""

import numpy as np

from .. import backend as _bk

_state = {"rng": None}


def _rng():
    if _state["rng"] is None:
        lib = _bk.host_lib()
        _state["rng"] = lib.tnsp_rng_create_host(1)
        import random as _pyrandom
        lib.tnsp_rng_seed_host(_state["rng"], 0, _pyrandom.SystemRandom().randrange(1 << 32))
    return _state["rng"]


def seed(seed):
    _bk.host_lib().tnsp_rng_seed_host(_rng(), 0, int(seed) & 0xFFFFFFFF)


class _Distribution:
    """what TAT.random.uniform_int / uniform_real / normal return: a callable OBJECT, like the builtin pybind11 hands out
    (PyTAT.hpp:96-126).  The reference stores it as a class attribute and calls `self.random_int()` (tetragono/utility.py:143-153):
    a plain Python function would be bound to the instance there, an object with __call__ is not."""
    __slots__ = ("_draw",)

    def __init__(self, draw):
        self._draw = draw

    def __call__(self):
        return self._draw()


def uniform_int(min=0, max=1):
    lo = np.array([min], dtype=np.int32)
    hi = np.array([max], dtype=np.int32)
    out = np.zeros(1, dtype=np.int32)

    def draw():
        _bk.host_lib().tnsp_rng_uniform_int_host(_rng(), lo.ctypes.data, hi.ctypes.data, None, out.ctypes.data)
        return int(out[0])

    return _Distribution(draw)


def uniform_real(min=0, max=1):
    out = np.zeros(1, dtype=np.float64)

    def draw():
        _bk.host_lib().tnsp_rng_uniform_real_host(_rng(), float(min), float(max), None, out.ctypes.data)
        return float(out[0])

    return _Distribution(draw)


def normal(mean=0, stddev=1):
    # libstdc++'s normal_distribution produces values in pairs and keeps the second one
    buf = []

    def draw():
        if not buf:
            out = np.zeros(2, dtype=np.float64)
            _bk.host_lib().tnsp_rng_normal_host(_rng(), 0, float(mean), float(stddev), 2, out.ctypes.data)
            buf.extend([out[1], out[0]])
        return float(buf.pop())

    return _Distribution(draw)


def _normal_fill(n, mean, stddev):
    out = np.zeros(n, dtype=np.float64)
    if n:
        _bk.host_lib().tnsp_rng_normal_host(_rng(), 0, float(mean), float(stddev), n, out.ctypes.data)
    return out


def _uniform_fill(n, lo, hi):
    out = np.zeros(n, dtype=np.float64)
    one = np.zeros(1, dtype=np.float64)
    lib = _bk.host_lib()
    for i in range(n):
        lib.tnsp_rng_uniform_real_host(_rng(), float(lo), float(hi), None, one.ctypes.data)
        out[i] = one[0]
    return out

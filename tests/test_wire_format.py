"""Binary tensor format (io.hpp:686-760): dumps of this repository load in the UNMODIFIED reference and the
reference's dumps load here, for all eight symmetry types; pickles are those bytes on both sides (PyTAT.hpp:768-771)."""
import pickle

import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from helpers import FERMI, SYMS, describe, make_tensor, rand_edge, storage


def _random_pair(ref, sym, seed):
    rng = np.random.default_rng(seed)
    rank = int(rng.integers(1, 5))
    names = [f"leg{i}" * (1 + i % 2) for i in range(rank)]
    edges = [rand_edge(rng, sym) for _ in range(rank)]
    mine, theirs = make_tensor(TAT, sym, names, edges), make_tensor(ref, sym, names, edges)
    vals = rng.standard_normal(len(storage(theirs)))
    if len(vals):
        mine.storage = vals
        theirs.storage = vals
    return mine, theirs


@pytest.mark.parametrize("sym", SYMS)
@pytest.mark.parametrize("seed", range(4))
def test_dump_load_both_directions(ref_tat, sym, seed):
    mine, theirs = _random_pair(ref_tat, sym, 900 + seed)
    # ours -> reference
    back = getattr(ref_tat, sym).D.Tensor().load(mine.dump())
    d_back, d_ref = describe(back, sym), describe(theirs, sym)
    assert d_back == d_ref
    assert np.array_equal(storage(back), storage(theirs))
    # reference -> ours
    got = getattr(TAT, sym).D.Tensor(["x"], [rand_edge(np.random.default_rng(0), sym)]).load(theirs.dump())
    d_got = describe(got, sym)
    assert d_got == d_ref
    assert np.array_equal(storage(got), storage(theirs))
    # same length, and identical bytes wherever the reference's are defined (its padding bytes are uninitialised memory)
    a, b = mine.dump(), bytes(theirs.dump())
    assert len(a) == len(b)
    assert getattr(TAT, sym).D.Tensor(["x"], [rand_edge(np.random.default_rng(0), sym)]).load(a).dump() == a


@pytest.mark.parametrize("sym", ["No", "BoseU1", "FermiU1BoseU1", "FermiU1FermiU1"])
def test_pickle_is_the_wire_format(ref_tat, sym):
    mine, theirs = _random_pair(ref_tat, sym, 77)
    clone = pickle.loads(pickle.dumps(mine))
    d_clone, d_mine = describe(clone, sym), describe(mine, sym)
    assert d_clone == d_mine and np.array_equal(storage(clone), storage(mine))
    # the state object of the pickle protocol is the dump on both sides
    assert mine.__getstate__() == mine.dump()
    state_ref = theirs.__getstate__()
    fresh = getattr(TAT, sym).D.Tensor.__new__(getattr(TAT, sym).D.Tensor)
    fresh.__setstate__(state_ref)
    d_fresh, d_ref = describe(fresh, sym), describe(theirs, sym)
    assert d_fresh == d_ref and np.array_equal(storage(fresh), storage(theirs))


def test_reference_pickle_loads_after_install_as_TAT(ref_tat):
    """a pickle written by the reference names the class `TAT.BoseU1.D.Tensor`; with the alias installed it resolves here"""
    import sys
    _, theirs = _random_pair(ref_tat, "BoseU1", 5)
    saved = {k: v for k, v in sys.modules.items() if k == "TAT" or k.startswith("TAT.")}
    try:
        # the reference build is loaded under a private module name (oracle/ref.py); give it its real one for pickling
        sys.modules["TAT"], sys.modules["TAT.BoseU1"], sys.modules["TAT.BoseU1.D"] = ref_tat, ref_tat.BoseU1, ref_tat.BoseU1.D
        blob = pickle.dumps(theirs)
        for k in ("TAT", "TAT.BoseU1", "TAT.BoseU1.D"):
            del sys.modules[k]
        TAT.install_as_TAT()
        got = pickle.loads(blob)
    finally:
        for k in [k for k in sys.modules if k == "TAT" or k.startswith("TAT.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    assert isinstance(got, TAT.BoseU1.D.Tensor)
    d_got, d_ref = describe(got, "BoseU1"), describe(theirs, "BoseU1")
    assert d_got == d_ref and np.array_equal(storage(got), storage(theirs))


def test_dump_rejects_batches_and_bad_input():
    t = TAT.No.D.Tensor.from_batch(["a"], [TAT.No.Edge(3)], np.zeros((2, 3)))
    with pytest.raises(RuntimeError):
        t.dump()
    with pytest.raises(RuntimeError):
        TAT.No.D.Tensor(["a"], [3]).load(b"XYZ" + bytes(40))
    good = TAT.No.D.Tensor(["a"], [3]).dump()
    with pytest.raises(RuntimeError):
        TAT.No.D.Tensor(["a"], [3]).load(good[:3] + b"\x07\x00" + good[5:])

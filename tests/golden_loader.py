"""Rebuild a model + reference results from a tests/golden/*.npz fixture (see make_golden.py)."""
import json
import os

import numpy as np

import tnsp_b200.TAT as TAT
from tnsp_b200.tetragono.state import AbstractLattice, AbstractState, SamplingLattice

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(f[:-4] for f in os.listdir(HERE) if f.endswith(".npz") and not f.startswith(("driver_", "gauge_", "direct_", "simple_", "long_", "hamiltonian_", "common_", "model_", "bench_")))
DRIVER_CASES = sorted(f[:-4] for f in os.listdir(HERE) if f.endswith(".npz") and f.startswith("driver_"))


def load(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


def _edge(mod, d):
    S = mod.Symmetry
    return mod.Edge([(S(*s), n) for s, n in d["segments"]], d["arrow"])


def tensor_from(mod, d, z):
    t = mod.D.Tensor(d["names"], [_edge(mod, e) for e in d["edges"]])
    t.storage = z[d["storage"]]
    return t


def build_lattice(meta, z):
    mod = getattr(TAT, meta["symmetry"])
    state = AbstractState(mod.D.Tensor, meta["L1"], meta["L2"])
    state.total_symmetry = mod.Symmetry(*meta["total_symmetry"])
    for l1 in range(meta["L1"]):
        for l2 in range(meta["L2"]):
            for o, e in meta["physics_edges"][l1][l2].items():
                state.physics_edges[l1, l2, int(o)] = _edge(mod, e)
    for h in meta["hamiltonians"]:
        state._hamiltonians[tuple(tuple(p) for p in h["positions"])] = tensor_from(mod, h["tensor"], z)
    lat = AbstractLattice(state)
    for l1 in range(meta["L1"]):
        for l2 in range(meta["L2"]):
            d = meta["sites"][l1][l2]
            for n, e in zip(d["names"], d["edges"]):
                if n in "UDLR":
                    lat._virtual_bond[l1][l2][n] = _edge(mod, e)
    TAT.random.seed(0)
    lat = SamplingLattice(lat)
    for l1 in range(meta["L1"]):
        for l2 in range(meta["L2"]):
            t = tensor_from(mod, meta["sites"][l1][l2], z)
            assert lat[l1, l2].names == t.names and lat[l1, l2]._edges == t._edges   # same structure as the reference built
            lat[l1, l2] = t
    return lat


def config_points(meta):
    mod = getattr(TAT, meta["symmetry"])
    return [[{int(o): (mod.Symmetry(*p[0]), p[1]) for o, p in meta["config"][l1][l2].items()} for l2 in range(meta["L2"])]
            for l1 in range(meta["L1"])]

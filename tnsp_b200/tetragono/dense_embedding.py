"""Charge-dense embedding of a bosonic-symmetric PEPS for the lock-step batch engine.

In a symmetric model the sampled physical charge of every site sits in the dim-1 ``P_l1_l2_o`` edges of
the environment tensors (lattice.py:319-339), so the block structure of every boundary / strip tensor
depends on the chain's configuration and chains of one batch never share a structure.  The union of
all those structures is the dense one: each site tensor is expanded with ``clear_symmetry``
(clear_symmetry.hpp) to a NoSymmetry tensor with exact zeros in the forbidden blocks.  Zeros stay
exact through every kernel (0*x in the GEMMs, no Jacobi rotation between exactly orthogonal columns,
Householder vectors supported on one sector), and the top-Dc cut of the dense singular values is the
greedy cross-sector cut of svd.hpp:452-470, so amplitudes, energies and gradients equal those of the
symmetric evaluation to rounding.  The factorisation kernels recover the sector structure on the
device from the zero pattern (csrc/factor.cu, "discovered sectors").
"""
from __future__ import annotations

import numpy as np

from .. import TAT
from .. import backend as _bk
from .state import AbstractLattice, AbstractState, SamplingLattice


_SAVED: dict = {}


def release_backend_flags():
    """undo the process-wide backend switches `embed_lattice` sets (sector discovery, zero-fragment skipping): a truly dense model
    run later in the same process must not take the block-sparse paths.  The optimisation driver calls this when it returns."""
    B = _bk.get()
    if "sector_discovery" in _SAVED:
        B.sector_discovery = _SAVED.pop("sector_discovery")
    if _SAVED.pop("skip", False) and hasattr(B, "lib") and hasattr(B.lib, "tnsp_gemm_skip_zero_fragments"):
        B.lib.tnsp_gemm_skip_zero_fragments(0)


def embed_lattice(lattice, NoTensor=None):
    """NoSymmetry SamplingLattice with the same PEPS (dense site tensors), physical edges and
    Hamiltonian terms as the symmetric `lattice`."""
    No = NoTensor if NoTensor is not None else TAT.No.D.Tensor
    # from here on every single-descriptor QR / SVD finds its symmetry sectors on the device (zero pattern)
    B = _bk.get()
    if "sector_discovery" not in _SAVED:
        _SAVED["sector_discovery"] = getattr(B, "sector_discovery", False)
    B.sector_discovery = True
    # ... and the dense contractions test their operand fragments for the exact zeros of charge conservation (block-sparse
    # operands: 13 % of the fragment pairs of cfg2's heaviest contraction are non-zero); truly dense models keep it off
    if hasattr(B, "lib") and hasattr(B.lib, "tnsp_gemm_skip_zero_fragments"):
        B.lib.tnsp_gemm_skip_zero_fragments(2)      # coarse tests: marginally the fastest of the four variants (profiles/r01_s9_mb_gemm_sparse.txt)
        _SAVED["skip"] = True
    state = AbstractState(No, lattice.L1, lattice.L2)
    for (l1, l2, orbit), edge in lattice.physics_edges:
        state.physics_edges[l1, l2, orbit] = edge.dimension
    dense_terms = {}
    for positions, h in lattice._hamiltonians.items():
        d = dense_terms.get(id(h))
        if d is None:
            d = dense_terms[id(h)] = h.clear_symmetry()
        state._set_hamiltonian(positions, d)
    state.attribute = dict(lattice.attribute)
    abstract = AbstractLattice(state)
    for l1, l2 in lattice.sites():
        for direction in ("R", "D"):
            if direction in lattice._virtual_bond[l1][l2] and lattice._virtual_bond[l1][l2][direction] is not None:
                abstract.virtual_bond[l1, l2, direction] = lattice._virtual_bond[l1][l2][direction].dimension
    dense = SamplingLattice.__new__(SamplingLattice)
    dense._init_by_copy(abstract)
    dense._lattice = [[None] * lattice.L2 for _ in range(lattice.L1)]
    for l1, l2 in lattice.sites():
        t = lattice[l1, l2].clear_symmetry()
        if "T" in t.names:   # the dim-1 total-symmetry edge carries no index in the dense picture
            keep = [n for n in t.names if n != "T"]
            t = t.transpose(keep + ["T"])
            t = type(t).from_batch(keep, t._edges[:-1], t.data)
        dense._lattice[l1][l2] = t
    return dense


def embed_configuration(lattice, points):
    """edge points of the symmetric lattice -> total physical indices [L1, L2, orbits] of the dense one"""
    max_orbit = max(orbit for (l1, l2, orbit), _ in lattice.physics_edges)
    out = np.zeros((lattice.L1, lattice.L2, max_orbit + 1), dtype=np.int64) - 1
    for (l1, l2, orbit), edge in lattice.physics_edges:
        sym, off = points[l1][l2][orbit]
        out[l1, l2, orbit] = edge.index_by_point((sym, int(np.asarray(off).reshape(-1)[0])))
    return out


def project_gradient(lattice, dense_gradient):
    """dense per-site gradient tensors -> tensors shaped like the symmetric site tensors (entries outside
    the blocks are exactly zero up to rounding and are dropped)"""
    out = [[None] * lattice.L2 for _ in range(lattice.L1)]
    for l1, l2 in lattice.sites():
        target = lattice[l1, l2].same_shape()
        g = dense_gradient[l1][l2]
        names = [n for n in target.names if n != "T"]
        g = g.transpose(names) if g.names != names else g
        out[l1][l2] = target.fill_from_dense(np.atleast_2d(np.asarray(g._host())))
    return out

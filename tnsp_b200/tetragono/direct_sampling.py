"""Direct sampling: configurations drawn site by site from the conditional distribution of |psi|^2, the
reference's default `sampling_method` (sampling_lattice/sampling.py:252-371).  A proposal-free sampler: every call returns an
independent configuration together with its (approximate) probability, which the Observer uses for reweighting.

Two environments are needed, both built from the existing TAT operations (no kernels of their own):
* the norm <psi|psi> of the rows BELOW the row being sampled as a double-layer boundary MPS
  (auxiliaries/double_layer_auxiliaries.py:387-392, 416-496): `double_layer_rows_from_below`;
* for the row being sampled a five-line column transfer -- sampled boundary above (ket, bra), the row itself (ket, bra) and the
  double-layer boundary below -- compressed column by column (auxiliaries/three_line_auxiliaries.py:60-160): `_FiveLineRow`.
The truncation sequences follow the reference step for step, because with a finite cut the sampled probabilities depend on them
(parity: tests/test_direct_sampling.py compares trajectories and probabilities with the unmodified reference).

One chain per call (`nb = 1`): the conditional probabilities are read back site by site, a lock-step batch would gain nothing
for symmetric tensors whose structure changes with every sampled charge.  All symmetry types, fermionic ones included.
"""
from __future__ import annotations

import numpy as np

from ..TAT import random as _random
from .configuration import Configuration
from .sampling import Sampling


def _contract(t1, t2, pairs, physics=False):
    """contract on the pairs whose names exist on both sides; `physics` adds every shared name that starts with "P"
    (utility.py:340-358, `contract_all_physics_edges`)"""
    n1, n2 = t1.names, t2.names
    wanted = {(a, b) for a, b in pairs if a in n1 and b in n2}
    if physics:
        wanted |= {(n, n) for n in n1 if n.startswith("P") and n in n2}
    return t1.contract(t2, wanted)


def _rename(t, name_map):
    names = t.names
    return t.edge_rename({k: v for k, v in name_map.items() if k in names})


def _compress(sites, left, right, cut, normalize):
    """QR sweep along `right`, truncated SVD sweep back; on entry the bonds are doubled (left + "1" / "0", right + "1" / "0")"""
    left1, left0, right1, right0 = left + "1", left + "0", right + "1", right + "0"
    n = len(sites)
    for i in range(n - 1):
        q, r = sites[i].qr("r", {name for name in (right1, right0) if name in sites[i].names}, right, left)
        sites[i] = q
        sites[i + 1] = _contract(sites[i + 1], r, {(left1, right1), (left0, right0)})
    for i in range(n - 1, 0, -1):
        u, s, v = sites[i].svd({left}, right, left, left, right, cut)
        if normalize:
            s /= s.norm_sum()
        sites[i] = v
        sites[i - 1] = _contract(_contract(sites[i - 1], u, {(right, left)}), s, {(right, left)})
    return sites


def double_layer_rows_from_below(owner, cut, normalize=True):
    """rows[l1][l2], l1 = L1 .. 0: boundary MPS (edges L, R, UN, UC) of <psi|psi> restricted to the rows l1 .. L1-1; every row is
    absorbed in two stages, ket layer then bra layer, each followed by a compression to `cut`"""
    L1, L2 = owner.L1, owner.L2
    one = owner.Tensor(1)
    rows = {L1: [one] * L2}
    for l1 in range(L1 - 1, -1, -1):
        below = rows[l1 + 1]
        ket = [owner[l1, l2] for l2 in range(L2)]
        bra = [t.conjugate() for t in ket]
        stage = [_contract(_rename(below[i], {"L": "L1", "R": "R1"}), _rename(ket[i], {"L": "L0", "R": "R0", "U": "UN"}), {("UN", "D")})
                 for i in range(L2)]
        stage = _compress(stage, "L", "R", cut, normalize)
        stage = [_contract(_rename(stage[i], {"L": "L1", "R": "R1"}), _rename(bra[i], {"L": "L0", "R": "R0", "U": "UC"}),
                           {("T", "T"), ("UC", "D")}, physics=True) for i in range(L2)]
        rows[l1] = _compress(stage, "L", "R", cut, normalize)
    return rows


class _FiveLineRow:
    """The row being sampled between the sampled boundary above and the double-layer boundary below.  A column holds five
    tensors (top ket, row ket, bottom double layer, row bra, top bra); `absorb` pushes a column into an accumulated environment
    of the same five-piece shape and compresses it, `hole` closes everything except the physical legs of one site."""

    def __init__(self, L2, Tensor, cut):
        self.L2, self.cut = L2, cut
        self.ones = [Tensor(1)] * 5
        self.columns = [None] * L2

    def set_column(self, l2, top, site, bottom):
        self.columns[l2] = [top, site, bottom, site.conjugate(), top.conjugate()]

    def set_site(self, l2, site):
        col = self.columns[l2]
        col[1], col[3] = site, site.conjugate()

    def absorb(self, env, column, left, right):
        """environment `env` grows by `column` in the direction left -> right (names: ("L", "R") or ("R", "L"))"""
        right_n, right_c = right + "N", right + "C"
        keep = {right} if right in column[0].names else set()
        # ket half: top boundary tensor split by a QR, the rest joined with the row's ket tensor
        top_n, rest_n = column[0].qr("q", keep, "D", "U")
        ket = _contract(_rename(rest_n, {"D": "D2"}), _rename(env[0], {"D": "D1"}), {(left, right)})
        ket = _contract(ket, _rename(env[1], {"D": "D1"}), {("D1", "U")})
        ket = _contract(ket, _rename(column[1], {"D": "D2"}), {(right, left), ("D2", "U")})
        top_c, rest_c = column[4].qr("q", keep, "D", "U")
        bra = _contract(_rename(rest_c, {"D": "D2"}), _rename(env[4], {"D": "D1"}), {(left, right)})
        bra = _contract(bra, _rename(env[3], {"D": "D1"}), {("D1", "U")})
        bra = _contract(bra, _rename(column[3], {"D": "D2"}), {(right, left), ("D2", "U")})
        # close over the double-layer boundary below
        core = _contract(_rename(ket, {right: right_n, "U": "UN"}), _rename(env[2], {"UC": "UC1"}), {("D1", "UN")})
        core = _contract(core, _rename(column[2], {"UC": "UC2"}), {("D2", "UN"), (right, left)})
        core = _contract(core, _rename(bra, {right: right_c, "U": "UC"}), {("UC1", "D1"), ("UC2", "D2"), ("T", "T")}, physics=True)
        # split the ket and the bra piece off again, each truncated to the cut
        u, s, v = core.svd({"UN", right_n}, "D", "UN", "UN", "D", self.cut)
        row_n = _rename(u, {"UN": "U", right_n: right})
        core = _contract(v, s, {("UN", "D")})
        u, s, v = core.svd({"UC", right_c}, "D", "UC", "UC", "D", self.cut)
        row_c = _rename(u, {"UC": "U", right_c: right})
        core = _contract(v, s, {("UC", "D")})
        return [top_n, row_n, core, row_c, top_c]

    def hole(self, l2, left_env, right_env):
        """reduced density matrix of the unsampled orbits of site l2: edges O{orbit} (ket) and I{orbit} (bra)"""
        top, site, bottom, site_c, top_c = self.columns[l2]
        site = _rename(site, {n: "O" + n[1:] for n in site.names if n.startswith("P") and not n.startswith("P_")})
        site_c = _rename(site_c, {n: "I" + n[1:] for n in site_c.names if n.startswith("P") and not n.startswith("P_")})
        line = [top, site, bottom, site_c, top_c]
        r = _contract(_rename(left_env[0], {"D": "D1"}), _rename(line[0], {"D": "D2"}), {("R", "L")})
        r = _contract(r, _rename(right_env[0], {"D": "D3"}), {("R", "L")})
        r = _contract(r, _rename(left_env[1], {"D": "D1"}), {("D1", "U")})
        r = _contract(r, _rename(line[1], {"D": "D2"}), {("D2", "U"), ("R", "L")})
        r = _contract(r, _rename(right_env[1], {"D": "D3"}), {("D3", "U"), ("R", "L")})
        r = _contract(r, _rename(left_env[2], {"UC": "U1"}), {("D1", "UN")})
        r = _contract(r, _rename(line[2], {"UC": "U2"}), {("D2", "UN"), ("R", "L")})
        r = _contract(r, _rename(right_env[2], {"UC": "U3"}), {("D3", "UN"), ("R", "L")})
        r = _contract(r, _rename(left_env[3], {"U": "U1"}), {("U1", "D")})
        r = _contract(r, _rename(line[3], {"U": "U2"}), {("U2", "D"), ("R", "L"), ("T", "T")})
        r = _contract(r, _rename(right_env[3], {"U": "U3"}), {("U3", "D"), ("R", "L")})
        r = _contract(r, _rename(left_env[4], {"U": "U1"}), {("U1", "D")})
        r = _contract(r, _rename(line[4], {"U": "U2"}), {("U2", "D"), ("R", "L"), ("T", "T")}, physics=True)
        r = _contract(r, _rename(right_env[4], {"U": "U3"}), {("U3", "D"), ("R", "L")})
        return r


class DirectSampling(Sampling):
    """`DirectSampling(owner, cut_dimension, restrict_subspace, double_layer_cut_dimension)` as the reference; `__call__`
    returns `(possibility, configuration)`.

    `nb` > 1 (new, models without symmetry): nb independent configurations per call as a lock-step batch -- the environments
    carry a chain axis, every chain draws from its own engine (`rng`, a `ChainRng`), and chain c reproduces the single-chain
    run with the same seed."""

    def __init__(self, owner, cut_dimension, restrict_subspace, double_layer_cut_dimension, *, nb=1, rng=None):
        super().__init__(owner, cut_dimension, restrict_subspace)
        if nb != 1:
            if owner.Tensor.Symmetry.length != 0:
                raise NotImplementedError("a lock-step batch needs one block structure for all chains: symmetric lattices sample "
                                          "one chain per call (or go through dense_embedding)")
            if restrict_subspace is not None:
                raise NotImplementedError("restrict_subspace callbacks are evaluated per chain; use nb=1")
            if rng is None:
                from .sampling import ChainRng
                rng = ChainRng(nb)
                rng.seed_like_reference()
        self.nb, self.rng = nb, rng
        self._double_layer_cut_dimension = double_layer_cut_dimension
        self.refresh_all()

    def refresh_all(self):
        self._below = double_layer_rows_from_below(self.owner, self._double_layer_cut_dimension, True)

    @staticmethod
    def _choice(p, rho):
        i = 0
        for i, r in enumerate(rho):
            p -= r
            if p < 0:
                return i
        return i

    def _draw(self, hole, hole_edge, uniform):
        """(choice index per chain, its probability per chain) from the diagonal of the reduced density matrix, or None when
        some chain has no weight left"""
        alpha = self.owner.attribute.get("alpha", 1)
        if self.nb == 1:
            rho = []
            for symmetry, _ in hole_edge.segments:
                rho.extend(np.diagonal(hole.const_blocks[[("I", -symmetry), ("O", symmetry)]]))
            rho = np.maximum(np.array(rho).real, 0)
            if np.sum(rho) == 0:
                return None
            rho = rho**alpha
            rho = rho / np.sum(rho)
            choice = self._choice(uniform(), rho)
            return choice, rho[choice]
        d = hole_edge.dimension
        rho = np.broadcast_to(np.asarray(hole.storage).reshape(-1, d, d), (self.nb, d, d))     # the first site is the same for all chains
        rho = np.maximum(np.diagonal(rho, axis1=1, axis2=2), 0)
        if (rho.sum(axis=1) == 0).any():
            return None
        rho = rho**alpha
        rho = rho / rho.sum(axis=1, keepdims=True)
        p = self.rng.uniform_real(None)
        choice = np.full(self.nb, d - 1, dtype=np.int64)
        open_ = np.ones(self.nb, dtype=bool)
        for i in range(d):                     # the same sequential subtraction as `_choice`, chain by chain
            p = p - rho[:, i]
            hit = open_ & (p < 0)
            choice[hit] = i
            open_ &= ~hit
        return choice, rho[np.arange(self.nb), choice]

    def __call__(self):
        owner = self.owner
        configuration = Configuration(owner, self._cut_dimension, self.nb)
        uniform = _random.uniform_real(0, 1)
        possibility = 1.0 if self.nb == 1 else np.ones(self.nb)
        for l1 in range(owner.L1):
            row = _FiveLineRow(owner.L2, owner.Tensor, self._cut_dimension)
            for l2 in range(owner.L2):
                row.set_column(l2, configuration._up_to_down_site[l1 - 1, l2](), owner[l1, l2], self._below[l1 + 1][l2])
            # environments right of every site (unsampled sites), then the left one grows as the sites are sampled
            right_env = {owner.L2: row.ones}
            for l2 in range(owner.L2 - 1, 0, -1):
                right_env[l2] = row.absorb(right_env[l2 + 1], row.columns[l2], "R", "L")
            left_env = row.ones
            for l2 in range(owner.L2):
                shrunk = owner[l1, l2]
                config = {}
                shrinkers = configuration._get_shrinker((l1, l2), config)
                site_hole = row.hole(l2, left_env, right_env[l2 + 1])
                unsampled = set(owner.physics_edges[l1, l2])
                for orbit in owner.physics_edges[l1, l2]:
                    unsampled.remove(orbit)
                    hole = (site_hole.trace({(f"I{o}", f"O{o}") for o in unsampled}).edge_rename({f"I{orbit}": "I", f"O{orbit}": "O"})
                            .transpose(["I", "O"]))
                    hole_edge = hole.edge_by_name("O")
                    drawn = self._draw(hole, hole_edge, uniform)
                    if drawn is None:
                        return self()          # block mismatch or vanishing weight: draw again, like the reference
                    choice, weight = drawn
                    possibility = possibility * weight
                    if self.nb == 1:
                        configuration[l1, l2, orbit] = hole_edge.point_by_index(choice)
                    else:
                        configuration[l1, l2, orbit] = Configuration._point_by_index(hole_edge, choice)
                    config[orbit] = configuration[l1, l2, orbit]      # normalised (symmetry, index array) form
                    _, shrinker = next(shrinkers)
                    shrunk = shrunk.contract(shrinker.edge_rename({"P": f"P{orbit}"}), {(f"P{orbit}", "Q")})
                    site_hole = (site_hole.contract(shrinker.edge_rename({"P": f"O{orbit}"}), {(f"O{orbit}", "Q")})
                                 .contract(shrinker.conjugate().edge_rename({"P": f"I{orbit}"}), {(f"I{orbit}", "Q")})
                                 .trace({(f"I{orbit}", f"O{orbit}")}))
                    row.set_site(l2, shrunk)
                if l2 + 1 < owner.L2:
                    left_env = row.absorb(left_env, row.columns[l2], "L", "R")
        if self._restrict_subspace is not None and not self._restrict_subspace(configuration):
            return self()
        return possibility, configuration

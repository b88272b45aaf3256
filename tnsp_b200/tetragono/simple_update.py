"""Simple update (second consumer of contract / qr / svd): imaginary-time evolution of a PEPS with a diagonal
singular-value "environment" tensor on every bond, and the conversions that hand its result to the sampling-VMC path.

Reference: tetragono/tetragono/simple_update_lattice.py (class and `environment` handler :34-137, `update` :252-344, the per-term
updates :372-677, `_try_multiple` :679-722) and tetragono/tetragono/conversion.py:24-68.  Same class name, method names, argument
meaning and errors.  Data convention (the reference's data_version >= 2): a site tensor holds the bare site times the environments
of all its bonds; every bond update divides the shared environment out of ONE of the two sites first.

What differs, on purpose:
* the horizontal / vertical nearest-neighbour updates are one routine parametrised by the bond direction;
* the reference scatters the independent terms of a bundle over MPI ranks and broadcasts the touched tensors; here every rank
  performs every update (replicas only -- identical, deterministic results on all ranks, no exchange step);
* `observe` / `observe_energy` (double-layer environment with a hole, outside the sampling path) raise NotImplementedError: convert
  with `simple_update_lattice_to_sampling_lattice` and measure with the sampling Observer instead.
All tensor work goes through the device-backed TAT operations; `Tensor.exponential` of the small Hamiltonian terms is set-up work.
"""
from .state import AbstractLattice, SamplingLattice

_OPPOSITE = {"L": "R", "R": "L", "U": "D", "D": "U"}


class SimpleUpdateLatticeEnvironment:
    """`lattice.environment[l1, l2, direction]`: the bond tensor next to site (l1, l2); None when unset or outside the lattice
    (simple_update_lattice.py:34-123)."""

    __slots__ = ["owner"]

    def __init__(self, owner):
        self.owner = owner

    def _slot(self, where):
        l1, l2, direction = where
        if direction not in _OPPOSITE:
            raise ValueError("Invalid direction")
        o = self.owner
        if direction in "LR":
            l2 -= direction == "L"
            return (o._environment_h, l1, l2) if 0 <= l1 < o.L1 and 0 <= l2 < o.L2 - 1 else None
        l1 -= direction == "U"
        return (o._environment_v, l1, l2) if 0 <= l1 < o.L1 - 1 and 0 <= l2 < o.L2 else None

    def __getitem__(self, where):
        slot = self._slot(where)
        return None if slot is None else slot[0][slot[1]][slot[2]]

    def __setitem__(self, where, value):
        slot = self._slot(where)
        if slot is None:
            raise ValueError("Environment out of lattice")
        slot[0][slot[1]][slot[2]] = value


class SimpleUpdateLattice(AbstractLattice):
    def __init__(self, abstract):
        self._init_by_copy(abstract)
        if self._virtual_bond is None:
            self._virtual_bond = [[self._default_bonds(l1, l2) for l2 in range(self.L2)] for l1 in range(self.L1)]
        self._lattice = [[self._construct_tensor(l1, l2) for l2 in range(self.L2)] for l1 in range(self.L1)]
        self._environment_h = [[None] * (self.L2 - 1) for _ in range(self.L1)]
        self._environment_v = [[None] * self.L2 for _ in range(self.L1 - 1)]

    def __getitem__(self, l1l2):
        return self._lattice[l1l2[0]][l1l2[1]]

    def __setitem__(self, l1l2, value):
        self._lattice[l1l2[0]][l1l2[1]] = value

    @property
    def environment(self):
        return SimpleUpdateLatticeEnvironment(self)

    # -- driver ----------------------------------------------------------------------------------
    def update(self, total_step, delta_tau, new_dimension):
        """`total_step` second-order Trotter steps (all terms forward, then backward) of exp(-delta_tau H_i); every bond svd keeps
        `new_dimension` values (int) or the values above that relative threshold (float in (0, 1)); simple_update_lattice.py:252-344"""
        updaters = []
        for positions, term in self.hamiltonians:
            coordinates, index_and_orbit = [], []
            for l1, l2, orbit in positions:
                if (l1, l2) not in coordinates:
                    coordinates.append((l1, l2))
                index_and_orbit.append((coordinates.index((l1, l2)), orbit))
            gate = (-delta_tau * term).exponential({(f"I{i}", f"O{i}") for i in range(len(positions))})
            updaters.append((coordinates, index_and_orbit, gate))
        # bundles of mutually independent terms (:287-315).  The order is part of the result: a step runs the bundles forward and
        # then in reverse order, the terms INSIDE a bundle in their original order both times (:325-340)
        bundles = []
        while updaters:
            taken, bundle, rest = set(), [], []
            for item in updaters:
                if any(c in taken for c in item[0]):
                    rest.append(item)
                else:
                    taken.update(item[0])
                    bundle.append(item)
            bundles.append(bundle)
            updaters = rest
        sequence = [item for bundle in bundles for item in bundle] + [item for bundle in reversed(bundles) for item in bundle]
        for _ in range(total_step):
            for coordinates, index_and_orbit, gate in sequence:
                self._single_term_simple_update(coordinates, index_and_orbit, gate, new_dimension)
        for l1, l2 in self.sites():
            if l1 != self.L1 - 1:
                self.virtual_bond[l1, l2, "D"] = self[l1, l2].edge_by_name("D")
            if l2 != self.L2 - 1:
                self.virtual_bond[l1, l2, "R"] = self[l1, l2].edge_by_name("R")

    def _single_term_simple_update(self, coordinates, index_and_orbit, evolution_operator, new_dimension):
        if len(coordinates) == 1:
            orbits = [orbit for _, orbit in index_and_orbit]
            self[coordinates[0]] = (self[coordinates[0]]
                                    .contract(evolution_operator, {(f"P{orbit}", f"I{rank}") for rank, orbit in enumerate(orbits)})
                                    .edge_rename({f"O{rank}": f"P{orbit}" for rank, orbit in enumerate(orbits)}))
            return
        if len(coordinates) == 2:
            (a1, a2), (b1, b2) = coordinates
            if (a1 == b1 and abs(a2 - b2) == 1) or (a2 == b2 and abs(a1 - b1) == 1):
                first_is_0 = (a1, a2) < (b1, b2)  # the site to the left / above goes first
                first = (a1, a2) if first_is_0 else (b1, b2)
                legs = [[(rank, orbit) for rank, (index, orbit) in enumerate(index_and_orbit) if index == which]
                        for which in ((0, 1) if first_is_0 else (1, 0))]
                return self._nearest_neighbour(first, "R" if a1 == b1 else "D", legs, len(index_and_orbit), evolution_operator, new_dimension)
            return self._long_range(coordinates, index_and_orbit, evolution_operator, new_dimension)
        raise NotImplementedError("Unsupported simple update style")

    def _nearest_neighbour(self, first, toward, legs, body, gate, new_dimension):
        """bond `toward` ("R" or "D") of site `first` (:430-572): divide the bond environment out of the second site, reduce both
        sites to their R factors (physical legs of the term + the bond), apply the gate, svd, store the normalised singular
        values as the new environment and multiply them into BOTH new factors."""
        back = _OPPOSITE[toward]
        i, j = first
        second = (i, j + 1) if toward == "R" else (i + 1, j)
        legs_1, legs_2 = legs
        site_1 = self[first]
        site_2 = self._try_multiple(self[second], *second, back, division=True)
        q_1, r_1 = site_1.qr("r", {*(f"P{orbit}" for _, orbit in legs_1), toward}, toward, back)
        q_2, r_2 = site_2.qr("r", {*(f"P{orbit}" for _, orbit in legs_2), back}, back, toward)
        u, s, v = (r_1.edge_rename({f"P{orbit}": f"P{rank}" for rank, orbit in legs_1})
                   .contract(r_2.edge_rename({f"P{orbit}": f"P{rank}" for rank, orbit in legs_2}), {(toward, back)})
                   .contract(gate, {(f"P{rank}", f"I{rank}") for rank in range(body)})
                   .svd({*(f"O{rank}" for rank, _ in legs_1), back}, toward, back, back, toward, new_dimension))
        s /= s.norm_2()
        self.environment[i, j, toward] = s
        u = self._try_multiple(u, *first, toward)
        self[first] = u.contract(q_1, {(back, toward)}).edge_rename({f"O{rank}": f"P{orbit}" for rank, orbit in legs_1})
        v = self._try_multiple(v, *second, back)
        self[second] = v.contract(q_2, {(toward, back)}).edge_rename({f"O{rank}": f"P{orbit}" for rank, orbit in legs_2})

    def _long_range(self, coordinates, index_and_orbit, gate, new_dimension):
        """two sites that are not neighbours (:574-677): the gate is applied on site 1 with the R factor of site 2's physical legs
        attached through a carrier edge "V" (+ physical passengers "VP{orbit}"), which is then moved bond by bond (rows first, then
        columns) to site 2, re-truncating every bond it crosses."""
        here, target = coordinates
        far_legs = [(rank, orbit) for rank, (index, orbit) in enumerate(index_and_orbit) if index == 1]
        near_legs = [(rank, orbit) for rank, (index, orbit) in enumerate(index_and_orbit) if index == 0]
        passengers = {f"VP{orbit}" for _, orbit in far_legs}
        q_far, r_far = self[target].qr("r", {f"P{orbit}" for _, orbit in far_legs}, "V", "V")
        r_far = r_far.edge_rename({f"P{orbit}": f"VP{orbit}" for _, orbit in far_legs})
        self[here] = (self[here]
                      .contract(gate, {(f"P{orbit}", f"I{rank}") for rank, orbit in near_legs})
                      .edge_rename({f"O{rank}": f"P{orbit}" for rank, orbit in near_legs})
                      .contract(r_far, {(f"I{rank}", f"VP{orbit}") for rank, orbit in far_legs})
                      .edge_rename({f"O{rank}": f"VP{orbit}" for rank, orbit in far_legs}))
        self[target] = q_far
        while here != target:
            if here[0] != target[0]:
                toward = "D" if here[0] < target[0] else "U"
                nxt = (here[0] + (1 if toward == "D" else -1), here[1])
            else:
                toward = "R" if here[1] < target[1] else "L"
                nxt = (here[0], here[1] + (1 if toward == "R" else -1))
            back = _OPPOSITE[toward]
            if nxt == target:
                q_here, r_here = self[here].qr("r", {"V", toward} | passengers, toward, back)
                q_next, r_next = self[nxt].qr("r", {back, "V"}, back, toward)
                big = self._try_multiple(r_here, *here, toward).contract(r_next, {(toward, back), ("V", "V")})
                u, s, v = big.svd({back}, toward, back, back, toward, new_dimension)
                s /= s.norm_2()
                self.environment[(*here, toward)] = s
                u = self._try_multiple(u, *here, toward)
                v = self._try_multiple(v, *nxt, back)
                self[here] = q_here.contract(u, {(toward, back)})
                self[nxt] = q_next.contract(v, {(back, toward)})
            else:
                u, s, v = self[here].svd({"V", toward} | passengers, back, toward, toward, back, new_dimension)
                u = self._try_multiple(u, *here, toward, division=True)
                s /= s.norm_2()
                self.environment[(*here, toward)] = s
                u = self._try_multiple(u, *nxt, back)
                v = self._try_multiple(v, *here, toward)
                self[here] = v
                self[nxt] = self[nxt].contract(u, {(back, toward)})
            here = nxt
        self[target] = self[target].edge_rename({f"VP{orbit}": f"P{orbit}" for _, orbit in far_legs})

    def _try_multiple(self, tensor, i, j, direction, *, division=False, square_root=False):
        """multiply (or divide by) the environment of bond `direction` of site (i, j), or its square root, into `tensor`; the
        tensor is returned unchanged when there is no environment there (:679-722).  For the square root the two ends of a bond
        get different factors so that the fermionic signs carried by the environment tensor are used exactly once."""
        env = self.environment[i, j, direction]
        if env is None:
            return tensor
        if division:
            env = env.reciprocal()
        if square_root:
            root = env.sqrt()
            if direction in ("D", "R"):
                env = env.same_shape().identity_({tuple(env.names)}) * root
            else:
                env = env * root.reciprocal()
        return tensor.contract(env, {(direction, _OPPOSITE[direction])})

    # -- outside the sampling path ---------------------------------------------------------------------
    def initialize_auxiliaries(self, cut_dimension):
        raise NotImplementedError("double-layer observation of a simple update lattice is outside the sampling-VMC path: convert with "
                                  "simple_update_lattice_to_sampling_lattice and use the sampling Observer")

    def observe(self, positions, observer):
        self.initialize_auxiliaries(None)

    def observe_energy(self):
        self.initialize_auxiliaries(None)


def simple_update_lattice_to_sampling_lattice(state):
    """every bond environment is shared out as a square root to its two ends (conversion.py:24-46)"""
    if not isinstance(state, SimpleUpdateLattice):
        raise ValueError("Conversion input type mismatch")
    result = SamplingLattice(state)
    for l1, l2 in state.sites():
        this = state[l1, l2]
        for direction in "LURD":
            this = state._try_multiple(this, l1, l2, direction, division=True, square_root=True)
        result[l1, l2] = this
    return result


def sampling_lattice_to_simple_update_lattice(state):
    """site tensors taken over as they are, no environments (conversion.py:49-67)"""
    if not isinstance(state, SamplingLattice):
        raise ValueError("Conversion input type mismatch")
    result = SimpleUpdateLattice(state)
    for l1, l2 in state.sites():
        result[l1, l2] = state[l1, l2]
    return result

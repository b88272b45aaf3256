"""Model description and PEPS container with the reference's attribute names.

Mirrors (behaviour, not code) tetragono/tetragono/abstract_state.py:241-557 (``AbstractState``:
physics edges, Hamiltonian terms, total symmetry), abstract_lattice.py:82-242 (``AbstractLattice``:
virtual bonds, site tensor names ``P{orbit}, T, U, D, L, R``) and sampling_lattice/lattice.py:699-998
(``SamplingLattice``: one tensor per site, created with ``randn_`` in row-major site order so that
``TAT.random.seed`` gives the same PEPS as the reference).
"""
from __future__ import annotations

_DIRECTIONS = {"L": (0, -1, "R"), "R": (0, 1, "L"), "U": (-1, 0, "D"), "D": (1, 0, "U")}


class _PhysicsEdges:
    def __init__(self, owner):
        self.owner = owner

    def __getitem__(self, key):
        if len(key) == 3:
            return self.owner._physics_edges[key[0]][key[1]][key[2]]
        return self.owner._physics_edges[key[0]][key[1]]

    def __setitem__(self, key, edge):
        o = self.owner
        edge = o._construct_edge(edge)
        if key is ...:
            o._physics_edges = [[{0: edge} for _ in range(o.L2)] for _ in range(o.L1)]
        elif len(key) == 3:
            l1, l2, orbit = key
            o._physics_edges[l1][l2][orbit] = edge
            o._physics_edges[l1][l2] = dict(sorted(o._physics_edges[l1][l2].items()))
        else:
            o._physics_edges[key[0]][key[1]] = {0: edge}

    def __iter__(self):
        for l1, l2 in self.owner.sites():
            for orbit, edge in self.owner._physics_edges[l1][l2].items():
                yield (l1, l2, orbit), edge


class _Hamiltonians:
    def __init__(self, owner):
        self.owner = owner

    def __getitem__(self, points):
        return self.owner._hamiltonians[tuple(p if len(p) == 3 else (p[0], p[1], 0) for p in points)]

    def __iter__(self):
        return iter(sorted(self.owner._hamiltonians.items()))

    def __len__(self):
        return len(self.owner._hamiltonians)

    def __contains__(self, key):
        return key in self.owner._hamiltonians

    def _regroup(self, transform):
        result = {}
        for key, value in self:
            value, key = transform(value, key)
            result[key] = result[key] + value if key in result else value
        self.owner._hamiltonians = result
        return self

    def trace_repeated(self):
        """terms whose point list names a point more than once: contract the output of an earlier occurrence with the input of
        the next one, so that every point appears once (abstract_state.py:200-212, utility.py:421-465)"""
        return self._regroup(trace_repeated)

    def sort_points(self):
        """points of every term in ascending order (edges renamed accordingly), terms on the same points summed
        (abstract_state.py:214-226, utility.py:468-480)"""
        return self._regroup(sort_points)

    def check_hermite(self, threshold):
        for _, value in self:
            body = len(value.names) // 2
            swap = {f"I{i}": f"O{i}" for i in range(body)} | {f"O{i}": f"I{i}" for i in range(body)}
            if float((value - value.conjugate().edge_rename(swap)).norm_max()) > threshold:
                raise ValueError("The Hamiltonian is not Hermitian")
        return self

    def __setitem__(self, arg, tensor):
        o = self.owner
        if isinstance(arg, str):
            for l1, l2 in o.sites():
                if arg == "single_site":
                    o._set_hamiltonian(((l1, l2, 0),), tensor)
                elif arg == "vertical_bond":
                    if l1 != o.L1 - 1:
                        o._set_hamiltonian(((l1, l2, 0), (l1 + 1, l2, 0)), tensor)
                elif arg == "horizontal_bond":
                    if l2 != o.L2 - 1:
                        o._set_hamiltonian(((l1, l2, 0), (l1, l2 + 1, 0)), tensor)
                else:
                    raise ValueError("Unknown kind of hamiltonian")
        else:
            o._set_hamiltonian(tuple(p if len(p) == 3 else (p[0], p[1], 0) for p in arg), tensor)


_REGROUP_POOL = {}


def trace_repeated(tensor, points):
    """-> (tensor with one (I, O) pair per distinct point, the distinct points in order of first appearance); utility.py:421-465.
    Results are kept per (tensor identity, pattern): the element tables of the samplers are cached by tensor identity too."""
    uniques, to_unique, first_index = [], [], []
    for index, point in enumerate(points):
        if point not in uniques:
            first_index.append(index)
            uniques.append(point)
        to_unique.append(uniques.index(point))
    key = ("trace", id(tensor), tuple(to_unique))
    if key not in _REGROUP_POOL:
        trace_set, rename = set(), {}
        for u in range(len(uniques)):
            group = [i for i, v in enumerate(to_unique) if v == u]
            trace_set.update((f"I{former}", f"O{latter}") for former, latter in zip(group[:-1], group[1:]))
            rename[f"I{group[-1]}"] = f"I{group[0]}"
        result = tensor.trace(trace_set).edge_rename(rename)
        result = result.edge_rename({f"{d}{old}": f"{d}{new}" for new, old in enumerate(first_index) for d in "IO"})
        _REGROUP_POOL[key] = (tensor, result)
    return _REGROUP_POOL[key][1], tuple(uniques)


def sort_points(tensor, points):
    """-> (tensor with its (I, O) pairs renumbered to the ascending order of the points, the sorted points); utility.py:468-480"""
    ordered = tuple(sorted(points))
    order = tuple(ordered.index(point) for point in points)
    key = ("sort", id(tensor), order)
    if key not in _REGROUP_POOL:
        _REGROUP_POOL[key] = (tensor, tensor.edge_rename({f"{d}{before}": f"{d}{after}" for before, after in enumerate(order) for d in "IO"}))
    return _REGROUP_POOL[key][1], ordered


class AbstractState:
    def __init__(self, Tensor, L1, L2):
        self.Tensor = Tensor
        self.L1, self.L2 = L1, L2
        self._physics_edges = [[{} for _ in range(L2)] for _ in range(L1)]
        self._hamiltonians = {}
        self._total_symmetry = Tensor.model.Symmetry()
        self.attribute = {}

    def _init_by_copy(self, other):
        self.Tensor = other.Tensor
        self.L1, self.L2 = other.L1, other.L2
        self._physics_edges = [[dict(other._physics_edges[l1][l2]) for l2 in range(self.L2)] for l1 in range(self.L1)]
        self._hamiltonians = dict(other._hamiltonians)
        self._total_symmetry = other._total_symmetry
        self.attribute = dict(other.attribute)

    @property
    def Edge(self):
        return self.Tensor.model.Edge

    @property
    def Symmetry(self):
        return self.Tensor.model.Symmetry

    def sites(self):
        for l1 in range(self.L1):
            for l2 in range(self.L2):
                yield l1, l2

    def _construct_symmetry(self, value):
        return self.Symmetry(value)

    def _construct_edge(self, value):
        return value if isinstance(value, self.Edge) else self.Edge(value)

    @property
    def total_symmetry(self):
        return self._total_symmetry

    @total_symmetry.setter
    def total_symmetry(self, value):
        self._total_symmetry = self._construct_symmetry(value)

    @property
    def _total_symmetry_edge(self):
        return self.Edge([(-self._total_symmetry, 1)], False)

    @property
    def physics_edges(self):
        return _PhysicsEdges(self)

    @property
    def hamiltonians(self):
        return _Hamiltonians(self)

    @property
    def site_number(self):
        return sum(1 for _ in self.physics_edges)

    def _set_hamiltonian(self, points, tensor):
        body = len(points)
        if not isinstance(tensor, self.Tensor):
            raise TypeError("Wrong hamiltonian type")
        if set(tensor.names) != {f"{io}{j}" for io in "IO" for j in range(body)}:
            raise ValueError("Wrong hamiltonian name")
        for i in range(body):
            edge_out, edge_in = tensor.edge_by_name(f"O{i}"), tensor.edge_by_name(f"I{i}")
            if edge_out != self.physics_edges[points[i]]:
                raise ValueError("Wrong hamiltonian edge")
            if edge_out.conjugate() != edge_in:
                raise ValueError("Wrong hamiltonian edge")
        if tensor.norm_max() != 0:
            self._hamiltonians[points] = tensor


class _VirtualBond:
    def __init__(self, owner):
        self.owner = owner

    def __getitem__(self, where):
        if len(where) == 3:
            return self.owner._virtual_bond[where[0]][where[1]][where[2]]
        return self.owner._virtual_bond[where[0]][where[1]]

    def __setitem__(self, where, value):
        o = self.owner
        if isinstance(where, str):
            for l1, l2 in o.sites():
                o._set_virtual_bond((l1, l2, where), value)
        else:
            o._set_virtual_bond(where, value)


class AbstractLattice(AbstractState):
    def __init__(self, abstract):
        self._init_by_copy(abstract)
        if not hasattr(self, "_virtual_bond") or self._virtual_bond is None:
            self._virtual_bond = [[self._default_bonds(l1, l2) for l2 in range(self.L2)] for l1 in range(self.L1)]

    def _init_by_copy(self, other):
        super()._init_by_copy(other)
        vb = getattr(other, "_virtual_bond", None)
        self._virtual_bond = None if vb is None else [[dict(vb[l1][l2]) for l2 in range(self.L2)] for l1 in range(self.L1)]

    def _default_bonds(self, l1, l2):
        # insertion order T, U, D, L, R defines the site tensor's name order (abstract_lattice.py:172-183)
        result = {}
        if l1 == l2 == 0:
            result["T"] = self._total_symmetry_edge
        if l1 != 0:
            result["U"] = None
        if l1 != self.L1 - 1:
            result["D"] = None
        if l2 != 0:
            result["L"] = None
        if l2 != self.L2 - 1:
            result["R"] = None
        return result

    @property
    def virtual_bond(self):
        return _VirtualBond(self)

    def _set_one_side(self, l1, l2, direction, edge):
        if 0 <= l1 < self.L1 and 0 <= l2 < self.L2 and direction in self._virtual_bond[l1][l2]:
            self._virtual_bond[l1][l2][direction] = edge

    def _set_virtual_bond(self, where, edge):
        l1, l2, direction = where
        if direction not in _DIRECTIONS:
            raise ValueError("Invalid direction when setting virtual bond")
        edge = self._construct_edge(edge)
        self._set_one_side(l1, l2, direction, edge)
        d1, d2, opposite = _DIRECTIONS[direction]
        self._set_one_side(l1 + d1, l2 + d2, opposite, edge.conjugate())

    def _construct_tensor(self, l1, l2):
        names, edges = [], []
        for orbit, edge in self._physics_edges[l1][l2].items():
            names.append(f"P{orbit}")
            edges.append(edge)
        if l1 == l2 == 0:
            self._virtual_bond[0][0]["T"] = self._total_symmetry_edge
        for direction, edge in self._virtual_bond[l1][l2].items():
            if edge is not None:
                names.append(direction)
                edges.append(edge)
        return self.Tensor(names, edges).randn_()


class SamplingLattice(AbstractLattice):
    def __init__(self, abstract):
        self._init_by_copy(abstract)
        if self._virtual_bond is None:
            self._virtual_bond = [[self._default_bonds(l1, l2) for l2 in range(self.L2)] for l1 in range(self.L1)]
        self._lattice = [[self._construct_tensor(l1, l2) for l2 in range(self.L2)] for l1 in range(self.L1)]

    def __getitem__(self, l1l2):
        return self._lattice[l1l2[0]][l1l2[1]]

    def __setitem__(self, l1l2, value):
        self._lattice[l1l2[0]][l1l2[1]] = value

    def lattice_dot(self, a=None, b=None):
        """sum over sites of <a, b> with the trivial metric; the lattice's own tensors where None is given (lattice.py:921-934)"""
        a = self._lattice if a is None else a
        b = self._lattice if b is None else b
        total = 0.0
        for l1, l2 in self.sites():
            x, y = a[l1][l2], b[l1][l2]
            total += float(x.conjugate(True).contract(y, {(name, name) for name in x.names}))
        return total

    def bcast_lattice(self, root=0):
        """all ranks take rank `root`'s site tensors: one broadcast of one flat buffer (lattice.py:950-954)"""
        from .. import dist
        dist.broadcast_tensors([self[l1, l2] for l1, l2 in self.sites()], root=root)

    def apply_gradient(self, gradient, step_size):
        """theta <- theta - step * g  (lattice.py:921-948, plain update)"""
        for l1, l2 in self.sites():
            self._lattice[l1][l2] = self._lattice[l1][l2] - gradient[l1][l2] * step_size

    # -- gauge fixing / bond expansion (lattice.py:821-919) ---------------------------------------
    def expand_dimension(self, new_dimension, epsilon):
        """Re-factorise every bond: QR of both neighbours, SVD of the product of the two triangular factors (optionally
        perturbed by `epsilon` x noise and truncated / enlarged to `new_dimension`: an int, or a float factor of the present
        dimension), the square roots of the singular values shared between the two sites.  With `new_dimension == 1.0` and
        `epsilon == 0` this only fixes the gauge (the state is unchanged).  Bond order as the reference: vertical bonds below odd
        rows, below even rows, then horizontal bonds right of odd columns, of even columns; every rank does all bonds, the
        arithmetic is deterministic (and the noise comes from the shared TAT.random engine), so no broadcast is needed.
        All symmetry types (the fermionic signs sit in `Tensor.identity_`, identity.hpp)."""
        for parity in (0, 1):
            for l1, l2 in self.sites():
                if l1 != 0 and l1 % 2 == parity:
                    self._refactor_bond((l1 - 1, l2), (l1, l2), "D", "U", new_dimension, epsilon)
        for parity in (0, 1):
            for l1, l2 in self.sites():
                if l2 != 0 and l2 % 2 == parity:
                    self._refactor_bond((l1, l2 - 1), (l1, l2), "R", "L", new_dimension, epsilon)

    def _refactor_bond(self, first, second, bond_1, bond_2, new_dimension, epsilon):
        """`bond_1` of site `first` is joined to `bond_2` of site `second`"""
        a, b = self[first], self[second]
        if isinstance(new_dimension, float):
            new_dimension = round(a.edge_by_name(bond_1).dimension * new_dimension)
        keep_1, keep_2 = {bond_1}, {bond_2}
        if epsilon != 0:       # with noise the physical legs take part, so that the enlarged bond can carry new directions
            keep_1 |= {n for n in a.names if n.startswith("P")}
            keep_2 |= {n for n in b.names if n.startswith("P")}
        a_q, a_r = a.qr("r", keep_1, bond_1, bond_2)
        b_q, b_r = b.qr("r", keep_2, bond_2, bond_1)
        a_r = a_r.edge_rename({n: f"A_{n}" for n in a_r.names})
        b_r = b_r.edge_rename({n: f"B_{n}" for n in b_r.names})
        core = a_r.contract(b_r, {(f"A_{bond_1}", f"B_{bond_2}")})
        noise = core.same_shape().randn_()      # drawn unconditionally, as the reference does (lattice.py:881, 909): the global stream stays in step
        if epsilon != 0:
            core = core + noise * (epsilon * float(core.norm_max()))
        u, sv, v = core.svd({n for n in core.names if n.startswith("A_")}, bond_1, bond_2, bond_2, bond_1, new_dimension)
        root = sv.sqrt()
        eye = sv.same_shape().identity_({(bond_2, bond_1)})
        eye *= root
        sv *= root.reciprocal()
        a_new = a_q.contract(u, {(bond_1, f"A_{bond_2}")}).contract(sv, {(bond_1, bond_2)})
        b_new = b_q.contract(v, {(bond_2, f"B_{bond_1}")}).contract(eye, {(bond_2, bond_1)})
        self[first] = a_new.edge_rename({n: n[2:] for n in a_new.names if n.startswith("A_")})
        self[second] = b_new.edge_rename({n: n[2:] for n in b_new.names if n.startswith("B_")})

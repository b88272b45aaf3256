"""Sampled configuration(s) with their boundary-MPS environments.

Mirrors ``Configuration`` of the reference (tetragono/tetragono/sampling_lattice/lattice.py:27-416):
``configuration[l1, l2, orbit] = edge_point``, ``replace``, ``hole``/``holes``, ``copy`` (warm caches),
``import_/export_configuration`` -- with one extension: a configuration object may hold ``nb``
Monte-Carlo chains at once (lock-step batch).  All chains of a batch share the symmetry sector of
every physical index (always true without symmetry); the index inside the sector is per chain.
An edge point is ``(Symmetry, index)`` where ``index`` is an int or an int array of length nb.
"""
from __future__ import annotations

import numpy as np

from .. import backend as _bk
from ..TAT.tensor import Tensor as _NativeTensor
from .auxiliaries import SingleLayerAuxiliaries, safe_rename


def is_native(Tensor):
    """True for this repository's device tensors; False for any other PyTAT-compatible class (the
    drivers then use only the public PyTAT API, one chain at a time -- used to time the reference)."""
    return isinstance(Tensor, type) and issubclass(Tensor, _NativeTensor)


def is_no_symmetry(Tensor):
    return Tensor.model.__name__.rsplit(".", 1)[-1] in ("No", "Normal")


# Lock-step batches of a symmetric model: "sector" = sector-compact tensors with per-chain symmetry sectors (TAT/ragged.py: the
# reference's sector plan, chain by chain); "dense" = the charge-dense embedding of round 1 (dense_embedding.py, bosonic only).
LOCKSTEP_ENGINE = "sector"


def sector_engine_supported(Tensor):
    kinds = Tensor.Symmetry.kinds
    return len(kinds) != 0 and all(k == "U1" for k in kinds)


class _RaggedFactory:
    """what SingleLayerAuxiliaries needs of a tensor class (`Tensor(1)`) when the environments are sector-compact tensors"""

    def __init__(self, Tensor):
        self.Symmetry, self.model = Tensor.Symmetry, Tensor.model

    def __call__(self, number):
        from ..TAT.ragged import RTensor, fermi_mask
        return RTensor.scalar_one(number, fermi_mask(self.Symmetry) if self.Symmetry.is_fermi_symmetry else 0)


class Configuration(SingleLayerAuxiliaries):
    def __init__(self, owner, cut_dimension, nb=1, engine=None):
        native = is_native(owner.Tensor)
        if engine is None:
            engine = "sector" if (native and nb > 1 and owner.Tensor.Symmetry.length != 0 and LOCKSTEP_ENGINE == "sector") else "plain"
        if engine == "sector" and not (native and sector_engine_supported(owner.Tensor)):
            raise NotImplementedError("lock-step chains of this symmetry type: only integer (U1-type) symmetries have per-chain sectors")
        self._ragged = engine == "sector"
        super().__init__(owner.L1, owner.L2, cut_dimension, False, _RaggedFactory(owner.Tensor) if self._ragged else owner.Tensor)
        self.owner = owner
        self.nb = nb
        self._native = native
        if not self._native and nb != 1:
            raise NotImplementedError("lock-step batches need the device-backed TAT tensors")
        # per site: {orbit: (Symmetry, int array [nb]) | None}
        self._configuration = [[{orbit: None for orbit in owner.physics_edges[l1, l2]} for l2 in range(owner.L2)] for l1 in range(owner.L1)]
        self._set_site_without_orbit()
        self._holes = None

    def _set_site_without_orbit(self):
        for l1, l2 in self.owner.sites():
            if len(self.owner.physics_edges[l1, l2]) == 0:
                SingleLayerAuxiliaries.__setitem__(self, (l1, l2), self.owner[l1, l2])

    def copy(self, cp=None):
        result = super().copy(cp=cp)
        result.owner = self.owner
        result.nb = self.nb
        result._native = self._native
        result._ragged = self._ragged
        result._configuration = [[dict(self._configuration[l1][l2]) for l2 in range(self.owner.L2)] for l1 in range(self.owner.L1)]
        result._holes = self._holes
        return result

    # -- edge points -----------------------------------------------------------------------------
    def _construct_edge_point(self, value, edge=None):
        if isinstance(value, tuple):
            symmetry, index = value
        else:
            symmetry, index = self.owner.Symmetry(), value
        index = np.asarray(index, dtype=np.int32).reshape(-1)
        if index.size == 1 and self.nb != 1:
            index = np.repeat(index, self.nb)
        if index.size != self.nb:
            raise ValueError("edge point index must be a scalar or one index per chain")
        if self._ragged:
            # per-chain sectors: a point is (None, total index per chain)
            if symmetry is not None:
                index = (edge.index_by_point((self.owner._construct_symmetry(symmetry), 0)) + index).astype(np.int32)
            return (None, index)
        return (self.owner._construct_symmetry(symmetry), index)

    @staticmethod
    def _same_point(a, b):
        return a is not None and b is not None and a[0] == b[0] and np.array_equal(a[1], b[1])

    def site_valid(self, l1, l2):
        return all(v is not None for v in self._configuration[l1][l2].values())

    def valid(self):
        return all(self.site_valid(l1, l2) for l1, l2 in self.owner.sites())

    def __getitem__(self, l1l2o):
        l1, l2, orbit = l1l2o
        return self._configuration[l1][l2][orbit]

    def __setitem__(self, l1l2o, value):
        l1, l2, orbit = l1l2o
        if value is None:
            self._configuration[l1][l2][orbit] = None
            SingleLayerAuxiliaries.__setitem__(self, (l1, l2), None)
            self._holes = None
            return
        point = self._construct_edge_point(value, self.owner.physics_edges[l1, l2, orbit])
        changed = not self._same_point(point, self._configuration[l1][l2][orbit])
        if changed:
            self._configuration[l1][l2][orbit] = point
        if self._lattice[l1][l2]() is None or changed:
            if self.site_valid(l1, l2):
                SingleLayerAuxiliaries.__setitem__(self, (l1, l2), self._shrink_configuration((l1, l2), self._configuration[l1][l2]))
                self._holes = None

    def __delitem__(self, l1l2o):
        self.__setitem__(l1l2o, None)

    def import_configuration(self, config):
        """config[l1][l2][orbit] = total edge index (>= 0), -1 absent, -2 None; an extra leading axis of
        length nb gives one configuration per chain (lattice.py:181-207)."""
        config = np.asarray(config)
        batched = config.ndim == 4
        for l1, l2 in self.owner.sites():
            for orbit in self.owner.physics_edges[l1, l2]:
                idx = config[:, l1, l2, orbit] if batched else config[l1, l2, orbit:orbit + 1]
                if (idx == -2).all():
                    self[l1, l2, orbit] = None
                elif (idx == -1).all():
                    continue
                elif (idx >= 0).all():
                    self[l1, l2, orbit] = self._point_by_index(self.owner.physics_edges[l1, l2, orbit], idx)
                else:
                    raise RuntimeError("Invalid edge index")

    def export_configuration(self):
        """int64 array [nb, L1, L2, orbits] (squeezed to [L1, L2, orbits] for a single chain)"""
        max_orbit = max(orbit for (l1, l2, orbit), _ in self.owner.physics_edges)
        result = np.zeros([self.nb, self.owner.L1, self.owner.L2, max_orbit + 1], dtype=np.int64) - 1
        for (l1, l2, orbit), edge in self.owner.physics_edges:
            point = self[l1, l2, orbit]
            result[:, l1, l2, orbit] = -2 if point is None else self._index_by_point(edge, point)
        return result[0] if self.nb == 1 else result

    @staticmethod
    def _point_by_index(edge, idx):
        """total index -> (symmetry, offset) when all chains land in one segment, else (None, total index per chain): the
        form only the sector-compact engine accepts"""
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        p0, _ = edge.coord_by_index(int(idx[0]))
        start = sum(d for _, d in edge.segments[:p0])
        off = idx - start
        if (off < 0).any() or (off >= edge.segments[p0][1]).any():
            return (None, idx.astype(np.int32))
        return (edge.segments[p0][0], off.astype(np.int32))

    @staticmethod
    def _index_by_point(edge, point):
        sym, off = point
        if sym is None:
            return np.asarray(off, dtype=np.int64)
        return edge.index_by_point((sym, 0)) + np.asarray(off, dtype=np.int64)

    # -- shrinking -------------------------------------------------------------------------------
    def _get_shrinker(self, l1l2, configuration):
        """one-hot (P, Q) tensors selecting the sampled physical slice (lattice.py:291-317)"""
        l1, l2 = l1l2
        for orbit in self.owner.physics_edges[l1, l2]:
            edge = self.owner.physics_edges[l1, l2, orbit]
            symmetry, index = configuration[orbit]
            if self._ragged:
                yield orbit, self._ragged_shrinker(edge, index)
                continue
            cedge = edge.conjugate()
            t = self.Tensor(["P", "Q"], [[(symmetry, 1)], cedge])
            if not self._native:
                t.zero_()
                t[{"Q": (-symmetry, int(index[0])), "P": (symmetry, 0)}] = 1
                yield orbit, t
                continue
            # the only block is (P = symmetry, Q = -symmetry), 1 x d: one-hot at `index`, one row per chain
            b = t._table.block_by_positions((0, cedge.position_by_symmetry(-symmetry)))
            onehot = np.zeros((len(index), t._table.size))
            onehot[np.arange(len(index)), int(t._table.offsets[b]) + index] = 1.0
            yield orbit, type(t).from_batch(t.names, t._edges, onehot)

    def _ragged_shrinker(self, edge, index):
        """sector-compact (P, Q): P a unit edge carrying the sampled charge of every chain, Q the conjugated physical edge,
        one-hot at the sampled index"""
        from ..TAT import ragged
        B = _bk.get()
        labels = np.concatenate([np.full(d, ragged.pack_symmetry(s), dtype=np.int32) for s, d in edge.segments])
        d = edge.dimension
        chosen = labels[index].astype(np.int32)
        onehot = np.zeros((len(index), d))
        onehot[np.arange(len(index)), index] = 1.0
        edges = [ragged.Edge(1, None, 1, False, chosen), ragged.Edge(d, B.from_numpy(labels.reshape(1, -1)), -1, not edge.arrow)]
        return ragged.RTensor.from_dense(["P", "Q"], edges, onehot, (-chosen).astype(np.int32),
                                         fermi=ragged.fermi_mask(self.owner.Symmetry) if self.owner.Symmetry.is_fermi_symmetry else 0)

    def _ragged_site(self, l1, l2):
        """sector-compact copy of the owner's site tensor (rebuilt when the owner's tensor object changes)"""
        from ..TAT import ragged
        cache = self.owner.__dict__.setdefault("_ragged_sites", {})
        tensor = self.owner[l1, l2]
        got = cache.get((l1, l2))
        if got is None or got[0] is not tensor or got[2] is not tensor._data:
            got = cache[(l1, l2)] = (tensor, ragged.RTensor.from_symmetric(tensor, unit_names=("T",)), tensor._data)
        return got[1]

    def _ragged_variants(self, l1, l2, orbit):
        """the d shrunk versions of a single-orbit site tensor (one per physical index), stacked: every chain of a batch then
        SELECTS its version with one row gather instead of contracting with its own one-hot tensor (lattice.py:319-339)"""
        from ..TAT import ragged
        B = _bk.get()
        site = self._ragged_site(l1, l2)
        cache = self.owner.__dict__.setdefault("_ragged_variants", {})
        got = cache.get((l1, l2))
        if got is not None and got[0] is site:
            return got[1]
        edge = self.owner.physics_edges[l1, l2, orbit]
        forms, names = [], None
        for p in range(edge.dimension):
            shr = self._ragged_shrinker(edge, np.array([p], dtype=np.int32))
            v = site.contract(shr.edge_rename({"P": f"P_{l1}_{l2}_{orbit}"}), {(f"P{orbit}", "Q")})
            forms.append(v._primary())
            names = v.names
            proto = v
        import torch
        cap = max(f.data.shape[1] for f in forms)
        cap += cap & 1
        # one table row per physical index: [stored data | pairing table viewed as float64], so that ONE row gather selects both
        match = torch.cat([f.match for f in forms], dim=0).contiguous().view(torch.float64)
        table = B.zeros(len(forms), cap + match.shape[1])
        for p, f in enumerate(forms):
            table[p, :f.data.shape[1]] = f.data[0]
        table[:, cap:] = match
        labels = np.concatenate([np.full(d, ragged.pack_symmetry(sy), dtype=np.int32) for sy, d in edge.segments])
        out = (proto, table, labels, cap)
        cache[(l1, l2)] = (site, out)
        return out

    def _shrink_by_selection(self, l1, l2, orbit, index):
        from ..TAT import ragged
        import torch
        B = _bk.get()
        proto, table, labels, cap = self._ragged_variants(l1, l2, orbit)
        idx = B.from_numpy(np.ascontiguousarray(index, dtype=np.int32))
        sel = B.gather_rows(table, table.shape[1], idx)          # [nb, cap + 66]: data and pairing table of every chain's version
        sel_data = sel[:, :cap]
        sel_match = sel[:, cap:].view(torch.int32)
        chosen = labels[index].astype(np.int32)
        f = proto._primary()
        pcore = proto.core
        edges = [e if not (e.unit and n == f"P_{l1}_{l2}_{orbit}") else ragged.Edge(1, None, e.sign, e.arrow, chosen)
                 for n, e in zip(proto.names, pcore.edges)]
        # target of the shrunk tensor: the site's own (the T edge) minus the sampled physical charge, per chain
        unit_sum = np.zeros(len(index), dtype=np.int64)
        for e in edges:
            if e.unit:
                unit_sum = unit_sum + e.sign * np.asarray(e.harr, dtype=np.int64).reshape(-1)
        core = ragged.Core(edges, len(index), B.from_numpy((-unit_sum).astype(np.int32)), 1, pcore.fermi)
        core.tables = dict(pcore.tables)
        core.set_primary(ragged.Form(f.rows, f.cols, f.rt, f.rs, f.ct, f.cs, sel_match, sel_data, f.M, f.N))
        return ragged.RTensor(proto.names, core, 1)

    def _shrink_configuration(self, l1l2, configuration):
        l1, l2 = l1l2
        if self._ragged:
            orbits = list(self.owner.physics_edges[l1, l2])
            if len(orbits) == 1:
                return self._shrink_by_selection(l1, l2, orbits[0], configuration[orbits[0]][1])
            tensor = self._ragged_site(l1, l2)
            for orbit, shrinker in self._get_shrinker(l1l2, configuration):
                tensor = tensor.contract(shrinker.edge_rename({"P": f"P_{l1}_{l2}_{orbit}"}), {(f"P{orbit}", "Q")})
            return tensor
        tensor = self.owner[l1l2]
        orbits = list(self.owner.physics_edges[l1, l2])
        # fast path: no symmetry, single orbit stored first -> a row gather of the site tensor
        if self._native and self.Tensor.Symmetry.length == 0 and orbits == [0] and tensor.names[0] == "P0" and tensor.nb == 1:
            _, index = configuration[0]
            B = _bk.get()
            d = tensor._edges[0].dimension
            row = tensor.storage.size // d
            data = B.gather_rows(tensor.data, row, B.from_numpy(np.ascontiguousarray(index, dtype=np.int32)))
            names = tensor.names[1:] + [f"P_{l1}_{l2}_0"]
            edges = tensor._edges[1:] + (self.owner.Edge(1),)
            return type(tensor).from_batch(names, edges, data)
        for orbit, shrinker in self._get_shrinker(l1l2, configuration):
            tensor = tensor.contract(shrinker.edge_rename({"P": f"P_{l1}_{l2}_{orbit}"}), {(f"P{orbit}", "Q")})
        return tensor

    def refresh_site(self, l1l2o):
        configuration = self[l1l2o]
        del self[l1l2o]
        self[l1l2o] = configuration

    def refresh_all(self):
        for l1, l2 in self.owner.sites():
            for orbit in self.owner.physics_edges[l1, l2]:
                self.refresh_site((l1, l2, orbit))

    # -- amplitudes ------------------------------------------------------------------------------
    def replace(self, replacement, *, hint=None):
        """<s'|psi> with several physical indices replaced (lattice.py:231-267)."""
        grouped = {}
        for (l1, l2, orbit), point in replacement.items():
            grouped.setdefault((l1, l2), {})[orbit] = self._construct_edge_point(point, self.owner.physics_edges[l1, l2, orbit])
        base = {}
        for l1l2, site in grouped.items():
            l1, l2 = l1l2
            changed = False
            for orbit, current in self._configuration[l1][l2].items():
                if orbit not in site:
                    site[orbit] = current
                elif not self._same_point(site[orbit], current):
                    changed = True
            if changed:
                base[l1l2] = self._shrink_configuration(l1l2, site)
        return SingleLayerAuxiliaries.replace(self, base, hint=hint)

    def holes(self):
        """<psi|s|d_x psi> / <psi|s|psi> for every site tensor x (lattice.py:362-416)."""
        if self._holes is None:
            owner = self.owner
            ws = self.hole(())
            inv_ws_conj = ws / (ws.norm_2()**2)
            inv_ws = inv_ws_conj.conjugate()
            all_name = {("T", "T")} | {(f"P_{l1}_{l2}_{orbit}",) * 2 for l1, l2 in owner.sites() for orbit in owner.physics_edges[l1, l2]}
            holes = [[None] * owner.L2 for _ in range(owner.L1)]
            for l1, l2 in owner.sites():
                hole = self.hole(((l1, l2),))
                names = set(all_name)
                for orbit in owner.physics_edges[l1, l2]:
                    names.discard((f"P_{l1}_{l2}_{orbit}",) * 2)
                if "T" not in hole.names:
                    names.discard(("T", "T"))
                hole = hole.contract(inv_ws, names)
                hole = safe_rename(hole, {"L0": "R", "R0": "L", "U0": "D", "D0": "U",
                                          **{f"P_{l1}_{l2}_{orbit}": f"P{orbit}" for orbit in owner.physics_edges[l1, l2]}})
                for orbit, shrinker in self._get_shrinker((l1, l2), self._configuration[l1][l2]):
                    hole = hole.contract(shrinker, {(f"P{orbit}", "P")}).edge_rename({"Q": f"P{orbit}"})
                holes[l1][l2] = hole
            self._holes = holes
        return self._holes


class ConfigurationPool:
    """Amplitudes of configurations that differ from a sampled one OUTSIDE the 2x2 window of `Configuration.replace`
    (long-range observables; reference: sampling_lattice/lattice.py:473-700 `ConfigurationPool.wss`).  The replacement is split as
    the reference splits it (`_split_replacement`, :648-684: the last cluster of changed sites that fits a 2x2 window stays a
    replacement, the rest is applied to a copy of the configuration), the half-replaced configurations are kept so that the
    other matrix elements of the same observable reuse their environments.  Where the reference may also answer from the
    "nearest" configuration it has seen (a different but equivalent contraction route -- equal up to the boundary truncation),
    this pool always takes the split route.  Only copies made here are stored, never the sampler's live configuration."""

    def __init__(self, owner):
        self.owner = owner
        self.tree = {}

    def _key(self, configuration, replacement=None):
        config = np.array(configuration.export_configuration())
        if replacement:
            for (l1, l2, orbit), point in replacement.items():
                point = configuration._construct_edge_point(point, self.owner.physics_edges[l1, l2, orbit])
                config[..., l1, l2, orbit] = Configuration._index_by_point(self.owner.physics_edges[l1, l2, orbit], point)
        return config.tobytes()

    def _split_replacement(self, replacement):
        first, second = {}, {}
        up, down, left, right = self.owner.L1, -1, self.owner.L2, -1
        for l1, l2 in self.owner.sites():
            for orbit in self.owner.physics_edges[l1, l2]:
                site = (l1, l2, orbit)
                if site not in replacement:
                    continue
                if down - 2 < l1 < up + 2 and right - 2 < l2 < left + 2:
                    second[site] = replacement[site]
                    up, down, left, right = min(up, l1), max(down, l1), min(left, l2), max(right, l2)
                else:
                    first[site] = replacement[site]
        return first, second

    def wss(self, configuration, replacement):
        wss = configuration.replace(replacement)
        if wss is not None:
            return wss
        whole = self._key(configuration, replacement)
        if whole in self.tree:
            return self.tree[whole].hole(())
        first, second = self._split_replacement(replacement)
        key = self._key(configuration, first)
        if key not in self.tree:
            half = configuration.copy()
            for site, point in first.items():
                half[site] = point
            self.tree[key] = half
        wss = self.tree[key].replace(second)
        if wss is None:
            raise NotImplementedError("not implemented replace style")
        return wss

"""Markov-chain samplers over PEPS configurations.

Mirrors ``SweepSampling`` / ``ErgodicSampling`` of the reference
(tetragono/tetragono/sampling_lattice/sampling.py:74-249): same sweep order over the Hamiltonian
terms, same proposal rule (uniform choice among the connected s' of the term, Metropolis ratio
|<s'|psi>/<s|psi>|^(2 alpha) * n_hop(s)/n_hop(s')), same random-number consumption per chain.
Extension: ``nb`` chains advance in lock step; chain c draws from its own ``mt19937_64`` exactly as
MPI rank c of the reference would (seed recipe utility.py:146-150).
"""
from __future__ import annotations


import numpy as np

from .. import backend as _bk
from ..TAT import random as _random
from .configuration import Configuration
from .tensor_element import element_table


class ChainRng:
    """One libstdc++ ``mt19937_64`` per chain (host part of the C-ABI)."""

    def __init__(self, nb):
        self.nb = nb
        self.lib = _bk.host_lib()
        self.handle = self.lib.tnsp_rng_create_host(nb)
        self._one = np.ones(nb, dtype=np.uint8)

    def __del__(self):
        try:
            self.lib.tnsp_rng_destroy_host(self.handle)
        except Exception:
            pass

    def seed_like_reference(self):
        """chain c gets the seed MPI rank c would get from ``seed_differ`` (utility.py:146-150):
        (global uniform_int(0, 2^31-1) + c) mod 2^31, then one uniform_real is discarded."""
        base = _random.uniform_int(0, 2**31 - 1)()
        for c in range(self.nb):
            self.lib.tnsp_rng_seed_host(self.handle, c, (base + c) % 2**31)
        self.uniform_real(None)
        return base

    def seed(self, seeds):
        for c, s in enumerate(seeds):
            self.lib.tnsp_rng_seed_host(self.handle, c, int(s))

    def uniform_int(self, hi, active):
        """per chain uniform_int_distribution<int>(0, hi[c]) where active[c]"""
        lo = np.zeros(self.nb, dtype=np.int32)
        hi = np.ascontiguousarray(hi, dtype=np.int32)
        out = np.zeros(self.nb, dtype=np.int32)
        act = self._one if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        self.lib.tnsp_rng_uniform_int_host(self.handle, lo.ctypes.data, hi.ctypes.data, act.ctypes.data, out.ctypes.data)
        return out

    def uniform_real(self, active):
        out = np.zeros(self.nb, dtype=np.float64)
        act = self._one if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        self.lib.tnsp_rng_uniform_real_host(self.handle, 0.0, 1.0, act.ctypes.data, out.ctypes.data)
        return out


class _GlobalRng:
    """single chain drawing from the global TAT.random engine (what the reference does per process)"""
    nb = 1

    def uniform_int(self, hi, active):
        if active is not None and not active[0]:
            return np.zeros(1, dtype=np.int32)
        return np.array([_random.uniform_int(0, int(hi[0]))()], dtype=np.int32)

    def uniform_real(self, active):
        if active is not None and not active[0]:
            return np.zeros(1)
        return np.array([_random.uniform_real(0, 1)()])


_CAPACITY_BUMPS = {"n": 0, "dropped": 0}


def _check_capacity():
    """sector-compact engine: a chain whose sectors outgrew a learnt buffer capacity was stored empty (TAT/ragged.py), i.e. it carried
    amplitude zero through this sweep (its proposals were rejected / its sample has weight zero).  Loud, and self-correcting: the
    capacities grow by half for everything allocated from now on; after three such corrections the run stops."""
    from ..TAT import ragged
    B = _bk.get()
    dropped = B.rt_overflow()
    if dropped:
        _CAPACITY_BUMPS["n"] += 1
        _CAPACITY_BUMPS["dropped"] += int(dropped)
        if _CAPACITY_BUMPS["n"] > 3:
            raise RuntimeError(f"sector-compact engine: {dropped} (chain, tensor) pairs exceeded their learnt capacity again after three "
                               "corrections; raise tnsp_b200.TAT.ragged.CAP_FACTOR (now %g) or set CAPS_ENABLED = False" % ragged.CAP_FACTOR)
        import warnings
        ragged.CAP_FACTOR *= 1.5
        warnings.warn(f"sector-compact engine: {dropped} (chain, tensor) pairs exceeded their learnt capacity in this sweep (stored empty: "
                      f"amplitude zero); CAP_FACTOR raised to {ragged.CAP_FACTOR:g}", RuntimeWarning)


def _amplitude_values(ws):
    """host float array [nb] of a one-element (batched) tensor"""
    return np.atleast_1d(np.asarray(ws.storage, dtype=np.float64).reshape(-1))


def calibrate_sector_engine(owner, cut_dimension, configuration, hopping_hamiltonians=None, chains=148, sweeps=1, observer_options=None):
    """Learn the buffer capacities of the sector-compact engine (TAT/ragged.py) on a small throw-away batch BEFORE a large lock-step
    batch allocates anything: `chains` chains with their own random streams (the caller's engines are not touched) sweep and are
    observed once from the start `configuration` ([L1, L2, orbits] total physical indices).  No effect on any result."""
    from .observer import Observer
    rng = ChainRng(chains)
    rng.seed([(911 + 7 * c) % 2**31 for c in range(chains)])
    s = SweepSampling(owner, cut_dimension, None, hopping_hamiltonians, nb=chains, rng=rng)
    conf = np.asarray(configuration)
    s.configuration.import_configuration(np.broadcast_to(conf, (chains,) + conf.shape) if conf.ndim == 3 else conf[:chains])
    from ..TAT import ragged
    ragged._LEARN["all"], ragged._LEARN["cycles"] = True, 0
    obs = Observer(owner, **(observer_options or dict(enable_energy=True, enable_gradient=True)))
    with obs:
        for _ in range(sweeps):
            p, c = s()
            obs(p, c)
    ragged.freeze_capacities()


class Sampling:
    def __init__(self, owner, cut_dimension, restrict_subspace):
        self.owner = owner
        self._cut_dimension = cut_dimension
        self._restrict_subspace = restrict_subspace

    def refresh_all(self):
        raise NotImplementedError("Not implement in abstract sampling")

    def __call__(self):
        raise NotImplementedError("Not implement in abstract sampling")


class SweepSampling(Sampling):
    def __init__(self, owner, cut_dimension, restrict_subspace=None, hopping_hamiltonians=None, *, nb=1, rng=None, engine=None):
        super().__init__(owner, cut_dimension, restrict_subspace)
        self.nb = nb
        self.configuration = Configuration(owner, cut_dimension, nb, engine=engine)
        self._hopping_hamiltonians = hopping_hamiltonians if hopping_hamiltonians is not None else owner._hamiltonians
        self._sweep_order = self._get_proper_position_order()
        self.rng = rng if rng is not None else (_GlobalRng() if nb == 1 else ChainRng(nb))
        if restrict_subspace is not None and nb != 1:
            raise NotImplementedError("restrict_subspace callbacks are evaluated per chain; use nb=1")

    def _get_proper_position_order(self):
        """row-major: single-site and horizontal terms, then column-major vertical terms (sampling.py:156-190).

        Ties inside one (site, neighbour) group are broken by the iteration order of a Python ``set``
        of position tuples in the reference; the same sequence of set operations is performed here so
        that the order -- and with it the random-number consumption of a chain -- is identical."""
        L1, L2 = self.owner.L1, self.owner.L2
        pending = set(self._hopping_hamiltonians.keys())
        result = []

        def take(allowed):
            nonlocal pending
            rest = set()
            for ps in pending:
                if all(p[:2] in allowed for p in ps):
                    result.append(ps)
                else:
                    rest.add(ps)
            pending = rest

        for l1 in range(L1):
            for l2 in range(L2):
                # the reference's first pass tests `p not in ((l1, l2))`, i.e. membership in the int pair itself,
                # so it never matches (sampling.py:163-168) but still rebuilds the pending set
                take(())
                take(((l1, l2), (l1, l2 + 1)))
        for l2 in range(L2):
            for l1 in range(L1):
                take(((l1, l2), (l1 + 1, l2)))
        if pending:
            raise NotImplementedError("Not implemented hamiltonian")
        return result

    def _current_flat(self, table, positions):
        conf = self.configuration
        idx = [Configuration._index_by_point(table.edges[i], conf[positions[i]]) for i in range(table.body)]
        return table.flatten(idx)

    def _single_term(self, positions, hamiltonian, ws, ws_val):
        owner = self.owner
        table = element_table(hamiltonian, [owner.physics_edges[p] for p in positions])
        cur = self._current_flat(table, positions)
        n_hop = table.count[cur]
        active = n_hop > 0
        if not active.any():
            return ws, ws_val
        pick = self.rng.uniform_int(np.maximum(n_hop - 1, 0), active)
        target = np.where(active, table.targets[cur, pick], cur)
        n_hop_s = table.count[target]
        new_idx = table.unflatten(target)
        replacement = {positions[i]: Configuration._point_by_index(table.edges[i], new_idx[i]) for i in range(table.body)}
        if self._restrict_subspace is not None and not self._restrict_subspace(self.configuration, replacement):
            return ws, ws_val
        wss = self.configuration.replace(replacement)
        wss_val = _amplitude_values(wss)
        alpha = owner.attribute.get("alpha", 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            p = np.abs(wss_val / ws_val)**(2 * alpha) * n_hop / np.maximum(n_hop_s, 1)
        u = self.rng.uniform_real(active)
        accept = active & (u < p)
        if accept.any():
            if accept.all():
                ws, ws_val = wss, wss_val
                for i in range(table.body):
                    self.configuration[positions[i]] = replacement[positions[i]]
            else:
                if not getattr(self.configuration, "_ragged", False):   # (the amplitude tensor itself is only carried along)
                    B = _bk.get()
                    mask = B.from_numpy(accept.astype(np.uint8))
                    ws = type(ws).from_batch(ws.names, ws._edges, B.select(mask, wss.data, ws.data))
                ws_val = np.where(accept, wss_val, ws_val)
                cur_idx = table.unflatten(cur)
                for i in range(table.body):
                    merged = np.where(accept, new_idx[i], cur_idx[i])
                    self.configuration[positions[i]] = Configuration._point_by_index(table.edges[i], merged)
        return ws, ws_val

    def __call__(self):
        if not self.configuration.valid():
            raise RuntimeError("Configuration not initialized")
        ws = self.configuration.hole(())
        ws_val = _amplitude_values(ws)
        for positions in self._sweep_order:
            ws, ws_val = self._single_term(positions, self._hopping_hamiltonians[positions], ws, ws_val)
        alpha = self.owner.attribute.get("alpha", 1)
        possibility = np.abs(ws_val)**(2 * alpha)
        if getattr(self.configuration, "_ragged", False):
            _check_capacity()
        return (float(possibility[0]) if self.nb == 1 else possibility), self.configuration.copy()

    def refresh_all(self):
        self.configuration.refresh_all()


class ErgodicSampling(Sampling):
    """Enumerate every configuration (sampling.py:196-249); rank r of `size` takes r, r+size, ...

    `nb` > 1 (new): one call yields nb consecutive configurations of this rank's sequence as a lock-step batch, so that an exact
    enumeration keeps the GPU busy; the concatenation of the batches is the nb = 1 sequence.  When the sequence is exhausted
    inside a batch the surplus chains get possibility = inf, i.e. weight zero in the Observer (the reference uses the same
    device for configurations outside a restricted subspace, sampling.py:244-247)."""

    def __init__(self, owner, cut_dimension, restrict_subspace=None, *, rank=0, size=1, nb=1):
        super().__init__(owner, cut_dimension, restrict_subspace)
        self.nb = nb
        self.configuration = Configuration(owner, cut_dimension, nb)
        self._rank, self._size = rank, size
        self.total_step = 1
        self._digits = []        # (l1, l2, orbit, edge) in the order the reference increments them: first site fastest
        for l1, l2 in owner.sites():
            for orbit, edge in owner.physics_edges[l1, l2].items():
                self.total_step *= edge.dimension
                self._digits.append((l1, l2, orbit, edge))
        if nb != 1 and restrict_subspace is not None:
            raise NotImplementedError("restrict_subspace callbacks are evaluated per chain; use nb=1")
        self._counter = rank     # number of the configuration the (first) chain currently holds
        self._served = 0         # configurations of this rank's sequence handed out so far
        if nb == 1:
            self._zero_configuration()
            for _ in range(rank):
                self._next_configuration()

    def _zero_configuration(self):
        for l1, l2 in self.owner.sites():
            for orbit, edge in self.owner.physics_edges[l1, l2].items():
                self.configuration[l1, l2, orbit] = edge.point_by_index(0)

    def _next_configuration(self):
        for l1, l2 in self.owner.sites():
            for orbit, edge in self.owner.physics_edges[l1, l2].items():
                sym, off = self.configuration[l1, l2, orbit]
                index = edge.index_by_point((sym, int(off[0]))) + 1
                if index == edge.dimension:
                    self.configuration[l1, l2, orbit] = edge.point_by_index(0)
                else:
                    self.configuration[l1, l2, orbit] = edge.point_by_index(index)
                    return

    def refresh_all(self):
        self.configuration.refresh_all()

    @property
    def calls(self):
        """calls over ALL ranks (the driver gives rank r the calls with step % size == r, like the reference's `total_step` loop)
        that cover every configuration once: every rank must get ceil(longest rank sequence / nb) calls -- rank 0's sequence,
        ceil(total / size) configurations, is the longest; surplus chains of the last batch carry possibility = inf"""
        longest = -(-self.total_step // self._size)
        return self._size * (-(-longest // self.nb))

    def __call__(self):
        if self.nb == 1:
            for _ in range(self._size):
                self._next_configuration()
            possibility = 1.0
            if self._restrict_subspace is not None and not self._restrict_subspace(self.configuration):
                possibility = np.inf
            return possibility, self.configuration.copy()
        # chain c holds configuration number rank + (served + c + 1) * size (mod total), decoded digit by digit
        numbers = self._rank + (self._served + np.arange(self.nb, dtype=np.int64) + 1) * self._size
        mine = -(-(self.total_step - self._rank) // self._size) if self._size > 1 else self.total_step   # length of this rank's sequence
        possibility = np.where(self._served + np.arange(self.nb) < mine, 1.0, np.inf)
        rest = numbers % self.total_step
        for l1, l2, orbit, edge in self._digits:
            self.configuration[l1, l2, orbit] = Configuration._point_by_index(edge, rest % edge.dimension)
            rest = rest // edge.dimension
        self._served += self.nb
        return possibility, self.configuration.copy()

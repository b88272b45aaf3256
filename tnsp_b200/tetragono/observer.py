"""Local-energy / log-derivative observer for sampled configurations.

Mirrors ``Observer`` of the reference (tetragono/tetragono/sampling_lattice/observer.py:28-920):
reweighted moments of every observable term, E_s = sum_s' <s'|H|s> <s'|psi>/<s|psi>, the
log-derivatives Delta = <holes>, E*Delta, the gradient 2 conj(<E Delta> - <E><Delta>) and the SR
natural gradient by matrix-free conjugate gradient -- with chains processed as a lock-step batch
and the per-rank MPI reductions replaced by one allreduce over GPUs (``tnsp_b200.dist``).

Device-side pieces: amplitudes / holes through the TAT kernels, the fused accumulation kernel
``tnsp_grad_accumulate_f64`` (Delta += w*hole, EDelta += w*E*hole over all chains of the batch) and
the grouped-GEMM kernel for the CG mat-vecs.
"""
from __future__ import annotations

import numpy as np

from .. import backend as _bk
from .. import dist as _dist
from .configuration import Configuration, is_native, is_no_symmetry
from .tensor_element import element_table


def _values(t):
    return np.atleast_1d(np.asarray(t.storage, dtype=np.float64).reshape(-1))


_BLOCK_PLANS = {}


def _batched_element(observer, table, cur_idx, new_idx, values):
    """one-element Hamiltonian tensors <s'|H|s> of a lock-step batch as ONE sector-compact tensor: names and arrows of the term
    tensor, every edge a unit edge carrying the chain's charge on it (tensor_element.py:41-53 per chain), data = matrix element"""
    from ..TAT import ragged
    labels = {}
    for i, e in enumerate(table.edges):
        lab = np.concatenate([np.full(d, ragged.pack_symmetry(sy), dtype=np.int32) for sy, d in e.segments])
        labels[f"I{i}"] = (-lab[np.asarray(cur_idx[i], dtype=np.int64)]).astype(np.int32)
        labels[f"O{i}"] = lab[np.asarray(new_idx[i], dtype=np.int64)].astype(np.int32)
    names = list(observer.names)
    edges = [ragged.Edge(1, None, 1, observer.edge_by_name(n).arrow, labels[n]) for n in names]
    S = observer.Symmetry
    return ragged.RTensor.from_dense(names, edges, np.asarray(values, dtype=np.float64).reshape(-1, 1), None,
                                     fermi=ragged.fermi_mask(S) if S.is_fermi_symmetry else 0)


def _blocks_of(hole, target):
    """storage [nb, size] of a sector-compact tensor in the block layout of the symmetric tensor `target` (same names, same edges;
    core.hpp:162-190): expand to the dense index space, then one pack over the blocks"""
    from ..TAT.plan import PackPlan
    B = _bk.get()
    dense = hole.to_dense()
    key = (type(target), target._edges)
    plan = _BLOCK_PLANS.get(key)
    if plan is None:
        table = target._table
        dims = [e.dimension for e in target._edges]
        strides, acc = [], 1
        for d in reversed(dims):
            strides.append(acc)
            acc *= d
        strides.reverse()
        starts = target._segment_starts()
        rows = []
        for b, pos in enumerate(table.positions):
            bd = [int(d) for d in table.dims[b]]
            if 0 in bd:
                continue
            src_off = sum(int(starts[i][int(p)]) * strides[i] for i, p in enumerate(pos))
            dstr, a = [], 1
            for d in reversed(bd):
                dstr.append(a)
                a *= d
            dstr.reverse()
            keep = [i for i, d in enumerate(bd) if d != 1]
            rows.append((src_off, int(table.offsets[b]), 0, [bd[i] for i in keep], [strides[i] for i in keep], [dstr[i] for i in keep]))
        plan = _BLOCK_PLANS[key] = PackPlan(tuple(target.names), target._edges, table, rows, acc)
    out = B.empty(dense.shape[0], target._table.size) if plan.covers_all else B.zeros(dense.shape[0], target._table.size)
    B.pack(plan, dense, out)
    return out


class Observer:
    def __init__(self, owner, *, observer_set=None, enable_energy=False, enable_gradient=False, enable_natural_gradient=False,
                 cache_natural_delta=None, cache_configuration=False, restrict_subspace=None, classical_energy=None):
        self.owner = owner
        self._observer = dict(observer_set) if observer_set is not None else {}
        self._enable_gradient = False
        self._enable_natural = False
        self._restrict_subspace = restrict_subspace
        self._classical_energy = classical_energy
        self._start = False
        if cache_natural_delta is not None:
            raise NotImplementedError("the natural-gradient delta cache is outside the hot path (SURVEY.md 8f)")
        if cache_configuration not in (False, True, "drop"):
            raise ValueError("cache_configuration must be False, True or 'drop'")
        # observer.py:198,305-331: with a configuration cache, observables beyond the 2x2 replace window are measured through
        # ConfigurationPool.wss; "drop" starts an empty pool at every sample
        self._cache_configuration = cache_configuration
        self._pool = None
        if enable_energy:
            self.add_energy()
        if enable_gradient:
            self.enable_gradient()
        if enable_natural_gradient:
            self.enable_natural_gradient()

    def set_classical_energy(self, classical_energy=None):
        """a function of the Configuration whose value is added to the energy (observer.py:213-222)"""
        self._classical_energy = classical_energy

    def restrict_subspace(self, restrict_subspace):
        if self._start:
            raise RuntimeError("Cannot set restrict subspace after sampling start")
        self._restrict_subspace = restrict_subspace

    def cache_configuration(self, cache_configuration):
        if self._start:
            raise RuntimeError("Cannot enable caching after sampling start")
        if cache_configuration not in (False, True, "drop"):
            raise ValueError("cache_configuration must be False, True or 'drop'")
        self._cache_configuration = cache_configuration

    def add_observer(self, name, observers):
        if self._start:
            raise RuntimeError("Cannot enable hole after sampling start")
        self._observer[name] = observers

    def add_energy(self):
        self.add_observer("energy", self.owner._hamiltonians)

    def enable_gradient(self):
        if self._start:
            raise RuntimeError("Cannot enable gradient after sampling start")
        if "energy" not in self._observer:
            self.add_energy()
        self._enable_gradient = True

    def enable_natural_gradient(self):
        if self._start:
            raise RuntimeError("Cannot enable natural gradient after sampling start")
        if not self._enable_gradient:
            self.enable_gradient()
        self._enable_natural = True

    # -- accumulation window -----------------------------------------------------------------------
    def __enter__(self):
        self._start = True
        z = lambda: {name: {positions: 0.0 for positions in obs} for name, obs in self._observer.items()}  # noqa: E731
        self._result_reweight, self._result_reweight_square, self._result_square_reweight_square = z(), z(), z()
        self._count = 0
        self._pool = None
        self._total_weight = 0.0
        self._total_weight_square = 0.0
        self._total_log_ws = 0.0
        self._whole_result_reweight = {name: 0.0 for name in self._observer}
        self._whole_result_reweight_square = {name: 0.0 for name in self._observer}
        self._whole_result_square_reweight_square = {name: 0.0 for name in self._observer}
        self._total_imaginary_energy_reweight = 0.0
        if self._enable_gradient:
            owner = self.owner
            self._Delta = [[owner[l1, l2].same_shape().conjugate().zero_() for l2 in range(owner.L2)] for l1 in range(owner.L1)]
            self._EDelta = [[owner[l1, l2].same_shape().conjugate().zero_() for l2 in range(owner.L2)] for l1 in range(owner.L1)]
            if self._enable_natural:
                self._Deltas = []   # list of (weights [nb], energies [nb], device matrix [nb, Np])
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        """reduce over GPUs (observer.py:83-126): one packed scalar vector + one flat Delta||EDelta buffer"""
        if exc_type is not None:
            return False
        buffer = []
        for name, observers in self._observer.items():
            for positions in observers:
                buffer += [self._result_reweight[name][positions], self._result_reweight_square[name][positions],
                           self._result_square_reweight_square[name][positions]]
        buffer += [self._count, self._total_weight, self._total_weight_square, self._total_log_ws]
        for name in self._observer:
            buffer += [self._whole_result_reweight[name], self._whole_result_reweight_square[name],
                       self._whole_result_square_reweight_square[name]]
        buffer.append(self._total_imaginary_energy_reweight)
        buffer = _dist.allreduce_host(np.array(buffer, dtype=np.float64)).tolist()
        self._total_imaginary_energy_reweight = buffer.pop()
        for name in reversed(list(self._observer)):
            self._whole_result_square_reweight_square[name] = buffer.pop()
            self._whole_result_reweight_square[name] = buffer.pop()
            self._whole_result_reweight[name] = buffer.pop()
        self._total_log_ws = buffer.pop()
        self._total_weight_square = buffer.pop()
        self._total_weight = buffer.pop()
        self._count = buffer.pop()
        for name, observers in reversed(list(self._observer.items())):
            for positions in reversed(list(observers)):
                self._result_square_reweight_square[name][positions] = buffer.pop()
                self._result_reweight_square[name][positions] = buffer.pop()
                self._result_reweight[name][positions] = buffer.pop()
        if self._enable_gradient and _dist.world_size() > 1:
            tensors = [t for row in self._Delta for t in row] + [t for row in self._EDelta for t in row]
            _dist.allreduce_tensors(tensors)

    # -- one sample (or one batch of samples) ------------------------------------------------------
    def __call__(self, possibility, configuration):
        owner = self.owner
        nb = configuration.nb
        if self._cache_configuration and (self._pool is None or self._cache_configuration == "drop"):
            from .configuration import ConfigurationPool
            self._pool = ConfigurationPool(owner)
        self._count += nb
        ws = configuration.hole(())
        ws_val = _values(ws)
        possibility = np.broadcast_to(np.asarray(possibility, dtype=np.float64), (nb,))
        alive = ws_val != 0
        if not alive.any():
            return
        with np.errstate(divide="ignore", invalid="ignore"):
            reweight = np.where(alive, ws_val**2 / possibility, 0.0)
        self._total_weight += float(reweight.sum())
        self._total_weight_square += float((reweight**2).sum())
        self._total_log_ws += float(np.log(np.abs(ws_val[alive])).sum())

        ragged = getattr(configuration, "_ragged", False)
        ragged_fermi = ragged and owner.Tensor.Symmetry.is_fermi_symmetry
        # amplitudes of bosonic models are plain numbers: the tensor form below only matters for the fermionic P-edge signs
        no_symmetry = is_no_symmetry(owner.Tensor) or (ragged and not ragged_fermi)
        native = is_native(owner.Tensor)
        if not no_symmetry:
            inv_ws_conj = ws / (ws.norm_2()**2)
            all_name = {("T", "T")} | {(f"P_{l1}_{l2}_{orbit}",) * 2 for l1, l2 in owner.sites() for orbit in owner.physics_edges[l1, l2]}
        Es = None
        for name, observers in self._observer.items():
            whole = np.zeros(nb)
            per_term = []
            for positions, observer in observers.items():
                body = len(positions)
                pending = []
                table = element_table(observer, [owner.physics_edges[p] for p in positions])
                cur = table.flatten([Configuration._index_by_point(table.edges[i], configuration[positions[i]]) for i in range(body)])
                count = table.count[cur]
                total = np.zeros(nb)
                for k in range(int(count.max()) if len(count) else 0):
                    act = (k < count) & alive
                    if not act.any():
                        continue
                    target = np.where(act, table.targets[cur, np.minimum(k, table.kmax - 1)], cur)
                    h = np.where(act, table.values[cur, np.minimum(k, table.kmax - 1)], 0.0)
                    new_idx = table.unflatten(target)
                    replacement = {positions[i]: Configuration._point_by_index(table.edges[i], new_idx[i]) for i in range(body)}
                    if self._restrict_subspace is not None and not self._restrict_subspace(configuration, replacement):
                        continue
                    if self._pool is not None:
                        wss = self._pool.wss(configuration, replacement)
                    else:
                        wss = configuration.replace(replacement)
                        if wss is None:
                            raise NotImplementedError("not implemented replace style, set cache_configuration to True to calculate it")
                    if no_symmetry:
                        # <psi|s'> H_{s's} / <psi|s>  for real amplitudes.  Device tensors: the amplitudes are only queued here
                        # and read back ONCE per observable set (one device -> host copy instead of one blocking copy per
                        # replaced configuration: the host keeps issuing kernels while the GPU works)
                        if native:
                            pending.append((act, h, wss))
                        else:
                            with np.errstate(divide="ignore", invalid="ignore"):
                                total += np.where(act, h * _values(wss) / ws_val, 0.0)
                    elif ragged_fermi:
                        # tensor form for a lock-step batch: the one-element Hamiltonian tensor of every chain (its unit edges carry
                        # that chain's in / out charges, hence parities) contracted like the reference does (observer.py:377-383)
                        cur_idx = table.unflatten(cur)
                        shrunk = _batched_element(observer, table, cur_idx, new_idx, h)
                        pn = [f"P_{l1}_{l2}_{orbit}" for l1, l2, orbit in positions]
                        value = (inv_ws_conj.contract(shrunk, {(pn[i], f"I{i}") for i in range(body)})
                                 .edge_rename({f"O{i}": pn[i] for i in range(body)}).contract(wss.conjugate(), all_name))
                        pending.append((act, np.ones(nb), value))
                    else:
                        # tensor form keeps the fermionic signs of the P edges (observer.py:377-383); one chain
                        if float(wss.norm_max()) == 0:
                            continue
                        edge_in = tuple(configuration[positions[i]] for i in range(body))
                        key_in = tuple((s, int(o[0])) for s, o in edge_in)
                        key_out = tuple((s, int(o[0])) for s, o in (replacement[positions[i]] for i in range(body)))
                        from .tensor_element import tensor_element
                        shrunk = tensor_element(observer)[key_in][key_out]
                        pn = [f"P_{l1}_{l2}_{orbit}" for l1, l2, orbit in positions]
                        value = (inv_ws_conj.contract(shrunk, {(pn[i], f"I{i}") for i in range(body)})
                                 .edge_rename({f"O{i}": pn[i] for i in range(body)}).contract(wss.conjugate(), all_name))
                        total += _values(value)
                per_term.append((positions, total, pending))
            if any(pend for _, _, pend in per_term):  # noqa: E501
                import torch
                flat = [(w.scalar().t.reshape(-1) if ragged else w.data.reshape(-1)).expand(nb) for _, _, pend in per_term for _, _, w in pend]
                host = torch.stack(flat).cpu().numpy()      # [replaced configurations, nb]: the only read-back
                row = 0
                for _, total, pend in per_term:
                    for act, h, _ in pend:
                        with np.errstate(divide="ignore", invalid="ignore"):
                            total += np.where(act, host[row], 0.0) if ragged_fermi else np.where(act, h * host[row] / ws_val, 0.0)
                        row += 1
            for positions, total, _ in per_term:
                r, rr, rsr = self._result_reweight[name], self._result_reweight_square[name], self._result_square_reweight_square[name]
                r[positions] += float((total * reweight).sum())
                rr[positions] += float((total * reweight**2).sum())
                rsr[positions] += float((total**2 * reweight**2).sum())
                whole += total
            if name == "energy" and self._classical_energy is not None:
                whole = whole + self._classical_energy(configuration)
            self._whole_result_reweight[name] += float((whole * reweight).sum())
            self._whole_result_reweight_square[name] += float((whole * reweight**2).sum())
            self._whole_result_square_reweight_square[name] += float((whole**2 * reweight**2).sum())
            if name == "energy":
                Es = whole
        if self._enable_gradient and Es is not None:
            holes = configuration.holes()
            if not is_native(owner.Tensor):
                # generic PyTAT path (one chain): plain tensor arithmetic as the reference (observer.py:399-415)
                w, e = float(reweight[0]), float(Es[0])
                rows = []
                for l1, l2 in owner.sites():
                    hole = holes[l1][l2] * w
                    self._Delta[l1][l2] += hole
                    self._EDelta[l1][l2] += e * hole
                    if self._enable_natural:
                        rows.append(np.array(holes[l1][l2].transpose(self._Delta[l1][l2].names).storage))
                if self._enable_natural:
                    self._Deltas.append((reweight.copy(), Es.copy(), np.concatenate(rows).reshape(1, -1)))
                return
            B = _bk.get()
            w_dev = B.from_numpy(np.ascontiguousarray(reweight))
            e_dev = B.from_numpy(np.ascontiguousarray(Es))
            rows = []
            for l1, l2 in owner.sites():
                hole = holes[l1][l2]
                target = self._Delta[l1][l2]
                if hole.names != target.names:
                    hole = hole.transpose(target.names)
                data = _blocks_of(hole, target) if ragged else hole.data
                if data.shape[0] != nb:
                    data = data.expand(nb, data.shape[1]).contiguous()
                B.grad_accumulate(data, w_dev, e_dev, target.data, self._EDelta[l1][l2].data)
                if self._enable_natural:
                    if not alive.all():
                        # a chain with zero amplitude has 0 / 0 holes: the reference never stores such a sample
                        # (observer.py:333-335 returns early); keep its row exactly zero so that 0 weight x NaN cannot poison the CG
                        import torch
                        data = torch.where(B.from_numpy(alive)[:, None], data, torch.zeros((), dtype=data.dtype, device=data.device))
                    rows.append(data)
            if self._enable_natural:
                import torch
                self._Deltas.append((reweight.copy(), Es.copy(), torch.cat(rows, dim=1)))
        if ragged:
            from ..TAT import ragged as _ragged
            from .sampling import _check_capacity
            _check_capacity()
            _ragged.learning_cycle_done()

    # -- results -------------------------------------------------------------------------------------
    def _expect_and_deviation(self, total_reweight, total_reweight_square, total_square_reweight_square):
        if total_reweight == 0.0 or self._total_weight == 0.0:
            return 0.0, 0.0
        R, ER, RR, ERR, EERR = self._total_weight, total_reweight, self._total_weight_square, total_reweight_square, total_square_reweight_square
        expect = ER / R
        variance = (EERR - 2 * ERR * expect + RR * expect**2) / R**2
        return expect, (variance**0.5 if variance > 0 else 0.0)

    @property
    def instability(self):
        N = self._count
        expect = self._total_weight / N
        variance = self._total_weight_square / N - expect**2
        return (variance**0.5 if variance > 0 else 0.0) / expect

    @property
    def result(self):
        return {name: {positions: self._expect_and_deviation(self._result_reweight[name][positions],
                                                             self._result_reweight_square[name][positions],
                                                             self._result_square_reweight_square[name][positions])
                       for positions in data} for name, data in self._observer.items()}

    @property
    def whole_result(self):
        return {name: self._expect_and_deviation(self._whole_result_reweight[name], self._whole_result_reweight_square[name],
                                                 self._whole_result_square_reweight_square[name]) for name in self._observer}

    @property
    def total_energy(self):
        return self.whole_result["energy"]

    @property
    def energy(self):
        expect, deviation = self.total_energy
        n = self.owner.site_number
        return expect / n, deviation / n

    def _total_energy_value(self):
        return self._whole_result_reweight["energy"] / self._total_weight

    @property
    def gradient(self):
        """2 * conj(<E Delta> - <E><Delta>) per site tensor (observer.py:542-555)"""
        energy = self._total_energy_value()
        owner = self.owner
        out = [[None] * owner.L2 for _ in range(owner.L1)]
        for l1, l2 in owner.sites():
            b = (self._EDelta[l1][l2] / self._total_weight - self._Delta[l1][l2] * (energy / self._total_weight)) * 2.0
            out[l1][l2] = b.conjugate(True)   # lattice_conjugate: trivial metric (utility.py:230-231)
        return out

    def _delta_to_array(self, delta):
        import torch
        return torch.cat([delta[l1][l2].transpose(self._Delta[l1][l2].names).data.reshape(-1) for l1, l2 in self.owner.sites()])

    def _array_to_delta(self, array):
        owner = self.owner
        out = [[None] * owner.L2 for _ in range(owner.L1)]
        index = 0
        for l1, l2 in owner.sites():
            t = self._Delta[l1][l2].same_shape()
            size = t.storage.size
            t._data = array[index:index + size].reshape(1, size).contiguous()
            index += size
            out[l1][l2] = t
        return out

    def natural_gradient_by_conjugate_gradient(self, step, error):
        """SR natural gradient, matrix free (observer.py:576-675): solve (D~^T D~) x = D~^T E~, return 2x."""
        if not is_native(self.owner.Tensor):
            return self._natural_gradient_host(step, error)
        import torch
        B = _bk.get()
        energy = self._total_energy_value()
        delta = self._delta_to_array(self._Delta) / self._total_weight
        if self._Deltas:
            w = np.concatenate([d[0] for d in self._Deltas])
            es = np.concatenate([d[1] for d in self._Deltas])
            rows = torch.cat([d[2] for d in self._Deltas], dim=0)
            param = B.from_numpy(np.sqrt(w / self._total_weight))
            Delta = (rows - delta.reshape(1, -1)) * param.reshape(-1, 1)
            Energy = B.from_numpy((es - energy) * np.sqrt(w / self._total_weight))
        else:
            Delta = B.zeros(0, delta.shape[0])
            Energy = B.zeros(1, 0).reshape(-1)
        self._Deltas = None
        mv = _MatVec(B, Delta)

        def DT(v):
            return _dist.allreduce_device(mv.t(v))

        b = DT(Energy)
        b_square = float(torch.dot(b, b))
        x = torch.zeros_like(b)
        r = b.clone()
        p = r
        r_square = float(torch.dot(r, r))
        t = 0
        while True:
            if t == step:
                break
            if error != 0.0 and b_square != 0 and error**2 > r_square / b_square:
                break
            Dp = mv.n(p)
            # numpy scalars: 0 / 0 of a degenerate sample set gives nan like the reference's numpy arithmetic, no exception
            with np.errstate(divide="ignore", invalid="ignore"):
                alpha = float(np.float64(r_square) / np.float64(_dist.allreduce_number(float(torch.dot(Dp, Dp)))))
            x = x + alpha * p
            r = r - alpha * DT(Dp)
            new_r_square = float(torch.dot(r, r))
            with np.errstate(divide="ignore", invalid="ignore"):
                beta = float(np.float64(new_r_square) / np.float64(r_square))
            r_square = new_r_square
            p = r + beta * p
            t += 1
        x = 2 * x
        out = self._array_to_delta(x)
        return [[t.conjugate(True) for t in row] for row in out]

    def natural_gradient_by_direct_pseudo_inverse(self, r_pinv, a_pinv, libraries=None):
        """SR natural gradient through the pseudo inverse of the Ns x Ns Gram matrix (observer.py:697-899; the reference does this
        with ScaLAPACK pgemm / pheevd, `libraries` names its shared objects and is ignored here):

            D_s = Delta_s - <Delta>,  e_s = conj(E_s) - conj(<E>)        (rows NOT reweighted, as in the reference)
            T = D D^H = U diag(L) U^H,   l^+ = 1 / (l (1 + (num / l)^6)) for l > 0 else 0,   num = r_pinv L_max + a_pinv
            NG = D^H U diag(l^+) U^H e,   return 2 NG

        Off the hot path: the Gram product and the symmetric eigen-decomposition are library calls (cuBLAS / cuSOLVER through
        torch).  Under several ranks the rows are all-gathered first, so every rank solves the same full problem."""
        import torch
        if not self._enable_natural:
            raise RuntimeError("natural gradient is not enabled on this observer")
        energy = self._total_energy_value()
        if is_native(self.owner.Tensor):
            delta = self._delta_to_array(self._Delta) / self._total_weight
            rows = torch.cat([d[2] for d in self._Deltas], dim=0) if self._Deltas else delta.new_zeros((0, delta.shape[0]))
        else:
            owner = self.owner
            delta = torch.from_numpy(np.concatenate([np.array(self._Delta[l1][l2].storage, dtype=np.float64)
                                                     for l1, l2 in owner.sites()]) / self._total_weight)
            rows = torch.from_numpy(np.concatenate([np.asarray(d[2]) for d in self._Deltas], axis=0))
        es = np.concatenate([d[1] for d in self._Deltas]) if self._Deltas else np.zeros(0)
        self._Deltas = None
        D = rows - delta.reshape(1, -1)
        e = torch.from_numpy(np.ascontiguousarray(es - energy)).to(D.device).reshape(-1, 1)
        packed = _dist.allgather_rows(torch.cat([D, e], dim=1))           # one exchange: rows and their energies together
        D, e = packed[:, :-1], packed[:, -1]
        T = D @ D.T
        L, U = torch.linalg.eigh(T)
        num = r_pinv * float(L[-1]) + a_pinv
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            l = L.cpu().numpy()
            l_inv = np.where(l > 0, 1.0 / (l * (1.0 + (num / np.where(l > 0, l, 1.0))**6)), 0.0)
        x = 2.0 * (D.T @ (U @ (torch.from_numpy(l_inv).to(U.device) * (U.T @ e))))
        if is_native(self.owner.Tensor):
            out = self._array_to_delta(x.contiguous())
        else:
            owner = self.owner
            out = [[None] * owner.L2 for _ in range(owner.L1)]
            index, xs = 0, x.cpu().numpy()
            for l1, l2 in owner.sites():
                t_ = self._Delta[l1][l2].same_shape()
                size = len(np.array(self._Delta[l1][l2].storage))
                t_.storage = xs[index:index + size]
                index += size
                out[l1][l2] = t_
        return [[t.conjugate(True) for t in row] for row in out]

    def _natural_gradient_host(self, step, error):
        """the same CG on host arrays, for PyTAT-compatible tensor classes other than this repository's
        (used to time the reference's CPU path through the public PyTAT API only)"""
        owner = self.owner
        energy = self._total_energy_value()
        delta = np.concatenate([np.array(self._Delta[l1][l2].storage, dtype=np.float64) for l1, l2 in owner.sites()]) / self._total_weight
        w = np.concatenate([d[0] for d in self._Deltas])
        es = np.concatenate([d[1] for d in self._Deltas])
        rows = np.concatenate([np.asarray(d[2]) for d in self._Deltas], axis=0)
        self._Deltas = None
        param = np.sqrt(w / self._total_weight)
        Delta = (rows - delta.reshape(1, -1)) * param.reshape(-1, 1)
        Energy = (es - energy) * param

        def DT(v):
            return _dist.allreduce_host(Delta.T @ v)

        b = DT(Energy)
        b_square = float(b @ b)
        x = np.zeros_like(b)
        r = b.copy()
        p = r
        r_square = float(r @ r)
        t = 0
        while t != step:
            if error != 0.0 and b_square != 0 and error**2 > r_square / b_square:
                break
            Dp = Delta @ p
            with np.errstate(divide="ignore", invalid="ignore"):
                alpha = float(np.float64(r_square) / np.float64(_dist.allreduce_number(float(Dp @ Dp))))
            x = x + alpha * p
            r = r - alpha * DT(Dp)
            new_r_square = float(r @ r)
            with np.errstate(divide="ignore", invalid="ignore"):
                beta = float(np.float64(new_r_square) / np.float64(r_square))
            r_square = new_r_square
            p = r + beta * p
            t += 1
        x = 2 * x
        out = [[None] * owner.L2 for _ in range(owner.L1)]
        index = 0
        for l1, l2 in owner.sites():
            t_ = self._Delta[l1][l2].same_shape()
            size = len(np.array(self._Delta[l1][l2].storage))
            t_.storage = x[index:index + size]
            index += size
            out[l1][l2] = t_.conjugate(True)
        return out

    def normalize_lattice(self, target=None):
        """rescale every site tensor by exp(<log|ws|>/(L1 L2)) (observer.py:909-920); `target`: the lattice to rescale when the
        observer watched its charge-dense embedding"""
        mean_log_ws = self._total_log_ws / self._count
        param = float(np.exp(mean_log_ws / (self.owner.L1 * self.owner.L2)))
        owner = self.owner if target is None else target
        for l1, l2 in owner.sites():
            owner[l1, l2] = owner[l1, l2] / param


class _MatVec:
    """D~ v and D~^T v with the grouped GEMM kernel (Ns_local x Np matrix, vectors as n = 1 GEMMs)."""

    def __init__(self, B, matrix):
        self.B = B
        self.M = matrix.contiguous()
        self.ns, self.np_ = self.M.shape

    class _P:
        _dev = None

    def _gemm(self, m, n, k, flags, a, b):
        p = self._P()
        p.gemm = np.array([[m, n, k, 0, 0, 0, flags, 1]], dtype=np.int64)
        p._dev = None
        out = self.B.zeros(1, m * n)
        if m and n and k:
            self.B.gemm(p, a.reshape(1, -1), b.reshape(1, -1), out)
        return out.reshape(-1)

    def n(self, v):      # [ns]
        return self._gemm(self.ns, 1, self.np_, 0, self.M, v.contiguous())

    def t(self, v):      # [np]  (A stored [k x m] with k = ns, m = np)
        return self._gemm(self.np_, 1, self.ns, 1, self.M, v.contiguous())

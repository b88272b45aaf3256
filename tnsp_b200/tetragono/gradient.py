"""Driver loop of the sampling VMC: `gradient_descent`, the generator behind the reference's
`gm_run` (sampling_lattice/gradient.py:93-445), with the same keyword names, defaults and update rules.

What changes against the reference is only what the B200 design needs:
* `chains` (new, default 1): Markov chains per GPU advanced in lock step.  One "sampling step" of
  the reference is one sample; here `sampling_total_step` still counts SAMPLES over all ranks, every call of the
  sampler yields `chains` of them, rank r of G takes every G-th call (gradient.py:353) -- so with chains == 1 and one
  rank the loop is the reference's loop, sample for sample.
* the energy / Delta / E-Delta / SR-CG exchanges are `torch.distributed` collectives inside `Observer`
  (observer.py:83-126, 639-664 of the reference), launched by `Observer.__exit__` and the CG iteration.
* every rank applies the identical all-reduced update, `bcast_lattice` (lattice.py:950-954) is therefore a no-op kept
  for the call sequence.

Entry points that SURVEY.md section 8 marks out of the path raise NotImplementedError instead of silently doing something
else: numerical check-difference (debug aid).  Direct sampling (8f-1, the reference's default) is `direct_sampling.py`:
`chains` configurations per call for models without symmetry.  Gauge fixing (8f-3) is `SamplingLattice.expand_dimension`.  The pseudo-inverse SR (8f-4) runs on
library eigen-solvers instead of ScaLAPACK.
State and configuration files are written in the reference's own formats (checkpoint.py).
"""
from __future__ import annotations

from datetime import datetime

import numpy as np

from .. import dist as _dist
from ..TAT import random as _random
from .observer import Observer
from .sampling import ChainRng, ErgodicSampling, SweepSampling


def _call_or_import(what, attribute):
    """`get_imported_function` of the reference (utility.py:323-337): a callable, or a module name exporting `attribute`"""
    if isinstance(what, str):
        import importlib
        return getattr(importlib.import_module(what), attribute)
    return what


def lattice_dot(a, b):
    """sum over sites of <a, b> (lattice.py:921-934 with `lattice_conjugate` / `lattice_prod_sum`, utility.py:277-300)"""
    total = 0.0
    for row_a, row_b in zip(a, b):
        for ta, tb in zip(row_a, row_b):
            total += float(ta.conjugate().contract(tb, {(n, n) for n in ta.names}))
    return total


def _lattice_of(state):
    return [[state[l1, l2] for l2 in range(state.L2)] for l1 in range(state.L1)]


def _randomize(grad, rng):
    """keep the sign of every element, draw its modulus uniformly in [0, 1) (`lattice_randomize`, utility.py:303-320)"""
    out = []
    for row in grad:
        new_row = []
        for t in row:
            r = t.copy()
            s = np.asarray(t.storage)
            r.storage = np.sign(s) * rng.random(s.shape)
            new_row.append(r)
        out.append(new_row)
    return out


def line_search(state, observer, grad, energy_observer, configuration_pool, step_size, line_search_amplitude):
    """one probe step: grow the step if the gradient still points the same way after it, shrink otherwise
    (gradient.py:55-91)"""
    saved = _lattice_of(state)
    # every rank holds the same all-reduced gradients: the reference's `bcast_number` needs no collective here
    dot_begin = lattice_dot(grad, observer.gradient)
    if dot_begin > 0:
        for l1, l2 in state.sites():
            state[l1, l2] = state[l1, l2] - grad[l1][l2] * step_size
        with energy_observer:
            for possibility, configuration in configuration_pool:
                configuration.refresh_all()
                energy_observer(possibility, configuration)
        dot_end = lattice_dot(grad, energy_observer.gradient)
        for l1, l2 in state.sites():
            state[l1, l2] = saved[l1][l2]
        if dot_end > 0:
            step_size *= line_search_amplitude
        else:
            step_size /= line_search_amplitude
    return step_size


def gradient_descent(
        state,
        sampling_total_step=0,
        grad_total_step=1,
        grad_step_size=0,
        *,
        # lock-step batch (new)
        chains=1,
        chain_seeds=None,
        # About observer
        cache_configuration=False,
        classical_energy=None,
        # About sampling
        sampling_method="direct",
        configuration_cut_dimension=None,
        direct_sampling_cut_dimension=4,
        sampling_configurations=None,
        sweep_hopping_hamiltonians=None,
        # About subspace
        restrict_subspace=None,
        # About gradient method
        use_check_difference=False,
        use_line_search=False,
        use_fix_relative_step_size=False,
        use_random_gradient=False,
        momentum_parameter=0.0,
        # About natural gradient
        use_natural_gradient=False,
        conjugate_gradient_method_step=20,
        conjugate_gradient_method_error=0.0,
        cache_natural_delta=None,
        use_natural_gradient_by_direct_pseudo_inverse=False,
        scalapack_libraries="libscalapack.so",
        natural_gradient_r_pinv=1e-12,
        natural_gradient_a_pinv=0,
        # About gauge fixing
        fix_gauge=False,
        # About log and save state
        log_file=None,
        save_state_file=None,
        save_configuration_file=None,
        # About line search
        line_search_amplitude=1.2,
        line_search_parameter=0.6,
        # About momentum
        orthogonalize_momentum=False,
        # About Measurement
        measurement=None):
    """Generator: one `(whole_result, result)` pair per optimisation step, like the reference.

    `sampling_configurations` is the start configuration of the sweep sampler: an int64 array `[L1, L2, orbits]` of total
    physical edge indices (-1: no such orbit), or `[chains, L1, L2, orbits]` with one configuration per chain, as written by
    `Configuration.export_configuration`; the last configurations of every step are written back into it (resized in place
    when it held a single configuration, gradient.py:360-364)."""
    if sampling_method not in ("sweep", "ergodic", "direct"):
        raise ValueError("Invalid sampling method")
    if use_check_difference:
        raise NotImplementedError("check_difference is a debugging aid outside the hot path")

    time_str = datetime.now().strftime("%Y-%m-%d-%H:%M:%S")
    rank, size = _dist.rank(), _dist.world_size()
    use_gradient = grad_step_size != 0
    if not use_gradient:
        grad_total_step = 1

    restrict = _call_or_import(restrict_subspace, "restrict") if restrict_subspace is not None else None
    if classical_energy is not None:
        classical_energy = _call_or_import(classical_energy, "classical_energy")

    observer = Observer(state, enable_energy=True, enable_gradient=use_gradient, enable_natural_gradient=use_natural_gradient,
                        cache_natural_delta=cache_natural_delta, cache_configuration=cache_configuration,
                        restrict_subspace=restrict, classical_energy=classical_energy)
    if measurement:
        if isinstance(measurement, str):
            measurement = measurement.split(",")
        if not isinstance(measurement, list):
            measurement = [measurement]
        for term in measurement:
            if isinstance(term, str):
                observer.add_observer(term, _call_or_import(term, "measurement")(state))
            else:
                observer.add_observer(term.__name__, term(state))
    need_energy_observer = use_gradient and use_line_search
    if need_energy_observer:
        energy_observer = Observer(state, enable_energy=True, enable_gradient=True, cache_configuration=cache_configuration,
                                   restrict_subspace=restrict, classical_energy=classical_energy)

    # Lock-step chains of a bosonic-symmetric model run on its charge-dense embedding: the embedding is
    # rebuilt from the symmetric parameters at the start of every step, sampled and observed, and the gradient is projected back
    # onto the symmetric blocks (it is exactly zero outside them), so the optimisation itself stays in the symmetric picture.
    from . import configuration as _configuration
    sector = (chains > 1 and state.Tensor.Symmetry.length != 0 and _configuration.LOCKSTEP_ENGINE == "sector"
              and _configuration.sector_engine_supported(state.Tensor))
    embedded = chains > 1 and state.Tensor.Symmetry.length != 0 and not sector
    if embedded:
        if state.Tensor.Symmetry.is_fermi_symmetry:
            raise NotImplementedError("lock-step chains of fermionic lattices need the sector-compact engine (integer symmetries)")
        if sampling_method != "sweep" or measurement or use_line_search or restrict is not None or classical_energy is not None:
            raise NotImplementedError("the embedded lock-step path supports sweep sampling with energy / gradient / SR only")
        from . import dense_embedding

    # random engines: the reference re-seeds every process around the sampling phase of each step (`seed_differ`,
    # utility.py:138-160): seed = global uniform_int + rank, one uniform_real discarded; afterwards all processes are
    # put back on a common seed.  Chain c of rank r plays the role of process r * chains + c.
    rng = ChainRng(chains) if (chains > 1 or chain_seeds is not None) else None
    if rng is not None and chain_seeds is not None:
        rng.seed(list(chain_seeds))
        rng.uniform_real(None)
    max_int = 2**31
    random_int = _random.uniform_int(0, max_int - 1)

    def seed_differ_enter():
        base = random_int()
        if rng is None:
            _random.seed((base + rank) % max_int)
            _random.uniform_real(0, 1)()
        elif chain_seeds is None:
            rng.seed([(base + rank * chains + c) % max_int for c in range(chains)])
            rng.uniform_real(None)

    def seed_differ_exit():
        _random.seed(int(_dist.allreduce_number(random_int() // size)))

    host_rng = np.random.default_rng(2333 + rank)
    total_grad = None
    configuration = None

    for grad_step in range(grad_total_step):
        configuration_pool = []
        seed_differ_enter()
        work = state
        if embedded:
            work = dense_embedding.embed_lattice(state)
            observer = Observer(work, enable_energy=True, enable_gradient=use_gradient, enable_natural_gradient=use_natural_gradient)
        with observer:
            if sampling_method == "sweep":
                hopping = None
                if sweep_hopping_hamiltonians is not None:
                    hopping = _call_or_import(sweep_hopping_hamiltonians, "hopping_hamiltonians")(work)
                sampling = SweepSampling(work, configuration_cut_dimension, restrict, hopping, nb=chains, rng=rng)
                if configuration is not None:
                    sampling.configuration.import_configuration(configuration.export_configuration())
                elif sampling_configurations is not None and np.size(sampling_configurations) != 0:
                    # [L1, L2, orbits] (one start configuration for every chain) or [chains, L1, L2, orbits] as written by
                    # `Configuration.export_configuration`; import_configuration broadcasts the former itself
                    conf = np.asarray(sampling_configurations)
                    if conf.ndim == 4 and conf.shape[0] != chains:
                        raise ValueError(f"sampling_configurations holds {conf.shape[0]} configurations for {chains} chains")
                    sampling.configuration.import_configuration(conf)
                else:
                    raise RuntimeError("sweep sampling needs an initial configuration (sampling_configurations)")
                calls = -(-sampling_total_step // chains)
            elif sampling_method == "direct":
                from .direct_sampling import DirectSampling
                sampling = DirectSampling(state, configuration_cut_dimension, restrict, direct_sampling_cut_dimension, nb=chains, rng=rng)
                calls = -(-sampling_total_step // chains)
            else:
                sampling = ErgodicSampling(state, configuration_cut_dimension, restrict, rank=rank, size=size, nb=chains)
                calls = sampling.calls          # == total_step for chains == 1
            for sampling_step in range(calls):
                if sampling_step % size == rank:
                    possibility, configuration = sampling()
                    observer(possibility, configuration)
                    if need_energy_observer:
                        configuration_pool.append((possibility, configuration))
            if sampling_method != "ergodic" and configuration is not None and isinstance(sampling_configurations, np.ndarray):
                # keep the caller's array current (gradient.py:360-364 resizes it in place too): the configuration file written
                # below and the next call of the driver start from the last configurations of this step
                new_conf = configuration.export_configuration()
                if sampling_configurations.shape != new_conf.shape:
                    try:
                        sampling_configurations.resize(new_conf.shape, refcheck=False)
                    except ValueError:
                        sampling_configurations = np.array(new_conf)
                np.copyto(sampling_configurations, new_conf)
        seed_differ_exit()

        measurement_result = observer.result
        measurement_whole_result = observer.whole_result
        if measurement is not None and rank == 0:
            for term in measurement:
                if isinstance(term, str):
                    _call_or_import(term, "save_result")(state, measurement_result[term], measurement_whole_result[term])
        if log_file and rank == 0:
            with open(log_file.replace("%t", time_str), "a", encoding="utf-8") as file:
                print(*observer.energy, file=file)

        if use_gradient:
            if use_natural_gradient and use_natural_gradient_by_direct_pseudo_inverse:
                grad = observer.natural_gradient_by_direct_pseudo_inverse(natural_gradient_r_pinv, natural_gradient_a_pinv,
                                                                          scalapack_libraries.split(","))
            elif use_natural_gradient:
                grad = observer.natural_gradient_by_conjugate_gradient(conjugate_gradient_method_step, conjugate_gradient_method_error)
            else:
                grad = observer.gradient
            if embedded:
                grad = dense_embedding.project_gradient(state, grad)

            if use_line_search:
                scale = (lattice_dot(_lattice_of(state), _lattice_of(state)) / lattice_dot(grad, grad))**0.5
                grad = [[g * scale for g in row] for row in grad]
                grad_step_size = line_search(state, observer, grad, energy_observer, configuration_pool, grad_step_size,
                                             line_search_amplitude)
                state.apply_gradient(grad, grad_step_size * line_search_parameter)
            else:
                if grad_step == 0 or momentum_parameter == 0.0:
                    total_grad = grad
                else:
                    if orthogonalize_momentum:
                        mine = _lattice_of(state)
                        param = lattice_dot(mine, total_grad) / lattice_dot(mine, mine)
                        total_grad = [[t - s * param for t, s in zip(row_t, row_s)] for row_t, row_s in zip(total_grad, mine)]
                    total_grad = [[t * momentum_parameter + g * (1 - momentum_parameter) for t, g in zip(row_t, row_g)]
                                  for row_t, row_g in zip(total_grad, grad)]
                this_grad = _randomize(total_grad, host_rng) if use_random_gradient else total_grad
                if use_fix_relative_step_size:
                    scale = (lattice_dot(_lattice_of(state), _lattice_of(state)) / lattice_dot(this_grad, this_grad))**0.5
                    this_grad = [[g * scale for g in row] for row in this_grad]
                    if not use_random_gradient:
                        total_grad = this_grad   # the reference scales in place: the momentum carries the scaled update
                state.apply_gradient(this_grad, grad_step_size)

            if fix_gauge:
                state.expand_dimension(1.0, 0)
            observer.normalize_lattice(state if embedded else None)
            bcast_lattice(state)

        if embedded:
            # the embedding's process-wide backend switches end with the step (the next step's embed_lattice sets them again): a truly
            # dense model evaluated by the caller between two steps, or after the last one, takes the dense paths
            dense_embedding.release_backend_flags()
        yield (measurement_whole_result, measurement_result)

        # checkpoints in the reference's own formats (utility.py:340-418): its `pickle.load` / `read_configurations` read them
        if save_state_file and rank == 0:
            from .checkpoint import save_reference_state
            save_reference_state(state, save_state_file.replace("%s", str(grad_step)).replace("%t", time_str))
        if save_configuration_file and isinstance(sampling_configurations, np.ndarray):
            from .checkpoint import write_configurations
            write_configurations(sampling_configurations, save_configuration_file.replace("%s", str(grad_step)).replace("%t", time_str))

def bcast_lattice(state, root=0):
    """lattice.py:950-954.  Every rank has applied the same all-reduced update, the parameters are already identical;
    under a multi-rank run they are still broadcast once so that rounding differences of rank-local reductions can never
    accumulate."""
    if _dist.world_size() > 1:
        _dist.broadcast_tensors([state[l1, l2] for l1, l2 in state.sites()], root=root)

"""bench.py -- VMC samples/sec of the sampling hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA path)
    python bench.py --impl reference --steps K --warmup W    # reference CPU implementation on the host cores

One "step" = one lock-step sweep of all chains of a rank (SweepSampling.__call__) + one Observer call
(local energy + log-derivative accumulation), i.e. `chains` samples per rank; value = samples of all
ranks / max-over-ranks device time.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: dict(L1, L2, D, Dc, sym, J2, sr (SR natural gradient by CG with `cg` iterations per step), chains, desc)
    "cfg1": dict(L1=4, L2=4, D=4, Dc=16, sym="No", J2=0.0, sr=False, cg=0, chains=4096,
                 desc="tetragono sampling VMC 4x4 Heisenberg square lattice, no symmetry, D=4, Dc=16, float64"),
    # chains: 20 x 148; measured 2258 samples/s at 2960 chains (120 GB) against 1970 - 2090 at 2368 (104 - 111 GB): the kernels of a larger
    # batch outlast more of the host's per-launch time (profiles/r02_bench_cfg2_nb2960.json)
    "cfg2": dict(L1=6, L2=6, D=6, Dc=36, sym="BoseU1", J2=0.5, sr=True, cg=20, chains=2960,
                 desc="6x6 J1-J2 Heisenberg (J2=0.5) with U(1) symmetry, D=6 (2+2+2), Dc=36, sweep sampling + SR natural gradient (CG 20), float64"),
    "cfg2s": dict(L1=4, L2=4, D=3, Dc=9, sym="BoseU1", J2=0.5, sr=True, cg=4, chains=64,
                  desc="4x4 J1-J2 Heisenberg with U(1) symmetry, D=3, Dc=9, sweep + SR (smoke size of cfg2)"),
    "heis6": dict(L1=6, L2=6, D=6, Dc=36, sym="No", J2=0.0, sr=False, cg=0, chains=148,
                  desc="6x6 Heisenberg, no symmetry (dense tensors), D=6, Dc=36, float64"),
    "tiny": dict(L1=3, L2=3, D=2, Dc=4, sym="No", J2=0.0, sr=False, cg=0, chains=64, desc="3x3 Heisenberg, no symmetry, D=2, Dc=4 (smoke size)"),
    # fermionic lock-step chains (sector-compact engine; bond profiles as SURVEY.md 8d prescribes, tetragono/models.py)
    "cfg3": dict(model="hubbard_ff", L1=8, L2=8, D=8, Dc=64, sym="FermiU1FermiU1", T=64, U=8.0, sr=False, cg=0, chains=74,
                 desc="8x8 fermionic Hubbard (FermiU1 x FermiU1 edges, half filling, U/t = 8), D=8 (2+1+1+1+1+1+1 over the charge "
                      "fluctuations of a bond), Dc=64, sweep sampling + gradient, float64"),
    "cfg3s": dict(model="hubbard_ff", L1=4, L2=4, D=8, Dc=16, sym="FermiU1FermiU1", T=16, U=8.0, sr=False, cg=0, chains=64,
                  desc="4x4 fermionic Hubbard (FermiU1 x FermiU1), D=8, Dc=16 (smoke size of cfg3)"),
    "cfg4": dict(model="tJ", L1=10, L2=10, D=10, Dc=100, sym="FermiU1BoseU1", T=40, J=0.4, sr=True, cg=20, chains=37,
                 desc="10x10 t-J model (FermiU1 x BoseU1 edges, 80 particles), D=10 (1,1,1,1,2,1,1,1,1), Dc=100, sweep sampling "
                      "(ergodic enumeration of 3^100 configurations is not feasible) + gradient + SR natural gradient (CG 20), float64"),
    "cfg4s": dict(model="tJ", L1=4, L2=4, D=10, Dc=16, sym="FermiU1BoseU1", T=2, J=0.4, sr=True, cg=4, chains=64,
                  desc="4x4 t-J model (FermiU1 x BoseU1), D=10 (1,1,1,1,2,1,1,1,1), Dc=16 (smoke size of cfg4)"),
}


def build_workload(tat, wl, state_classes=None):
    """(SamplingLattice on `tat`'s tensors, sweep hopping terms or None, Neel edge points).  The same function
    builds the model for this repository's device tensors and for the reference PyTAT classes."""
    from tnsp_b200.tetragono import models
    from tnsp_b200.tetragono.state import SamplingLattice
    if wl.get("model") in ("hubbard_ff", "tJ"):
        return build_fermionic_workload(tat, wl)
    T = getattr(tat, wl["sym"]).D.Tensor
    abstract = models.j1j2_abstract_lattice(T, wl["L1"], wl["L2"], wl["D"], 1.0, wl["J2"])
    tat.random.seed(2333)
    lat = SamplingLattice(abstract)
    hopping = models.nearest_neighbour_terms(lat) if wl["J2"] != 0 else None
    return lat, hopping, models.neel_points(lat)


def fermionic_start(wl):
    """total physical indices [L1, L2, 1] of the start configuration: the particles of every row staggered, rows shifted by one site.
    Hubbard: physical edge (empty, down, up, double) -> alternating up / down (half filling); t-J: (hole, down, up), T / L1 up and
    T / L1 down particles per row, the holes spread evenly (for 4 x 4, T = 2: the fixture's pattern, particles in rows 0 and 2)"""
    L1, L2 = wl["L1"], wl["L2"]
    out = np.zeros((L1, L2, 1), dtype=np.int64)
    if wl["model"] == "hubbard_ff":
        for l1 in range(L1):
            for l2 in range(L2):
                out[l1, l2, 0] = 2 if (l1 + l2) % 2 == 0 else 1
        return out
    per_row = 2 * wl["T"] // L1
    if per_row == 0 or (2 * wl["T"]) % L1:
        rows = [r for r in range(L1) if r % 2 == 0][:wl["T"]]          # one up + one down in every other row
        for k, r in enumerate(rows):
            out[r, 0, 0], out[r, 2, 0] = (2, 1) if k % 2 == 0 else (1, 2)
        return out
    for l1 in range(L1):
        holes = L2 - per_row
        hole_at = {int((h + 1) * L2 / holes) - 1 for h in range(holes)} if holes else set()
        k = l1
        for l2 in range(L2):
            if l2 in hole_at:
                continue
            out[l1, l2, 0] = 2 if k % 2 == 0 else 1
            k += 1
    return out


def build_fermionic_workload(tat, wl):
    """cfg3 / cfg4 families on this repository's tensors (the reference arm builds the same model with the reference's own tetraku
    modules, `_reference_fermionic_workload`)"""
    from tnsp_b200.tetragono import models
    from tnsp_b200.tetragono.state import SamplingLattice
    L1, L2 = wl["L1"], wl["L2"]
    if wl["model"] == "hubbard_ff":
        D = models.HUBBARD_D8 if wl["D"] == 8 else wl["D"]
        abstract = models.hubbard_fermi_fermi_abstract_lattice(L1, L2, D, wl["T"], 1.0, wl["U"])
    else:
        D = models.TJ_D10 if wl["D"] == 10 else wl["D"]
        abstract = models.tJ_abstract_lattice(L1, L2, D, wl["T"], 1.0, wl["J"])
    tat.random.seed(2333)
    lat = SamplingLattice(abstract)
    start = fermionic_start(wl)
    points = [[{0: lat.physics_edges[l1, l2, 0].point_by_index(int(start[l1, l2, 0]))} for l2 in range(L2)] for l1 in range(L1)]
    return lat, None, points


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own C++ TAT (oracle/_ref, LAPACK/BLAS per sector on the CPU) driven by
# the same-semantics Python drivers, one independent Markov chain per host core
# ---------------------------------------------------------------------------------------------------
def _stock_reference_lattice(wl):
    """the workload built with the reference's OWN modules (TAT = oracle/_ref, tetragono / tetraku = oracle/_ref/site): same model,
    same PEPS (TAT.random.seed(2333) + randn_ in site order) as `build_workload` builds on this repository's tensors"""
    import TAT
    import tetragono as tet
    L1, L2 = wl["L1"], wl["L2"]
    model = wl.get("model")
    if model == "hubbard_ff":
        from tetraku.models.hubbard.fermi_fermi import abstract_state
        T = wl["T"]
        state = tet.AbstractLattice(abstract_state(L1, L2, T, 1.0, wl["U"]))
        half = T // 2
        prof = {(0, 0): 2, (1, 0): 1, (-1, 0): 1, (0, 1): 1, (0, -1): 1, (1, -1): 1, (-1, 1): 1} if wl["D"] == 8 else \
            {(a, b): wl["D"] for a in (-1, 0, 1) for b in (-1, 0, 1)}
        total = sum(prof.values())

        def charged(Q):
            return [((Q + a, Q + b), prof[a, b]) for a in (-1, 0, 1) for b in (-1, 0, 1) if prof.get((a, b), 0) > 0]

        per_row = half / L1
        for l1 in range(L1 - 1):
            state.virtual_bond[l1, 0, "D"] = charged(int(half * (L1 - l1 - 1) / L1))
            for l2 in range(1, L2):
                state.virtual_bond[l1, l2, "D"] = [((0, 0), total)]
        for l1 in range(L1):
            for l2 in range(L2 - 1):
                state.virtual_bond[l1, l2, "R"] = charged(int(per_row * (L2 - l2 - 1) / L2))
        hopping = None
    elif model == "tJ":
        from tetraku.models.tJ import abstract_state
        T = wl["T"]
        state = tet.AbstractLattice(abstract_state(L1, L2, T, 1.0, wl["J"]))
        prof = [1, 1, 1, 1, 2, 1, 1, 1, 1] if wl["D"] == 10 else [wl["D"]] * 9
        sectors = [(dn, ds) for dn, spins in ((-2, (0,)), (-1, (-1, 1)), (0, (-2, 0, 2)), (1, (-1, 1)), (2, (0,))) for ds in spins]

        def charged(Q):
            return [((2 * Q + dn, ds), d) for (dn, ds), d in zip(sectors, prof) if d > 0]

        per_row = T / L1
        for l1 in range(L1 - 1):
            state.virtual_bond[l1, 0, "D"] = charged(int(T * (L1 - l1 - 1) / L1))
            for l2 in range(1, L2):
                state.virtual_bond[l1, l2, "D"] = [((0, 0), sum(prof))]
        for l1 in range(L1):
            for l2 in range(L2 - 1):
                state.virtual_bond[l1, l2, "R"] = charged(int(per_row * (L2 - l2 - 1) / L2))
        hopping = None
    elif wl["sym"] == "No":
        from tetraku.models.heisenberg import abstract_lattice
        state = abstract_lattice(L1, L2, wl["D"], 1.0)
        hopping = None
    else:
        # J1-J2 with U(1) (2 Sz) tensors: the reference ships this model without symmetry only
        Tn = TAT.BoseU1.D.Tensor
        st = tet.AbstractState(Tn, L1, L2)
        pe, cpe = [(+1, 1), (-1, 1)], [(-1, 1), (+1, 1)]
        st.physics_edges[...] = pe
        SS = Tn(["I0", "I1", "O0", "O1"], [cpe, cpe, pe, pe]).zero_()
        up, dn = (1, 0), (-1, 0)
        for i0, i1, o0, o1, v in ((up, up, up, up, 0.25), (dn, dn, dn, dn, 0.25), (up, dn, up, dn, -0.25), (dn, up, dn, up, -0.25),
                                  (up, dn, dn, up, 0.5), (dn, up, up, dn, 0.5)):
            SS[{"I0": (-i0[0], 0), "I1": (-i1[0], 0), "O0": o0, "O1": o1}] = v
        H = -1.0 * SS
        st.hamiltonians["vertical_bond"] = H
        st.hamiltonians["horizontal_bond"] = H
        state = tet.AbstractLattice(st)
        d = wl["D"] // 3
        state.virtual_bond["R"] = [(-1, d), (0, d), (+1, d)]
        state.virtual_bond["D"] = [(-1, d), (0, d), (+1, d)]
        hopping = "nn"
    TAT.random.seed(2333)
    lat = tet.SamplingLattice(state)
    if hopping == "nn":
        H1 = lat._hamiltonians[((0, 0, 0), (0, 1, 0))]
        hopping = dict(lat._hamiltonians)
        if wl.get("J2", 0.0) != 0:
            H2 = wl["J2"] * H1
            for l1 in range(L1 - 1):
                for l2 in range(L2 - 1):
                    lat.hamiltonians[(l1, l2, 0), (l1 + 1, l2 + 1, 0)] = H2
                    lat.hamiltonians[(l1, l2 + 1, 0), (l1 + 1, l2, 0)] = H2
        else:
            hopping = None
    return lat, hopping


def _stock_reference_worker(args):
    """ONE independent Markov chain of the UNMODIFIED reference: its C++ TAT (oracle/_ref/TAT*.so), its tetragono / tetraku / lazy
    (oracle/_ref/site, installed there by `make -C oracle pyref`) and the single-rank mpi4py stand-in.  Nothing of tnsp_b200 is
    imported in this process."""
    workload, seed, n_warm, n_samples = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    ref = os.path.join(ROOT, "oracle", "_ref")
    sys.path[:0] = [ref, os.path.join(ref, "site")]
    import TAT
    import tetragono as tet
    wl = dict(WORKLOADS[workload])
    wl.setdefault("J2", 0.0)
    lat, hopping = _stock_reference_lattice(wl)
    TAT.random.seed(seed)
    s = tet.sampling_lattice.SweepSampling(lat, wl["Dc"], None, hopping)
    if wl.get("model") in ("hubbard_ff", "tJ"):
        start = fermionic_start(wl)
    else:
        start = np.array([[[(l1 + l2) % 2] for l2 in range(wl["L2"])] for l1 in range(wl["L1"])])
    for l1 in range(wl["L1"]):
        for l2 in range(wl["L2"]):
            s.configuration[l1, l2, 0] = lat.physics_edges[l1, l2, 0].point_by_index(int(start[l1, l2, 0]))
    obs = tet.sampling_lattice.Observer(lat, enable_energy=True, enable_gradient=True, enable_natural_gradient=wl["sr"])
    with obs:
        for _ in range(n_warm):
            p, c = s()
            obs(p, c)
    t0 = time.perf_counter()
    with obs:
        for _ in range(n_samples):
            p, c = s()
            obs(p, c)
    g = obs.natural_gradient_by_conjugate_gradient(wl["cg"], 0.0) if wl["sr"] else obs.gradient   # noqa: F841
    dt = time.perf_counter() - t0
    return n_samples, dt, obs.energy[0]


def _reference_worker(args):
    """fallback when oracle/_ref/site is absent: the reference's C++ TAT (or, without it, the numpy port) under this repository's
    drivers -- measured equivalent to the stock path, but not the stock code"""
    workload, seed, n_warm, n_samples, use_ref = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    wl = dict(WORKLOADS[workload])
    wl.setdefault("J2", 0.0)
    if use_ref:
        from oracle.ref import load_reference_tat
        tat = load_reference_tat()
    else:
        tat = None
    from tnsp_b200.tetragono.observer import Observer
    from tnsp_b200.tetragono.sampling import SweepSampling
    import tnsp_b200.TAT.random as rnd
    if tat is None:
        # "port": the numpy checker backend under this repository's planner
        from oracle import numpy_backend
        numpy_backend.install()
        import tnsp_b200.TAT as tat
    lat, hopping, points = build_workload(tat, wl)
    rnd.seed(seed)
    s = SweepSampling(lat, wl["Dc"], None, hopping)
    for l1, row in enumerate(points):
        for l2, site in enumerate(row):
            for o, pt in site.items():
                s.configuration[l1, l2, o] = pt
    obs = Observer(lat, enable_energy=True, enable_gradient=True, enable_natural_gradient=wl["sr"])
    with obs:
        for _ in range(n_warm):
            p, c = s()
            obs(p, c)
    t0 = time.perf_counter()
    with obs:
        for _ in range(n_samples):
            p, c = s()
            obs(p, c)
    g = obs.natural_gradient_by_conjugate_gradient(wl["cg"], 0.0) if wl["sr"] else obs.gradient   # noqa: F841
    dt = time.perf_counter() - t0
    return n_samples, dt, obs.energy[0]


def _plan_flops_worker(args):
    """GEMM flops of ONE sample (sweep + observe) as the contraction plans enumerate them: sum over sector GEMMs of 2 m n k
    (SURVEY.md 8d).  Runs this repository's host planner on the CPU checker backend (oracle/, cpu_baseline leg only) twice:
    on the block-symmetric U(1) tensors -- the reference's own sector plan -- and on their charge-dense embedding, which is
    what the lock-step GPU engine executes."""
    workload, = args
    wl = WORKLOADS[workload]
    from oracle import numpy_backend
    numpy_backend.install()
    from tnsp_b200 import backend
    import tnsp_b200.TAT as tat
    import tnsp_b200.TAT.random as rnd
    from tnsp_b200.tetragono import dense_embedding, models
    from tnsp_b200.tetragono.observer import Observer
    from tnsp_b200.tetragono.sampling import SweepSampling
    B = backend.get()
    count = {"flops": 0.0}
    gemm0, gather0 = B.gemm, B.gemm_gather

    def gemm(plan, a, b, c):
        count["flops"] += float(sum(2.0 * int(m) * int(n) * int(k) for m, n, k, *_ in plan.gemm))
        return gemm0(plan, a, b, c)

    def gemm_gather(plan, a, b, c):
        _, _, m, n, k = plan.gather
        count["flops"] += 2.0 * m * n * k
        return gather0(plan, a, b, c)

    B.gemm, B.gemm_gather = gemm, gemm_gather
    sym_lat, hopping, points = build_workload(tat, wl)
    out = {}
    for which in ("sector_plan", "dense_embedding"):
        if which == "sector_plan":
            lat, hop = sym_lat, hopping
            s = SweepSampling(lat, wl["Dc"], None, hop)
            for l1, row in enumerate(points):
                for l2, site in enumerate(row):
                    for o, pt in site.items():
                        s.configuration[l1, l2, o] = pt
        else:
            if wl["sym"] == "No":
                out[which] = out["sector_plan"]
                continue
            lat = dense_embedding.embed_lattice(sym_lat)
            hop = models.nearest_neighbour_terms(lat) if hopping is not None else None
            s = SweepSampling(lat, wl["Dc"], None, hop)
            s.configuration.import_configuration(dense_embedding.embed_configuration(sym_lat, points))
        rnd.seed(4321)
        obs = Observer(lat, enable_energy=True, enable_gradient=True)
        with obs:                      # first sample: environments built from scratch
            p, c = s()
            obs(p, c)
        count["flops"] = 0.0
        with obs:                      # steady state: one sweep + one observation
            p, c = s()
            obs(p, c)
        out[which] = count["flops"]
    return out


def run_reference(workload, steps, warmup, samples_per_step, cores=None, plan_flops=False):
    """returns dict(value, cores, kind, sample, ms_per_step[, plan_flops])"""
    import glob
    use_ref = bool(glob.glob(os.path.join(ROOT, "oracle", "_ref", "TAT*.so")))
    stock = use_ref and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "site", "tetragono"))
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    per_step = []
    flops = None
    with ctx.Pool(cores) as pool:
        if use_ref:
            ok = pool.map(_probe_ref, range(1))[0]
            use_ref = ok
            stock = stock and ok
        for st in range(warmup + steps):
            t0 = time.perf_counter()
            if stock:
                res = pool.map(_stock_reference_worker, [(workload, 1000 + 97 * st + c, 1, samples_per_step) for c in range(cores)])
            else:
                res = pool.map(_reference_worker, [(workload, 1000 + 97 * st + c, 1, samples_per_step, use_ref) for c in range(cores)])
            wall = time.perf_counter() - t0
            # throughput of the step: every core runs one independent chain (sum of per-chain rates)
            rate = sum(n / dt for n, dt, _ in res)
            if st >= warmup:
                per_step.append((rate, wall))
        if plan_flops:
            cache = os.path.join(ROOT, "profiles", "plan_flops.json")
            try:
                flops = json.load(open(cache)).get(workload)
            except Exception:
                flops = None
    value = float(np.mean([r for r, _ in per_step]))
    how = ("the UNMODIFIED reference: C++ TAT (oracle/_ref) + its own tetragono / tetraku / lazy (oracle/_ref/site), no tnsp_b200 code"
           if stock else ("reference C++ TAT under this repository's drivers (oracle/_ref/site missing)" if use_ref else "numpy port"))
    return {"value": value, "plan_flops": flops, "unit": "samples/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{samples_per_step} samples per chain x {cores} independent chains (one per host core, 1 BLAS thread each) per step, "
                      f"{steps} steps, workload {workload}; {how}",
            "ms_per_step": float(np.mean([w for _, w in per_step]) * 1e3)}


def _probe_ref(_):
    """does the reference extension load in a child process (libopenblas of the image present)?"""
    try:
        ref = os.path.join(ROOT, "oracle", "_ref")
        sys.path.insert(0, ref)
        import TAT  # noqa: F401
        return hasattr(TAT, "BoseU1")
    except Exception:
        return False


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        def loop():
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            while not self._stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a short collective timeout: a rank that leaves the lock-step fails the run within minutes, not after NCCL's 10
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    from tnsp_b200 import backend
    import tnsp_b200.TAT as TAT
    from tnsp_b200 import dist as tdist
    from tnsp_b200.tetragono import models
    from tnsp_b200.tetragono.observer import Observer
    from tnsp_b200.tetragono.sampling import ChainRng, SweepSampling
    B = backend.get()   # raises without the CUDA library / a GPU: no CPU fallback

    wl = WORKLOADS[args.workload]
    L1, L2, Dc, desc = wl["L1"], wl["L2"], wl["Dc"], wl["desc"]
    nb = args.chains or wl["chains"]
    if getattr(args, "strong", False):
        nb = max(1, nb // world)          # fixed total work: the curve whose limiter is the host path of a step
    sym_lat, hopping, points = build_workload(TAT, wl)
    sector = wl["sym"] != "No" and (args.engine == "sector" or wl["sym"].startswith("Fermi"))
    if wl["sym"] == "No":
        lat = sym_lat
        conf0 = models.neel_configuration(L1, L2)
    elif sector:
        # symmetric PEPS run on sector-compact tensors: per-chain symmetry sectors, the reference's own sector plan (TAT/ragged.py)
        from tnsp_b200.tetragono import dense_embedding
        lat = sym_lat
        conf0 = dense_embedding.embed_configuration(sym_lat, points)      # total physical indices [L1, L2, orbits]
    else:
        # symmetric PEPS enter the lock-step engine through the charge-dense embedding
        from tnsp_b200.tetragono import dense_embedding
        lat = dense_embedding.embed_lattice(sym_lat)
        conf0 = dense_embedding.embed_configuration(sym_lat, points)
        hopping = models.nearest_neighbour_terms(lat) if hopping is not None else None
    # ---- parity line: cache-cold amplitude and local energy of the start configuration, ALL through the engine being timed, against
    # the fixture the unmodified reference wrote for exactly this PEPS (tests/golden/make_golden.py cfg2size) -------------------------
    parity = None
    # (cfg2: tests/golden/make_golden.py cfg2size; the fermionic workloads: tests/golden/make_bench_fixtures.py, which also stores the
    # start configuration it evaluated)
    fixture = {"cfg2": "j1j2U1_6x6_d2_Dc36", "cfg2s": "j1j2U1_4x4_d1_Dc9", "cfg3s": "bench_cfg3s", "cfg4s": "bench_cfg4s", "cfg3": "bench_cfg3",
               "cfg4": "bench_cfg4"}.get(args.workload)
    fpath = os.path.join(ROOT, "tests", "golden", f"{fixture}.npz") if fixture else None
    # (EVERY rank runs it: the Observer's exit is a collective -- summed numerator and denominator leave the ratio unchanged)
    if fpath and os.path.exists(fpath) and (args.workload != "cfg2s") and sector:
        from tnsp_b200.tetragono.configuration import Configuration
        z = np.load(fpath)
        if "conf" in z.files and not np.array_equal(z["conf"], conf0):
            raise RuntimeError("bench parity check: the fixture was written for another start configuration")
        n_par = min(nb, 148)
        pc = Configuration(lat, Dc, n_par)
        pc.import_configuration(np.broadcast_to(conf0, (n_par,) + conf0.shape))
        ws_p = np.asarray(pc.hole(()).storage).reshape(-1)
        po = Observer(lat, enable_energy=True)
        with po:
            po(ws_p**2, pc)
        e_p = po._whole_result_reweight["energy"] / po._total_weight
        parity = {"fixture": f"tests/golden/{fixture}.npz (unmodified reference, cache-cold start configuration)", "chains": n_par,
                  "ws_reference": float(z["ws"][0]), "ws_max_rel_err": float(np.abs(ws_p - z["ws"][0]).max() / abs(z["ws"][0])),
                  "local_energy_reference": float(z["energy_s"][0]),
                  "local_energy_rel_err": float(abs(e_p - z["energy_s"][0]) / abs(z["energy_s"][0])), "tolerance": 1e-10}
        parity["ok"] = bool(parity["ws_max_rel_err"] <= 1e-10 and parity["local_energy_rel_err"] <= 1e-10)
        if not parity["ok"]:
            raise RuntimeError(f"bench parity check failed: {parity}")
        del pc, po
    # one normalisation pass so that amplitudes are O(1) (observer.normalize_lattice)
    s0 = SweepSampling(lat, Dc, None, hopping, nb=1, engine="sector" if sector else None)
    s0.configuration.import_configuration(conf0)
    TAT.random.seed(2333)
    o0 = Observer(lat, enable_energy=True)
    with o0:
        for _ in range(2):
            p, c = s0()
            o0(p, c)
    o0.normalize_lattice()
    del s0, o0
    if sector:
        # the single normalisation chain must not define the buffer capacities of the batch: learn again (calibration batch below, or
        # the first two warm-up steps)
        from tnsp_b200.TAT import ragged as _ragged
        _ragged._CAPS.clear()
        _ragged._LEARN.update(all=True, cycles=0)

    n_cal = min(148, max(32, nb // 8))
    if sector and wl["sym"].startswith("Fermi"):
        # few chains per GPU and strongly fluctuating sector sizes (hopping moves charge between the bonds): memory is not the limit
        # here, so batches up to 148 chains allocate the dense bound of every tensor (no learnt capacities, nothing can overflow)
        # (larger batches learn capacities; measured at 296 chains of cfg3: a factor of 3 needed three corrections up to 10, 89 GB -- start at 8)
        _ragged.CAP_FACTOR = 8.0
        _ragged.CAP_MIN_CHAINS = 160
    if sector and nb > n_cal and nb >= _ragged.CAP_MIN_CHAINS:
        # buffer capacities of the sector-compact engine are learnt on a small throw-away batch first (TAT/ragged.py)
        from tnsp_b200.tetragono.sampling import calibrate_sector_engine
        calibrate_sector_engine(lat, Dc, conf0, hopping, chains=n_cal, sweeps=2,
                                observer_options=dict(enable_energy=True, enable_gradient=True, enable_natural_gradient=wl["sr"]))
    rng = ChainRng(nb)
    rng.seed([(2333 + rank * nb + c) % 2**31 for c in range(nb)])
    rng.uniform_real(None)
    sampling = SweepSampling(lat, Dc, None, hopping, nb=nb, rng=rng)
    sampling.configuration.import_configuration(np.broadcast_to(conf0, (nb,) + conf0.shape))
    observer = Observer(lat, enable_energy=True, enable_gradient=True, enable_natural_gradient=wl["sr"])

    def step():
        """one optimisation step of the reference's gradient_descent loop (gradient.py:327-399) on this rank's
        chains: sweep + observe every chain, reduce, gradient (SR natural gradient by CG for cfg2)"""
        with observer:
            p, c = sampling()
            observer(p, c)
        return observer.natural_gradient_by_conjugate_gradient(wl["cg"], 0.0) if wl["sr"] else observer.gradient

    def sync_all():
        torch.cuda.synchronize()
        tdist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    # the long-lived objects (lattice, plan caches, lazy graphs: ~1e6 of them) leave the cyclic collector's generations, so that a full
    # collection inside a step only walks what the step itself created (one was seen to cost ~1 s of a 592-chain run)
    import gc
    gc.collect()
    gc.freeze()
    gc.set_threshold(50000, 20, 20)      # a step allocates ~5e5 container objects and almost no cycles: 10 young collections instead of 700
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    sync_all()
    l0 = B.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    launches = B.launch_count() - l0
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop() if rank == 0 else None
    energy = observer.energy
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    samples = nb * args.steps * world
    value = samples / (ms_max * 1e-3)

    # ---- end to end through the public API with HOST buffers --------------------------------------
    # every step: PEPS tensors host -> device (pinned), refresh the environments, sweep + observe,
    # gradient and energy device -> host
    host_sites = [[torch.from_numpy(np.ascontiguousarray(np.asarray(lat[l1, l2].storage))).pin_memory() for l2 in range(L2)] for l1 in range(L1)]
    h2d = sum(t_.numel() * 8 for row in host_sites for t_ in row)
    d2h = 0

    def e2e_step():
        nonlocal d2h
        for l1 in range(L1):
            for l2 in range(L2):
                lat[l1, l2]._data = host_sites[l1][l2].to("cuda", non_blocking=True).reshape(1, -1)
        sampling.configuration.refresh_all()
        grad = step()
        out = [g.data.cpu() for row in grad for g in row]
        e = observer.energy
        d2h = sum(o.numel() * 8 for o in out) + 16
        return e, out

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(1, args.steps // 2)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n_e2e):
        e2e_step()
    e1.record()
    sync_all()
    wall = time.perf_counter() - t0
    ms2 = max(e0.elapsed_time(e1), wall * 1e3)
    t = torch.tensor([ms2], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = nb * n_e2e * world / (float(t.item()) * 1e-3)

    # ---- roofline of the dominant kernel (instrumented pass, outside the timed regions) -----------
    # every rank runs the same steps (the observer's exchange is a collective); only rank 0 instruments its kernels
    roofline, breakdown, top_shapes = None, None, None
    n_prof = 3 if ms_max / args.steps < 2000 else 1
    prof = None
    step()
    torch.cuda.synchronize()
    rt_counts = None
    if rank == 0:
        from tnsp_b200 import profiling
        prof = profiling.KernelTimer(B)
        prof.enable()
        if sector:
            B.rt_stats(enable=1, reset=True)
    for _ in range(n_prof):
        step()
    torch.cuda.synchronize()
    if rank == 0:
        prof.disable()
        if sector:
            rt_counts = B.rt_stats(enable=0, read=True)
        breakdown = prof.summary()
        top_shapes = prof.shape_summary()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic_table = {}
        try:
            traffic_table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        roofline = profiling.roofline_of_dominant(breakdown, peaks, top_shapes, traffic_table)
        if sector:
            dgemm = profiling.measure_fp64_gemm_tflops()
            lines = profiling.sector_engine_rooflines(breakdown, rt_counts, peaks, dgemm)
            name = max(lines, key=lambda k: lines[k]["seconds"])
            top = lines[name]
            hbm_peak = peaks.get("hbm_gbs") or 6650.0
            if top["bound"] == "tensor":
                ach, peak, unit = top["algorithmic_tflops"], dgemm, "TFLOP/s"
            else:
                ach, peak, unit = top["gbs"], hbm_peak, "GB/s"
            roofline = {"kernel": name, "bound": top["bound"], "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "traffic": None,
                        "peak_source": ("MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6.65 TB/s (B200_PROFILING.md)")
                        if unit == "GB/s" else "cuBLAS DGEMM 4096^3 measured in this run",
                        "dgemm_peak_tflops": dgemm, "share_of_kernel_time": breakdown[name]["share"], "profiled_steps": n_prof,
                        "work_counted": "on the device over exactly the timed launches (tnsp_rt_stats): GEMM algorithmic flops = sum over the "
                                        "sector GEMMs of 2mnk (the reference plan, SURVEY 8d), executed = DMMA.8x8x4 issued x 512",
                        "classes": lines}

    out = None
    if rank == 0:
        out = {
            "metric": "VMC samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if getattr(args, "strong", False) else "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (randn_ PEPS seed 2333, Neel start, per-chain mt19937_64 seeds)",
            "config": {"workload": f"{args.workload}: {desc}", "chains_per_gpu": nb, "samples_per_step": nb * world,
                       "observer": "energy+gradient" + ("+SR natural gradient (CG %d)" % wl["cg"] if wl["sr"] else ""),
                       "symmetric_engine": None if wl["sym"] == "No" else ("sector-compact tensors: per-chain symmetry sectors planned on the device"
                                                                          if sector else "charge-dense embedding, sectors discovered on device"),
                       "l2": "working set of a step (all chains' environments) exceeds L2; no flush"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clock_info, "energy_per_site": energy[0], "parity_check": parity,
            "hbm_peak_allocated_gb": torch.cuda.max_memory_allocated() / 1e9,
            "roofline": roofline, "kernel_breakdown": breakdown, "top_shapes": top_shapes,
        }
    return out


def run_secondary(workloads=("cfg1", "cfg3s", "cfg4s"), steps=2, warmup=3):
    """short runs of the other BASELINE configurations (cfg1 at its full size; the fermionic cfg3 / cfg4 at their 4x4 smoke sizes --
    their full sizes are `--workload cfg3 | cfg4`, minutes per run), each in a fresh process: same metric, device-timed value and
    end-to-end value through host buffers.  Reported beside the headline line, never mixed into it."""
    out = {}
    for w in workloads:
        t0 = time.perf_counter()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", w, "--steps", str(steps), "--warmup", str(warmup),
                                "--no-cpu-baseline", "--no-secondary"], capture_output=True, text=True, timeout=600)
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out[w] = {"workload": d["config"]["workload"], "chains_per_gpu": d["config"]["chains_per_gpu"], "value": d["value"], "unit": d["unit"],
                      "e2e": d["e2e"]["value"], "ms_per_step": d["ms_per_step"], "gpu_launches": d["gpu_launches"],
                      "energy_per_site": d["energy_per_site"], "wall_s": time.perf_counter() - t0}
        except Exception as e:       # a report, never a reason to lose the headline line
            out[w] = {"error": repr(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0, help="Markov chains per GPU (lock-step batch); 0 = the workload's default")
    ap.add_argument("--ref-samples", type=int, default=8, help="samples per chain per step in the reference arm")
    ap.add_argument("--engine", default="sector", choices=["sector", "dense"],
                    help="lock-step engine of symmetric models: sector-compact tensors (default) or the charge-dense embedding of round 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the workload's chain count is the TOTAL over all ranks "
                                                          "(chains per GPU = total / world size) instead of the per-GPU count")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short runs of the other BASELINE configurations (`secondary` key)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        desc = WORKLOADS[args.workload]["desc"]
        r = run_reference(args.workload, args.steps, max(1, args.warmup // 3), args.ref_samples)
        line = {"impl": "reference", "metric": "VMC samples/sec", "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic (randn_ PEPS seed 2333, Neel start)",
                "config": {"workload": f"{args.workload}: {desc}", "observer": "energy+gradient" + ("+SR natural gradient" if WORKLOADS[args.workload]["sr"] else "")},
                "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    out = run_own(args)
    if rank == 0:
        if not args.no_cpu_baseline and int(os.environ.get("WORLD_SIZE", "1")) == 1:
            r = run_reference(args.workload, 2, 1, args.ref_samples, plan_flops=True)
            out["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            pf = r.get("plan_flops")
            if pf and "sector_plan" in pf and out.get("roofline") and out["roofline"].get("unit") == "TFLOP/s":
                # the roofline's `achieved` counts the flops the kernels execute (dense plan, SURVEY.md 8d quotes the dense
                # figure for this lattice); the block-sparse U(1) plan of the reference needs fewer: report that too
                rf = out["roofline"]
                ratio = pf["sector_plan"] / pf["dense_embedding"] if pf.get("dense_embedding") else None
                rf["plan_flops_per_sample"] = {"reference_sector_plan": pf["sector_plan"], "dense_embedding_executed": pf["dense_embedding"],
                                               "sector_over_dense": ratio,
                                               "note": "GEMM flops of one steady-state sample (sweep + observe) counted by the host planner on the "
                                                       "CPU checker; achieved x sector_over_dense = throughput in reference-plan flops"}
                if ratio:
                    rf["achieved_sector_plan"] = rf["achieved"] * ratio
                    rf["frac_sector_plan"] = rf["frac"] * ratio
        else:
            out["cpu_baseline"] = None
        if args.workload == "cfg2" and not args.no_secondary and int(os.environ.get("WORLD_SIZE", "1")) == 1:
            out["secondary"] = run_secondary()
        print(json.dumps(out))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Generates the golden fixtures of tests/golden/*.npz by running the UNMODIFIED reference
(PyTAT from oracle/_ref, tetragono/tetraku imported from /root/reference, single-rank mpi4py stub
from oracle/stubs) in THIS container.  The fixtures travel to the GPU box; the reference does not.

    python tests/golden/make_golden.py            # re-spawns itself with the reference environment

Each fixture holds a complete, self-describing model (edges, Hamiltonian terms, PEPS tensors) plus
the reference's results on it:
  ws            amplitude <s|psi> of a fixed configuration, cache-cold (single_layer_auxiliaries.py:710-716)
  energy_s      local energy E_s of that configuration                  (observer.py:313-398)
  holes         <psi|s|d_x psi>/<psi|s|psi> for every site              (lattice.py:362-416)
  traj_*        a sweep-sampling trajectory from a fixed seed           (sampling.py:117-154)
  energy/gradient/natural gradient accumulated over that trajectory     (observer.py:83-126,542-675)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def _respawn():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([
        os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "oracle", "stubs"), f"{REF}/tetragono", f"{REF}/tetraku",
        f"{REF}/lazy_graph", f"{REF}/PyScalapack"
    ])
    env["TNSP_GOLDEN_CHILD"] = "1"
    env["OPENBLAS_NUM_THREADS"] = "1"
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))


if os.environ.get("TNSP_GOLDEN_CHILD") != "1":
    _respawn()

import numpy as np  # noqa: E402
import TAT  # noqa: E402
import tetragono as tet  # noqa: E402

FIELDS = {"No": (), "BoseZ2": ("z2",), "BoseU1": ("u1",), "FermiU1": ("fermi",), "FermiZ2": ("parity",),
          "FermiU1BoseZ2": ("fermi", "z2"), "FermiU1BoseU1": ("fermi", "u1"), "FermiU1FermiU1": ("fermi_0", "fermi_1")}


def sym_tuple(sym_name, s):
    return [int(getattr(s, f)) for f in FIELDS[sym_name]]


def edge_desc(sym_name, e):
    return {"segments": [[sym_tuple(sym_name, s), int(d)] for s, d in e.segments], "arrow": bool(e.arrow)}


def tensor_desc(sym_name, t, arrays, key):
    arrays[key] = np.array(t.storage, dtype=np.float64)
    return {"names": [str(n) for n in t.names], "edges": [edge_desc(sym_name, t.edge_by_name(n)) for n in t.names], "storage": key}


def point_desc(sym_name, p):
    return [sym_tuple(sym_name, p[0]), int(p[1])]


def dump_case(name, sym_name, lattice, Dc, config_points, seed, n_samples, cg_step=2, hopping=None):
    arrays = {}
    meta = {"symmetry": sym_name, "L1": lattice.L1, "L2": lattice.L2, "Dc": Dc,
            "total_symmetry": sym_tuple(sym_name, lattice.total_symmetry)}
    meta["physics_edges"] = [[{str(o): edge_desc(sym_name, e) for o, e in lattice.physics_edges[l1, l2].items()}
                              for l2 in range(lattice.L2)] for l1 in range(lattice.L1)]
    meta["hamiltonians"] = []
    for i, (positions, h) in enumerate(lattice._hamiltonians.items()):   # insertion order = sweep tie order
        meta["hamiltonians"].append({"positions": [list(p) for p in positions], "tensor": tensor_desc(sym_name, h, arrays, f"ham_{i}")})
    meta["sites"] = [[tensor_desc(sym_name, lattice[l1, l2], arrays, f"site_{l1}_{l2}") for l2 in range(lattice.L2)]
                     for l1 in range(lattice.L1)]
    meta["config"] = [[{str(o): point_desc(sym_name, config_points[l1][l2][o]) for o in config_points[l1][l2]}
                       for l2 in range(lattice.L2)] for l1 in range(lattice.L1)]

    # --- single fresh configuration: ws, E_s, holes -------------------------------------------------
    conf = tet.sampling_lattice.Configuration(lattice, Dc)
    for l1 in range(lattice.L1):
        for l2 in range(lattice.L2):
            for o, p in config_points[l1][l2].items():
                conf[l1, l2, o] = p
    ws = conf.hole(())
    arrays["ws"] = np.array([float(ws)])
    meta["ws_names"] = [str(n) for n in ws.names]
    obs = tet.sampling_lattice.Observer(lattice, enable_energy=True, enable_gradient=True)
    with obs:
        obs(float(ws)**2, conf)
    arrays["energy_s"] = np.array([obs._whole_result_reweight["energy"] / obs._total_weight])
    holes = conf.holes()
    meta["holes"] = [[tensor_desc(sym_name, holes[l1][l2], arrays, f"hole_{l1}_{l2}") for l2 in range(lattice.L2)]
                     for l1 in range(lattice.L1)]

    # --- sweep trajectory from a fixed seed -----------------------------------------------------------
    TAT.random.seed(seed)
    sampling = tet.sampling_lattice.SweepSampling(lattice, Dc, None, hopping)
    meta["sweep_nearest_neighbour_only"] = hopping is not None
    for l1 in range(lattice.L1):
        for l2 in range(lattice.L2):
            for o, p in config_points[l1][l2].items():
                sampling.configuration[l1, l2, o] = p
    obs = tet.sampling_lattice.Observer(lattice, enable_energy=True, enable_gradient=True, enable_natural_gradient=True)
    traj, poss = [], []
    with obs:
        for _ in range(n_samples):
            p, c = sampling()
            traj.append(c.export_configuration())
            poss.append(p)
            obs(p, c)
    arrays["traj_config"] = np.array(traj)
    arrays["traj_possibility"] = np.array(poss)
    arrays["traj_energy"] = np.array(obs.total_energy)
    grad = obs.gradient
    meta["gradient"] = [[tensor_desc(sym_name, grad[l1][l2], arrays, f"grad_{l1}_{l2}") for l2 in range(lattice.L2)]
                        for l1 in range(lattice.L1)]
    ng = obs.natural_gradient_by_conjugate_gradient(cg_step, 0.0)
    meta["natural_gradient"] = [[tensor_desc(sym_name, ng[l1][l2], arrays, f"ng_{l1}_{l2}") for l2 in range(lattice.L2)]
                                for l1 in range(lattice.L1)]
    meta["seed"], meta["n_samples"], meta["cg_step"] = seed, n_samples, cg_step
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, **arrays)
    print(name, "ws", arrays["ws"], "E_s", arrays["energy_s"], "traj E", arrays["traj_energy"], os.path.getsize(out), "bytes")


def dump_driver_case(name, sym_name, lattice, Dc, config_points, seed, method="sweep", **kwargs):
    """the reference's own optimisation loop (`gradient_descent`, sampling_lattice/gradient.py:93-445) from a fixed seed:
    energy of every step and the PEPS tensors after the last update"""
    arrays = {}
    meta = {"symmetry": sym_name, "L1": lattice.L1, "L2": lattice.L2, "Dc": Dc,
            "total_symmetry": sym_tuple(sym_name, lattice.total_symmetry)}
    meta["physics_edges"] = [[{str(o): edge_desc(sym_name, e) for o, e in lattice.physics_edges[l1, l2].items()}
                              for l2 in range(lattice.L2)] for l1 in range(lattice.L1)]
    meta["hamiltonians"] = []
    for i, (positions, h) in enumerate(lattice._hamiltonians.items()):
        meta["hamiltonians"].append({"positions": [list(p) for p in positions], "tensor": tensor_desc(sym_name, h, arrays, f"ham_{i}")})
    meta["sites"] = [[tensor_desc(sym_name, lattice[l1, l2], arrays, f"site_{l1}_{l2}") for l2 in range(lattice.L2)]
                     for l1 in range(lattice.L1)]
    conf = tet.sampling_lattice.Configuration(lattice, Dc)
    for l1 in range(lattice.L1):
        for l2 in range(lattice.L2):
            for o, p in config_points[l1][l2].items():
                conf[l1, l2, o] = p
    start = conf.export_configuration()
    arrays["start_configuration"] = np.array(start)
    sampling_configurations = np.array(start)
    TAT.random.seed(seed)
    energies = []
    from tetragono.sampling_lattice.gradient import gradient_descent as ref_gradient_descent
    for whole, _ in ref_gradient_descent(lattice, sampling_method=method, configuration_cut_dimension=Dc,
                                                          sampling_configurations=sampling_configurations, **kwargs):
        energies.append(whole["energy"])
    arrays["step_energy"] = np.array(energies)
    arrays["last_configuration"] = np.array(sampling_configurations)
    meta["final_sites"] = [[tensor_desc(sym_name, lattice[l1, l2], arrays, f"final_{l1}_{l2}") for l2 in range(lattice.L2)]
                           for l1 in range(lattice.L1)]
    meta["seed"], meta["kwargs"], meta["method"] = seed, kwargs, method
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, **arrays)
    print(name, "step energies", arrays["step_energy"], os.path.getsize(out), "bytes")


def neel(lattice):
    S = lattice.Symmetry
    return [[{0: (S(), (l1 + l2) % 2)} for l2 in range(lattice.L2)] for l1 in range(lattice.L1)]


def heisenberg(L1, L2, D):
    from tetraku.models.heisenberg import abstract_lattice
    TAT.random.seed(2333)
    return tet.SamplingLattice(abstract_lattice(L1, L2, D, 1.0))


def heisenberg_u1(L1, L2, d):
    """spin-1/2 Heisenberg with U(1) (2 Sz) symmetric tensors; virtual bonds {-1,0,+1} x d (cfg2 family)"""
    T = TAT.BoseU1.D.Tensor
    state = tet.AbstractState(T, L1, L2)
    pe = [(+1, 1), (-1, 1)]
    state.physics_edges[...] = pe
    cpe = [(-1, 1), (+1, 1)]
    SS = T(["I0", "I1", "O0", "O1"], [cpe, cpe, pe, pe]).zero_()
    up, dn = (1, 0), (-1, 0)
    def setel(i0, i1, o0, o1, v):
        SS[{"I0": (-i0[0], 0), "I1": (-i1[0], 0), "O0": o0, "O1": o1}] = v
    setel(up, up, up, up, 0.25)
    setel(dn, dn, dn, dn, 0.25)
    setel(up, dn, up, dn, -0.25)
    setel(dn, up, dn, up, -0.25)
    setel(up, dn, dn, up, 0.5)
    setel(dn, up, up, dn, 0.5)
    H = -1.0 * SS
    state.hamiltonians["vertical_bond"] = H
    state.hamiltonians["horizontal_bond"] = H
    lat = tet.AbstractLattice(state)
    ve = [(-1, d), (0, d), (+1, d)]
    lat.virtual_bond["R"] = ve
    lat.virtual_bond["D"] = ve
    TAT.random.seed(2333)
    return tet.SamplingLattice(lat)


def j1j2_u1(L1, L2, d, J2):
    """cfg2 family: J1-J2 Heisenberg with U(1) tensors; diagonal J2 terms are measured by
    the Observer through the 2x2 replace branch but excluded from the sweep (sampling.py:156-190)"""
    lat = heisenberg_u1(L1, L2, d)
    H1 = lat._hamiltonians[((0, 0, 0), (0, 1, 0))]
    H2 = J2 * H1
    for l1 in range(L1 - 1):
        for l2 in range(L2 - 1):
            lat.hamiltonians[(l1, l2, 0), (l1 + 1, l2 + 1, 0)] = H2
            lat.hamiltonians[(l1, l2 + 1, 0), (l1 + 1, l2, 0)] = H2
    return lat


def neel_u1(lattice):
    S = lattice.Symmetry
    return [[{0: (S(+1) if (l1 + l2) % 2 == 0 else S(-1), 0)} for l2 in range(lattice.L2)] for l1 in range(lattice.L1)]


def tJ(L1, L2, D, T):
    from tetraku.models.tJ import abstract_lattice
    TAT.random.seed(2333)
    return tet.SamplingLattice(abstract_lattice(L1, L2, D, T, 1.0, 0.4))


def hubbard_ff(L1, L2, D, T):
    """FermiU1FermiU1 Hubbard with bonds following the charge-flow pattern of the shipped FermiU1
    lattice (tetraku/models/hubbard/__init__.py:47-75), one (up, down) pair per row."""
    from tetraku.models.hubbard.fermi_fermi import abstract_state
    state = tet.AbstractLattice(abstract_state(L1, L2, T, 1.0, 4.0))
    half = T // 2
    def segs(qu, qd):
        return [((qu + a, qd + b), D) for a in (-1, 0, 1) for b in (-1, 0, 1)]
    tt = half / L1
    for l1 in range(L1 - 1):
        Q = int(half * (L1 - l1 - 1) / L1)
        state.virtual_bond[l1, 0, "D"] = segs(Q, Q)
    for l1 in range(L1 - 1):
        for l2 in range(1, L2):
            state.virtual_bond[l1, l2, "D"] = [((0, 0), D)]
    for l1 in range(L1):
        for l2 in range(L2 - 1):
            Q = int(tt * (L2 - l2 - 1) / L2)
            state.virtual_bond[l1, l2, "R"] = segs(Q, Q)
    TAT.random.seed(2333)
    return tet.SamplingLattice(state)


def dump_structure(name, sym_name, lattice):
    """model only: symmetry, edges, Hamiltonian terms and site tensors (what golden_loader.build_lattice needs)"""
    arrays = {}
    L1, L2 = lattice.L1, lattice.L2
    meta = {"symmetry": sym_name, "L1": L1, "L2": L2, "total_symmetry": sym_tuple(sym_name, lattice.total_symmetry)}
    meta["physics_edges"] = [[{str(o): edge_desc(sym_name, e) for o, e in lattice.physics_edges[l1, l2].items()} for l2 in range(L2)] for l1 in range(L1)]
    meta["hamiltonians"] = [{"positions": [list(p) for p in positions], "tensor": tensor_desc(sym_name, h, arrays, f"ham_{i}")}
                            for i, (positions, h) in enumerate(lattice._hamiltonians.items())]
    meta["sites"] = [[tensor_desc(sym_name, lattice[l1, l2], arrays, f"site_{l1}_{l2}") for l2 in range(L2)] for l1 in range(L1)]
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **arrays)
    print(name, len(meta["hamiltonians"]), "terms")


def main():
    if "cfg2size" in sys.argv[1:]:
        # the benched configuration at its REAL size (BASELINE cfg2: 6x6 J1-J2 U(1), D = 2+2+2, Dc = 36): cache-cold ws / E_s /
        # holes of the Neel configuration and a 3-sample sweep trajectory with gradient + SR-CG, the at-size pin of bench.py
        lat = j1j2_u1(6, 6, 2, 0.5)
        hop = {k: v for k, v in lat._hamiltonians.items() if k[0][0] == k[1][0] or k[0][1] == k[1][1]}
        dump_case("j1j2U1_6x6_d2_Dc36", "BoseU1", lat, 36, neel_u1(lat), seed=18, n_samples=3, hopping=hop)
        return
    if "j1j2model" in sys.argv[1:]:
        # the J1-J2 model the reference ships (tetraku/models/J1J2/__init__.py, NoSymmetry)
        from tetraku.models.J1J2 import abstract_lattice
        TAT.random.seed(2333)
        dump_structure("model_j1j2_3x4_D2", "No", tet.SamplingLattice(abstract_lattice(3, 4, 2, 1.0, 0.5)))
        return
    if "direct" in sys.argv[1:]:
        # direct sampling (sampling.py:252-371): configurations and their probabilities from a fixed seed
        out = {}
        for name, lat, Dc, dl_cut, seed in (("heis_3x3_D2_Dc4", heisenberg(3, 3, 2), 4, 4, 31), ("heis_4x4_D3_Dc5_truncating", heisenberg(4, 4, 3), 5, 3, 32),
                                            ("heisU1_4x4_d1_Dc6", heisenberg_u1(4, 4, 1), 6, 4, 33), ("tJ_4x4_D1_Dc8", tJ(4, 4, 1, 2), 8, 4, 34),
                                            ("hubbardFF_4x4_D1_Dc8", hubbard_ff(4, 4, 1, 8), 8, 4, 35)):
            TAT.random.seed(seed)
            sampling = tet.sampling_lattice.DirectSampling(lat, Dc, None, dl_cut)
            confs, poss = [], []
            for _ in range(6):
                p, c = sampling()
                confs.append(c.export_configuration())
                poss.append(p)
            out[name + "_conf"] = np.array(confs)
            out[name + "_poss"] = np.array(poss)
            out[name + "_par"] = np.array([Dc, dl_cut, seed])
            print(name, poss)
        np.savez(os.path.join(ROOT, "tests", "golden", "direct_sampling.npz"), **out)
        return
    if "simple" in sys.argv[1:]:
        # simple update (simple_update_lattice.py:252-344) from the fixture PEPS, then conversion.py:24-46 to a sampling lattice:
        # bond environments (singular values), |site tensor| (the sign gauge of the svd is not pinned) and the exact amplitude of
        # the fixture configuration of the converted state (gauge invariant)
        out = {}
        cases = (("heis_3x3_D2_Dc4", heisenberg(3, 3, 2), "neel", 3, 0.05, 2), ("heis_4x4_D3_Dc5_truncating", heisenberg(4, 4, 3), "neel", 2, 0.1, 2),
                 ("heisU1_4x4_d1_Dc6", heisenberg_u1(4, 4, 1), "u1", 2, 0.05, 3), ("j1j2U1_4x4_d1_Dc9", j1j2_u1(4, 4, 1, 0.5), "u1", 1, 0.05, 3),
                 ("tJ_4x4_D1_Dc8", tJ(4, 4, 1, 2), None, 2, 0.05, 8), ("hubbardFF_4x4_D1_Dc8", hubbard_ff(4, 4, 1, 8), None, 2, 0.05, 6),
                 ("heis_3x3_D2_Dc4:relative", heisenberg(3, 3, 2), "neel", 2, 0.05, 0.3))
        gauge = np.load(os.path.join(ROOT, "tests", "golden", "gauge_fixing.npz"))
        for name, lat, points, steps, tau, dim in cases:
            fixture = name.split(":")[0]
            su = tet.conversion.sampling_lattice_to_simple_update_lattice(lat)
            su.update(steps, tau, dim)
            for l1 in range(su.L1):
                for l2 in range(su.L2):
                    out[f"{name}_site_{l1}_{l2}"] = np.abs(np.array(su[l1, l2].storage))
                    out[f"{name}_dims_{l1}_{l2}"] = np.array([su[l1, l2].edge_by_name(n).dimension for n in su[l1, l2].names])
                    for d in "RD":
                        env = su.environment[l1, l2, d]
                        if env is not None:
                            out[f"{name}_env_{l1}_{l2}_{d}"] = np.array(env.storage)
            back = tet.conversion.simple_update_lattice_to_sampling_lattice(su)
            conf = tet.sampling_lattice.Configuration(back, 256)
            if points is None:
                conf.import_configuration(gauge[fixture + "_conf"])
            else:
                pts = neel_u1(back) if points == "u1" else neel(back)
                for l1 in range(back.L1):
                    for l2 in range(back.L2):
                        for o, p in pts[l1][l2].items():
                            conf[l1, l2, o] = p
            out[name + "_par"] = np.array([steps, tau, dim])
            out[name + "_ws"] = np.array([float(conf.hole(()))])
            print(name, out[name + "_ws"], [int(x) for x in out[f"{name}_dims_1_1"]])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "simple_update.npz"), **out)
        return
    if "gauge" in sys.argv[1:]:
        # gauge fixing (SamplingLattice.expand_dimension(1.0, 0), lattice.py:821-919): the TRUNCATED amplitude of the fixture
        # configuration afterwards -- it depends on the gauge the reference fixes, not only on the state
        out = {}
        for name, lat, Dc, points in (("heis_3x3_D2_Dc4", heisenberg(3, 3, 2), 4, None), ("heis_4x4_D3_Dc5_truncating", heisenberg(4, 4, 3), 5, None),
                                      ("heisU1_4x4_d1_Dc6", heisenberg_u1(4, 4, 1), 6, "u1")):
            pts = neel_u1(lat) if points == "u1" else neel(lat)
            def amplitude(cut):
                conf = tet.sampling_lattice.Configuration(lat, cut)
                for l1 in range(lat.L1):
                    for l2 in range(lat.L2):
                        for o, p in pts[l1][l2].items():
                            conf[l1, l2, o] = p
                return float(conf.hole(()))
            before = amplitude(Dc)
            lat.expand_dimension(1.0, 0)
            out[name] = np.array([before, amplitude(Dc), amplitude(64)])
            print(name, out[name])
        for name, lat, Dc, seed in (("tJ_4x4_D1_Dc8", tJ(4, 4, 1, 2), 8, 41), ("hubbardFF_4x4_D1_Dc8", hubbard_ff(4, 4, 1, 8), 8, 42)):
            TAT.random.seed(seed)
            _, drawn = tet.sampling_lattice.DirectSampling(lat, Dc, None, 4)()
            conf_array = drawn.export_configuration()
            def amplitude(cut):
                conf = tet.sampling_lattice.Configuration(lat, cut)
                conf.import_configuration(conf_array)
                return float(conf.hole(()))
            before = amplitude(Dc)
            lat.expand_dimension(1.0, 0)
            out[name] = np.array([before, amplitude(Dc), amplitude(64)])
            out[name + "_conf"] = np.array(conf_array)
            print(name, out[name])
        np.savez(os.path.join(ROOT, "tests", "golden", "gauge_fixing.npz"), **out)
        return
    if "longrange" in sys.argv[1:]:
        # observables beyond the 2x2 replace window (ConfigurationPool.wss, lattice.py:562-614; Observer(cache_configuration=True)):
        # the first two-site Hamiltonian tensor placed on distant pairs of sites, measured on the first sweep samples after the fixture
        # configuration; Dc large enough that every contraction route is exact
        out = {}
        gauge = np.load(os.path.join(ROOT, "tests", "golden", "gauge_fixing.npz"))
        for name, lat, points, seed in (("heis_3x3_D2_Dc4", heisenberg(3, 3, 2), "neel", 31), ("heisU1_4x4_d1_Dc6", heisenberg_u1(4, 4, 1), "u1", 32),
                                        ("tJ_4x4_D1_Dc8", tJ(4, 4, 1, 2), None, 33)):
            L1, L2 = lat.L1, lat.L2
            term = [h for p, h in lat.hamiltonians if len(p) == 2][0]
            pairs = [((0, 0, 0), (L1 - 1, L2 - 1, 0)), ((0, 0, 0), (0, L2 - 1, 0)), ((0, 1, 0), (L1 - 1, 1, 0)), ((L1 - 1, 0, 0), (0, L2 - 1, 0)),
                     ((0, 0, 0), (0, 1, 0))]
            TAT.random.seed(seed)
            sampling = tet.sampling_lattice.SweepSampling(lat, 64, None, None)
            if points is None:
                sampling.configuration.import_configuration(gauge[name + "_conf"])
            else:
                pts = neel_u1(lat) if points == "u1" else neel(lat)
                for l1 in range(L1):
                    for l2 in range(L2):
                        for o, p in pts[l1][l2].items():
                            sampling.configuration[l1, l2, o] = p
            for mode in (True, "drop"):
                obs = tet.sampling_lattice.Observer(lat, cache_configuration=mode)
                obs.add_observer("far", {pair: term for pair in pairs})
                confs = []
                with obs:
                    for _ in range(3):
                        p, c = sampling()
                        confs.append(c.export_configuration())
                        obs(p, c)
                tag = name + ("_drop" if mode == "drop" else "")
                out[tag + "_conf"] = np.array(confs)
                out[tag + "_far"] = np.array([obs._result_reweight["far"][pair] / obs._total_weight for pair in pairs])
                print(tag, out[tag + "_far"])
            out[name + "_pairs"] = np.array(pairs)
            out[name + "_seed"] = np.array([seed])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "long_range.npz"), **out)
        return
    if "hamiltonian" in sys.argv[1:]:
        # Hamiltonian bookkeeping (abstract_state.py:200-240, utility.py:421-480) and lattice_dot (lattice.py:921-934) on a
        # fermionic model, where the renames / traces carry signs
        arrays, meta = {}, {}
        lat = tJ(4, 4, 1, 2)
        pe = lat.physics_edges[0, 0, 0]
        T = lat.Tensor
        TAT.random.seed(77)
        three = T(["I0", "I1", "I2", "O0", "O1", "O2"], [pe.conjugate()] * 3 + [pe] * 3).randn_()
        four = T(["O0", "I0", "O1", "I1", "O2", "I2", "O3", "I3"], [pe, pe.conjugate()] * 4).randn_()
        meta["three"] = tensor_desc("FermiU1BoseU1", three, arrays, "three")
        meta["four"] = tensor_desc("FermiU1BoseU1", four, arrays, "four")
        meta["cases"] = []
        for kind, tensor, points in (("trace", "three", ((0, 0, 0), (0, 1, 0), (0, 0, 0))), ("trace", "three", ((0, 1, 0), (0, 1, 0), (0, 0, 0))),
                                     ("trace", "four", ((1, 1, 0), (0, 0, 0), (1, 1, 0), (0, 0, 0))), ("trace", "four", ((1, 1, 0), (1, 1, 0), (1, 1, 0), (0, 0, 0))),
                                     ("sort", "three", ((1, 0, 0), (0, 0, 0), (0, 1, 0))), ("sort", "four", ((3, 0, 0), (0, 2, 0), (0, 1, 0), (2, 2, 0)))):
            fn = tet.utility.trace_repeated if kind == "trace" else tet.utility.sort_points
            result, new_points = fn({"three": three, "four": four}[tensor], points)
            i = len(meta["cases"])
            meta["cases"].append({"kind": kind, "tensor": tensor, "points": [list(p) for p in points], "new_points": [list(p) for p in new_points],
                                  "result": tensor_desc("FermiU1BoseU1", result, arrays, f"result_{i}")})
        arrays["lattice_dot"] = np.array([lat.lattice_dot()])
        herm = three + three.conjugate().edge_rename({f"I{i}": f"O{i}" for i in range(3)} | {f"O{i}": f"I{i}" for i in range(3)})
        meta["hermitian"] = tensor_desc("FermiU1BoseU1", herm, arrays, "hermitian")
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hamiltonian_tools.npz"), **arrays)
        print(arrays["lattice_dot"], [c["result"]["names"] for c in meta["cases"]])
        return
    if "hubbard" in sys.argv[1:]:
        # structure of the Hubbard lattice the reference ships (tetraku/models/hubbard/__init__.py), FermiU1, 4x4, D = 1, T = 8
        from tetraku.models.hubbard import abstract_lattice
        TAT.random.seed(2333)
        lattice = tet.SamplingLattice(abstract_lattice(4, 4, 1, 8, 1.0, 4.0))
        arrays, sym_name = {}, "FermiU1"
        meta = {"symmetry": sym_name, "L1": 4, "L2": 4, "total_symmetry": sym_tuple(sym_name, lattice.total_symmetry)}
        meta["physics_edges"] = [[{str(o): edge_desc(sym_name, e) for o, e in lattice.physics_edges[l1, l2].items()} for l2 in range(4)] for l1 in range(4)]
        meta["hamiltonians"] = [{"positions": [list(p) for p in positions], "tensor": tensor_desc(sym_name, h, arrays, f"ham_{i}")}
                                for i, (positions, h) in enumerate(lattice._hamiltonians.items())]
        meta["sites"] = [[tensor_desc(sym_name, lattice[l1, l2], arrays, f"site_{l1}_{l2}") for l2 in range(4)] for l1 in range(4)]
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "model_hubbard_4x4_D1.npz"), **arrays)
        print("hubbard", len(meta["hamiltonians"]), "terms")
        # ... and a 2x2 one with 2 particles measured exactly (ergodic enumeration): energy and gradient.  Its physical edge has
        # a segment of dimension 2 (singly occupied: up / down), unlike the other fixtures
        TAT.random.seed(2333)
        lattice = tet.SamplingLattice(abstract_lattice(2, 2, 2, 2, 1.0, 4.0))
        arrays = {}
        meta = {"symmetry": sym_name, "L1": 2, "L2": 2, "total_symmetry": sym_tuple(sym_name, lattice.total_symmetry)}
        meta["physics_edges"] = [[{str(o): edge_desc(sym_name, e) for o, e in lattice.physics_edges[l1, l2].items()} for l2 in range(2)] for l1 in range(2)]
        meta["hamiltonians"] = [{"positions": [list(p) for p in positions], "tensor": tensor_desc(sym_name, h, arrays, f"ham_{i}")}
                                for i, (positions, h) in enumerate(lattice._hamiltonians.items())]
        meta["sites"] = [[tensor_desc(sym_name, lattice[l1, l2], arrays, f"site_{l1}_{l2}") for l2 in range(2)] for l1 in range(2)]
        sampling = tet.sampling_lattice.ErgodicSampling(lattice, 16, None)
        obs = tet.sampling_lattice.Observer(lattice, enable_energy=True, enable_gradient=True)
        count = 0
        with obs:
            for _ in range(sampling.total_step):
                p, c = sampling()
                obs(p, c)
                count += 1
        grad = obs.gradient
        meta["gradient"] = [[tensor_desc(sym_name, grad[l1][l2], arrays, f"grad_{l1}_{l2}") for l2 in range(2)] for l1 in range(2)]
        arrays["energy"] = np.array(obs.total_energy)
        arrays["count"] = np.array([count])
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "model_hubbard_2x2_D2.npz"), **arrays)
        print("2x2", arrays["energy"], count)
        return
    if "common" in sys.argv[1:]:
        # every real operator tensor of the reference's common_tensor modules, as its models take them (`.to(float)`)
        arrays, meta = {}, {}
        syms = {"No": "No", "Fermi": "FermiU1", "Parity": "FermiZ2", "Fermi_Hubbard": "FermiU1", "Parity_Hubbard": "FermiZ2", "FermiU1_Hubbard": "FermiU1BoseU1", "FermiFermi_Hubbard": "FermiU1FermiU1", "FermiU1_tJ": "FermiU1BoseU1"}
        for module, sym in syms.items():
            mod = getattr(tet.common_tensor, module)
            meta[module] = {"symmetry": sym, "tensors": {}}
            def visit(prefix, holder):
                for attr in sorted(vars(holder)):
                    value = getattr(holder, attr)
                    if attr.startswith("_"):
                        continue
                    if isinstance(value, mod.Tensor):
                        raw = np.array(value.storage)
                        if np.abs(raw.imag).max() == 0:
                            meta[module]["tensors"][prefix + attr] = tensor_desc(sym, value.to(float), arrays, f"{module}.{prefix}{attr}")
                    elif isinstance(value, type) and attr in ("Up", "Down"):
                        visit(attr + ".", value)
            visit("", mod)
            print(module, sorted(meta[module]["tensors"]))
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "common_tensor.npz"), **arrays)
        return
    if "sustate" in sys.argv[1:]:
        # a simple-update checkpoint exactly as the reference writes it: pickle of the SimpleUpdateLattice after an update
        import pickle
        lat = heisenberg_u1(4, 4, 1)
        su = tet.conversion.sampling_lattice_to_simple_update_lattice(lat)
        su.update(2, 0.05, 3)
        with open(os.path.join(ROOT, "tests", "golden", "state_su_heisU1_4x4_d1.pkl"), "wb") as f:
            pickle.dump(su, f)
        return
    if "state" in sys.argv[1:]:
        # checkpoints exactly as the reference writes them (utility.py:365-388): pickle of the SamplingLattice
        import pickle
        for name, lat in (("state_heisU1_4x4_d1", heisenberg_u1(4, 4, 1)), ("state_tJ_4x4_D1", tJ(4, 4, 1, 2)), ("state_heis_3x3_D2", heisenberg(3, 3, 2))):
            with open(os.path.join(ROOT, "tests", "golden", name + ".pkl"), "wb") as f:
                pickle.dump(lat, f)
            print(name, os.path.getsize(os.path.join(ROOT, "tests", "golden", name + ".pkl")), "bytes")
        return
    if "driver" in sys.argv[1:]:
        lat = heisenberg(3, 3, 2)
        dump_driver_case("driver_heis_3x3_D2_Dc4_sr_momentum", "No", lat, 4, neel(lat), seed=21, sampling_total_step=12, grad_total_step=3,
                         grad_step_size=0.002, use_natural_gradient=True, conjugate_gradient_method_step=2, momentum_parameter=0.5,
                         use_fix_relative_step_size=True)
        lat = heisenberg_u1(4, 4, 1)
        dump_driver_case("driver_heisU1_4x4_d1_Dc6_line_search", "BoseU1", lat, 6, neel_u1(lat), seed=22, sampling_total_step=4,
                         grad_total_step=2, grad_step_size=0.02, use_line_search=True)
        lat = heisenberg(3, 3, 2)
        dump_driver_case("driver_heis_3x3_D2_Dc4_direct", "No", lat, 4, neel(lat), seed=24, sampling_total_step=5, grad_total_step=2,
                         grad_step_size=0.01, method="direct", direct_sampling_cut_dimension=4)
        lat = heisenberg(3, 3, 2)
        dump_driver_case("driver_heis_3x3_D2_Dc4_plain", "No", lat, 4, neel(lat), seed=23, sampling_total_step=5, grad_total_step=3,
                         grad_step_size=0.01, momentum_parameter=0.3, orthogonalize_momentum=True)
        return
    if "j1j2" in sys.argv[1:]:
        lat = j1j2_u1(4, 4, 1, 0.5)
        hop = {k: v for k, v in lat._hamiltonians.items() if k[0][0] == k[1][0] or k[0][1] == k[1][1]}
        dump_case("j1j2U1_4x4_d1_Dc9", "BoseU1", lat, 9, neel_u1(lat), seed=17, n_samples=6, hopping=hop)
        return
    lat = heisenberg(3, 3, 2)
    dump_case("heis_3x3_D2_Dc4", "No", lat, 4, neel(lat), seed=11, n_samples=12)
    lat = heisenberg(4, 4, 4)
    dump_case("heis_4x4_D4_Dc16", "No", lat, 16, neel(lat), seed=12, n_samples=6)
    lat = heisenberg(4, 4, 3)
    dump_case("heis_4x4_D3_Dc5_truncating", "No", lat, 5, neel(lat), seed=13, n_samples=6)
    lat = heisenberg_u1(4, 4, 1)
    dump_case("heisU1_4x4_d1_Dc6", "BoseU1", lat, 6, neel_u1(lat), seed=14, n_samples=6)
    # t-J, 4x4, 4 up + 4 down... tetraku's T is the half particle number
    lat = tJ(4, 4, 1, 2)
    S = lat.Symmetry
    # physical edge of the t-J model: (0,0) hole, (1,+1) up, (1,-1) down; one up and one down in rows 0 and 2
    hole, up, dn = (S(0, 0), 0), (S(1, 1), 0), (S(1, -1), 0)
    rows = [[up, hole, dn, hole], [hole, hole, hole, hole], [dn, hole, up, hole], [hole, hole, hole, hole]]
    dump_case("tJ_4x4_D1_Dc8", "FermiU1BoseU1", lat, 8, [[{0: rows[l1][l2]} for l2 in range(4)] for l1 in range(4)], seed=15, n_samples=6)
    lat = hubbard_ff(4, 4, 1, 8)
    S = lat.Symmetry
    e, u, d_, ud = (S(0, 0), 0), (S(1, 0), 0), (S(0, 1), 0), (S(1, 1), 0)
    rows = [[u, e, d_, e], [e, u, e, d_], [d_, e, u, e], [e, d_, e, u]]
    dump_case("hubbardFF_4x4_D1_Dc8", "FermiU1FermiU1", lat, 8, [[{0: rows[l1][l2]} for l2 in range(4)] for l1 in range(4)], seed=16, n_samples=6)


if __name__ == "__main__":
    main()

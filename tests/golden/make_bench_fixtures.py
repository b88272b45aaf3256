"""Fixtures for the bench workloads beyond cfg2, written by the UNMODIFIED reference (its C++ TAT = oracle/_ref, its own tetragono /
tetraku = oracle/_ref/site, exactly what `bench.py --impl reference` runs):

    python tests/golden/make_bench_fixtures.py cfg3s cfg4s [cfg3 ...]

For every workload: the lattice `bench._stock_reference_lattice` builds (the model of the reference arm, PEPS from TAT.random.seed(2333)),
the bench's start configuration, and its cache-cold amplitude and local energy -> tests/golden/bench_<workload>.npz.  bench.py checks the
first evaluation of its own engine against them (`parity_check`), tests/test_bench_workloads.py does the same on the CPU checker."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ref = os.path.join(ROOT, "oracle", "_ref")
    sys.path[:0] = [ROOT, ref, os.path.join(ref, "site")]
    import torch  # noqa: F401  (before the reference extension, see oracle/ref.py)
    import TAT
    import tetragono as tet
    import bench
    for name in sys.argv[1:]:
        wl = dict(bench.WORKLOADS[name])
        wl.setdefault("J2", 0.0)
        t0 = time.time()
        lat, _ = bench._stock_reference_lattice(wl)
        if wl.get("model") in ("hubbard_ff", "tJ"):
            start = bench.fermionic_start(wl)
        else:
            start = np.array([[[(l1 + l2) % 2] for l2 in range(wl["L2"])] for l1 in range(wl["L1"])])
        conf = tet.sampling_lattice.Configuration(lat, wl["Dc"])
        for l1 in range(wl["L1"]):
            for l2 in range(wl["L2"]):
                conf[l1, l2, 0] = lat.physics_edges[l1, l2, 0].point_by_index(int(start[l1, l2, 0]))
        ws = float(conf.hole(()))
        obs = tet.sampling_lattice.Observer(lat, enable_energy=True)
        with obs:
            obs(ws**2, conf)
        e = obs.energy[0] * wl["L1"] * wl["L2"]          # one sample: the energy IS its local energy
        print(name, "ws", ws, "E_s", e, "seconds", round(time.time() - t0, 1), flush=True)
        np.savez(os.path.join(ROOT, "tests", "golden", f"bench_{name}.npz"), ws=np.array([ws]), energy_s=np.array([e]), conf=start)


if __name__ == "__main__":
    main()

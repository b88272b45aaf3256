"""The bench workloads beyond cfg2 against fixtures written by the UNMODIFIED reference for exactly the model `bench.py` builds
(tests/golden/make_bench_fixtures.py: the reference arm's own lattice builder, PEPS from seed 2333, the bench's start configuration):
cache-cold amplitude and local energy of a lock-step batch on the sector-compact engine, 1e-10.  The same numbers are the
`parity_check` of the bench line on the B200."""
import os
import sys

import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from tnsp_b200.tetragono.configuration import Configuration
from tnsp_b200.tetragono.observer import Observer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def bench_workload_cold_check(name, nb):
    import bench
    z = np.load(os.path.join(ROOT, "tests", "golden", f"bench_{name}.npz"))
    wl = bench.WORKLOADS[name]
    lat, _, _ = bench.build_workload(TAT, wl)
    conf = Configuration(lat, wl["Dc"], nb)
    conf.import_configuration(np.broadcast_to(z["conf"], (nb,) + z["conf"].shape) if nb > 1 else z["conf"])
    ws = np.asarray(conf.hole(()).storage).reshape(-1)
    assert ws.shape == (nb,)
    assert np.abs(ws - z["ws"][0]).max() <= 1e-10 * abs(z["ws"][0])
    obs = Observer(lat, enable_energy=True)
    with obs:
        obs(ws**2 if nb > 1 else float(ws[0])**2, conf)
    e = obs._whole_result_reweight["energy"] / obs._total_weight
    assert abs(e - z["energy_s"][0]) <= 1e-10 * abs(z["energy_s"][0])
    return float(np.abs(ws - z["ws"][0]).max() / abs(z["ws"][0])), float(abs(e - z["energy_s"][0]) / abs(z["energy_s"][0]))


@pytest.mark.parametrize("name", ["cfg3s", "cfg4s"])
def test_fermionic_bench_workloads_match_the_reference(name):
    bench_workload_cold_check(name, 3)      # lock-step batch on the sector-compact engine (fermionic signs per chain)
    bench_workload_cold_check(name, 1)      # one block-symmetric chain


def test_cfg3_at_full_size_matches_the_reference():
    """BASELINE cfg3 at its REAL size (8x8 Hubbard, FermiU1 x FermiU1, D = 8, Dc = 64): two lock-step chains on the sector-compact engine
    against the unmodified reference's cache-cold amplitude and local energy (observed on the CPU checker: 1e-14 both)"""
    bench_workload_cold_check("cfg3", 2)


@pytest.mark.skipif(not os.environ.get("TNSP_SLOW_TESTS"), reason="15 minutes on the CPU checker; set TNSP_SLOW_TESTS=1")
def test_cfg4_at_full_size_matches_the_reference():
    """BASELINE cfg4 at its REAL size (10x10 t-J, FermiU1 x BoseU1, D = 10, Dc = 100), as above.  Last run on the CPU checker (round 2):
    ws 3.5e-14, local energy 6.5e-14 relative to the reference, 911 s"""
    bench_workload_cold_check("cfg4", 2)

"""The Python seam: `tnsp_b200.TAT.install_as_TAT()` must host UNMODIFIED reference code.

* the reference's own PyTAT test-suite (130 tests, all scalar types and symmetries) runs green on this module;
* the unmodified reference tetragono + tetraku + lazy sweep / observe on this module and reproduce the fixture the reference
  produced on its own PyTAT (heis_3x3_D2_Dc4).
Both need /root/reference (this container only; skipped on the GPU box) and run in child processes (the module tree is renamed)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "PyTAT", "tests")), reason="the reference tree exists in the build container only")


@needs_reference
def test_reference_pytat_suite_runs_on_this_module():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_pytat_suite.py")], capture_output=True, text=True, cwd="/tmp", timeout=1500)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    assert r.returncode == 0, tail
    assert "130 passed" in tail, tail


@needs_reference
def test_unmodified_tetragono_runs_on_this_module_and_matches_the_fixture():
    from golden_loader import load
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "oracle", "stubs"), f"{REF}/tetragono", f"{REF}/tetraku", f"{REF}/lazy_graph",
                                         f"{REF}/PyScalapack"])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_tetragono_on_our_tat.py")], capture_output=True, text=True, env=env,
                       cwd="/tmp", timeout=1500)
    assert r.returncode == 0, r.stderr[-1500:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    got = json.loads(line[7:])
    meta, z = load("heis_3x3_D2_Dc4")
    assert abs(got["ws"] - z["ws"][0]) <= 1e-10 * abs(z["ws"][0])
    assert abs(got["energy_s"] - z["energy_s"][0]) <= 1e-10 * abs(z["energy_s"][0])
    assert np.array_equal(np.array(got["traj_config"]), z["traj_config"])
    assert np.allclose(got["traj_possibility"], z["traj_possibility"], rtol=1e-9, atol=0)
    assert np.allclose(got["traj_energy"], z["traj_energy"], rtol=1e-9, atol=0)
    gs = max(np.abs(z[meta["gradient"][l1][l2]["storage"]]).max() for l1 in range(3) for l2 in range(3))
    for l1 in range(3):
        for l2 in range(3):
            assert got["gradient_names"][l1][l2] == meta["gradient"][l1][l2]["names"]
            assert np.abs(np.array(got["gradient"][l1][l2]) - z[meta["gradient"][l1][l2]["storage"]]).max() <= 1e-9 * gs

"""CPU-side checks of the boundary and of the chain-parallel plumbing.

* the C-ABI library loads without a GPU and exports every symbol include/tnsp_b200.h declares
  (no compute call is made);
* the product refuses to run without its CUDA library / a GPU (no silent CPU fallback);
* world_size = 2 over gloo: two ranks, each owning half of the Markov chains, must reproduce the energy
  and the gradient of one rank owning all chains (the exchange of Observer.__exit__ and of the SR-CG,
  reference observer.py:83-126, 639, 664).
"""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tnsp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnsp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tnsp_b200 import backend
    if not os.path.exists(backend.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tnsp_b200", "csrc")], check=True, capture_output=True)
    lib = ctypes.CDLL(backend.LIB_PATH)          # loads on a machine without a CUDA driver (cudart is lazy)
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/tnsp_b200.h but not exported"
    lib.tnsp_abi_version.restype = ctypes.c_int
    assert lib.tnsp_abi_version() >= 1
    host = ctypes.CDLL(backend.HOST_LIB_PATH)
    for name in names:
        if name.endswith("_host") or name in ("tnsp_abi_version", "tnsp_last_error", "tnsp_launch_count"):
            assert hasattr(host, name)


def test_no_cpu_fallback():
    """without CUDA the product backend must raise, never compute on the CPU"""
    import torch
    from tnsp_b200 import backend
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(backend.TnspError):
        backend.CudaBackend()


_WORKER = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
if world > 1:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
from oracle import numpy_backend
numpy_backend.install()
import tnsp_b200.TAT as TAT
from tnsp_b200.tetragono import models
from tnsp_b200.tetragono.observer import Observer
from tnsp_b200.tetragono.sampling import ChainRng, SweepSampling
from tnsp_b200.tetragono.state import SamplingLattice
L1, L2, D, Dc, total = 3, 3, 2, 4, 4
TAT.random.seed(2333)
lat = SamplingLattice(models.j1j2_abstract_lattice(TAT.No.D.Tensor, L1, L2, D, 1.0, 0.0))
nb = total // world
rng = ChainRng(nb)
rng.seed([100 + rank * nb + c for c in range(nb)])
s = SweepSampling(lat, Dc, nb=nb, rng=rng)
conf0 = models.neel_configuration(L1, L2)
s.configuration.import_configuration(np.broadcast_to(conf0, (nb,) + conf0.shape))
obs = Observer(lat, enable_energy=True, enable_gradient=True, enable_natural_gradient=True)
with obs:
    for _ in range(3):
        p, c = s()
        obs(p, c)
g = obs.gradient
saved = list(obs._Deltas)
ng = obs.natural_gradient_by_conjugate_gradient(3, 0.0)
obs._Deltas = saved                     # the pseudo-inverse SR of the same sample set (rows all-gathered over the ranks)
pg = obs.natural_gradient_by_direct_pseudo_inverse(1e-6, 0.0, [])
out = dict(energy=list(obs.total_energy), count=obs._count,
           grad=[np.asarray(t.storage).tolist() for row in g for t in row],
           ngrad=[np.asarray(t.storage).tolist() for row in ng for t in row],
           pgrad=[np.asarray(t.storage).tolist() for row in pg for t in row])
original = float(np.asarray(lat[0, 0].storage).sum())
lat[0, 0] = lat[0, 0] * float(rank + 2)            # the ranks drift apart ...
lat.bcast_lattice(root=world - 1)                   # ... and take the last rank's tensors (lattice.py:950-954)
out["bcast_ratio"] = float(np.asarray(lat[0, 0].storage).sum()) / original
if rank == 0:
    json.dump(out, open({out!r}, "w"))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
"""


def _run(world, port, out):
    code = _WORKER.format(root=ROOT, port=port, out=out)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        o, _ = p.communicate(timeout=600)
        assert p.returncode == 0, o.decode()[-2000:]


def test_two_ranks_over_gloo_equal_one_rank(tmp_path):
    import json
    f1, f2 = str(tmp_path / "w1.json"), str(tmp_path / "w2.json")
    _run(1, 29611, f1)
    _run(2, 29613, f2)
    a, b = json.load(open(f1)), json.load(open(f2))
    assert a["count"] == b["count"] == 12
    assert abs(a["bcast_ratio"] - 2.0) < 1e-12 and abs(b["bcast_ratio"] - 3.0) < 1e-12
    assert np.allclose(a["energy"], b["energy"], rtol=1e-12, atol=0)
    for key, tol in (("grad", 1e-11), ("ngrad", 1e-9), ("pgrad", 1e-7)):
        scale = max(np.abs(np.array(x)).max() for x in a[key])
        for x, y in zip(a[key], b[key]):
            assert np.abs(np.array(x) - np.array(y)).max() <= tol * scale

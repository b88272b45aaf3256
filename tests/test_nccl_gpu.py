"""-m gpu, needs TWO GPUs (skipped on a one-GPU box): two NCCL ranks on the CUDA path give the results of one rank holding all
chains -- energies, gradient, SR natural gradient (CG and pseudo-inverse), parameter broadcast.  The host-side twin over gloo on the
CPU checker is tests/test_abi_and_dist.py::test_two_ranks_over_gloo_equal_one_rank."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
import tnsp_b200.TAT as TAT
from tnsp_b200 import backend
backend.get()                                        # the CUDA library, or an error: no fallback
from tnsp_b200.tetragono import models
from tnsp_b200.tetragono.observer import Observer
from tnsp_b200.tetragono.sampling import ChainRng, SweepSampling
from tnsp_b200.tetragono.state import SamplingLattice
L1, L2, D, Dc, total = 3, 3, 2, 4, 8
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
obs._Deltas = saved
pg = obs.natural_gradient_by_direct_pseudo_inverse(1e-6, 0.0, [])
out = dict(energy=list(obs.total_energy), count=obs._count,
           grad=[np.asarray(t.storage).tolist() for row in g for t in row],
           ngrad=[np.asarray(t.storage).tolist() for row in ng for t in row],
           pgrad=[np.asarray(t.storage).tolist() for row in pg for t in row])
original = float(np.asarray(lat[0, 0].storage).sum())
lat[0, 0] = lat[0, 0] * float(rank + 2)
lat.bcast_lattice(root=world - 1)
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
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        o, _ = p.communicate(timeout=900)
        assert p.returncode == 0, o.decode()[-2000:]


def test_two_nccl_ranks_equal_one_rank(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    f1, f2 = str(tmp_path / "w1.json"), str(tmp_path / "w2.json")
    _run(1, 29621, f1)
    _run(2, 29623, f2)
    a, b = json.load(open(f1)), json.load(open(f2))
    assert a["count"] == b["count"] == 24
    assert abs(a["bcast_ratio"] - 2.0) < 1e-12 and abs(b["bcast_ratio"] - 3.0) < 1e-12
    assert np.allclose(a["energy"], b["energy"], rtol=1e-12, atol=0)
    for key, tol in (("grad", 1e-11), ("ngrad", 1e-9), ("pgrad", 1e-7)):
        scale = max(np.abs(np.array(x)).max() for x in a[key])
        for x, y in zip(a[key], b[key]):
            assert np.abs(np.array(x) - np.array(y)).max() <= tol * scale

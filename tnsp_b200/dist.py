"""Chain-parallel plumbing: one process per GPU, statistics combined by ONE exchange step.

Replaces the reference's mpi4py helpers (tetragono/tetragono/utility.py:53-102: Allreduce /
Iallreduce / Bcast on numpy views of tensor storage) by ``torch.distributed`` collectives --
NCCL over NVLink on the GPUs, gloo in the CPU tests.  Markov chains are independent, so there is
no data-path collective inside a sample; the exchange points are
  * Observer.__exit__   : packed scalars + one flat Delta||EDelta buffer   (observer.py:83-126)
  * SR conjugate gradient: one Np-vector + one scalar per iteration        (observer.py:639,664)
"""
from __future__ import annotations

import numpy as np
import torch


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def world_size():
    d = _dist()
    return d.get_world_size() if d else 1


def rank():
    d = _dist()
    return d.get_rank() if d else 0


def _device_for_collective():
    d = _dist()
    if d and d.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_host(array):
    """sum a host float64 vector over ranks (packed observer scalars)"""
    d = _dist()
    if not d:
        return array
    t = torch.from_numpy(np.ascontiguousarray(array)).to(_device_for_collective())
    d.all_reduce(t)
    return t.cpu().numpy()


def allreduce_number(x):
    return float(allreduce_host(np.array([x], dtype=np.float64))[0])


def allreduce_device(t):
    """sum a device vector over ranks in place (CG vectors)"""
    d = _dist()
    if not d:
        return t
    if t.device.type == "cuda" or d.get_backend() != "nccl":
        d.all_reduce(t)
        return t
    u = t.cuda()
    d.all_reduce(u)
    return u.to(t.device)


def allreduce_tensors(tensors):
    """sum the storages of many TAT tensors over ranks with one collective on one flat buffer"""
    d = _dist()
    if not d:
        return
    flat = torch.cat([t.data.reshape(-1) for t in tensors])
    flat = allreduce_device(flat)
    index = 0
    for t in tensors:
        n = t.data.numel()
        t._data = flat[index:index + n].reshape(t.data.shape).clone()
        index += n


def broadcast_tensors(tensors, root=0):
    d = _dist()
    if not d:
        return
    flat = torch.cat([t.data.reshape(-1) for t in tensors])
    d.broadcast(flat, src=root)
    index = 0
    for t in tensors:
        n = t.data.numel()
        t._data = flat[index:index + n].reshape(t.data.shape).clone()
        index += n


def barrier():
    d = _dist()
    if d:
        d.barrier()


def allgather_rows(t):
    """concatenate a [rows, n] tensor over ranks along the rows (every rank may hold a different number of rows)"""
    d = _dist()
    if not d:
        return t
    dev = t.device
    if d.get_backend() == "nccl" and dev.type != "cuda":
        t = t.cuda()
    counts = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(d.get_world_size())]
    d.all_gather(counts, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    most = int(max(int(c.item()) for c in counts))
    padded = torch.zeros((most, t.shape[1]), dtype=t.dtype, device=t.device)
    padded[:t.shape[0]] = t
    parts = [torch.empty_like(padded) for _ in range(d.get_world_size())]
    d.all_gather(parts, padded)
    return torch.cat([p_[:int(c.item())] for p_, c in zip(parts, counts)], dim=0).to(dev)

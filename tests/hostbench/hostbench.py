"""Host-overhead benchmark of the lock-step engine WITHOUT a GPU (development tool, not a product or test path).

The Python drivers + TAT/ragged.py + backend.py run exactly as on the B200, but the C-ABI is a library of no-ops
(tests/hostbench/stub.c, generated from the exported symbols) and the buffers are CPU tensors holding garbage: nothing is
computed, only the interpreter time per lock-step step is measured.  On the B200 a cfg2 step is host-bound (the same ~1.3 s at
148 and at 2368 chains), so this number IS the step time there.

    python tests/hostbench/hostbench.py [workload] [steps] [--profile]
"""
import ctypes
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # tests/hostbench -> repository root
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
np.seterr(all="ignore")
torch.set_num_threads(1)

from tnsp_b200 import backend  # noqa: E402


def build_stub():
    import subprocess
    so = "/tmp/libtnsp_stub.so"
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-w", "-o", so, os.path.join(ROOT, "tests", "hostbench", "stub.c"),
                           os.path.join(ROOT, "tnsp_b200", "lib", "capi.o"), "-lstdc++"])
    return so


class _Zeros:
    """stands in for the allocation prototypes: zero-filled buffers, so that nothing the host reads back is garbage"""

    def __init__(self, dtype):
        self.dtype = dtype

    def new_empty(self, shape):
        return torch.zeros(shape, dtype=self.dtype)


class DryBackend(backend.CudaBackend):
    name = "dry"

    def __init__(self):
        self.lib = lib = backend._declare_host(ctypes.CDLL(build_stub()))
        self.device = torch.device("cpu")
        self._dev_index = 0
        self._pf64 = torch.zeros(0, dtype=torch.float64)      # float buffers are never read by the host here (rt_scalar is overridden)
        self._pi32 = _Zeros(torch.int32)
        self._raw_stream = lambda idx: 0
        self.gather_gemm = True
        self.sector_discovery = False
        self._declare_kernels()
        self._rt_declare()

    def from_numpy(self, array):
        return torch.from_numpy(np.ascontiguousarray(array)).clone()

    upload = from_numpy

    def to_numpy(self, t):
        return t.detach().numpy()

    def synchronize(self):
        pass

    def rt_scalar(self, data, match):
        super().rt_scalar(data, match)
        return torch.full((max(data.shape[0], match.shape[0]),), 0.7, dtype=torch.float64)      # deterministic amplitudes: the same accept / reject path in every run

    def rt_overflow(self, clear=True):
        return 0

    def launch_count(self):
        self.lib.tnsp_stub_calls.restype = ctypes.c_longlong
        return int(self.lib.tnsp_stub_calls())


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 2
    # the model (PEPS, Hamiltonian element tables) is built with real arithmetic on the numpy checker; only the lock-step engine runs dry
    from oracle.numpy_backend import NumpyBackend
    backend.set_backend(NumpyBackend())
    import bench
    import tnsp_b200.TAT as TAT
    from tnsp_b200.TAT import ragged
    from tnsp_b200.tetragono import dense_embedding
    from tnsp_b200.tetragono.observer import Observer
    from tnsp_b200.tetragono.sampling import ChainRng, SweepSampling
    from tnsp_b200.tetragono.tensor_element import element_table
    ragged.CAPS_ENABLED = False
    wl = bench.WORKLOADS[workload]
    lat, hopping, points = bench.build_workload(TAT, wl)
    for terms in (lat._hamiltonians, hopping or {}):
        for positions, h in terms.items():
            element_table(h, [lat.physics_edges[p] for p in positions])
    B = DryBackend()
    backend.set_backend(B)
    backend._host_lib = B.lib
    conf0 = dense_embedding.embed_configuration(lat, points)
    nb = 2
    rng = ChainRng(nb)
    rng.seed([2333 + c for c in range(nb)])
    s = SweepSampling(lat, wl["Dc"], None, hopping, nb=nb, rng=rng, engine="sector")
    s.configuration.import_configuration(np.broadcast_to(conf0, (nb,) + conf0.shape))
    obs = Observer(lat, enable_energy=True, enable_gradient=True, enable_natural_gradient=wl["sr"])

    def step():
        with obs:
            p, c = s()
            obs(p, c)
        return obs.natural_gradient_by_conjugate_gradient(wl["cg"], 0.0) if wl["sr"] else obs.gradient

    step()
    step()
    l0 = B.launch_count()
    if "--profile" in sys.argv:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        step()
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(45)
        return
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    n = (B.launch_count() - l0) / steps
    dt = min(times)
    print({k: v for k, v in ragged.STATS.items()})
    print(f"{workload}: min {dt * 1e3:.1f} ms / median {sorted(times)[len(times) // 2] * 1e3:.1f} ms per step on the host, {n:.0f} C-ABI calls per step, "
          f"{dt / n * 1e6:.1f} us per call")


if __name__ == "__main__":
    main()

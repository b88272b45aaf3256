"""cfg5 microbenchmark (BASELINE.json configs[4]): block-symmetric contract / qr / svd of tnsp_b200.TAT on random U(1) rank-5
tensors, one call for a whole batch of samples, CUDA-event timed; the unmodified reference PyTAT (oracle/_ref, one sample per
call on one host core) is timed beside it when present.
    python scripts/mb_cfg5.py [Dc nb [Dc nb ...]]         default: 64 4096  128 1024"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnsp_b200 import backend, cfg5
import tnsp_b200.TAT as TAT


def timeit(fn, budget_s=6.0, max_reps=5):
    fn(); torch.cuda.synchronize()
    times = []
    t_begin = time.perf_counter()
    while len(times) < max_reps and time.perf_counter() - t_begin < budget_s:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return min(times)


def main():
    B = backend.get()
    args = [int(x) for x in sys.argv[1:]] or [64, 4096, 128, 1024]
    ref = None
    try:
        from oracle.ref import load_reference_tat
        ref = load_reference_tat()
    except Exception:
        pass
    for Dc, nb in zip(args[0::2], args[1::2]):
        a1, a2, v1, v2 = cfg5.random_batch(TAT, Dc, nb)
        T = a1.contract(a2, {("D", "U")})
        size = T.storage.size if nb == 1 else np.asarray(T.data.shape)[-1]
        print(f"Dc={Dc} nb={nb}: T has {int(size)} elements per sample ({int(size) * 8 * nb / 1e9:.2f} GB per batch)", flush=True)
        l0 = B.launch_count()
        ms_c = timeit(lambda: a1.contract(a2, {("D", "U")}))
        ms_q = timeit(lambda: T.qr("r", {"R1", "R2"}, "R", "L"))
        ms_s = timeit(lambda: T.svd({"L1", "L2"}, "R", "L", "L", "R", Dc))
        launches = B.launch_count() - l0
        by = int(size) * 8 * nb
        print(f"  contract {ms_c:9.3f} ms  ({by / ms_c / 1e6:7.0f} GB/s of result bytes)   {nb / ms_c * 1e3:10.0f} samples/s")
        print(f"  qr       {ms_q:9.3f} ms  ({2 * by / ms_q / 1e6:7.0f} GB/s of in+out bytes)   {nb / ms_q * 1e3:10.0f} samples/s")
        print(f"  svd      {ms_s:9.3f} ms  ({by / ms_s / 1e6:7.0f} GB/s of input bytes)    {nb / ms_s * 1e3:10.0f} samples/s   [{launches} launches in the timing loops]", flush=True)
        if ref is not None:
            (n1, e1), (n2, e2) = cfg5.structures(ref, Dc)
            r1 = ref.BoseU1.D.Tensor(n1, e1); r1.storage = v1[0]
            r2 = ref.BoseU1.D.Tensor(n2, e2); r2.storage = v2[0]
            rT = r1.contract(r2, {("D", "U")})
            out = []
            for fn in (lambda: r1.contract(r2, {("D", "U")}), lambda: rT.qr("r", {"R1", "R2"}, "R", "L"),
                       lambda: rT.svd({"L1", "L2"}, "R", "L", "L", "R", Dc)):
                t0 = time.perf_counter(); fn(); out.append(time.perf_counter() - t0)
            print(f"  reference PyTAT, 1 sample on 1 host core: contract {out[0] * 1e3:.2f} ms, qr {out[1] * 1e3:.2f} ms, svd {out[2] * 1e3:.2f} ms"
                  f"  -> per core {1 / out[0]:.0f} / {1 / out[1]:.0f} / {1 / out[2]:.1f} samples/s", flush=True)
        del a1, a2, T
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

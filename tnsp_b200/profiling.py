"""Per-kernel-class device timing with CUDA events on the launching stream, plus the algorithmic
work each launch performs (bytes / flops as defined in DESIGN.md), for bench.py's roofline object.
Used only in an instrumented pass OUTSIDE the timed regions (the event pairs add launch overhead)."""
from __future__ import annotations

import numpy as np
import torch


def _nb(t):
    return t.shape[0]


class KernelTimer:
    CLASSES = ("pack", "gemm", "gemm_gather", "qr", "svd", "svd_cut", "norm", "scale", "binary", "unary", "gather_rows", "select", "grad_accumulate",
               "diag_scatter", "block_sign", "svd_mask",
               # sector-compact engine (TAT/ragged.py): algorithmic work of these classes is counted ON THE DEVICE (per-chain sector
               # sizes are never known to the host): backend.rt_stats
               "rt_sort", "rt_match", "rt_match_many", "rt_repack", "rt_repack_pair", "rt_gemm", "rt_dot", "rt_factor_plan", "rt_qr_work", "rt_svd_work", "rt_svd_finish", "rt_svd_scatter",
               "rt_scale", "rt_binary", "rt_norm", "rt_scalar")

    def __init__(self, backend):
        self.B = backend
        self.records = {k: [] for k in self.CLASSES}
        self.shapes = {}   # (class, shape signature) -> list of (e0, e1, flops, bytes)
        self._orig = {}

    def _work(self, name, args):
        """(algorithmic bytes, flops) of one launch"""
        if name == "pack":
            plan, src, dst = args[:3]
            return 16.0 * plan.total * _nb(dst), 0.0
        if name == "gemm":
            plan, a, b, c = args[:4]
            nb = _nb(c)
            g = plan.gemm
            flops = float((2 * g[:, 0] * g[:, 1] * g[:, 2]).sum()) * nb
            by = 8.0 * (float((g[:, 0] * g[:, 2]).sum()) * _nb(a) + float((g[:, 2] * g[:, 1]).sum()) * _nb(b) + float((g[:, 0] * g[:, 1]).sum()) * nb)
            return by, flops
        if name == "gemm_gather":
            plan, a, b, c = args[:4]
            _, _, m, n, k = plan.gather
            nb = _nb(c)
            return 8.0 * (m * k * _nb(a) + k * n * _nb(b) + m * n * nb), 2.0 * m * n * k * nb
        if name in ("qr", "svd"):
            plan, a = args[:2]
            s = plan.sectors
            m, n, k = s[:, 0], s[:, 1], s[:, 2]
            nb = _nb(a)
            if name == "qr":
                return 8.0 * float((2 * m * n + m * k + k * n).sum()) * nb, float((4 * m * n * k).sum()) * nb
            return 8.0 * float((m * n + m * k + k * n + k).sum()) * nb, 0.0
        if name in ("scale", "unary", "block_sign"):
            x = args[1] if name == "block_sign" else args[0]
            return 16.0 * x.shape[1] * _nb(x), 0.0
        if name in ("binary", "select"):
            x = args[1] if name == "select" else args[0]
            return 24.0 * x.shape[1] * max(_nb(x), 1), 0.0
        if name == "norm":
            return 8.0 * args[0].numel(), 0.0
        if name == "grad_accumulate":
            return 8.0 * args[0].numel(), 2.0 * args[0].numel()
        if name == "gather_rows":
            return 16.0 * args[1] * _nb(args[2]), 0.0
        return 0.0, 0.0

    def enable(self):
        for name in self.CLASSES:
            fn = getattr(self.B, name, None)
            if fn is None:
                continue
            self._orig[name] = fn

            def wrapped(*args, _fn=fn, _name=name, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = _fn(*args, **kw)
                e1.record()
                by, fl = self._work(_name, args)
                self.records[_name].append((e0, e1, by, fl))
                if _name == "rt_gemm":
                    sig = (_name, 0, (int(args[0].M), int(args[1].N), int(args[0].N)), int(args[4]))
                    self.shapes.setdefault(sig, []).append((e0, e1, fl, by))
                elif _name == "rt_sort":
                    M = 1
                    for _, _, d in args[0]:
                        M *= int(d)
                    sig = (_name, len(args[0]), (M, 0, 0), max([int(a.shape[0]) for a, _, _ in args[0]] + [1]))
                    self.shapes.setdefault(sig, []).append((e0, e1, fl, by))
                elif _name == "rt_repack":
                    dst = args[2]
                    dims = (int(dst.M), int(dst.N), 0) if hasattr(dst, "M") else (int(dst.shape[1]), 0, -1)
                    sig = (_name, 0, dims, int(dst.data.shape[0]) if hasattr(dst, "M") else int(dst.shape[0]))
                    self.shapes.setdefault(sig, []).append((e0, e1, fl, by))
                elif _name == "rt_repack_pair":
                    d0, d1 = args[2], args[6]
                    sig = (_name, 0, (int(d0.M * d0.N), int(d1.M * d1.N), 0), int(d0.data.shape[0]))
                    self.shapes.setdefault(sig, []).append((e0, e1, fl, by))
                elif _name in ("rt_qr_work", "rt_svd_work"):
                    sig = (_name, 0, (int(args[0].M), int(args[0].N), 0), int(args[-1]))
                    self.shapes.setdefault(sig, []).append((e0, e1, fl, by))
                elif _name in ("gemm", "gemm_gather", "qr", "svd"):
                    tab = args[0].gemm if _name == "gemm" else ([args[0].gather[2:5]] if _name == "gemm_gather" else args[0].sectors)
                    sig = (_name, len(tab), tuple(int(x) for x in tab[0][:3]), int(args[3 if _name.startswith('gemm') else 1].shape[0]))
                    self.shapes.setdefault(sig, []).append((e0, e1, fl, by))
                return r

            setattr(self.B, name, wrapped)

    def disable(self):
        for name, fn in self._orig.items():
            try:
                delattr(self.B, name)   # wrapped functions are instance attributes shadowing the methods
            except AttributeError:
                pass
        self._orig = {}

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            if not recs:
                continue
            ms = [a.elapsed_time(b) for a, b, _, _ in recs]
            out[name] = {"launches": len(recs), "ms": float(np.sum(ms)), "bytes": float(sum(r[2] for r in recs)),
                         "flops": float(sum(r[3] for r in recs))}
        total = sum(v["ms"] for v in out.values()) or 1.0
        for v in out.values():
            v["share"] = v["ms"] / total
        return out

    def shape_summary(self, top=40):
        """heaviest (kernel class, #descriptors, first (m, n, k), chains) signatures"""
        rows = []
        for sig, recs in self.shapes.items():
            ms = float(sum(a.elapsed_time(b) for a, b, _, _ in recs))
            rows.append({"kernel": sig[0], "descriptors": sig[1], "mnk": list(sig[2]), "chains": sig[3], "launches": len(recs), "ms": ms,
                         "tflops": float(sum(r[2] for r in recs)) / (ms * 1e9) if ms else 0.0,
                         "gbs": float(sum(r[3] for r in recs)) / (ms * 1e6) if ms else 0.0})
        rows.sort(key=lambda r: -r["ms"])
        return rows[:top]


def measure_fp64_gemm_tflops(n=4096, reps=3):
    """cuBLAS DGEMM throughput of this GPU (denominator of FP64 tensor-bound kernels; MEASURED_PEAKS.json has no FP64 entry)"""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n**3 / (best * 1e-3) / 1e12


def roofline_of_dominant(breakdown, peaks, top_shapes=None, traffic_table=None):
    """roofline object of bench.py for the kernel class with the largest share of device time.

    achieved = algorithmic bytes (flops) of all its launches in the instrumented pass / their CUDA-event time;
    traffic  = DRAM bytes per launch of the class' heaviest shape from the ncu --set full capture recorded in
               profiles/ncu_traffic.json (None if that shape was not captured), next to its algorithmic bytes."""
    if not breakdown:
        return None
    name = max(breakdown, key=lambda k: breakdown[k]["ms"])
    v = breakdown[name]
    sec = v["ms"] * 1e-3
    hbm_peak = peaks.get("hbm_gbs")
    which = "of measured (MEASURED_PEAKS.json hbm_gbs)"
    if hbm_peak is None:
        hbm_peak, which = 6650.0, "of fallback (B200_PROFILING.md 6.65 TB/s)"
    gbs = v["bytes"] / sec / 1e9
    intensity = v["flops"] / max(v["bytes"], 1.0)
    extra = {"launches": v["launches"], "avg_launch_us": v["ms"] * 1e3 / v["launches"], "share_of_kernel_time": v["share"]}
    traffic = None
    if top_shapes:
        mine = [r for r in top_shapes if r["kernel"] == name]
        if mine:
            r = mine[0]
            key = "%s:%s:%d" % (name, "x".join(str(x) for x in r["mnk"]), r["chains"])
            extra["heaviest_shape"] = {"mnk": r["mnk"], "chains": r["chains"], "launches": r["launches"], "ms": r["ms"],
                                       "achieved_gbs": r["gbs"], "achieved_tflops": r["tflops"]}
            scale = 1.0
            if traffic_table and key not in traffic_table:
                # the ncu capture may have been taken at another chain count: DRAM bytes per launch are linear in the
                # number of chains (every chain streams its own operands), scale and say so
                prefix = key.rsplit(":", 1)[0] + ":"
                other = [k for k in traffic_table if k.startswith(prefix)]
                if other:
                    scale = r["chains"] / float(other[0].rsplit(":", 1)[1])
                    key = other[0]
            if traffic_table and key in traffic_table:
                t = traffic_table[key]
                traffic = t["dram_bytes_per_launch"] * scale
                alg = t.get("algorithmic_bytes_per_launch")
                extra["heaviest_shape"]["algorithmic_bytes_per_launch"] = alg * scale if alg is not None else None
                extra["heaviest_shape"]["ncu_report"] = t.get("report")
                if scale != 1.0:
                    extra["heaviest_shape"]["traffic_scaled_from"] = key
    if name in ("gemm", "gemm_gather") and intensity > 6.0:
        peak = measure_fp64_gemm_tflops()
        ach = v["flops"] / sec / 1e12
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (no FP64 entry in MEASURED_PEAKS.json)",
                "flop_count": "dense plan 2 m n k per launch; with zero-fragment skipping (symmetric models) fewer tensor-core "
                              "instructions are issued, so a shape's rate may exceed the pipe's peak", **extra}
    return {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": traffic,
            "peak_source": which, **extra}


def sector_engine_rooflines(breakdown, stats, peaks, dgemm_peak):
    """roofline lines of the sector-compact engine from THIS run's timers and device counters (algorithmic GEMM work
    = sum over the reference plan's sector GEMMs of 2 m n k; executed = DMMA.8x8x4 issued x 512):
    stats = backend.rt_stats counters accumulated over exactly the launches `breakdown` timed."""
    hbm = peaks.get("hbm_gbs") or 6650.0
    out = {}

    def sec(name):
        return breakdown[name]["ms"] * 1e-3 if name in breakdown else 0.0

    t = sec("rt_gemm")
    if t:
        alg, issued, by = float(stats[0]), float(stats[1]), float(stats[2])
        out["rt_gemm"] = {"bound": "hbm" if alg / max(by, 1.0) < 6.0 else "tensor", "seconds": t, "sector_gemms": int(stats[8]),
                          "algorithmic_flops": alg, "executed_flops": issued, "executed_over_algorithmic": issued / alg if alg else None,
                          "algorithmic_tflops": alg / t / 1e12, "frac_of_dgemm_peak": alg / t / 1e12 / dgemm_peak if dgemm_peak else None,
                          "executed_frac_of_dgemm_peak": issued / t / 1e12 / dgemm_peak if dgemm_peak else None,
                          "algorithmic_bytes": by, "gbs": by / t / 1e9, "frac_of_hbm_peak": by / t / 1e9 / hbm}
    t = sec("rt_repack") + sec("rt_repack_pair")
    if t:
        out["rt_repack"] = {"bound": "hbm", "seconds": t, "algorithmic_bytes": 16.0 * stats[3], "gbs": 16.0 * stats[3] / t / 1e9,
                            "frac_of_hbm_peak": 16.0 * stats[3] / t / 1e9 / hbm}
    t = sec("rt_qr_work")
    if t:
        out["rt_qr_work"] = {"bound": "hbm", "seconds": t, "algorithmic_bytes": float(stats[4]), "gbs": stats[4] / t / 1e9,
                             "frac_of_hbm_peak": stats[4] / t / 1e9 / hbm, "flops": float(stats[5]), "tflops": stats[5] / t / 1e12}
    t = sec("rt_svd_work")
    if t:
        out["rt_svd_work"] = {"bound": "hbm", "seconds": t, "algorithmic_bytes": float(stats[6]), "gbs": stats[6] / t / 1e9,
                              "frac_of_hbm_peak": stats[6] / t / 1e9 / hbm, "note": "bytes of one pass; the Jacobi sweeps run in shared memory"}
    return out

"""Thin ctypes layer over the C-ABI of ``libtnsp_b200.so`` (include/tnsp_b200.h).

torch tensors are used only as device buffers (allocation, stream, data_ptr); every numerical
operation of the hot path is a kernel of this repository's CUDA library.  There is NO CPU
fallback: ``get()`` raises if the CUDA library or a GPU is missing.

Tests of the host-side logic (planner, TAT API, VMC drivers) on machines without a GPU install a
checker backend explicitly with ``set_backend`` (see oracle/numpy_backend.py); the product never
selects it by itself.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtnsp_b200.so")
HOST_LIB_PATH = os.path.join(LIB_DIR, "libtnsp_host.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p

_backend = None
_host_lib = None


class TnspError(RuntimeError):
    pass


class RtForm(ctypes.Structure):
    """tnsp_rt_form of include/tnsp_b200.h: one sector-compact storage (or a dense array when rt is NULL)"""
    _fields_ = [("data", c_vp), ("data_stride", c_i64), ("rt", c_vp), ("rt_stride", c_i64), ("M", c_i64), ("ct", c_vp), ("ct_stride", c_i64),
                ("N", c_i64), ("match", c_vp), ("match_stride", c_i64)]


class RtMatchSpec(ctypes.Structure):
    """tnsp_rt_match_spec: sector pairing computed inside the consumer kernel"""
    _fields_ = [("rs", c_int), ("cs", c_int), ("t1", c_vp), ("t1_stride", c_int), ("s1", c_int), ("t2", c_vp), ("t2_stride", c_int), ("s2", c_int),
                ("match_out", c_vp), ("match_out_stride", c_i64), ("tsum_out", c_vp), ("cap", c_i64)]


class RtMatchJob(ctypes.Structure):
    """tnsp_rt_match_job: one pairing of tnsp_rt_match_multi_i32"""
    _fields_ = [("rt", c_vp), ("rt_stride", c_i64), ("rs", ctypes.c_int32), ("ct", c_vp), ("ct_stride", c_i64), ("cs", ctypes.c_int32),
                ("t1", c_vp), ("t1_stride", ctypes.c_int32), ("s1", ctypes.c_int32), ("t2", c_vp), ("t2_stride", ctypes.c_int32),
                ("s2", ctypes.c_int32), ("match", c_vp), ("tsum", c_vp), ("cap", c_i64)]


RT_SMAX = 64
RT_HDR = 3 + 2 * RT_SMAX
RT_MSTRIDE = 4 + 2 * RT_SMAX


def _declare_host(lib):
    lib.tnsp_abi_version.restype = c_int
    lib.tnsp_last_error.restype = ctypes.c_char_p
    lib.tnsp_launch_count.restype = c_i64
    lib.tnsp_rng_create_host.restype = c_vp
    lib.tnsp_rng_create_host.argtypes = [c_int]
    lib.tnsp_rng_destroy_host.argtypes = [c_vp]
    lib.tnsp_rng_seed_host.argtypes = [c_vp, c_int, ctypes.c_uint32]
    lib.tnsp_rng_uniform_int_host.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.tnsp_rng_uniform_real_host.argtypes = [c_vp, c_dbl, c_dbl, c_vp, c_vp]
    lib.tnsp_rng_normal_host.argtypes = [c_vp, c_int, c_dbl, c_dbl, c_i64, c_vp]
    return lib


def host_lib():
    """Host-only part of the C-ABI (RNG with libstdc++ semantics); needs no GPU."""
    global _host_lib
    if _host_lib is None:
        path = LIB_PATH if (_backend is not None and isinstance(_backend, CudaBackend)) else HOST_LIB_PATH
        if not os.path.exists(path):
            raise TnspError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C tnsp_b200/csrc)")
        _host_lib = _declare_host(ctypes.CDLL(path))
    return _host_lib


class CudaBackend:
    name = "cuda"

    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise TnspError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() / make -C tnsp_b200/csrc")
        if not torch.cuda.is_available():
            raise TnspError("tnsp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = lib = _declare_host(ctypes.CDLL(LIB_PATH))
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._dev_index = self.device.index
        # allocation prototypes: `proto.new_empty(shape)` skips the dtype / device argument parsing of torch.empty (the lock-step
        # engine allocates ~30 k buffers per step; the host is the bottleneck of a step)
        self._pf64 = torch.empty(0, dtype=torch.float64, device=self.device)
        self._pi32 = torch.empty(0, dtype=torch.int32, device=self.device)
        self._raw_stream = torch._C._cuda_getCurrentRawStream
        self._declare_kernels()
        self._rt_declare()

    def _declare_kernels(self):
        lib, P = self.lib, c_vp
        lib.tnsp_pack_f64.argtypes = [P, P, c_int, c_i64, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_pack_tiled_f64.argtypes = [c_i64] * 9 + [c_int, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_gemm_grouped_f64.argtypes = [P, c_int, P, P, c_i64, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_qr_batched_f64.argtypes = [P, c_int, P, P, c_i64, P, c_i64, P, c_i64, c_int, c_int, P]
        lib.tnsp_svd_work_size.restype = c_i64
        lib.tnsp_svd_work_size.argtypes = [P, c_int]
        lib.tnsp_svd_batched_f64.argtypes = [P, c_int, P, P, c_i64, P, c_i64, P, c_i64, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_qr_sectors_f64.argtypes = [P, P, P, c_i64, P, c_i64, P, c_i64, c_int, c_int, P]
        lib.tnsp_svd_sectors_f64.argtypes = [P, P, P, c_i64, P, c_i64, P, c_i64, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_gemm_gather_f64.argtypes = [P, c_i64, c_i64, c_i64, c_int, c_dbl, P, c_i64, P, c_i64, P, c_i64, c_int, P]
        # contract of dense tensors reads its operands in place (no pack); tests switch it off to cover the packed path
        self.gather_gemm = True
        lib.tnsp_qr_sectors_gather_f64.argtypes = [P, P, P, P, c_i64, P, c_i64, P, c_i64, c_int, c_int, P]
        lib.tnsp_svd_sectors_gather_f64.argtypes = [P, P, P, P, c_i64, P, c_i64, P, c_i64, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_sector_queue_min.restype = c_i64
        lib.tnsp_sector_queue_min.argtypes = [c_i64]
        lib.tnsp_factor_desc_kernels.argtypes = [c_int]
        lib.tnsp_gemm_skip_zero_fragments.argtypes = [c_int]
        lib.tnsp_jacobi_cached_norms.argtypes = [c_int]
        # set by tetragono.dense_embedding: single-descriptor factorisations discover their sectors on the device
        self.sector_discovery = False
        lib.tnsp_svd_cut_f64.argtypes = [P, c_int, c_i64, P, c_i64, c_i64, c_dbl, P, c_int, P]
        lib.tnsp_svd_mask_f64.argtypes = [P, c_int, P, P, c_i64, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_diag_scatter_f64.argtypes = [P, c_int, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_norm_f64.argtypes = [P, c_i64, c_i64, c_int, P, c_int, P]
        lib.tnsp_scale_f64.argtypes = [P, c_i64, P, c_i64, c_int, P, c_i64, c_i64, c_int, P]
        lib.tnsp_binary_f64.argtypes = [P, c_i64, P, c_i64, c_int, P, c_i64, c_i64, c_int, P]
        lib.tnsp_unary_f64.argtypes = [P, c_i64, c_int, P, c_i64, c_i64, c_int, P]
        lib.tnsp_grad_accumulate_f64.argtypes = [P, c_i64, P, P, P, P, c_i64, c_int, P]
        lib.tnsp_block_sign_f64.argtypes = [P, c_int, P, c_i64, P, c_i64, c_i64, c_int, P]
        lib.tnsp_gather_rows_f64.argtypes = [P, c_i64, P, P, c_i64, c_int, P]
        lib.tnsp_select_f64.argtypes = [P, P, c_i64, P, c_i64, P, c_i64, c_i64, c_int, P]

    # -- sector-compact lock-step tensors (TAT/ragged.py; csrc/ragged.cu, csrc/factor_sector.cu) ------------------------
    def _rt_declare(self):
        lib, P = self.lib, c_vp
        FP = ctypes.POINTER(RtForm)
        lib.tnsp_rt_sort_i32.argtypes = [c_int, P, P, P, P, c_i64, P, c_int, P]
        lib.tnsp_rt_match_i32.argtypes = [P, c_i64, c_int, P, c_i64, c_int, P, c_int, c_int, P, c_int, c_int, P, P, c_int, c_i64, P]
        lib.tnsp_rt_match_multi_i32.argtypes = [P, c_int, c_int, P]
        SP = ctypes.POINTER(RtMatchSpec)
        lib.tnsp_rt_repack_f64.argtypes = [P, FP, FP, SP, P, c_i64, c_i64, c_int, P]
        lib.tnsp_rt_repack_signed_f64.argtypes = [P, FP, FP, SP, P, c_i64, P, P, c_i64, c_int, P, P, c_int, c_int, P]
        lib.tnsp_rt_repack_pair_f64.argtypes = [P, FP, FP, SP, P, c_i64, c_i64, P, FP, FP, SP, P, c_i64, c_i64, c_int, P]
        lib.tnsp_rt_gemm_f64.argtypes = [FP, FP, FP, SP, P, c_i64, c_int, c_int, P]
        lib.tnsp_rt_dot_f64.argtypes = [P, FP, FP, P, c_int, c_int, P, c_int, c_int, P, c_i64, P, P, c_int, P]
        lib.tnsp_rt_factor_ws_ints.restype = c_i64
        lib.tnsp_rt_factor_ws_ints.argtypes = [c_i64]
        lib.tnsp_rt_svd_work_doubles.restype = c_i64
        lib.tnsp_rt_svd_work_doubles.argtypes = [c_i64, c_i64]
        lib.tnsp_rt_factor_plan.argtypes = [FP, c_int, c_int, P, c_int, c_int, c_i64, P, P, c_i64, c_int, P]
        lib.tnsp_rt_qr_f64.argtypes = [FP, c_int, P, c_int, c_int, P, c_i64, P, P, c_i64, P, P, c_i64, c_int, P]
        lib.tnsp_rt_svd_work_f64.argtypes = [FP, P, c_i64, P, c_i64, c_int, P]
        lib.tnsp_rt_svd_finish_f64.argtypes = [FP, c_int, P, c_int, c_int, c_i64, c_i64, c_dbl, P, c_i64, P, P, c_i64, c_int, P]
        lib.tnsp_rt_svd_scatter_f64.argtypes = [FP, c_int, P, c_int, c_int, P, c_i64, P, P, c_i64, P, P, c_i64, P, P, c_i64, P, c_i64, P, c_i64,
                                                c_int, P]
        lib.tnsp_rt_scale_f64.argtypes = [P, c_i64, P, c_i64, P, c_int, c_int, P, c_i64, c_i64, c_int, P]
        lib.tnsp_rt_binary_f64.argtypes = [P, c_i64, P, c_i64, P, c_i64, c_int, P, c_i64, c_i64, c_int, P]
        lib.tnsp_rt_norm_f64.argtypes = [P, c_i64, P, c_i64, c_int, P, c_int, P]
        lib.tnsp_rt_scalar_f64.argtypes = [P, c_i64, P, c_i64, P, c_int, P]
        lib.tnsp_rt_stats.argtypes = [c_int, P, c_int]
        # rt_factor evaluates these two size formulas on the host (two C calls less per factorisation): they must agree with the library
        if int(lib.tnsp_rt_factor_ws_ints(7)) != 8 + 6 * RT_SMAX + 7 or int(lib.tnsp_rt_svd_work_doubles(5, 9)) != 5 + 5 * 5 + 5 * 9 + 8:
            raise TnspError("libtnsp_b200.so: workspace size formulas differ from backend.py (rebuild the library)")

    @staticmethod
    def _st(t):
        """chain stride in elements of a [nb, size] buffer (0 broadcasts)"""
        return 0 if t.shape[0] == 1 else t.stride(0)

    def _form(self, f, data=None):
        """ctypes view of a ragged.Form (or of a dense device array)"""
        if isinstance(f, torch.Tensor):
            return RtForm(f.data_ptr(), self._st(f), None, 0, 0, None, 0, 0, None, 0)
        if data is None:
            c = f._c                  # a Form's arrays never change once its match table exists: the ctypes view is built once
            if c is None:
                d, rt, ct, m = f.data, f.rt, f.ct, f.match
                c = f._c = RtForm(d.data_ptr(), 0 if d.shape[0] == 1 else d.stride(0), rt.data_ptr(), 0 if rt.shape[0] == 1 else rt.stride(0),
                                  f.M, ct.data_ptr(), 0 if ct.shape[0] == 1 else ct.stride(0), f.N, m.data_ptr(),
                                  0 if m.shape[0] == 1 else m.stride(0))
            return c
        d = data
        return RtForm(d.data_ptr(), self._st(d), f.rt.data_ptr(), self._st(f.rt), f.M, f.ct.data_ptr(), self._st(f.ct), f.N,
                      f.match.data_ptr(), self._st(f.match))

    def rt_alloc(self, nb, size):
        # even row length: every chain's storage starts 16-byte aligned (the GEMM brings whole B operands in by TMA bulk copies)
        return self._pf64.new_empty((nb, ((max(int(size), 1) + 1) & ~1) + RT_SMAX))

    _SORT_ARRAYS: dict = {}

    def rt_sort(self, edges):
        n = len(edges)
        nbT, M = 1, 1
        for a, _, d in edges:
            if a.shape[0] > nbT:
                nbT = a.shape[0]
            M *= d
        table = self._pi32.new_empty((nbT, RT_HDR + 2 * M))
        m = n if n > 0 else 1
        types = self._SORT_ARRAYS.get(m)
        if types is None:
            types = self._SORT_ARRAYS[m] = (c_vp * m, c_i64 * m, ctypes.c_int32 * m)
        ptrs, strides, dims, signs = types[0](), types[1](), types[2](), types[2]()
        for i, (a, sg, d) in enumerate(edges):
            ptrs[i] = a.data_ptr()
            strides[i] = 0 if a.shape[0] == 1 else a.stride(0)
            dims[i] = d
            signs[i] = sg
        self._ck(self.lib.tnsp_rt_sort_i32(n, ptrs, strides, dims, signs, M, table.data_ptr(), nbT, self._stream()))
        return table

    def rt_match(self, rt, rs, ct, cs, t1, s1, t2, s2, nbm, cap=0):
        nbm = max(int(nbm), rt.shape[0], ct.shape[0], 1 if t1 is None else t1.shape[0], 1 if t2 is None else t2.shape[0])
        match = self._pi32.new_empty((nbm, RT_MSTRIDE))
        tsum = self._pi32.new_empty((nbm,)) if (t1 is not None or t2 is not None) else None
        self._ck(self.lib.tnsp_rt_match_i32(rt.data_ptr(), self._st(rt), int(rs), ct.data_ptr(), self._st(ct), int(cs),
                                            None if t1 is None else t1.data_ptr(), 0 if t1 is None or t1.shape[0] == 1 else 1, int(s1),
                                            None if t2 is None else t2.data_ptr(), 0 if t2 is None or t2.shape[0] == 1 else 1, int(s2),
                                            match.data_ptr(), None if tsum is None else tsum.data_ptr(), nbm, int(cap), self._stream()))
        return match, tsum

    def rt_match_many(self, jobs, nb):
        """several pairings in ONE launch (the pairing tables of a factorisation's factors); jobs = [(rt, rs, ct, cs, t1, s1, t2, s2, cap)]
        -> list of match tables [nb, MSTRIDE] (views of one allocation)"""
        n = len(jobs)
        out = self._pi32.new_empty((n, nb, RT_MSTRIDE))
        views = [out[i] for i in range(n)]
        # (structs built by their constructor: 15 attribute assignments through ctypes cost ~1 us each)
        arr = (RtMatchJob * n)(*[
            RtMatchJob(rt.data_ptr(), 0 if rt.shape[0] == 1 else rt.stride(0), int(rs), ct.data_ptr(), 0 if ct.shape[0] == 1 else ct.stride(0), int(cs),
                       None if t1 is None else t1.data_ptr(), 0 if t1 is None or t1.shape[0] == 1 else 1, int(s1),
                       None if t2 is None else t2.data_ptr(), 0 if t2 is None or t2.shape[0] == 1 else 1, int(s2), m.data_ptr(), None, int(cap))
            for (rt, rs, ct, cs, t1, s1, t2, s2, cap), m in zip(jobs, views)])
        self._ck(self.lib.tnsp_rt_match_multi_i32(arr, n, nb, self._stream()))
        return views

    def _spec(self, f, spec, nbm, want_tsum):
        """spec = (rs, cs, t1, s1, t2, s2): allocate the match table of form `f` (and the summed target), to be filled by the kernel"""
        rs, cs, t1, s1, t2, s2 = spec
        nbm = max(int(nbm), f.rt.shape[0], f.ct.shape[0], 1 if t1 is None else t1.shape[0], 1 if t2 is None else t2.shape[0])
        f.match = self._pi32.new_empty((nbm, RT_MSTRIDE))
        f._c = None
        tsum = self._pi32.new_empty((nbm,)) if (want_tsum and (t1 is not None or t2 is not None)) else None
        c = RtMatchSpec(int(rs), int(cs), None if t1 is None else t1.data_ptr(), 0 if t1 is None or t1.shape[0] == 1 else 1, int(s1),
                        None if t2 is None else t2.data_ptr(), 0 if t2 is None or t2.shape[0] == 1 else 1, int(s2),
                        f.match.data_ptr(), self._st(f.match), None if tsum is None else tsum.data_ptr(), int(f.data.shape[1]))
        return c, tsum

    def rt_repack(self, plan, src, dst, match_spec=None, sign=None):
        """regroup src -> dst; with match_spec the destination's sector pairing is computed by the same launch (dst.match is set);
        sign = (quad, per_chain, [(label array, dim)...], fermi mask): fermionic sign of every element (TAT/ragged_fermi.py)"""
        if sign is not None:
            return self._rt_repack_signed(plan, src, dst, match_spec, sign)
        dst_dense = isinstance(dst, torch.Tensor)
        out = dst if dst_dense else dst.data
        nb = out.shape[0]
        if not isinstance(src, torch.Tensor):
            nb = max(nb, src.match.shape[0])
        spec = None
        if match_spec is not None:
            spec, _ = self._spec(dst, match_spec, 1, False)
            nb = max(nb, dst.match.shape[0])
        work = out.shape[1] if dst_dense else dst.M * dst.N + RT_SMAX
        fs, fd = self._form(src), self._form(dst)
        self._ck(self.lib.tnsp_rt_repack_f64(plan.data_ptr(), ctypes.byref(fs), ctypes.byref(fd), None if spec is None else ctypes.byref(spec),
                                             out.data_ptr(), out.stride(0), int(work), nb, self._stream()))

    def _rt_repack_signed(self, plan, src, dst, match_spec, sign):
        quad, per_chain, labels, fermi = sign
        spec, _ = self._spec(dst, match_spec, 1, False)
        nb = max(dst.data.shape[0], src.match.shape[0], dst.match.shape[0], per_chain.shape[0])
        n = len(labels)
        ptrs = (c_vp * max(n, 1))(*[a.data_ptr() for a, _ in labels])
        strides = (c_i64 * max(n, 1))(*[self._st(a) for a, _ in labels])
        fs, fd = self._form(src), self._form(dst)
        self._ck(self.lib.tnsp_rt_repack_signed_f64(plan.data_ptr(), ctypes.byref(fs), ctypes.byref(fd), ctypes.byref(spec), dst.data.data_ptr(),
                                                    dst.data.stride(0), quad.data_ptr(), per_chain.data_ptr(), 0 if per_chain.shape[0] == 1 else 1,
                                                    n, ptrs, strides, int(fermi), nb, self._stream()))

    def rt_repack_pair(self, plan0, src0, dst0, spec0, plan1, src1, dst1, spec1):
        """the two regroupings of a contraction in one launch"""
        s0, _ = self._spec(dst0, spec0, 1, False)
        s1, _ = self._spec(dst1, spec1, 1, False)
        nb = max(dst0.data.shape[0], dst1.data.shape[0], src0.match.shape[0], src1.match.shape[0], dst0.match.shape[0], dst1.match.shape[0])
        f = [self._form(x) for x in (src0, dst0, src1, dst1)]
        self._ck(self.lib.tnsp_rt_repack_pair_f64(plan0.data_ptr(), ctypes.byref(f[0]), ctypes.byref(f[1]), ctypes.byref(s0), dst0.data.data_ptr(),
                                                  dst0.data.stride(0), dst0.M * dst0.N + RT_SMAX, plan1.data_ptr(), ctypes.byref(f[2]),
                                                  ctypes.byref(f[3]), ctypes.byref(s1), dst1.data.data_ptr(), dst1.data.stride(0),
                                                  dst1.M * dst1.N + RT_SMAX, nb, self._stream()))

    def rt_gemm(self, A, B, C, ksign, nb, match_spec=None):
        """C = A B per (chain, sector); with match_spec C's sector pairing (and summed target, returned) come out of the same launch"""
        spec, tsum = (None, None)
        if match_spec is not None:
            spec, tsum = self._spec(C, match_spec, 1, True)
        fa, fb, fc = self._form(A), self._form(B), self._form(C)
        self._ck(self.lib.tnsp_rt_gemm_f64(ctypes.byref(fa), ctypes.byref(fb), ctypes.byref(fc), None if spec is None else ctypes.byref(spec),
                                           C.data.data_ptr(), C.data.stride(0), int(ksign), nb, self._stream()))
        return tsum

    def rt_dot(self, plan, src, dst, t1, s1, t2, s2, nb):
        """full contraction of two tensors -> (data [nb, 2 + SMAX], match, summed target or None)"""
        out = self.rt_alloc(nb, 2)
        match = self._pi32.new_empty((nb, RT_MSTRIDE))
        tsum = self._pi32.new_empty((nb,)) if (t1 is not None or t2 is not None) else None
        fs, fd = self._form(src), self._form(dst)
        self._ck(self.lib.tnsp_rt_dot_f64(plan.data_ptr(), ctypes.byref(fs), ctypes.byref(fd), None if t1 is None else t1.data_ptr(),
                                          0 if t1 is None or t1.shape[0] == 1 else 1, int(s1), None if t2 is None else t2.data_ptr(),
                                          0 if t2 is None or t2.shape[0] == 1 else 1, int(s2), out.data_ptr(), out.stride(0), match.data_ptr(),
                                          None if tsum is None else tsum.data_ptr(), nb, self._stream()))
        return out, match, tsum

    # the factorisation is a short fixed sequence of launches; each step is a method of its own so that the bench's instrumented
    # pass can time the kernel classes separately (profiling.KernelTimer)
    def rt_factor_plan(self, ff, code, frs, t1p, t1st, t1s, kd, labels, ws, wss, nb):
        self._ck(self.lib.tnsp_rt_factor_plan(ctypes.byref(ff), code, frs, t1p, t1st, t1s, kd, labels.data_ptr(), ws.data_ptr(), wss, nb, self._stream()))

    def rt_qr_work(self, ff, frs, t1p, t1st, t1s, tab, m_first, first, m_second, second, nb):
        self._ck(self.lib.tnsp_rt_qr_f64(ctypes.byref(ff), frs, t1p, t1st, t1s, tab.data_ptr(), self._st(tab), m_first.data_ptr(), first.data_ptr(),
                                         first.stride(0), m_second.data_ptr(), second.data_ptr(), second.stride(0), nb, self._stream()))

    def rt_svd_work(self, ff, work, ws, wss, nb):
        self._ck(self.lib.tnsp_rt_svd_work_f64(ctypes.byref(ff), work.data_ptr(), work.stride(0), ws.data_ptr(), wss, nb, self._stream()))

    def rt_svd_finish(self, ff, frs, t1p, t1st, t1s, kd, remain_cut, relative_cut, work, labels, ws, wss, nb):
        self._ck(self.lib.tnsp_rt_svd_finish_f64(ctypes.byref(ff), frs, t1p, t1st, t1s, kd, int(min(remain_cut, 1 << 40)), float(relative_cut),
                                                 work.data_ptr(), work.stride(0), labels.data_ptr(), ws.data_ptr(), wss, nb, self._stream()))

    def rt_svd_scatter(self, ff, frs, t1p, t1st, t1s, tab, m_first, first, m_s, s_data, m_second, second, work, ws, wss, nb):
        self._ck(self.lib.tnsp_rt_svd_scatter_f64(ctypes.byref(ff), frs, t1p, t1st, t1s, tab.data_ptr(), self._st(tab), m_first.data_ptr(),
                                                  first.data_ptr(), first.stride(0), m_s.data_ptr(), s_data.data_ptr(), s_data.stride(0),
                                                  m_second.data_ptr(), second.data_ptr(), second.stride(0), work.data_ptr(), work.stride(0),
                                                  ws.data_ptr(), wss, nb, self._stream()))

    def rt_factor(self, kind, F, fsign, tt, tts, t1, t1s, kdim, remain_cut, relative_cut, nb, caps=None):
        lib = self.lib
        kd = max(int(kdim), 1)
        kfull = max(min(F.M, F.N), 1)
        frs = int(fsign) * int(F.rs)
        ff = self._form(F)
        t1p = None if t1 is None else t1.data_ptr()
        t1st = 0 if t1 is None or t1.shape[0] == 1 else 1
        t1s = int(t1s) if t1 is not None else 0
        wss = 8 + 6 * RT_SMAX + kfull          # = tnsp_rt_factor_ws_ints(kfull)
        ws = self._pi32.new_empty((nb, wss))
        labels = self._pi32.new_empty((nb, kd))
        code = 0 if kind == "qr" else 2
        self.rt_factor_plan(ff, code, frs, t1p, t1st, t1s, kd, labels, ws, wss, nb)
        if code == 2:
            work = self._pf64.new_empty((nb, kfull + F.M * kfull + kfull * F.N + 8))     # = tnsp_rt_svd_work_doubles(M, N)
            self.rt_svd_work(ff, work, ws, wss, nb)
            self.rt_svd_finish(ff, frs, t1p, t1st, t1s, kd, remain_cut, relative_cut, work, labels, ws, wss, nb)
        tab = self.rt_sort([(labels, 1, kd)])
        first = self.rt_alloc(nb, caps[0] if caps else F.M * kd)
        second = self.rt_alloc(nb, caps[1] if caps else kd * F.N)
        # the pairing tables of the factors (and of the singular-value tensor): one launch
        jobs = [(F.rt, F.rs * fsign, tab, 1, t1, t1s, None, 0, first.shape[1]),
                (tab, -1, F.ct, F.cs * fsign, tt, tts, t1, -t1s if t1 is not None else 0, second.shape[1])]
        if code != 0:
            s_data = self.rt_alloc(nb, kd * kd)
            jobs.append((tab, -1, tab, 1, None, 0, None, 0, s_data.shape[1]))
        ms = self.rt_match_many(jobs, nb)
        m_first, m_second = ms[0], ms[1]
        out = {"labels": labels, "bond_col": (tab, 1), "bond_row": (tab, -1), "first": (m_first, first), "second": (m_second, second)}
        if code == 0:
            self.rt_qr_work(ff, frs, t1p, t1st, t1s, tab, m_first, first, m_second, second, nb)
            return out
        m_s = ms[2]
        self.rt_svd_scatter(ff, frs, t1p, t1st, t1s, tab, m_first, first, m_s, s_data, m_second, second, work, ws, wss, nb)
        out["s"] = (m_s, s_data)
        return out

    def rt_overflow(self, clear=True):
        """number of (chain, tensor) pairs dropped because their sectors did not fit a learnt capacity (synchronises)"""
        out = (ctypes.c_uint64 * 16)()
        self._ck(self.lib.tnsp_rt_stats(-1, out, 2 if clear else 0))
        return int(out[15])

    def rt_stats(self, enable=-1, read=False, reset=False):
        """device work counters of the sector-compact kernels (include/tnsp_b200.h: tnsp_rt_stats); reading synchronises"""
        out = (ctypes.c_uint64 * 16)() if read else None
        self._ck(self.lib.tnsp_rt_stats(int(enable), out, int(bool(reset))))
        return list(out) if read else None

    def rt_scale(self, data, match, vec, op):
        nb = max(data.shape[0], match.shape[0], vec.shape[0])
        out = self._pf64.new_empty((nb, data.shape[1]))
        self._ck(self.lib.tnsp_rt_scale_f64(data.data_ptr(), self._st(data), match.data_ptr(), self._st(match), vec.data_ptr(),
                                            0 if vec.shape[0] == 1 else 1, int(op), out.data_ptr(), out.stride(0), data.shape[1], nb, self._stream()))
        return out

    def rt_binary(self, a, b, match, op):
        nb = max(a.shape[0], b.shape[0], match.shape[0])
        out = self._pf64.new_empty((nb, a.shape[1]))
        self._ck(self.lib.tnsp_rt_binary_f64(a.data_ptr(), self._st(a), b.data_ptr(), self._st(b), match.data_ptr(), self._st(match), int(op),
                                             out.data_ptr(), out.stride(0), a.shape[1], nb, self._stream()))
        return out

    def rt_norm(self, data, match, kind):
        nb = max(data.shape[0], match.shape[0])
        out = self._pf64.new_empty((nb,))
        self._ck(self.lib.tnsp_rt_norm_f64(data.data_ptr(), self._st(data), match.data_ptr(), self._st(match), int(kind), out.data_ptr(), nb,
                                           self._stream()))
        return out

    def rt_scalar(self, data, match):
        nb = max(data.shape[0], match.shape[0])
        out = self._pf64.new_empty((nb,))
        self._ck(self.lib.tnsp_rt_scalar_f64(data.data_ptr(), self._st(data), match.data_ptr(), self._st(match), out.data_ptr(), nb, self._stream()))
        return out

    # -- buffers ------------------------------------------------------------------------------
    def empty(self, nb, size):
        return self._pf64.new_empty((nb, size))

    def zeros(self, nb, size):
        return torch.zeros((nb, size), dtype=torch.float64, device=self.device)

    def from_numpy(self, array):
        return torch.from_numpy(np.ascontiguousarray(array)).to(self.device)

    def to_numpy(self, t):
        return t.detach().cpu().numpy()

    def upload(self, array):
        return torch.from_numpy(np.ascontiguousarray(array)).to(self.device)

    def _stream(self):
        return self._raw_stream(self._dev_index)

    def _ck(self, rc):
        if rc != 0:
            raise TnspError(self.lib.tnsp_last_error().decode())

    @staticmethod
    def _bs(t):
        """batch stride in elements (0 broadcasts a single entry)"""
        return 0 if t.shape[0] == 1 else t.stride(0)

    def launch_count(self):
        return int(self.lib.tnsp_launch_count())

    def synchronize(self):
        torch.cuda.synchronize()

    # -- kernels ------------------------------------------------------------------------------
    def pack(self, plan, src, dst):
        dev = plan._dev
        if dev is None:
            dev = plan._dev = (self.upload(plan.desc), self.upload(plan.estart))
        nb = dst.shape[0]
        self._ck(self.lib.tnsp_pack_f64(dev[0].data_ptr(), dev[1].data_ptr(), len(plan.desc), plan.total, src.data_ptr(), self._bs(src),
                                        dst.data_ptr(), dst.stride(0), nb, self._stream()))

    def gemm(self, plan, a, b, c):
        dev = plan._dev
        if dev is None:
            dev = plan._dev = self.upload(plan.gemm)
        nb = c.shape[0]
        self._ck(self.lib.tnsp_gemm_grouped_f64(dev.data_ptr(), len(plan.gemm), plan.gemm.ctypes.data, a.data_ptr(), self._bs(a),
                                                b.data_ptr(), self._bs(b), c.data_ptr(), c.stride(0), nb, self._stream()))

    def gemm_gather(self, plan, a, b, c):
        tab, flags, m, n, k = plan.gather
        dev = plan._gdev
        if dev is None:
            dev = plan._gdev = self.upload(tab)
        nb = c.shape[0]
        self._ck(self.lib.tnsp_gemm_gather_f64(dev.data_ptr(), m, n, k, flags, 1.0, a.data_ptr(), self._bs(a), b.data_ptr(), self._bs(b),
                                               c.data_ptr(), c.stride(0), nb, self._stream()))

    def qr_destroys_input(self, plan):
        """only the descriptor-driven kernel for matrices beyond shared memory factorises in place"""
        return not (self.sector_discovery and len(plan.sectors) == 1)

    def _sect(self, plan):
        dev = plan._dev
        if dev is None:
            dev = plan._dev = self.upload(plan.sectors)
        return dev

    def factor_in_place(self, plan):
        """qr / svd of a dense(-embedded) tensor can read the operand through the plan's offset table (no merged copy)
        when the per-sector work-queue kernels apply"""
        return (self.sector_discovery and self.gather_gemm and len(plan.sectors) == 1
                and int(plan.sectors[0][0]) * int(plan.sectors[0][1]) >= int(self.lib.tnsp_sector_queue_min(-1)))

    def _rc(self, plan):
        dev = plan._rcdev
        if dev is None:
            dev = plan._rcdev = self.upload(plan.rc_tab)
        return dev

    def qr(self, plan, a, out1, out2, in_place=False):
        dev = self._sect(plan)
        nb = a.shape[0]
        if in_place:
            self._ck(self.lib.tnsp_qr_sectors_gather_f64(dev.data_ptr(), plan.sectors.ctypes.data, self._rc(plan).data_ptr(), a.data_ptr(),
                                                         self._bs(a), out1.data_ptr(), out1.stride(0), out2.data_ptr(), out2.stride(0),
                                                         int(plan.flag), nb, self._stream()))
            return
        if self.sector_discovery and len(plan.sectors) == 1:
            self._ck(self.lib.tnsp_qr_sectors_f64(dev.data_ptr(), plan.sectors.ctypes.data, a.data_ptr(), a.stride(0), out1.data_ptr(),
                                                  out1.stride(0), out2.data_ptr(), out2.stride(0), int(plan.flag), nb, self._stream()))
            return
        self._ck(self.lib.tnsp_qr_batched_f64(dev.data_ptr(), len(plan.sectors), plan.sectors.ctypes.data, a.data_ptr(), a.stride(0),
                                              out1.data_ptr(), out1.stride(0), out2.data_ptr(), out2.stride(0), int(plan.flag), nb,
                                              self._stream()))

    def svd(self, plan, a, out1, s, out2, in_place=False):
        dev = self._sect(plan)
        nb = a.shape[0]
        wsize = int(self.lib.tnsp_svd_work_size(plan.sectors.ctypes.data, len(plan.sectors)))
        work = self.empty(nb, max(wsize, 1))
        if in_place:
            self._ck(self.lib.tnsp_svd_sectors_gather_f64(dev.data_ptr(), plan.sectors.ctypes.data, self._rc(plan).data_ptr(), a.data_ptr(),
                                                          self._bs(a), out1.data_ptr(), out1.stride(0), s.data_ptr(), s.stride(0),
                                                          out2.data_ptr(), out2.stride(0), work.data_ptr(), work.stride(0), nb,
                                                          self._stream()))
            return
        if self.sector_discovery and len(plan.sectors) == 1:
            self._ck(self.lib.tnsp_svd_sectors_f64(dev.data_ptr(), plan.sectors.ctypes.data, a.data_ptr(), a.stride(0), out1.data_ptr(),
                                                   out1.stride(0), s.data_ptr(), s.stride(0), out2.data_ptr(), out2.stride(0),
                                                   work.data_ptr(), work.stride(0), nb, self._stream()))
            return
        self._ck(self.lib.tnsp_svd_batched_f64(dev.data_ptr(), len(plan.sectors), plan.sectors.ctypes.data, a.data_ptr(), a.stride(0),
                                               out1.data_ptr(), out1.stride(0), s.data_ptr(), s.stride(0), out2.data_ptr(),
                                               out2.stride(0), work.data_ptr(), work.stride(0), nb, self._stream()))

    def svd_cut(self, plan, s, remain_cut, relative_cut):
        dev = self._sect(plan)
        nb = s.shape[0]
        counts = torch.empty((nb, len(plan.sectors)), dtype=torch.int32, device=self.device)
        self._ck(self.lib.tnsp_svd_cut_f64(dev.data_ptr(), len(plan.sectors), plan.s_total, s.data_ptr(), s.stride(0), int(remain_cut),
                                           float(relative_cut), counts.data_ptr(), nb, self._stream()))
        return counts

    def svd_mask(self, plan, counts, out1, s, out2):
        dev = self._sect(plan)
        nb = s.shape[0]
        self._ck(self.lib.tnsp_svd_mask_f64(dev.data_ptr(), len(plan.sectors), counts.data_ptr(), out1.data_ptr(), out1.stride(0),
                                            s.data_ptr(), s.stride(0), out2.data_ptr(), out2.stride(0), nb, self._stream()))

    def diag_scatter(self, blk, s, dst):
        if len(blk) == 0:
            return
        d = self.upload(blk)
        self._ck(self.lib.tnsp_diag_scatter_f64(d.data_ptr(), len(blk), s.data_ptr(), s.stride(0), dst.data_ptr(), dst.stride(0),
                                                dst.shape[0], self._stream()))

    def norm(self, x, kind):
        nb = x.shape[0]
        out = self._pf64.new_empty((nb,))
        self._ck(self.lib.tnsp_norm_f64(x.data_ptr(), x.stride(0), x.shape[1], kind, out.data_ptr(), nb, self._stream()))
        return out

    def scale(self, x, alpha, op, nb=None):
        """y[b] = x[b] * alpha[b] (op 0) or x[b] / alpha[b] (op 1); alpha is a device vector of nb or 1 entries."""
        nb = max(x.shape[0], alpha.shape[0]) if nb is None else nb
        y = self.empty(nb, x.shape[1])
        self._ck(self.lib.tnsp_scale_f64(x.data_ptr(), self._bs(x), alpha.data_ptr(), 0 if alpha.shape[0] == 1 else 1, op, y.data_ptr(),
                                         y.stride(0), x.shape[1], nb, self._stream()))
        return y

    def binary(self, a, b, op):
        nb = max(a.shape[0], b.shape[0])
        z = self.empty(nb, a.shape[1])
        self._ck(self.lib.tnsp_binary_f64(a.data_ptr(), self._bs(a), b.data_ptr(), self._bs(b), op, z.data_ptr(), z.stride(0), a.shape[1],
                                          nb, self._stream()))
        return z

    def unary(self, a, op):
        z = self.empty(a.shape[0], a.shape[1])
        self._ck(self.lib.tnsp_unary_f64(a.data_ptr(), a.stride(0), op, z.data_ptr(), z.stride(0), a.shape[1], a.shape[0], self._stream()))
        return z

    def block_sign(self, blk, x):
        y = self.empty(x.shape[0], x.shape[1])
        d = self.upload(blk)
        self._ck(self.lib.tnsp_block_sign_f64(d.data_ptr(), len(blk), x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), x.shape[1],
                                              x.shape[0], self._stream()))
        return y

    def gather_rows(self, src, row_size, index):
        """dst[b] = src.view(-1,row_size)[index[b]]; index int32 device vector"""
        nb = index.shape[0]
        dst = self.empty(nb, row_size)
        self._ck(self.lib.tnsp_gather_rows_f64(src.data_ptr(), row_size, index.data_ptr(), dst.data_ptr(), dst.stride(0), nb,
                                               self._stream()))
        return dst

    def select(self, mask, a, b):
        nb = mask.shape[0]
        dst = self.empty(nb, a.shape[1])
        self._ck(self.lib.tnsp_select_f64(mask.data_ptr(), a.data_ptr(), self._bs(a), b.data_ptr(), self._bs(b), dst.data_ptr(),
                                          dst.stride(0), a.shape[1], nb, self._stream()))
        return dst

    def grad_accumulate(self, holes, weight, energy, delta, edelta):
        self._ck(self.lib.tnsp_grad_accumulate_f64(holes.data_ptr(), holes.stride(0), weight.data_ptr(), energy.data_ptr(),
                                                   delta.data_ptr(), edelta.data_ptr(), holes.shape[1], holes.shape[0], self._stream()))


def set_backend(backend):
    """Install a backend object explicitly (tests only)."""
    global _backend, _host_lib
    _backend = backend
    _host_lib = None


def get():
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend

// K2: grouped FP64 GEMM over (symmetry sector x Monte-Carlo chain).
//
// Replaces detail::gemm_batch (TAT/include/TAT/implement/contract.hpp:194-250), i.e. one ?gemm_
// call per sector per contract, by ONE launch that walks the descriptor list the host planner
// built from contract.hpp:539-616 (symmetric) / :826-852 (no symmetry), for all chains of a batch.
//
// FP64 has no tcgen05/UMMA kind on Blackwell; the FP64 tensor path is the warp-level DMMA
// (mma.sync.m8n8k4.f64).  Operand tiles are staged in padded shared memory (conflict-free for the
// DMMA fragment layout) with register prefetch of the next K-slab while the current one is
// multiplied.
//
// Roofline: FP64 tensor pipe when min(m,n,k) >~ 128, HBM otherwise; algorithmic flops = sum 2mnk.
#include "common.cuh"
#include <climits>
#include <cstdint>

namespace tnsp {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int BK = 16;

// 4 warps as 2x2; warp tile (8*TM) x (8*TN); CTA tile BM = 16*TM, BN = 16*TN.
template <int TM, int TN>
__global__ void __launch_bounds__(128) gemm_grouped_kernel(const int64_t* __restrict__ desc, int ng, const double* __restrict__ a,
                                                           int64_t abs_, const double* __restrict__ b, int64_t bbs,
                                                           double* __restrict__ c, int64_t cbs, int nb) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    constexpr int LDA = BK + 4;   // As[BM][LDA]  : (row * 20 + k) -> conflict-free A fragments
    constexpr int LDB = BN + 4;   // Bs[BK][LDB]  : (k * (BN+4) + n) -> conflict-free B fragments
    __shared__ double As[BM * LDA];
    __shared__ double Bs[BK * LDB];

    // locate (descriptor, tile) of this CTA
    int64_t t = blockIdx.x;
    int g = 0;
    int64_t m = 0, n = 0, k = 0, tiles_n = 0;
    for (; g < ng; ++g) {
        m = desc[g * TNSP_GEMM_COLS + 0];
        n = desc[g * TNSP_GEMM_COLS + 1];
        k = desc[g * TNSP_GEMM_COLS + 2];
        tiles_n = (n + BN - 1) / BN;
        const int64_t tiles = ((m + BM - 1) / BM) * tiles_n;
        if (t < tiles) break;
        t -= tiles;
    }
    if (g >= ng) return;
    const int64_t* d = desc + g * TNSP_GEMM_COLS;
    const int64_t a_off = d[3], b_off = d[4], c_off = d[5], flags = d[6];
    const double alpha = (double)d[7];
    const bool a_km = flags & 1;   // A stored [k x m]
    const bool b_nk = flags & 2;   // B stored [n x k]
    const int64_t row0 = (t / tiles_n) * BM, col0 = (t % tiles_n) * BN;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 8 * TM, wn = (warp & 1) * 8 * TN;
    const int gid = lane >> 2, tig = lane & 3;

    constexpr int A_PER = (BM * BK) / 128;  // elements per thread per slab
    constexpr int B_PER = (BK * BN) / 128;

    for (int bi = blockIdx.y; bi < nb; bi += gridDim.y) {
        const double* A = a + (int64_t)bi * abs_ + a_off;
        const double* B = b + (int64_t)bi * bbs + b_off;
        double* C = c + (int64_t)bi * cbs + c_off;

        double acc[TM][TN][2];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        double ra[A_PER], rb[B_PER];
        auto load_slab = [&](int64_t k0) {
#pragma unroll
            for (int u = 0; u < A_PER; ++u) {
                const int e = tid + u * 128;
                int r, kk;
                if (a_km) { r = e % BM; kk = e / BM; } else { kk = e % BK; r = e / BK; }
                const int64_t gr = row0 + r, gk = k0 + kk;
                double v = 0.0;
                if (gr < m && gk < k) v = a_km ? __ldg(A + gk * m + gr) : __ldg(A + gr * k + gk);
                ra[u] = v;
            }
#pragma unroll
            for (int u = 0; u < B_PER; ++u) {
                const int e = tid + u * 128;
                int kk, cc;
                if (b_nk) { kk = e % BK; cc = e / BK; } else { cc = e % BN; kk = e / BN; }
                const int64_t gk = k0 + kk, gc = col0 + cc;
                double v = 0.0;
                if (gk < k && gc < n) v = b_nk ? __ldg(B + gc * k + gk) : __ldg(B + gk * n + gc);
                rb[u] = v;
            }
        };
        auto store_slab = [&]() {
#pragma unroll
            for (int u = 0; u < A_PER; ++u) {
                const int e = tid + u * 128;
                int r, kk;
                if (a_km) { r = e % BM; kk = e / BM; } else { kk = e % BK; r = e / BK; }
                As[r * LDA + kk] = ra[u];
            }
#pragma unroll
            for (int u = 0; u < B_PER; ++u) {
                const int e = tid + u * 128;
                int kk, cc;
                if (b_nk) { kk = e % BK; cc = e / BK; } else { cc = e % BN; kk = e / BN; }
                Bs[kk * LDB + cc] = rb[u];
            }
        };

        load_slab(0);
        for (int64_t k0 = 0; k0 < k; k0 += BK) {
            __syncthreads();   // previous slab fully consumed
            store_slab();
            __syncthreads();
            if (k0 + BK < k) load_slab(k0 + BK);   // prefetch next slab into registers
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                double fa[TM], fb[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) fa[i] = As[(wm + i * 8 + gid) * LDA + kk + tig];
#pragma unroll
                for (int j = 0; j < TN; ++j) fb[j] = Bs[(kk + tig) * LDB + wn + j * 8 + gid];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
            }
        }
        // epilogue: thread holds C[row = gid][col = 2*tig, 2*tig+1] of each 8x8 tile
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int64_t r = row0 + wm + i * 8 + gid;
            if (r >= m) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int64_t cc = col0 + wn + j * 8 + 2 * tig;
                if (cc < n) C[r * n + cc] = alpha * acc[i][j][0];
                if (cc + 1 < n) C[r * n + cc + 1] = alpha * acc[i][j][1];
            }
        }
        __syncthreads();
    }
}


// Small-sector path (m, n <= 16): one WARP computes the whole C of one (sector, chain) with operand
// fragments loaded straight from global memory in the DMMA register layout -- no shared memory, no
// barrier; a CTA of 4 warps covers 4 chains.  This is the shape class of almost every GEMM of the
// cfg1-sized lattices (16x16x16, 4x4x16, 16x4x64 ...), where launch latency, not bandwidth, rules.
__global__ void __launch_bounds__(128) gemm_small_kernel(const int64_t* __restrict__ desc, int ng, const double* __restrict__ a,
                                                         int64_t abs_, const double* __restrict__ b, int64_t bbs,
                                                         double* __restrict__ c, int64_t cbs, int nb) {
    const int g = blockIdx.x;
    const int64_t* d = desc + g * TNSP_GEMM_COLS;
    const int m = (int)d[0], n = (int)d[1], k = (int)d[2];
    const int64_t a_off = d[3], b_off = d[4], c_off = d[5], flags = d[6];
    const double alpha = (double)d[7];
    const bool a_km = flags & 1, b_nk = flags & 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    for (int bi = blockIdx.y * 4 + warp; bi < nb; bi += gridDim.y * 4) {
        const double* A = a + (int64_t)bi * abs_ + a_off;
        const double* B = b + (int64_t)bi * bbs + b_off;
        double* C = c + (int64_t)bi * cbs + c_off;
        double acc[2][2][2] = {};
        for (int k0 = 0; k0 < k; k0 += 4) {
            const int kk = k0 + tig;
            double fa[2], fb[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = i * 8 + gid;
                fa[i] = (r < m && kk < k) ? (a_km ? __ldg(A + (int64_t)kk * m + r) : __ldg(A + (int64_t)r * k + kk)) : 0.0;
                const int cc = i * 8 + gid;
                fb[i] = (cc < n && kk < k) ? (b_nk ? __ldg(B + (int64_t)cc * k + kk) : __ldg(B + (int64_t)kk * n + cc)) : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = i * 8 + gid;
            if (r >= m) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int cc = j * 8 + 2 * tig;
                if (cc < n) C[(int64_t)r * n + cc] = alpha * acc[i][j][0];
                if (cc + 1 < n) C[(int64_t)r * n + cc + 1] = alpha * acc[i][j][1];
            }
        }
    }
}

// Long-K, few-outputs path (amplitude closures <left strip | right strip>: m = n = 1, k = Dc^2 D^2 ...; SR-CG
// mat-vecs with few rows): ONE CTA per output element C[i][j], all 256 threads stride over k, fixed-order block
// reduction.  HBM bound: reads 2k doubles per output.
__global__ void __launch_bounds__(256) gemm_dot_kernel(const int64_t* __restrict__ desc, int ng, const double* __restrict__ a,
                                                       int64_t abs_, const double* __restrict__ b, int64_t bbs,
                                                       double* __restrict__ c, int64_t cbs, int nb) {
    __shared__ double red[32];
    int64_t t = blockIdx.x;
    int g = 0;
    int64_t m = 0, n = 0, k = 0;
    for (; g < ng; ++g) {
        m = desc[g * TNSP_GEMM_COLS + 0];
        n = desc[g * TNSP_GEMM_COLS + 1];
        k = desc[g * TNSP_GEMM_COLS + 2];
        if (t < m * n) break;
        t -= m * n;
    }
    if (g >= ng) return;
    const int64_t* d = desc + g * TNSP_GEMM_COLS;
    const int64_t a_off = d[3], b_off = d[4], c_off = d[5], flags = d[6];
    const double alpha = (double)d[7];
    const int64_t i = t / n, j = t - i * n;
    const int64_t a0 = (flags & 1) ? i : i * k, as = (flags & 1) ? m : 1;
    const int64_t b0 = (flags & 2) ? j * k : j, bs = (flags & 2) ? 1 : n;
    for (int bi = blockIdx.y; bi < nb; bi += gridDim.y) {
        const double* A = a + (int64_t)bi * abs_ + a_off + a0;
        const double* B = b + (int64_t)bi * bbs + b_off + b0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int64_t kk = threadIdx.x;
        for (; kk + 768 < k; kk += 1024) {
            s0 += __ldg(A + kk * as) * __ldg(B + kk * bs);
            s1 += __ldg(A + (kk + 256) * as) * __ldg(B + (kk + 256) * bs);
            s2 += __ldg(A + (kk + 512) * as) * __ldg(B + (kk + 512) * bs);
            s3 += __ldg(A + (kk + 768) * as) * __ldg(B + (kk + 768) * bs);
        }
        for (; kk < k; kk += 256) s0 += __ldg(A + kk * as) * __ldg(B + kk * bs);
        const double tot = block_sum((s0 + s1) + (s2 + s3), red);
        if (threadIdx.x == 0) c[(int64_t)bi * cbs + c_off + t] = alpha * tot;
    }
}

// ------------------------------------------------------------------------------------------------
// Streaming kernel for the boundary-MPS contractions of the larger lattices (cfg2: m = 216 .. 7776,
// n, k = 6 .. 216, one GEMM per chain): these are HBM bound (intensity <= n/12 flop/B), so the kernel is
// organised around memory flow, not around the tensor pipe:
//   * CTA = 4 warps stacked along M (BM = 64 rows), every warp owns 16 rows x BN = 8*NT columns, NT chosen per
//     launch so that BN covers n with little padding (n = 36 -> NT = 5);
//   * operand slabs (BK = 16) go global -> shared with 8-byte cp.async (LDGSTS, zero-fill out of range),
//     double buffered, so no registers are tied up by loads and 4-5 CTAs stay resident per SM;
//   * DMMA m8n8k4 on padded (conflict-free) shared tiles; C is written straight from the fragments
//     (each store instruction covers whole 32-byte sectors).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int SBM = 64, SBK = 16, SLDA = SBK + 4;

template <int NT>
__global__ void __launch_bounds__(128, 4) gemm_stream_kernel(const int64_t* __restrict__ desc, int ng, const double* __restrict__ a,
                                                             int64_t abs_, const double* __restrict__ b, int64_t bbs,
                                                             double* __restrict__ c, int64_t cbs, int nb) {
    constexpr int BN = 8 * NT, LDB = BN + 4;
    constexpr int A_PER = SBM * SBK / 128;                 // 8
    constexpr int B_ELEMS = SBK * BN, B_PER = (B_ELEMS + 127) / 128;
    __shared__ double As[2][SBM * SLDA];
    __shared__ double Bs[2][SBK * LDB];

    int64_t t = blockIdx.x;
    int g = 0;
    int m = 0, n = 0, k = 0, tiles_n = 0;
    for (; g < ng; ++g) {
        m = (int)desc[g * TNSP_GEMM_COLS + 0];
        n = (int)desc[g * TNSP_GEMM_COLS + 1];
        k = (int)desc[g * TNSP_GEMM_COLS + 2];
        tiles_n = (n + BN - 1) / BN;
        const int64_t tiles = (int64_t)((m + SBM - 1) / SBM) * tiles_n;
        if (t < tiles) break;
        t -= tiles;
    }
    if (g >= ng) return;
    const int64_t* d = desc + g * TNSP_GEMM_COLS;
    const int64_t a_off = d[3], b_off = d[4], c_off = d[5], flags = d[6];
    const double alpha = (double)d[7];
    const bool a_km = flags & 1, b_nk = flags & 2;
    const int row0 = (int)(t / tiles_n) * SBM, col0 = (int)(t % tiles_n) * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp * 16;

    // per-thread slab element coordinates (fixed for the whole K loop)
    int ar, ak, a_du;            // A element u: row ar + (a_km ? 0 : u * 8), k index ak + (a_km ? u * 2 : 0)
    if (a_km) { ar = tid & 63; ak = tid >> 6; } else { ak = tid & 15; ar = tid >> 4; }
    (void)a_du;
    int bk, bc;                  // B element u: e = tid + u * 128
    const int nslab = (k + SBK - 1) / SBK;

    for (int bi = blockIdx.y; bi < nb; bi += gridDim.y) {
        const double* A = a + (int64_t)bi * abs_ + a_off;
        const double* B = b + (int64_t)bi * bbs + b_off;
        double* C = c + (int64_t)bi * cbs + c_off;

        auto issue = [&](int slab, int buf) {
            const int k0 = slab * SBK;
#pragma unroll
            for (int u = 0; u < A_PER; ++u) {
                const int r = a_km ? ar : ar + u * 8;
                const int kk = a_km ? ak + u * 2 : ak;
                const int gr = row0 + r, gk = k0 + kk;
                const bool ok = gr < m && gk < k;
                const double* src = ok ? (a_km ? A + (int64_t)gk * m + gr : A + (int64_t)gr * k + gk) : A;
                cp_async8(&As[buf][r * SLDA + kk], src, ok);
            }
#pragma unroll
            for (int u = 0; u < B_PER; ++u) {
                const int e = tid + u * 128;
                if (B_ELEMS % 128 == 0 || e < B_ELEMS) {
                    if (b_nk) { bk = e % SBK; bc = e / SBK; } else { bc = e % BN; bk = e / BN; }
                    const int gk = k0 + bk, gc = col0 + bc;
                    const bool ok = gk < k && gc < n;
                    const double* src = ok ? (b_nk ? B + (int64_t)gc * k + gk : B + (int64_t)gk * n + gc) : B;
                    cp_async8(&Bs[buf][bk * LDB + bc], src, ok);
                }
            }
            cp_async_commit();
        };

        double acc[2][NT][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        issue(0, 0);
        for (int sl = 0; sl < nslab; ++sl) {
            const int buf = sl & 1;
            if (sl + 1 < nslab) { issue(sl + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            __syncthreads();
            const double* as = As[buf];
            const double* bs = Bs[buf];
#pragma unroll
            for (int kk = 0; kk < SBK; kk += 4) {
                double fa[2], fb[NT];
#pragma unroll
                for (int i = 0; i < 2; ++i) fa[i] = as[(wm + i * 8 + gid) * SLDA + kk + tig];
#pragma unroll
                for (int j = 0; j < NT; ++j) fb[j] = bs[(kk + tig) * LDB + j * 8 + gid];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
            }
            __syncthreads();   // this buffer is refilled by the issue of the next iteration
        }
        const bool vec2 = ((n & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = row0 + wm + i * 8 + gid;
            if (r >= m) continue;
            double* crow = C + (int64_t)r * n;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int cc = col0 + j * 8 + 2 * tig;
                if (vec2 && cc + 1 < n) {
                    *reinterpret_cast<double2*>(crow + cc) = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
                } else {
                    if (cc < n) crow[cc] = alpha * acc[i][j][0];
                    if (cc + 1 < n) crow[cc + 1] = alpha * acc[i][j][1];
                }
            }
        }
    }
}

template <int NT>
static int launch_stream(const int64_t* desc, int ng, const int64_t* dh, const double* a, int64_t abs_, const double* b, int64_t bbs,
                         double* c, int64_t cbs, int nb, cudaStream_t st) {
    constexpr int BN = 8 * NT;
    int64_t tiles = 0;
    for (int g = 0; g < ng; ++g) {
        const int64_t m = dh[g * TNSP_GEMM_COLS], n = dh[g * TNSP_GEMM_COLS + 1];
        tiles += ((m + SBM - 1) / SBM) * ((n + BN - 1) / BN);
    }
    if (tiles == 0) return 0;
    const int gy = nb > 65535 ? 65535 : nb;
    gemm_stream_kernel<NT><<<dim3((unsigned)tiles, gy), 128, 0, st>>>(desc, ng, a, abs_, b, bbs, c, cbs, nb);
    return check_launch("tnsp_gemm_grouped_f64(stream)");
}

template <int TM, int TN>
static int launch_gemm(const int64_t* desc, int ng, const int64_t* dh, const double* a, int64_t abs_, const double* b, int64_t bbs,
                       double* c, int64_t cbs, int nb, cudaStream_t st) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    int64_t tiles = 0;
    for (int g = 0; g < ng; ++g) {
        const int64_t m = dh[g * TNSP_GEMM_COLS], n = dh[g * TNSP_GEMM_COLS + 1];
        tiles += ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
    }
    if (tiles == 0) return 0;
    const int gy = nb > 65535 ? 65535 : nb;
    gemm_grouped_kernel<TM, TN><<<dim3((unsigned)tiles, gy), 128, 0, st>>>(desc, ng, a, abs_, b, bbs, c, cbs, nb);
    return check_launch("tnsp_gemm_grouped_f64");
}

}  // namespace tnsp

using namespace tnsp;

extern "C" int tnsp_gemm_grouped_f64(const int64_t* desc, int ng, const int64_t* desc_host, const double* a, int64_t abs_,
                                     const double* b, int64_t bbs, double* c, int64_t cbs, int nb, void* stream) {
    if (ng == 0 || nb == 0) return 0;
    int64_t mmax = 0, nmax = 0, kmin = INT64_MAX, outputs = 0;
    for (int g = 0; g < ng; ++g) {
        const int64_t m = desc_host[g * TNSP_GEMM_COLS], n = desc_host[g * TNSP_GEMM_COLS + 1], k = desc_host[g * TNSP_GEMM_COLS + 2];
        if (m > mmax) mmax = m;
        if (n > nmax) nmax = n;
        if (m * n > 0 && k < kmin) kmin = k;
        outputs += m * n;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (outputs > 0 && kmin >= 1024 && outputs * 8 <= kmin && outputs <= 4096) {
        // few outputs, long contraction: one CTA per output element
        int64_t gy = nb;
        const int64_t want = (8 * kSMs + outputs - 1) / outputs;
        if (gy > want) gy = want;
        if (gy > 65535) gy = 65535;
        if (gy < 1) gy = 1;
        gemm_dot_kernel<<<dim3((unsigned)outputs, (unsigned)gy), 256, 0, st>>>(desc, ng, a, abs_, b, bbs, c, cbs, nb);
        return check_launch("tnsp_gemm_grouped_f64(dot)");
    }
    if (mmax >= 48) {
        // n-tile count with the least padding of the widest descriptor (prefer wide tiles: fewer re-reads of A)
        int best = 8;
        if (nmax <= 64) best = (int)((nmax + 7) / 8);
        else {
            int64_t best_waste = INT64_MAX;
            for (int nt = 8; nt >= 5; --nt) {
                const int64_t bn = 8 * nt, waste = ((nmax + bn - 1) / bn) * bn - nmax;
                if (waste < best_waste) { best_waste = waste; best = nt; }
            }
        }
        switch (best) {
            case 1: return launch_stream<1>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
            case 2: return launch_stream<2>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
            case 3: return launch_stream<3>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
            case 4: return launch_stream<4>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
            case 5: return launch_stream<5>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
            case 6: return launch_stream<6>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
            case 7: return launch_stream<7>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
            default: return launch_stream<8>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
        }
    }
    const bool big_m = mmax > 32, big_n = nmax > 32, mid_m = mmax > 16, mid_n = nmax > 16;
    if (big_m && big_n) return launch_gemm<4, 4>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
    if (big_m) return mid_n ? launch_gemm<4, 2>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st)
                            : launch_gemm<4, 1>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
    if (big_n) return mid_m ? launch_gemm<2, 4>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st)
                            : launch_gemm<1, 4>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
    if (mid_m && mid_n) return launch_gemm<2, 2>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
    if (mid_m) return launch_gemm<2, 1>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
    if (mid_n) return launch_gemm<1, 2>(desc, ng, desc_host, a, abs_, b, bbs, c, cbs, nb, st);
    {
        int64_t gy = (nb + 3) / 4;
        if (gy > 65535) gy = 65535;
        gemm_small_kernel<<<dim3(ng, (unsigned)gy), 128, 0, st>>>(desc, ng, a, abs_, b, bbs, c, cbs, nb);
        return check_launch("tnsp_gemm_grouped_f64(small)");
    }
}

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
    int64_t mmax = 0, nmax = 0;
    for (int g = 0; g < ng; ++g) {
        if (desc_host[g * TNSP_GEMM_COLS] > mmax) mmax = desc_host[g * TNSP_GEMM_COLS];
        if (desc_host[g * TNSP_GEMM_COLS + 1] > nmax) nmax = desc_host[g * TNSP_GEMM_COLS + 1];
    }
    cudaStream_t st = (cudaStream_t)stream;
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

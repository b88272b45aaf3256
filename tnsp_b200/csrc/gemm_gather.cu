// K2g: contraction of two dense(-embedded) tensors WITHOUT materialising the merged / transposed operands.
//
// The reference's contract (contract.hpp:622-857) first runs edge_operator on both operands (merge the free
// and the common edges, transpose to a matrix: two full read+write passes over HBM) and then one ?gemm_.
// For a dense tensor the merged-matrix element (r, kk) of an operand sits at
//        base + row_off[r] + col_off[kk]
// in the ORIGINAL tensor (both offsets are sums of index * stride over the edges of the group), so the GEMM
// can gather its tiles in place: the host planner builds the four int32 offset tables once per plan
// (tnsp_b200/TAT/plan.py::contract_plan) and this kernel feeds its cp.async pipeline through them.
//
// Organisation (the shapes of the boundary-MPS path are tall and skinny, m = 216..7776, n, k = 6..216, one GEMM
// per Markov chain, i.e. HBM/L2 bound except for k = n = 216):
//   * persistent CTAs (4 warps, 64 x 8*NT tile), each walks its share of the (chain, m-tile, n-tile) items;
//   * ONE cp.async pipeline (STAGES K-slabs of 16) runs across item boundaries, so the memory system never
//     drains between tiles -- the old per-tile kernel spent most of its life in prologue / epilogue latency;
//   * one block barrier per slab; DMMA m8n8k4 (the only FP64 tensor instruction of sm_100a: DMMA.8x8x4) on padded,
//     conflict-free shared tiles; C is written straight from the accumulator fragments as 16-byte stores.
//
// Roofline: FP64 tensor pipe when min(n, k) >~ 128, else HBM; algorithmic flops 2mnk, bytes 8(mk + kn + mn) per chain.
#include "common.cuh"

namespace tnsp {

namespace {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int GBM = 64, GBK = 16, GLDA = GBK + 4;
constexpr int GSTAGES = 3;

}  // namespace

// tab = row_off_A[m] | col_off_A[k] | row_off_B[k] | col_off_B[n]   (int32, element offsets)
// flags bit0: consecutive kk of A are the memory-contiguous direction (else consecutive r)
//       bit1: consecutive c  of B are the memory-contiguous direction (else consecutive kk)
template <int NT>
__global__ void __launch_bounds__(128, 3) gemm_gather_kernel(const int* __restrict__ tab, int m, int n, int k, int flags, double alpha,
                                                             const double* __restrict__ a, int64_t abs_, const double* __restrict__ b,
                                                             int64_t bbs, double* __restrict__ c, int64_t cbs, int nb, int tiles_m,
                                                             int tiles_n) {
    constexpr int BN = 8 * NT, LDB = BN + 4;
    constexpr int A_PER = GBM * GBK / 128;                 // 8
    constexpr int B_ELEMS = GBK * BN, B_PER = (B_ELEMS + 127) / 128;
    constexpr int A_STAGE = GBM * GLDA, B_STAGE = GBK * LDB;
    extern __shared__ __align__(16) double gsm[];
    double* As = gsm;                                      // [GSTAGES][A_STAGE]
    double* Bs = gsm + GSTAGES * A_STAGE;                  // [GSTAGES][B_STAGE]

    const int* aro = tab;
    const int* aco = tab + m;
    const int* bro = tab + m + k;
    const int* bco = tab + m + 2 * k;
    const bool a_fast_k = flags & 1, b_fast_n = flags & 2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp * 16;
    const int tiles = tiles_m * tiles_n;
    const int64_t total_items = (int64_t)nb * tiles;
    if ((int64_t)blockIdx.x >= total_items) return;
    const int n_my = (int)((total_items - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int nslab = (k + GBK - 1) / GBK;
    const int64_t total_q = (int64_t)n_my * nslab;

    // element coordinates of this thread inside a slab (fixed for the whole kernel)
    int ar, ak;
    if (a_fast_k) { ak = tid & 15; ar = tid >> 4; } else { ar = tid & 63; ak = tid >> 6; }
    int bk[B_PER], bc[B_PER];
#pragma unroll
    for (int u = 0; u < B_PER; ++u) {
        const int e = tid + u * 128;
        if (b_fast_n) { bc[u] = e % BN; bk[u] = e / BN; } else { bk[u] = e % GBK; bc[u] = e / GBK; }
    }

    // state of the issue cursor's current item: operand bases and the offsets that do not change from slab to slab
    // (-1 marks an out-of-range row / column, copied as zero fill)
    const double* iA = a;
    const double* iB = b;
    int a_off[A_PER];       // a_fast_k: row offsets of this thread's 8 rows; else a_off[0] = offset of its single row
    int b_off[B_PER];       // column offsets of this thread's B elements
    auto begin_item = [&](int it) {
        const int64_t t = blockIdx.x + (int64_t)it * gridDim.x;
        const int bi = (int)(t / tiles);
        const int rem = (int)(t - (int64_t)bi * tiles);
        const int row0 = (rem / tiles_n) * GBM, col0 = (rem % tiles_n) * BN;
        iA = a + (int64_t)bi * abs_;
        iB = b + (int64_t)bi * bbs;
        if (a_fast_k) {
#pragma unroll
            for (int u = 0; u < A_PER; ++u) { const int gr = row0 + ar + u * 8; a_off[u] = gr < m ? __ldg(aro + gr) : -1; }
        } else {
            const int gr = row0 + ar;
            a_off[0] = gr < m ? __ldg(aro + gr) : -1;
        }
#pragma unroll
        for (int u = 0; u < B_PER; ++u) { const int gc = col0 + bc[u]; b_off[u] = gc < n ? __ldg(bco + gc) : -1; }
    };
    auto issue = [&](int slab, int buf) {
        const int k0 = slab * GBK;
        double* as = As + buf * A_STAGE;
        double* bs = Bs + buf * B_STAGE;
        if (a_fast_k) {
            const int gk = k0 + ak;
            const int co = gk < k ? __ldg(aco + gk) : -1;
#pragma unroll
            for (int u = 0; u < A_PER; ++u) {
                const bool ok = (co | a_off[u]) >= 0;
                cp_async8(&as[(ar + u * 8) * GLDA + ak], ok ? iA + (a_off[u] + co) : iA, ok);
            }
        } else {
            const int ro = a_off[0];
#pragma unroll
            for (int u = 0; u < A_PER; ++u) {
                const int kk = ak + u * 2, gk = k0 + kk;
                const int co = gk < k ? __ldg(aco + gk) : -1;
                const bool ok = (co | ro) >= 0;
                cp_async8(&as[ar * GLDA + kk], ok ? iA + (ro + co) : iA, ok);
            }
        }
#pragma unroll
        for (int u = 0; u < B_PER; ++u) {
            if (B_ELEMS % 128 == 0 || tid + u * 128 < B_ELEMS) {
                const int gk = k0 + bk[u];
                const int ro = gk < k ? __ldg(bro + gk) : -1;
                const bool ok = (ro | b_off[u]) >= 0;
                cp_async8(&bs[bk[u] * LDB + bc[u]], ok ? iB + (ro + b_off[u]) : iB, ok);
            }
        }
    };

    double acc[2][NT][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue: slabs 0 .. GSTAGES-2 of the flattened (item, slab) sequence
    int i_it = 0, i_sl = 0;     // issue cursor
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
        if (i_it < n_my) {
            if (i_sl == 0) begin_item(i_it);
            issue(i_sl, s);
            if (++i_sl == nslab) { i_sl = 0; ++i_it; }
        }
        cp_async_commit();
    }
    int c_it = 0, c_sl = 0;     // compute cursor
    for (int64_t q = 0; q < total_q; ++q) {
        const int buf = (int)(q % GSTAGES);
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();        // slab q has landed for every thread; everyone is done with slab q-1
        if (i_it < n_my) {
            if (i_sl == 0) begin_item(i_it);
            issue(i_sl, (int)((q + GSTAGES - 1) % GSTAGES));
            if (++i_sl == nslab) { i_sl = 0; ++i_it; }
        }
        cp_async_commit();
        const double* as = As + buf * A_STAGE;
        const double* bs = Bs + buf * B_STAGE;
        const int kvalid = k - c_sl * GBK;      // k steps of this slab that hold data (the last slab may be short)
#pragma unroll
        for (int kk = 0; kk < GBK; kk += 4) {
            if (kk >= kvalid) break;
            double fa[2], fb[NT];
#pragma unroll
            for (int i = 0; i < 2; ++i) fa[i] = as[(wm + i * 8 + gid) * GLDA + kk + tig];
#pragma unroll
            for (int j = 0; j < NT; ++j) fb[j] = bs[(kk + tig) * LDB + j * 8 + gid];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
        if (++c_sl == nslab) {
            // epilogue of item c_it
            const int64_t t = blockIdx.x + (int64_t)c_it * gridDim.x;
            const int bi = (int)(t / tiles);
            const int rem = (int)(t - (int64_t)bi * tiles);
            const int row0 = (rem / tiles_n) * GBM, col0 = (rem % tiles_n) * BN;
            double* C = c + (int64_t)bi * cbs;
            const bool vec2 = ((n & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = row0 + wm + i * 8 + gid;
                double* crow = C + (int64_t)r * n;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int cc = col0 + j * 8 + 2 * tig;
                    if (r < m) {
                        if (vec2 && cc + 1 < n) {
                            *reinterpret_cast<double2*>(crow + cc) = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
                        } else {
                            if (cc < n) crow[cc] = alpha * acc[i][j][0];
                            if (cc + 1 < n) crow[cc + 1] = alpha * acc[i][j][1];
                        }
                    }
                    acc[i][j][0] = acc[i][j][1] = 0.0;
                }
            }
            c_sl = 0;
            ++c_it;
        }
    }
    cp_async_wait<0>();
}

template <int NT>
static int launch_gather(const int* tab, int m, int n, int k, int flags, double alpha, const double* a, int64_t abs_, const double* b,
                         int64_t bbs, double* c, int64_t cbs, int nb, cudaStream_t st) {
    constexpr int BN = 8 * NT, LDB = BN + 4;
    constexpr int smem = GSTAGES * (GBM * GLDA + GBK * LDB) * 8;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_gather_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    const int tiles_m = (m + GBM - 1) / GBM, tiles_n = (n + BN - 1) / BN;
    const int64_t items = (int64_t)nb * tiles_m * tiles_n;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gemm_gather_kernel<NT>, 128, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)kSMs * per_sm;
    if (grid > items) grid = items;
    gemm_gather_kernel<NT><<<(unsigned)grid, 128, smem, st>>>(tab, m, n, k, flags, alpha, a, abs_, b, bbs, c, cbs, nb, tiles_m, tiles_n);
    return check_launch("tnsp_gemm_gather_f64");
}


// ------------------------------------------------------------------------------------------------
// Row-stream kernel for the HBM-bound class (k * n small enough for the whole B of a chain to sit in shared
// memory: 1296x36x36, 7776x36x36, 1296x36x216, 1296x216x6, 216x36x36 ... of the boundary-MPS path).
//   * a CTA (4 warps) owns one chain and a range of 16-row strips; B (gathered through its offset tables, zero
//     padded) is staged in shared memory ONCE per CTA;
//   * the A operand never touches shared memory: every lane loads its DMMA fragment elements straight from
//     global memory (a warp-wide load covers 8 rows x 32 B or 4 k x 64 B: whole sectors in either layout),
//     chunks of 32 k double-buffered in registers so the loads of the next chunk / strip fly under the DMMAs;
//   * C goes straight from the accumulator fragments to global memory (16-byte stores).
// Per 16 x 36 x 36 strip a warp issues ~190 instructions (18 loads, 90 DMMA, 45 LDS, 10 stores) where the tiled
// kernels needed ~1300-2000: the kernel is limited by HBM, not by instruction issue.
// ------------------------------------------------------------------------------------------------
// k-steps (of 4) per register chunk: 8 for the narrow tiles; 4 for NT >= 6, whose 4 * NT accumulator registers would
// otherwise push the kernel past 128 registers (two 256-thread CTAs per SM need <= 128)
// zero-fragment skipping of the row-stream GEMM (tnsp_gemm_skip_zero_fragments): off for dense models (the tests cost 45 % on
// operands without zeros), switched on by tetragono.dense_embedding for block-sparse ones
static int g_skip_zero_fragments = 0;
constexpr int kSkipMinK = 96;

template <int NT> struct RowstreamCfg { static constexpr int RKS = NT >= 6 ? 4 : 8; static constexpr int MINB = 2; };

// grid (x: slices of the n-passes, y: chain, z: strip ranges of a chain) -- the slices of one strip range are launched
// next to each other so that their re-reads of the same A rows hit L2; blockDim 128 or 256
// SKIP: 0 dense (branch free); 1 fine tests (k-step x tile from the B map, 8-row block of A by warp vote); 2 coarse tests (k-step x
// pass, 8-row block; the NT instructions behind a test stay one independent group); 3 / 4 = 1 / 2 + the A fragments of k-steps no tile
// of the pass needs are not fetched
template <int NT, int SKIP>
__global__ void __launch_bounds__(256, RowstreamCfg<NT>::MINB) gemm_rowstream_kernel(const int* __restrict__ tab, int m, int n, int k, double alpha,
                                                             const double* __restrict__ a, int64_t abs_, const double* __restrict__ b,
                                                             int64_t bbs, double* __restrict__ c, int64_t cbs, int npass_total,
                                                             int passes_per_cta, int strips_per_cta) {
    constexpr int RKS = RowstreamCfg<NT>::RKS;
    extern __shared__ __align__(16) double gsm[];
    double* Bs = gsm;                                   // [kpad][ldb], zero padded
    const int* aro = tab;
    const int* aco = tab + m;
    const int* bro = tab + m + k;
    const int* bco = tab + m + 2 * k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int bi = blockIdx.y;
    const double* A = a + (int64_t)bi * abs_;
    const double* B = b + (int64_t)bi * bbs;
    double* C = c + (int64_t)bi * cbs;
    const int kpad = (k + 3) & ~3;
    const int pass_begin = blockIdx.x * passes_per_cta;
    const int npass = min(passes_per_cta, npass_total - pass_begin);
    const int ncols = passes_per_cta * 8 * NT;
    const int ldb = ncols + 4;
    const unsigned long long* bmask_ld = reinterpret_cast<const unsigned long long*>(Bs + (size_t)kpad * ldb);   // zero-fragment map, filled below
    const int colbase = pass_begin * 8 * NT;

    // stage this CTA's slice of B: element (kk, cc) <- B[bro[kk] + bco[colbase + cc]]
    for (int e = tid; e < kpad * ncols; e += nthreads) {
        const int kk = e / ncols, cc = e - kk * ncols;
        const bool ok = kk < k && colbase + cc < n;
        cp_async8(&Bs[kk * ldb + cc], ok ? B + (__ldg(bro + kk) + __ldg(bco + colbase + cc)) : B, ok);
    }
    cp_async_commit();

    const int nstrips = (m + 15) >> 4;
    const int s_begin = blockIdx.z * strips_per_cta;
    const int s_end = min(nstrips, s_begin + strips_per_cta);
    const int nchunk = (kpad / 4 + RKS - 1) / RKS;
    const int avail = s_end - s_begin - warp;
    const int my_strips = avail > 0 ? (avail + nwarps - 1) / nwarps : 0;   // strips s_begin + warp, + nwarps, ...
    const int total = my_strips * npass * nchunk;

    double fa[2][2][RKS];       // [buffer][row block][k-step]
    int ro0 = -1, ro1 = -1;     // row offsets of the strip the NEXT load belongs to
    auto load_chunk = [&](int seq, int buf) {
        const int ch = seq % nchunk;
        const int strip = s_begin + warp + nwarps * (seq / (nchunk * npass));
        if (ch == 0) {
            const int r0 = strip * 16 + gid, r1 = r0 + 8;
            ro0 = r0 < m ? __ldg(aro + r0) : -1;
            ro1 = r1 < m ? __ldg(aro + r1) : -1;
        }
        unsigned long long want = ~0ull;                 // bit ks: some tile of this chunk's pass needs k-step ks
        if constexpr (SKIP >= 3) {
            const int ps = (seq / nchunk) % npass;
            want = 0ull;
#pragma unroll
            for (int ks = 0; ks < RKS; ++ks) {
                const int kstep = ch * RKS + ks;
                if (kstep < kpad / 4 && ((unsigned)(bmask_ld[kstep] >> (ps * NT)) & ((1u << NT) - 1u)) != 0u) want |= 1ull << ks;
            }
        }
#pragma unroll
        for (int ks = 0; ks < RKS; ++ks) {
            const int kk = (ch * RKS + ks) * 4 + tig;
            const int co = (kk < k && ((want >> ks) & 1ull)) ? __ldg(aco + kk) : -1;
            fa[buf][0][ks] = (co | ro0) >= 0 ? __ldg(A + (ro0 + co)) : 0.0;
            fa[buf][1][ks] = (co | ro1) >= 0 ? __ldg(A + (ro1 + co)) : 0.0;
        }
    };

    if (SKIP < 3 && total > 0) load_chunk(0, 0);      // overlaps the staging of B
    cp_async_wait<0>();
    __syncthreads();

    // Zero-fragment map of the staged B slice: bit t of bmask[kstep] is set when the 4 x 8 fragment (k-step, n-tile t) holds a
    // non-zero.  Symmetric tensors in the charge-dense embedding are block sparse: at cfg2 only 23 % of the B fragments of the
    // 1296 x 216 x 216 contraction are non-zero (13 % when the A fragment's emptiness is counted too), and a DMMA on an
    // all-zero fragment cannot change the accumulator -- it is skipped.  The branch is uniform over the CTA (B) / warp (A).
    // SKIP is a template parameter: the dense instantiation is the branch-free kernel, instruction for instruction (a run-time
    // switch inside the unrolled loop cost the k = 36 shapes 38 %).  Measured at cfg2 (2368 chains): 1296 x 216 x 216 10.9 -> 8.1 ms.
    unsigned long long* bmask = reinterpret_cast<unsigned long long*>(Bs + (size_t)kpad * ldb);
    const int ntiles = ncols >> 3;                      // <= 64 whenever the host selects SKIP
    if constexpr (SKIP != 0) {
        for (int ksb = tid; ksb < kpad / 4; ksb += nthreads) {
            unsigned long long word = 0ull;
            for (int t = 0; t < ntiles; ++t) {
                bool nz = false;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const double* row = Bs + (ksb * 4 + r) * ldb + t * 8;
#pragma unroll
                    for (int cc = 0; cc < 8; cc += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(row + cc);
                        nz |= (v.x != 0.0) | (v.y != 0.0);
                    }
                }
                if (nz) word |= 1ull << t;
            }
            bmask[ksb] = word;
        }
        __syncthreads();
    }
    if (SKIP >= 3 && total > 0) load_chunk(0, 0);

    double acc[2][NT][2];
    int seq = 0;
    for (int si = 0; si < my_strips; ++si) {
        const int strip = s_begin + warp + nwarps * si;
        for (int pass = 0; pass < npass; ++pass) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            const double* bcol = Bs + pass * (8 * NT) + gid + tig * ldb;
            for (int ch = 0; ch < nchunk; ++ch, ++seq) {
                const double* bch = bcol + ch * (RKS * 4) * ldb;
                const int ksteps = min(RKS, kpad / 4 - ch * RKS);
                // the two register buffers alternate; both branches are fully unrolled so that fa[] stays in registers
                if ((seq & 1) == 0) {
                    if (seq + 1 < total) load_chunk(seq + 1, 1);
#pragma unroll
                    for (int ks = 0; ks < RKS; ++ks) {
                        if (ks < ksteps) {
                            if constexpr (SKIP == 1 || SKIP == 3) {
                                const unsigned tiles = (unsigned)(bmask[ch * RKS + ks] >> (pass * NT)) & ((1u << NT) - 1u);
                                if (tiles == 0u) continue;
                                const bool a0 = __any_sync(0xffffffffu, fa[0][0][ks] != 0.0);
                                const bool a1 = __any_sync(0xffffffffu, fa[0][1][ks] != 0.0);
                                if (!(a0 | a1)) continue;
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    if (tiles & (1u << j)) {
                                        const double fbj = bch[ks * 4 * ldb + j * 8];
                                        if (a0) dmma884(acc[0][j][0], acc[0][j][1], fa[0][0][ks], fbj);
                                        if (a1) dmma884(acc[1][j][0], acc[1][j][1], fa[0][1][ks], fbj);
                                    }
                                }
                            } else if constexpr (SKIP == 2 || SKIP == 4) {
                                if (((unsigned)(bmask[ch * RKS + ks] >> (pass * NT)) & ((1u << NT) - 1u)) == 0u) continue;
                                const bool a0 = __any_sync(0xffffffffu, fa[0][0][ks] != 0.0);
                                const bool a1 = __any_sync(0xffffffffu, fa[0][1][ks] != 0.0);
                                if (!(a0 | a1)) continue;
                                double fb[NT];
#pragma unroll
                                for (int j = 0; j < NT; ++j) fb[j] = bch[ks * 4 * ldb + j * 8];
                                if (a0) {
#pragma unroll
                                    for (int j = 0; j < NT; ++j) dmma884(acc[0][j][0], acc[0][j][1], fa[0][0][ks], fb[j]);
                                }
                                if (a1) {
#pragma unroll
                                    for (int j = 0; j < NT; ++j) dmma884(acc[1][j][0], acc[1][j][1], fa[0][1][ks], fb[j]);
                                }
                            } else {
                                double fb[NT];
#pragma unroll
                                for (int j = 0; j < NT; ++j) fb[j] = bch[ks * 4 * ldb + j * 8];
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    dmma884(acc[0][j][0], acc[0][j][1], fa[0][0][ks], fb[j]);
                                    dmma884(acc[1][j][0], acc[1][j][1], fa[0][1][ks], fb[j]);
                                }
                            }
                        }
                    }
                } else {
                    if (seq + 1 < total) load_chunk(seq + 1, 0);
#pragma unroll
                    for (int ks = 0; ks < RKS; ++ks) {
                        if (ks < ksteps) {
                            if constexpr (SKIP == 1 || SKIP == 3) {
                                const unsigned tiles = (unsigned)(bmask[ch * RKS + ks] >> (pass * NT)) & ((1u << NT) - 1u);
                                if (tiles == 0u) continue;
                                const bool a0 = __any_sync(0xffffffffu, fa[1][0][ks] != 0.0);
                                const bool a1 = __any_sync(0xffffffffu, fa[1][1][ks] != 0.0);
                                if (!(a0 | a1)) continue;
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    if (tiles & (1u << j)) {
                                        const double fbj = bch[ks * 4 * ldb + j * 8];
                                        if (a0) dmma884(acc[0][j][0], acc[0][j][1], fa[1][0][ks], fbj);
                                        if (a1) dmma884(acc[1][j][0], acc[1][j][1], fa[1][1][ks], fbj);
                                    }
                                }
                            } else if constexpr (SKIP == 2 || SKIP == 4) {
                                if (((unsigned)(bmask[ch * RKS + ks] >> (pass * NT)) & ((1u << NT) - 1u)) == 0u) continue;
                                const bool a0 = __any_sync(0xffffffffu, fa[1][0][ks] != 0.0);
                                const bool a1 = __any_sync(0xffffffffu, fa[1][1][ks] != 0.0);
                                if (!(a0 | a1)) continue;
                                double fb[NT];
#pragma unroll
                                for (int j = 0; j < NT; ++j) fb[j] = bch[ks * 4 * ldb + j * 8];
                                if (a0) {
#pragma unroll
                                    for (int j = 0; j < NT; ++j) dmma884(acc[0][j][0], acc[0][j][1], fa[1][0][ks], fb[j]);
                                }
                                if (a1) {
#pragma unroll
                                    for (int j = 0; j < NT; ++j) dmma884(acc[1][j][0], acc[1][j][1], fa[1][1][ks], fb[j]);
                                }
                            } else {
                                double fb[NT];
#pragma unroll
                                for (int j = 0; j < NT; ++j) fb[j] = bch[ks * 4 * ldb + j * 8];
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    dmma884(acc[0][j][0], acc[0][j][1], fa[1][0][ks], fb[j]);
                                    dmma884(acc[1][j][0], acc[1][j][1], fa[1][1][ks], fb[j]);
                                }
                            }
                        }
                    }
                }
            }
            const int col0 = colbase + pass * (8 * NT);
            const bool vec2 = ((n & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = strip * 16 + i * 8 + gid;
                if (r < m) {
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
    }
}

constexpr int64_t kRowstreamSmemMax = 104 * 1024;   // two CTAs per SM

// the zero-padded B slice of one CTA (passes_per_cta n-passes of 8 * NT columns + 4 pad, k rounded up to 4) must fit
// in kRowstreamSmemMax; wide B operands are split over blockIdx.z
static bool rowstream_shape(int64_t n, int64_t k, int& nt, int& npass, int& passes_per_cta, int64_t& smem) {
    if (n <= 64) { nt = (int)((n + 7) / 8); npass = 1; }
    else {
        int64_t best_waste = INT64_MAX;
        nt = 8;
        for (int t = 8; t >= 5; --t) {
            const int64_t bn = 8 * t, waste = ((n + bn - 1) / bn) * bn - n;
            if (waste < best_waste) { best_waste = waste; nt = t; }
        }
        npass = (int)((n + 8 * nt - 1) / (8 * nt));
    }
    const int64_t kpad = (k + 3) / 4 * 4;
    passes_per_cta = npass;
    while (passes_per_cta > 1 && kpad * (passes_per_cta * 8 * nt + 4) * 8 + (kpad / 4) * 8 > kRowstreamSmemMax) passes_per_cta = (passes_per_cta + 1) / 2;
    smem = kpad * ((int64_t)passes_per_cta * 8 * nt + 4) * 8 + (kpad / 4) * 8;     // B slice + zero-fragment map
    return smem <= kRowstreamSmemMax;
}

template <int NT>
static int launch_rowstream(const int* tab, int m, int n, int k, double alpha, const double* a, int64_t abs_, const double* b, int64_t bbs,
                            double* c, int64_t cbs, int nb, int npass, int passes_per_cta, int64_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_rowstream_kernel<NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowstreamSmemMax);
        if constexpr (NT >= 6) {
            cudaFuncSetAttribute(gemm_rowstream_kernel<NT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowstreamSmemMax);
            cudaFuncSetAttribute(gemm_rowstream_kernel<NT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowstreamSmemMax);
            cudaFuncSetAttribute(gemm_rowstream_kernel<NT, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowstreamSmemMax);
            cudaFuncSetAttribute(gemm_rowstream_kernel<NT, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowstreamSmemMax);
        }
        attr_set = true;
    }
    // a large B slice limits the CTAs per SM: use 8 warps per CTA then, so that enough loads stay in flight
    const int threads = smem > 40 * 1024 ? 256 : 128;
    const int warps = threads / 32;
    const int nstrips = (m + 15) / 16;
    // about 5 strips per warp amortise the staging of B; never more CTAs than needed to give every warp a strip
    int ctas = (nstrips + 5 * warps - 1) / (5 * warps);
    if (ctas < 1) ctas = 1;
    const int strips_per_cta = (nstrips + ctas - 1) / ctas;
    ctas = (nstrips + strips_per_cta - 1) / strips_per_cta;
    const int slices = (npass + passes_per_cta - 1) / passes_per_cta;
    if (nb > 65535) { set_error("tnsp_gemm_gather_f64: more than 65535 chains"); return 1; }
    // the fragment tests pay for themselves on the tensor-bound shapes (long k, wide n); the others keep the branch-free kernel
    const int mode = (NT >= 6 && k >= kSkipMinK && passes_per_cta * NT <= 64) ? g_skip_zero_fragments : 0;
#define TNSP_ROWSTREAM_LAUNCH(S)                                                                                                         \
    gemm_rowstream_kernel<NT, S><<<dim3(slices, nb, ctas), threads, smem, st>>>(tab, m, n, k, alpha, a, abs_, b, bbs, c, cbs, npass,    \
                                                                                passes_per_cta, strips_per_cta)
    if constexpr (NT >= 6) {
        switch (mode) {
            case 1: TNSP_ROWSTREAM_LAUNCH(1); break;
            case 2: TNSP_ROWSTREAM_LAUNCH(2); break;
            case 3: TNSP_ROWSTREAM_LAUNCH(3); break;
            case 4: TNSP_ROWSTREAM_LAUNCH(4); break;
            default: TNSP_ROWSTREAM_LAUNCH(0); break;
        }
    } else {
        TNSP_ROWSTREAM_LAUNCH(0);
    }
#undef TNSP_ROWSTREAM_LAUNCH
    return check_launch("tnsp_gemm_gather_f64(rowstream)");
}

}  // namespace tnsp

using namespace tnsp;

extern "C" int tnsp_gemm_skip_zero_fragments(int enable) {
    const int old = g_skip_zero_fragments;
    if (enable >= 0) g_skip_zero_fragments = enable > 4 ? 1 : enable;      // 1 .. 4 select the variant (see the kernel)
    return old;
}

extern "C" int tnsp_gemm_gather_f64(const int32_t* tab, int64_t m, int64_t n, int64_t k, int flags, double alpha, const double* a,
                                    int64_t abs_, const double* b, int64_t bbs, double* c, int64_t cbs, int nb, void* stream) {
    if (nb == 0 || m == 0 || n == 0) return 0;
    if (k == 0) { set_error("tnsp_gemm_gather_f64: k == 0 (the caller zero-fills)"); return 1; }
    if (m > INT32_MAX / 2 || n > INT32_MAX / 2 || k > INT32_MAX / 2) { set_error("tnsp_gemm_gather_f64: dimension too large"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    {
        int nt, npass, ldb;   // ldb: n-passes per CTA
        int64_t smem;
        if (rowstream_shape(n, k, nt, npass, ldb, smem)) {
            const int mi = (int)m, ni = (int)n, ki = (int)k;
            switch (nt) {
                case 1: return launch_rowstream<1>(tab, mi, ni, ki, alpha, a, abs_, b, bbs, c, cbs, nb, npass, ldb, smem, st);
                case 2: return launch_rowstream<2>(tab, mi, ni, ki, alpha, a, abs_, b, bbs, c, cbs, nb, npass, ldb, smem, st);
                case 3: return launch_rowstream<3>(tab, mi, ni, ki, alpha, a, abs_, b, bbs, c, cbs, nb, npass, ldb, smem, st);
                case 4: return launch_rowstream<4>(tab, mi, ni, ki, alpha, a, abs_, b, bbs, c, cbs, nb, npass, ldb, smem, st);
                case 5: return launch_rowstream<5>(tab, mi, ni, ki, alpha, a, abs_, b, bbs, c, cbs, nb, npass, ldb, smem, st);
                case 6: return launch_rowstream<6>(tab, mi, ni, ki, alpha, a, abs_, b, bbs, c, cbs, nb, npass, ldb, smem, st);
                case 7: return launch_rowstream<7>(tab, mi, ni, ki, alpha, a, abs_, b, bbs, c, cbs, nb, npass, ldb, smem, st);
                default: return launch_rowstream<8>(tab, mi, ni, ki, alpha, a, abs_, b, bbs, c, cbs, nb, npass, ldb, smem, st);
            }
        }
    }
    // n-tile width with the least padding (prefer wide tiles: fewer re-reads of A)
    int best = 8;
    if (n <= 64) best = (int)((n + 7) / 8);
    else {
        int64_t best_waste = INT64_MAX;
        for (int nt = 8; nt >= 5; --nt) {
            const int64_t bn = 8 * nt, waste = ((n + bn - 1) / bn) * bn - n;
            if (waste < best_waste) { best_waste = waste; best = nt; }
        }
    }
    const int mi = (int)m, ni = (int)n, ki = (int)k;
    switch (best) {
        case 1: return launch_gather<1>(tab, mi, ni, ki, flags, alpha, a, abs_, b, bbs, c, cbs, nb, st);
        case 2: return launch_gather<2>(tab, mi, ni, ki, flags, alpha, a, abs_, b, bbs, c, cbs, nb, st);
        case 3: return launch_gather<3>(tab, mi, ni, ki, flags, alpha, a, abs_, b, bbs, c, cbs, nb, st);
        case 4: return launch_gather<4>(tab, mi, ni, ki, flags, alpha, a, abs_, b, bbs, c, cbs, nb, st);
        case 5: return launch_gather<5>(tab, mi, ni, ki, flags, alpha, a, abs_, b, bbs, c, cbs, nb, st);
        case 6: return launch_gather<6>(tab, mi, ni, ki, flags, alpha, a, abs_, b, bbs, c, cbs, nb, st);
        case 7: return launch_gather<7>(tab, mi, ni, ki, flags, alpha, a, abs_, b, bbs, c, cbs, nb, st);
        default: return launch_gather<8>(tab, mi, ni, ki, flags, alpha, a, abs_, b, bbs, c, cbs, nb, st);
    }
}

// K3s/K4s: QR/LQ and Jacobi SVD with ON-DEVICE SECTOR DISCOVERY for the lock-step batch engine.
//
// In the charge-dense embedding every chain of a batch stores a symmetric
// tensor as a dense array with exact zeros outside its symmetry blocks; which rows / columns of the
// merged matrix form a block differs from chain to chain (it follows the sampled physical charges),
// so no host-side descriptor can name the blocks.  These kernels recover them from the zero pattern:
//
//   pass 1  one bit mask of the non-zero columns per row (ballot), kept in shared memory
//   pass 2  union-find over the columns: all columns sharing a row are merged; the resulting connected
//           components of the bipartite row/column graph are the reference's symmetry sectors
//           (contract.hpp:539-580 / qr.hpp:419-429 / svd.hpp:405-427) -- blocks need not be full
//           (triangular R factors and rank-deficient bonds leave staircase patterns inside a sector)
//   per sector: gather the compact p x q block into shared memory, factorise it there (Householder
//           QR as ?geqrf/?orgqr, qr.hpp:178-304; one-sided Jacobi as ?gesvd 'S','S', svd.hpp:104-211),
//           scatter the factors back into the dense layout.
//
// One CTA owns one chain's matrix and walks its sectors; the new bond index of sector s is the
// contiguous range [K_s, K_s + min(p_s, q_s)) (QR) or the global descending rank of the singular
// value (SVD), so the first `cut` bond indices are the greedy cross-sector cut of svd.hpp:429-481.
// Outputs must be zero-initialised by the caller (only sector blocks are written).
//
// Roofline: HBM; algorithmic bytes 8*(2mn + mk + kn) (QR), 8*(mn + mk + kn + k) (SVD) per chain.
#include <stdlib.h>

#include "common.cuh"
#include "ragged.cuh"
#include <algorithm>
#include <utility>
#include <vector>

namespace tnsp {

constexpr int kSecThreads = 256;
constexpr int kSecWarps = kSecThreads / 32;
constexpr int kSecSmemDoubles = 25 * 1024;   // 200 KiB working set per CTA
constexpr int kSecMaxDim = 8192;
constexpr int kWarpSectorMax = 40;   // sectors up to this many columns are diagonalised by one warp each, concurrently

struct SecMap {
    int* colkey;         // [n]
    int* rowkey;         // [m]
    uint16_t* colsec;    // [n] sector id or 0xFFFF
    uint16_t* rowsec;    // [m]
    uint16_t* collist;   // [n] columns grouped by sector, ascending inside a sector
    uint16_t* rowlist;   // [m]
    uint16_t* coltmp;    // [n]
    int* cstart;         // [S+1]
    int* rstart;         // [S+1]
    int* kstart;         // [S+1] prefix of min(p_s, q_s)
    int* repkey;         // [n] representative row of every first-non-zero column (discovery only)
    int S;
};

__host__ __device__ inline int64_t secmap_bytes(int64_t m, int64_t n) {
    const int64_t s = (m < n ? m : n) + 2;
    int64_t b = 4 * (2 * n + m) + 2 * 2 * (n + m) + 2 * n + 3 * 4 * s;
    return (b + 15) / 16 * 16;
}

__device__ inline void secmap_carve(SecMap& sm, unsigned char* base, int m, int n) {
    const int s = (m < n ? m : n) + 2;
    int* ip = reinterpret_cast<int*>(base);
    sm.colkey = ip; ip += n;
    sm.rowkey = ip; ip += m;
    sm.cstart = ip; ip += s;
    sm.rstart = ip; ip += s;
    sm.kstart = ip; ip += s;
    sm.repkey = ip; ip += n;
    uint16_t* hp = reinterpret_cast<uint16_t*>(ip);
    sm.colsec = hp; hp += n;
    sm.rowsec = hp; hp += m;
    sm.collist = hp; hp += n;
    sm.rowlist = hp; hp += m;
    sm.coltmp = hp;
}

// 64-bit read-only load that the compiler cannot fold into a predicate or reorder against its neighbours
__device__ __forceinline__ unsigned long long ldg_bits(const double* p) {
    unsigned long long v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// Discover the sectors of the m x n row-major matrix A = connected components of the bipartite row/column
// graph of its non-zeros.  The zero pattern is first condensed into one bit mask per row (shared memory,
// `masks`, m x nw words; may alias the later working set), then a lock-free union-find over the COLUMNS
// merges, row by row, all columns that share a row (hooking larger roots under smaller ones with atomicMin,
// repeated until a pass changes nothing).  Sector order = ascending smallest column of the component.
// All threads of the CTA call this; `mask_cap_words` is the room available for the masks.
__device__ __forceinline__ int uf_find(const volatile int* parent, int x) {
    int r = parent[x];
    while (true) { const int q = parent[r]; if (q == r) break; r = q; }
    return r;
}

// ro / co (optional): element (i, j) of the matrix is A[ro[i] + co[j]] (operand read in place from a tensor of any
// index order, see tnsp_qr_sectors_gather_f64); nullptr: plain row-major A[i * n + j].
__device__ void discover_sectors(SecMap& sm, const double* __restrict__ A, int m, int n, int* sh_flag, uint32_t* masks,
                                 int64_t mask_cap_words, const int* __restrict__ ro = nullptr, const int* __restrict__ co = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = (n + 31) >> 5;
    int* parent = sm.colkey;
    if ((int64_t)m * nw > mask_cap_words) {
        // no room for the masks: one dense sector holding every row and column
        for (int j = tid; j < n; j += kSecThreads) { sm.colsec[j] = 0; sm.collist[j] = (uint16_t)j; }
        for (int i = tid; i < m; i += kSecThreads) { sm.rowsec[i] = 0; sm.rowlist[i] = (uint16_t)i; }
        if (tid == 0) {
            sm.cstart[0] = 0; sm.cstart[1] = n; sm.rstart[0] = 0; sm.rstart[1] = m;
            sm.kstart[0] = 0; sm.kstart[1] = m < n ? m : n;
        }
        sm.S = 1;
        __syncthreads();
        return;
    }
    for (int j = tid; j < n; j += kSecThreads) { parent[j] = j; sm.repkey[j] = 0; }
    __syncthreads();
    // row masks (ballot per 32 columns), first non-zero column and population count of every row; the row with the
    // most columns among those starting at column f becomes the representative of f.
    // A warp handles two rows and eight 32-column words per step: 16 loads per lane, all UNCONDITIONAL (column index
    // clamped, validity applied to the ballot) and compared as integers afterwards -- with conditional loads the
    // compiler turned every value into a predicate right away, one load in flight per warp (SASS: LDG, DSETP, LDG, ...;
    // ncu: 0.96 TB/s, 70 % of the stall samples on the first use of a loaded value)
    for (int i = warp; i < m; i += 2 * kSecWarps) {
        const int i2 = i + kSecWarps;
        const bool has2 = i2 < m;
        const double* row0 = A + (ro ? (int64_t)__ldg(ro + i) : (int64_t)i * n);
        const double* row1 = has2 ? A + (ro ? (int64_t)__ldg(ro + i2) : (int64_t)i2 * n) : row0;
        int first0 = n, pc0 = 0, first1 = n, pc1 = 0;
        for (int w0 = 0; w0 < nw; w0 += 8) {
            int cj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = min(((w0 + u) << 5) + lane, n - 1);
                cj[u] = co ? __ldg(co + j) : j;
            }
            unsigned long long v0[8], v1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                v0[u] = ldg_bits(row0 + cj[u]);
                v1[u] = ldg_bits(row1 + cj[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool valid = ((w0 + u) << 5) + lane < n;
                // (bits << 1) != 0: non-zero, -0.0 counts as zero like `v != 0.0`
                const unsigned b0 = __ballot_sync(0xffffffffu, valid && (v0[u] << 1) != 0ull);
                const unsigned b1 = __ballot_sync(0xffffffffu, valid && has2 && (v1[u] << 1) != 0ull);
                if (w0 + u < nw && lane == 0) {
                    masks[(int64_t)i * nw + w0 + u] = b0;
                    if (has2) masks[(int64_t)i2 * nw + w0 + u] = b1;
                }
                if (b0 && first0 == n) first0 = ((w0 + u) << 5) + __ffs(b0) - 1;
                if (b1 && first1 == n) first1 = ((w0 + u) << 5) + __ffs(b1) - 1;
                pc0 += __popc(b0);
                pc1 += __popc(b1);
            }
        }
        if (lane == 0) {
            sm.rowkey[i] = first0;
            if (first0 < n) atomicMax(&sm.repkey[first0], (pc0 << 16) | (0xFFFF - i));
            if (has2) {
                sm.rowkey[i2] = first1;
                if (first1 < n) atomicMax(&sm.repkey[first1], (pc1 << 16) | (0xFFFF - i2));
            }
        }
    }
    __syncthreads();
    // rows whose columns are a subset of their representative's need no processing of their own:
    // rowsec[i] = 1 marks the rows the union-find has to walk
    for (int i = warp; i < m; i += kSecWarps) {
        const int f = sm.rowkey[i];
        int proc = 0;
        if (f < n) {
            const int rep = 0xFFFF - (sm.repkey[f] & 0xFFFF);
            if (rep == i) proc = 1;
            else {
                const uint32_t* mi = masks + (int64_t)i * nw;
                const uint32_t* mr = masks + (int64_t)rep * nw;
                int extra = 0;
                for (int w = lane; w < nw; w += 32) extra |= (mi[w] & ~mr[w]) != 0;
                proc = __any_sync(0xffffffffu, extra);
            }
        }
        if (lane == 0) sm.rowsec[i] = (uint16_t)proc;
    }
    __syncthreads();
    // union-find over columns, walking the marked rows
    while (true) {
        if (tid == 0) sh_flag[0] = 0;
        __syncthreads();
        int changed = 0;
        for (int i = warp; i < m; i += kSecWarps) {
            if (!sm.rowsec[i]) continue;
            const uint32_t* mk = masks + (int64_t)i * nw;
            int rmin = n;
            for (int w = 0; w < nw; ++w)
                if ((mk[w] >> lane) & 1u) rmin = min(rmin, uf_find(parent, (w << 5) + lane));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
            for (int w = 0; w < nw; ++w)
                if ((mk[w] >> lane) & 1u) {
                    const int j = (w << 5) + lane;
                    const int r = uf_find(parent, j);
                    if (r != rmin) { atomicMin(&parent[r], rmin); changed = 1; }
                    if (j != rmin && parent[j] > rmin) atomicMin(&parent[j], rmin);   // path compression (monotone)
                }
        }
        if (changed) sh_flag[0] = 1;
        __syncthreads();
        const int any = sh_flag[0];
        __syncthreads();
        if (!any) break;
    }
    // flatten by pointer jumping (monotone, safe under concurrent updates): parent[j] = root
    while (true) {
        int changed = 0;
        for (int j = tid; j < n; j += kSecThreads) {
            const int p1 = ((volatile int*)parent)[j];
            const int p2 = ((volatile int*)parent)[p1];
            if (p2 != p1) { parent[j] = p2; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    // flag the roots that own at least one row
    for (int j = tid; j < n; j += kSecThreads) sm.colsec[j] = 0;
    __syncthreads();
    for (int i = tid; i < m; i += kSecThreads) {
        const int f = sm.rowkey[i];
        if (f < n) { const int r = parent[f]; sm.rowkey[i] = r; sm.colsec[r] = 1; } else sm.rowkey[i] = -1;
    }
    __syncthreads();
    // sector id of root r = number of flagged roots below r (warp 0 scans the n flags) -> coltmp[r]
    if (warp == 0) {
        int run = 0;
        for (int base = 0; base < n; base += 32) {
            const int j = base + lane;
            const int f = (j < n) ? sm.colsec[j] : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (j < n) sm.coltmp[j] = (uint16_t)(run + __popc(bal & ((1u << lane) - 1)));
            run += __popc(bal);
        }
        if (lane == 0) sh_flag[1] = run;
    }
    __syncthreads();
    const int S = sh_flag[1];
    sm.S = S;
    for (int i = tid; i < m; i += kSecThreads) sm.rowsec[i] = (sm.rowkey[i] >= 0) ? sm.coltmp[sm.rowkey[i]] : (uint16_t)0xFFFF;
    // all-zero columns are their own, unflagged, roots
    for (int j = tid; j < n; j += kSecThreads) { const int r = parent[j]; sm.collist[j] = sm.colsec[r] ? sm.coltmp[r] : (uint16_t)0xFFFF; }
    __syncthreads();
    for (int j = tid; j < n; j += kSecThreads) sm.colsec[j] = sm.collist[j];
    for (int s = tid; s <= S; s += kSecThreads) { sm.cstart[s] = 0; sm.rstart[s] = 0; }
    __syncthreads();
    // counts (integer atomics: deterministic result)
    for (int j = tid; j < n; j += kSecThreads) if (sm.colsec[j] != 0xFFFF) atomicAdd(&sm.cstart[sm.colsec[j] + 1], 1);
    for (int i = tid; i < m; i += kSecThreads) if (sm.rowsec[i] != 0xFFFF) atomicAdd(&sm.rstart[sm.rowsec[i] + 1], 1);
    __syncthreads();
    if (tid == 0) {
        sm.kstart[0] = 0;
        for (int s = 0; s < S; ++s) {
            const int c = sm.cstart[s + 1], r = sm.rstart[s + 1];
            sm.kstart[s + 1] = sm.kstart[s] + (c < r ? c : r);
            sm.cstart[s + 1] += sm.cstart[s];
            sm.rstart[s + 1] += sm.rstart[s];
        }
    }
    __syncthreads();
    // stable compaction: warp w fills the lists of sectors w, w + 8, ...
    for (int s = warp; s < S; s += kSecWarps) {
        int run = sm.cstart[s];
        for (int base = 0; base < n; base += 32) {
            const int j = base + lane;
            const int f = (j < n) && (sm.colsec[j] == s);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (f) sm.collist[run + __popc(bal & ((1u << lane) - 1))] = (uint16_t)j;
            run += __popc(bal);
        }
        run = sm.rstart[s];
        for (int base = 0; base < m; base += 32) {
            const int i = base + lane;
            const int f = (i < m) && (sm.rowsec[i] == s);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (f) sm.rowlist[run + __popc(bal & ((1u << lane) - 1))] = (uint16_t)i;
            run += __popc(bal);
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Householder QR of the compact p x q matrix W (row stride ld, unit column stride), k = min(p, q):
// on exit W(:, 0:k) holds the explicit Q and Rout (k x q, row stride q) the upper-trapezoidal R.
// One warp per trailing column, lanes over rows.  LAPACK dlarfg / dorg2r conventions.
// ------------------------------------------------------------------------------------------------
__device__ void householder_qr(double* W, int ld, int p, int q, int k, double* tau, double* Rout, double* red) {
    // ONE block barrier per column step.  The warp that updates column j+1 in step j also accumulates the norm of
    // its new sub-column and derives reflector j+1 right away, so the norm / sqrt / divisions are off the critical
    // path of the other warps.  The reflector is kept unscaled in W (v_i = W[i][j] * scl[j], v_j = 1) and the
    // diagonal of R in dia[j]; tau / scl / dia live in the caller's `tau` array (3k doubles).
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    double* scl = tau + k;
    double* dia = tau + 2 * k;
    (void)red;
    // all lanes of one warp: reflector of column j from the squared norm of its rows below the diagonal
    auto make_reflector = [&](int j, double xnorm2) {
        const double alpha = W[j * ld + j];
        double tj = 0.0, scale = 0.0, beta = alpha;
        if (xnorm2 != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            tj = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        if (lane == 0) { tau[j] = tj; scl[j] = scale; dia[j] = beta; }
    };
    if (warp == 0 && k > 0) {
        double part = 0.0;
        for (int i = 1 + lane; i < p; i += 32) { const double v = W[i * ld]; part += v * v; }
        make_reflector(0, warp_sum(part));
    }
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        const double tj = tau[j], scale = scl[j];
        const double* vj = W + j;            // column j: vj[i * ld]
        for (int c = j + 1 + warp; c < q; c += nwarps) {
            double* wc = W + c;
            const bool next = (c == j + 1) && (c < k);
            double part = 0.0;
            if (tj != 0.0) {
                double w = 0.0;
                for (int i = j + 1 + lane; i < p; i += 32) w += vj[i * ld] * wc[i * ld];
                w = (warp_sum(w) * scale + wc[j * ld]) * tj;
                const double ws = w * scale;
                if (next) {
                    for (int i = j + 1 + lane; i < p; i += 32) {
                        const double nv = wc[i * ld] - ws * vj[i * ld];
                        wc[i * ld] = nv;
                        if (i > c) part += nv * nv;
                    }
                } else {
                    for (int i = j + 1 + lane; i < p; i += 32) wc[i * ld] -= ws * vj[i * ld];
                }
                __syncwarp();
                if (lane == 0) wc[j * ld] -= w;
            } else if (next) {
                for (int i = c + 1 + lane; i < p; i += 32) { const double v = wc[i * ld]; part += v * v; }
            }
            if (next) {
                __syncwarp();
                make_reflector(c, warp_sum(part));
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < k * q; e += nthreads) {
        const int i = e / q, j = e - i * q;
        Rout[e] = (j > i) ? W[i * ld + j] : (j == i ? dia[i] : 0.0);
    }
    __syncthreads();
    // explicit Q (dorg2r): columns k-1 .. 0.  Column j+1 (the reflector of the previous step) is turned into a column
    // of Q by the warp that is about to apply H_j to it (same lane <-> row mapping, so no extra barrier).
    auto finalize = [&](int j) {             // one warp; lanes own rows j + lane, j + lane + 32, ...
        const double tj = tau[j];
        const double f = -tj * scl[j];
        double* wj = W + j;
        for (int i = j + lane; i < p; i += 32) wj[i * ld] = (i == j) ? 1.0 - tj : wj[i * ld] * f;
        for (int i = lane; i < j; i += 32) wj[i * ld] = 0.0;
    };
    for (int j = k - 1; j >= 0; --j) {
        const double tj = tau[j], scale = scl[j];
        const double* vj = W + j;
        for (int c = j + 1 + warp; c < k; c += nwarps) {
            double* wc = W + c;
            if (c == j + 1) { finalize(c); __syncwarp(); }
            if (tj != 0.0) {
                double w = 0.0;
                for (int i = j + 1 + lane; i < p; i += 32) w += vj[i * ld] * wc[i * ld];
                w = (warp_sum(w) * scale + wc[j * ld]) * tj;
                const double ws = w * scale;
                for (int i = j + 1 + lane; i < p; i += 32) wc[i * ld] -= ws * vj[i * ld];
                __syncwarp();
                if (lane == 0) wc[j * ld] -= w;
            }
        }
        __syncthreads();
    }
    if (warp == 0 && k > 0) finalize(0);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Blocked Householder QR (LAPACK dgeqrf / dorgqr structure) for the work-queue kernel: panels of 8 columns are
// factorised with the unblocked code above, the trailing columns are updated with the compact-WY block reflector
//      A := (I - V T' V^T) A         Z = V^T A  (8 x 8, K = rows)   Y = T' Z   A -= V Y
// on the FP64 tensor pipe (DMMA m8n8k4): the 8 reflectors of a panel are exactly the M = 8 of the instruction.
// ncu of the unblocked kernel showed the column-by-column updates to be issue bound (~110 instructions per
// column and step); a block update costs ~10 instructions per 8 x 8 sub-block.
//   Vp : [P8][8] explicit reflectors of the current panel (unit diagonal, zeros above, zero padded rows), P8 = rows
//        below j0 rounded up to 8;   Tm : [2][8][8] Gram matrix and triangular factor.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_f64(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// C fragment (row gid, cols 2 tig, 2 tig + 1) -> the two B fragments (rows tig / tig + 4, col gid) of the same 8 x 8 matrix
__device__ __forceinline__ void cfrag_to_bfrag(double z0, double z1, int gid, int tig, double& b_lo, double& b_hi) {
    const int src_lo = tig * 4 + (gid >> 1), src_hi = (tig + 4) * 4 + (gid >> 1);
    const double l0 = __shfl_sync(0xffffffffu, z0, src_lo), l1 = __shfl_sync(0xffffffffu, z1, src_lo);
    const double h0 = __shfl_sync(0xffffffffu, z0, src_hi), h1 = __shfl_sync(0xffffffffu, z1, src_hi);
    b_lo = (gid & 1) ? l1 : l0;
    b_hi = (gid & 1) ? h1 : h0;
}

// explicit reflectors of panel [j0, j0 + jb) and its triangular factor T (forward, columnwise: H_1 .. H_jb = I - V T V^T)
__device__ void build_panel_vt(const double* W, int ld, int p, int j0, int jb, const double* tau, const double* scl, double* Vp,
                               double* Tm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int gid = lane >> 2, tig = lane & 3;
    const int P8 = ((p - j0) + 7) & ~7;
    for (int e = tid; e < P8 * 8; e += nthreads) {
        const int i = e >> 3, jj = e & 7;
        const int gi = j0 + i, gj = j0 + jj;
        double v = 0.0;
        if (jj < jb && gi < p) v = (gi == gj) ? 1.0 : (gi > gj ? W[gi * ld + gj] * scl[gj] : 0.0);
        Vp[e] = v;
    }
    __syncthreads();
    if (warp == 0) {
        double g0 = 0.0, g1 = 0.0;
        for (int i0 = 0; i0 < P8; i0 += 4) {
            const double a = Vp[(i0 + tig) * 8 + gid];
            dmma_f64(g0, g1, a, a);                   // G = V^T V
        }
        double* Gs = Tm;          // [8][8]
        double* T = Tm + 64;      // [8][8]
        Gs[gid * 8 + 2 * tig] = g0;
        Gs[gid * 8 + 2 * tig + 1] = g1;
        __syncwarp();
        if (lane < 8) {
            // row `lane` of T: T[i][jj] = -tau_jj * sum_{l = i .. jj-1} T[i][l] G[l][jj]  (i < jj), T[jj][jj] = tau_jj
            double row[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const double tj = jj < jb ? tau[j0 + jj] : 0.0;
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < 8; ++l)
                    if (l < jj && l >= lane) acc += row[l] * Gs[l * 8 + jj];
                row[jj] = (lane == jj) ? tj : (lane < jj ? -tj * acc : 0.0);
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) T[lane * 8 + jj] = row[jj];
        }
    }
    __syncthreads();
}

// W[j0 : p, c_begin : c_end) := (I - V Tm V^T) W[...]  with Tm = T^T (transpose_t, factorisation) or T (forming Q)
__device__ void block_reflect(double* W, int ld, int p, int j0, int c_begin, int c_end, const double* Vp, const double* Tm,
                              bool transpose_t) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int P8 = ((p - j0) + 7) & ~7;
    const double* T = Tm + 64;
    // A fragments of the 8 x 8 factor: a(kk) = Tm'[gid][kk0 + tig]
    const double t_lo = transpose_t ? T[tig * 8 + gid] : T[gid * 8 + tig];
    const double t_hi = transpose_t ? T[(tig + 4) * 8 + gid] : T[gid * 8 + tig + 4];
    for (int c0 = c_begin + 8 * warp; c0 < c_end; c0 += 8 * nwarps) {
        const int cb = c0 + gid;                        // column of this lane's B fragments
        const bool cb_ok = cb < c_end;
        double z0 = 0.0, z1 = 0.0;
        for (int i0 = 0; i0 < P8; i0 += 4) {
            const int gi = j0 + i0 + tig;
            const double a = Vp[(i0 + tig) * 8 + gid];
            const double b = (cb_ok && gi < p) ? W[gi * ld + cb] : 0.0;
            dmma_f64(z0, z1, a, b);                     // Z = V^T A
        }
        double b_lo, b_hi;
        cfrag_to_bfrag(z0, z1, gid, tig, b_lo, b_hi);
        double y0 = 0.0, y1 = 0.0;
        dmma_f64(y0, y1, t_lo, b_lo);
        dmma_f64(y0, y1, t_hi, b_hi);                   // Y = Tm' Z
        cfrag_to_bfrag(-y0, -y1, gid, tig, b_lo, b_hi);
        const int cc = c0 + 2 * tig;                    // columns of this lane's C fragment
        for (int i0 = 0; i0 < P8; i0 += 8) {
            const int gi = j0 + i0 + gid;
            const bool r_ok = gi < p;
            double* wr = W + gi * ld + cc;
            double w0 = (r_ok && cc < c_end) ? wr[0] : 0.0;
            double w1 = (r_ok && cc + 1 < c_end) ? wr[1] : 0.0;
            dmma_f64(w0, w1, Vp[(i0 + gid) * 8 + tig], b_lo);
            dmma_f64(w0, w1, Vp[(i0 + gid) * 8 + tig + 4], b_hi);   // A -= V Y
            if (r_ok && cc < c_end) wr[0] = w0;
            if (r_ok && cc + 1 < c_end) wr[1] = w1;
        }
    }
}

// The same update for a W that lives in GLOBAL memory (sectors beyond shared memory): Vp / Tm are in shared memory and the
// loads of W are issued four k-steps ahead of the DMMAs that consume them (the plain loop above has one dependent L2 round
// trip per step).
__device__ __noinline__ void block_reflect_g(double* __restrict__ W, int ld, int p, int j0, int c_begin, int c_end, const double* Vp, const double* Tm,
                                bool transpose_t, double* Zp) {
    // Work units are (8-column block, row part): with fewer column blocks than warps the rows of a block are split over
    // several warps (every warp then has its own loads in flight), the partial Z = V^T A meet in `Zp` (64 doubles per
    // warp) behind one block barrier and are summed in a fixed order.  ALL threads of the CTA must call this.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int P8 = ((p - j0) + 7) & ~7;
    const double* T = Tm + 64;
    const double t_lo = transpose_t ? T[tig * 8 + gid] : T[gid * 8 + tig];
    const double t_hi = transpose_t ? T[(tig + 4) * 8 + gid] : T[gid * 8 + tig + 4];
    const int nblk = (c_end - c_begin + 7) >> 3;
    int parts = (Zp != nullptr && nblk < nwarps) ? nwarps / nblk : 1;
    if (parts > (P8 + 63) / 64) parts = (P8 + 63) / 64;          // at least 64 rows per part
    if (parts < 1) parts = 1;
    const int rows_per_part = (((P8 + parts - 1) / parts) + 31) & ~31;

    auto z_partial = [&](int c0, int i_begin, int i_end, double& z0, double& z1) {
        const int cb = c0 + gid;
        const bool cb_ok = cb < c_end;
        z0 = 0.0; z1 = 0.0;
        for (int i0 = i_begin; i0 < i_end; i0 += 16) {
            double bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int gi = j0 + i0 + 4 * u + tig;
                bv[u] = (cb_ok && gi < p && i0 + 4 * u < i_end) ? W[(int64_t)gi * ld + cb] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + 4 * u < i_end) dmma_f64(z0, z1, Vp[(i0 + 4 * u + tig) * 8 + gid], bv[u]);     // Z = V^T A
        }
    };
    auto update_rows = [&](int c0, int i_begin, int i_end, double z0, double z1) {
        double b_lo, b_hi;
        cfrag_to_bfrag(z0, z1, gid, tig, b_lo, b_hi);
        double y0 = 0.0, y1 = 0.0;
        dmma_f64(y0, y1, t_lo, b_lo);
        dmma_f64(y0, y1, t_hi, b_hi);                   // Y = Tm' Z
        cfrag_to_bfrag(-y0, -y1, gid, tig, b_lo, b_hi);
        const int cc = c0 + 2 * tig;
        const bool c_ok0 = cc < c_end, c_ok1 = cc + 1 < c_end;
        for (int i0 = i_begin; i0 < i_end; i0 += 32) {
            double w0[4], w1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int gi = j0 + i0 + 8 * u + gid;
                const bool r_ok = gi < p && i0 + 8 * u < i_end;
                const double* wr = W + (int64_t)gi * ld + cc;
                w0[u] = (r_ok && c_ok0) ? wr[0] : 0.0;
                w1[u] = (r_ok && c_ok1) ? wr[1] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + 8 * u < i_end) {
                    const int gi = j0 + i0 + 8 * u + gid;
                    dmma_f64(w0[u], w1[u], Vp[(i0 + 8 * u + gid) * 8 + tig], b_lo);
                    dmma_f64(w0[u], w1[u], Vp[(i0 + 8 * u + gid) * 8 + tig + 4], b_hi);   // A -= V Y
                    double* wr = W + (int64_t)gi * ld + cc;
                    if (gi < p && c_ok0) wr[0] = w0[u];
                    if (gi < p && c_ok1) wr[1] = w1[u];
                }
            }
        }
    };

    if (parts == 1) {
        for (int c0 = c_begin + 8 * warp; c0 < c_end; c0 += 8 * nwarps) {
            double z0, z1;
            z_partial(c0, 0, P8, z0, z1);
            update_rows(c0, 0, P8, z0, z1);
        }
        return;
    }
    const int blk = warp % nblk, part = warp / nblk;
    const bool active = part < parts;
    const int c0 = c_begin + 8 * blk;
    const int i_begin = part * rows_per_part;
    const int i_end = (i_begin + rows_per_part < P8) ? i_begin + rows_per_part : P8;
    if (active && i_begin < P8) {
        double z0, z1;
        z_partial(c0, i_begin, i_end, z0, z1);
        double* zp = Zp + (part * nblk + blk) * 64;
        zp[gid * 8 + 2 * tig] = z0;
        zp[gid * 8 + 2 * tig + 1] = z1;
    } else if (active) {
        double* zp = Zp + (part * nblk + blk) * 64;
        zp[gid * 8 + 2 * tig] = 0.0;
        zp[gid * 8 + 2 * tig + 1] = 0.0;
    }
    __syncthreads();
    if (active && i_begin < P8) {
        double z0 = 0.0, z1 = 0.0;
        for (int pt = 0; pt < parts; ++pt) {            // fixed order: deterministic
            const double* zp = Zp + (pt * nblk + blk) * 64;
            z0 += zp[gid * 8 + 2 * tig];
            z1 += zp[gid * 8 + 2 * tig + 1];
        }
        update_rows(c0, i_begin, i_end, z0, z1);
    }
    __syncthreads();                                     // Zp is reused by the next call
}

__host__ __device__ inline int64_t qr_blocked_extra(int64_t p) { return 8 * ((p + 7) & ~(int64_t)7) + 128; }

// Pan (optional): [P8][8] shared-memory buffer for a W in GLOBAL memory.  Every panel is then staged there, factorised in
// shared memory and written back once, the explicit reflectors are built in place (Vp must alias Pan) and the trailing
// update streams W through block_reflect_g.
__device__ void householder_qr_blocked(double* W, int ld, int p, int q, int k, double* tau, double* Rout, double* Vp, double* Tm,
                                       double* Pan = nullptr, double* Zp = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    double* scl = tau + k;
    double* dia = tau + 2 * k;
    double* Wp = W;      // what the panel code addresses: W itself, or the staged panel shifted so that Wp[i * lp + j] is element (i, j)
    int lp = ld;
    auto stage_panel = [&](int j0, int jb) {
        for (int e = tid; e < (p - j0) * 8; e += nthreads) {
            const int i = e >> 3, jj = e & 7;
            Pan[e] = jj < jb ? W[(int64_t)(j0 + i) * ld + j0 + jj] : 0.0;
        }
        Wp = Pan - (j0 * 8 + j0);
        lp = 8;
        __syncthreads();
    };
    auto unstage_panel = [&](int j0, int jb) {
        for (int e = tid; e < (p - j0) * 8; e += nthreads) {
            const int i = e >> 3, jj = e & 7;
            if (jj < jb) W[(int64_t)(j0 + i) * ld + j0 + jj] = Pan[e];
        }
        __syncthreads();
    };
    auto make_reflector = [&](int j, double xnorm2) {
        const double alpha = Wp[j * lp + j];
        double tj = 0.0, scale = 0.0, beta = alpha;
        if (xnorm2 != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            tj = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        if (lane == 0) { tau[j] = tj; scl[j] = scale; dia[j] = beta; }
    };
    for (int j0 = 0; j0 < k; j0 += 8) {
        const int jb = (k - j0 < 8) ? k - j0 : 8, jend = j0 + jb;
        if (Pan) stage_panel(j0, jb);
        if (warp == 0) {
            double part = 0.0;
            for (int i = j0 + 1 + lane; i < p; i += 32) { const double v = Wp[i * lp + j0]; part += v * v; }
            make_reflector(j0, warp_sum(part));
        }
        __syncthreads();
        for (int j = j0; j < jend; ++j) {
            const double tj = tau[j], scale = scl[j];
            const double* vj = Wp + j;
            for (int c = j + 1 + warp; c < jend; c += nwarps) {
                double* wc = Wp + c;
                const bool next = (c == j + 1);
                double part = 0.0;
                if (tj != 0.0) {
                    double w = 0.0;
                    for (int i = j + 1 + lane; i < p; i += 32) w += vj[i * lp] * wc[i * lp];
                    w = (warp_sum(w) * scale + wc[j * lp]) * tj;
                    const double ws = w * scale;
                    if (next) {
                        for (int i = j + 1 + lane; i < p; i += 32) {
                            const double nv = wc[i * lp] - ws * vj[i * lp];
                            wc[i * lp] = nv;
                            if (i > c) part += nv * nv;
                        }
                    } else {
                        for (int i = j + 1 + lane; i < p; i += 32) wc[i * lp] -= ws * vj[i * lp];
                    }
                    __syncwarp();
                    if (lane == 0) wc[j * lp] -= w;
                } else if (next) {
                    for (int i = c + 1 + lane; i < p; i += 32) { const double v = wc[i * lp]; part += v * v; }
                }
                if (next) {
                    __syncwarp();
                    make_reflector(c, warp_sum(part));
                }
            }
            __syncthreads();
        }
        if (Pan) unstage_panel(j0, jb);
        if (jend < q) {
            build_panel_vt(Wp, lp, p, j0, jb, tau, scl, Vp, Tm);      // in place when Vp aliases the staged panel
            if (Pan) block_reflect_g(W, ld, p, j0, jend, q, Vp, Tm, true, Zp);
            else block_reflect(W, ld, p, j0, jend, q, Vp, Tm, true);
            __syncthreads();
        }
    }
    Wp = W;
    lp = ld;
    for (int e = tid; e < k * q; e += nthreads) {
        const int i = e / q, j = e - i * q;
        Rout[e] = (j > i) ? W[i * ld + j] : (j == i ? dia[i] : 0.0);
    }
    __syncthreads();
    auto finalize = [&](int j, int j0) {     // one warp; lanes own rows j + lane, j + lane + 32, ...
        const double tj = tau[j];
        const double f = -tj * scl[j];
        double* wj = Wp + j;
        for (int i = j + lane; i < p; i += 32) wj[i * lp] = (i == j) ? 1.0 - tj : wj[i * lp] * f;
        const int first = Pan ? j0 : 0;        // rows above a staged panel are zeroed in W itself
        for (int i = first + lane; i < j; i += 32) wj[i * lp] = 0.0;
        if (Pan) for (int i = lane; i < j0; i += 32) W[(int64_t)i * ld + j] = 0.0;
    };
    for (int j0 = ((k - 1) >> 3) << 3; j0 >= 0; j0 -= 8) {
        const int jb = (k - j0 < 8) ? k - j0 : 8, jend = j0 + jb;
        if (jend < k) {
            if (Pan) stage_panel(j0, jb);
            build_panel_vt(Wp, lp, p, j0, jb, tau, scl, Vp, Tm);
            if (Pan) block_reflect_g(W, ld, p, j0, jend, k, Vp, Tm, false, Zp);
            else block_reflect(W, ld, p, j0, jend, k, Vp, Tm, false);
            __syncthreads();
        }
        if (Pan) stage_panel(j0, jb);          // (again: the reflectors were expanded in place)
        for (int j = jend - 1; j >= j0; --j) {
            const double tj = tau[j], scale = scl[j];
            const double* vj = Wp + j;
            for (int c = j + 1 + warp; c < jend; c += nwarps) {
                double* wc = Wp + c;
                if (c == j + 1) { finalize(c, j0); __syncwarp(); }
                if (tj != 0.0) {
                    double w = 0.0;
                    for (int i = j + 1 + lane; i < p; i += 32) w += vj[i * lp] * wc[i * lp];
                    w = (warp_sum(w) * scale + wc[j * lp]) * tj;
                    const double ws = w * scale;
                    for (int i = j + 1 + lane; i < p; i += 32) wc[i * lp] -= ws * vj[i * lp];
                    __syncwarp();
                    if (lane == 0) wc[j * lp] -= w;
                }
            }
            __syncthreads();
        }
        if (warp == 0) finalize(j0, j0);
        __syncthreads();
        if (Pan) unstage_panel(j0, jb);
    }
}

// sect[8] = (m, n, k, a_off, out1_off, out2_off, -, -)
__global__ void __launch_bounds__(kSecThreads) qr_sector_kernel(const int64_t* __restrict__ sect, const double* __restrict__ a, int64_t abs_,
                                                                double* __restrict__ out1, int64_t o1bs, double* __restrict__ out2,
                                                                int64_t o2bs, int use_qr, int nb, double* __restrict__ scratch,
                                                                int64_t scratch_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ int sh_flag[2];
    const int m = (int)sect[0], n = (int)sect[1], k = (int)sect[2];
    const int tid = threadIdx.x;
    SecMap sm;
    secmap_carve(sm, smem_raw, m, n);
    double* work = reinterpret_cast<double*>(smem_raw + secmap_bytes(m, n));
    const int64_t cap = kSecSmemDoubles - secmap_bytes(m, n) / 8;
    double* gscratch = scratch + (int64_t)blockIdx.x * scratch_per_cta;

    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const double* A = a + (int64_t)b * abs_ + sect[3];
        double* O1 = out1 + (int64_t)b * o1bs + sect[4];   // m x k
        double* O2 = out2 + (int64_t)b * o2bs + sect[5];   // k x n
        discover_sectors(sm, A, m, n, sh_flag, reinterpret_cast<uint32_t*>(work), cap * 2);
        const int S = sm.S;
        for (int s = 0; s < S; ++s) {
            const int r0 = sm.rstart[s], c0 = sm.cstart[s];
            const int ms = sm.rstart[s + 1] - r0, ns = sm.cstart[s + 1] - c0;
            if (ms == 0 || ns == 0) continue;
            const int p = use_qr ? ms : ns, q = use_qr ? ns : ms;
            const int ks = p < q ? p : q;
            const int K0 = sm.kstart[s];
            const int ld = q | 1;
            const int64_t need = (int64_t)p * ld + 3 * ks + (int64_t)ks * q;
            double* base = (need <= cap) ? work : gscratch;
            double* W = base;
            double* tau = W + (int64_t)p * ld;
            double* Rc = tau + 3 * ks;
            // gather X: use_qr X = M_s ; else X = M_s^T
            if (use_qr) {
                for (int e = tid; e < ms * ns; e += kSecThreads) {
                    const int r = e / ns, c = e - r * ns;
                    W[(int64_t)r * ld + c] = A[(int64_t)sm.rowlist[r0 + r] * n + sm.collist[c0 + c]];
                }
            } else {
                for (int e = tid; e < ms * ns; e += kSecThreads) {
                    const int r = e / ns, c = e - r * ns;
                    W[(int64_t)c * ld + r] = A[(int64_t)sm.rowlist[r0 + r] * n + sm.collist[c0 + c]];
                }
            }
            __syncthreads();
            householder_qr(W, ld, p, q, ks, tau, Rc, red);
            if (use_qr) {
                // Q (m x k): rows of the sector, bond K0..K0+ks ; R (k x n): columns of the sector
                for (int e = tid; e < ms * ks; e += kSecThreads) {
                    const int r = e / ks, t = e - r * ks;
                    O1[(int64_t)sm.rowlist[r0 + r] * k + K0 + t] = W[(int64_t)r * ld + t];
                }
                for (int e = tid; e < ks * ns; e += kSecThreads) {
                    const int t = e / ns, c = e - t * ns;
                    O2[(int64_t)(K0 + t) * n + sm.collist[c0 + c]] = Rc[(int64_t)t * q + c];
                }
            } else {
                // M = L Q : L (m x k) = R^T, Q (k x n) = Qx^T ; here p = ns, q = ms
                for (int e = tid; e < ms * ks; e += kSecThreads) {
                    const int r = e / ks, t = e - r * ks;
                    O1[(int64_t)sm.rowlist[r0 + r] * k + K0 + t] = Rc[(int64_t)t * q + r];
                }
                for (int e = tid; e < ks * ns; e += kSecThreads) {
                    const int t = e / ns, c = e - t * ns;
                    O2[(int64_t)(K0 + t) * n + sm.collist[c0 + c]] = W[(int64_t)c * ld + t];
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// One-sided Jacobi SVD of a compact sector.  X (p x q, p >= q) is held column-wise: G[c*ldp + r];
// V[c*ldq + t] accumulates the rotations.  Pairs of a round-robin round are distributed over
// sub-warp groups of `gs` lanes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double gsum(double v, int gs) {
    for (int o = gs >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ void jacobi_svd(double* G, int ldp, double* V, int ldq, int p, int q, int* sh_rot) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    int gs = 32;
    while (gs > 4 && (gs >> 1) >= p) gs >>= 1;
    // as many lanes per pair as still let all pairs of a round run in one pass
    while (gs > 4 && ((q + 1) >> 1) * gs > nthreads) gs >>= 1;
    const int groups = nthreads / gs, grp = tid / gs, gl = tid % gs;
    const int qe = q + (q & 1), npairs = qe / 2;
    const double tol = fmax(1e-15, sqrt((double)p) * 2.3e-16), tol2 = tol * tol;
    for (int e = tid; e < q * q; e += nthreads) V[(e / q) * ldq + (e % q)] = ((e / q) == (e % q)) ? 1.0 : 0.0;
    __syncthreads();
    for (int sweep = 0; sweep < 60 && q > 1; ++sweep) {
        if (tid == 0) *sh_rot = 0;
        __syncthreads();
        for (int round = 0; round < qe - 1; ++round) {
            for (int base = 0; base < npairs; base += groups) {
                const int pr = base + grp;
                int i = 0, j = 0;
                bool valid = pr < npairs;
                if (valid) {
                    if (pr == 0) { i = qe - 1; j = round; }
                    else { i = (round + pr) % (qe - 1); j = (round - pr + (qe - 1)) % (qe - 1); }
                    valid = i < q && j < q;
                    if (i > j) { const int t = i; i = j; j = t; }
                }
                double* gi = G + (int64_t)i * ldp;
                double* gj = G + (int64_t)j * ldp;
                double aa = 0.0, bb = 0.0, cc = 0.0;
                if (valid)
                    for (int r = gl; r < p; r += gs) { const double x = gi[r], y = gj[r]; aa += x * x; bb += y * y; cc += x * y; }
                aa = gsum(aa, gs); bb = gsum(bb, gs); cc = gsum(cc, gs);
                if (valid && cc * cc > tol2 * (aa * bb) && aa * bb > 0.0) {
                    // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = (bb - aa) / (2 cc), written with
                    // one square root, one division and one reciprocal square root
                    const double dd = bb - aa, c2 = 2.0 * cc;
                    const double hh = sqrt(dd * dd + c2 * c2);
                    const double tt = (dd >= 0.0 ? c2 : -c2) / (fabs(dd) + hh);
                    const double cs = rsqrt(1.0 + tt * tt), sn = cs * tt;
                    for (int r = gl; r < p; r += gs) { const double x = gi[r], y = gj[r]; gi[r] = cs * x - sn * y; gj[r] = sn * x + cs * y; }
                    double* vi = V + (int64_t)i * ldq;
                    double* vj = V + (int64_t)j * ldq;
                    for (int r = gl; r < q; r += gs) { const double x = vi[r], y = vj[r]; vi[r] = cs * x - sn * y; vj[r] = sn * x + cs * y; }
                    if (gl == 0) *sh_rot = 1;
                }
            }
            __syncthreads();
        }
        const int any = *sh_rot;
        __syncthreads();
        if (!any) break;
    }
}

// Second generation of the CTA-wide Jacobi (used by the work-queue kernel).  Same mathematics, reorganised around
// the two limits the first version hit on B200 (ncu: 33 % of the shared-memory wavefronts were bank conflicts and the
// issue slots were the bottleneck):
//   * columns are padded to an even length and 16-byte aligned, every lane owns PAIRS of rows and moves them with
//     128-bit loads / stores -- a group of 8 lanes touches 128 contiguous bytes, i.e. exactly one conflict-free
//     wavefront, and the number of memory instructions halves;
//   * the rotation is derived with one rsqrt-based square root, one reciprocal and one rsqrt;
//   * a sweep whose largest rotation was below 1e-8 (relative) ends the iteration: the cyclic Jacobi method converges
//     quadratically, so the remaining off-diagonal couplings are O(1e-16) and the confirming sweep is not needed.
// ldp, ldq even; rows p..ldp-1 of G and q..ldq-1 of V must be (and stay) zero.
// 1: the Jacobi sweeps carry the squared column norms (tnsp_jacobi_cached_norms); 0 (default): three dot products per pair.
// Measured on B200 (scripts/mb_sector.py, 296 matrices of 216 x 216 in ~7 sectors): 3.26 ms with the carried norms against
// 2.67 ms without -- the extra shared-memory reads and the per-sweep norm pass cost more than the two saved dot products
// at these column lengths (30 - 70), so the variant stays off and is kept for longer columns / differential tests.
__device__ int g_jacobi_cached_norms = 0;

__device__ void jacobi_svd2(double* G, int ldp, double* V, int ldq, int p, int q, int* sh_rot, double* nrm = nullptr) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int pe = p + (p & 1), qe_rows = q + (q & 1);
    const int qe = q + (q & 1), npairs = qe / 2;
    int gs = 32;
    while (gs > 4 && gs >= pe) gs >>= 1;                       // no more lanes than row pairs
    while (gs > 4 && npairs * gs > nthreads) gs >>= 1;         // all pairs of a round in one pass when possible
    const int groups = nthreads / gs, grp = tid / gs, gl = tid % gs;
    const double tol = fmax(1e-15, sqrt((double)p) * 2.3e-16), tol2 = tol * tol;
    for (int e = tid; e < q * ldq; e += nthreads) { const int cidx = e / ldq, t = e - cidx * ldq; V[e] = (cidx == t) ? 1.0 : 0.0; }
    __syncthreads();
    for (int sweep = 0; sweep < 60 && q > 1; ++sweep) {
        if (tid == 0) *sh_rot = 0;
        if (nrm) {
            // squared column norms, exact at the start of every sweep and carried through the rotations of the sweep
            // (aa' = aa - t cc, bb' = bb + t cc): a pair then costs ONE dot product and one reduction instead of three
            for (int c0 = 0; c0 < q; c0 += groups) {      // uniform trip count: the reduction shuffles are warp-wide
                const int c = c0 + grp;
                double a2 = 0.0;
                if (c < q) {
                    const double2* gc = reinterpret_cast<const double2*>(G + c * ldp);
                    for (int r = gl; 2 * r < pe; r += gs) { const double2 x = gc[r]; a2 = fma(x.x, x.x, a2); a2 = fma(x.y, x.y, a2); }
                }
                a2 = gsum(a2, gs);
                if (c < q && gl == 0) nrm[c] = a2;
            }
        }
        __syncthreads();
        for (int round = 0; round < qe - 1; ++round) {
            for (int base = 0; base < npairs; base += groups) {
                const int pr = base + grp;
                int i = 0, j = 0;
                bool valid = pr < npairs;
                if (valid) {
                    if (pr == 0) { i = qe - 1; j = round; }
                    else { i = round + pr; if (i >= qe - 1) i -= qe - 1; j = round - pr; if (j < 0) j += qe - 1; }
                    valid = i < q && j < q;
                    if (i > j) { const int t = i; i = j; j = t; }
                }
                double2* gi = reinterpret_cast<double2*>(G + i * ldp);
                double2* gj = reinterpret_cast<double2*>(G + j * ldp);
                double aa = 0.0, bb = 0.0, cc = 0.0;
                if (nrm) {
                    if (valid) {
                        for (int r = gl; 2 * r < pe; r += gs) {
                            const double2 x = gi[r], y = gj[r];
                            cc = fma(x.x, y.x, cc); cc = fma(x.y, y.y, cc);
                        }
                        aa = fmax(nrm[i], 0.0); bb = fmax(nrm[j], 0.0);
                    }
                    cc = gsum(cc, gs);
                } else {
                    if (valid)
                        for (int r = gl; 2 * r < pe; r += gs) {
                            const double2 x = gi[r], y = gj[r];
                            aa = fma(x.x, x.x, aa); aa = fma(x.y, x.y, aa);
                            bb = fma(y.x, y.x, bb); bb = fma(y.y, y.y, bb);
                            cc = fma(x.x, y.x, cc); cc = fma(x.y, y.y, cc);
                        }
                    aa = gsum(aa, gs); bb = gsum(bb, gs); cc = gsum(cc, gs);
                }
                const double ab = aa * bb, c2q = cc * cc;
                if (valid && c2q > tol2 * ab && ab > 0.0) {
                    const double dd = bb - aa, c2 = 2.0 * cc;
                    const double x2 = fma(dd, dd, c2 * c2);
                    const double hh = x2 > 1e-280 ? x2 * rsqrt(x2) : sqrt(x2);
                    const double tt = (dd >= 0.0 ? c2 : -c2) * __drcp_rn(fabs(dd) + hh);
                    const double cs = rsqrt(fma(tt, tt, 1.0)), sn = cs * tt;
                    for (int r = gl; 2 * r < pe; r += gs) {
                        const double2 x = gi[r], y = gj[r];
                        gi[r] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
                        gj[r] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
                    }
                    double2* vi = reinterpret_cast<double2*>(V + i * ldq);
                    double2* vj = reinterpret_cast<double2*>(V + j * ldq);
                    for (int r = gl; 2 * r < qe_rows; r += gs) {
                        const double2 x = vi[r], y = vj[r];
                        vi[r] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
                        vj[r] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
                    }
                    if (nrm && gl == 0) { nrm[i] = aa - tt * cc; nrm[j] = bb + tt * cc; }
                    if (gl == 0 && c2q > 1e-16 * ab) *sh_rot = 1;
                }
            }
            __syncthreads();
        }
        const int any = *sh_rot;
        __syncthreads();
        if (!any) break;
    }
}

// Third variant of the CTA-wide Jacobi: the scalar part of a rotation (one square root, one reciprocal, one reciprocal square
// root in double precision: ~45 FP64 instructions) is the same for all lanes of a pair's group, so in jacobi_svd2 every warp spends
// 40 - 50 % of its FP64 issue slots recomputing what its neighbours compute.  Here a round has three phases: (A) dot products per group,
// leaders park (aa, bb, cc) in shared memory; (B) ONE thread per pair derives (cs, sn); (C) the groups apply them.  Two more block
// barriers per round against ~45 % fewer FP64 instructions.  prm: 5 doubles per pair (3 q doubles reserved by svd_sector_need).
__device__ void jacobi_svd3(double* G, int ldp, double* V, int ldq, int p, int q, int* sh_rot, double* prm) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int pe = p + (p & 1), qe_rows = q + (q & 1);
    const int qe = q + (q & 1), npairs = qe / 2;
    int gs = 32;
    while (gs > 4 && gs >= pe) gs >>= 1;
    while (gs > 4 && npairs * gs > nthreads) gs >>= 1;
    const int groups = nthreads / gs, grp = tid / gs, gl = tid % gs;
    const double tol = fmax(1e-15, sqrt((double)p) * 2.3e-16), tol2 = tol * tol;
    for (int e = tid; e < q * ldq; e += nthreads) { const int cidx = e / ldq, t = e - cidx * ldq; V[e] = (cidx == t) ? 1.0 : 0.0; }
    __syncthreads();
    for (int sweep = 0; sweep < 60 && q > 1; ++sweep) {
        if (tid == 0) *sh_rot = 0;
        __syncthreads();
        for (int round = 0; round < qe - 1; ++round) {
            // (A) dot products
            for (int base = 0; base < npairs; base += groups) {
                const int pr = base + grp;
                int i = 0, j = 0;
                bool valid = pr < npairs;
                if (valid) {
                    if (pr == 0) { i = qe - 1; j = round; }
                    else { i = round + pr; if (i >= qe - 1) i -= qe - 1; j = round - pr; if (j < 0) j += qe - 1; }
                    valid = i < q && j < q;
                    if (i > j) { const int t = i; i = j; j = t; }
                }
                const double2* gi = reinterpret_cast<const double2*>(G + i * ldp);
                const double2* gj = reinterpret_cast<const double2*>(G + j * ldp);
                double aa = 0.0, bb = 0.0, cc = 0.0;
                if (valid)
                    for (int r = gl; 2 * r < pe; r += gs) {
                        const double2 x = gi[r], y = gj[r];
                        aa = fma(x.x, x.x, aa); aa = fma(x.y, x.y, aa);
                        bb = fma(y.x, y.x, bb); bb = fma(y.y, y.y, bb);
                        cc = fma(x.x, y.x, cc); cc = fma(x.y, y.y, cc);
                    }
                aa = gsum(aa, gs); bb = gsum(bb, gs); cc = gsum(cc, gs);
                if (gl == 0 && pr < npairs) { prm[5 * pr] = aa; prm[5 * pr + 1] = bb; prm[5 * pr + 2] = valid ? cc : 0.0; }
            }
            __syncthreads();
            // (B) one thread per pair: rotation (cs, sn); sn == 0 means "leave the pair alone"
            for (int pr = tid; pr < npairs; pr += nthreads) {
                const double aa = prm[5 * pr], bb = prm[5 * pr + 1], cc = prm[5 * pr + 2];
                const double ab = aa * bb, c2q = cc * cc;
                double cs = 1.0, sn = 0.0;
                if (c2q > tol2 * ab && ab > 0.0) {
                    const double dd = bb - aa, c2 = 2.0 * cc;
                    const double x2 = fma(dd, dd, c2 * c2);
                    const double hh = x2 > 1e-280 ? x2 * rsqrt(x2) : sqrt(x2);
                    const double tt = (dd >= 0.0 ? c2 : -c2) * __drcp_rn(fabs(dd) + hh);
                    cs = rsqrt(fma(tt, tt, 1.0));
                    sn = cs * tt;
                    if (c2q > 1e-16 * ab) *sh_rot = 1;
                }
                prm[5 * pr + 3] = cs; prm[5 * pr + 4] = sn;
            }
            __syncthreads();
            // (C) apply
            for (int base = 0; base < npairs; base += groups) {
                const int pr = base + grp;
                if (pr >= npairs) continue;
                const double cs = prm[5 * pr + 3], sn = prm[5 * pr + 4];
                if (sn == 0.0) continue;
                int i, j;
                if (pr == 0) { i = qe - 1; j = round; }
                else { i = round + pr; if (i >= qe - 1) i -= qe - 1; j = round - pr; if (j < 0) j += qe - 1; }
                if (i > j) { const int t = i; i = j; j = t; }
                double2* gi = reinterpret_cast<double2*>(G + i * ldp);
                double2* gj = reinterpret_cast<double2*>(G + j * ldp);
                for (int r = gl; 2 * r < pe; r += gs) {
                    const double2 x = gi[r], y = gj[r];
                    gi[r] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
                    gj[r] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
                }
                double2* vi = reinterpret_cast<double2*>(V + i * ldq);
                double2* vj = reinterpret_cast<double2*>(V + j * ldq);
                for (int r = gl; 2 * r < qe_rows; r += gs) {
                    const double2 x = vi[r], y = vj[r];
                    vi[r] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
                    vj[r] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
                }
            }
            __syncthreads();
        }
        const int any = *sh_rot;
        __syncthreads();
        if (!any) break;
    }
}

// ------------------------------------------------------------------------------------------------
// Block Jacobi for a sector whose G (q columns of p rows) and V (q x q) do not fit shared memory: the columns are cut into
// blocks of B, every pair of blocks is staged in shared memory (G and V panels of 2 B columns), swept once there and written
// back; an outer sweep visits all block pairs.  The first version ran the plain sweeps on the global scratch: one dependent
// L2 round trip per row chunk and pair, 0.9 ms for one dense 216 x 216 matrix.
// `jacobi_panel_sweep`: one cyclic sweep (all pairs once) over nc staged columns; *sh_rot != 0 afterwards if a rotation above
// the convergence threshold happened.
// ------------------------------------------------------------------------------------------------
__device__ void jacobi_panel_sweep(double* G, int ldp, double* V, int ldq, int p, int vrows, int nc, int* sh_rot) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int pe = p + (p & 1), ve = vrows + (vrows & 1);
    const int qe = nc + (nc & 1), npairs = qe / 2;
    int gs = 32;
    while (gs > 4 && gs >= pe) gs >>= 1;
    while (gs > 4 && npairs * gs > nthreads) gs >>= 1;
    const int groups = nthreads / gs, grp = tid / gs, gl = tid % gs;
    const double tol = fmax(1e-15, sqrt((double)p) * 2.3e-16), tol2 = tol * tol;
    if (tid == 0) *sh_rot = 0;
    __syncthreads();
    for (int round = 0; round < qe - 1; ++round) {
        for (int base = 0; base < npairs; base += groups) {
            const int pr = base + grp;
            int i = 0, j = 0;
            bool valid = pr < npairs;
            if (valid) {
                if (pr == 0) { i = qe - 1; j = round; }
                else { i = round + pr; if (i >= qe - 1) i -= qe - 1; j = round - pr; if (j < 0) j += qe - 1; }
                valid = i < nc && j < nc;
                if (i > j) { const int t = i; i = j; j = t; }
            }
            double2* gi = reinterpret_cast<double2*>(G + i * ldp);
            double2* gj = reinterpret_cast<double2*>(G + j * ldp);
            double aa = 0.0, bb = 0.0, cc = 0.0;
            if (valid)
                for (int r = gl; 2 * r < pe; r += gs) {
                    const double2 x = gi[r], y = gj[r];
                    aa = fma(x.x, x.x, aa); aa = fma(x.y, x.y, aa);
                    bb = fma(y.x, y.x, bb); bb = fma(y.y, y.y, bb);
                    cc = fma(x.x, y.x, cc); cc = fma(x.y, y.y, cc);
                }
            aa = gsum(aa, gs); bb = gsum(bb, gs); cc = gsum(cc, gs);
            const double ab = aa * bb, c2q = cc * cc;
            if (valid && c2q > tol2 * ab && ab > 0.0) {
                const double dd = bb - aa, c2 = 2.0 * cc;
                const double x2 = fma(dd, dd, c2 * c2);
                const double hh = x2 > 1e-280 ? x2 * rsqrt(x2) : sqrt(x2);
                const double tt = (dd >= 0.0 ? c2 : -c2) * __drcp_rn(fabs(dd) + hh);
                const double cs = rsqrt(fma(tt, tt, 1.0)), sn = cs * tt;
                for (int r = gl; 2 * r < pe; r += gs) {
                    const double2 x = gi[r], y = gj[r];
                    gi[r] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
                    gj[r] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
                }
                double2* vi = reinterpret_cast<double2*>(V + i * ldq);
                double2* vj = reinterpret_cast<double2*>(V + j * ldq);
                for (int r = gl; 2 * r < ve; r += gs) {
                    const double2 x = vi[r], y = vj[r];
                    vi[r] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
                    vj[r] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
                }
                if (gl == 0 && c2q > 1e-16 * ab) *sh_rot = 1;
            }
        }
        __syncthreads();
    }
}

// largest even block width whose pair of G and V panels fits `cap` doubles (0: not even two pairs of columns fit)
__host__ __device__ inline int jacobi_block_width(int64_t ldp, int64_t ldq, int64_t cap) {
    const int64_t b = cap / (2 * (ldp + ldq));
    return (int)(b & ~(int64_t)1);
}

// G, V in global memory (ldp, ldq even, pad rows zero, 16-byte aligned columns); panels: `cap` doubles of shared memory; flags: 2 ints
__device__ void jacobi_svd_blocked(double* G, int ldp, double* V, int ldq, int p, int q, double* panels, int64_t cap, int* flags) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    for (int e = tid; e < q * ldq; e += nthreads) { const int cidx = e / ldq, t = e - cidx * ldq; V[e] = (cidx == t) ? 1.0 : 0.0; }
    const int B = jacobi_block_width(ldp, ldq, cap);
    const int nblk = (q + B - 1) / B;
    __syncthreads();
    for (int outer = 0; outer < 60 && q > 1; ++outer) {
        if (tid == 0) flags[1] = 0;
        __syncthreads();
        for (int I = 0; I < nblk; ++I) {
            for (int J = (nblk == 1 ? I : I + 1); J < nblk; ++J) {
                const int i0 = I * B, ni = min(B, q - i0);
                const int j0 = J * B, nj = (J == I) ? 0 : min(B, q - j0);
                const int nc = ni + nj;
                double* Gs = panels;
                double* Vs = panels + (int64_t)nc * ldp;
                // columns are contiguous in G and V: straight 16-byte copies
                for (int e = tid; e < nc * (ldp >> 1); e += nthreads) {
                    const int c = e / (ldp >> 1), r = e - c * (ldp >> 1);
                    const int gc = c < ni ? i0 + c : j0 + (c - ni);
                    reinterpret_cast<double2*>(Gs + (int64_t)c * ldp)[r] = reinterpret_cast<const double2*>(G + (int64_t)gc * ldp)[r];
                }
                for (int e = tid; e < nc * (ldq >> 1); e += nthreads) {
                    const int c = e / (ldq >> 1), r = e - c * (ldq >> 1);
                    const int gc = c < ni ? i0 + c : j0 + (c - ni);
                    reinterpret_cast<double2*>(Vs + (int64_t)c * ldq)[r] = reinterpret_cast<const double2*>(V + (int64_t)gc * ldq)[r];
                }
                __syncthreads();
                jacobi_panel_sweep(Gs, ldp, Vs, ldq, p, q, nc, &flags[0]);
                if (tid == 0 && flags[0]) flags[1] = 1;
                for (int e = tid; e < nc * (ldp >> 1); e += nthreads) {
                    const int c = e / (ldp >> 1), r = e - c * (ldp >> 1);
                    const int gc = c < ni ? i0 + c : j0 + (c - ni);
                    reinterpret_cast<double2*>(G + (int64_t)gc * ldp)[r] = reinterpret_cast<const double2*>(Gs + (int64_t)c * ldp)[r];
                }
                for (int e = tid; e < nc * (ldq >> 1); e += nthreads) {
                    const int c = e / (ldq >> 1), r = e - c * (ldq >> 1);
                    const int gc = c < ni ? i0 + c : j0 + (c - ni);
                    reinterpret_cast<double2*>(V + (int64_t)gc * ldq)[r] = reinterpret_cast<const double2*>(Vs + (int64_t)c * ldq)[r];
                }
                __syncthreads();
            }
        }
        const int any = flags[1];
        __syncthreads();
        if (!any) break;
    }
}

// Warp-level variant: ONE warp owns a whole (small) sector, so the sectors of a matrix are diagonalised
// concurrently by the warps of the CTA with no block barrier inside the sweeps.  `gs` lanes per column pair.
__device__ void jacobi_svd_warp(double* G, int ldp, double* V, int ldq, int p, int q) {
    const int lane = threadIdx.x & 31;
    const int gs = (p <= 96) ? 4 : 8;
    const int groups = 32 / gs, grp = lane / gs, gl = lane % gs;
    const int qe = q + (q & 1), npairs = qe / 2;
    const double tol = fmax(1e-15, sqrt((double)p) * 2.3e-16), tol2 = tol * tol;
    for (int e = lane; e < q * q; e += 32) V[(e / q) * ldq + (e % q)] = ((e / q) == (e % q)) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60 && q > 1; ++sweep) {
        bool rotated = false;
        for (int round = 0; round < qe - 1; ++round) {
            for (int base = 0; base < npairs; base += groups) {
                const int pr = base + grp;
                int i = 0, j = 0;
                bool valid = pr < npairs;
                if (valid) {
                    if (pr == 0) { i = qe - 1; j = round; }
                    else { i = round + pr; if (i >= qe - 1) i -= qe - 1; j = round - pr; if (j < 0) j += qe - 1; }
                    valid = i < q && j < q;
                    if (i > j) { const int t = i; i = j; j = t; }
                }
                double* gi = G + i * ldp;
                double* gj = G + j * ldp;
                double aa = 0.0, bb = 0.0, cc = 0.0;
                if (valid)
                    for (int r = gl; r < p; r += gs) { const double x = gi[r], y = gj[r]; aa += x * x; bb += y * y; cc += x * y; }
                aa = gsum(aa, gs); bb = gsum(bb, gs); cc = gsum(cc, gs);
                if (valid && cc * cc > tol2 * (aa * bb) && aa * bb > 0.0) {
                    // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = (bb - aa) / (2 cc), written with
                    // one square root, one division and one reciprocal square root
                    const double dd = bb - aa, c2 = 2.0 * cc;
                    const double hh = sqrt(dd * dd + c2 * c2);
                    const double tt = (dd >= 0.0 ? c2 : -c2) / (fabs(dd) + hh);
                    const double cs = rsqrt(1.0 + tt * tt), sn = cs * tt;
                    for (int r = gl; r < p; r += gs) { const double x = gi[r], y = gj[r]; gi[r] = cs * x - sn * y; gj[r] = sn * x + cs * y; }
                    double* vi = V + i * ldq;
                    double* vj = V + j * ldq;
                    for (int r = gl; r < q; r += gs) { const double x = vi[r], y = vj[r]; vi[r] = cs * x - sn * y; vj[r] = sn * x + cs * y; }
                    rotated = true;
                }
            }
            __syncwarp();
        }
        if (!__any_sync(0xffffffffu, rotated)) break;
    }
}

__host__ __device__ inline int64_t svd_sector_work(int64_t m, int64_t n) {
    const int64_t k = m < n ? m : n, p = m < n ? n : m;
    // sigma(k) + sector tags(k) + staged U (m*k) + staged Vt (k*n) + one over-sized sector (p*(k|1) + k*(k|1) + k)
    return 2 * k + m * k + k * n + p * (k | 1) + k * (k | 1) + k + 8;
}

// sect[8] = (m, n, k, a_off, out1_off, out2_off, s_off, -)
__global__ void __launch_bounds__(kSecThreads) svd_sector_kernel(const int64_t* __restrict__ sect, const double* __restrict__ a, int64_t abs_,
                                                                 double* __restrict__ out1, int64_t o1bs, double* __restrict__ sv,
                                                                 int64_t sbs, double* __restrict__ out2, int64_t o2bs,
                                                                 double* __restrict__ workg, int64_t wbs, int nb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_flag[2];
    __shared__ int sh_rot;
    const int m = (int)sect[0], n = (int)sect[1], k = (int)sect[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SecMap sm;
    secmap_carve(sm, smem_raw, m, n);
    double* work = reinterpret_cast<double*>(smem_raw + secmap_bytes(m, n));
    const int64_t cap = kSecSmemDoubles - secmap_bytes(m, n) / 8;

    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const double* A = a + (int64_t)b * abs_ + sect[3];
        double* O1 = out1 + (int64_t)b * o1bs + sect[4];   // U  m x k
        double* O2 = out2 + (int64_t)b * o2bs + sect[5];   // Vt k x n
        double* Sg = sv + (int64_t)b * sbs + sect[6];
        double* wg = workg + (int64_t)b * wbs;
        double* sig_all = wg;                  // [k] singular values in staging order
        double* Ust = wg + 2 * k;              // staged U blocks, sector s at ms*? offsets (below)
        double* Vst = Ust + (int64_t)m * k;
        double* big = Vst + (int64_t)k * n;
        discover_sectors(sm, A, m, n, sh_flag, reinterpret_cast<uint32_t*>(work), cap * 2);
        const int S = sm.S;
        const int ktot = sm.kstart[S];
        // Sectors are processed in batches of consecutive small sectors, one warp per sector (no block barrier
        // inside the Jacobi sweeps); a sector too large for that is diagonalised by the whole CTA.
        // Staged results: U_s (ms x ks) row-major at Ust + uoff, Vt_s (ks x ns) row-major at Vst + voff.
        int64_t uoff = 0, voff = 0;
        int s = 0;
        while (s < S) {
            const int ms0 = sm.rstart[s + 1] - sm.rstart[s], ns0 = sm.cstart[s + 1] - sm.cstart[s];
            const int p0 = ms0 >= ns0 ? ms0 : ns0, q0 = ms0 >= ns0 ? ns0 : ms0;
            const int64_t need0 = (int64_t)q0 * (p0 | 1) + (int64_t)q0 * (q0 | 1) + q0;
            if (q0 > kWarpSectorMax || need0 > cap / 2) {
                // ---- whole-CTA mode for one large sector ----
                const int r0 = sm.rstart[s], c0 = sm.cstart[s];
                const int ms = ms0, ns = ns0;
                const bool tall = ms >= ns;
                const int p = p0, q = q0;
                const int K0 = sm.kstart[s];
                const int ldp = p | 1, ldq = q | 1;
                double* base = (need0 <= cap) ? work : big;
                double* G = base;
                double* V = G + (int64_t)q * ldp;
                double* sig = V + (int64_t)q * ldq;
                for (int e = tid; e < ms * ns; e += kSecThreads) {
                    const int r = e / ns, c = e - r * ns;
                    const double v = A[(int64_t)sm.rowlist[r0 + r] * n + sm.collist[c0 + c]];
                    if (tall) G[(int64_t)c * ldp + r] = v; else G[(int64_t)r * ldp + c] = v;
                }
                __syncthreads();
                jacobi_svd(G, ldp, V, ldq, p, q, &sh_rot);
                for (int c = warp; c < q; c += kSecWarps) {
                    double s2 = 0.0;
                    for (int r = lane; r < p; r += 32) s2 += G[(int64_t)c * ldp + r] * G[(int64_t)c * ldp + r];
                    s2 = warp_sum(s2);
                    if (lane == 0) { sig[c] = sqrt(s2); sig_all[K0 + c] = sqrt(s2); }
                }
                __syncthreads();
                double* Us = Ust + uoff;
                double* Vs = Vst + voff;
                for (int e = tid; e < ms * q; e += kSecThreads) {
                    const int r = e / q, c = e - r * q;
                    if (tall) { const double sg = sig[c]; Us[e] = sg > 0.0 ? G[(int64_t)c * ldp + r] / sg : 0.0; }
                    else Us[e] = V[(int64_t)c * ldq + r];
                }
                for (int e = tid; e < q * ns; e += kSecThreads) {
                    const int c = e / ns, t = e - c * ns;
                    if (tall) Vs[e] = V[(int64_t)c * ldq + t];
                    else { const double sg = sig[c]; Vs[e] = sg > 0.0 ? G[(int64_t)c * ldp + t] / sg : 0.0; }
                }
                uoff += (int64_t)ms * q;
                voff += (int64_t)q * ns;
                __syncthreads();
                ++s;
                continue;
            }
            // ---- batch of small sectors [s, e): warp w takes sector s + w ----
            int e_ = s;
            int64_t off = 0, my_off = -1, my_u = 0, my_v = 0, bu = uoff, bv = voff;
            while (e_ < S && e_ - s < kSecWarps) {
                const int ms = sm.rstart[e_ + 1] - sm.rstart[e_], ns = sm.cstart[e_ + 1] - sm.cstart[e_];
                const int p = ms >= ns ? ms : ns, q = ms >= ns ? ns : ms;
                const int64_t need = (int64_t)q * (p | 1) + (int64_t)q * (q | 1) + q;
                if (q > kWarpSectorMax || need > cap / 2 || off + need > cap) break;
                if (e_ - s == warp) { my_off = off; my_u = bu; my_v = bv; }
                off += need;
                bu += (int64_t)ms * q;
                bv += (int64_t)q * ns;
                ++e_;
            }
            if (my_off >= 0) {
                const int sw = s + warp;
                const int r0 = sm.rstart[sw], c0 = sm.cstart[sw];
                const int ms = sm.rstart[sw + 1] - r0, ns = sm.cstart[sw + 1] - c0;
                const bool tall = ms >= ns;
                const int p = tall ? ms : ns, q = tall ? ns : ms;
                const int K0 = sm.kstart[sw];
                const int ldp = p | 1, ldq = q | 1;
                double* G = work + my_off;
                double* V = G + q * ldp;
                double* sig = V + q * ldq;
                for (int e = lane; e < ms * ns; e += 32) {
                    const int r = e / ns, c = e - r * ns;
                    const double v = A[(int64_t)sm.rowlist[r0 + r] * n + sm.collist[c0 + c]];
                    if (tall) G[c * ldp + r] = v; else G[r * ldp + c] = v;
                }
                __syncwarp();
                jacobi_svd_warp(G, ldp, V, ldq, p, q);
                for (int c = 0; c < q; ++c) {
                    double s2 = 0.0;
                    for (int r = lane; r < p; r += 32) s2 += G[c * ldp + r] * G[c * ldp + r];
                    s2 = warp_sum(s2);
                    if (lane == 0) { sig[c] = sqrt(s2); sig_all[K0 + c] = sqrt(s2); }
                }
                __syncwarp();
                double* Us = Ust + my_u;
                double* Vs = Vst + my_v;
                for (int e = lane; e < ms * q; e += 32) {
                    const int r = e / q, c = e - r * q;
                    if (tall) { const double sg = sig[c]; Us[e] = sg > 0.0 ? G[c * ldp + r] / sg : 0.0; }
                    else Us[e] = V[c * ldq + r];
                }
                for (int e = lane; e < q * ns; e += 32) {
                    const int c = e / ns, t = e - c * ns;
                    if (tall) Vs[e] = V[c * ldq + t];
                    else { const double sg = sig[c]; Vs[e] = sg > 0.0 ? G[c * ldp + t] / sg : 0.0; }
                }
            }
            uoff = bu;
            voff = bv;
            s = e_;
            __syncthreads();
        }
        __threadfence_block();
        __syncthreads();
        // global descending rank of every staged singular value (ties: staging order, i.e. sector order
        // then position, as the greedy cut of svd.hpp:455-461 resolves them)
        uoff = 0; voff = 0;
        for (int s = 0; s < S; ++s) {
            const int r0 = sm.rstart[s], c0 = sm.cstart[s];
            const int ms = sm.rstart[s + 1] - r0, ns = sm.cstart[s + 1] - c0;
            if (ms == 0 || ns == 0) continue;
            const int q = ms < ns ? ms : ns;
            const int K0 = sm.kstart[s];
            const double* Us = Ust + uoff;
            const double* Vs = Vst + voff;
            // ranks of this sector's values into shared scratch (reuse `work` head as int array)
            int* rk = reinterpret_cast<int*>(work);
            for (int c = tid; c < q; c += kSecThreads) {
                const double v = sig_all[K0 + c];
                int r = 0;
                for (int o = 0; o < ktot; ++o) { const double w = sig_all[o]; r += (w > v) || (w == v && o < K0 + c); }
                rk[c] = r;
                Sg[r] = v;
            }
            __syncthreads();
            for (int e = tid; e < ms * q; e += kSecThreads) {
                const int r = e / q, c = e - r * q;
                O1[(int64_t)sm.rowlist[r0 + r] * k + rk[c]] = Us[e];
            }
            for (int e = tid; e < q * ns; e += kSecThreads) {
                const int c = e / ns, t = e - c * ns;
                O2[(int64_t)rk[c] * n + sm.collist[c0 + t]] = Vs[e];
            }
            uoff += (int64_t)ms * q;
            voff += (int64_t)q * ns;
            __syncthreads();
        }
        __syncthreads();
    }
}


// ================================================================================================
// Work-queue variant for the larger matrices (the boundary-MPS bonds of cfg2: 216 x 216, 216 x 1296 ...).
//
// The single-kernel variants above give one CTA a whole chain; a chain has only ~5 sectors of ~40 columns,
// so an SM holds 8 warps that mostly wait on shared-memory latency.  Here the work is split by SECTOR:
//
//   sector_discover_kernel   one CTA per chain: zero pattern -> sector map in global memory (gmap) and one
//                            work item (chain, sector) per non-empty sector, pushed into a queue by cost class
//   qr_work_kernel /         persistent CTAs pop items (heaviest class first) and factorise ONE compact sector
//   svd_work_kernel          in shared memory; several CTAs per SM (72 KiB class) or one (200 KiB class)
//   svd_finish_kernel        per chain: global descending rank of the staged singular values (the greedy
//                            cross-sector cut order of svd.hpp:455-461) and scatter of U / Vt into the dense layout
// ================================================================================================
constexpr int kQSmallDoubles = 9 * 1024;    // 72 KiB  -> 3 CTAs / SM
constexpr int kQSmallThreads = 256;
constexpr int kQBigDoubles = 25 * 1024;     // 200 KiB -> 1 CTA / SM
constexpr int kQBigThreads = 1024;
constexpr int kQMidDoubles = 7040;          // 55 KiB  -> 4 CTAs / SM (the register file allows no more at 256 threads x ~58 registers)
constexpr int kQClasses = 3;                // by shared-memory footprint: 0: beyond 72 KiB (big kernel), 1: 55 - 72 KiB, 2: up to 55 KiB
// qctl layout: [0..2] item counts per class, [3..5] tickets of the kernels that consume class 0 / 1 / 2.
// The one-sided Jacobi and the panel factorisations are latency chains: what a SM delivers is set by how many independent sectors
// it holds, i.e. by the footprint class -- most sectors of cfg2 (<= 58 x 58) fit the 55 KiB class and run four to a SM.

__host__ __device__ inline int gmap_smax(int m, int n) { return (m < n ? m : n) + 2; }
__host__ __device__ inline int64_t gmap_stride(int m, int n) { return 2 + 5 * (int64_t)gmap_smax(m, n) + 2 * ((int64_t)n + m); }

struct GMap {
    const int* base;
    int smax, n;
    __device__ GMap(const int* g, int m_, int n_) : base(g), smax(gmap_smax(m_, n_)), n(n_) {}
    __device__ int S() const { return base[0]; }
    __device__ int ktot() const { return base[1]; }
    __device__ const int* cstart() const { return base + 2; }
    __device__ const int* rstart() const { return base + 2 + smax; }
    __device__ const int* kstart() const { return base + 2 + 2 * smax; }
    __device__ const int* uoff() const { return base + 2 + 3 * smax; }
    __device__ const int* voff() const { return base + 2 + 4 * smax; }
    __device__ const int* collist() const { return base + 2 + 5 * smax; }
    __device__ const int* rowlist() const { return base + 2 + 5 * smax + n; }
    // source offsets of the sorted columns / rows: element (rowlist[r], collist[c]) of the matrix is A[aoff_r[r] + aoff_c[c]]
    __device__ const int* aoff_c(int m_) const { return base + 2 + 5 * smax + n + m_; }
    __device__ const int* aoff_r(int m_) const { return base + 2 + 5 * smax + 2 * n + m_; }
};

__host__ __device__ inline int64_t qr_sector_need(int64_t p, int64_t q) {
    const int64_t k = p < q ? p : q;
    return p * (q | 1) + 3 * k + k * q + 8 * ((p + 7) & ~(int64_t)7) + 128;   // W | tau, scl, dia | R | panel reflectors | G, T
}
// G | V | sigma | 3 q doubles of per-pair scratch (jacobi_svd3: the rotation parameters of a round are computed pair-parallel)
__host__ __device__ inline int64_t svd_sector_need(int64_t p, int64_t q) { return q * (p + (p & 1)) + q * (q + (q & 1)) + q + (q & 1) + 3 * q; }

// kind: 0 = QR of M_s, 1 = LQ of M_s (QR of its transpose), 2 = SVD
__global__ void __launch_bounds__(kSecThreads, 2) sector_discover_kernel(const int64_t* __restrict__ sect, const double* __restrict__ a,
                                                                      int64_t abs_, int nb, int kind, int64_t mask_cap_words,
                                                                      int* __restrict__ gmap, int64_t gstride, int* __restrict__ qctl,
                                                                      int2* __restrict__ qitems, int64_t qcap, const int* __restrict__ rc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_flag[2];
    const int m = (int)sect[0], n = (int)sect[1];
    const int tid = threadIdx.x;
    SecMap sm;
    secmap_carve(sm, smem_raw, m, n);
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem_raw + secmap_bytes(m, n));
    const int smax = gmap_smax(m, n);
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const double* A = a + (int64_t)b * abs_ + sect[3];
        const int* ro = rc, *co = rc ? rc + m : nullptr;
        discover_sectors(sm, A, m, n, sh_flag, masks, mask_cap_words, ro, co);
        const int S = sm.S;
        int* g = gmap + (int64_t)b * gstride;
        int* g_c = g + 2, *g_r = g_c + smax, *g_k = g_r + smax, *g_u = g_k + smax, *g_v = g_u + smax;
        int* g_cl = g_v + smax, *g_rl = g_cl + n;
        for (int s = tid; s <= S; s += kSecThreads) { g_c[s] = sm.cstart[s]; g_r[s] = sm.rstart[s]; g_k[s] = sm.kstart[s]; }
        int* g_ac = g_rl + m, *g_ar = g_ac + n;
        const int nc_used = sm.cstart[S], nr_used = sm.rstart[S];
        for (int j = tid; j < n; j += kSecThreads) {
            const int cj = sm.collist[j];
            g_cl[j] = cj;
            g_ac[j] = j < nc_used ? (co ? __ldg(co + cj) : cj) : 0;
        }
        for (int i = tid; i < m; i += kSecThreads) {
            const int ri = sm.rowlist[i];
            g_rl[i] = ri;
            g_ar[i] = i < nr_used ? (ro ? __ldg(ro + ri) : ri * n) : 0;
        }
        if (tid == 0) {
            g[0] = S;
            g[1] = sm.kstart[S];
            int uo = 0, vo = 0;
            for (int s = 0; s < S; ++s) {
                const int ms = sm.rstart[s + 1] - sm.rstart[s], ns = sm.cstart[s + 1] - sm.cstart[s];
                g_u[s] = uo; g_v[s] = vo;
                if (ms == 0 || ns == 0) continue;
                int64_t need, cost;
                int qq;
                if (kind == 2) {
                    const int pp = ms >= ns ? ms : ns; qq = ms >= ns ? ns : ms;
                    need = svd_sector_need(pp, qq);
                    cost = (int64_t)pp * qq * qq;
                    uo += ms * qq; vo += qq * ns;
                } else {
                    const int pp = kind == 0 ? ms : ns; qq = kind == 0 ? ns : ms;
                    need = qr_sector_need(pp, qq);
                    cost = (int64_t)pp * qq * (pp < qq ? pp : qq);
                }
                const int cls = need > kQSmallDoubles ? 0 : (need > kQMidDoubles ? 1 : 2);
                const int at = atomicAdd(&qctl[cls], 1);
                qitems[(int64_t)cls * qcap + at] = make_int2(b, s);
            }
        }
        __syncthreads();
    }
}

// pops the next work item of a persistent CTA of the kernel that consumes class `which`
__device__ __forceinline__ bool pop_item(int which, int* qctl, const int2* qitems, int64_t qcap, int* sh_ticket, int2& item) {
    __syncthreads();
    if (threadIdx.x == 0) *sh_ticket = atomicAdd(&qctl[3 + which], 1);
    __syncthreads();
    const int t = *sh_ticket;
    if (t >= qctl[which]) return false;
    item = qitems[(int64_t)which * qcap + t];
    return true;
}

// STAGED: the instantiation of the big class (over-sized sectors keep W in the global scratch and stage the panel in shared memory);
// the 72 / 55 KiB classes use the plain instantiation, whose register count the staging code must not raise
template <bool STAGED>
__global__ void __launch_bounds__(kQBigThreads) qr_work_kernel(const int64_t* __restrict__ sect, const double* __restrict__ a, int64_t abs_, double* __restrict__ out1,
                               int64_t o1bs, double* __restrict__ out2, int64_t o2bs, int use_qr, const int* __restrict__ gmap,
                               int64_t gstride, int* qctl, const int2* __restrict__ qitems, int64_t qcap, int which, int64_t cap,
                               double* __restrict__ scratch, int64_t scratch_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_ticket;
    double* work = reinterpret_cast<double*>(smem_raw);
    const int m = (int)sect[0], n = (int)sect[1], k = (int)sect[2];
    const int tid = threadIdx.x, nt = blockDim.x;
    int2 item;
    while (pop_item(which, qctl, qitems, qcap, &sh_ticket, item)) {
        const int b = item.x, s = item.y;
        const GMap g(gmap + (int64_t)b * gstride, m, n);
        const double* A = a + (int64_t)b * abs_ + sect[3];
        double* O1 = out1 + (int64_t)b * o1bs + sect[4];   // m x k
        double* O2 = out2 + (int64_t)b * o2bs + sect[5];   // k x n
        const int r0 = g.rstart()[s], c0 = g.cstart()[s];
        const int ms = g.rstart()[s + 1] - r0, ns = g.cstart()[s + 1] - c0;
        const int* rl = g.rowlist() + r0;
        const int* cl = g.collist() + c0;
        const int p = use_qr ? ms : ns, q = use_qr ? ns : ms;
        const int ks = p < q ? p : q;
        const int K0 = g.kstart()[s];
        const int ld = q | 1;
        const int64_t need = qr_sector_need(p, q);
        double* W = (need <= cap) ? work : scratch + (int64_t)blockIdx.x * scratch_per_cta;
        double* tau = W + (int64_t)p * ld;
        double* Rc = tau + 3 * ks;
        double* Vp = Rc + (int64_t)ks * q;
        double* Tm = Vp + 8 * ((p + 7) & ~7);
        double* Pan = nullptr, *Zp = nullptr;
        if (STAGED && need > cap && 128 + 2 * (int64_t)nt + 3 * (int64_t)ks + 8 * (int64_t)((p + 7) & ~7) <= cap) {
            // an over-sized sector (a dense matrix, for instance): W and R stay in the global scratch, the panel / block reflector,
            // T, tau and the partial Z are staged in shared memory like in the descriptor kernels
            Tm = work; Zp = work + 128; tau = Zp + 2 * nt; Pan = tau + 3 * ks; Vp = Pan;
        }
        const int* ar = g.aoff_r(m) + r0;
        const int* ac = g.aoff_c(m) + c0;
        if (use_qr) {
            for (int e = tid; e < ms * ns; e += nt) {
                const int r = e / ns, c = e - r * ns;
                W[(int64_t)r * ld + c] = __ldg(A + ((int64_t)ar[r] + ac[c]));
            }
        } else {
            for (int e = tid; e < ms * ns; e += nt) {
                const int r = e / ns, c = e - r * ns;
                W[(int64_t)c * ld + r] = __ldg(A + ((int64_t)ar[r] + ac[c]));
            }
        }
        __syncthreads();
        if (q > 8) {
            if constexpr (STAGED) householder_qr_blocked(W, ld, p, q, ks, tau, Rc, Vp, Tm, Pan, Zp);
            else householder_qr_blocked(W, ld, p, q, ks, tau, Rc, Vp, Tm);
        } else householder_qr(W, ld, p, q, ks, tau, Rc, nullptr);
        if (use_qr) {
            for (int e = tid; e < ms * ks; e += nt) {
                const int r = e / ks, t = e - r * ks;
                O1[(int64_t)rl[r] * k + K0 + t] = W[(int64_t)r * ld + t];
            }
            for (int e = tid; e < ks * ns; e += nt) {
                const int t = e / ns, c = e - t * ns;
                O2[(int64_t)(K0 + t) * n + cl[c]] = Rc[(int64_t)t * q + c];
            }
        } else {
            for (int e = tid; e < ms * ks; e += nt) {
                const int r = e / ks, t = e - r * ks;
                O1[(int64_t)rl[r] * k + K0 + t] = Rc[(int64_t)t * q + r];
            }
            for (int e = tid; e < ks * ns; e += nt) {
                const int t = e / ns, c = e - t * ns;
                O2[(int64_t)(K0 + t) * n + cl[c]] = W[(int64_t)c * ld + t];
            }
        }
    }
}

// work (per chain): sigma in staging order [k] | - [k] | staged U blocks [m*k] | staged Vt blocks [k*n] | ...
template <bool BLOCKED>      // the big-class instantiation carries the block Jacobi for over-sized sectors
__global__ void __launch_bounds__(kQBigThreads) svd_work_kernel(const int64_t* __restrict__ sect, const double* __restrict__ a, int64_t abs_, double* __restrict__ workg,
                                int64_t wbs, const int* __restrict__ gmap, int64_t gstride, int* qctl, const int2* __restrict__ qitems,
                                int64_t qcap, int which, int64_t cap, double* __restrict__ scratch, int64_t scratch_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_ticket;
    __shared__ int sh_rot;
    __shared__ int sh_flags[2];
    double* work = reinterpret_cast<double*>(smem_raw);
    const int m = (int)sect[0], n = (int)sect[1], k = (int)sect[2];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    int2 item;
    while (pop_item(which, qctl, qitems, qcap, &sh_ticket, item)) {
        const int b = item.x, s = item.y;
        const GMap g(gmap + (int64_t)b * gstride, m, n);
        const double* A = a + (int64_t)b * abs_ + sect[3];
        double* wg = workg + (int64_t)b * wbs;
        double* sig_all = wg;
        double* Us = wg + 2 * k + g.uoff()[s];
        double* Vs = wg + 2 * k + (int64_t)m * k + g.voff()[s];
        const int r0 = g.rstart()[s], c0 = g.cstart()[s];
        const int ms = g.rstart()[s + 1] - r0, ns = g.cstart()[s + 1] - c0;
        const int* rl = g.rowlist() + r0;
        const int* cl = g.collist() + c0;
        const bool tall = ms >= ns;
        const int p = tall ? ms : ns, q = tall ? ns : ms;
        const int K0 = g.kstart()[s];
        const int ldp = p + (p & 1), ldq = q + (q & 1);
        double* G = (svd_sector_need(p, q) <= cap) ? work : scratch + (int64_t)blockIdx.x * scratch_per_cta;
        double* V = G + (int64_t)q * ldp;
        double* sig = V + (int64_t)q * ldq;
        const int* ar = g.aoff_r(m) + r0;
        const int* ac = g.aoff_c(m) + c0;
        for (int e = tid; e < ms * ns; e += nt) {
            const int r = e / ns, c = e - r * ns;
            const double v = __ldg(A + ((int64_t)ar[r] + ac[c]));
            if (tall) G[(int64_t)c * ldp + r] = v; else G[(int64_t)r * ldp + c] = v;
        }
        if (p & 1) for (int c = tid; c < q; c += nt) G[(int64_t)c * ldp + p] = 0.0;
        __syncthreads();
        if (!BLOCKED || svd_sector_need(p, q) <= cap || jacobi_block_width(ldp, ldq, cap) < 2)
            jacobi_svd2(G, ldp, V, ldq, p, q, &sh_rot, g_jacobi_cached_norms ? sig : nullptr);
        else
            jacobi_svd_blocked(G, ldp, V, ldq, p, q, work, cap, sh_flags);      // over-sized sector: block pairs staged in shared memory
        for (int c = warp; c < q; c += nwarps) {
            double s2 = 0.0;
            for (int r = lane; r < p; r += 32) s2 += G[(int64_t)c * ldp + r] * G[(int64_t)c * ldp + r];
            s2 = warp_sum(s2);
            if (lane == 0) { sig[c] = sqrt(s2); sig_all[K0 + c] = sqrt(s2); }
        }
        __syncthreads();
        for (int e = tid; e < ms * q; e += nt) {
            const int r = e / q, c = e - r * q;
            if (tall) { const double sg = sig[c]; Us[e] = sg > 0.0 ? G[(int64_t)c * ldp + r] / sg : 0.0; }
            else Us[e] = V[(int64_t)c * ldq + r];
        }
        for (int e = tid; e < q * ns; e += nt) {
            const int c = e / ns, t = e - c * ns;
            if (tall) Vs[e] = V[(int64_t)c * ldq + t];
            else { const double sg = sig[c]; Vs[e] = sg > 0.0 ? G[(int64_t)c * ldp + t] / sg : 0.0; }
        }
    }
}

// grid (nb, split): every CTA ranks the chain's staged singular values (value desc, staging order on ties = sector
// order then position, svd.hpp:455-461); CTA y scatters sectors y, y + split, ...
__global__ void __launch_bounds__(256) svd_finish_kernel(const int64_t* __restrict__ sect, double* __restrict__ out1, int64_t o1bs,
                                                         double* __restrict__ sv, int64_t sbs, double* __restrict__ out2, int64_t o2bs,
                                                         const double* __restrict__ workg, int64_t wbs, const int* __restrict__ gmap,
                                                         int64_t gstride, int rank_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = (int)sect[0], n = (int)sect[1], k = (int)sect[2];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b = blockIdx.x;
    const GMap g(gmap + (int64_t)b * gstride, m, n);
    double* O1 = out1 + (int64_t)b * o1bs + sect[4];   // U  m x k
    double* O2 = out2 + (int64_t)b * o2bs + sect[5];   // Vt k x n
    double* Sg = sv + (int64_t)b * sbs + sect[6];
    const double* wg = workg + (int64_t)b * wbs;
    const double* sig_all = wg;
    const int S = g.S(), ktot = g.ktot();
    double* sg_s = reinterpret_cast<double*>(smem_raw);          // [ktot] when rank_in_smem
    int* rk = reinterpret_cast<int*>(sg_s + (rank_in_smem ? k : 0));   // [ktot]
    int* rk_g = reinterpret_cast<int*>(const_cast<double*>(wg) + k);    // [k] ints inside the second k doubles of work
    if (rank_in_smem) {
        for (int c = tid; c < ktot; c += nt) sg_s[c] = sig_all[c];
        __syncthreads();
        for (int c = tid; c < ktot; c += nt) {
            const double v = sg_s[c];
            int r = 0;
            for (int o = 0; o < ktot; ++o) { const double w = sg_s[o]; r += (w > v) || (w == v && o < c); }
            rk[c] = r;
            if (blockIdx.y == 0) Sg[r] = v;
        }
        __syncthreads();
    } else {
        // very wide bonds: ranks are computed by the y == 0 CTA only into the work buffer (gridDim.y == 1 then)
        for (int c = tid; c < ktot; c += nt) {
            const double v = sig_all[c];
            int r = 0;
            for (int o = 0; o < ktot; ++o) { const double w = sig_all[o]; r += (w > v) || (w == v && o < c); }
            rk_g[c] = r;
            Sg[r] = v;
        }
        __threadfence_block();
        __syncthreads();
        rk = rk_g;
    }
    for (int s = blockIdx.y; s < S; s += gridDim.y) {
        const int r0 = g.rstart()[s], c0 = g.cstart()[s];
        const int ms = g.rstart()[s + 1] - r0, ns = g.cstart()[s + 1] - c0;
        if (ms == 0 || ns == 0) continue;
        const int q = ms < ns ? ms : ns;
        const int K0 = g.kstart()[s];
        const double* Us = wg + 2 * k + g.uoff()[s];
        const double* Vs = wg + 2 * k + (int64_t)m * k + g.voff()[s];
        const int* rl = g.rowlist() + r0;
        const int* cl = g.collist() + c0;
        for (int e = tid; e < ms * q; e += nt) {
            const int r = e / q, c = e - r * q;
            O1[(int64_t)rl[r] * k + rk[K0 + c]] = Us[e];
        }
        for (int e = tid; e < q * ns; e += nt) {
            const int c = e / ns, t = e - c * ns;
            O2[(int64_t)rk[K0 + c] * n + cl[t]] = Vs[e];
        }
    }
}

// ---- workspace of the work-queue path (grown on demand, reused by every call on the stream) ----
struct QueueWs {
    int* gmap = nullptr; int64_t gmap_cap = 0;
    int* qctl = nullptr;
    int2* qitems = nullptr; int64_t qitems_cap = 0;
    double* scratch = nullptr; int64_t scratch_cap = 0;
};
static QueueWs g_qws;

template <class T>
static bool grow(T*& ptr, int64_t& cap, int64_t need) {
    if (need <= cap) return true;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    if (cudaMalloc(&ptr, sizeof(T) * need) != cudaSuccess) { cudaGetLastError(); return false; }
    cap = need;
    return true;
}

static int64_t g_queue_min = -1;
static int64_t queue_min_elems() {
    if (g_queue_min < 0) {
        const char* e = getenv("TNSP_SECTOR_QUEUE_MIN");
        g_queue_min = e ? atoll(e) : 2048;
    }
    return g_queue_min;
}

// common front end: sector discovery + queue fill.  Returns 0 ok, 1 error, -1 not applicable.
static int queue_discover(const int64_t* sect, int64_t m, int64_t n, const double* a, int64_t abs_, int nb, int kind, int64_t per_cta_scratch,
                          int big_grid, cudaStream_t st, int64_t& gstride, int64_t& qcap, const int* rc) {
    if (m > kSecMaxDim || n > kSecMaxDim) return -1;
    if (secmap_bytes(m, n) + 4096 > (int64_t)kSecSmemDoubles * 8) return -1;
    gstride = gmap_stride((int)m, (int)n);
    qcap = (int64_t)nb * gmap_smax((int)m, (int)n);
    if (!g_qws.qctl && cudaMalloc(&g_qws.qctl, 8 * sizeof(int)) != cudaSuccess) { set_error("sector queue: cudaMalloc"); return 1; }
    if (!grow(g_qws.gmap, g_qws.gmap_cap, gstride * nb) || !grow(g_qws.qitems, g_qws.qitems_cap, kQClasses * qcap) ||
        !grow(g_qws.scratch, g_qws.scratch_cap, per_cta_scratch * big_grid)) {
        set_error("sector queue: cudaMalloc of the workspace failed");
        return 1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(sector_discover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSecSmemDoubles * 8);
        cudaFuncSetAttribute(qr_work_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(qr_work_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(svd_work_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(svd_work_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(svd_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set = true;
    }
    cudaMemsetAsync(g_qws.qctl, 0, 8 * sizeof(int), st);
    const int64_t nw = (n + 31) / 32;
    const int64_t room = (int64_t)kSecSmemDoubles * 8 - secmap_bytes(m, n);
    int64_t mask_bytes = m * nw * 4;
    int64_t mask_cap_words = m * nw;
    if (mask_bytes > room) { mask_bytes = 0; mask_cap_words = 0; }   // falls back to one dense sector
    const int grid = nb < 8 * kSMs ? nb : 8 * kSMs;
    sector_discover_kernel<<<grid, kSecThreads, secmap_bytes(m, n) + mask_bytes, st>>>(sect, a, abs_, nb, kind, mask_cap_words, g_qws.gmap,
                                                                                       gstride, g_qws.qctl, g_qws.qitems, qcap, rc);
    return check_launch("tnsp sector discovery");
}

static int qr_queue_launch(const int64_t* sect, const int64_t* sh, const double* a, int64_t abs_, double* out1, int64_t o1bs, double* out2,
                           int64_t o2bs, int use_qr, int nb, cudaStream_t st, const int* rc = nullptr) {
    const int64_t m = sh[0], n = sh[1];
    const int64_t p = use_qr ? m : n, q = use_qr ? n : m;
    const int64_t full = qr_sector_need(p, q);   // the largest sector possible
    const int64_t per_cta = full > kQBigDoubles ? full + 8 : 0;
    int64_t gstride, qcap;
    const int err = queue_discover(sect, m, n, a, abs_, nb, use_qr ? 0 : 1, per_cta, kSMs, st, gstride, qcap, rc);
    if (err != 0) return err;
    if (full > kQSmallDoubles) {
        qr_work_kernel<true><<<kSMs, kQBigThreads, kQBigDoubles * 8, st>>>(sect, a, abs_, out1, o1bs, out2, o2bs, use_qr, g_qws.gmap, gstride,
                                                                     g_qws.qctl, g_qws.qitems, qcap, 0, kQBigDoubles, g_qws.scratch, per_cta);
        if (check_launch("tnsp_qr_sectors_f64(big)")) return 1;
    }
    qr_work_kernel<false><<<3 * kSMs, kQSmallThreads, kQSmallDoubles * 8, st>>>(sect, a, abs_, out1, o1bs, out2, o2bs, use_qr, g_qws.gmap, gstride,
                                                                         g_qws.qctl, g_qws.qitems, qcap, 1, kQSmallDoubles, nullptr, 0);
    if (check_launch("tnsp_qr_sectors_f64(72 KiB class)")) return 1;
    qr_work_kernel<false><<<4 * kSMs, kQSmallThreads, kQMidDoubles * 8, st>>>(sect, a, abs_, out1, o1bs, out2, o2bs, use_qr, g_qws.gmap, gstride,
                                                                       g_qws.qctl, g_qws.qitems, qcap, 2, kQMidDoubles, nullptr, 0);
    return check_launch("tnsp_qr_sectors_f64(55 KiB class)");
}

static int svd_queue_launch(const int64_t* sect, const int64_t* sh, const double* a, int64_t abs_, double* out1, int64_t o1bs, double* s,
                            int64_t sbs, double* out2, int64_t o2bs, double* work, int64_t wbs, int nb, cudaStream_t st,
                            const int* rc = nullptr) {
    const int64_t m = sh[0], n = sh[1], k = sh[2];
    const int64_t p = m >= n ? m : n, q = m >= n ? n : m;
    const int64_t full = svd_sector_need(p, q);
    const int64_t per_cta = full > kQBigDoubles ? ((full + 9) & ~(int64_t)1) : 0;   // even: 16-byte aligned slices
    int64_t gstride, qcap;
    const int err = queue_discover(sect, m, n, a, abs_, nb, 2, per_cta, kSMs, st, gstride, qcap, rc);
    if (err != 0) return err;
    if (full > kQSmallDoubles) {
        svd_work_kernel<true><<<kSMs, kQBigThreads, kQBigDoubles * 8, st>>>(sect, a, abs_, work, wbs, g_qws.gmap, gstride, g_qws.qctl, g_qws.qitems,
                                                                      qcap, 0, kQBigDoubles, g_qws.scratch, per_cta);
        if (check_launch("tnsp_svd_sectors_f64(big)")) return 1;
    }
    svd_work_kernel<false><<<3 * kSMs, kQSmallThreads, kQSmallDoubles * 8, st>>>(sect, a, abs_, work, wbs, g_qws.gmap, gstride, g_qws.qctl,
                                                                          g_qws.qitems, qcap, 1, kQSmallDoubles, nullptr, 0);
    if (check_launch("tnsp_svd_sectors_f64(72 KiB class)")) return 1;
    svd_work_kernel<false><<<4 * kSMs, kQSmallThreads, kQMidDoubles * 8, st>>>(sect, a, abs_, work, wbs, g_qws.gmap, gstride, g_qws.qctl,
                                                                        g_qws.qitems, qcap, 2, kQMidDoubles, nullptr, 0);
    if (check_launch("tnsp_svd_sectors_f64(55 KiB class)")) return 1;
    const int rank_in_smem = k * 12 <= 96 * 1024;
    const int split = rank_in_smem ? 4 : 1;
    svd_finish_kernel<<<dim3(nb, split), 256, rank_in_smem ? k * 12 : 0, st>>>(sect, out1, o1bs, s, sbs, out2, o2bs, work, wbs, g_qws.gmap,
                                                                                gstride, rank_in_smem);
    return check_launch("tnsp_svd_sectors_f64(finish)");
}

static double* g_qr_scratch = nullptr;
static int64_t g_qr_scratch_cap = 0;


// ------------------------------------------------------------------------------------------------
// Descriptor-driven sectors beyond the warp class (block-symmetric tensors whose sector shapes the plan names:
// cfg3 - cfg5, several hundred to several thousand rows).  The first kernels for these (factor.cu: qr_kernel,
// svd_kernel) worked column by column out of a global scratch: 9 GB/s for the qr and 2 GB/s for the svd of cfg5.
// Here one (sector, chain) item at a time goes through the blocked Householder code of the work-queue kernel (panel
// of 8 columns, compact-WY trailing update on the FP64 tensor pipe); the svd of a sector that does not fit shared
// memory is preconditioned by that QR (X = Q R, one-sided Jacobi on the small triangular factor, U = Q U_R on the
// tensor pipe), so the Jacobi sweeps run on q x q instead of p x q.
// Items are dealt round-robin in the host's order (largest sector shape first, chains of one shape adjacent).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQBigThreads) qr_desc_kernel(const int64_t* __restrict__ sect, const int* __restrict__ order, int nsel,
                                                              const double* __restrict__ a, int64_t abs_, double* __restrict__ out1,
                                                              int64_t o1bs, double* __restrict__ out2, int64_t o2bs, int use_qr, int nb,
                                                              int64_t cap, double* __restrict__ scratch, int64_t scratch_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* work = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t total = (int64_t)nsel * nb;
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const int64_t* sc = sect + (int64_t)order[item / nb] * TNSP_SECT_COLS;
        const int b = (int)(item % nb);
        const int m = (int)sc[0], n = (int)sc[1], k = (int)sc[2];
        const double* A = a + (int64_t)b * abs_ + sc[3];
        double* O1 = out1 + (int64_t)b * o1bs + sc[4];   // m x k
        double* O2 = out2 + (int64_t)b * o2bs + sc[5];   // k x n
        const int p = use_qr ? m : n, q = use_qr ? n : m;
        const int ks = p < q ? p : q;                     // == k
        const int ld = q | 1;
        const int64_t need = qr_sector_need(p, q);
        double* W = (need <= cap) ? work : scratch + (int64_t)blockIdx.x * scratch_per_cta;
        double* tau = W + (int64_t)p * ld;
        double* Rc = tau + 3 * ks;
        double* Vp = Rc + (int64_t)ks * q;
        double* Tm = Vp + 8 * ((p + 7) & ~7);
        double* Pan = nullptr, *Zp = nullptr;
        if (need > cap && 128 + 2 * (int64_t)nt + 3 * (int64_t)ks + 8 * (int64_t)((p + 7) & ~7) <= cap) {
            // W, R in the global scratch; the panel / block reflector, T, tau and the partial Z in shared memory
            Tm = work; Zp = work + 128; tau = Zp + 2 * nt; Pan = tau + 3 * ks; Vp = Pan;
        }
        if (use_qr) {
            for (int e = tid; e < m * n; e += nt) { const int r = e / n, c = e - r * n; W[(int64_t)r * ld + c] = __ldg(A + e); }
        } else {
            for (int e = tid; e < m * n; e += nt) { const int r = e / n, c = e - r * n; W[(int64_t)c * ld + r] = __ldg(A + e); }
        }
        __syncthreads();
        if (q > 8) householder_qr_blocked(W, ld, p, q, ks, tau, Rc, Vp, Tm, Pan, Zp);
        else householder_qr(W, ld, p, q, ks, tau, Rc, nullptr);
        if (use_qr) {
            for (int e = tid; e < m * k; e += nt) { const int r = e / k, t = e - r * k; O1[e] = W[(int64_t)r * ld + t]; }
            for (int e = tid; e < k * n; e += nt) O2[e] = Rc[e];
        } else {
            for (int e = tid; e < m * k; e += nt) { const int r = e / k, t = e - r * k; O1[e] = Rc[(int64_t)t * q + r]; }
            for (int e = tid; e < k * n; e += nt) { const int t = e / n, c = e - t * n; O2[e] = W[(int64_t)c * ld + t]; }
        }
        __syncthreads();
    }
}

// room of the Jacobi stage of the preconditioned svd: G2 (q x ldq), V (q x ldq), sigma, ranks
__host__ __device__ inline int64_t svd_desc_small_need(int64_t q) { const int64_t ldq = q + (q & 1); return 2 * q * ldq + 2 * ldq; }
// whole footprint of one item in the preconditioned path: the QR working set, then the Jacobi stage
__host__ __device__ inline int64_t svd_desc_pre_need(int64_t p, int64_t q) { return ((qr_sector_need(p, q) + 1) & ~(int64_t)1) + svd_desc_small_need(q); }
// direct path (everything in shared memory): G (q x ldp), V (q x ldq), sigma, ranks
__host__ __device__ inline int64_t svd_desc_direct_need(int64_t p, int64_t q) { return svd_sector_need(p, q) + q + (q & 1) + 2; }

__global__ void __launch_bounds__(kQBigThreads) svd_desc_kernel(const int64_t* __restrict__ sect, const int* __restrict__ order, int nsel,
                                                               const double* __restrict__ a, int64_t abs_, double* __restrict__ out1,
                                                               int64_t o1bs, double* __restrict__ sv, int64_t sbs, double* __restrict__ out2,
                                                               int64_t o2bs, int nb, int64_t cap, double* __restrict__ scratch,
                                                               int64_t scratch_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_rot;
    __shared__ int sh_flags[2];
    double* work = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int64_t total = (int64_t)nsel * nb;
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const int64_t* sc = sect + (int64_t)order[item / nb] * TNSP_SECT_COLS;
        const int b = (int)(item % nb);
        const int m = (int)sc[0], n = (int)sc[1], k = (int)sc[2];
        const double* A = a + (int64_t)b * abs_ + sc[3];
        double* O1 = out1 + (int64_t)b * o1bs + sc[4];   // m x k
        double* O2 = out2 + (int64_t)b * o2bs + sc[5];   // k x n
        double* S = sv + (int64_t)b * sbs + sc[6];
        const bool tall = m >= n;
        const int p = tall ? m : n, q = tall ? n : m;     // X (p x q) = M or M^T, q == k
        const int ldp = p + (p & 1), ldq = q + (q & 1);
        double* gscr = scratch + (int64_t)blockIdx.x * scratch_per_cta;
        if (svd_desc_direct_need(p, q) <= cap) {
            // ---- whole sector in shared memory: Jacobi on X itself ----
            double* G = work;
            double* V = G + (int64_t)q * ldp;
            double* sig = V + (int64_t)q * ldq;
            int* rnk = reinterpret_cast<int*>(sig + ldq);
            for (int e = tid; e < m * n; e += nt) {
                const int r = e / n, c = e - r * n;
                const double v = __ldg(A + e);
                if (tall) G[(int64_t)c * ldp + r] = v; else G[(int64_t)r * ldp + c] = v;
            }
            if (p & 1) for (int c = tid; c < q; c += nt) G[(int64_t)c * ldp + p] = 0.0;
            __syncthreads();
            jacobi_svd2(G, ldp, V, ldq, p, q, &sh_rot, g_jacobi_cached_norms ? sig : nullptr);
            for (int c = warp; c < q; c += nwarps) {
                double s2 = 0.0;
                for (int r = lane; r < p; r += 32) s2 += G[(int64_t)c * ldp + r] * G[(int64_t)c * ldp + r];
                s2 = warp_sum(s2);
                if (lane == 0) sig[c] = sqrt(s2);
            }
            __syncthreads();
            for (int c = tid; c < q; c += nt) {
                const double s = sig[c];
                int r = 0;
                for (int o = 0; o < q; ++o) { const double so = sig[o]; r += (so > s) || (so == s && o < c); }
                rnk[c] = r;
            }
            __syncthreads();
            for (int c = tid; c < q; c += nt) S[rnk[c]] = sig[c];
            // tall: U[r][j] = G[c][r] / sigma_c, Vt[j][t] = V[c][t];  wide: U[t][j] = V[c][t], Vt[j][r] = G[c][r] / sigma_c
            for (int e = tid; e < q * p; e += nt) {
                const int r = e / q, c = e - r * q;         // r: long index (0 .. p), c: column of X
                const double s = sig[c];
                const double v = s > 0.0 ? G[(int64_t)c * ldp + r] / s : 0.0;
                if (tall) O1[(int64_t)r * k + rnk[c]] = v; else O2[(int64_t)rnk[c] * n + r] = v;
            }
            for (int e = tid; e < q * q; e += nt) {
                const int c = e / q, t = e - c * q;
                const double v = V[(int64_t)c * ldq + t];
                if (tall) O2[(int64_t)rnk[c] * n + t] = v; else O1[(int64_t)t * k + rnk[c]] = v;
            }
            __syncthreads();
            continue;
        }
        // ---- preconditioned: X = Q R (blocked Householder out of the global scratch), Jacobi on R ----
        const int ld = q | 1;
        double* W = gscr;
        double* tau = W + (int64_t)p * ld;
        double* Rc = tau + 3 * q;
        double* Vp = Rc + (int64_t)q * q;
        double* Tm = Vp + 8 * ((p + 7) & ~7);
        double* Pan = nullptr, *Zp = nullptr;
        if (128 + 2 * (int64_t)nt + 3 * (int64_t)q + 8 * (int64_t)((p + 7) & ~7) <= cap) {
            // QR stage: panel / block reflector, T, tau and the partial Z in shared memory (the Jacobi stage reuses it afterwards)
            Tm = work; Zp = work + 128; tau = Zp + 2 * nt; Pan = tau + 3 * q; Vp = Pan;
        }
        double* small = (svd_desc_small_need(q) <= cap) ? work : gscr + ((qr_sector_need(p, q) + 1) & ~(int64_t)1);
        double* G2 = small;                                  // column c of R at G2[c * ldq]
        double* V = G2 + (int64_t)q * ldq;
        double* sig = V + (int64_t)q * ldq;
        int* rnk = reinterpret_cast<int*>(sig + ldq);
        for (int e = tid; e < m * n; e += nt) {
            const int r = e / n, c = e - r * n;
            const double v = __ldg(A + e);
            if (tall) W[(int64_t)r * ld + c] = v; else W[(int64_t)c * ld + r] = v;
        }
        __syncthreads();
        if (q > 8) householder_qr_blocked(W, ld, p, q, q, tau, Rc, Vp, Tm, Pan, Zp);
        else householder_qr(W, ld, p, q, q, tau, Rc, nullptr);
        __syncthreads();
        for (int e = tid; e < q * ldq; e += nt) {
            const int c = e / ldq, r = e - c * ldq;
            G2[e] = r < q ? Rc[(int64_t)r * q + c] : 0.0;
        }
        __syncthreads();
        if (small == work || jacobi_block_width(ldq, ldq, cap) < 2)
            jacobi_svd2(G2, ldq, V, ldq, q, q, &sh_rot, g_jacobi_cached_norms ? sig : nullptr);
        else
            jacobi_svd_blocked(G2, ldq, V, ldq, q, q, work, cap, sh_flags);     // q x q factor beyond shared memory
        for (int c = warp; c < q; c += nwarps) {
            double s2 = 0.0;
            for (int r = lane; r < q; r += 32) s2 += G2[(int64_t)c * ldq + r] * G2[(int64_t)c * ldq + r];
            s2 = warp_sum(s2);
            if (lane == 0) sig[c] = sqrt(s2);
        }
        __syncthreads();
        for (int c = tid; c < q; c += nt) {
            const double s = sig[c];
            int r = 0;
            for (int o = 0; o < q; ++o) { const double so = sig[o]; r += (so > s) || (so == s && o < c); }
            rnk[c] = r;
        }
        __syncthreads();
        for (int c = tid; c < q; c += nt) S[rnk[c]] = sig[c];
        for (int e = tid; e < q * q; e += nt) {
            const int c = e / q, t = e - c * q;
            const double v = V[(int64_t)c * ldq + t];
            if (tall) O2[(int64_t)rnk[c] * n + t] = v; else O1[(int64_t)t * k + rnk[c]] = v;
        }
        // left vectors of X: U_X = Q (R V) / sigma = W[:, :q] * G2^T / sigma, 8 rows x 32 columns per warp turn on the tensor pipe
        const int row_blocks = (p + 7) >> 3, col_groups = (q + 31) >> 5;
        for (int wt = warp; wt < row_blocks * col_groups; wt += nwarps) {
            const int i0 = (wt / col_groups) << 3, c0 = (wt % col_groups) << 5;
            double acc[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = 0.0;
            const int ri = i0 + gid;
            const double* wrow = W + (int64_t)(ri < p ? ri : p - 1) * ld;
            for (int t0 = 0; t0 < q; t0 += 4) {
                const int t = t0 + tig;
                const double av = (ri < p && t < q) ? wrow[t] : 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int cb = c0 + 8 * j + gid;
                    const double bv = (cb < q && t < q) ? G2[(int64_t)cb * ldq + t] : 0.0;
                    dmma_f64(acc[j][0], acc[j][1], av, bv);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = c0 + 8 * j + 2 * tig + h;
                    if (ri < p && c < q) {
                        const double s = sig[c];
                        const double v = s > 0.0 ? acc[j][h] / s : 0.0;
                        if (tall) O1[(int64_t)ri * k + rnk[c]] = v; else O2[(int64_t)rnk[c] * n + ri] = v;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---- host side of the descriptor-driven kernels ----
struct DescOrder {
    int* dev = nullptr;
    int64_t cap = 0;
};
static DescOrder g_desc_order;

// sectors of class `big` (working set beyond the 72 KiB CTA) or not, largest first; returns their count
static int desc_select(const int64_t* sh, int ns, bool svd, int use_qr, bool big, std::vector<int>& out, int64_t& scratch_need) {
    std::vector<std::pair<int64_t, int>> sel;
    scratch_need = 0;
    for (int i = 0; i < ns; ++i) {
        const int64_t m = sh[i * TNSP_SECT_COLS], n = sh[i * TNSP_SECT_COLS + 1];
        if (m * n == 0) continue;
        int64_t p, q, need_smem, need_scr = 0;
        if (svd) {
            p = m >= n ? m : n; q = m >= n ? n : m;
            need_smem = svd_desc_direct_need(p, q);
            if (need_smem > kQBigDoubles) need_scr = svd_desc_pre_need(p, q);
        } else {
            p = use_qr ? m : n; q = use_qr ? n : m;
            need_smem = qr_sector_need(p, q);
            if (need_smem > kQBigDoubles) need_scr = need_smem;
        }
        const bool is_big = need_smem > kQSmallDoubles;
        if (is_big != big) continue;
        if (need_scr > scratch_need) scratch_need = need_scr;
        sel.push_back({-(p * q * q), i});
    }
    std::sort(sel.begin(), sel.end());
    out.clear();
    for (auto& e : sel) out.push_back(e.second);
    return (int)out.size();
}

static int desc_upload(const std::vector<int>& a, const std::vector<int>& b, cudaStream_t st) {
    const int64_t total = (int64_t)a.size() + b.size();
    if (total > g_desc_order.cap) {
        if (g_desc_order.dev) cudaFree(g_desc_order.dev);
        g_desc_order.cap = total < 256 ? 256 : 2 * total;
        if (cudaMalloc(&g_desc_order.dev, sizeof(int) * g_desc_order.cap) != cudaSuccess) { set_error("descriptor sectors: cudaMalloc"); return 1; }
    }
    // pageable source: the copy is staged before the call returns, the vectors may go out of scope
    if (!a.empty()) cudaMemcpyAsync(g_desc_order.dev, a.data(), sizeof(int) * a.size(), cudaMemcpyHostToDevice, st);
    if (!b.empty()) cudaMemcpyAsync(g_desc_order.dev + a.size(), b.data(), sizeof(int) * b.size(), cudaMemcpyHostToDevice, st);
    return 0;
}

static bool desc_applicable(const int64_t* sh, int ns) {
    for (int i = 0; i < ns; ++i) {
        const int64_t m = sh[i * TNSP_SECT_COLS], n = sh[i * TNSP_SECT_COLS + 1];
        if (m * n >= (int64_t)1 << 30) return false;      // 32-bit element indices inside the kernels
    }
    return true;
}

// ================================================================================================
// Sector-compact lock-step tensors (tnsp_b200/TAT/ragged.py, csrc/ragged.cuh): the sectors of every chain are NAMED by the
// chain's label tables, stored as contiguous row-major matrices, and factorised from the same footprint-class work queue as
// above -- no zero-pattern discovery, no gather, and the factors are written sector by sector in the compact layout.
// qr.hpp:178-304 / 419-429, svd.hpp:104-211 / 429-481 with the sample axis added.
//   ws (int32 per chain): [0] staged bond size, [1] kept bond size, then per row sector 6 ints
//   (k0 = staging offset of its bond indices, uoff, voff, q = min(m, n) or 0, kept, new bond offset), then dest[kfull]:
//   position of a staged singular triplet inside its sector's kept block (-1: cut)
// ================================================================================================
constexpr bool kRtUseWarpClass = false;
constexpr int kRtTinyDoubles = 2560;          // 20 KiB: class 3 of the queue when the warp class is off
constexpr int kRtTinyThreads = 128;
constexpr int kRtClasses = 4;                 // work queue of the rt_* kernels: counts at qctl[0..3], tickets at qctl[4..7]
constexpr int kRtQrWarpDoubles = 6144;        // 48 KiB per warp, 4 warps per CTA
constexpr int kRtQrWarps = 4;
constexpr int kRtSvdWarpDoubles = 3072;       // 24 KiB per warp, 8 warps per CTA
constexpr int kRtSvdWarps = 8;
__host__ __device__ inline int64_t rt_qr_warp_need(int64_t p, int64_t q, int64_t k) { return p * (q | 1) + k + 2; }
__host__ __device__ inline int64_t rt_svd_warp_need(int64_t p, int64_t q) { return q * (p + (p & 1)) + q * (q + (q & 1)) + q + 2; }

// pops the next item of class `which` (rt layout of qctl) for a whole CTA / for one warp
__device__ __forceinline__ bool rt_pop_cta(int which, int* qctl, const int2* qitems, long long qcap, int* sh_ticket, int2& item) {
    __syncthreads();
    if (threadIdx.x == 0) *sh_ticket = atomicAdd(&qctl[4 + which], 1);
    __syncthreads();
    const int t = *sh_ticket;
    if (t >= qctl[which]) return false;
    item = qitems[(long long)which * qcap + t];
    return true;
}
__device__ __forceinline__ bool rt_pop_warp(int which, int* qctl, const int2* qitems, long long qcap, int2& item) {
    int t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(&qctl[4 + which], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= qctl[which]) return false;
    item = qitems[(long long)which * qcap + t];
    return true;
}

constexpr int kRtJacobiDefault = 2;
constexpr int RT_WS_HDR = 8;
constexpr int RT_WS_SEC = 6;
__host__ __device__ inline int64_t rt_ws_stride(int64_t kfull) { return RT_WS_HDR + RT_WS_SEC * RT_SMAX + kfull; }

// one thread per chain: bond layout of the chain + one work item per non-empty sector
__global__ void __launch_bounds__(128) rt_factor_plan_kernel(RtForm F, int kind, int frs, const int* __restrict__ t1, int t1st, int t1s, int kdim,
                                                             int* __restrict__ labels, int* __restrict__ ws, long long wss, int* qctl,
                                                             int2* __restrict__ qitems, long long qcap, int nb, unsigned long long* stats,
                                                             int mid_cap) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const RtTab R(F.rt + b * F.rts), C(F.ct + b * F.cts);
    const RtMatch Mt(F.match + b * F.mts);
    int* w = ws + (long long)b * wss;
    const int tq = t1 ? t1s * t1[(long long)b * t1st] : 0;
    const int nsec = max(R.nsec(), 0);
    int k0 = 0, uo = 0, vo = 0;
    unsigned long long st_bytes = 0, st_flops = 0, st_n = 0;
    for (int i = 0; i < nsec; ++i) {
        int* e = w + RT_WS_HDR + RT_WS_SEC * i;
        const int j = Mt.mcol(i);
        const int m = R.count(i), n = j >= 0 ? C.count(j) : 0;
        const int q = m < n ? m : n;
        e[0] = k0; e[1] = uo; e[2] = vo; e[3] = q; e[4] = q; e[5] = k0;
        if (q == 0) continue;
        if (kind == 0) {
            const int lam = tq - frs * R.skey(i);
            for (int t = 0; t < q; ++t)
                if (k0 + t < kdim) labels[(long long)b * kdim + k0 + t] = lam;
        }
        int64_t need;
        if (kind == 2) { const int pp = m >= n ? m : n; need = svd_sector_need(pp, q); uo += m * q; vo += q * n; }
        else need = qr_sector_need(m, n);
        st_n += 1;
        if (kind == 2) st_bytes += 8ull * ((unsigned long long)m * n + (unsigned long long)m * q + (unsigned long long)q * n + q);
        else { st_bytes += 8ull * (2ull * m * n + (unsigned long long)m * q + (unsigned long long)q * n); st_flops += 4ull * m * n * q; }
        k0 += q;
        int cls = need > kQSmallDoubles ? 0 : (need > mid_cap ? 1 : 2);
        // optional class 3: one WARP per small sector, no block barrier inside the factorisation.  MEASURED SLOWER on cfg2 (B200, 2368
        // chains: QR 1296 x 216 10.5 vs 6.5 ms, SVD 216 x 216 10.0 vs 8.9 ms per launch: a single warp's dependent shuffle chains
        // take longer than the barriers they save, and 48 KiB per sector leave 4 sectors per SM), so it is switched off.
        if (kRtUseWarpClass) {
            if (kind == 2) { const int pp = m >= n ? m : n; if (q <= kWarpSectorMax && rt_svd_warp_need(pp, q) <= kRtSvdWarpDoubles) cls = 3; }
            else if (rt_qr_warp_need(m, n, q) <= kRtQrWarpDoubles) cls = 3;
        } else if (need <= kRtTinyDoubles) {
            cls = 3;       // tiny sectors (cfg2: the 31 x 31 sectors of a 216 x 216 bond matrix): 128-thread CTAs, ~8 per SM
        }
        const int at = atomicAdd(&qctl[cls], 1);
        qitems[(long long)cls * qcap + at] = make_int2(b, i);
    }
    w[0] = k0; w[1] = k0;
    if (stats) {
        atomicAdd(&stats[kind == 2 ? 6 : 4], st_bytes);
        if (kind != 2) atomicAdd(&stats[5], st_flops);
        atomicAdd(&stats[7], st_n);
    }
    if (kind == 0)
        for (int t = k0; t < kdim; ++t) labels[(long long)b * kdim + t] = 1 << 30;    // dead bond indices
}

// QR of every queued (chain, sector): A_s (m x n, contiguous) = Q_s R_s; Q_s -> first factor at its row sector, R_s -> second
// factor at the bond sector carrying this sector's label
template <bool STAGED>
__global__ void __launch_bounds__(kQBigThreads) rt_qr_work_kernel(RtForm F, int frs, const int* __restrict__ t1, int t1st, int t1s,
                                                                  const int* __restrict__ bond, long long bonds, const int* __restrict__ m1,
                                                                  double* __restrict__ first, long long fst, const int* __restrict__ m2,
                                                                  double* __restrict__ second, long long sst, int* qctl,
                                                                  const int2* __restrict__ qitems, long long qcap, int which, int64_t cap,
                                                                  double* __restrict__ scratch, int64_t scratch_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_ticket;
    double* work = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    int2 item;
    while (rt_pop_cta(which, qctl, qitems, qcap, &sh_ticket, item)) {
        const int b = item.x, i = item.y;
        const RtTab R(F.rt + b * F.rts), C(F.ct + b * F.cts), Bd(bond + b * bonds);
        const RtMatch Mt(F.match + b * F.mts), M1(m1 + (long long)b * RT_MSTRIDE), M2(m2 + (long long)b * RT_MSTRIDE);
        const int j = Mt.mcol(i);
        const int ms = R.count(i), ns = C.count(j);
        const int p = ms, q = ns;
        const int ks = p < q ? p : q;
        const double* A = F.data + (long long)b * F.dstride + Mt.moff(i);
        const int tq = t1 ? t1s * t1[(long long)b * t1st] : 0;
        const int ib = Bd.find(tq - frs * R.skey(i));
        // (a factor whose sectors did not fit its learnt capacity is stored empty: mcol < 0, nothing is written)
        double* O1 = M1.mcol(i) >= 0 ? first + (long long)b * fst + M1.moff(i) : nullptr;                      // ms x ks
        double* O2 = (ib >= 0 && M2.mcol(ib) >= 0) ? second + (long long)b * sst + M2.moff(ib) : nullptr;  // ks x ns
        const int ld = q | 1;
        const int64_t need = qr_sector_need(p, q);
        double* W = (need <= cap) ? work : scratch + (int64_t)blockIdx.x * scratch_per_cta;
        double* tau = W + (int64_t)p * ld;
        double* Rc = tau + 3 * ks;
        double* Vp = Rc + (int64_t)ks * q;
        double* Tm = Vp + 8 * ((p + 7) & ~7);
        double* Pan = nullptr, *Zp = nullptr;
        if (STAGED && need > cap && 128 + 2 * (int64_t)nt + 3 * (int64_t)ks + 8 * (int64_t)((p + 7) & ~7) <= cap) {
            Tm = work; Zp = work + 128; tau = Zp + 2 * nt; Pan = tau + 3 * ks; Vp = Pan;
        }
        for (int e = tid; e < ms * ns; e += nt) {
            const int r = e / ns, c = e - r * ns;
            W[(int64_t)r * ld + c] = __ldg(A + e);
        }
        __syncthreads();
        if (q > 8) {
            if constexpr (STAGED) householder_qr_blocked(W, ld, p, q, ks, tau, Rc, Vp, Tm, Pan, Zp);
            else householder_qr_blocked(W, ld, p, q, ks, tau, Rc, Vp, Tm);
        } else householder_qr(W, ld, p, q, ks, tau, Rc, nullptr);
        if (O1) {
            for (int e = tid; e < ms * ks; e += nt) {
                const int r = e / ks, t = e - r * ks;
                O1[e] = W[(int64_t)r * ld + t];
            }
            if (tid == 0 && ((ms * ks) & 1)) O1[ms * ks] = 0.0;
        }
        if (O2) {
            for (int e = tid; e < ks * ns; e += nt) O2[e] = Rc[e];
            if (tid == 0 && ((ks * ns) & 1)) O2[ks * ns] = 0.0;
        }
    }
}

// SVD of every queued (chain, sector) into the chain's staging buffer: sigma at k0, U_s (m x q) at uoff, Vt_s (q x n) at voff
template <bool BLOCKED>
__global__ void __launch_bounds__(kQBigThreads) rt_svd_work_kernel(RtForm F, double* __restrict__ workg, long long wbs, const int* __restrict__ ws,
                                                                   long long wss, int kfull, int* qctl, const int2* __restrict__ qitems,
                                                                   long long qcap, int which, int64_t cap, double* __restrict__ scratch,
                                                                   int64_t scratch_per_cta, int variant) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_ticket;
    __shared__ int sh_rot;
    __shared__ int sh_flags[2];
    double* work = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    int2 item;
    while (rt_pop_cta(which, qctl, qitems, qcap, &sh_ticket, item)) {
        const int b = item.x, i = item.y;
        const RtTab R(F.rt + b * F.rts), C(F.ct + b * F.cts);
        const RtMatch Mt(F.match + b * F.mts);
        const int* e6 = ws + (long long)b * wss + RT_WS_HDR + RT_WS_SEC * i;
        const int j = Mt.mcol(i);
        const int ms = R.count(i), ns = C.count(j);
        const double* A = F.data + (long long)b * F.dstride + Mt.moff(i);
        double* wg = workg + (long long)b * wbs;
        double* sig_all = wg + e6[0];
        double* Us = wg + kfull + e6[1];
        double* Vs = wg + kfull + (long long)F.M * kfull + e6[2];
        const bool tall = ms >= ns;
        const int p = tall ? ms : ns, q = tall ? ns : ms;
        const int ldp = p + (p & 1), ldq = q + (q & 1);
        double* G = (svd_sector_need(p, q) <= cap) ? work : scratch + (int64_t)blockIdx.x * scratch_per_cta;
        double* V = G + (int64_t)q * ldp;
        double* sig = V + (int64_t)q * ldq;
        for (int e = tid; e < ms * ns; e += nt) {
            const int r = e / ns, c = e - r * ns;
            const double v = __ldg(A + e);
            if (tall) G[(int64_t)c * ldp + r] = v; else G[(int64_t)r * ldp + c] = v;
        }
        if (p & 1) for (int c = tid; c < q; c += nt) G[(int64_t)c * ldp + p] = 0.0;
        __syncthreads();
        if (!BLOCKED || svd_sector_need(p, q) <= cap || jacobi_block_width(ldp, ldq, cap) < 2) {
            if (variant == 3 && q >= 8) jacobi_svd3(G, ldp, V, ldq, p, q, &sh_rot, sig + q + (q & 1));
            else jacobi_svd2(G, ldp, V, ldq, p, q, &sh_rot, g_jacobi_cached_norms ? sig : nullptr);
        } else
            jacobi_svd_blocked(G, ldp, V, ldq, p, q, work, cap, sh_flags);
        for (int c = warp; c < q; c += nwarps) {
            double s2 = 0.0;
            for (int r = lane; r < p; r += 32) s2 += G[(int64_t)c * ldp + r] * G[(int64_t)c * ldp + r];
            s2 = warp_sum(s2);
            if (lane == 0) { sig[c] = sqrt(s2); sig_all[c] = sqrt(s2); }
        }
        __syncthreads();
        for (int e = tid; e < ms * q; e += nt) {
            const int r = e / q, c = e - r * q;
            if (tall) { const double sg = sig[c]; Us[e] = sg > 0.0 ? G[(int64_t)c * ldp + r] / sg : 0.0; }
            else Us[e] = V[(int64_t)c * ldq + r];
        }
        for (int e = tid; e < q * ns; e += nt) {
            const int c = e / ns, t = e - c * ns;
            if (tall) Vs[e] = V[(int64_t)c * ldq + t];
            else { const double sg = sig[c]; Vs[e] = sg > 0.0 ? G[(int64_t)c * ldp + t] / sg : 0.0; }
        }
    }
}

// one WARP per small sector (class 3 of the queue): unblocked Householder QR in the warp's private slice of shared memory, lanes
// over rows (tall) or over columns (wide); no block barrier anywhere (cfg2: 185 x 31 sectors, ~30 us each instead of ~170 us on a
// 256-thread CTA whose every panel column costs a barrier)
__global__ void __launch_bounds__(32 * kRtQrWarps) rt_qr_warp_kernel(RtForm F, int frs, const int* __restrict__ t1, int t1st, int t1s,
                                                                     const int* __restrict__ bond, long long bonds, const int* __restrict__ m1,
                                                                     double* __restrict__ first, long long fst, const int* __restrict__ m2,
                                                                     double* __restrict__ second, long long sst, int* qctl,
                                                                     const int2* __restrict__ qitems, long long qcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* X = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kRtQrWarpDoubles;
    int2 item;
    while (rt_pop_warp(3, qctl, qitems, qcap, item)) {
        const int b = item.x, i = item.y;
        const RtTab R(F.rt + b * F.rts), C(F.ct + b * F.cts), Bd(bond + b * bonds);
        const RtMatch Mt(F.match + b * F.mts), M1(m1 + (long long)b * RT_MSTRIDE), M2(m2 + (long long)b * RT_MSTRIDE);
        const int j = Mt.mcol(i);
        const int p = R.count(i), q = C.count(j);
        const int k = p < q ? p : q;
        const int ld = q | 1;
        double* tau = X + p * ld;
        const double* A = F.data + (long long)b * F.dstride + Mt.moff(i);
        const int tq = t1 ? t1s * t1[(long long)b * t1st] : 0;
        const int ib = Bd.find(tq - frs * R.skey(i));
        double* O1 = M1.mcol(i) >= 0 ? first + (long long)b * fst + M1.moff(i) : nullptr;
        double* O2 = (ib >= 0 && M2.mcol(ib) >= 0) ? second + (long long)b * sst + M2.moff(ib) : nullptr;
        for (int e = lane; e < p * q; e += 32) X[(e / q) * ld + (e % q)] = __ldg(A + e);
        __syncwarp();
        const bool by_rows = p >= q;
        for (int c0 = 0; c0 < k; ++c0) {
            double part = 0.0;
            for (int r = c0 + 1 + lane; r < p; r += 32) { const double v = X[r * ld + c0]; part += v * v; }
            const double xnorm2 = warp_sum(part);
            const double alpha = X[c0 * ld + c0];
            double tj = 0.0, scale = 0.0, beta = alpha;
            if (xnorm2 != 0.0) {
                beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
                tj = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            __syncwarp();
            if (tj != 0.0) {
                for (int r = c0 + 1 + lane; r < p; r += 32) X[r * ld + c0] *= scale;
                __syncwarp();
                if (by_rows) {
                    for (int c = c0 + 1; c < q; ++c) {
                        double w = 0.0;
                        for (int r = c0 + 1 + lane; r < p; r += 32) w += X[r * ld + c0] * X[r * ld + c];
                        w = (warp_sum(w) + X[c0 * ld + c]) * tj;
                        for (int r = c0 + 1 + lane; r < p; r += 32) X[r * ld + c] -= w * X[r * ld + c0];
                        __syncwarp();
                        if (lane == 0) X[c0 * ld + c] -= w;
                    }
                } else {
                    for (int c = c0 + 1 + lane; c < q; c += 32) {
                        double w = X[c0 * ld + c];
                        for (int r = c0 + 1; r < p; ++r) w += X[r * ld + c0] * X[r * ld + c];
                        w *= tj;
                        X[c0 * ld + c] -= w;
                        for (int r = c0 + 1; r < p; ++r) X[r * ld + c] -= w * X[r * ld + c0];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) { tau[c0] = tj; X[c0 * ld + c0] = beta; }
            __syncwarp();
        }
        if (O2) {
            for (int e = lane; e < k * q; e += 32) { const int r = e / q, c = e - r * q; O2[e] = (c >= r) ? X[r * ld + c] : 0.0; }
            if (lane == 0 && ((k * q) & 1)) O2[k * q] = 0.0;
        }
        __syncwarp();
        for (int c0 = k - 1; c0 >= 0; --c0) {
            const double tj = tau[c0];
            if (by_rows) {
                for (int c = c0 + 1; c < k; ++c) {
                    double w = 0.0;
                    for (int r = c0 + 1 + lane; r < p; r += 32) w += X[r * ld + c0] * X[r * ld + c];
                    w = (warp_sum(w) + X[c0 * ld + c]) * tj;
                    for (int r = c0 + 1 + lane; r < p; r += 32) X[r * ld + c] -= w * X[r * ld + c0];
                    __syncwarp();
                    if (lane == 0) X[c0 * ld + c] -= w;
                }
            } else {
                for (int c = c0 + 1 + lane; c < k; c += 32) {
                    double w = X[c0 * ld + c];
                    for (int r = c0 + 1; r < p; ++r) w += X[r * ld + c0] * X[r * ld + c];
                    w *= tj;
                    X[c0 * ld + c] -= w;
                    for (int r = c0 + 1; r < p; ++r) X[r * ld + c] -= w * X[r * ld + c0];
                }
            }
            __syncwarp();
            for (int r = c0 + 1 + lane; r < p; r += 32) X[r * ld + c0] *= -tj;
            for (int r = lane; r < c0; r += 32) X[r * ld + c0] = 0.0;
            if (lane == 0) X[c0 * ld + c0] = 1.0 - tj;
            __syncwarp();
        }
        if (O1) {
            for (int e = lane; e < p * k; e += 32) { const int r = e / k, c = e - r * k; O1[e] = X[r * ld + c]; }
            if (lane == 0 && ((p * k) & 1)) O1[p * k] = 0.0;
        }
        __syncwarp();
    }
}

// one WARP per small sector: one-sided Jacobi without block barriers (jacobi_svd_warp), results staged like rt_svd_work_kernel
__global__ void __launch_bounds__(32 * kRtSvdWarps) rt_svd_warp_kernel(RtForm F, double* __restrict__ workg, long long wbs, const int* __restrict__ ws,
                                                                       long long wss, int kfull, int* qctl, const int2* __restrict__ qitems,
                                                                       long long qcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* G = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kRtSvdWarpDoubles;
    int2 item;
    while (rt_pop_warp(3, qctl, qitems, qcap, item)) {
        const int b = item.x, i = item.y;
        const RtTab R(F.rt + b * F.rts), C(F.ct + b * F.cts);
        const RtMatch Mt(F.match + b * F.mts);
        const int* e6 = ws + (long long)b * wss + RT_WS_HDR + RT_WS_SEC * i;
        const int j = Mt.mcol(i);
        const int ms = R.count(i), ns = C.count(j);
        const double* A = F.data + (long long)b * F.dstride + Mt.moff(i);
        double* wg = workg + (long long)b * wbs;
        double* sig_all = wg + e6[0];
        double* Us = wg + kfull + e6[1];
        double* Vs = wg + kfull + (long long)F.M * kfull + e6[2];
        const bool tall = ms >= ns;
        const int p = tall ? ms : ns, q = tall ? ns : ms;
        const int ldp = p + (p & 1), ldq = q + (q & 1);
        double* V = G + (size_t)q * ldp;
        double* sig = V + (size_t)q * ldq;
        for (int e = lane; e < ms * ns; e += 32) {
            const int r = e / ns, c = e - r * ns;
            const double v = __ldg(A + e);
            if (tall) G[c * ldp + r] = v; else G[r * ldp + c] = v;
        }
        if (p & 1) for (int c = lane; c < q; c += 32) G[c * ldp + p] = 0.0;
        __syncwarp();
        jacobi_svd_warp(G, ldp, V, ldq, p, q);
        __syncwarp();
        for (int c = 0; c < q; ++c) {
            double s2 = 0.0;
            for (int r = lane; r < p; r += 32) s2 += G[c * ldp + r] * G[c * ldp + r];
            s2 = warp_sum(s2);
            if (lane == 0) { sig[c] = sqrt(s2); sig_all[c] = sqrt(s2); }
        }
        __syncwarp();
        for (int e = lane; e < ms * q; e += 32) {
            const int r = e / q, c = e - r * q;
            if (tall) { const double sg = sig[c]; Us[e] = sg > 0.0 ? G[c * ldp + r] / sg : 0.0; }
            else Us[e] = V[c * ldq + r];
        }
        for (int e = lane; e < q * ns; e += 32) {
            const int c = e / ns, t = e - c * ns;
            if (tall) Vs[e] = V[c * ldq + t];
            else { const double sg = sig[c]; Vs[e] = sg > 0.0 ? G[c * ldp + t] / sg : 0.0; }
        }
        __syncwarp();
    }
}

// one CTA per chain: global descending rank of the staged singular values (ties: staging order = sector order, then position:
// svd.hpp:455-461), greedy cut (svd.hpp:429-481: the first remain_cut values above relative_cut * sigma_max), kept counts and the
// labels of the new bond (kept values sector by sector, descending inside a sector)
__global__ void __launch_bounds__(256) rt_svd_finish_kernel(RtForm F, int frs, const int* __restrict__ t1, int t1st, int t1s, int kdim, int kfull,
                                                            long long remain_cut, double relative_cut, const double* __restrict__ workg,
                                                            long long wbs, int* __restrict__ labels, int* __restrict__ ws, long long wss) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[32];
    __shared__ int sec_kept[RT_SMAX], sec_new[RT_SMAX + 1];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const RtTab R(F.rt + b * F.rts);
    int* w = ws + (long long)b * wss;
    const double* sig = workg + (long long)b * wbs;
    const int ktot = w[0];
    const int nsec = max(R.nsec(), 0);
    double* sg = reinterpret_cast<double*>(smem_raw);       // [ktot]
    int* keep = reinterpret_cast<int*>(sg + kfull);         // [ktot] 1 / 0
    int* dest = w + RT_WS_HDR + RT_WS_SEC * RT_SMAX;
    const int tq = t1 ? t1s * t1[(long long)b * t1st] : 0;
    double mx = 0.0;
    for (int c = tid; c < ktot; c += nt) { sg[c] = sig[c]; mx = fmax(mx, sig[c]); }
    mx = block_max(mx, red);
    const double thr = relative_cut * mx;
    for (int c = tid; c < ktot; c += nt) {
        const double v = sg[c];
        int r = 0;
        for (int o = 0; o < ktot; ++o) { const double u = sg[o]; r += (u > v) || (u == v && o < c); }
        keep[c] = (r < remain_cut && v > thr) ? 1 : 0;
    }
    for (int i = tid; i < RT_SMAX; i += nt) sec_kept[i] = 0;
    __syncthreads();
    // position inside the sector: number of kept sector-mates with a larger value (or equal and earlier)
    for (int i = 0; i < nsec; ++i) {
        const int* e6 = w + RT_WS_HDR + RT_WS_SEC * i;
        const int k0 = e6[0], q = e6[3];
        for (int c = tid; c < q; c += nt) {
            int pos = -1;
            if (keep[k0 + c]) {
                pos = 0;
                const double v = sg[k0 + c];
                for (int o = 0; o < q; ++o) { const double u = sg[k0 + o]; pos += keep[k0 + o] && ((u > v) || (u == v && o < c)); }
                atomicAdd(&sec_kept[i], 1);
            }
            dest[k0 + c] = pos;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < nsec; ++i) { sec_new[i] = acc; acc += sec_kept[i]; }
        sec_new[nsec] = acc;
        w[1] = acc;
    }
    __syncthreads();
    for (int i = tid; i < nsec; i += nt) {
        int* e6 = w + RT_WS_HDR + RT_WS_SEC * i;
        e6[4] = sec_kept[i];
        e6[5] = sec_new[i];
    }
    for (int i = 0; i < nsec; ++i) {
        const int lam = tq - frs * R.skey(i);
        for (int t = tid; t < sec_kept[i]; t += nt)
            if (sec_new[i] + t < kdim) labels[(long long)b * kdim + sec_new[i] + t] = lam;
    }
    for (int t = sec_new[nsec] + tid; t < kdim; t += nt) labels[(long long)b * kdim + t] = 1 << 30;
}

// grid (chain, split): kept columns of U_s, rows of Vt_s and the diagonal S_s into the compact factors
__global__ void __launch_bounds__(256) rt_svd_scatter_kernel(RtForm F, int frs, const int* __restrict__ t1, int t1st, int t1s,
                                                             const int* __restrict__ bond, long long bonds, const int* __restrict__ m1,
                                                             double* __restrict__ first, long long fst, const int* __restrict__ m3,
                                                             double* __restrict__ sdat, long long sst, const int* __restrict__ m2,
                                                             double* __restrict__ second, long long cst, const double* __restrict__ workg,
                                                             long long wbs, const int* __restrict__ ws, long long wss, int kfull) {
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const RtTab R(F.rt + b * F.rts), C(F.ct + b * F.cts), Bd(bond + b * bonds);
    const RtMatch Mt(F.match + b * F.mts), M1(m1 + (long long)b * RT_MSTRIDE), M2(m2 + (long long)b * RT_MSTRIDE), M3(m3 + (long long)b * RT_MSTRIDE);
    const int* w = ws + (long long)b * wss;
    const int* dest = w + RT_WS_HDR + RT_WS_SEC * RT_SMAX;
    const double* wg = workg + (long long)b * wbs;
    const int tq = t1 ? t1s * t1[(long long)b * t1st] : 0;
    const int nsec = max(R.nsec(), 0);
    for (int i = blockIdx.y; i < nsec; i += gridDim.y) {
        const int* e6 = w + RT_WS_HDR + RT_WS_SEC * i;
        const int q = e6[3], kc = e6[4];
        if (q == 0 || kc == 0) continue;
        const int j = Mt.mcol(i);
        const int ms = R.count(i), ns = C.count(j);
        const int ib = Bd.find(tq - frs * R.skey(i));
        if (ib < 0 || M1.mcol(i) < 0 || M2.mcol(ib) < 0 || M3.mcol(ib) < 0) continue;     // (dropped by a capacity check)
        const double* sig = wg + e6[0];
        const double* Us = wg + kfull + e6[1];
        const double* Vs = wg + kfull + (long long)F.M * kfull + e6[2];
        const int* dst = dest + e6[0];
        double* O1 = first + (long long)b * fst + M1.moff(i);        // ms x kc
        double* O2 = second + (long long)b * cst + M2.moff(ib);      // kc x ns
        double* O3 = sdat + (long long)b * sst + M3.moff(ib);        // kc x kc
        for (int e = tid; e < ms * q; e += nt) {
            const int r = e / q, c = e - r * q;
            const int d = dst[c];
            if (d >= 0) O1[(long long)r * kc + d] = Us[e];
        }
        for (int e = tid; e < q * ns; e += nt) {
            const int c = e / ns, t = e - c * ns;
            const int d = dst[c];
            if (d >= 0) O2[(long long)d * ns + t] = Vs[e];
        }
        for (int e = tid; e < kc * kc; e += nt) O3[e] = 0.0;
        __syncthreads();
        for (int c = tid; c < q; c += nt) {
            const int d = dst[c];
            if (d >= 0) O3[(long long)d * kc + d] = sig[c];
        }
        if (tid == 0) {
            if ((ms * kc) & 1) O1[ms * kc] = 0.0;
            if ((kc * ns) & 1) O2[kc * ns] = 0.0;
            if ((kc * kc) & 1) O3[kc * kc] = 0.0;
        }
        __syncthreads();
    }
}

}  // namespace tnsp

using namespace tnsp;

// descriptor-driven sectors beyond the warp class; returns 0 ok, 1 error, -1 not applicable
int tnsp_qr_desc_launch(const int64_t* sect, const int64_t* sh, int ns, const double* a, int64_t abs_, double* out1, int64_t o1bs,
                        double* out2, int64_t o2bs, int use_qr, int nb, cudaStream_t st) {
    if (!desc_applicable(sh, ns)) return -1;
    std::vector<int> big, small;
    int64_t scr_big = 0, scr_small = 0;
    desc_select(sh, ns, false, use_qr, true, big, scr_big);
    desc_select(sh, ns, false, use_qr, false, small, scr_small);
    if (desc_upload(big, small, st)) return 1;
    const int64_t per_cta = scr_big ? scr_big + 8 : 0;
    if (!grow(g_qws.scratch, g_qws.scratch_cap, per_cta * kSMs)) { set_error("descriptor sectors: cudaMalloc of the scratch failed"); return 1; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(qr_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(svd_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        attr_set = true;
    }
    if (!big.empty()) {
        const int64_t items = (int64_t)big.size() * nb;
        qr_desc_kernel<<<(unsigned)(items < kSMs ? items : kSMs), kQBigThreads, kQBigDoubles * 8, st>>>(
            sect, g_desc_order.dev, (int)big.size(), a, abs_, out1, o1bs, out2, o2bs, use_qr, nb, kQBigDoubles, g_qws.scratch, per_cta);
        if (check_launch("tnsp_qr_batched_f64(descriptor, big)")) return 1;
    }
    if (!small.empty()) {
        const int64_t items = (int64_t)small.size() * nb;
        qr_desc_kernel<<<(unsigned)(items < 3 * kSMs ? items : 3 * kSMs), kQSmallThreads, kQSmallDoubles * 8, st>>>(
            sect, g_desc_order.dev + big.size(), (int)small.size(), a, abs_, out1, o1bs, out2, o2bs, use_qr, nb, kQSmallDoubles, nullptr, 0);
        if (check_launch("tnsp_qr_batched_f64(descriptor, small)")) return 1;
    }
    return 0;
}

int tnsp_svd_desc_launch(const int64_t* sect, const int64_t* sh, int ns, const double* a, int64_t abs_, double* out1, int64_t o1bs,
                         double* s, int64_t sbs, double* out2, int64_t o2bs, int nb, cudaStream_t st) {
    if (!desc_applicable(sh, ns)) return -1;
    std::vector<int> big, small;
    int64_t scr_big = 0, scr_small = 0;
    desc_select(sh, ns, true, 0, true, big, scr_big);
    desc_select(sh, ns, true, 0, false, small, scr_small);
    if (desc_upload(big, small, st)) return 1;
    const int64_t per_cta = scr_big ? ((scr_big + 9) & ~(int64_t)1) : 0;
    if (!grow(g_qws.scratch, g_qws.scratch_cap, per_cta * kSMs)) { set_error("descriptor sectors: cudaMalloc of the scratch failed"); return 1; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(qr_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(svd_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        attr_set = true;
    }
    if (!big.empty()) {
        const int64_t items = (int64_t)big.size() * nb;
        svd_desc_kernel<<<(unsigned)(items < kSMs ? items : kSMs), kQBigThreads, kQBigDoubles * 8, st>>>(
            sect, g_desc_order.dev, (int)big.size(), a, abs_, out1, o1bs, s, sbs, out2, o2bs, nb, kQBigDoubles, g_qws.scratch, per_cta);
        if (check_launch("tnsp_svd_batched_f64(descriptor, big)")) return 1;
    }
    if (!small.empty()) {
        const int64_t items = (int64_t)small.size() * nb;
        svd_desc_kernel<<<(unsigned)(items < 3 * kSMs ? items : 3 * kSMs), kQSmallThreads, kQSmallDoubles * 8, st>>>(
            sect, g_desc_order.dev + big.size(), (int)small.size(), a, abs_, out1, o1bs, s, sbs, out2, o2bs, nb, kQSmallDoubles, nullptr, 0);
        if (check_launch("tnsp_svd_batched_f64(descriptor, small)")) return 1;
    }
    return 0;
}


// Launchers shared by tnsp_qr_batched_f64 / tnsp_svd_batched_f64 (factor.cu: a single LARGE descriptor) and by
// tnsp_qr_sectors_f64 / tnsp_svd_sectors_f64 below (any size).  Return -1 if the shape is not handled here.
int tnsp_qr_sector_launch(const int64_t* sect, const int64_t* sh, const double* a, int64_t abs_, double* out1, int64_t o1bs,
                          double* out2, int64_t o2bs, int use_qr, int nb, cudaStream_t st) {
    const int64_t m = sh[0], n = sh[1];
    if (m > kSecMaxDim || n > kSecMaxDim) return -1;
    if (secmap_bytes(m, n) / 8 + 4096 > kSecSmemDoubles) return -1;
    if (m * n >= queue_min_elems()) {
        const int rc = qr_queue_launch(sect, sh, a, abs_, out1, o1bs, out2, o2bs, use_qr, nb, st);
        if (rc >= 0) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(qr_sector_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSecSmemDoubles * 8);
        attr_set = true;
    }
    int grid = nb < 2 * kSMs ? nb : 2 * kSMs;
    const int64_t p = use_qr ? m : n, q = use_qr ? n : m, k = p < q ? p : q;
    const int64_t per_cta = p * (q | 1) + 3 * k + k * q + 8;
    const int64_t need = per_cta * grid;
    if (need > g_qr_scratch_cap) {
        if (g_qr_scratch) cudaFree(g_qr_scratch);
        g_qr_scratch_cap = need;
        if (cudaMalloc(&g_qr_scratch, sizeof(double) * need) != cudaSuccess) {
            g_qr_scratch = nullptr; g_qr_scratch_cap = 0;
            set_error("tnsp_qr_batched_f64(sector): cudaMalloc of the spill scratch failed");
            return 1;
        }
    }
    qr_sector_kernel<<<grid, kSecThreads, kSecSmemDoubles * 8, st>>>(sect, a, abs_, out1, o1bs, out2, o2bs, use_qr, nb, g_qr_scratch, per_cta);
    return check_launch("tnsp_qr_batched_f64(sector)");
}

int64_t tnsp_svd_sector_work(int64_t m, int64_t n) { return svd_sector_work(m, n); }

int tnsp_svd_sector_launch(const int64_t* sect, const int64_t* sh, const double* a, int64_t abs_, double* out1, int64_t o1bs,
                           double* s, int64_t sbs, double* out2, int64_t o2bs, double* work, int64_t wbs, int nb, cudaStream_t st) {
    const int64_t m = sh[0], n = sh[1];
    if (m > kSecMaxDim || n > kSecMaxDim) return -1;
    if (secmap_bytes(m, n) / 8 + 4096 > kSecSmemDoubles) return -1;
    if (work == nullptr || wbs < svd_sector_work(m, n)) { set_error("tnsp_svd_batched_f64(sector): scratch too small"); return 1; }
    if (m * n >= queue_min_elems()) {
        const int rc = svd_queue_launch(sect, sh, a, abs_, out1, o1bs, s, sbs, out2, o2bs, work, wbs, nb, st);
        if (rc >= 0) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(svd_sector_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSecSmemDoubles * 8);
        attr_set = true;
    }
    const int grid = nb < 2 * kSMs ? nb : 2 * kSMs;
    svd_sector_kernel<<<grid, kSecThreads, kSecSmemDoubles * 8, st>>>(sect, a, abs_, out1, o1bs, s, sbs, out2, o2bs, work, wbs, nb);
    return check_launch("tnsp_svd_batched_f64(sector)");
}

// C-ABI: factorisation of ONE dense(-embedded) matrix per chain with the symmetry sectors discovered on the
// device (see the header of this file).  Same arguments as tnsp_qr_batched_f64 / tnsp_svd_batched_f64 with ns == 1.
extern "C" int tnsp_qr_sectors_f64(const int64_t* sect, const int64_t* sect_host, double* a, int64_t abs_, double* out1, int64_t o1bs,
                                   double* out2, int64_t o2bs, int use_qr, int nb, void* stream) {
    if (nb == 0 || sect_host[0] * sect_host[1] == 0) return 0;
    const int rc = tnsp_qr_sector_launch(sect, sect_host, a, abs_, out1, o1bs, out2, o2bs, use_qr, nb, (cudaStream_t)stream);
    if (rc < 0) { set_error("tnsp_qr_sectors_f64: matrix too large for the discovered-sector kernel"); return 1; }
    return rc;
}

extern "C" int tnsp_svd_sectors_f64(const int64_t* sect, const int64_t* sect_host, const double* a, int64_t abs_, double* out1,
                                    int64_t o1bs, double* s, int64_t sbs, double* out2, int64_t o2bs, double* work, int64_t wbs,
                                    int nb, void* stream) {
    if (nb == 0 || sect_host[0] * sect_host[1] == 0) return 0;
    const int rc = tnsp_svd_sector_launch(sect, sect_host, a, abs_, out1, o1bs, s, sbs, out2, o2bs, work, wbs, nb, (cudaStream_t)stream);
    if (rc < 0) { set_error("tnsp_svd_sectors_f64: matrix too large for the discovered-sector kernel"); return 1; }
    return rc;
}

// Tuning knob: matrices with at least `min_elems` elements take the per-sector work-queue path, smaller ones the
// one-CTA-per-chain kernel.  Returns the previous value; a negative argument only queries.
extern "C" int tnsp_jacobi_cached_norms(int enable) {
    static int current = 0;
    const int old = current;
    if (enable >= 0 && enable != current) {
        current = enable ? 1 : 0;
        cudaDeviceSynchronize();
        if (cudaMemcpyToSymbol(g_jacobi_cached_norms, &current, sizeof(int)) != cudaSuccess) { set_error("tnsp_jacobi_cached_norms: cudaMemcpyToSymbol"); return -1; }
    }
    return old;
}

extern "C" int64_t tnsp_sector_queue_min(int64_t min_elems) {
    const int64_t old = queue_min_elems();
    if (min_elems >= 0) g_queue_min = min_elems;
    return old;
}

// The same factorisations with the m x n operand READ IN PLACE from a dense tensor of any index order: element (i, j)
// of the matrix is a[b * abs + rc[i] + rc[m + j]] (int32 element offsets built by the host planner from the edge strides,
// merge order of edge_operator.hpp:321-404).  Replaces the merge / transpose copy in front of qr.hpp:309-508 /
// svd.hpp:259-538.  sect[3] (a_off) must be 0.  Always the work-queue path.
extern "C" int tnsp_qr_sectors_gather_f64(const int64_t* sect, const int64_t* sect_host, const int32_t* rc, const double* a, int64_t abs_,
                                          double* out1, int64_t o1bs, double* out2, int64_t o2bs, int use_qr, int nb, void* stream) {
    if (nb == 0 || sect_host[0] * sect_host[1] == 0) return 0;
    if (sect_host[0] > kSecMaxDim || sect_host[1] > kSecMaxDim) { set_error("tnsp_qr_sectors_gather_f64: matrix too large"); return 1; }
    const int err = qr_queue_launch(sect, sect_host, a, abs_, out1, o1bs, out2, o2bs, use_qr, nb, (cudaStream_t)stream, rc);
    if (err < 0) { set_error("tnsp_qr_sectors_gather_f64: matrix too large for the discovered-sector kernels"); return 1; }
    return err;
}

extern "C" int tnsp_svd_sectors_gather_f64(const int64_t* sect, const int64_t* sect_host, const int32_t* rc, const double* a, int64_t abs_,
                                           double* out1, int64_t o1bs, double* s, int64_t sbs, double* out2, int64_t o2bs, double* work,
                                           int64_t wbs, int nb, void* stream) {
    if (nb == 0 || sect_host[0] * sect_host[1] == 0) return 0;
    if (sect_host[0] > kSecMaxDim || sect_host[1] > kSecMaxDim) { set_error("tnsp_svd_sectors_gather_f64: matrix too large"); return 1; }
    if (work == nullptr || wbs < svd_sector_work(sect_host[0], sect_host[1])) { set_error("tnsp_svd_sectors_gather_f64: scratch too small"); return 1; }
    const int err = svd_queue_launch(sect, sect_host, a, abs_, out1, o1bs, s, sbs, out2, o2bs, work, wbs, nb, (cudaStream_t)stream, rc);
    if (err < 0) { set_error("tnsp_svd_sectors_gather_f64: matrix too large for the discovered-sector kernels"); return 1; }
    return err;
}

// ---- sector-compact entry points (tnsp_b200/TAT/ragged.py) ----
// launch shapes of the work-queue classes (threads per CTA); TNSP_RT_* environment variables override them for experiments
static int rt_tune(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int x = atoi(v);
    return (x >= 32 && x <= 1024 && x % 32 == 0) ? x : dflt;
}
// capacity of class 2 in doubles (default 55 KiB: 4 CTAs per SM) and the number of such CTAs an SM holds
static int rt_mid_doubles() {
    static int v = 0;
    if (!v) { const char* e = getenv("TNSP_RT_MID_DOUBLES"); v = (e && atoi(e) >= 2560 && atoi(e) <= kQSmallDoubles) ? atoi(e) : kQMidDoubles; }
    return v;
}
// Jacobi variant of the sector SVD: 2 = rotations derived per group (default), 3 = pair-parallel rotation phase (TNSP_RT_JACOBI)
static int rt_jacobi_variant() {
    static int v = 0;
    if (!v) { const char* e = getenv("TNSP_RT_JACOBI"); v = (e && atoi(e) == 3) ? 3 : kRtJacobiDefault; }
    return v;
}
static int rt_mid_ctas() { return std::max(1, (int)(227 * 1024 / (rt_mid_doubles() * 8 + 1024))); }
static int rt_svd_threads(int cls) {
    static int t[4] = {0, 0, 0, 0};
    if (!t[0]) { t[0] = rt_tune("TNSP_RT_SVD_T0", kQBigThreads); t[1] = rt_tune("TNSP_RT_SVD_T1", kQSmallThreads); t[2] = rt_tune("TNSP_RT_SVD_T2", kQSmallThreads);
                 t[3] = rt_tune("TNSP_RT_SVD_T3", kRtTinyThreads); }
    return t[cls];
}
static int rt_qr_threads(int cls) {
    static int t[4] = {0, 0, 0, 0};
    if (!t[0]) { t[0] = rt_tune("TNSP_RT_QR_T0", kQBigThreads); t[1] = rt_tune("TNSP_RT_QR_T1", kQSmallThreads); t[2] = rt_tune("TNSP_RT_QR_T2", kQSmallThreads);
                 t[3] = rt_tune("TNSP_RT_QR_T3", kRtTinyThreads); }
    return t[cls];
}
static int rt_queue_prepare(int nb, int64_t per_cta_scratch, cudaStream_t st, int64_t& qcap) {
    qcap = (int64_t)nb * RT_SMAX;
    if (!g_qws.qctl && cudaMalloc(&g_qws.qctl, 8 * sizeof(int)) != cudaSuccess) { set_error("sector queue: cudaMalloc"); return 1; }
    if (!grow(g_qws.qitems, g_qws.qitems_cap, kRtClasses * qcap) || !grow(g_qws.scratch, g_qws.scratch_cap, per_cta_scratch * kSMs)) {
        set_error("sector queue: cudaMalloc of the workspace failed");
        return 1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(rt_qr_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtQrWarps * kRtQrWarpDoubles * 8);
        cudaFuncSetAttribute(rt_svd_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtSvdWarps * kRtSvdWarpDoubles * 8);
        cudaFuncSetAttribute(rt_qr_work_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(rt_qr_work_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(rt_svd_work_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(rt_svd_work_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQBigDoubles * 8);
        cudaFuncSetAttribute(rt_svd_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set = true;
    }
    return 0;
}

extern "C" int64_t tnsp_rt_factor_ws_ints(int64_t kfull) { return rt_ws_stride(kfull); }
extern "C" int64_t tnsp_rt_svd_work_doubles(int64_t M, int64_t N) {
    const int64_t k = M < N ? M : N;
    return k + M * k + k * N + 8;
}

static int64_t rt_scratch_need(int64_t M, int64_t N, int kind) {
    if (kind == 2) {
        const int64_t p = M >= N ? M : N, q = M >= N ? N : M;
        const int64_t full = svd_sector_need(p, q);
        return full > kQBigDoubles ? ((full + 9) & ~(int64_t)1) : 0;
    }
    const int64_t full = qr_sector_need(M, N);
    return full > kQBigDoubles ? full + 8 : 0;
}

extern "C" int tnsp_rt_factor_plan(const tnsp_rt_form* f, int kind, int fsign_rs, const int32_t* t1, int t1_stride, int t1s, int64_t kdim,
                                   int32_t* labels, int32_t* ws, int64_t ws_stride, int nb, void* stream) {
    if (nb == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t qcap;
    if (rt_queue_prepare(nb, rt_scratch_need(f->M, f->N, kind), st, qcap)) return 1;
    cudaMemsetAsync(g_qws.qctl, 0, 8 * sizeof(int), st);
    rt_factor_plan_kernel<<<(nb + 127) / 128, 128, 0, st>>>(to_form(f), kind, fsign_rs, t1, t1_stride, t1s, (int)kdim, labels, ws, ws_stride,
                                                            g_qws.qctl, g_qws.qitems, qcap, nb, rt_stats_ptr(), rt_mid_doubles());
    return check_launch("tnsp_rt_factor_plan");
}

extern "C" int tnsp_rt_qr_f64(const tnsp_rt_form* f, int fsign_rs, const int32_t* t1, int t1_stride, int t1s, const int32_t* bond,
                              int64_t bond_stride, const int32_t* m_first, double* first, int64_t first_stride, const int32_t* m_second,
                              double* second, int64_t second_stride, int nb, void* stream) {
    if (nb == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const RtForm F = to_form(f);
    const int64_t qcap = (int64_t)nb * RT_SMAX;
    const int64_t per_cta = rt_scratch_need(f->M, f->N, 0);
    if (kRtUseWarpClass) {
        rt_qr_warp_kernel<<<kSMs, 32 * kRtQrWarps, kRtQrWarps * kRtQrWarpDoubles * 8, st>>>(F, fsign_rs, t1, t1_stride, t1s, bond, bond_stride, m_first,
                                                                                            first, first_stride, m_second, second, second_stride,
                                                                                            g_qws.qctl, g_qws.qitems, qcap);
        if (check_launch("tnsp_rt_qr_f64(warp class)")) return 1;
        if (rt_qr_warp_need(f->M, f->N, f->M < f->N ? f->M : f->N) <= kRtQrWarpDoubles) return 0;     // nothing can be larger
    }
    if (qr_sector_need(f->M, f->N) > kQSmallDoubles) {
        rt_qr_work_kernel<true><<<kSMs, rt_qr_threads(0), kQBigDoubles * 8, st>>>(F, fsign_rs, t1, t1_stride, t1s, bond, bond_stride, m_first, first,
                                                                               first_stride, m_second, second, second_stride, g_qws.qctl,
                                                                               g_qws.qitems, qcap, 0, kQBigDoubles, g_qws.scratch, per_cta);
        if (check_launch("tnsp_rt_qr_f64(big)")) return 1;
    }
    if (qr_sector_need(f->M, f->N) > rt_mid_doubles()) {
        rt_qr_work_kernel<false><<<3 * kSMs, rt_qr_threads(1), kQSmallDoubles * 8, st>>>(F, fsign_rs, t1, t1_stride, t1s, bond, bond_stride, m_first,
                                                                                       first, first_stride, m_second, second, second_stride,
                                                                                       g_qws.qctl, g_qws.qitems, qcap, 1, kQSmallDoubles, nullptr, 0);
        if (check_launch("tnsp_rt_qr_f64(72 KiB class)")) return 1;
    }
    if (!kRtUseWarpClass) {
        rt_qr_work_kernel<false><<<8 * kSMs, rt_qr_threads(3), kRtTinyDoubles * 8, st>>>(F, fsign_rs, t1, t1_stride, t1s, bond, bond_stride, m_first,
                                                                                       first, first_stride, m_second, second, second_stride,
                                                                                       g_qws.qctl, g_qws.qitems, qcap, 3, kRtTinyDoubles, nullptr, 0);
        if (check_launch("tnsp_rt_qr_f64(tiny class)")) return 1;
        if (qr_sector_need(f->M, f->N) <= kRtTinyDoubles) return 0;
    }
    rt_qr_work_kernel<false><<<rt_mid_ctas() * kSMs, rt_qr_threads(2), rt_mid_doubles() * 8, st>>>(F, fsign_rs, t1, t1_stride, t1s, bond, bond_stride,
                                                                                 m_first, first, first_stride, m_second, second, second_stride,
                                                                                 g_qws.qctl, g_qws.qitems, qcap, 2, rt_mid_doubles(), nullptr, 0);
    return check_launch("tnsp_rt_qr_f64(55 KiB class)");
}

extern "C" int tnsp_rt_svd_work_f64(const tnsp_rt_form* f, double* work, int64_t work_stride, const int32_t* ws, int64_t ws_stride, int nb,
                                    void* stream) {
    if (nb == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const RtForm F = to_form(f);
    const int64_t qcap = (int64_t)nb * RT_SMAX;
    const int64_t per_cta = rt_scratch_need(f->M, f->N, 2);
    const int kfull = (int)(f->M < f->N ? f->M : f->N);
    const int64_t p = f->M >= f->N ? f->M : f->N, q = f->M >= f->N ? f->N : f->M;
    const int64_t full = svd_sector_need(p, q);
    if (kRtUseWarpClass) {
        rt_svd_warp_kernel<<<kSMs, 32 * kRtSvdWarps, kRtSvdWarps * kRtSvdWarpDoubles * 8, st>>>(F, work, work_stride, ws, ws_stride, kfull, g_qws.qctl,
                                                                                                g_qws.qitems, qcap);
        if (check_launch("tnsp_rt_svd_work_f64(warp class)")) return 1;
        if (q <= kWarpSectorMax && rt_svd_warp_need(p, q) <= kRtSvdWarpDoubles) return 0;
    }
    if (full > kQSmallDoubles) {
        rt_svd_work_kernel<true><<<kSMs, rt_svd_threads(0), kQBigDoubles * 8, st>>>(F, work, work_stride, ws, ws_stride, kfull, g_qws.qctl, g_qws.qitems,
                                                                                qcap, 0, kQBigDoubles, g_qws.scratch, per_cta, rt_jacobi_variant());
        if (check_launch("tnsp_rt_svd_work_f64(big)")) return 1;
    }
    if (full > rt_mid_doubles()) {
        rt_svd_work_kernel<false><<<3 * kSMs, rt_svd_threads(1), kQSmallDoubles * 8, st>>>(F, work, work_stride, ws, ws_stride, kfull, g_qws.qctl,
                                                                                        g_qws.qitems, qcap, 1, kQSmallDoubles, nullptr, 0, rt_jacobi_variant());
        if (check_launch("tnsp_rt_svd_work_f64(72 KiB class)")) return 1;
    }
    if (!kRtUseWarpClass) {
        rt_svd_work_kernel<false><<<8 * kSMs, rt_svd_threads(3), kRtTinyDoubles * 8, st>>>(F, work, work_stride, ws, ws_stride, kfull, g_qws.qctl,
                                                                                        g_qws.qitems, qcap, 3, kRtTinyDoubles, nullptr, 0, rt_jacobi_variant());
        if (check_launch("tnsp_rt_svd_work_f64(tiny class)")) return 1;
        if (full <= kRtTinyDoubles) return 0;
    }
    rt_svd_work_kernel<false><<<rt_mid_ctas() * kSMs, rt_svd_threads(2), rt_mid_doubles() * 8, st>>>(F, work, work_stride, ws, ws_stride, kfull,
                                                                                  g_qws.qctl, g_qws.qitems, qcap, 2, rt_mid_doubles(), nullptr, 0, rt_jacobi_variant());
    return check_launch("tnsp_rt_svd_work_f64(55 KiB class)");
}

extern "C" int tnsp_rt_svd_finish_f64(const tnsp_rt_form* f, int fsign_rs, const int32_t* t1, int t1_stride, int t1s, int64_t kdim,
                                      int64_t remain_cut, double relative_cut, const double* work, int64_t work_stride, int32_t* labels,
                                      int32_t* ws, int64_t ws_stride, int nb, void* stream) {
    if (nb == 0) return 0;
    const int kfull = (int)(f->M < f->N ? f->M : f->N);
    const size_t smem = (size_t)kfull * 12 + 16;
    if (smem > 96 * 1024) { set_error("tnsp_rt_svd_finish_f64: bond beyond 8190 singular values per chain"); return 1; }
    rt_svd_finish_kernel<<<nb, 256, smem, (cudaStream_t)stream>>>(to_form(f), fsign_rs, t1, t1_stride, t1s, (int)kdim, kfull, remain_cut, relative_cut,
                                                                  work, work_stride, labels, ws, ws_stride);
    return check_launch("tnsp_rt_svd_finish_f64");
}

extern "C" int tnsp_rt_svd_scatter_f64(const tnsp_rt_form* f, int fsign_rs, const int32_t* t1, int t1_stride, int t1s, const int32_t* bond,
                                       int64_t bond_stride, const int32_t* m_first, double* first, int64_t first_stride, const int32_t* m_s,
                                       double* s, int64_t s_stride, const int32_t* m_second, double* second, int64_t second_stride,
                                       const double* work, int64_t work_stride, const int32_t* ws, int64_t ws_stride, int nb, void* stream) {
    if (nb == 0) return 0;
    const int kfull = (int)(f->M < f->N ? f->M : f->N);
    rt_svd_scatter_kernel<<<dim3(nb, 4), 256, 0, (cudaStream_t)stream>>>(to_form(f), fsign_rs, t1, t1_stride, t1s, bond, bond_stride, m_first, first,
                                                                        first_stride, m_s, s, s_stride, m_second, second, second_stride, work,
                                                                        work_stride, ws, ws_stride, kfull);
    return check_launch("tnsp_rt_svd_scatter_f64");
}

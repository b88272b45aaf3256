// K1: symmetry-sector pairing + transpose ("pack") kernel, and the small streaming kernels.
//
// One launch moves every block of one edge operation (split / reverse / transpose / merge with the
// fermi sign folded in) for all chains of a batch.  The descriptor table is produced on the host
// from the reference's integer rules (TAT/include/TAT/implement/edge_operator.hpp:486-690); the
// copy loop this replaces is utility/multidimension_span.hpp:250-383.
//
// Roofline: HBM bound, algorithmic bytes = 2 * 8 B * elements (read once, write once).
#include "common.cuh"

namespace tnsp {

constexpr int R = TNSP_PACK_MAX_RANK;
constexpr int COLS = TNSP_PACK_COLS;

// Generic path: thread <-> destination element (coalesced stores), source gathered through the
// descriptor strides; the index decode is amortised over the chains handled by the thread.
__global__ void __launch_bounds__(256) pack_generic_kernel(
    const int64_t* __restrict__ desc, const int64_t* __restrict__ estart, int n_desc, int64_t total,
    const double* __restrict__ src, int64_t sbs, double* __restrict__ dst, int64_t dbs, int nb, int b_per_block) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= total) return;
    int lo = 0, hi = n_desc;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(estart + mid) <= e) lo = mid; else hi = mid;
    }
    const int64_t* d = desc + (int64_t)lo * COLS;
    int64_t so = __ldg(d), dof = __ldg(d + 1);
    const int64_t sr = __ldg(d + 2);
    const bool neg = sr & 1;
    const int rank = (int)(sr >> 1);
    uint64_t local = (uint64_t)(e - __ldg(estart + lo));
    if (local <= 0xffffffffull) {
        uint32_t l32 = (uint32_t)local;
        for (int i = rank - 1; i >= 0; --i) {
            const uint32_t dim = (uint32_t)__ldg(d + 3 + i);
            const uint32_t q = l32 / dim, r = l32 - q * dim;
            so += (int64_t)r * __ldg(d + 3 + R + i);
            dof += (int64_t)r * __ldg(d + 3 + 2 * R + i);
            l32 = q;
        }
    } else {
        for (int i = rank - 1; i >= 0; --i) {
            const uint64_t dim = (uint64_t)__ldg(d + 3 + i);
            const uint64_t q = local / dim, r = local - q * dim;
            so += (int64_t)r * __ldg(d + 3 + R + i);
            dof += (int64_t)r * __ldg(d + 3 + 2 * R + i);
            local = q;
        }
    }
    const int b0 = blockIdx.y * b_per_block;
    const int b1 = min(nb, b0 + b_per_block);
    const double* s = src + so + (int64_t)b0 * sbs;
    double* t = dst + dof + (int64_t)b0 * dbs;
    if (neg) {
        for (int b = b0; b < b1; ++b, s += sbs, t += dbs) *t = -__ldg(s);
    } else {
        for (int b = b0; b < b1; ++b, s += sbs, t += dbs) *t = __ldg(s);
    }
}

// Tiled path for a single 2-D transposing descriptor replicated over an outer axis:
// dst[o][j][i] = src[o][i][j] style moves where the source-contiguous axis differs from the
// destination-contiguous axis.  32x32 tiles staged through shared memory (+1 padding) so both the
// loads and the stores are coalesced.
struct Tile2D {
    int64_t src_off, dst_off;
    int64_t n_outer, n_i, n_j;          // dst order: [outer][j... ] see below
    int64_t s_outer, s_i, s_j;          // source strides (s_j == 1)
    int64_t d_outer, d_i, d_j;          // destination strides (d_i == 1)
    int neg;
};

__global__ void __launch_bounds__(256) pack_tiled_kernel(Tile2D t, const double* __restrict__ src, int64_t sbs,
                                                         double* __restrict__ dst, int64_t dbs, int nb) {
    __shared__ double tile[32][33];
    const int64_t tiles_i = (t.n_i + 31) >> 5, tiles_j = (t.n_j + 31) >> 5;
    const int64_t tiles_per_outer = tiles_i * tiles_j;
    const int64_t n_tiles = tiles_per_outer * t.n_outer;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int64_t w = blockIdx.x; w < n_tiles * nb; w += gridDim.x) {
        const int64_t b = w / n_tiles;
        int64_t r = w - b * n_tiles;
        const int64_t o = r / tiles_per_outer;
        r -= o * tiles_per_outer;
        const int64_t ti = r / tiles_j, tj = r - ti * tiles_j;
        const double* s = src + b * sbs + t.src_off + o * t.s_outer;
        double* d = dst + b * dbs + t.dst_off + o * t.d_outer;
        const int64_t i0 = ti << 5, j0 = tj << 5;
        // load: j contiguous in the source
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = i0 + ty + k * 8, j = j0 + tx;
            if (i < t.n_i && j < t.n_j) tile[ty + k * 8][tx] = __ldg(s + i * t.s_i + j * t.s_j);
        }
        __syncthreads();
        // store: i contiguous in the destination
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t j = j0 + ty + k * 8, i = i0 + tx;
            if (i < t.n_i && j < t.n_j) {
                const double v = tile[tx][ty + k * 8];
                d[i * t.d_i + j * t.d_j] = t.neg ? -v : v;
            }
        }
        __syncthreads();
    }
}

// ---- streaming kernels -------------------------------------------------------------------------
__global__ void norm_kernel(const double* __restrict__ x, int64_t xbs, int64_t size, int kind, double* __restrict__ out) {
    __shared__ double red[32];
    const double* p = x + (int64_t)blockIdx.x * xbs;
    double acc = 0.0;
    if (kind == -1) {
        for (int64_t i = threadIdx.x; i < size; i += blockDim.x) acc = fmax(acc, fabs(p[i]));
        acc = block_max(acc, red);
    } else if (kind == 1) {
        for (int64_t i = threadIdx.x; i < size; i += blockDim.x) acc += fabs(p[i]);
        acc = block_sum(acc, red);
    } else {
        for (int64_t i = threadIdx.x; i < size; i += blockDim.x) acc += p[i] * p[i];
        acc = sqrt(block_sum(acc, red));
    }
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

__global__ void scale_kernel(const double* __restrict__ x, int64_t xbs, const double* __restrict__ alpha, int64_t as, int op,
                             double* __restrict__ y, int64_t ybs, int64_t size, int nb) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= size) return;
    for (int b = blockIdx.y; b < nb; b += gridDim.y) {
        const double a = alpha[(int64_t)b * as];
        const double v = x[(int64_t)b * xbs + i];
        y[(int64_t)b * ybs + i] = op ? v / a : v * a;
    }
}

__global__ void binary_kernel(const double* __restrict__ a, int64_t abs_, const double* __restrict__ b, int64_t bbs, int op,
                              double* __restrict__ z, int64_t zbs, int64_t size, int nb) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= size) return;
    for (int c = blockIdx.y; c < nb; c += gridDim.y) {
        const double u = a[(int64_t)c * abs_ + i], v = b[(int64_t)c * bbs + i];
        double r;
        switch (op) {
            case 0: r = u + v; break;
            case 1: r = u - v; break;
            case 2: r = u * v; break;
            default: r = u / v; break;
        }
        z[(int64_t)c * zbs + i] = r;
    }
}

__global__ void unary_kernel(const double* __restrict__ a, int64_t abs_, int op, double* __restrict__ z, int64_t zbs, int64_t size, int nb) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= size) return;
    for (int c = blockIdx.y; c < nb; c += gridDim.y) {
        const double u = a[(int64_t)c * abs_ + i];
        double r;
        switch (op) {
            case 0: r = sqrt(fabs(u)); break;
            case 1: r = (u == 0.0) ? 0.0 : 1.0 / u; break;
            case 2: r = -u; break;
            default: r = fabs(u); break;
        }
        z[(int64_t)c * zbs + i] = r;
    }
}

// Fused log-derivative accumulation over the chains of a batch, in two deterministic stages:
// stage 1: CTA (x, y) sums chains [y*chunk, (y+1)*chunk) of elements x*128.. into partial[y]
// stage 2: partials are added in fixed order into the running Delta / E*Delta accumulators.
__global__ void grad_partial_kernel(const double* __restrict__ holes, int64_t hbs, const double* __restrict__ weight,
                                    const double* __restrict__ energy, double* __restrict__ partial, int64_t size, int nb, int chunk) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= size) return;
    const int b0 = blockIdx.y * chunk, b1 = min(nb, b0 + chunk);
    double d = 0.0, ed = 0.0;
    for (int b = b0; b < b1; ++b) {
        const double w = weight[b];
        if (w == 0.0) continue;   // dead chain (zero amplitude): its hole is 0/0
        const double h = holes[(int64_t)b * hbs + i] * w;
        d += h;
        ed += h * energy[b];
    }
    partial[((int64_t)blockIdx.y * 2) * size + i] = d;
    partial[((int64_t)blockIdx.y * 2 + 1) * size + i] = ed;
}

__global__ void grad_final_kernel(const double* __restrict__ partial, int nchunks, double* __restrict__ delta,
                                  double* __restrict__ edelta, int64_t size) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= size) return;
    double d = 0.0, ed = 0.0;
    for (int y = 0; y < nchunks; ++y) {
        d += partial[((int64_t)y * 2) * size + i];
        ed += partial[((int64_t)y * 2 + 1) * size + i];
    }
    delta[i] += d;
    edelta[i] += ed;
}

__global__ void block_sign_kernel(const int64_t* __restrict__ blk, int nblk, const double* __restrict__ x, int64_t xbs,
                                  double* __restrict__ y, int64_t ybs, int64_t size, int nb) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= size) return;
    int lo = 0, hi = nblk;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (blk[mid * 3] <= i) lo = mid; else hi = mid;
    }
    const bool neg = blk[lo * 3 + 2] != 0;
    for (int b = blockIdx.y; b < nb; b += gridDim.y) {
        const double v = x[(int64_t)b * xbs + i];
        y[(int64_t)b * ybs + i] = neg ? -v : v;
    }
}

__global__ void gather_rows_kernel(const double* __restrict__ src, int64_t row, const int32_t* __restrict__ index,
                                   double* __restrict__ dst, int64_t dbs, int nb) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= row) return;
    for (int b = blockIdx.y; b < nb; b += gridDim.y) dst[(int64_t)b * dbs + i] = src[(int64_t)index[b] * row + i];
}

__global__ void select_kernel(const uint8_t* __restrict__ mask, const double* __restrict__ a, int64_t abs_,
                              const double* __restrict__ b_, int64_t bbs, double* __restrict__ dst, int64_t dbs, int64_t size, int nb) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= size) return;
    for (int b = blockIdx.y; b < nb; b += gridDim.y)
        dst[(int64_t)b * dbs + i] = mask[b] ? a[(int64_t)b * abs_ + i] : b_[(int64_t)b * bbs + i];
}

__global__ void diag_scatter_kernel(const int64_t* __restrict__ blk, int nblk, const double* __restrict__ s, int64_t sbs,
                                    double* __restrict__ dst, int64_t dbs, int nb) {
    const int j = blockIdx.x;
    const int64_t s_off = blk[j * 4], d_off = blk[j * 4 + 1], r = blk[j * 4 + 2];
    const bool neg = blk[j * 4 + 3] != 0;
    for (int b = blockIdx.y; b < nb; b += gridDim.y) {
        const double* sp = s + (int64_t)b * sbs + s_off;
        double* dp = dst + (int64_t)b * dbs + d_off;
        for (int64_t i = threadIdx.x; i < r; i += blockDim.x) dp[i * (r + 1)] = neg ? -sp[i] : sp[i];
    }
}

static inline dim3 grid_1d_batch(int64_t size, int nb) {
    const int64_t gx = (size + 255) / 256;
    int64_t gy = (4 * kSMs + gx - 1) / gx;
    if (gy > nb) gy = nb;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)gx, (unsigned)gy);
}

}  // namespace tnsp

using namespace tnsp;

extern "C" int tnsp_pack_f64(const int64_t* desc, const int64_t* estart, int n_desc, int64_t total, const double* src,
                             int64_t sbs, double* dst, int64_t dbs, int nb, void* stream) {
    if (n_desc == 0 || total == 0 || nb == 0) return 0;
    const int64_t gx = (total + 255) / 256;
    // enough CTAs to fill the machine, but keep several chains per thread to amortise the index decode
    int64_t gy = (8 * kSMs + gx - 1) / gx;
    if (gy > nb) gy = nb;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const int bpb = (int)((nb + gy - 1) / gy);
    gy = (nb + bpb - 1) / bpb;
    pack_generic_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(desc, estart, n_desc, total, src, sbs, dst,
                                                                                           dbs, nb, bpb);
    return check_launch("tnsp_pack_f64");
}

// Host-side descriptor variant used when the plan is a single 2-D transposition (the planner
// detects it); exported for the Python layer as an optimisation of the same operation.
extern "C" int tnsp_pack_tiled_f64(int64_t src_off, int64_t dst_off, int64_t n_outer, int64_t n_i, int64_t n_j, int64_t s_outer,
                                   int64_t s_i, int64_t d_outer, int64_t d_j, int neg, const double* src, int64_t sbs, double* dst,
                                   int64_t dbs, int nb, void* stream) {
    if (n_outer * n_i * n_j == 0 || nb == 0) return 0;
    Tile2D t{src_off, dst_off, n_outer, n_i, n_j, s_outer, s_i, 1, d_outer, 1, d_j, neg};
    const int64_t tiles = ((n_i + 31) / 32) * ((n_j + 31) / 32) * n_outer * nb;
    const int64_t g = tiles < (int64_t)kSMs * 16 ? tiles : (int64_t)kSMs * 16;
    pack_tiled_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(t, src, sbs, dst, dbs, nb);
    return check_launch("tnsp_pack_tiled_f64");
}

extern "C" int tnsp_norm_f64(const double* x, int64_t xbs, int64_t size, int kind, double* out, int nb, void* stream) {
    if (nb == 0) return 0;
    int threads = size >= 1024 ? 256 : (size >= 128 ? 128 : 32);
    norm_kernel<<<nb, threads, 0, (cudaStream_t)stream>>>(x, xbs, size, kind, out);
    return check_launch("tnsp_norm_f64");
}

extern "C" int tnsp_scale_f64(const double* x, int64_t xbs, const double* alpha, int64_t as, int op, double* y, int64_t ybs,
                              int64_t size, int nb, void* stream) {
    if (nb == 0 || size == 0) return 0;
    scale_kernel<<<grid_1d_batch(size, nb), 256, 0, (cudaStream_t)stream>>>(x, xbs, alpha, as, op, y, ybs, size, nb);
    return check_launch("tnsp_scale_f64");
}

extern "C" int tnsp_binary_f64(const double* a, int64_t abs_, const double* b, int64_t bbs, int op, double* z, int64_t zbs,
                               int64_t size, int nb, void* stream) {
    if (nb == 0 || size == 0) return 0;
    binary_kernel<<<grid_1d_batch(size, nb), 256, 0, (cudaStream_t)stream>>>(a, abs_, b, bbs, op, z, zbs, size, nb);
    return check_launch("tnsp_binary_f64");
}

extern "C" int tnsp_unary_f64(const double* a, int64_t abs_, int op, double* z, int64_t zbs, int64_t size, int nb, void* stream) {
    if (nb == 0 || size == 0) return 0;
    unary_kernel<<<grid_1d_batch(size, nb), 256, 0, (cudaStream_t)stream>>>(a, abs_, op, z, zbs, size, nb);
    return check_launch("tnsp_unary_f64");
}

extern "C" int tnsp_grad_accumulate_f64(const double* holes, int64_t hbs, const double* weight, const double* energy, double* delta,
                                        double* edelta, int64_t size, int nb, void* stream) {
    if (nb == 0 || size == 0) return 0;
    static double* scratch = nullptr;
    static int64_t scratch_cap = 0;
    const int64_t gx = (size + 127) / 128;
    int64_t nchunks = (2 * kSMs + gx - 1) / gx;       // enough CTAs to fill the machine
    if (nchunks > nb) nchunks = nb;
    if (nchunks < 1) nchunks = 1;
    const int chunk = (int)((nb + nchunks - 1) / nchunks);
    nchunks = (nb + chunk - 1) / chunk;
    const int64_t need = nchunks * 2 * size;
    if (need > scratch_cap) {
        if (scratch) cudaFree(scratch);
        scratch_cap = need * 2;
        if (cudaMalloc(&scratch, sizeof(double) * scratch_cap) != cudaSuccess) { set_error("tnsp_grad_accumulate_f64: cudaMalloc"); return 1; }
    }
    cudaStream_t st = (cudaStream_t)stream;
    grad_partial_kernel<<<dim3((unsigned)gx, (unsigned)nchunks), 128, 0, st>>>(holes, hbs, weight, energy, scratch, size, nb, chunk);
    if (check_launch("tnsp_grad_accumulate_f64(partial)")) return 1;
    grad_final_kernel<<<(unsigned)gx, 128, 0, st>>>(scratch, (int)nchunks, delta, edelta, size);
    return check_launch("tnsp_grad_accumulate_f64(final)");
}

extern "C" int tnsp_block_sign_f64(const int64_t* blk, int nblk, const double* x, int64_t xbs, double* y, int64_t ybs, int64_t size,
                                   int nb, void* stream) {
    if (nb == 0 || size == 0 || nblk == 0) return 0;
    block_sign_kernel<<<grid_1d_batch(size, nb), 256, 0, (cudaStream_t)stream>>>(blk, nblk, x, xbs, y, ybs, size, nb);
    return check_launch("tnsp_block_sign_f64");
}

extern "C" int tnsp_gather_rows_f64(const double* src, int64_t row, const int32_t* index, double* dst, int64_t dbs, int nb,
                                    void* stream) {
    if (nb == 0 || row == 0) return 0;
    gather_rows_kernel<<<grid_1d_batch(row, nb), 256, 0, (cudaStream_t)stream>>>(src, row, index, dst, dbs, nb);
    return check_launch("tnsp_gather_rows_f64");
}

extern "C" int tnsp_select_f64(const uint8_t* mask, const double* a, int64_t abs_, const double* b_, int64_t bbs, double* dst,
                               int64_t dbs, int64_t size, int nb, void* stream) {
    if (nb == 0 || size == 0) return 0;
    select_kernel<<<grid_1d_batch(size, nb), 256, 0, (cudaStream_t)stream>>>(mask, a, abs_, b_, bbs, dst, dbs, size, nb);
    return check_launch("tnsp_select_f64");
}

extern "C" int tnsp_diag_scatter_f64(const int64_t* blk, int nblk, const double* s, int64_t sbs, double* dst, int64_t dbs, int nb,
                                     void* stream) {
    if (nb == 0 || nblk == 0) return 0;
    int gy = nb > 65535 ? 65535 : nb;
    diag_scatter_kernel<<<dim3(nblk, gy), 64, 0, (cudaStream_t)stream>>>(blk, nblk, s, sbs, dst, dbs, nb);
    return check_launch("tnsp_diag_scatter_f64");
}

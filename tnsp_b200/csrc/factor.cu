// K3/K4: batched Householder QR/LQ and batched one-sided Jacobi SVD over (sector x chain), plus the
// reference's greedy cross-sector truncation.
//
// Replaces the per-sector LAPACK calls of the reference:
//   qr.hpp:178-304   (?geqrf + ?orgqr  /  ?gelqf + ?orglq, explicit thin Q)
//   svd.hpp:104-211  (?gesvd 'S','S')
//   svd.hpp:429-481  (global greedy cut across sectors)
// One CTA owns one (sector, chain) matrix; the working set lives in shared memory when it fits
// (<= kSmemDoubles) and in an L2-resident global scratch otherwise.
//
// Roofline: HBM/L2 bound streaming of small matrices; algorithmic bytes are stated in DESIGN.md.
#include "common.cuh"

namespace tnsp {

constexpr int kFactorThreads = 256;
constexpr int kSmemDoubles = 24 * 1024;   // 192 KiB of dynamic shared memory for the working set

// ------------------------------------------------------------------------------------------------
// QR: logical matrix X (p x q), element (i, j) at W[i * rs + j * cs].
//   use_qr : X = M        (p = m, q = n, rs = n, cs = 1)   out1 = Q (m x k), out2 = R (k x n)
//   LQ     : X = M^T      (p = n, q = m, rs = 1, cs = n)   out1 = L = R^T (m x k), out2 = Q^T (k x n)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFactorThreads) qr_kernel(const int64_t* __restrict__ sect, double* __restrict__ a, int64_t abs_,
                                                            double* __restrict__ out1, int64_t o1bs, double* __restrict__ out2,
                                                            int64_t o2bs, int use_qr, int nb) {
    extern __shared__ double smem[];
    __shared__ double red[32];
    __shared__ double sh_tau, sh_scale, sh_beta;
    const int64_t* sc = sect + (int64_t)blockIdx.x * TNSP_SECT_COLS;
    const int64_t m = sc[0], n = sc[1], k = sc[2];
    if (m * n == 0) return;
    const int64_t p = use_qr ? m : n, q = use_qr ? n : m;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* tau = smem;                 // [k]
    double* buf = smem + k;             // [p*q] when it fits
    const bool in_smem = (p * q + k) <= kSmemDoubles;

    for (int b = blockIdx.y; b < nb; b += gridDim.y) {
        double* A = a + (int64_t)b * abs_ + sc[3];
        double* W;
        int64_t rs, cs;
        if (in_smem) {
            // stage row-major X into shared memory (X[i][j] at buf[i*q + j])
            W = buf; rs = q; cs = 1;
            if (use_qr) {
                for (int64_t e = tid; e < p * q; e += nt) buf[e] = A[e];
            } else {
                // X = M^T : X[i][j] = M[j][i] = A[j*n + i]; read coalesced over A, scatter into smem
                for (int64_t e = tid; e < m * n; e += nt) {
                    const int64_t j = e / n, i = e - j * n;
                    buf[i * q + j] = A[e];
                }
            }
        } else {
            W = A;
            if (use_qr) { rs = n; cs = 1; } else { rs = 1; cs = n; }
        }
        __syncthreads();

        for (int64_t j = 0; j < k; ++j) {
            // Householder vector of column j (LAPACK dlarfg convention: v0 = 1)
            double part = 0.0;
            for (int64_t i = j + 1 + tid; i < p; i += nt) {
                const double v = W[i * rs + j * cs];
                part += v * v;
            }
            const double xnorm2 = block_sum(part, red);
            if (tid == 0) {
                const double alpha = W[j * rs + j * cs];
                if (xnorm2 == 0.0) {
                    sh_tau = 0.0; sh_scale = 0.0; sh_beta = alpha;
                } else {
                    const double beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
                    sh_tau = (beta - alpha) / beta;
                    sh_scale = 1.0 / (alpha - beta);
                    sh_beta = beta;
                }
                tau[j] = sh_tau;
                W[j * rs + j * cs] = sh_beta;
            }
            __syncthreads();
            const double tj = sh_tau, scale = sh_scale;
            if (tj != 0.0) {
                for (int64_t i = j + 1 + tid; i < p; i += nt) W[i * rs + j * cs] *= scale;
                __syncthreads();
                // trailing update, one thread per column (rows serial)
                for (int64_t cidx = j + 1 + tid; cidx < q; cidx += nt) {
                    double w = W[j * rs + cidx * cs];
                    for (int64_t i = j + 1; i < p; ++i) w += W[i * rs + j * cs] * W[i * rs + cidx * cs];
                    w *= tj;
                    W[j * rs + cidx * cs] -= w;
                    for (int64_t i = j + 1; i < p; ++i) W[i * rs + cidx * cs] -= w * W[i * rs + j * cs];
                }
            }
            __syncthreads();
        }

        // R (k x q): upper trapezoid of W
        if (use_qr) {
            double* Rm = out2 + (int64_t)b * o2bs + sc[5];   // k x n
            for (int64_t e = tid; e < k * q; e += nt) {
                const int64_t i = e / q, j = e - i * q;
                Rm[e] = (j >= i) ? W[i * rs + j * cs] : 0.0;
            }
        } else {
            double* Lm = out1 + (int64_t)b * o1bs + sc[4];   // m x k, L[r][i] = R[i][r], q == m
            for (int64_t e = tid; e < q * k; e += nt) {
                const int64_t r = e / k, i = e - r * k;
                Lm[e] = (r >= i) ? W[i * rs + r * cs] : 0.0;
            }
        }
        __syncthreads();

        // explicit Q in place (LAPACK dorg2r), first k columns of W
        for (int64_t j = k - 1; j >= 0; --j) {
            const double tj = tau[j];
            // apply H(j) to W(j:p, j+1:k) from the left with v = [1; W(j+1:p, j)]
            for (int64_t cidx = j + 1 + tid; cidx < k; cidx += nt) {
                double w = W[j * rs + cidx * cs];
                for (int64_t i = j + 1; i < p; ++i) w += W[i * rs + j * cs] * W[i * rs + cidx * cs];
                w *= tj;
                W[j * rs + cidx * cs] -= w;
                for (int64_t i = j + 1; i < p; ++i) W[i * rs + cidx * cs] -= w * W[i * rs + j * cs];
            }
            __syncthreads();
            for (int64_t i = j + 1 + tid; i < p; i += nt) W[i * rs + j * cs] *= -tj;
            if (tid == 0) W[j * rs + j * cs] = 1.0 - tj;
            for (int64_t i = tid; i < j; i += nt) W[i * rs + j * cs] = 0.0;
            __syncthreads();
        }
        if (use_qr) {
            double* Qm = out1 + (int64_t)b * o1bs + sc[4];   // m x k
            for (int64_t e = tid; e < p * k; e += nt) {
                const int64_t i = e / k, j = e - i * k;
                Qm[e] = W[i * rs + j * cs];
            }
        } else {
            double* Qm = out2 + (int64_t)b * o2bs + sc[5];   // k x n, Q[j][i] = Qx[i][j], p == n
            for (int64_t e = tid; e < k * p; e += nt) {
                const int64_t j = e / p, i = e - j * p;
                Qm[e] = W[i * rs + j * cs];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// One-sided Jacobi (Hestenes) SVD.  Logical X (p x q, p >= q): X = M if m >= n else M^T.
// G[c][0..p) holds column c of X contiguously, V[c][0..q) column c of the accumulated rotations.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFactorThreads) svd_kernel(const int64_t* __restrict__ sect, const double* __restrict__ a,
                                                             int64_t abs_, double* __restrict__ out1, int64_t o1bs,
                                                             double* __restrict__ sv, int64_t sbs, double* __restrict__ out2,
                                                             int64_t o2bs, double* __restrict__ work, int64_t wbs,
                                                             const int64_t* __restrict__ work_off, int nb) {
    extern __shared__ double smem[];
    __shared__ int sh_rot;
    const int64_t* sc = sect + (int64_t)blockIdx.x * TNSP_SECT_COLS;
    const int64_t m = sc[0], n = sc[1], k = sc[2];
    if (m * n == 0) return;
    const bool tall = m >= n;
    const int64_t p = tall ? m : n, q = tall ? n : m;   // q == k
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int64_t need = q * p + q * q + 2 * q;
    const bool in_smem = need <= kSmemDoubles;

    for (int b = blockIdx.y; b < nb; b += gridDim.y) {
        const double* A = a + (int64_t)b * abs_ + sc[3];
        double* base = in_smem ? smem : (work + (int64_t)b * wbs + work_off[blockIdx.x]);
        double* G = base;
        double* V = base + q * p;
        double* sig = V + q * q;
        int* rnk = reinterpret_cast<int*>(sig + q);

        if (tall) {
            // G[c][r] = M[r][c]
            for (int64_t e = tid; e < m * n; e += nt) {
                const int64_t r = e / n, cidx = e - r * n;
                G[cidx * p + r] = A[e];
            }
        } else {
            // X = M^T: column c of X is row c of M (contiguous)
            for (int64_t e = tid; e < m * n; e += nt) G[e] = A[e];
        }
        for (int64_t e = tid; e < q * q; e += nt) V[e] = ((e / q) == (e % q)) ? 1.0 : 0.0;
        __syncthreads();

        const int64_t qe = q + (q & 1);   // even number of players
        const double tol = fmax(1e-15, sqrt((double)p) * 2.3e-16);
        for (int sweep = 0; sweep < 60 && q > 1; ++sweep) {
            if (tid == 0) sh_rot = 0;
            __syncthreads();
            for (int64_t round = 0; round < qe - 1; ++round) {
                for (int64_t pr = warp; pr < qe / 2; pr += nw) {
                    int64_t i, j;
                    if (pr == 0) { i = qe - 1; j = round; }
                    else { i = (round + pr) % (qe - 1); j = (round - pr + (qe - 1)) % (qe - 1); }
                    if (i >= q || j >= q) continue;
                    if (i > j) { const int64_t t = i; i = j; j = t; }
                    double* gi = G + i * p;
                    double* gj = G + j * p;
                    double aa = 0.0, bb = 0.0, cc = 0.0;
                    for (int64_t r = lane; r < p; r += 32) {
                        const double x = gi[r], y = gj[r];
                        aa += x * x; bb += y * y; cc += x * y;
                    }
                    aa = warp_sum(aa); bb = warp_sum(bb); cc = warp_sum(cc);
                    if (fabs(cc) > tol * sqrt(aa * bb) && aa * bb > 0.0) {
                        const double zeta = (bb - aa) / (2.0 * cc);
                        const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                        for (int64_t r = lane; r < p; r += 32) {
                            const double x = gi[r], y = gj[r];
                            gi[r] = cs * x - sn * y;
                            gj[r] = sn * x + cs * y;
                        }
                        double* vi = V + i * q;
                        double* vj = V + j * q;
                        for (int64_t r = lane; r < q; r += 32) {
                            const double x = vi[r], y = vj[r];
                            vi[r] = cs * x - sn * y;
                            vj[r] = sn * x + cs * y;
                        }
                        if (lane == 0) sh_rot = 1;
                    }
                }
                __syncthreads();
            }
            const int any = sh_rot;
            __syncthreads();
            if (!any) break;
        }

        // singular values and their descending order
        for (int64_t cidx = warp; cidx < q; cidx += nw) {
            double s2 = 0.0;
            for (int64_t r = lane; r < p; r += 32) s2 += G[cidx * p + r] * G[cidx * p + r];
            s2 = warp_sum(s2);
            if (lane == 0) sig[cidx] = sqrt(s2);
        }
        __syncthreads();
        for (int64_t cidx = tid; cidx < q; cidx += nt) {
            const double s = sig[cidx];
            int r = 0;
            for (int64_t o = 0; o < q; ++o) {
                const double so = sig[o];
                r += (so > s) || (so == s && o < cidx);
            }
            rnk[cidx] = r;
        }
        __syncthreads();
        double* S = sv + (int64_t)b * sbs + sc[6];
        double* O1 = out1 + (int64_t)b * o1bs + sc[4];   // m x k
        double* O2 = out2 + (int64_t)b * o2bs + sc[5];   // k x n
        for (int64_t cidx = tid; cidx < q; cidx += nt) S[rnk[cidx]] = sig[cidx];
        if (tall) {
            // U[r][jj] = G[c][r] / sigma_c ; Vt[jj][t] = V[c][t]
            for (int64_t e = tid; e < m * k; e += nt) {
                const int64_t cidx = e / m, r = e - cidx * m;
                const double s = sig[cidx];
                O1[r * k + rnk[cidx]] = s > 0.0 ? G[cidx * p + r] / s : 0.0;
            }
            for (int64_t e = tid; e < k * n; e += nt) {
                const int64_t cidx = e / n, t = e - cidx * n;
                O2[(int64_t)rnk[cidx] * n + t] = V[cidx * q + t];
            }
        } else {
            // U[t][jj] = V[c][t] ; Vt[jj][r] = G[c][r] / sigma_c
            for (int64_t e = tid; e < m * k; e += nt) {
                const int64_t cidx = e / m, t = e - cidx * m;
                O1[t * k + rnk[cidx]] = V[cidx * q + t];
            }
            for (int64_t e = tid; e < k * n; e += nt) {
                const int64_t cidx = e / n, r = e - cidx * n;
                const double s = sig[cidx];
                O2[(int64_t)rnk[cidx] * n + r] = s > 0.0 ? G[cidx * p + r] / s : 0.0;
            }
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------
// Warp-per-matrix variants for the small sector matrices that dominate the VMC path (cfg1: 16x16,
// 64x16, 4x64 ...).  One warp owns one (sector, chain) matrix in its private slice of shared memory
// (padded leading dimension -> conflict-free), so the factorization needs no block barrier at all;
// a CTA of kWarpsPerCta warps works on kWarpsPerCta chains at once.
// ------------------------------------------------------------------------------------------------
constexpr int kWarpsPerCta = 4;
constexpr int kWarpDoubles = 1536;   // 12 KiB per warp

__device__ __forceinline__ double group_sum(double v, int width) {
    for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__host__ __device__ inline int64_t qr_warp_need(int64_t p, int64_t q, int64_t k) { return p * (q | 1) + k; }

__global__ void __launch_bounds__(32 * kWarpsPerCta) qr_warp_kernel(const int64_t* __restrict__ sect, const double* __restrict__ a,
                                                                    int64_t abs_, double* __restrict__ out1, int64_t o1bs,
                                                                    double* __restrict__ out2, int64_t o2bs, int use_qr, int nb) {
    extern __shared__ double smem[];
    const int64_t* sc = sect + (int64_t)blockIdx.x * TNSP_SECT_COLS;
    const int m = (int)sc[0], n = (int)sc[1], k = (int)sc[2];
    if (m * n == 0) return;
    const int p = use_qr ? m : n, q = use_qr ? n : m;
    const int ld = q | 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* X = smem + (int64_t)warp * kWarpDoubles;   // X[i*ld + j]
    double* tau = X + p * ld;
    for (int b = blockIdx.y * kWarpsPerCta + warp; b < nb; b += gridDim.y * kWarpsPerCta) {
        const double* A = a + (int64_t)b * abs_ + sc[3];
        if (use_qr) {
            for (int e = lane; e < m * n; e += 32) X[(e / n) * ld + (e % n)] = A[e];
        } else {
            for (int e = lane; e < m * n; e += 32) X[(e % n) * ld + (e / n)] = A[e];   // X = M^T
        }
        __syncwarp();
        const bool by_rows = p >= q;   // lanes over rows (tall) or over columns (wide)
        for (int j = 0; j < k; ++j) {
            double part = 0.0;
            for (int i = j + 1 + lane; i < p; i += 32) { const double v = X[i * ld + j]; part += v * v; }
            const double xnorm2 = warp_sum(part);
            const double alpha = X[j * ld + j];
            double tj = 0.0, scale = 0.0, beta = alpha;
            if (xnorm2 != 0.0) {
                beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
                tj = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            __syncwarp();
            if (tj != 0.0) {
                for (int i = j + 1 + lane; i < p; i += 32) X[i * ld + j] *= scale;
                __syncwarp();
                if (by_rows) {
                    for (int c = j + 1; c < q; ++c) {
                        double w = 0.0;
                        for (int i = j + 1 + lane; i < p; i += 32) w += X[i * ld + j] * X[i * ld + c];
                        w = (warp_sum(w) + X[j * ld + c]) * tj;
                        for (int i = j + 1 + lane; i < p; i += 32) X[i * ld + c] -= w * X[i * ld + j];
                        __syncwarp();
                        if (lane == 0) X[j * ld + c] -= w;
                    }
                } else {
                    for (int c = j + 1 + lane; c < q; c += 32) {
                        double w = X[j * ld + c];
                        for (int i = j + 1; i < p; ++i) w += X[i * ld + j] * X[i * ld + c];
                        w *= tj;
                        X[j * ld + c] -= w;
                        for (int i = j + 1; i < p; ++i) X[i * ld + c] -= w * X[i * ld + j];
                    }
                }
            }
            if (lane == 0) { tau[j] = tj; X[j * ld + j] = beta; }
            __syncwarp();
        }
        if (use_qr) {
            double* Rm = out2 + (int64_t)b * o2bs + sc[5];
            for (int e = lane; e < k * q; e += 32) { const int i = e / q, j = e - i * q; Rm[e] = (j >= i) ? X[i * ld + j] : 0.0; }
        } else {
            double* Lm = out1 + (int64_t)b * o1bs + sc[4];
            for (int e = lane; e < q * k; e += 32) { const int r = e / k, i = e - r * k; Lm[e] = (r >= i) ? X[i * ld + r] : 0.0; }
        }
        __syncwarp();
        for (int j = k - 1; j >= 0; --j) {
            const double tj = tau[j];
            if (by_rows) {
                for (int c = j + 1; c < k; ++c) {
                    double w = 0.0;
                    for (int i = j + 1 + lane; i < p; i += 32) w += X[i * ld + j] * X[i * ld + c];
                    w = (warp_sum(w) + X[j * ld + c]) * tj;
                    for (int i = j + 1 + lane; i < p; i += 32) X[i * ld + c] -= w * X[i * ld + j];
                    __syncwarp();
                    if (lane == 0) X[j * ld + c] -= w;
                }
            } else {
                for (int c = j + 1 + lane; c < k; c += 32) {
                    double w = X[j * ld + c];
                    for (int i = j + 1; i < p; ++i) w += X[i * ld + j] * X[i * ld + c];
                    w *= tj;
                    X[j * ld + c] -= w;
                    for (int i = j + 1; i < p; ++i) X[i * ld + c] -= w * X[i * ld + j];
                }
            }
            __syncwarp();
            for (int i = j + 1 + lane; i < p; i += 32) X[i * ld + j] *= -tj;
            for (int i = lane; i < j; i += 32) X[i * ld + j] = 0.0;
            if (lane == 0) X[j * ld + j] = 1.0 - tj;
            __syncwarp();
        }
        if (use_qr) {
            double* Qm = out1 + (int64_t)b * o1bs + sc[4];
            for (int e = lane; e < p * k; e += 32) { const int i = e / k, j = e - i * k; Qm[e] = X[i * ld + j]; }
        } else {
            double* Qm = out2 + (int64_t)b * o2bs + sc[5];
            for (int e = lane; e < k * p; e += 32) { const int j = e / p, i = e - j * p; Qm[e] = X[i * ld + j]; }
        }
        __syncwarp();
    }
}

__host__ __device__ inline int64_t svd_warp_need(int64_t p, int64_t q) { return q * (p | 1) + q * (q | 1) + 2 * q; }

// Jacobi SVD, one warp per matrix; every column pair of a round-robin round is rotated by its own
// sub-warp group of `gs` lanes (gs = 32 / pairs-in-flight), all rounds separated by __syncwarp only.
__global__ void __launch_bounds__(32 * kWarpsPerCta) svd_warp_kernel(const int64_t* __restrict__ sect, const double* __restrict__ a,
                                                                     int64_t abs_, double* __restrict__ out1, int64_t o1bs,
                                                                     double* __restrict__ sv, int64_t sbs, double* __restrict__ out2,
                                                                     int64_t o2bs, int nb) {
    extern __shared__ double smem[];
    const int64_t* sc = sect + (int64_t)blockIdx.x * TNSP_SECT_COLS;
    const int m = (int)sc[0], n = (int)sc[1], k = (int)sc[2];
    if (m * n == 0) return;
    const bool tall = m >= n;
    const int p = tall ? m : n, q = tall ? n : m;
    const int ldp = p | 1, ldq = q | 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* G = smem + (int64_t)warp * kWarpDoubles;   // G[c*ldp + r]
    double* V = G + q * ldp;                            // V[c*ldq + t]
    double* sig = V + q * ldq;
    int* rnk = reinterpret_cast<int*>(sig + q);
    int gs = 32;                                        // lanes per pair: smallest power of two >= p, at least 4
    while (gs > 4 && (gs >> 1) >= p) gs >>= 1;
    const int ngroups = 32 / gs, grp = lane / gs, gl = lane % gs;
    const int qe = q + (q & 1), npairs = qe / 2;
    const double tol = fmax(1e-15, sqrt((double)p) * 2.3e-16);
    for (int b = blockIdx.y * kWarpsPerCta + warp; b < nb; b += gridDim.y * kWarpsPerCta) {
        const double* A = a + (int64_t)b * abs_ + sc[3];
        if (tall) {
            for (int e = lane; e < m * n; e += 32) G[(e % n) * ldp + (e / n)] = A[e];
        } else {
            for (int e = lane; e < m * n; e += 32) G[(e / n) * ldp + (e % n)] = A[e];
        }
        for (int e = lane; e < q * q; e += 32) V[(e / q) * ldq + (e % q)] = ((e / q) == (e % q)) ? 1.0 : 0.0;
        __syncwarp();
        for (int sweep = 0; sweep < 60 && q > 1; ++sweep) {
            bool rotated = false;
            for (int round = 0; round < qe - 1; ++round) {
                for (int base = 0; base < npairs; base += ngroups) {
                    const int pr = base + grp;
                    int i = 0, j = 0;
                    bool valid = pr < npairs;
                    if (valid) {
                        if (pr == 0) { i = qe - 1; j = round; }
                        else { i = (round + pr) % (qe - 1); j = (round - pr + (qe - 1)) % (qe - 1); }
                        valid = i < q && j < q;
                        if (i > j) { const int t = i; i = j; j = t; }
                    }
                    double* gi = G + i * ldp;
                    double* gj = G + j * ldp;
                    double aa = 0.0, bb = 0.0, cc = 0.0;
                    if (valid)
                        for (int r = gl; r < p; r += gs) { const double x = gi[r], y = gj[r]; aa += x * x; bb += y * y; cc += x * y; }
                    aa = group_sum(aa, gs); bb = group_sum(bb, gs); cc = group_sum(cc, gs);
                    if (valid && fabs(cc) > tol * sqrt(aa * bb) && aa * bb > 0.0) {
                        const double zeta = (bb - aa) / (2.0 * cc);
                        const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
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
        for (int c = 0; c < q; ++c) {
            double s2 = 0.0;
            for (int r = lane; r < p; r += 32) s2 += G[c * ldp + r] * G[c * ldp + r];
            s2 = warp_sum(s2);
            if (lane == 0) sig[c] = sqrt(s2);
        }
        __syncwarp();
        for (int c = lane; c < q; c += 32) {
            const double s = sig[c];
            int r = 0;
            for (int o = 0; o < q; ++o) { const double so = sig[o]; r += (so > s) || (so == s && o < c); }
            rnk[c] = r;
        }
        __syncwarp();
        double* S = sv + (int64_t)b * sbs + sc[6];
        double* O1 = out1 + (int64_t)b * o1bs + sc[4];
        double* O2 = out2 + (int64_t)b * o2bs + sc[5];
        for (int c = lane; c < q; c += 32) S[rnk[c]] = sig[c];
        if (tall) {
            for (int e = lane; e < m * k; e += 32) {
                const int c = e % k, r = e / k;   // consecutive lanes -> consecutive source columns of one row
                const double s = sig[c];
                O1[r * k + rnk[c]] = s > 0.0 ? G[c * ldp + r] / s : 0.0;
            }
            for (int e = lane; e < k * n; e += 32) { const int c = e / n, t = e - c * n; O2[rnk[c] * n + t] = V[c * ldq + t]; }
        } else {
            for (int e = lane; e < m * k; e += 32) { const int c = e % k, t = e / k; O1[t * k + rnk[c]] = V[c * ldq + t]; }
            for (int e = lane; e < k * n; e += 32) {
                const int c = e / n, r = e - c * n;
                const double s = sig[c];
                O2[rnk[c] * n + r] = s > 0.0 ? G[c * ldp + r] / s : 0.0;
            }
        }
        __syncwarp();
    }
}

// Greedy cross-sector truncation (svd.hpp:429-481) as a global ranking: value (i,t) is kept iff
// fewer than remain_cut values precede it in (value desc, sector asc, position asc) order and it
// exceeds relative_cut * max.  One CTA per chain.
__global__ void svd_cut_kernel(const int64_t* __restrict__ sect, int ns, int64_t s_total, const double* __restrict__ s, int64_t sbs,
                               int64_t remain_cut, double relative_cut, int32_t* __restrict__ counts) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    const double* S = s + (int64_t)b * sbs;
    int32_t* cnt = counts + (int64_t)b * ns;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) cnt[i] = 0;
    double mx = 0.0;
    for (int64_t e = threadIdx.x; e < s_total; e += blockDim.x) mx = fmax(mx, S[e]);
    mx = block_max(mx, red);
    const double thr = relative_cut * mx;
    for (int64_t e = threadIdx.x; e < s_total; e += blockDim.x) {
        const double v = S[e];
        if (!(v > thr)) continue;
        // sector of e
        int sec = 0;
        for (; sec < ns; ++sec) {
            const int64_t off = sect[sec * TNSP_SECT_COLS + 6], kk = sect[sec * TNSP_SECT_COLS + 2];
            if (e >= off && e < off + kk) break;
        }
        int64_t before = 0;
        for (int64_t o = 0; o < s_total; ++o) {
            const double w = S[o];
            before += (w > v) || (w == v && o < e);
        }
        if (before < remain_cut) atomicAdd(cnt + sec, 1);
    }
}

__global__ void svd_mask_kernel(const int64_t* __restrict__ sect, const int32_t* __restrict__ counts, int ns, double* __restrict__ out1,
                                int64_t o1bs, double* __restrict__ s, int64_t sbs, double* __restrict__ out2, int64_t o2bs, int nb) {
    const int64_t* sc = sect + (int64_t)blockIdx.x * TNSP_SECT_COLS;
    const int64_t m = sc[0], n = sc[1], k = sc[2];
    for (int b = blockIdx.y; b < nb; b += gridDim.y) {
        const int64_t keep = counts[(int64_t)b * ns + blockIdx.x];
        if (keep >= k) continue;
        double* O1 = out1 + (int64_t)b * o1bs + sc[4];
        double* O2 = out2 + (int64_t)b * o2bs + sc[5];
        double* S = s + (int64_t)b * sbs + sc[6];
        const int64_t w = k - keep;
        for (int64_t e = threadIdx.x; e < m * w; e += blockDim.x) O1[(e / w) * k + keep + (e % w)] = 0.0;
        for (int64_t e = threadIdx.x; e < w * n; e += blockDim.x) O2[keep * n + e] = 0.0;
        for (int64_t e = threadIdx.x; e < w; e += blockDim.x) S[keep + e] = 0.0;
    }
}

static int64_t svd_need(int64_t m, int64_t n) {
    const int64_t p = m >= n ? m : n, q = m >= n ? n : m;
    return q * p + q * q + 2 * q;
}

}  // namespace tnsp

using namespace tnsp;

// factor_sector.cu: discovered-sector kernels for one large dense(-embedded) matrix per chain
int tnsp_qr_sector_launch(const int64_t* sect, const int64_t* sh, const double* a, int64_t abs_, double* out1, int64_t o1bs,
                          double* out2, int64_t o2bs, int use_qr, int nb, cudaStream_t st);
int64_t tnsp_svd_sector_work(int64_t m, int64_t n);
int tnsp_svd_sector_launch(const int64_t* sect, const int64_t* sh, const double* a, int64_t abs_, double* out1, int64_t o1bs,
                           double* s, int64_t sbs, double* out2, int64_t o2bs, double* work, int64_t wbs, int nb, cudaStream_t st);
// factor_sector.cu: descriptor-driven sectors beyond the warp class (blocked Householder on the tensor pipe, QR-preconditioned Jacobi)
int tnsp_qr_desc_launch(const int64_t* sect, const int64_t* sh, int ns, const double* a, int64_t abs_, double* out1, int64_t o1bs,
                        double* out2, int64_t o2bs, int use_qr, int nb, cudaStream_t st);
int tnsp_svd_desc_launch(const int64_t* sect, const int64_t* sh, int ns, const double* a, int64_t abs_, double* out1, int64_t o1bs,
                         double* s, int64_t sbs, double* out2, int64_t o2bs, int nb, cudaStream_t st);

// the first-generation column-by-column kernels stay selectable for differential tests (tnsp_factor_desc_kernels(0))
static int g_desc_kernels = 1;
extern "C" int tnsp_factor_desc_kernels(int enable) {
    const int old = g_desc_kernels;
    if (enable >= 0) g_desc_kernels = enable;
    return old;
}

extern "C" int tnsp_qr_batched_f64(const int64_t* sect, int ns, const int64_t* sh, double* a, int64_t abs_, double* out1, int64_t o1bs,
                                   double* out2, int64_t o2bs, int use_qr, int nb, void* stream) {
    if (ns == 0 || nb == 0) return 0;
    int64_t smem = 0;
    for (int i = 0; i < ns; ++i) {
        const int64_t m = sh[i * TNSP_SECT_COLS], n = sh[i * TNSP_SECT_COLS + 1], k = sh[i * TNSP_SECT_COLS + 2];
        int64_t need = (m * n + k <= kSmemDoubles) ? (m * n + k) : k;
        if (k > kSmemDoubles) { set_error("tnsp_qr_batched_f64: k too large"); return 1; }
        if (need > smem) smem = need;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDoubles * 8);
        cudaFuncSetAttribute(qr_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarpsPerCta * kWarpDoubles * 8);
        attr_set = true;
    }
    bool small = true;
    for (int i = 0; i < ns; ++i) {
        const int64_t m = sh[i * TNSP_SECT_COLS], n = sh[i * TNSP_SECT_COLS + 1], k = sh[i * TNSP_SECT_COLS + 2];
        if (m * n == 0) continue;
        if (qr_warp_need(use_qr ? m : n, use_qr ? n : m, k) > kWarpDoubles) small = false;
    }
    if (small) {
        int64_t gy = (nb + kWarpsPerCta - 1) / kWarpsPerCta;
        if (gy > 65535) gy = 65535;
        qr_warp_kernel<<<dim3(ns, (unsigned)gy), 32 * kWarpsPerCta, kWarpsPerCta * kWarpDoubles * 8, (cudaStream_t)stream>>>(
            sect, a, abs_, out1, o1bs, out2, o2bs, use_qr, nb);
        return check_launch("tnsp_qr_batched_f64(warp)");
    }
    if (ns == 1) {
        const int rc = tnsp_qr_sector_launch(sect, sh, a, abs_, out1, o1bs, out2, o2bs, use_qr, nb, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    if (g_desc_kernels) {
        const int rc = tnsp_qr_desc_launch(sect, sh, ns, a, abs_, out1, o1bs, out2, o2bs, use_qr, nb, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    const int gy = nb > 65535 ? 65535 : nb;
    qr_kernel<<<dim3(ns, gy), kFactorThreads, smem * 8, (cudaStream_t)stream>>>(sect, a, abs_, out1, o1bs, out2, o2bs, use_qr, nb);
    return check_launch("tnsp_qr_batched_f64");
}

extern "C" int64_t tnsp_svd_work_size(const int64_t* sh, int ns) {
    int64_t total = 0;
    for (int i = 0; i < ns; ++i) total += svd_need(sh[i * TNSP_SECT_COLS], sh[i * TNSP_SECT_COLS + 1]);
    if (ns == 1) {
        const int64_t w = tnsp_svd_sector_work(sh[0], sh[1]);
        if (w > total) total = w;
    }
    return total;
}

extern "C" int tnsp_svd_batched_f64(const int64_t* sect, int ns, const int64_t* sh, const double* a, int64_t abs_, double* out1,
                                    int64_t o1bs, double* s, int64_t sbs, double* out2, int64_t o2bs, double* work, int64_t wbs,
                                    int nb, void* stream) {
    if (ns == 0 || nb == 0) return 0;
    int64_t smem = 0;
    // per-sector offsets into the scratch (device copy lives at the tail of `work`'s first chain? no:
    // offsets are small, pass them through a tiny device buffer allocated once per call site)
    static int64_t* d_off = nullptr;
    static int d_off_cap = 0;
    if (ns > d_off_cap) {
        if (d_off) cudaFree(d_off);
        d_off_cap = ns < 64 ? 64 : 2 * ns;
        if (cudaMalloc(&d_off, sizeof(int64_t) * d_off_cap) != cudaSuccess) { set_error("tnsp_svd_batched_f64: cudaMalloc"); return 1; }
    }
    int64_t offs[4096];
    if (ns > 4096) { set_error("tnsp_svd_batched_f64: too many sectors"); return 1; }
    int64_t acc = 0;
    bool need_global = false;
    for (int i = 0; i < ns; ++i) {
        offs[i] = acc;
        const int64_t need = svd_need(sh[i * TNSP_SECT_COLS], sh[i * TNSP_SECT_COLS + 1]);
        acc += need;
        if (sh[i * TNSP_SECT_COLS] * sh[i * TNSP_SECT_COLS + 1] == 0) continue;
        if (need <= kSmemDoubles) { if (need > smem) smem = need; } else need_global = true;
    }
    if (need_global && (work == nullptr || wbs < acc)) { set_error("tnsp_svd_batched_f64: scratch too small"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    if (need_global) cudaMemcpyAsync(d_off, offs, sizeof(int64_t) * ns, cudaMemcpyHostToDevice, st);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(svd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemDoubles * 8);
        cudaFuncSetAttribute(svd_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarpsPerCta * kWarpDoubles * 8);
        attr_set = true;
    }
    bool small = true;
    for (int i = 0; i < ns; ++i) {
        const int64_t m = sh[i * TNSP_SECT_COLS], n = sh[i * TNSP_SECT_COLS + 1];
        if (m * n == 0) continue;
        if (svd_warp_need(m >= n ? m : n, m >= n ? n : m) > kWarpDoubles) small = false;
    }
    if (small) {
        int64_t gyw = (nb + kWarpsPerCta - 1) / kWarpsPerCta;
        if (gyw > 65535) gyw = 65535;
        svd_warp_kernel<<<dim3(ns, (unsigned)gyw), 32 * kWarpsPerCta, kWarpsPerCta * kWarpDoubles * 8, st>>>(
            sect, a, abs_, out1, o1bs, s, sbs, out2, o2bs, nb);
        return check_launch("tnsp_svd_batched_f64(warp)");
    }
    if (ns == 1) {
        const int rc = tnsp_svd_sector_launch(sect, sh, a, abs_, out1, o1bs, s, sbs, out2, o2bs, work, wbs, nb, st);
        if (rc >= 0) return rc;
    }
    if (g_desc_kernels) {
        const int rc = tnsp_svd_desc_launch(sect, sh, ns, a, abs_, out1, o1bs, s, sbs, out2, o2bs, nb, st);
        if (rc >= 0) return rc;
    }
    const int gy = nb > 65535 ? 65535 : nb;
    svd_kernel<<<dim3(ns, gy), kFactorThreads, smem * 8, st>>>(sect, a, abs_, out1, o1bs, s, sbs, out2, o2bs, work, wbs, d_off, nb);
    return check_launch("tnsp_svd_batched_f64");
}

extern "C" int tnsp_svd_cut_f64(const int64_t* sect, int ns, int64_t s_total, const double* s, int64_t sbs, int64_t remain_cut,
                                double relative_cut, int32_t* counts, int nb, void* stream) {
    if (nb == 0 || ns == 0) return 0;
    svd_cut_kernel<<<nb, 128, 0, (cudaStream_t)stream>>>(sect, ns, s_total, s, sbs, remain_cut, relative_cut, counts);
    return check_launch("tnsp_svd_cut_f64");
}

extern "C" int tnsp_svd_mask_f64(const int64_t* sect, int ns, const int32_t* counts, double* out1, int64_t o1bs, double* s,
                                 int64_t sbs, double* out2, int64_t o2bs, int nb, void* stream) {
    if (nb == 0 || ns == 0) return 0;
    const int gy = nb > 65535 ? 65535 : nb;
    svd_mask_kernel<<<dim3(ns, gy), 128, 0, (cudaStream_t)stream>>>(sect, counts, ns, out1, o1bs, s, sbs, out2, o2bs, nb);
    return check_launch("tnsp_svd_mask_f64");
}

/*
 * tnsp_b200.h -- C-ABI of the B200 (sm_100a) kernels behind the TAT/tetragono sampling-VMC hot path.
 *
 * This is the drop-in boundary below the Python `TAT` module (reference: PyTAT/PyTAT.cpp:58-166,
 * PyTAT/PyTAT.hpp:569-1147).  The reference's own lower seam is the Fortran BLAS/LAPACK ABI that
 * TAT calls once per symmetry sector (contract.hpp:28-160, svd.hpp:30-97, qr.hpp:28-121); each
 * entry point below replaces one such per-sector loop by a single launch over
 * (sector x Monte-Carlo chain).  Plain pointers and sizes only; all pointers are DEVICE pointers
 * unless the name ends in `_host`.  `stream` is a cudaStream_t passed as void*.
 *
 * Conventions
 *   - all data are float64, row-major, dense inside a block;
 *   - a "batch" is a set of `nb` Monte-Carlo chains sharing one block structure; batch entry b of
 *     a tensor lives at base + b * bstride (bstride == 0 broadcasts one tensor to all chains);
 *   - descriptor tables are int64 arrays in device memory, produced by the host planner
 *     (tnsp_b200/TAT/plan.py) from the reference's integer rules;
 *   - every function returns 0 on success, non-zero on error (tnsp_last_error() has the text).
 */
#ifndef TNSP_B200_H
#define TNSP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNSP_PACK_MAX_RANK 8
#define TNSP_PACK_COLS (3 + 3 * TNSP_PACK_MAX_RANK)
#define TNSP_GEMM_COLS 8
#define TNSP_SECT_COLS 8

int tnsp_abi_version(void);
const char* tnsp_last_error(void);
/* number of kernels launched through this library since load (bench.py's gpu_launches) */
int64_t tnsp_launch_count(void);

/* ---- K1: sector pairing + transpose (replaces Tensor::edge_operator_implement's copy loop,
 * TAT/include/TAT/implement/edge_operator.hpp:651-688 + utility/multidimension_span.hpp:250-383).
 * desc[n][TNSP_PACK_COLS] = (src_off, dst_off, sign | rank<<1, dims[8], src_stride[8], dst_stride[8]);
 * estart[n+1] = prefix sum of elements per descriptor; total = estart[n]. */
int tnsp_pack_f64(const int64_t* desc, const int64_t* estart, int n_desc, int64_t total,
                  const double* src, int64_t src_bstride, double* dst, int64_t dst_bstride, int nb, void* stream);

/* single 2-D transposing descriptor, 32x32 shared-memory tiles (same operation, coalesced both ways):
 * dst[o*d_outer + j*d_j + i] = (+/-) src[o*s_outer + i*s_i + j] for o < n_outer, i < n_i, j < n_j */
int tnsp_pack_tiled_f64(int64_t src_off, int64_t dst_off, int64_t n_outer, int64_t n_i, int64_t n_j, int64_t s_outer,
                        int64_t s_i, int64_t d_outer, int64_t d_j, int neg, const double* src, int64_t src_bstride,
                        double* dst, int64_t dst_bstride, int nb, void* stream);

/* ---- K2: grouped GEMM over (sector x chain) (replaces detail::gemm_batch, contract.hpp:194-250,
 * fed by the descriptor loop contract.hpp:539-616 / :826-852).
 * desc[ng][8] = (m, n, k, a_off, b_off, c_off, flags, alpha); C[m x n] = alpha * op(A) * op(B);
 * flags bit0: A stored [k x m]; bit1: B stored [n x k]; otherwise [m x k] / [k x n]. */
int tnsp_gemm_grouped_f64(const int64_t* desc, int ng, const int64_t* desc_host,
                          const double* a, int64_t a_bstride, const double* b, int64_t b_bstride,
                          double* c, int64_t c_bstride, int nb, void* stream);

/* ---- K2g: contraction of two DENSE(-embedded) tensors read in place (replaces the two edge_operator merges
 * AND the gemm of contract.hpp:622-857 by one pass): C[b] (m x n, row-major) = alpha * A_g[b] * B_g[b] with
 *   A_g(r, kk) = a[b * a_bstride + tab[r] + tab[m + kk]],  B_g(kk, c) = b[b * b_bstride + tab[m + k + kk] + tab[m + 2k + c]]
 * tab: int32[m + 2k + n] element offsets built by the host planner from the edge strides (merge order of
 * edge_operator.hpp:321-404).  flags bit0: consecutive kk of A_g are (mostly) contiguous in memory, else consecutive r;
 * bit1: consecutive c of B_g are contiguous, else consecutive kk (decides which index runs along the lanes). */
int tnsp_gemm_gather_f64(const int32_t* tab, int64_t m, int64_t n, int64_t k, int flags, double alpha,
                         const double* a, int64_t a_bstride, const double* b, int64_t b_bstride,
                         double* c, int64_t c_bstride, int nb, void* stream);

/* The row-stream kernel behind tnsp_gemm_gather_f64 skips tensor-core instructions whose A or B fragment is all zero
 * (block-sparse operands of the charge-dense embedding: 13 % of the fragment pairs of cfg2's 1296 x 216 x 216 contraction are
 * non-zero; 1296 x 216 x 216 x 2368 chains: 10.9 -> 7.9 ms).  Results are identical (a product with an all-zero fragment adds
 * exact zeros).  Off by default -- on operands without zeros the tests cost 45 % -- and switched on by the dense embedding of
 * symmetric models; < 0 only queries.  Returns the previous setting. */
int tnsp_gemm_skip_zero_fragments(int enable);

/* ---- K3: batched QR / LQ with explicit Q (replaces ?geqrf/?orgqr and ?gelqf/?orglq per sector,
 * qr.hpp:178-304).  sect[ns][8] = (m, n, k, a_off, out1_off, out2_off, s_off, -); the m x n input
 * at a_off is destroyed.  use_qr != 0: out1 = Q (m x k), out2 = R (k x n); else out1 = L, out2 = Q. */
int tnsp_qr_batched_f64(const int64_t* sect, int ns, const int64_t* sect_host, double* a, int64_t a_bstride,
                        double* out1, int64_t o1_bstride, double* out2, int64_t o2_bstride,
                        int use_qr, int nb, void* stream);

/* ---- K4: batched one-sided Jacobi SVD (replaces ?gesvd('S','S') per sector, svd.hpp:104-211).
 * out1 = U (m x k), s = singular values sorted descending at s_off, out2 = V^T (k x n).
 * work: scratch of tnsp_svd_work_size() doubles per chain. */
int64_t tnsp_svd_work_size(const int64_t* sect_host, int ns);
int tnsp_svd_batched_f64(const int64_t* sect, int ns, const int64_t* sect_host, const double* a, int64_t a_bstride,
                         double* out1, int64_t o1_bstride, double* s, int64_t s_bstride,
                         double* out2, int64_t o2_bstride, double* work, int64_t w_bstride,
                         int nb, void* stream);

/* ---- K3s/K4s: the same factorisations for ONE dense(-embedded) matrix per chain (ns == 1) whose symmetry
 * sectors are NOT named by the descriptor but discovered on the device from the zero pattern (connected
 * components of the row/column graph): the lock-step batch engine stores symmetric tensors densely because
 * the chains of a batch differ in their sector structure.  Per discovered sector this is qr.hpp:178-304 /
 * svd.hpp:104-211; the bond index of a sector's factors is a contiguous range (QR) or the global descending
 * rank of the singular value (SVD), so the first `cut` indices are the greedy cut of svd.hpp:429-481.
 * Null vectors are zeros (a sector absent on one side does not appear, qr.hpp:419-429).  Outputs must be
 * zero-initialised; work: tnsp_svd_work_size() doubles per chain. */
int tnsp_qr_sectors_f64(const int64_t* sect, const int64_t* sect_host, double* a, int64_t a_bstride,
                        double* out1, int64_t o1_bstride, double* out2, int64_t o2_bstride, int use_qr, int nb, void* stream);
int tnsp_svd_sectors_f64(const int64_t* sect, const int64_t* sect_host, const double* a, int64_t a_bstride,
                         double* out1, int64_t o1_bstride, double* s, int64_t s_bstride, double* out2, int64_t o2_bstride,
                         double* work, int64_t w_bstride, int nb, void* stream);

/* The same two factorisations with the m x n operand READ IN PLACE from a dense tensor of any index order: element (i, j)
 * of the matrix is a[b * a_bstride + rc[i] + rc[m + j]] (int32 element offsets from the host planner).  Replaces the
 * merge / transpose copy (edge_operator.hpp:651-688) in front of qr.hpp:309-508 / svd.hpp:259-538. */
int tnsp_qr_sectors_gather_f64(const int64_t* sect, const int64_t* sect_host, const int32_t* rc, const double* a, int64_t a_bstride,
                               double* out1, int64_t o1_bstride, double* out2, int64_t o2_bstride, int use_qr, int nb, void* stream);
int tnsp_svd_sectors_gather_f64(const int64_t* sect, const int64_t* sect_host, const int32_t* rc, const double* a, int64_t a_bstride,
                                double* out1, int64_t o1_bstride, double* s, int64_t s_bstride, double* out2, int64_t o2_bstride,
                                double* work, int64_t w_bstride, int nb, void* stream);

/* Tuning knob of the two entry points above: matrices with >= min_elems elements are factorised sector by
 * sector from a device-side work queue (several CTAs per SM), smaller ones by one CTA per chain.  Returns the
 * previous threshold; a negative argument only queries. */
int64_t tnsp_sector_queue_min(int64_t min_elems);

/* Descriptor-driven sectors beyond the warp class go through blocked Householder on the FP64 tensor pipe and a
 * QR-preconditioned Jacobi (factor_sector.cu); enable = 0 selects the first-generation column-by-column kernels
 * (kept for differential tests), < 0 only queries.  Returns the previous setting. */
int tnsp_factor_desc_kernels(int enable);

/* One-sided Jacobi of the work-queue / descriptor kernels: 1 carries the squared column norms through a sweep (exact again
 * at the start of the next: one dot product per column pair), 0 (default: faster at the column lengths of cfg2, measured)
 * recomputes all three dot products per pair.  Both converge to the same factorisation.  Returns the previous setting. */
int tnsp_jacobi_cached_norms(int enable);

/* ---- greedy cross-sector truncation (svd.hpp:429-481): counts[b][i] = kept values of sector i. */
int tnsp_svd_cut_f64(const int64_t* sect, int ns, int64_t s_total, const double* s, int64_t s_bstride,
                     int64_t remain_cut, double relative_cut, int32_t* counts, int nb, void* stream);
/* zero singular triplets beyond counts[b][i] (keeps one block structure for all chains of a batch) */
int tnsp_svd_mask_f64(const int64_t* sect, int ns, const int32_t* counts, double* out1, int64_t o1_bstride,
                      double* s, int64_t s_bstride, double* out2, int64_t o2_bstride, int nb, void* stream);
/* S as diagonal blocks (svd.hpp:213-254): blk[nblk][4] = (s_off, dst_off, r, sign); dst pre-zeroed */
int tnsp_diag_scatter_f64(const int64_t* blk, int nblk, const double* s, int64_t s_bstride,
                          double* dst, int64_t dst_bstride, int nb, void* stream);

/* ---- elementwise / reductions (scalar.hpp:46-118, tensor.hpp:631-660, conjugate.hpp:99-116) ---- */
/* per-chain norms: kind -1 max|x|, 1 sum|x|, 2 sqrt(sum x^2); out[nb] */
int tnsp_norm_f64(const double* x, int64_t x_bstride, int64_t size, int kind, double* out, int nb, void* stream);
/* y[b] = alpha[b*alpha_stride] * x[b] (+ beta[b*beta_stride] * y[b] if beta != NULL); op: 0 mul, 1 div by alpha */
int tnsp_scale_f64(const double* x, int64_t x_bstride, const double* alpha, int64_t alpha_stride, int op,
                   double* y, int64_t y_bstride, int64_t size, int nb, void* stream);
/* z = a (op) b elementwise; op 0 +, 1 -, 2 *, 3 / ; bstride 0 broadcasts */
int tnsp_binary_f64(const double* a, int64_t a_bstride, const double* b, int64_t b_bstride, int op,
                    double* z, int64_t z_bstride, int64_t size, int nb, void* stream);
/* z = f(a): op 0 sqrt|a|, 1 reciprocal-except-zero, 2 negate, 3 abs */
int tnsp_unary_f64(const double* a, int64_t a_bstride, int op, double* z, int64_t z_bstride, int64_t size, int nb, void* stream);
/* y += w[b] * x  and  ey += w[b] * e[b] * x, reduced over the nb chains of the batch into ONE
 * accumulator (fused local-energy / log-derivative accumulation, observer.py:399-407) */
int tnsp_grad_accumulate_f64(const double* holes, int64_t h_bstride, const double* weight, const double* energy,
                             double* delta, double* edelta, int64_t size, int nb, void* stream);
/* per-block sign flip copy (conjugate of real fermionic tensors): blk[nblk][3] = (off, size, sign) */
int tnsp_block_sign_f64(const int64_t* blk, int nblk, const double* x, int64_t x_bstride,
                        double* y, int64_t y_bstride, int64_t size, int nb, void* stream);
/* gather rows: dst[b] = src[index[b]] (select the sampled physical slice per chain, lattice.py:319-339) */
int tnsp_gather_rows_f64(const double* src, int64_t row_size, const int32_t* index, double* dst, int64_t dst_bstride,
                         int nb, void* stream);
/* dst[b] = mask[b] ? a[b] : b_[b] (Metropolis accept, sampling.py:138-142) */
int tnsp_select_f64(const uint8_t* mask, const double* a, int64_t a_bstride, const double* b_, int64_t b_bstride,
                    double* dst, int64_t dst_bstride, int64_t size, int nb, void* stream);

/* ==== sector-compact lock-step tensors (tnsp_b200/TAT/ragged.py) ============================================================
 * A lock-step batch of Markov chains holds block-symmetric tensors whose sectors differ from chain to chain.  Every index of an
 * edge carries an int32 charge label per chain (|label| >= 2^29: the index is dead); a tensor is stored, per chain, as the
 * sector matrices of one grouping rows | cols of its edges, back to back (even offsets), row-major.  The integer planning the
 * reference does on the host per tensor (edge_operator.hpp:34-692, contract.hpp:306-620, qr.hpp:309-508, svd.hpp:259-538)
 * runs on the device from the labels; nothing below needs a device -> host copy.
 *   group table  int32 [nbT][TNSP_RT_HDR + 2 M]: nsec (-1: more than TNSP_RT_SMAX sectors), nvalid, skey[SMAX] ascending,
 *                sstart[SMAX + 1], perm[M] (merged indices sorted by (charge, index); dead ones last), inv[M]
 *   match table  int32 [nb][TNSP_RT_MSTRIDE]: stored elements, overflow flag, moff[SMAX + 1] (offset of row sector i),
 *                mcol[SMAX] (column sector paired with row sector i or -1)
 * A stride of 0 broadcasts one table / one tensor to all chains. */
#define TNSP_RT_SMAX 64
#define TNSP_RT_HDR (3 + 2 * TNSP_RT_SMAX)
#define TNSP_RT_MSTRIDE (4 + 2 * TNSP_RT_SMAX)
typedef struct {
    const double* data; int64_t data_stride;      /* sector matrices (or a dense array when rt == NULL) */
    const int32_t* rt; int64_t rt_stride; int64_t M;   /* group table of the rows, its chain stride, merged dimension */
    const int32_t* ct; int64_t ct_stride; int64_t N;
    const int32_t* match; int64_t match_stride;
} tnsp_rt_form;

/* sector pairing computed INSIDE a consumer kernel instead of a tnsp_rt_match_i32 launch of its own: same rule, the table is
 * written to match_out (chain stride match_out_stride) and the right-hand side to tsum_out (may be NULL) */
typedef struct {
    int rs, cs;
    const int32_t* t1; int t1_stride, s1;
    const int32_t* t2; int t2_stride, s2;
    int32_t* match_out; int64_t match_out_stride;
    int32_t* tsum_out;
    int64_t cap;   /* elements the destination holds per chain; a chain that needs more is stored EMPTY and counted (0: unchecked) */
} tnsp_rt_match_spec;

/* work counters for the bench's roofline (instrumented pass only): enable 0 / 1 (< 0: unchanged); out16 != NULL: copy the 16
 * uint64 counters to the host (15: chains dropped by a capacity check, cleared by reset == 2 only; 0 gemm algorithmic flops = sum 2mnk over the sectors, 1 gemm executed flops = DMMA issued x 512,
 * 2 gemm algorithmic bytes, 3 repacked elements, 4 qr bytes, 5 qr flops, 6 svd bytes, 7 sectors factorised, 8 gemm sectors) */
int tnsp_rt_stats(int enable, uint64_t* out16, int reset);
/* merged edge of a group of <= 8 edges (edge_operator.hpp:321-404, per chain): key(r) = sum_e signs[e] * labels[e][chain][r_e] */
int tnsp_rt_sort_i32(int n_edges, const int32_t* const* labels, const int64_t* lstrides, const int32_t* dims, const int32_t* signs,
                     int64_t M, int32_t* table, int nbT, void* stream);
/* sector pairing rs * rowkey + cs * colkey = s1 * t1[chain] + s2 * t2[chain] (core.hpp:162-190; NULL targets count as 0);
 * tsum (may be NULL) receives the right-hand side */
int tnsp_rt_match_i32(const int32_t* rt, int64_t rt_stride, int rs, const int32_t* ct, int64_t ct_stride, int cs, const int32_t* t1,
                      int t1_stride, int s1, const int32_t* t2, int t2_stride, int s2, int32_t* match, int32_t* tsum, int nbm, int64_t cap,
                      void* stream);
/* up to four pairings in ONE launch (the two / three pairing tables a factorisation needs for its factors: qr.hpp:419-429,
 * svd.hpp:405-427): job j uses entry j of every array; same rule and outputs as tnsp_rt_match_i32 (tsum entries may be NULL) */
typedef struct {
    const int32_t* rt; int64_t rt_stride; int32_t rs;
    const int32_t* ct; int64_t ct_stride; int32_t cs;
    const int32_t* t1; int32_t t1_stride; int32_t s1;
    const int32_t* t2; int32_t t2_stride; int32_t s2;
    int32_t* match; int32_t* tsum; int64_t cap;
} tnsp_rt_match_job;
int tnsp_rt_match_multi_i32(const tnsp_rt_match_job* jobs, int n_jobs, int nbm, void* stream);
/* regroup (edge_operator.hpp:651-688): plan = int32 [2 + 3 (nr + nc)]: nr, nc, then per destination edge (rows, then cols,
 * slowest first) dimension, 1 if the edge sits in the source's column group, stride inside that source group.  src->rt == NULL:
 * dense source; dst->rt == NULL: dense destination of `work` elements; else `work` bounds the stored elements (grid size). */
int tnsp_rt_repack_f64(const int32_t* plan, const tnsp_rt_form* src, const tnsp_rt_form* dst, const tnsp_rt_match_spec* dst_match,
                       double* dst_data, int64_t dst_stride, int64_t work, int nb, void* stream);
/* regrouping of a FERMIONIC tensor with the sign the reference attaches to every moved block (edge_operator.hpp:497-555, 591, 612),
 * evaluated per element: sign = s0 + lin . x + sum_{k<j} Q_kj x_k x_j (mod 2) over the parity bits x of the n_entries indexed edges
 * of the plan (dimension-1 edges included); quad[k] = bit mask of the j > k with Q_kj = 1; per_chain[chain] = lin bits | s0 << 31
 * (the host folds the dimension-1 edges with host-known charges into them); labels[k] / lstrides[k]: label array of entry k and its
 * chain stride; fermi_mask bit i: component i of a packed label (16 bits each) is fermionic */
int tnsp_rt_repack_signed_f64(const int32_t* plan, const tnsp_rt_form* src, const tnsp_rt_form* dst, const tnsp_rt_match_spec* dst_match,
                              double* dst_data, int64_t dst_stride, const int32_t* quad, const int32_t* per_chain, int64_t per_chain_stride,
                              int n_entries, const int32_t* const* labels, const int64_t* lstrides, int fermi_mask, int nb, void* stream);
/* two regroupings (the two operands of a contraction) in one launch; dst_match as above (NULL: dst->match is valid) */
int tnsp_rt_repack_pair_f64(const int32_t* plan0, const tnsp_rt_form* src0, const tnsp_rt_form* dst0, const tnsp_rt_match_spec* match0,
                            double* dst_data0, int64_t dst_stride0, int64_t work0, const int32_t* plan1, const tnsp_rt_form* src1,
                            const tnsp_rt_form* dst1, const tnsp_rt_match_spec* match1, double* dst_data1, int64_t dst_stride1, int64_t work1,
                            int nb, void* stream);
/* ONE grouped GEMM over (chain x sector) (contract.hpp:582-616): for every row sector i of c (rows of a, columns of b):
 * C_i = A_i B_i', i' = the row sector of b whose charge is ksign * (charge of the column sector a pairs with i); zeros when absent */
int tnsp_rt_gemm_f64(const tnsp_rt_form* a, const tnsp_rt_form* b, const tnsp_rt_form* c, const tnsp_rt_match_spec* c_match, double* c_data,
                     int64_t c_stride, int ksign, int nb, void* stream);
/* contraction over ALL edges of two tensors (closing contraction of a strip, amplitude x hole): out[chain][0] = sum over the stored
 * elements of dst of dst[e] * src[same multi-index]; plan as for tnsp_rt_repack_f64 with dst's grouping as the destination.  No
 * merged group over the whole tensor is built.  match / tsum: pairing table ([nb][TNSP_RT_MSTRIDE]) and summed target of the
 * one-element result (targets as in tnsp_rt_match_i32; the element exists iff the sum vanishes). */
int tnsp_rt_dot_f64(const int32_t* plan, const tnsp_rt_form* src, const tnsp_rt_form* dst, const int32_t* t1, int t1_stride, int s1,
                    const int32_t* t2, int t2_stride, int s2, double* out, int64_t out_stride, int32_t* match, int32_t* tsum, int nb,
                    void* stream);
/* per-sector QR (kind 0; qr.hpp:178-304, common edge qr.hpp:419-429) / SVD with the global greedy cut (kind 2; svd.hpp:104-211,
 * 429-481) of every (chain, sector) matrix of f.  The bond label of row sector i on the first factor is
 * t1s * t1[chain] - fsign_rs * rowkey(i).  Three calls: _plan (labels of the QR bond, work queue, work-buffer layout), then the
 * caller sorts / matches the bond, then _qr writes Q | R, or _svd_work + _svd_finish (labels of the kept bond) + sort / match +
 * _svd_scatter write U | S | V. */
int64_t tnsp_rt_factor_ws_ints(int64_t kfull);   /* int32 entries of `ws` per chain, kfull = min(M, N) */
int64_t tnsp_rt_svd_work_doubles(int64_t M, int64_t N);
int tnsp_rt_factor_plan(const tnsp_rt_form* f, int kind, int fsign_rs, const int32_t* t1, int t1_stride, int t1s, int64_t kdim,
                        int32_t* labels, int32_t* ws, int64_t ws_stride, int nb, void* stream);
int tnsp_rt_qr_f64(const tnsp_rt_form* f, int fsign_rs, const int32_t* t1, int t1_stride, int t1s, const int32_t* bond, int64_t bond_stride,
                   const int32_t* m_first, double* first, int64_t first_stride, const int32_t* m_second, double* second,
                   int64_t second_stride, int nb, void* stream);
int tnsp_rt_svd_work_f64(const tnsp_rt_form* f, double* work, int64_t work_stride, const int32_t* ws, int64_t ws_stride, int nb, void* stream);
int tnsp_rt_svd_finish_f64(const tnsp_rt_form* f, int fsign_rs, const int32_t* t1, int t1_stride, int t1s, int64_t kdim, int64_t remain_cut,
                           double relative_cut, const double* work, int64_t work_stride, int32_t* labels, int32_t* ws,
                           int64_t ws_stride, int nb, void* stream);
int tnsp_rt_svd_scatter_f64(const tnsp_rt_form* f, int fsign_rs, const int32_t* t1, int t1_stride, int t1s, const int32_t* bond,
                            int64_t bond_stride, const int32_t* m_first, double* first, int64_t first_stride, const int32_t* m_s, double* s,
                            int64_t s_stride, const int32_t* m_second, double* second, int64_t second_stride, const double* work,
                            int64_t work_stride, const int32_t* ws, int64_t ws_stride, int nb, void* stream);
/* elementwise over the stored sectors (scalar.hpp:46-118, tensor.hpp:631-660): op 0 multiply / 1 divide by alpha[chain];
 * binary op 0 + 1 - 2 * 3 /; norms kind -1 max 1 sum 2 euclid; scalar: the single element of an all-dimension-1 tensor (0 when
 * the sector is absent) */
int tnsp_rt_scale_f64(const double* x, int64_t x_stride, const int32_t* match, int64_t match_stride, const double* alpha, int alpha_stride,
                      int op, double* y, int64_t y_stride, int64_t cap, int nb, void* stream);
int tnsp_rt_binary_f64(const double* x, int64_t x_stride, const double* w, int64_t w_stride, const int32_t* match, int64_t match_stride,
                       int op, double* y, int64_t y_stride, int64_t cap, int nb, void* stream);
int tnsp_rt_norm_f64(const double* x, int64_t x_stride, const int32_t* match, int64_t match_stride, int kind, double* out, int nb, void* stream);
int tnsp_rt_scalar_f64(const double* x, int64_t x_stride, const int32_t* match, int64_t match_stride, double* out, int nb, void* stream);

/* ---- host RNG with libstdc++ semantics (TAT.random, PyTAT.hpp:87-126): one mt19937_64 per chain ---- */
void* tnsp_rng_create_host(int n_chains);
void tnsp_rng_destroy_host(void* rng);
void tnsp_rng_seed_host(void* rng, int chain, uint32_t seed);
/* out[i] = uniform_int_distribution<int>(lo[i], hi[i]) on chain i where active[i] != 0 */
void tnsp_rng_uniform_int_host(void* rng, const int32_t* lo, const int32_t* hi, const uint8_t* active, int32_t* out);
void tnsp_rng_uniform_real_host(void* rng, double lo, double hi, const uint8_t* active, double* out);
void tnsp_rng_normal_host(void* rng, int chain, double mean, double stddev, int64_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif

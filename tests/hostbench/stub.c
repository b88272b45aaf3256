/* generated for tests/hostbench: no-op stand-ins of the C-ABI (host-overhead benchmark only; NOT a product or test path) */
#include <stdint.h>
static long long n_calls = 0;
int tnsp_binary_f64() { ++n_calls; return 0; }
int tnsp_block_sign_f64() { ++n_calls; return 0; }
int tnsp_diag_scatter_f64() { ++n_calls; return 0; }
int tnsp_factor_desc_kernels() { ++n_calls; return 0; }
int tnsp_gather_rows_f64() { ++n_calls; return 0; }
int tnsp_gemm_gather_f64() { ++n_calls; return 0; }
int tnsp_gemm_grouped_f64() { ++n_calls; return 0; }
int tnsp_gemm_skip_zero_fragments() { ++n_calls; return 0; }
int tnsp_grad_accumulate_f64() { ++n_calls; return 0; }
int tnsp_jacobi_cached_norms() { ++n_calls; return 0; }
int tnsp_norm_f64() { ++n_calls; return 0; }
int tnsp_pack_f64() { ++n_calls; return 0; }
int tnsp_pack_tiled_f64() { ++n_calls; return 0; }
int tnsp_qr_batched_f64() { ++n_calls; return 0; }
int tnsp_qr_sectors_f64() { ++n_calls; return 0; }
int tnsp_qr_sectors_gather_f64() { ++n_calls; return 0; }
int tnsp_rt_binary_f64() { ++n_calls; return 0; }
int tnsp_rt_dot_f64() { ++n_calls; return 0; }
int tnsp_rt_factor_plan() { ++n_calls; return 0; }
int tnsp_rt_gemm_f64() { ++n_calls; return 0; }
int tnsp_rt_match_i32() { ++n_calls; return 0; }
int tnsp_rt_match_multi_i32() { ++n_calls; return 0; }
int tnsp_rt_norm_f64() { ++n_calls; return 0; }
int tnsp_rt_qr_f64() { ++n_calls; return 0; }
int tnsp_rt_repack_f64() { ++n_calls; return 0; }
int tnsp_rt_repack_pair_f64() { ++n_calls; return 0; }
int tnsp_rt_repack_signed_f64() { ++n_calls; return 0; }
int tnsp_rt_scalar_f64() { ++n_calls; return 0; }
int tnsp_rt_scale_f64() { ++n_calls; return 0; }
int tnsp_rt_sort_i32() { ++n_calls; return 0; }
int tnsp_rt_stats() { ++n_calls; return 0; }
int tnsp_rt_svd_finish_f64() { ++n_calls; return 0; }
int tnsp_rt_svd_scatter_f64() { ++n_calls; return 0; }
int tnsp_rt_svd_work_f64() { ++n_calls; return 0; }
int tnsp_scale_f64() { ++n_calls; return 0; }
int tnsp_select_f64() { ++n_calls; return 0; }
int tnsp_svd_batched_f64() { ++n_calls; return 0; }
int tnsp_svd_cut_f64() { ++n_calls; return 0; }
int tnsp_svd_mask_f64() { ++n_calls; return 0; }
int tnsp_svd_sectors_f64() { ++n_calls; return 0; }
int tnsp_svd_sectors_gather_f64() { ++n_calls; return 0; }
int tnsp_unary_f64() { ++n_calls; return 0; }
int64_t tnsp_rt_factor_ws_ints(int64_t kfull) { return 8 + 6 * 64 + kfull; }
int64_t tnsp_rt_svd_work_doubles(int64_t M, int64_t N) { int64_t k = M < N ? M : N; return k + M * k + k * N + 8; }
int64_t tnsp_svd_work_size(const int64_t* s, int ns) { return 1024; }
int64_t tnsp_sector_queue_min(int64_t m) { return 2048; }
long long tnsp_stub_calls(void) { return n_calls; }

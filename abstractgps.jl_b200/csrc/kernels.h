// kernels.h -- launch-level interface between the engine (engine.cu) and the sm_100a kernels.
// All matrices are column-major.  TILE = 128 is the factorisation tile edge: every matrix the
// Cholesky touches is padded to a multiple of TILE (identity padding), so kernels on that path
// see no ragged edges; the generic GEMM still bounds-checks for the prediction path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AGP_TILE 128

struct GramParams {
  int family;        // AGP_SE ...
  double variance;   // sigma_f^2
  double linear_c;
  int symmetric;     // 1: Xb == Xa, exact-zero self distance, optional noise on the diagonal
  int lower_only;    // 1: skip 64x64 tiles strictly above the diagonal
  int64_t valid_a;   // rows >= valid_a are padding
  int64_t valid_b;   // cols >= valid_b are padding
  int noise_kind;    // -1 none, 0 scalar, 1 vector
  double noise_s;
  const void* noise_v;  // device, T
  const unsigned char* mask_a;  // optional per-row validity (extended posteriors); overrides valid_a
  const unsigned char* mask_b;
  int64_t noise_off;  // noise_v index offset for the diagonal (block Gram of an extension)
  int64_t diag_off;   // column index offset: column gj of this launch is global column gj + diag_off
};

// op(A) is M x K, op(B) is K x N, C is M x N (ldc).  C = beta*C + alpha*op(A)op(B), alpha in {+1,-1}.
struct GemmArgs {
  const void* A; int64_t lda; int a_kmajor;  // 0: A(m,k) at A[m + k*lda]   1: A(m,k) at A[k + m*lda]
  const void* B; int64_t ldb; int b_kmajor;  // 0: B(k,n) at B[n + k*ldb]   1: B(k,n) at B[k + n*ldb]
  void* C; int64_t ldc;
  int64_t M, N, K;
  int alpha_neg;     // 1 -> alpha = -1 else +1
  int beta_one;      // 1 -> beta = 1 else 0
  int lower_only;    // 1 -> skip tiles with tile_n > tile_m (SYRK on a diagonal-anchored C)
  int trmm_lower;    // 1 -> A is lower triangular M x K anchored at (0,0): limit k <= row tile end
  // block-cyclic column gather (multi-GPU trailing update): local column n of C takes its B rows from
  // n_src = (n / 128) * b_tile_stride + n % 128 + b_off; 0 = identity.  The lower-only test uses n_src.
  int64_t b_tile_stride, b_off;
  int64_t b_tile_width;  // width of a distribution block in columns (0 -> 128)
};

template <typename T> void launch_prep_points(const T* X, int layout, int64_t n, int64_t n_pad, int D,
                                              int transform, double scale, const T* ard, T* Xt,
                                              cudaStream_t s);
template <typename T> void launch_gram(const T* Xa, const T* Xb, int64_t na_pad, int64_t nb_pad, int D,
                                       T* K, int64_t ldk, const GramParams& p, cudaStream_t s);
template <typename T> void launch_kdiag(const T* Xt, int64_t n, int D, int family, double variance,
                                        double linear_c, T* out, cudaStream_t s);
// border rows: E[s, j] = Y[j + s*ldy] - mean_j  (j < n, s < S), 0 elsewhere; E is TILE x n_pad at
// rows [n_pad, n_pad+TILE) of the factor matrix (leading dimension lda).
template <typename T> void launch_border_init(T* A, int64_t lda, int64_t n, int64_t n_pad, const T* Y,
                                              int64_t ldy, int S, int mean_kind, double mean_c,
                                              const T* mean_v, cudaStream_t s);
// same for a block of `ncols` columns of a column-distributed factor: A points at the block's first
// column, border rows start at row_off, column c of the block is global point col0 + c (valid if < n)
template <typename T> void launch_border_init_cols(T* A, int64_t lda, int64_t row_off, int64_t col0, int64_t ncols,
                                                   int64_t n, const T* Y, int64_t ldy, int S, int mean_kind,
                                                   double mean_c, const T* mean_v, cudaStream_t s);
// diagonal block factorisation + inverse: A (TILE x TILE at Ablk, lda) -> L in place (upper zeroed),
// Dinv = inv(L) (TILE x TILE col-major, lower), logdet_part[blk] = sum log L_jj, info (first bad pivot, 1-based).
template <typename T> void launch_potrf_diag(T* Ablk, int64_t lda, T* Dinv, double* logdet_part, int blk,
                                             int* info, cudaStream_t s);
template <typename T> void launch_gemm(const GemmArgs& g, cudaStream_t s);
// fp64 split schedule: factor-only diagonal block, 8-CTA strip inverse (off the critical path), and the
// panel TRSM by blocked substitution that does not need the 128x128 inverse
int potrf_split_enabled();
void launch_potrf_factor_f64(double* Ablk, int64_t lda, double* logdet_part, int blk, int* info, cudaStream_t s);
void launch_trtri_f64(const double* Ablk, int64_t lda, double* Dinv, cudaStream_t s);
void launch_trsm_sub_f64(double* A21, int64_t lda, int64_t M, const double* Lkk, cudaStream_t s);
// v extraction from the border rows + sqmahal: r[s*n_pad + j] = E[s, j], sq[s] = sum_j E[s,j]^2
template <typename T> void launch_extract_v(const T* A, int64_t lda, int64_t n_pad, int S, T* r, double* sq,
                                            cudaStream_t s);
// whole backward substitution in one persistent launch; flags_and_ticket: nblk+1 ints (zeroed inside)
template <typename T> void launch_bwd_solve(const T* A, int64_t lda, const T* Dinv, int nblk, T* r,
                                            int* flags_and_ticket, cudaStream_t s);
// distributed (column-cyclic) backward substitution pieces
template <typename T> void launch_bwd_diag(const T* Dinv_i, const T* r_i, T* alpha_i, cudaStream_t s);
template <typename T> void launch_bwd_update_local(const T* Lloc, int64_t lda, int i_blk, const T* alpha_i, T* r, int nloc,
                                                   int rank, int nranks, int G, cudaStream_t s);
template <typename T> void launch_bwd_update_local_multi(const T* Lloc, int64_t lda, int i_lo, int Gn, const T* alpha_lo, T* r,
                                                         int nloc, int rank, int nranks, int G, int64_t j_min, int64_t j_max,
                                                         cudaStream_t s);
template <typename T> void launch_bwd_block_solve(const T* Lblk, int64_t lda, const T* Dinv_blk, const T* r_blk, T* alpha_blk,
                                                  int Gn, cudaStream_t s);
template <typename T> void launch_finalize_logpdf(const double* logdet_part, int nblk, const double* sq, int S,
                                                  int64_t n, T* out, double* logdet_out, cudaStream_t s);
// mu[j] = mean_j + sum_i B[i + j*ldb] * alpha[i]
template <typename T> void launch_gemv_t(const T* B, int64_t ldb, int64_t n, int64_t m, const T* alpha,
                                         int mean_kind, double mean_c, const T* mean_v, T* mu, cudaStream_t s);
// var[j] = kdiag[j] - sum_i V[i + j*ldv]^2 (+ noise)
template <typename T> void launch_colsumsq_var(const T* V, int64_t ldv, int64_t n, int64_t m, const T* kdiag,
                                               int noise_kind, double noise_s, const T* noise_v, T* var,
                                               cudaStream_t s);
// out[i + j*ldo] = (i<=j) ? L[j + i*lda] : 0     (U = L')
template <typename T> void launch_export_upper(const T* A, int64_t lda, int64_t n, T* U, int64_t ldo,
                                               cudaStream_t s);
template <typename T> void launch_add_mean_cols(T* out, int64_t ldo, int64_t n, int S, int mean_kind,
                                                double mean_c, const T* mean_v, cudaStream_t s);
// C[i + j*ldc] = Kss[i + j*ldc] - C[...]  and symmetrise (mean_and_cov epilogue)
template <typename T> void launch_cov_finish(T* C, int64_t ldc, const T* Kss, int64_t m, cudaStream_t s);
template <typename T> void launch_fill(T* p, int64_t n, double v, cudaStream_t s);
// out[i] = y[i] - mean_i
template <typename T> void launch_sub_mean(const T* y, int64_t n, int mean_kind, double mean_c, const T* mean_v, T* out, cudaStream_t s);
// y[m] += sum_n A[m + n*lda] * x[n]   (rows coalesced)
template <typename T> void launch_gemv_n_acc(const T* A, int64_t lda, int64_t m, int64_t n, const T* x, T* y, cudaStream_t s);
// out[j] += sign * sum_i V[i + j*ldv]^2
template <typename T> void launch_colsumsq_acc(const T* V, int64_t ldv, int64_t n, int64_t m, double sign, T* out, cudaStream_t s);
// zero the strict upper triangle of every TILE x TILE diagonal block of an n_pad x n_pad factor
template <typename T> void launch_zero_diag_upper(T* A, int64_t lda, int64_t n_pad, cudaStream_t s);
// U(i,j) = L(map[j], map[i]) for i <= j (map == nullptr -> identity)
template <typename T> void launch_export_upper_map(const T* A, int64_t lda, int64_t n, const int64_t* map, T* U, int64_t ldo, cudaStream_t s);
template <typename T> void launch_copy2d(const T* src, int64_t lds, T* dst, int64_t ldd, int64_t rows,
                                         int64_t cols, cudaStream_t s);
// generic small helpers for VFE
template <typename T> void launch_scale_cols(T* B, int64_t ldb, int64_t rows, int64_t cols, const T* colscale,
                                             cudaStream_t s);  // B[:,j] *= colscale[j]
// EXPERIMENTAL (grad.cu): fused reduction of the logpdf gradient over the lower triangle; sums has 5 + D doubles (zeroed
// by the caller), noise_diag (n, optional) receives 1/2 W_ii
template <typename T> void launch_grad_reduce(const T* Xt, int D, int64_t n, int64_t n_pad, const T* Cinv, int64_t ldc,
                                              const T* alpha, int family, double linear_c, int want_ard, double* sums,
                                              T* noise_diag, cudaStream_t s);
template <typename T> void launch_add_diag(T* A, int64_t lda, int64_t n, double v, cudaStream_t s);
template <typename T> void launch_sumsq(const T* p, int64_t n, double* out, cudaStream_t s);  // out += sum p^2
template <typename T> void launch_vfe_prep(const T* y, int64_t n, int mean_kind, double mean_c, const T* mean_v,
                                           int noise_kind, double noise_s, const T* noise_v, const T* kdiag,
                                           T* delta, T* inv_sqrt_noise, double* scal /*[0]=logdet_sy [1]=sum d^2 [2]=tr*/,
                                           cudaStream_t s);

// true exactly once per (call site, current device): guards cudaFuncSetAttribute, which is a per-device setting
static inline bool agp_first_use_on_device(uint64_t* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

int64_t agp_kernel_launches();
void agp_count_launch();  // global counter (all kernels above bump it)

/* agp.h -- C ABI of libagp.so, the Blackwell-native (sm_100a) exact-GP engine that sits
 * behind AbstractGPs.jl's dense hot path.
 *
 * The reference (pure Julia, /root/reference) has no FFI; the two seams a drop-in uses are
 * Julia multiple dispatch on FiniteGP / PosteriorGP.  Each entry point
 * below names the reference method(s) it replaces (path:line relative to /root/reference).
 * The Julia shim that `ccall`s these symbols is julia/AGPBlackwell.jl; the identical symbols
 * are driven from Python ctypes in abstractgps.jl_b200/_cabi.py (the only host toolchain in
 * this image).  See INTEGRATION.md.
 *
 * Conventions
 *   - all functions return int32_t status (AGP_OK == 0); no exception crosses the ABI.
 *   - every data pointer is HOST memory owned by the caller unless the ctx was switched
 *     with agp_set_memspace(ctx, AGP_MEM_DEVICE) (then X / Y / Xs / Z-normals / outputs that
 *     are arrays are DEVICE pointers on the ctx's GPU; scalars-out stay host).
 *   - element type is `dtype` (AGP_F32 / AGP_F64) for every array argument.
 *   - points: AGP_POINT_MAJOR   = ColVecs(X),  X is D x N column-major (a point is contiguous)
 *             AGP_FEATURE_MAJOR = RowVecs(X),  X is N x D column-major
 *   - matrices out are column-major (Julia layout).
 *   - calls are blocking (stream-synchronised before return) and a ctx is not re-entrant.
 */
#ifndef AGP_H
#define AGP_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct agp_ctx agp_ctx;           /* device, streams, workspace arena, (NCCL comm)      */
typedef struct agp_post agp_post;         /* device-resident factor L (=U'), alpha, x, kernel     */
typedef struct agp_vfe_post agp_vfe_post; /* device-resident VFE cache (m_e, Lambda, U, alpha, z) */

enum { AGP_F32 = 0, AGP_F64 = 1 };
enum { AGP_SE = 0, AGP_MATERN12 = 1, AGP_MATERN32 = 2, AGP_MATERN52 = 3, AGP_LINEAR = 4 };
enum { AGP_T_NONE = 0, AGP_T_SCALE = 1, AGP_T_ARD = 2 };
enum { AGP_POINT_MAJOR = 0, AGP_FEATURE_MAJOR = 1 };
enum { AGP_MEM_HOST = 0, AGP_MEM_DEVICE = 1 };

enum {
  AGP_OK = 0,
  AGP_ERR_NOT_POSDEF = 1,   /* -> Julia PosDefException(info); info via agp_last_info()      */
  AGP_ERR_DIM_MISMATCH = 2, /* -> DimensionMismatch (src/sparse_approximations.jl:290-294)   */
  AGP_ERR_UNSUPPORTED = 3,
  AGP_ERR_CUDA = 4,
  AGP_ERR_NCCL = 5,
  AGP_ERR_INVALID = 6
};

/* sigma_f^2 * (kappa o transform): KernelFunctions ScaledKernel / TransformedKernel with
 * ScaleTransform(s) | ARDTransform(v); with_lengthscale(k,l) == scale 1/l.
 * Reference call sites: src/base_gp.jl:70,72,74. */
typedef struct {
  int32_t family;    /* AGP_SE ... AGP_LINEAR */
  int32_t transform; /* AGP_T_* */
  double variance;   /* sigma_f^2 */
  double scale;      /* ScaleTransform s */
  double linear_c;   /* LinearKernel c */
  const void* ard;   /* ARDTransform v: D values in `dtype`, HOST memory always */
} agp_kernel;

/* ZeroMean / ConstMean / CustomMean-evaluated-to-a-vector (src/mean_function.jl:27,40,52-55) */
typedef struct {
  int32_t kind; /* 0 zero, 1 const, 2 vector */
  double c;
  const void* v; /* `dtype`, length = number of points; HOST memory always */
} agp_mean;

/* Diagonal Sigma_y: Fill(sigma^2) or per-point vector (src/finite_gp_projection.jl:13-21) */
typedef struct {
  int32_t kind; /* 0 scalar, 1 per-point vector */
  double s;
  const void* v; /* `dtype`; HOST memory always */
} agp_noise;

typedef struct {
  int32_t tile_nb;       /* OUTER panel width (multiple of 128); 0 -> auto (512 from n_pad >= 8192, else 128) */
  int32_t fp64_mode;     /* -1 auto (tcgen05 from n_pad >= 8192), 0 = DMMA mma.sync trailing update,
                            1 = int8-sliced (Ozaki) trailing update on tcgen05.mma.kind::i8 */
  int32_t fp32_mode;     /* -1 auto (tcgen05 from n_pad >= 4096), 0 = FFMA tile kernels, 1 = int8-sliced trailing update /
                            triangular solves on tcgen05.mma.kind::i8 (4 seven-bit slices cover the fp32 significand) */
  int32_t lookahead;     /* 0 off, 1: overlap the next panel with the bulk of the trailing update, 2 (default): additionally
                            split the bulk so that the chain waits only for the panel-after-next block (DMMA path) */
  int32_t use_graph;     /* reserved */
  int32_t ozaki_slices;  /* 5..8 seven-bit slices of the tcgen05 fp64 path; 0 -> 7 (~2^-49 of the row scale) */
  int32_t profile_kernels; /* 1: CUDA events around every trailing-update launch (agp_last_timings[7]); default 0 */
  int32_t reserved[9];
} agp_config; /* NULL -> defaults; env AGP_NB, AGP_FP64_MODE, AGP_FP32_MODE, AGP_LOOKAHEAD, AGP_OZAKI_S, AGP_OZAKI_S32 override at agp_init */

/* ---- context ------------------------------------------------------------------------- */
int32_t agp_init(agp_ctx** ctx, int32_t device, const agp_config* cfg);
/* one process per GPU, NCCL across processes (rank 0 creates the id, host code ships it) */
int32_t agp_nccl_unique_id(void* out128);
int32_t agp_init_dist(agp_ctx** ctx, int32_t device, int32_t rank, int32_t nranks,
                      int32_t grid_p, int32_t grid_q, const void* nccl_unique_id128,
                      const agp_config* cfg);
int32_t agp_destroy(agp_ctx* ctx);
const char* agp_last_error(const agp_ctx* ctx);
int64_t agp_last_info(const agp_ctx* ctx); /* LAPACK-style failing pivot (1-based) after NOT_POSDEF */
int32_t agp_set_memspace(agp_ctx* ctx, int32_t memspace);
int32_t agp_set_config(agp_ctx* ctx, const agp_config* cfg); /* change the tunables of a live ctx (bench / tests) */
int32_t agp_get_config(const agp_ctx* ctx, agp_config* out);
const char* agp_version(void);

/* instrumentation for bench.py: per-phase device times (CUDA events on the launching stream)
 * of the LAST call: out[0]=total [1]=h2d [2]=gram [3]=cholesky [4]=solves [5]=d2h [6]=predict
 * [7]=trailing-update kernels only (sum), all in ms; returns number of doubles written. */
int32_t agp_last_timings(const agp_ctx* ctx, double* out, int32_t n);
int64_t agp_launch_count(const agp_ctx* ctx); /* kernels launched by this ctx so far */

/* ---- Gram only: cov(f, x) / cov(f, x, z) / cov(fx)  ---------------------------------------
 * replaces kernelmatrix(k,x[,z]) at src/base_gp.jl:70,74 and `C + Sigma_y`
 * src/finite_gp_projection.jl:96,135.  Z == NULL -> symmetric N x N (+ noise on the diagonal
 * when noise != NULL); else N x M cross-Gram.  K_out column-major, leading dim N. */
int32_t agp_gram(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, int32_t layout, const void* X,
                 int64_t N, int32_t D, const void* Z, int64_t M, const agp_noise* noise,
                 void* K_out);

/* ---- fused fit: ONE Gram + ONE Cholesky -> logpdf for S columns of Y, alpha, posterior ------
 * replaces logpdf(::FiniteGP, Y) src/finite_gp_projection.jl:306-311 (+ _sqmahal :325-326,
 * tr_Xt_invA_X / diag_Xt_invA_X src/util/common_covmat_ops.jl:90,101) AND
 * posterior(fx, y) src/exact_gpr_posterior.jl:29-35 (alpha = C \ (y - m) for column 0 of Y).
 * Y: N x S column-major.  logpdf_out: S values (`dtype`).  alpha_out: N values or NULL.
 * post_out: NULL or receives a handle owning the device-resident factor. */
int32_t agp_fit(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, const agp_mean* mean,
                const agp_noise* noise, int32_t layout, const void* X, int64_t N, int32_t D,
                const void* Y, int32_t S, void* logpdf_out, void* alpha_out, agp_post** post_out);

/* mean_and_var(::PosteriorGP, x*) src/exact_gpr_posterior.jl:85-90 (+ mean :60-62, var :68-70);
 * with noise != NULL also mean_and_var(::FiniteGP) src/finite_gp_projection.jl:154-158.
 * mean_s == NULL -> the prior mean given at fit time (zero/const only). */
int32_t agp_post_mean_var(agp_post* p, int32_t layout, const void* Xs, int64_t M,
                          const agp_mean* mean_s, const agp_noise* noise_s, void* mean_out,
                          void* var_out);
/* mean_and_cov(::PosteriorGP, x*) src/exact_gpr_posterior.jl:78-83 (Xt_invA_X
 * src/util/common_covmat_ops.jl:54-58).  cov_out M x M column-major. */
int32_t agp_post_mean_cov(agp_post* p, int32_t layout, const void* Xs, int64_t M,
                          const agp_mean* mean_s, void* mean_out, void* cov_out);
/* logpdf(f_post(x*, Sigma*), Y): src/finite_gp_projection.jl:306-311 (matrix Y :313-318) for a
 * FiniteGP over a PosteriorGP -- mean_and_cov(f_post, x*) src/exact_gpr_posterior.jl:78-83 is formed on
 * the device, Sigma* (noise_s; NULL -> 1e-18) added, factored in place; the M x M covariance never visits
 * the host.  Y is M x S column-major, S <= 128; logpdf_out[S].  AGP_ERR_NOT_POSDEF as agp_fit. */
int32_t agp_post_logpdf(agp_post* p, int32_t layout, const void* Xs, int64_t M,
                        const agp_mean* mean_s, const agp_noise* noise_s, const void* Y, int32_t S,
                        void* logpdf_out);
/* rand(f_post(x*, Sigma*), S): src/finite_gp_projection.jl:233-240, out = m* + chol(C* + Sigma*).U' Z with
 * caller-supplied standard normals Z (M x S column-major) as agp_rand. */
int32_t agp_post_rand(agp_post* p, int32_t layout, const void* Xs, int64_t M, const agp_mean* mean_s,
                      const agp_noise* noise_s, const void* Z, int32_t S, void* out);
/* EXPERIMENTAL -- compiled, not yet validated on a device.  Gradient of logpdf(fx, y) with
 * respect to the hyper-parameters, from the factor and alpha that agp_fit left in the handle: what reverse-mode AD
 * returns through the reference's logpdf (test/finite_gp_projection.jl:152-178, examples/1-mauna-loa/script.jl:200-242).
 * grad_out (double, 5 + D entries): [0] d/d variance, [1] d/d ScaleTransform s, [2] d/d LinearKernel c,
 * [3] d/d sigma^2 (scalar noise; = 1/2 tr W), [4] d/d ConstMean c, [5 + d] d/d ARDTransform v_d.
 * noise_diag_out (N elements of the handle's dtype, or NULL): d/d sigma_i^2 for per-point noise, and -- via
 * d/d m_i = alpha_i -- the caller already holds the gradient w.r.t. a vector mean. */
int32_t agp_post_logpdf_grad(agp_post* p, double* grad_out, void* noise_diag_out);
/* V = U' \ B (N x nrhs, column-major): backs Xt_invA_X / diag_Xt_invA_X / Xt_invA_Y /
 * tr_Xt_invA_X on a device factor, src/util/common_covmat_ops.jl:54-60,90,101. */
int32_t agp_post_solve_lower(agp_post* p, const void* B, int64_t nrhs, void* V_out);
/* C.U for p.data.C.U compatibility (test/exact_gpr_posterior.jl:40): N x N column-major upper */
int32_t agp_post_factor_export(agp_post* p, void* U_out);
int32_t agp_post_logdet(agp_post* p, double* logdet_out); /* logdet(C) = 2 sum log U_ii */
int64_t agp_post_n(const agp_post* p);
/* sequential conditioning: posterior(fx::FiniteGP{<:PosteriorGP}, y) src/exact_gpr_posterior.jl:46-56
 * via update_chol src/util/common_covmat_ops.jl:38-42.  alpha_out receives the N1+N2 re-solved weights
 * (or NULL).  post_out != NULL: a NEW handle is returned and p stays valid (the reference's value
 * semantics -- both posteriors usable, both to be freed); post_out == NULL: p is extended in place. */
int32_t agp_post_extend(agp_post* p, int32_t layout, const void* X2, int64_t N2, const void* y2,
                        const agp_mean* mean2, const agp_noise* noise2, void* alpha_out,
                        agp_post** post_out);
int32_t agp_post_free(agp_post* p);

/* rand(rng, fx, S) / _rand! src/finite_gp_projection.jl:233-237,271-277: out = m + U' Z with the
 * caller's standard normals Z (N x S column-major) -- the RNG stream stays host-defined. */
int32_t agp_rand(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, const agp_mean* mean,
                 const agp_noise* noise, int32_t layout, const void* X, int64_t N, int32_t D,
                 const void* Z, int32_t S, void* out);

/* ---- VFE (Titsias) -------------------------------------------------------------------------
 * approx_log_evidence(::VFE)/elbo src/sparse_approximations.jl:248-254 with
 * _compute_intermediates :289-305 and tr_Cf_invSigma_y :307-313; dtc_out (optional) is the DTC
 * objective :282-286.  Only diagonal Sigma_y (as in the reference, :307-313). */
int32_t agp_vfe_elbo(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, const agp_mean* mean,
                     const agp_noise* noise, int32_t layout, const void* X, int64_t N, int32_t D,
                     const void* Zind, int64_t M, const agp_noise* jitter, const void* y,
                     void* elbo_out, void* dtc_out);
/* posterior(::VFE, fx, y) src/sparse_approximations.jl:58-75 */
int32_t agp_vfe_fit(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, const agp_mean* mean,
                    const agp_noise* noise, int32_t layout, const void* X, int64_t N, int32_t D,
                    const void* Zind, int64_t M, const agp_noise* jitter, const void* y,
                    agp_vfe_post** out);
/* mean_and_var(::ApproxPosteriorGP, x*) src/sparse_approximations.jl:212-217 */
int32_t agp_vfe_mean_var(agp_vfe_post* p, int32_t layout, const void* Xs, int64_t Ms,
                         void* mean_out, void* var_out);
/* EXPERIMENTAL -- composed of validated kernels, not yet run on a device.
 * mean_and_cov(::ApproxPosteriorGP, x*) src/sparse_approximations.jl:205-210 (cov :187-190); cov_out M x M column-major,
 * no observation noise.  The prior mean at x* is the handle's Zero/Const mean (a closure mean is added by the host). */
int32_t agp_vfe_mean_cov(agp_vfe_post* p, int32_t layout, const void* Xs, int64_t M, void* mean_out,
                         void* cov_out);
/* logpdf(f_approx_post(x*, Sigma*), Y) and rand(f_approx_post(x*, Sigma*), S): src/finite_gp_projection.jl:306-318,
 * :233-240 over the approximate posterior -- C* + Sigma* is formed and factored on the device like agp_post_logpdf. */
int32_t agp_vfe_post_logpdf(agp_vfe_post* p, int32_t layout, const void* Xs, int64_t M, const agp_noise* noise_s,
                            const void* Y, int32_t S, void* logpdf_out);
int32_t agp_vfe_post_rand(agp_vfe_post* p, int32_t layout, const void* Xs, int64_t M, const agp_noise* noise_s,
                          const void* Z, int32_t S, void* out);
int32_t agp_vfe_post_free(agp_vfe_post* p);

/* Device-pointer entry points (AGP_MEM_DEVICE and the agp_debug_* hooks): the library runs on its OWN non-blocking streams
 * and returns after synchronising them, so results are complete on return -- but the CALLER must make sure the device
 * buffers it passes in are complete (synchronise the stream that produced them) before the call. */
/* ---- test hook for the tcgen05 int8-sliced fp64 trailing update (csrc/umma_ozaki.cu): DEVICE pointers;
 * C (M x N, ldc, fp64) -= P P' (lower tiles when lower_only), P = M x K fp64 (lda), S in 5..8 slices. */
int32_t agp_debug_ozaki_syrk(agp_ctx* ctx, void* C_dev, int64_t ldc, const void* P_dev, int64_t lda, int64_t M,
                             int64_t N, int32_t K, int32_t S, int32_t lower_only);

/* general product on the same tcgen05 path: C (M x N, N % 128 == 0) += sign * A B' with fp32 or fp64 operands and output,
 * row-contiguous (element (r,k) at [r + k*ld]) or k-major ([k + r*ld]) operands; B_dev == NULL: B = A, lower tiles only.
 * S = number of 7-bit slices (3..5 for fp32 output, 4..8 for fp64). */
int32_t agp_debug_ozaki_gemm(agp_ctx* ctx, void* C_dev, int32_t c_is_float, int64_t ldc, const void* A_dev, int32_t a_is_float,
                             int32_t a_kmajor, int64_t lda, int64_t M, const void* B_dev, int32_t b_is_float, int32_t b_kmajor,
                             int64_t ldb, int64_t N, int32_t K, int32_t S, double sign);
/* same kernel through the block-cyclic column map of the multi-GPU trailing update (the strip-table tile enumeration):
 * P has m_panel rows; row r of C (M x N local columns, N % 128 == 0) pairs with panel row r + a_off, local column n with
 * panel row (n / b_tile_width) * b_tile_stride + n % b_tile_width + b_off; lower tiles only (relative to panel rows). */
int32_t agp_debug_ozaki_syrk_map(agp_ctx* ctx, void* C_dev, int64_t ldc, const void* P_dev, int64_t lda, int64_t m_panel,
                                 int64_t M, int64_t N, int32_t K, int32_t S, int64_t b_tile_stride, int64_t b_tile_width,
                                 int64_t b_off, int64_t a_off);

/* ---- host-only helpers of the 2D block-cyclic tile map (no GPU needed; used by the CPU
 * world_size-2 tests): owner rank of tile (i,j) on a P x Q grid and local tile counts. */
int32_t agp_bc_owner(int32_t ti, int32_t tj, int32_t grid_p, int32_t grid_q);
int64_t agp_bc_local_tiles(int32_t ntiles, int32_t rank, int32_t grid_p, int32_t grid_q);

#ifdef __cplusplus
}
#endif
#endif /* AGP_H */

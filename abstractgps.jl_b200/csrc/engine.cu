// engine.cu -- host side of libagp.so: context, stream-ordered workspace, the fused "fit"
// (Gram -> blocked right-looking Cholesky with bordered forward solve -> backward solve -> logpdf),
// prediction, sampling, and the extern "C" ABI declared in include/agp.h.
//
// Data layout in HBM: the factor lives in ONE column-major buffer of
// (n_pad + 128) x n_pad elements, n_pad = N rounded up to 128 with identity padding.  The Gram kernel
// writes the lower triangle directly into it, the Cholesky runs in place (A = L L', L = U'), and the
// extra 128-row "border" tile holds delta' = (Y - m)' so that the panel TRSM + trailing update
// perform the forward substitution v = L^-1 delta for free (up to 128 right-hand sides).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <nccl.h>

#include "agp.h"
#include "kernels.h"
#include "umma_ozaki.h"

#define TILE AGP_TILE

struct agp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // look-ahead side stream (bulk of the trailing update)
  cudaStream_t stream3 = nullptr;  // strip inverses of the diagonal blocks (needed only by the solves)
  cudaEvent_t ev_s3 = nullptr, ev_fac = nullptr;
  bool s3_dirty = false;
  std::vector<cudaEvent_t> dep_ev;  // dependency events of the look-ahead schedule
  agp_config cfg{};
  std::string err;
  int64_t info = 0;
  int memspace = AGP_MEM_HOST;
  bool out_dev_override = false;  // internal: outputs of the current call are device pointers whatever memspace says
  double timings[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t ev[8]{};
  std::vector<cudaEvent_t> prof_ev;  // pairs around every trailing-update launch
  int prof_used = 0;
  int profile = 1;
  int rank = 0, nranks = 1, grid_p = 1, grid_q = 1;
  ncclComm_t nccl = nullptr;
  OzakiWs oz{};            // slice workspace of the tcgen05 fp64 path (cached across fits of the same shape)
  OzakiWs oz2{};           // second slice buffer of the pipelined distributed schedule (panel k+1 is sliced while rest(k) runs)
  int64_t oz_rows = 0, oz2_rows = 0;
  cudaStream_t stream_comm = nullptr;  // panel broadcasts of the pipelined distributed schedule
  int oz_S = 7;
  int oz_chunk = 16;        // bounded-CTA size (tiles) of the rest updates that run beside a higher-priority stream; 0 = persistent
  int oz_S32 = 4;          // slices of the fp32 operands (4 x 7 bits >= the 24-bit significand)
};

struct agp_post {
  agp_ctx* ctx = nullptr;
  int dtype = AGP_F64;
  int64_t n = 0, n_pad = 0, lda = 0;
  int D = 0;
  void* L = nullptr;      // (n_pad+TILE) x n_pad
  void* Dinv = nullptr;   // nblk x TILE x TILE
  void* Xt = nullptr;     // n_pad x D transformed points
  void* alpha = nullptr;  // n_pad
  void* ard = nullptr;    // D (device) or null
  void* delta = nullptr;  // n_pad: y - m at the valid rows, 0 at padding
  unsigned char* valid = nullptr;  // n_pad row-validity mask; null while the valid rows are [0, n)
  std::vector<std::pair<int64_t, int64_t>> segs;  // (offset, length) of the valid row segments
  agp_kernel k{};
  int mean_kind = 0;
  double mean_c = 0.0;
  double logdet = 0.0;
  // distributed posterior (multi-GPU fit): the factor stays in its block-column-cyclic form (Lloc: lda x nloc*W, block
  // column jo = lj*R + me) until the first operation that needs it whole; post_replicate() then gathers it over NVLink
  // (one ncclBroadcast per block column from its owner) into L / Dinv and every rank holds a regular handle.
  int dist_R = 0, dist_me = 0, dist_G = 0, dist_nloc = 0;
  int64_t dist_W = 0;
  void* Lloc = nullptr;
  void* Dinv_loc = nullptr;
};

struct agp_vfe_post {
  agp_ctx* ctx = nullptr;
  int dtype = AGP_F32;
  int64_t m = 0, m_pad = 0, lda = 0;
  int D = 0;
  void* U = nullptr;      // factor of Kzz + jitter   (m_pad+TILE) x m_pad
  void* Udinv = nullptr;
  void* Lam = nullptr;    // factor of A A' + I
  void* Ldinv = nullptr;
  void* Zt = nullptr;     // m_pad x D
  void* m_e = nullptr;    // m_pad
  void* ard = nullptr;
  agp_kernel k{};
  int mean_kind = 0;
  double mean_c = 0.0;
};

namespace {

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t env_int64(const char* name, int64_t dflt) {
  const char* v = getenv(name);
  return v ? atoll(v) : dflt;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      char _b[512];                                                                       \
      snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      ctx->err = _b;                                                                      \
      return AGP_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define CKN(call)                                                                         \
  do {                                                                                    \
    ncclResult_t _r = (call);                                                             \
    if (_r != ncclSuccess) {                                                              \
      char _b[512];                                                                       \
      snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(_r)); \
      ctx->err = _b;                                                                      \
      return AGP_ERR_NCCL;                                                                \
    }                                                                                     \
  } while (0)

template <typename T> struct NcclType;
template <> struct NcclType<float> { static constexpr ncclDataType_t v = ncclFloat; };
template <> struct NcclType<double> { static constexpr ncclDataType_t v = ncclDouble; };

struct Scratch {  // stream-ordered allocations freed together
  agp_ctx* ctx;
  std::vector<void*> ptrs;
  explicit Scratch(agp_ctx* c) : ctx(c) {}
  ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, ctx->stream); }
  cudaError_t alloc(void** p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
  void release(void* p) {  // ownership moves out
    for (auto& q : ptrs) if (q == p) { q = ptrs.back(); ptrs.pop_back(); return; }
  }
};

template <typename T>
int upload(agp_ctx* ctx, Scratch& sc, const void* src, size_t count, bool always_host, T** out) {
  if (!src || count == 0) { *out = nullptr; return AGP_OK; }
  if (!always_host && ctx->memspace == AGP_MEM_DEVICE) { *out = (T*)src; return AGP_OK; }
  void* d = nullptr;
  CK(sc.alloc(&d, count * sizeof(T)));
  CK(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *out = (T*)d;
  return AGP_OK;
}

template <typename T>
int download(agp_ctx* ctx, void* dst, const T* src, size_t count, bool always_host) {
  if (!dst || count == 0) return AGP_OK;
  cudaMemcpyKind kind = (!always_host && (ctx->memspace == AGP_MEM_DEVICE || ctx->out_dev_override)) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CK(cudaMemcpyAsync(dst, src, count * sizeof(T), kind, ctx->stream));
  return AGP_OK;
}

int check_kernel(agp_ctx* ctx, const agp_kernel* k, int D) {
  if (!k) { ctx->err = "kernel spec is NULL"; return AGP_ERR_INVALID; }
  if (k->family < AGP_SE || k->family > AGP_LINEAR) { ctx->err = "unsupported kernel family"; return AGP_ERR_UNSUPPORTED; }
  if (k->transform < AGP_T_NONE || k->transform > AGP_T_ARD) { ctx->err = "unsupported transform"; return AGP_ERR_UNSUPPORTED; }
  if (k->transform == AGP_T_ARD && !k->ard) { ctx->err = "ARD transform without weights"; return AGP_ERR_INVALID; }
  if (D <= 0) { ctx->err = "D must be positive"; return AGP_ERR_DIM_MISMATCH; }
  return AGP_OK;
}

void prof_begin(agp_ctx* ctx) { ctx->prof_used = 0; }
cudaEvent_t prof_event(agp_ctx* ctx) {
  if (ctx->prof_used == (int)ctx->prof_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->prof_ev.push_back(e);
  }
  return ctx->prof_ev[ctx->prof_used++];
}
double prof_total_ms(agp_ctx* ctx) {
  double t = 0;
  for (int i = 0; i + 1 < ctx->prof_used; i += 2) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->prof_ev[i], ctx->prof_ev[i + 1]) == cudaSuccess) t += ms;
  }
  return t;
}

// ---- blocked right-looking Cholesky, in place, lower; rows include the border tile -------------
// Look-ahead (depth 1): after the panel solve of step k, the main stream updates only the NEXT panel
// column and immediately factors / solves panel k+1, while the side stream applies panel k to the rest
// of the trailing matrix.  Event edges: rest_k waits trsm_k; next-column update of step k+1 waits
// rest_k (both touch column block k+2).
static cudaEvent_t dep_event(agp_ctx* ctx, size_t i) {
  while (ctx->dep_ev.size() <= i) {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    ctx->dep_ev.push_back(e);
  }
  return ctx->dep_ev[i];
}

// Two-level blocking: an OUTER panel is G inner blocks (G*128 columns).  Inside it, each inner step is
// potrf -> panel TRSM -> rank-128 update of the remaining inner columns only; the trailing matrix then
// receives ONE rank-(G*128) update per outer panel (K = G*128 halves/quarters the read-modify-write
// passes over C and is what the tensor-core trailing kernels need to be compute-bound).
template <typename T>
void trailing_update(agp_ctx* ctx, T* L, int64_t lda, int64_t row0, int64_t col0, int64_t kcol0, int64_t K,
                     int64_t M, int64_t N, cudaStream_t st, const OzakiWs* oz = nullptr, int64_t oz_row0 = 0) {
  // C = L[row0.., col0..] (M x N, diagonal-anchored iff row0 == col0) -= L[row0.., kcol0..] * L[col0.., kcol0..]'
  if (M <= 0 || N <= 0) return;
  if (ctx->profile) cudaEventRecord(prof_event(ctx), st);
  bool done = false;
  if (oz) {  // tcgen05 int8-sliced path: the slices of panel rows [oz_row0, ...) are already in *oz
    if constexpr (std::is_same<T, double>::value) {
      ozaki_syrk(*oz, L + row0 + col0 * lda, lda, M, N, 1, 0, 0, col0 - oz_row0, row0 - oz_row0, st);
      done = true;
    } else {
      done = ozaki_update_ex(*oz, L + row0 + col0 * lda, 1, lda, M, N, 0, -1.0, 0, 0, col0 - oz_row0, row0 - oz_row0, st) == 0;
    }
  }
  if (!done) {
    GemmArgs u{};
    u.A = L + row0 + kcol0 * lda; u.lda = lda;
    u.B = L + col0 + kcol0 * lda; u.ldb = lda;
    u.C = L + row0 + col0 * lda; u.ldc = lda;
    u.M = M; u.N = N; u.K = K; u.alpha_neg = 1; u.beta_one = 1; u.lower_only = 1;
    if (row0 != col0) { u.b_tile_stride = TILE; u.b_off = col0 - row0; }  // lower-only test relative to row0
    launch_gemm<T>(u, st);
  }
  if (ctx->profile) cudaEventRecord(prof_event(ctx), st);
}

// the strip inverses run on stream3; every consumer of Dinv on the main stream joins them first
static void join_inverses(agp_ctx* ctx) {
  if (!ctx->s3_dirty) return;
  cudaEventRecord(ctx->ev_s3, ctx->stream3);
  cudaStreamWaitEvent(ctx->stream, ctx->ev_s3, 0);
  ctx->s3_dirty = false;
}

// resolve the outer panel width (in 128-blocks) and the fp64 trailing-update engine for a problem size:
// explicit config / env wins; "auto" = tcgen05 int8-sliced path with 512-wide panels from n_pad >= 8192
static int resolve_G(const agp_ctx* ctx, int64_t n_pad) {
  int nb = ctx->cfg.tile_nb;
  if (nb <= 0) nb = (n_pad >= 8192) ? 512 : TILE;
  int G = nb / TILE;
  if (G > 8) G = 8;  // kernels that stage a whole outer block (distributed backward solve) hold at most 8 inner blocks
  return G < 1 ? 1 : G;
}
static int resolve_fp64_mode(const agp_ctx* ctx, int64_t n_pad) {
  if (ctx->cfg.fp64_mode >= 0) return ctx->cfg.fp64_mode;
  return n_pad >= 8192 ? 1 : 0;
}
// fp32: 0 = FFMA tile kernels, 1 = int8-sliced tcgen05 path (4 slices); auto = tcgen05 from n_pad >= 4096
static int resolve_fp32_mode(const agp_ctx* ctx, int64_t n_pad) {
  if (ctx->cfg.fp32_mode >= 0) return ctx->cfg.fp32_mode;
  return n_pad >= 4096 ? 1 : 0;
}
template <typename T> static int resolve_tensor_mode(const agp_ctx* ctx, int64_t n_pad) {
  return std::is_same<T, double>::value ? resolve_fp64_mode(ctx, n_pad) : resolve_fp32_mode(ctx, n_pad);
}
template <typename T> static int slices_of(const agp_ctx* ctx) { return std::is_same<T, double>::value ? ctx->oz_S : ctx->oz_S32; }
// (re)size the cached slice workspace: rows x K bytes per slice, S slices
static bool ensure_oz2(agp_ctx* ctx, int64_t rows, int K, int S, cudaStream_t s) {  // second cached workspace (long-K products)
  if (!ctx->oz2.SL || ctx->oz2.K != K || ctx->oz2_rows < rows || ctx->oz2.S != S) {
    if (ctx->oz2.SL) ozaki_ws_destroy(&ctx->oz2, s);
    if (ozaki_ws_create(&ctx->oz2, rows, K, S, s) == 0) ctx->oz2_rows = rows;
    else { memset(&ctx->oz2, 0, sizeof(ctx->oz2)); ctx->oz2_rows = 0; }
  }
  return ctx->oz2.SL != nullptr;
}
static bool ensure_oz(agp_ctx* ctx, int64_t rows, int K, int S, cudaStream_t s) {
  if (!ctx->oz.SL || ctx->oz.K != K || ctx->oz_rows < rows || ctx->oz.S != S) {
    if (ctx->oz.SL) ozaki_ws_destroy(&ctx->oz, s);
    if (ozaki_ws_create(&ctx->oz, rows, K, S, s) == 0) ctx->oz_rows = rows;
    else { memset(&ctx->oz, 0, sizeof(ctx->oz)); ctx->oz_rows = 0; }
  }
  return ctx->oz.SL != nullptr;
}

// factor one outer panel in place: Lp points at its diagonal element; Gp inner 128-blocks; rows = rows from the
// panel's first row to the end of the (local) column storage (border rows included)
template <typename T>
void factor_panel_step(agp_ctx* ctx, T* Lp, int64_t lda, int Gp, int g, int64_t rows, T* Dinv_p, double* logdet_part,
                       int blk_base, int* info, cudaStream_t s) {
  {
    T* Akk = Lp + (int64_t)g * TILE + (int64_t)g * TILE * lda;
    const int64_t rows_below = rows - (int64_t)(g + 1) * TILE;
    bool split_done = false;
    if constexpr (std::is_same<T, double>::value) {
      if (potrf_split_enabled()) {
        launch_potrf_factor_f64(Akk, lda, logdet_part, blk_base + g, info, s);
        static const int64_t trsm_gemm_min = env_int64("AGP_TRSM_GEMM_MIN", 8192);
        if (rows_below >= trsm_gemm_min) {
          // tall panels: the substitution kernel (32 rows per CTA, 16 dependent steps) runs at ~4 TFLOP/s; the 128x128
          // inverse (8 CTAs, 14 us) followed by ONE in-place DMMA GEMM A21 <- A21 inv(L11)' is ~3x faster from ~8k rows
          launch_trtri_f64(Akk, lda, Dinv_p + (int64_t)g * TILE * TILE, s);
          GemmArgs t{};
          t.A = Akk + TILE; t.lda = lda; t.B = Dinv_p + (int64_t)g * TILE * TILE; t.ldb = TILE;
          t.C = Akk + TILE; t.ldc = lda; t.M = rows_below; t.N = TILE; t.K = TILE;
          launch_gemm<T>(t, s);
        } else {
          // short panels (latency-bound): factor on the main stream; the inverse (only the solves need it) on stream3;
          // the panel TRSM by blocked substitution straight from L11
          cudaEventRecord(ctx->ev_fac, s);
          cudaStreamWaitEvent(ctx->stream3, ctx->ev_fac, 0);
          launch_trtri_f64(Akk, lda, Dinv_p + (int64_t)g * TILE * TILE, ctx->stream3);
          ctx->s3_dirty = true;
          if (rows_below > 0) launch_trsm_sub_f64(Akk + TILE, lda, rows_below, Akk, s);
        }
        split_done = true;
      }
    }
    if (!split_done) {
      launch_potrf_diag<T>(Akk, lda, Dinv_p + (int64_t)g * TILE * TILE, logdet_part, blk_base + g, info, s);
      if (rows_below > 0) {
        GemmArgs t{};  // A21 <- A21 * inv(L11)'
        t.A = Akk + TILE; t.lda = lda; t.B = Dinv_p + (int64_t)g * TILE * TILE; t.ldb = TILE;
        t.C = Akk + TILE; t.ldc = lda; t.M = rows_below; t.N = TILE; t.K = TILE;
        launch_gemm<T>(t, s);
      }
    }
    if (rows_below <= 0) return;
    const int64_t ncols_in = (int64_t)(Gp - (g + 1)) * TILE;  // rank-128 update of the remaining inner columns
    if (ncols_in > 0) {
      GemmArgs u{};
      u.A = Akk + TILE; u.lda = lda; u.B = Akk + TILE; u.ldb = lda;
      u.C = Akk + TILE + (int64_t)TILE * lda; u.ldc = lda;
      u.M = rows_below; u.N = ncols_in; u.K = TILE; u.alpha_neg = 1; u.beta_one = 1; u.lower_only = 1;
      launch_gemm<T>(u, s);
    }
  }
}

// factor one outer panel in place (all inner blocks)
template <typename T>
void factor_panel(agp_ctx* ctx, T* Lp, int64_t lda, int Gp, int64_t rows, T* Dinv_p, double* logdet_part, int blk_base,
                  int* info, cudaStream_t s) {
  for (int g = 0; g < Gp; ++g) factor_panel_step<T>(ctx, Lp, lda, Gp, g, rows, Dinv_p, logdet_part, blk_base, info, s);
}

template <typename T>
void cholesky_inplace(agp_ctx* ctx, T* L, int64_t lda, int64_t n_pad, int64_t rows_total, T* Dinv,
                      double* logdet_part, int* info) {
  cudaStream_t s = ctx->stream, s2 = ctx->stream2;
  const int nblk = (int)(n_pad / TILE);
  const int fp64_mode = resolve_tensor_mode<T>(ctx, n_pad);  // 1: int8-sliced tcgen05 trailing update (fp64: 7 slices, fp32: 4)
  int G = resolve_G(ctx, n_pad);
  if (!std::is_same<T, double>::value && fp64_mode == 1 && ctx->cfg.tile_nb <= 0 && n_pad >= 4096) G = 4;  // 512-wide panels
  const bool oz_ok = fp64_mode == 1 && nblk > 2 * G && ensure_oz(ctx, rows_total, G * TILE, slices_of<T>(ctx), s) &&
                     (std::is_same<T, double>::value || ctx->oz.bulk == 2);
  const bool la = ctx->cfg.lookahead != 0 && nblk > 2 * G;
  const bool la2 = la && ctx->cfg.lookahead >= 2;
  bool rest_pending = false, restA_pending = false, last_rest_full = false;
  size_t ev_idx = 0, last_rest = 0, last_restA = 0;
  for (int ko = 0; ko < nblk; ko += G) {
    const int g_end = (ko + G < nblk) ? ko + G : nblk;  // inner blocks [ko, g_end)
    factor_panel<T>(ctx, L + (int64_t)ko * TILE + (int64_t)ko * TILE * lda, lda, g_end - ko, rows_total - (int64_t)ko * TILE,
                    Dinv + (int64_t)ko * TILE * TILE, logdet_part, ko, info, s);
    const int64_t t0 = (int64_t)g_end * TILE;           // first trailing row/column
    const int64_t cols_trail = n_pad - t0;
    if (cols_trail <= 0) continue;
    const int64_t K = (int64_t)(g_end - ko) * TILE, kc0 = (int64_t)ko * TILE;
    const OzakiWs* oz = nullptr;
    // tcgen05 path: needs the full outer-panel width it was sized for and enough trailing work to pay for slicing
    if (oz_ok && K == ctx->oz.K && cols_trail >= 2 * TILE) oz = &ctx->oz;
    constexpr int is_f32 = std::is_same<T, double>::value ? 0 : 1;
    if (!la) {
      if (oz) ozaki_prepare_ex(*oz, L + t0 + kc0 * lda, is_f32, 0, lda, rows_total - t0, 0, s);
      trailing_update<T>(ctx, L, lda, t0, t0, kc0, K, rows_total - t0, cols_trail, s, oz, t0);
      continue;
    }
    if (la2 && !oz) {
      // EXPERIMENTAL look-ahead depth 2 (cfg.lookahead = 2 / AGP_LOOKAHEAD=2, DMMA path only -- the tcgen05 path would
      // need a second slice buffer; not yet run on a device).  The rest update is split: restA = the panel after next,
      // restB = everything beyond.  The next-panel update of step k+1 needs only restA(k), so the latency-bound chain
      // (potrf -> TRSM -> next-panel update) no longer waits for the bulk of the previous rest update; restB(k) has two
      // chain steps to finish instead of one (it is serialised behind restB(k-1) and restA(k) on the side stream).
      cudaEvent_t e_panel = dep_event(ctx, ev_idx++), e_restA = dep_event(ctx, ev_idx++), e_restB = dep_event(ctx, ev_idx++);
      if (restA_pending) cudaStreamWaitEvent(s, dep_event(ctx, last_restA), 0);
      // a previous step may have used the depth-1 branch (tcgen05 panel followed by a short DMMA tail): its rest update
      // on the side stream touches the same block column as the update below
      if (rest_pending && last_rest_full) cudaStreamWaitEvent(s, dep_event(ctx, last_rest), 0);
      cudaEventRecord(e_panel, s);
      const int64_t next_cols = (cols_trail < (int64_t)G * TILE) ? cols_trail : (int64_t)G * TILE;
      trailing_update<T>(ctx, L, lda, t0, t0, kc0, K, rows_total - t0, next_cols, s, nullptr, t0);
      restA_pending = false;
      rest_pending = false;
      if (cols_trail > next_cols) {
        const int64_t r0 = t0 + next_cols;
        const int64_t colsA = (n_pad - r0 < (int64_t)G * TILE) ? (n_pad - r0) : (int64_t)G * TILE;
        cudaStreamWaitEvent(s2, e_panel, 0);
        trailing_update<T>(ctx, L, lda, r0, r0, kc0, K, rows_total - r0, colsA, s2, nullptr, t0);
        cudaEventRecord(e_restA, s2);
        restA_pending = true;
        last_restA = ev_idx - 2;
        const int64_t r1 = r0 + colsA;
        if (n_pad > r1) trailing_update<T>(ctx, L, lda, r1, r1, kc0, K, rows_total - r1, n_pad - r1, s2, nullptr, t0);
        cudaEventRecord(e_restB, s2);
        rest_pending = true;
        last_rest_full = false;  // restB never touches the next panel's columns
        last_rest = ev_idx - 1;
      }
      continue;
    }
    cudaEvent_t e_panel = dep_event(ctx, ev_idx++), e_rest = dep_event(ctx, ev_idx++);
    if (rest_pending) cudaStreamWaitEvent(s, dep_event(ctx, last_rest), 0);  // also frees the slice buffer
    if (restA_pending) { cudaStreamWaitEvent(s, dep_event(ctx, last_restA), 0); restA_pending = false; }
    if (oz) ozaki_prepare_ex(*oz, L + t0 + kc0 * lda, is_f32, 0, lda, rows_total - t0, 0, s);
    cudaEventRecord(e_panel, s);
    const int64_t next_cols = (cols_trail < (int64_t)G * TILE) ? cols_trail : (int64_t)G * TILE;
    trailing_update<T>(ctx, L, lda, t0, t0, kc0, K, rows_total - t0, next_cols, s, oz, t0);  // next outer panel first
    rest_pending = false;
    if (cols_trail > next_cols) {
      const int64_t r0 = t0 + next_cols;
      cudaStreamWaitEvent(s2, e_panel, 0);
      if (oz) oz->chunk_tiles = ctx->oz_chunk;  // the panel chain on the (higher-priority) main stream gets SMs between CTAs
      trailing_update<T>(ctx, L, lda, r0, r0, kc0, K, rows_total - r0, n_pad - r0, s2, oz, t0);
      if (oz) oz->chunk_tiles = 0;
      cudaEventRecord(e_rest, s2);
      rest_pending = true;
      last_rest_full = true;
      last_rest = ev_idx - 1;
    }
  }
  if (rest_pending) cudaStreamWaitEvent(s, dep_event(ctx, last_rest), 0);
  join_inverses(ctx);  // Dinv (stream3) is complete before any solve on the main stream
}

// V <- L^-1 V on the tensor cores: two-level blocked substitution.  Inside an outer block of W = 512 rows the 128-step
// loop runs on the tile GEMMs (W^2 ncols work); everything below the block receives ONE rank-W update
//   B[below] -= L[below, block] * B[block]
// on the int8-sliced tcgen05 kernel: the L panel (rows_below x W, row-contiguous) and B[block]' (ncols x W, k-major) are
// sliced into one workspace (A rows first, B rows behind them) and the product is a rectangular update of B[below].
// This is `C.U' \ X` of /root/reference/src/util/common_covmat_ops.jl:54,90 (prediction, N^2 M flops at config C3) and the
// A = U' \ K_zx solve of /root/reference/src/sparse_approximations.jl:296.
template <typename T>
bool forward_subst_multi_tc(agp_ctx* ctx, const T* L, int64_t lda, const T* Dinv, int64_t n_pad, T* B, int64_t ldb,
                            int64_t ncols) {
  cudaStream_t s = ctx->stream;
  constexpr int GW = 4;
  const int64_t W = (int64_t)GW * TILE;
  const int nblk = (int)(n_pad / TILE);
  constexpr int is_f32 = std::is_same<T, double>::value ? 0 : 1;
  const int64_t a_rows = round_up(n_pad, TILE);
  if (!ensure_oz(ctx, a_rows + ncols, (int)W, slices_of<T>(ctx), s) || ctx->oz.bulk != 2) return false;
  const OzakiWs& ws = ctx->oz;
  for (int ko = 0; ko < nblk; ko += GW) {
    const int k_end = (ko + GW < nblk) ? ko + GW : nblk;
    const int64_t blk_end = (int64_t)k_end * TILE;
    for (int k = ko; k < k_end; ++k) {  // in-block substitution (128-steps, rows of this outer block only)
      T* Bk = B + (int64_t)k * TILE;
      GemmArgs a{};
      a.A = Dinv + (int64_t)k * TILE * TILE; a.lda = TILE; a.a_kmajor = 0;
      a.B = Bk; a.ldb = ldb; a.b_kmajor = 1;
      a.C = Bk; a.ldc = ldb; a.M = TILE; a.N = ncols; a.K = TILE;
      launch_gemm<T>(a, s);
      const int64_t rows_in = blk_end - (int64_t)(k + 1) * TILE;
      if (rows_in <= 0) continue;
      GemmArgs u{};
      u.A = L + (int64_t)(k + 1) * TILE + (int64_t)k * TILE * lda; u.lda = lda; u.a_kmajor = 0;
      u.B = Bk; u.ldb = ldb; u.b_kmajor = 1;
      u.C = Bk + TILE; u.ldc = ldb; u.M = rows_in; u.N = ncols; u.K = TILE; u.alpha_neg = 1; u.beta_one = 1;
      launch_gemm<T>(u, s);
    }
    const int64_t rows_below = n_pad - blk_end, Kw = blk_end - (int64_t)ko * TILE;
    if (rows_below <= 0) continue;
    if (Kw != W) {  // ragged last outer block (n_pad not a multiple of 512): finish on the tile GEMM
      GemmArgs u{};
      u.A = L + blk_end + (int64_t)ko * TILE * lda; u.lda = lda; u.a_kmajor = 0;
      u.B = B + (int64_t)ko * TILE; u.ldb = ldb; u.b_kmajor = 1;
      u.C = B + blk_end; u.ldc = ldb; u.M = rows_below; u.N = ncols; u.K = Kw; u.alpha_neg = 1; u.beta_one = 1;
      launch_gemm<T>(u, s);
      continue;
    }
    ozaki_prepare_ex(ws, L + blk_end + (int64_t)ko * TILE * lda, is_f32, 0, lda, rows_below, 0, s);
    ozaki_prepare_ex(ws, B + (int64_t)ko * TILE, is_f32, 1, ldb, ncols, a_rows, s);
    if (ozaki_update_ex(ws, B + blk_end, is_f32, ldb, rows_below, ncols, 1, -1.0, 0, 0, a_rows, 0, s) != 0) return false;
  }
  return true;
}

// V <- L^-1 V for a n_pad x ncols block of right-hand sides (ncols multiple of 4), in place
template <typename T>
void forward_subst_multi(agp_ctx* ctx, const T* L, int64_t lda, const T* Dinv, int64_t n_pad, T* B, int64_t ldb,
                         int64_t ncols) {
  cudaStream_t s = ctx->stream;
  const int nblk = (int)(n_pad / TILE);
  static const int64_t tc_min_cols = env_int64("AGP_SOLVE_TC_MIN_COLS", 512);
  if (resolve_tensor_mode<T>(ctx, n_pad >= 2048 ? (int64_t)1 << 20 : 0) == 1 && n_pad >= 2048 && ncols % TILE == 0 &&
      ncols >= tc_min_cols) {
    if (forward_subst_multi_tc<T>(ctx, L, lda, Dinv, n_pad, B, ldb, ncols)) return;
  }
  for (int k = 0; k < nblk; ++k) {
    T* Bk = B + (int64_t)k * TILE;
    GemmArgs a{};
    a.A = Dinv + (int64_t)k * TILE * TILE; a.lda = TILE; a.a_kmajor = 0;
    a.B = Bk; a.ldb = ldb; a.b_kmajor = 1;
    a.C = Bk; a.ldc = ldb; a.M = TILE; a.N = ncols; a.K = TILE;
    launch_gemm<T>(a, s);
    const int64_t rows_below = n_pad - (int64_t)(k + 1) * TILE;
    if (rows_below <= 0) continue;
    GemmArgs u{};
    u.A = L + (int64_t)(k + 1) * TILE + (int64_t)k * TILE * lda; u.lda = lda; u.a_kmajor = 0;
    u.B = Bk; u.ldb = ldb; u.b_kmajor = 1;
    u.C = Bk + TILE; u.ldc = ldb; u.M = rows_below; u.N = ncols; u.K = TILE; u.alpha_neg = 1; u.beta_one = 1;
    launch_gemm<T>(u, s);
  }
}

template <typename T>
void fill_gram_params(GramParams& gp, const agp_kernel* k, int symmetric, int lower_only, int64_t va, int64_t vb,
                      const agp_noise* noise, const T* noise_v_dev) {
  gp.family = k->family;
  gp.variance = k->variance;
  gp.linear_c = k->linear_c;
  gp.symmetric = symmetric;
  gp.lower_only = lower_only;
  gp.valid_a = va;
  gp.valid_b = vb;
  gp.noise_kind = noise ? noise->kind : -1;
  gp.noise_s = noise ? noise->s : 0.0;
  gp.noise_v = noise_v_dev;
}

// transformed, padded, point-major copy of a point set on the device
template <typename T>
int prep_points(agp_ctx* ctx, Scratch& sc, const agp_kernel* k, const T* ard_dev, int layout, const void* X,
                int64_t n, int64_t n_pad, int D, T** Xt_out, bool keep) {
  T* Xd = nullptr;
  int rc = upload<T>(ctx, sc, X, (size_t)n * D, false, &Xd);
  if (rc) return rc;
  void* xt = nullptr;
  CK(cudaMallocAsync(&xt, (size_t)(n_pad > 0 ? n_pad : 1) * D * sizeof(T), ctx->stream));
  if (!keep) sc.ptrs.push_back(xt);
  launch_prep_points<T>(Xd, layout, n, n_pad, D, k->transform, k->scale, ard_dev, (T*)xt, ctx->stream);
  *Xt_out = (T*)xt;
  return AGP_OK;
}

template <typename T>
int fit_impl(agp_ctx* ctx, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise, int layout,
             const void* X, int64_t N, int D, const void* Y, int S, void* logpdf_out, void* alpha_out,
             agp_post** post_out, T** L_keep /*optional raw factor out for rand*/, int64_t* lda_out,
             Scratch* outer_sc) {
  int rc = check_kernel(ctx, k, D);
  if (rc) return rc;
  if (N <= 0) { ctx->err = "N must be positive"; return AGP_ERR_DIM_MISMATCH; }
  if (S < 0 || S > TILE) { ctx->err = "number of right-hand sides must be in [0,128]"; return AGP_ERR_UNSUPPORTED; }
  if (S > 0 && !Y) { ctx->err = "Y is NULL"; return AGP_ERR_INVALID; }
  static const agp_mean zero_mean{0, 0.0, nullptr};
  static const agp_noise default_noise{0, 1e-18, nullptr};  // default_sigma^2, finite_gp_projection.jl:17
  if (!mean) mean = &zero_mean;
  if (!noise) noise = &default_noise;
  if (mean->kind == 2 && !mean->v) { ctx->err = "mean vector is NULL"; return AGP_ERR_INVALID; }
  if (noise->kind == 1 && !noise->v) { ctx->err = "noise vector is NULL"; return AGP_ERR_INVALID; }

  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  const int64_t n_pad = round_up(N, TILE), lda = n_pad + TILE;
  const int nblk = (int)(n_pad / TILE);
  prof_begin(ctx);
  CK(cudaEventRecord(ctx->ev[0], s));

  // the factor is allocated FIRST: the stream-ordered pool then hands back the block the previous fit freed, before the small
  // staging buffers below can split it (a split forces a fresh multi-GB allocation from the OS: 0.1-0.6 s, seen as a
  // sporadic e2e-only slowdown in round 2, tools/dist1_probe.py)
  void* Lv = nullptr;
  CK(cudaMallocAsync(&Lv, (size_t)lda * n_pad * sizeof(T), s));
  struct BufGuard { void* p; cudaStream_t s; ~BufGuard() { if (p) cudaFreeAsync(p, s); } } lguard{Lv, s};  // error returns
  // ---- H2D
  T *ard_d = nullptr, *mean_d = nullptr, *noise_d = nullptr, *Yd = nullptr, *Xt = nullptr;
  if (k->transform == AGP_T_ARD) { rc = upload<T>(ctx, sc, k->ard, D, true, &ard_d); if (rc) return rc; }
  if (mean->kind == 2) { rc = upload<T>(ctx, sc, mean->v, N, true, &mean_d); if (rc) return rc; }
  if (noise->kind == 1) { rc = upload<T>(ctx, sc, noise->v, N, true, &noise_d); if (rc) return rc; }
  if (S > 0) { rc = upload<T>(ctx, sc, Y, (size_t)N * S, false, &Yd); if (rc) return rc; }
  const bool keep = (post_out != nullptr);
  rc = prep_points<T>(ctx, sc, k, ard_d, layout, X, N, n_pad, D, &Xt, keep);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[1], s));

  // ---- buffers
  void *Dinvv = nullptr, *alphav = nullptr;
  lguard.p = nullptr;  // ownership passes to the handle / scratch lists below
  CK(cudaMallocAsync(&Dinvv, (size_t)nblk * TILE * TILE * sizeof(T), s));
  CK(cudaMallocAsync(&alphav, (size_t)n_pad * sizeof(T), s));
  T* L = (T*)Lv; T* Dinv = (T*)Dinvv; T* alpha = (T*)alphav;
  agp_post* post = nullptr;
  struct PostGuard { agp_post* p; ~PostGuard() { if (p) agp_post_free(p); } } pguard{nullptr};  // early error returns
  if (keep) {
    post = new agp_post();
    pguard.p = post;
    post->ctx = ctx; post->dtype = sizeof(T) == 8 ? AGP_F64 : AGP_F32;
    post->n = N; post->n_pad = n_pad; post->lda = lda; post->D = D;
    post->L = Lv; post->Dinv = Dinvv; post->Xt = Xt; post->alpha = alphav;
    post->k = *k; post->k.ard = nullptr;
    post->mean_kind = mean->kind == 2 ? 0 : mean->kind; post->mean_c = mean->c;
    post->segs.push_back({0, N});
    CK(cudaMallocAsync(&post->delta, (size_t)n_pad * sizeof(T), s));
    if (ard_d) { sc.release(ard_d); post->ard = ard_d; }
  } else if (L_keep) {
    outer_sc->ptrs.push_back(Lv); sc.ptrs.push_back(Dinvv); sc.ptrs.push_back(alphav);
  } else {
    sc.ptrs.push_back(Lv); sc.ptrs.push_back(Dinvv); sc.ptrs.push_back(alphav);
  }
  double* dscal = nullptr;  // [0..nblk) logdet parts, [nblk..nblk+TILE) sqmahal, [+1] logdet
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)(nblk + TILE + 2) * sizeof(double)));
  dscal = (double*)tmp;
  int* dinfo = nullptr;
  CK(sc.alloc(&tmp, sizeof(int)));
  dinfo = (int*)tmp;
  CK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  T* rwork = nullptr;
  CK(sc.alloc(&tmp, (size_t)(S > 0 ? S : 1) * n_pad * sizeof(T)));
  rwork = (T*)tmp;
  int* dflags = nullptr;
  CK(sc.alloc(&tmp, (size_t)(nblk + 1) * sizeof(int)));
  dflags = (int*)tmp;
  T* lp_d = nullptr;
  CK(sc.alloc(&tmp, (size_t)TILE * sizeof(T)));
  lp_d = (T*)tmp;

  // ---- Gram (+ noise) straight into the factor buffer, border rows = delta'
  GramParams gp{};
  fill_gram_params<T>(gp, k, 1, 1, N, N, noise, noise_d);
  launch_gram<T>(Xt, Xt, n_pad, n_pad, D, L, lda, gp, s);
  launch_border_init<T>(L, lda, N, n_pad, Yd, N, S, mean->kind, mean->c, mean_d, s);
  if (keep) {
    if (S > 0) launch_extract_v<T>(L, lda, n_pad, 1, (T*)post->delta, dscal + nblk, s);  // delta = border row 0
    else CK(cudaMemsetAsync(post->delta, 0, (size_t)n_pad * sizeof(T), s));
  }
  CK(cudaEventRecord(ctx->ev[2], s));

  // ---- Cholesky (forward substitution of delta rides along in the border tile)
  cholesky_inplace<T>(ctx, L, lda, n_pad, lda, Dinv, dscal, dinfo);
  CK(cudaEventRecord(ctx->ev[3], s));

  // ---- sqmahal, alpha = L^-T v, logpdf
  if (S > 0) {
    launch_extract_v<T>(L, lda, n_pad, S, rwork, dscal + nblk, s);
    if (alpha_out || keep) {
      launch_bwd_solve<T>(L, lda, Dinv, nblk, rwork, dflags, s);
      CK(cudaMemcpyAsync(alpha, rwork, (size_t)n_pad * sizeof(T), cudaMemcpyDeviceToDevice, s));
    }
  }
  launch_finalize_logpdf<T>(dscal, nblk, dscal + nblk, S, N, lp_d, dscal + nblk + TILE, s);
  CK(cudaEventRecord(ctx->ev[4], s));

  // ---- D2H
  int h_info = 0;
  double h_logdet = 0.0;
  CK(cudaMemcpyAsync(&h_info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h_logdet, dscal + nblk + TILE, sizeof(double), cudaMemcpyDeviceToHost, s));
  if (S > 0 && logpdf_out) CK(cudaMemcpyAsync(logpdf_out, lp_d, (size_t)S * sizeof(T), cudaMemcpyDeviceToHost, s));
  if (S > 0 && alpha_out) { rc = download<T>(ctx, alpha_out, alpha, (size_t)N, false); if (rc) return rc; }
  CK(cudaEventRecord(ctx->ev[5], s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());

  float ms = 0;
  auto el = [&](int a, int b) { cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]); return (double)ms; };
  ctx->timings[0] = el(0, 5); ctx->timings[1] = el(0, 1); ctx->timings[2] = el(1, 2); ctx->timings[3] = el(2, 3);
  ctx->timings[4] = el(3, 4); ctx->timings[5] = el(4, 5); ctx->timings[6] = 0.0;
  ctx->timings[7] = ctx->profile ? prof_total_ms(ctx) : 0.0;

  if (h_info != 0) {
    ctx->info = h_info;
    char b[128];
    snprintf(b, sizeof(b), "matrix is not positive definite; Cholesky failed at pivot %d", h_info);
    ctx->err = b;
    return AGP_ERR_NOT_POSDEF;  // the guard releases the handle
  }
  pguard.p = nullptr;
  if (post) { post->logdet = h_logdet; *post_out = post; }
  if (L_keep) { *L_keep = L; *lda_out = lda; }
  return AGP_OK;
}

// logpdf(fx, Y::Matrix) has no limit on the number of columns (/root/reference/src/finite_gp_projection.jl:306-311).
// The border tile carries 128 right-hand sides through the factorisation; further columns reuse the SAME factor:
// V = L^-1 (Y_c - m) by the multi-RHS forward substitution, sqmahal = column sums of V.^2 -- ONE Gram and ONE Cholesky
// whatever S is.
template <typename T>
int fit_many_impl(agp_ctx* ctx, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise, int layout,
                  const void* X, int64_t N, int D, const void* Y, int S, void* logpdf_out, void* alpha_out,
                  agp_post** post_out) {
  if (S <= TILE) return fit_impl<T>(ctx, k, mean, noise, layout, X, N, D, Y, S, logpdf_out, alpha_out, post_out, nullptr, nullptr, nullptr);
  if (!Y || !logpdf_out) { ctx->err = "Y/logpdf_out is NULL"; return AGP_ERR_INVALID; }
  agp_post* p = nullptr;
  int rc = fit_impl<T>(ctx, k, mean, noise, layout, X, N, D, Y, TILE, logpdf_out, alpha_out, &p, nullptr, nullptr, nullptr);
  if (rc) return rc;
  struct Guard { agp_post* p; bool keep; ~Guard() { if (p && !keep) agp_post_free(p); } } guard{p, false};
  cudaStream_t s = ctx->stream;
  static const agp_mean zero_mean{0, 0.0, nullptr};
  if (!mean) mean = &zero_mean;
  const int64_t n_pad = p->n_pad;
  const double log2pi = 1.8378770664093454835606594728112;
  const int64_t chunk = 1024;
  Scratch sc(ctx);
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)n_pad * chunk * sizeof(T)));
  T* B = (T*)tmp;
  CK(sc.alloc(&tmp, (size_t)chunk * sizeof(T)));
  T* sq = (T*)tmp;
  T* mean_d = nullptr;
  if (mean->kind == 2) { rc = upload<T>(ctx, sc, mean->v, N, true, &mean_d); if (rc) return rc; }
  std::vector<T> h_sq((size_t)chunk);
  for (int64_t c0 = TILE; c0 < S; c0 += chunk) {
    const int64_t nc = (S - c0 < chunk) ? (S - c0) : chunk, nc_pad = round_up(nc, TILE);
    Scratch scc(ctx);
    T* Yd = nullptr;
    rc = upload<T>(ctx, scc, (const T*)Y + (size_t)c0 * N, (size_t)N * nc, false, &Yd);
    if (rc) return rc;
    CK(cudaMemsetAsync(B, 0, (size_t)n_pad * nc_pad * sizeof(T), s));
    CK(cudaMemsetAsync(sq, 0, (size_t)chunk * sizeof(T), s));
    for (int64_t j = 0; j < nc; ++j) launch_sub_mean<T>(Yd + j * N, N, mean->kind, mean->c, mean_d, B + j * n_pad, s);
    forward_subst_multi<T>(ctx, (const T*)p->L, p->lda, (const T*)p->Dinv, n_pad, B, n_pad, nc_pad);
    launch_colsumsq_acc<T>(B, n_pad, n_pad, nc, 1.0, sq, s);
    CK(cudaMemcpyAsync(h_sq.data(), sq, (size_t)nc * sizeof(T), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int64_t j = 0; j < nc; ++j)
      ((T*)logpdf_out)[c0 + j] = (T)(-0.5 * ((double)N * log2pi + p->logdet + (double)h_sq[(size_t)j]));
  }
  CK(cudaGetLastError());
  if (post_out) { *post_out = p; guard.keep = true; }
  return AGP_OK;
}

// rows [c0, c0 + mc) of a feature-major point set (M x D column-major, host or device per the context's memspace) gathered
// into a contiguous mc x D feature-major DEVICE block: chunked prediction over RowVecs inputs larger than one chunk
template <typename T>
int gather_feature_major_chunk(agp_ctx* ctx, Scratch& sc, const void* Xs, int64_t M, int D, int64_t c0, int64_t mc, T** out) {
  void* d = nullptr;
  CK(sc.alloc(&d, (size_t)mc * D * sizeof(T)));
  const cudaMemcpyKind kind = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CK(cudaMemcpy2DAsync(d, (size_t)mc * sizeof(T), (const T*)Xs + c0, (size_t)M * sizeof(T), (size_t)mc * sizeof(T), (size_t)D, kind,
                       ctx->stream));
  *out = (T*)d;
  return AGP_OK;
}

template <typename T>
int post_cross(agp_post* p, Scratch& sc, int layout, const void* Xs, int64_t M, int64_t m_pad, T** Xst, T** B) {
  agp_ctx* ctx = p->ctx;
  int rc = prep_points<T>(ctx, sc, &p->k, (const T*)p->ard, layout, Xs, M, m_pad, p->D, Xst, false);
  if (rc) return rc;
  void* b = nullptr;
  CK(sc.alloc(&b, (size_t)p->n_pad * m_pad * sizeof(T)));
  *B = (T*)b;
  GramParams gp{};
  fill_gram_params<T>(gp, &p->k, 0, 0, p->n, M, nullptr, nullptr);
  gp.mask_a = p->valid;
  launch_gram<T>((const T*)p->Xt, *Xst, p->n_pad, m_pad, p->D, *B, p->n_pad, gp, ctx->stream);
  return AGP_OK;
}

template <typename T>
int post_mean_var_impl(agp_post* p, int layout, const void* Xs, int64_t M, const agp_mean* mean_s,
                       const agp_noise* noise_s, void* mean_out, void* var_out) {
  agp_ctx* ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  if (M <= 0) return AGP_OK;
  CK(cudaEventRecord(ctx->ev[0], s));
  // chunk the test points so that the N x Mc cross-Gram stays <= ~4 GB
  int64_t cap = (int64_t)(4.0e9 / ((double)p->n_pad * sizeof(T)));
  { const int64_t c_env = env_int64("AGP_PREDICT_CHUNK", 0); if (c_env > 0) cap = c_env; }  // tests: force small chunks
  cap = cap / TILE * TILE;
  if (cap < TILE) cap = TILE;
  agp_mean mz{p->mean_kind, p->mean_c, nullptr};
  if (!mean_s) mean_s = &mz;
  for (int64_t c0 = 0; c0 < M; c0 += cap) {
    const int64_t mc = (M - c0 < cap) ? (M - c0) : cap;
    const int64_t m_pad = round_up(mc, TILE);
    Scratch sc(ctx);
    const char* xs_c = (const char*)Xs;
    const void* xs_chunk = nullptr;
    const int saved_memspace = ctx->memspace;
    if (layout == AGP_POINT_MAJOR) {
      xs_chunk = xs_c + (size_t)c0 * p->D * sizeof(T);
    } else if (c0 == 0 && mc == M) {
      xs_chunk = Xs;
    } else {  // RowVecs test set larger than one chunk: gather the chunk's rows on the device
      T* g = nullptr;
      int grc = gather_feature_major_chunk<T>(ctx, sc, Xs, M, p->D, c0, mc, &g);
      if (grc) return grc;
      xs_chunk = g;
      ctx->memspace = AGP_MEM_DEVICE;  // the gathered block is a device pointer (inputs only; outputs use the override below)
    }
    T *Xst = nullptr, *B = nullptr;
    int rc = post_cross<T>(p, sc, layout, xs_chunk, mc, m_pad, &Xst, &B);
    const bool out_dev_saved = ctx->out_dev_override;
    if (ctx->memspace != saved_memspace) { ctx->memspace = saved_memspace; }
    (void)out_dev_saved;
    if (rc) return rc;
    T *mean_d = nullptr, *noise_d = nullptr;
    if (mean_s->kind == 2) { rc = upload<T>(ctx, sc, (const T*)mean_s->v + c0, mc, true, &mean_d); if (rc) return rc; }
    if (noise_s && noise_s->kind == 1) { rc = upload<T>(ctx, sc, (const T*)noise_s->v + c0, mc, true, &noise_d); if (rc) return rc; }
    void* tmp = nullptr;
    CK(sc.alloc(&tmp, (size_t)m_pad * 3 * sizeof(T)));
    T* mu = (T*)tmp; T* var = mu + m_pad; T* kd = var + m_pad;
    launch_kdiag<T>(Xst, mc, p->D, p->k.family, p->k.variance, p->k.linear_c, kd, s);
    launch_gemv_t<T>(B, p->n_pad, p->n_pad, mc, (const T*)p->alpha, mean_s->kind, mean_s->c, mean_d, mu, s);
    if (var_out) {
      forward_subst_multi<T>(ctx, (const T*)p->L, p->lda, (const T*)p->Dinv, p->n_pad, B, p->n_pad, m_pad);
      launch_colsumsq_var<T>(B, p->n_pad, p->n_pad, mc, kd, noise_s ? noise_s->kind : -1, noise_s ? noise_s->s : 0.0,
                             noise_d, var, s);
    }
    if (mean_out) { rc = download<T>(ctx, (T*)mean_out + c0, mu, mc, false); if (rc) return rc; }
    if (var_out) { rc = download<T>(ctx, (T*)var_out + c0, var, mc, false); if (rc) return rc; }
    CK(cudaStreamSynchronize(s));
  }
  CK(cudaEventRecord(ctx->ev[1], s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  ctx->timings[6] = ms;
  ctx->timings[0] = ms;
  return AGP_OK;
}


// ---- distributed posterior: gather the block-column-cyclic factor so that every rank holds it whole (SURVEY s8e:
// "partition the M test points across GPUs if the factor is replicated").  Collective: every rank of the communicator
// calls it at the same point (all handle operations are SPMD, like agp_fit on a distributed context).
template <typename T>
int post_replicate(agp_post* p) {
  if (!p->Lloc) return AGP_OK;
  agp_ctx* ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  const int R = p->dist_R, me = p->dist_me, G = p->dist_G;
  const int64_t W = p->dist_W, lda = p->lda, n_pad = p->n_pad;
  const int nto = (int)(n_pad / W), nt = (int)(n_pad / TILE);
  void *Lf = nullptr, *Df = nullptr;
  CK(cudaMallocAsync(&Lf, (size_t)lda * n_pad * sizeof(T), s));
  CK(cudaMallocAsync(&Df, (size_t)nt * TILE * TILE * sizeof(T), s));
  for (int jo = 0; jo < nto; ++jo) {
    const int owner = jo % R, lj = jo / R;
    T* dstL = (T*)Lf + (int64_t)jo * W * lda;
    T* dstD = (T*)Df + (int64_t)jo * G * TILE * TILE;
    const T* srcL = owner == me ? (const T*)p->Lloc + (int64_t)lj * W * lda : dstL;
    const T* srcD = owner == me ? (const T*)p->Dinv_loc + (int64_t)lj * G * TILE * TILE : dstD;
    CKN(ncclBroadcast(srcL, dstL, (size_t)lda * W, NcclType<T>::v, owner, ctx->nccl, s));
    CKN(ncclBroadcast(srcD, dstD, (size_t)G * TILE * TILE, NcclType<T>::v, owner, ctx->nccl, s));
  }
  CK(cudaStreamSynchronize(s));
  cudaFreeAsync(p->Lloc, s);
  cudaFreeAsync(p->Dinv_loc, s);
  p->Lloc = nullptr; p->Dinv_loc = nullptr;
  p->L = Lf; p->Dinv = Df;
  return AGP_OK;
}

// mean_and_var over a distributed posterior: the factor is replicated (once), the test points are partitioned over the
// ranks (point-major inputs), every rank predicts its slice with the single-GPU path and the M-vectors are exchanged
// with one broadcast per rank.  Every rank returns the complete outputs.
template <typename T>
int post_mean_var_dist(agp_post* p, int layout, const void* Xs, int64_t M, const agp_mean* mean_s,
                       const agp_noise* noise_s, void* mean_out, void* var_out) {
  agp_ctx* ctx = p->ctx;
  int rc = post_replicate<T>(p);
  if (rc) return rc;
  const int R = p->dist_R, me = p->dist_me;
  if (layout != AGP_POINT_MAJOR || M < 2 * R)  // small or feature-major: every rank computes everything (same result)
    return post_mean_var_impl<T>(p, layout, Xs, M, mean_s, noise_s, mean_out, var_out);
  cudaStream_t s = ctx->stream;
  Scratch sc(ctx);
  const int64_t per = (M + R - 1) / R, lo = (int64_t)me * per < M ? (int64_t)me * per : M;
  const int64_t hi = lo + per < M ? lo + per : M;
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)2 * M * sizeof(T)));
  T* mu_all = (T*)tmp; T* var_all = mu_all + M;
  agp_mean ms; agp_noise ns;
  const agp_mean* msp = nullptr; const agp_noise* nsp = nullptr;
  if (mean_s) { ms = *mean_s; if (ms.kind == 2 && ms.v) ms.v = (const T*)ms.v + lo; msp = &ms; }
  if (noise_s) { ns = *noise_s; if (ns.kind == 1 && ns.v) ns.v = (const T*)ns.v + lo; nsp = &ns; }
  if (hi > lo) {
    ctx->out_dev_override = true;
    rc = post_mean_var_impl<T>(p, layout, (const char*)Xs + (size_t)lo * p->D * sizeof(T), hi - lo, msp, nsp,
                               mean_out ? mu_all + lo : nullptr, var_out ? var_all + lo : nullptr);
    ctx->out_dev_override = false;
    if (rc) return rc;
  }
  for (int r = 0; r < R; ++r) {
    const int64_t rlo = (int64_t)r * per < M ? (int64_t)r * per : M, rhi = rlo + per < M ? rlo + per : M;
    if (rhi <= rlo) continue;
    if (mean_out) CKN(ncclBroadcast(mu_all + rlo, mu_all + rlo, (size_t)(rhi - rlo), NcclType<T>::v, r, ctx->nccl, s));
    if (var_out) CKN(ncclBroadcast(var_all + rlo, var_all + rlo, (size_t)(rhi - rlo), NcclType<T>::v, r, ctx->nccl, s));
  }
  if (mean_out) { rc = download<T>(ctx, mean_out, mu_all, (size_t)M, false); if (rc) return rc; }
  if (var_out) { rc = download<T>(ctx, var_out, var_all, (size_t)M, false); if (rc) return rc; }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return AGP_OK;
}

template <typename T>
int post_mean_cov_impl(agp_post* p, int layout, const void* Xs, int64_t M, const agp_mean* mean_s, void* mean_out,
                       void* cov_out) {
  agp_ctx* ctx = p->ctx;
  { int rrc = post_replicate<T>(p); if (rrc) return rrc; }
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  if (M <= 0) return AGP_OK;
  const int64_t m_pad = round_up(M, TILE);
  Scratch sc(ctx);
  T *Xst = nullptr, *B = nullptr;
  int rc = post_cross<T>(p, sc, layout, Xs, M, m_pad, &Xst, &B);
  if (rc) return rc;
  agp_mean mz{p->mean_kind, p->mean_c, nullptr};
  if (!mean_s) mean_s = &mz;
  T* mean_d = nullptr;
  if (mean_s->kind == 2) { rc = upload<T>(ctx, sc, mean_s->v, M, true, &mean_d); if (rc) return rc; }
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)m_pad * sizeof(T)));
  T* mu = (T*)tmp;
  launch_gemv_t<T>(B, p->n_pad, p->n_pad, M, (const T*)p->alpha, mean_s->kind, mean_s->c, mean_d, mu, s);
  if (mean_out) { rc = download<T>(ctx, mean_out, mu, M, false); if (rc) return rc; }
  if (cov_out) {
    forward_subst_multi<T>(ctx, (const T*)p->L, p->lda, (const T*)p->Dinv, p->n_pad, B, p->n_pad, m_pad);
    CK(sc.alloc(&tmp, (size_t)m_pad * m_pad * sizeof(T) * 2));
    T* C = (T*)tmp; T* Kss = C + m_pad * m_pad;
    GemmArgs g{};  // C = V' V
    g.A = B; g.lda = p->n_pad; g.a_kmajor = 1;
    g.B = B; g.ldb = p->n_pad; g.b_kmajor = 1;
    g.C = C; g.ldc = m_pad; g.M = m_pad; g.N = m_pad; g.K = p->n_pad;
    launch_gemm<T>(g, s);
    GramParams gp{};
    fill_gram_params<T>(gp, &p->k, 1, 0, M, M, nullptr, nullptr);
    launch_gram<T>(Xst, Xst, m_pad, m_pad, p->D, Kss, m_pad, gp, s);
    launch_cov_finish<T>(C, m_pad, Kss, m_pad, s);
    cudaMemcpyKind kind = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CK(cudaMemcpy2DAsync(cov_out, (size_t)M * sizeof(T), C, (size_t)m_pad * sizeof(T), (size_t)M * sizeof(T), (size_t)M, kind, s));
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return AGP_OK;
}

// ---- logpdf / rand of a FiniteGP over a posterior: logpdf(f_post(x*, Sigma*), Y) and rand(f_post(x*, Sigma*), S)
// (/root/reference/src/finite_gp_projection.jl:306-311 and :233-237 applied to f = PosteriorGP, whose
// mean_and_cov is /root/reference/src/exact_gpr_posterior.jl:78-83).  The M x M posterior covariance never
// leaves the device: K** + Sigma* is generated straight into a factor buffer, V'V (V = L^-1 K(x, x*)) is
// subtracted by one GEMM, and the same in-place Cholesky as agp_fit runs with (Y - m*)' in the border tile.
template <typename T>
int post_cond_impl(agp_post* p, int layout, const void* Xs, int64_t M, const agp_mean* mean_s,
                   const agp_noise* noise_s, const void* Y, int S, void* logpdf_out, const void* Z, int Sz,
                   void* rand_out) {
  agp_ctx* ctx = p->ctx;
  { int rrc = post_replicate<T>(p); if (rrc) return rrc; }
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  if (M <= 0) { ctx->err = "M must be positive"; return AGP_ERR_DIM_MISMATCH; }
  if (!Xs) { ctx->err = "Xs is NULL"; return AGP_ERR_INVALID; }
  if (S < 0 || S > TILE) { ctx->err = "number of right-hand sides must be in [0,128]"; return AGP_ERR_UNSUPPORTED; }
  if (S > 0 && (!Y || !logpdf_out)) { ctx->err = "Y/logpdf_out is NULL"; return AGP_ERR_INVALID; }
  if (Sz > 0 && (!Z || !rand_out)) { ctx->err = "Z/out is NULL"; return AGP_ERR_INVALID; }
  static const agp_noise default_noise{0, 1e-18, nullptr};  // default_sigma^2, finite_gp_projection.jl:17
  if (!noise_s) noise_s = &default_noise;
  if (noise_s->kind == 1 && !noise_s->v) { ctx->err = "noise vector is NULL"; return AGP_ERR_INVALID; }
  agp_mean mz{p->mean_kind, p->mean_c, nullptr};
  if (!mean_s) mean_s = &mz;
  if (mean_s->kind == 2 && !mean_s->v) { ctx->err = "mean vector is NULL"; return AGP_ERR_INVALID; }

  const int64_t m_pad = round_up(M, TILE), ldf = m_pad + TILE;
  const int nblk = (int)(m_pad / TILE);
  Scratch sc(ctx);
  T *Xst = nullptr, *B = nullptr;
  int rc = post_cross<T>(p, sc, layout, Xs, M, m_pad, &Xst, &B);
  if (rc) return rc;
  T *mean_d = nullptr, *noise_d = nullptr, *Yd = nullptr;
  if (mean_s->kind == 2) { rc = upload<T>(ctx, sc, mean_s->v, M, true, &mean_d); if (rc) return rc; }
  if (noise_s->kind == 1) { rc = upload<T>(ctx, sc, noise_s->v, M, true, &noise_d); if (rc) return rc; }
  if (S > 0) { rc = upload<T>(ctx, sc, Y, (size_t)M * S, false, &Yd); if (rc) return rc; }
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)m_pad * sizeof(T)));
  T* mu = (T*)tmp;
  CK(cudaMemsetAsync(mu, 0, (size_t)m_pad * sizeof(T), s));
  // posterior mean m* = m(x*) + K(x*, x) alpha, then V = L^-1 K(x, x*) in place
  launch_gemv_t<T>(B, p->n_pad, p->n_pad, M, (const T*)p->alpha, mean_s->kind, mean_s->c, mean_d, mu, s);
  forward_subst_multi<T>(ctx, (const T*)p->L, p->lda, (const T*)p->Dinv, p->n_pad, B, p->n_pad, m_pad);
  // C* + Sigma* = K(x*, x*) + Sigma* - V'V, lower triangle, identity padding
  CK(sc.alloc(&tmp, (size_t)ldf * m_pad * sizeof(T)));
  T* Lf = (T*)tmp;
  CK(sc.alloc(&tmp, (size_t)nblk * TILE * TILE * sizeof(T)));
  T* Dinv = (T*)tmp;
  GramParams gp{};
  fill_gram_params<T>(gp, &p->k, 1, 1, M, M, noise_s, noise_d);
  launch_gram<T>(Xst, Xst, m_pad, m_pad, p->D, Lf, ldf, gp, s);
  {
    GemmArgs g{};
    g.A = B; g.lda = p->n_pad; g.a_kmajor = 1;
    g.B = B; g.ldb = p->n_pad; g.b_kmajor = 1;
    g.C = Lf; g.ldc = ldf; g.M = m_pad; g.N = m_pad; g.K = p->n_pad;
    g.alpha_neg = 1; g.beta_one = 1; g.lower_only = 1;
    launch_gemm<T>(g, s);
  }
  launch_border_init<T>(Lf, ldf, M, m_pad, Yd, M, S, 2, 0.0, (const T*)mu, s);  // border = (Y - m*)'
  CK(sc.alloc(&tmp, (size_t)(nblk + TILE + 2) * sizeof(double)));
  double* dscal = (double*)tmp;
  CK(sc.alloc(&tmp, sizeof(int)));
  int* dinfo = (int*)tmp;
  CK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  CK(sc.alloc(&tmp, (size_t)TILE * sizeof(T)));
  T* lp_d = (T*)tmp;
  prof_begin(ctx);
  cholesky_inplace<T>(ctx, Lf, ldf, m_pad, ldf, Dinv, dscal, dinfo);
  if (S > 0) {
    CK(sc.alloc(&tmp, (size_t)S * m_pad * sizeof(T)));
    launch_extract_v<T>(Lf, ldf, m_pad, S, (T*)tmp, dscal + nblk, s);
    launch_finalize_logpdf<T>(dscal, nblk, dscal + nblk, S, M, lp_d, dscal + nblk + TILE, s);
  }
  int h_info = 0;
  CK(cudaMemcpyAsync(&h_info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (h_info != 0) {
    ctx->info = h_info;
    char b[128];
    snprintf(b, sizeof(b), "posterior covariance is not positive definite; Cholesky failed at pivot %d", h_info);
    ctx->err = b;
    return AGP_ERR_NOT_POSDEF;
  }
  if (S > 0) CK(cudaMemcpyAsync(logpdf_out, lp_d, (size_t)S * sizeof(T), cudaMemcpyDeviceToHost, s));
  if (Sz > 0) {  // out = m* + L* Z  (C.U' * randn, finite_gp_projection.jl:235)
    const int64_t s_pad = round_up(Sz, 4);
    CK(sc.alloc(&tmp, (size_t)m_pad * s_pad * sizeof(T) * 2));
    T* Zd = (T*)tmp; T* Od = Zd + m_pad * s_pad;
    CK(cudaMemsetAsync(Zd, 0, (size_t)m_pad * s_pad * sizeof(T), s));
    cudaMemcpyKind kin = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    cudaMemcpyKind kout = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CK(cudaMemcpy2DAsync(Zd, (size_t)m_pad * sizeof(T), Z, (size_t)M * sizeof(T), (size_t)M * sizeof(T), (size_t)Sz, kin, s));
    GemmArgs g{};
    g.A = Lf; g.lda = ldf; g.a_kmajor = 0;
    g.B = Zd; g.ldb = m_pad; g.b_kmajor = 1;
    g.C = Od; g.ldc = m_pad; g.M = m_pad; g.N = s_pad; g.K = m_pad; g.trmm_lower = 1;
    launch_gemm<T>(g, s);
    launch_add_mean_cols<T>(Od, m_pad, M, Sz, 2, 0.0, (const T*)mu, s);
    CK(cudaMemcpy2DAsync(rand_out, (size_t)M * sizeof(T), Od, (size_t)m_pad * sizeof(T), (size_t)M * sizeof(T), (size_t)Sz, kout, s));
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return AGP_OK;
}

// ---- EXPERIMENTAL (compiles, not yet run on a device): gradient of logpdf(fx, y) w.r.t. the hyper-parameters from the
// factor agp_fit left in the handle.  V = L^-1 by the blocked forward
// substitution on the identity, C^-1 = V'V by the lower-only GEMM (both validated kernels), then grad.cu's fused
// reduction 1/2 sum (alpha alpha' - C^-1) o dC/dtheta.  Extra cost ~ 2 N^3 flop and two N^2 buffers.
template <typename T>
int post_logpdf_grad_impl(agp_post* p, double* grad_out, void* noise_diag_out) {
  agp_ctx* ctx = p->ctx;
  { int rrc = post_replicate<T>(p); if (rrc) return rrc; }
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  if (p->valid || p->segs.size() > 1) { ctx->err = "gradient of an extended (sequentially conditioned) posterior is unsupported"; return AGP_ERR_UNSUPPORTED; }
  const int64_t n = p->n, n_pad = p->n_pad;
  const int D = p->D;
  Scratch sc(ctx);
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)n_pad * n_pad * sizeof(T)));
  T* V = (T*)tmp;
  CK(sc.alloc(&tmp, (size_t)n_pad * n_pad * sizeof(T)));
  T* Cinv = (T*)tmp;
  CK(sc.alloc(&tmp, (size_t)(5 + D) * sizeof(double)));
  double* sums = (double*)tmp;
  T* noise_d = nullptr;
  if (noise_diag_out) { CK(sc.alloc(&tmp, (size_t)n_pad * sizeof(T))); noise_d = (T*)tmp; }
  CK(cudaMemsetAsync(V, 0, (size_t)n_pad * n_pad * sizeof(T), s));
  CK(cudaMemsetAsync(sums, 0, (size_t)(5 + D) * sizeof(double), s));
  launch_add_diag<T>(V, n_pad, n_pad, 1.0, s);
  forward_subst_multi<T>(ctx, (const T*)p->L, p->lda, (const T*)p->Dinv, n_pad, V, n_pad, n_pad);
  {
    GemmArgs g{};  // C^-1 = V'V, lower tiles
    g.A = V; g.lda = n_pad; g.a_kmajor = 1;
    g.B = V; g.ldb = n_pad; g.b_kmajor = 1;
    g.C = Cinv; g.ldc = n_pad; g.M = n_pad; g.N = n_pad; g.K = n_pad; g.lower_only = 1;
    launch_gemm<T>(g, s);
  }
  const int want_ard = (p->k.transform == AGP_T_ARD) ? 1 : 0;
  launch_grad_reduce<T>((const T*)p->Xt, D, n, n_pad, Cinv, n_pad, (const T*)p->alpha, p->k.family, p->k.linear_c, want_ard,
                        sums, noise_d, s);
  std::vector<double> h((size_t)5 + D);
  std::vector<T> ard_h((size_t)(D > 0 ? D : 1));
  CK(cudaMemcpyAsync(h.data(), sums, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (want_ard && p->ard) CK(cudaMemcpyAsync(ard_h.data(), p->ard, (size_t)D * sizeof(T), cudaMemcpyDeviceToHost, s));
  if (noise_diag_out) { int rc = download<T>(ctx, noise_diag_out, noise_d, (size_t)n, false); if (rc) return rc; }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  const bool linear = p->k.family == AGP_LINEAR;
  const double var = p->k.variance, sc_ = p->k.scale;
  grad_out[0] = 0.5 * h[0];
  grad_out[1] = (p->k.transform == AGP_T_SCALE) ? (linear ? var * h[1] / sc_ : 0.5 * var * h[1] / sc_) : 0.0;
  grad_out[2] = linear ? 0.5 * var * h[2] : 0.0;
  grad_out[3] = 0.5 * h[3];
  grad_out[4] = h[4];
  for (int d = 0; d < D; ++d)
    grad_out[5 + d] = want_ard ? (linear ? var : 0.5 * var) * h[(size_t)5 + d] / (double)ard_h[(size_t)d] : 0.0;
  return AGP_OK;
}

template <typename T>
int post_solve_lower_impl(agp_post* p, const void* Bh, int64_t nrhs, void* V_out) {
  agp_ctx* ctx = p->ctx;
  { int rrc = post_replicate<T>(p); if (rrc) return rrc; }
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  if (nrhs <= 0) return AGP_OK;
  const int64_t c_pad = round_up(nrhs, 4);
  Scratch sc(ctx);
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)p->n_pad * c_pad * sizeof(T)));
  T* B = (T*)tmp;
  CK(cudaMemsetAsync(B, 0, (size_t)p->n_pad * c_pad * sizeof(T), s));
  cudaMemcpyKind kin = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  cudaMemcpyKind kout = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  int64_t off = 0;
  for (auto& sg : p->segs) {  // compact rows -> padded row positions
    CK(cudaMemcpy2DAsync(B + sg.first, (size_t)p->n_pad * sizeof(T), (const T*)Bh + off, (size_t)p->n * sizeof(T),
                         (size_t)sg.second * sizeof(T), (size_t)nrhs, kin, s));
    off += sg.second;
  }
  forward_subst_multi<T>(ctx, (const T*)p->L, p->lda, (const T*)p->Dinv, p->n_pad, B, p->n_pad, c_pad);
  off = 0;
  for (auto& sg : p->segs) {
    CK(cudaMemcpy2DAsync((T*)V_out + off, (size_t)p->n * sizeof(T), B + sg.first, (size_t)p->n_pad * sizeof(T),
                         (size_t)sg.second * sizeof(T), (size_t)nrhs, kout, s));
    off += sg.second;
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return AGP_OK;
}

template <typename T>
int post_export_impl(agp_post* p, void* U_out) {
  agp_ctx* ctx = p->ctx;
  { int rrc = post_replicate<T>(p); if (rrc) return rrc; }
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  void* tmp = nullptr;
  const int64_t* map_d = nullptr;
  if (p->segs.size() > 1) {  // extended posterior: valid rows are not contiguous
    std::vector<int64_t> map;
    for (auto& sg : p->segs) for (int64_t i = 0; i < sg.second; ++i) map.push_back(sg.first + i);
    CK(sc.alloc(&tmp, map.size() * sizeof(int64_t)));
    CK(cudaMemcpyAsync(tmp, map.data(), map.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    map_d = (const int64_t*)tmp;
  }
  T* Ud = (T*)U_out;
  if (ctx->memspace != AGP_MEM_DEVICE) {
    CK(sc.alloc(&tmp, (size_t)p->n * p->n * sizeof(T)));
    Ud = (T*)tmp;
  }
  if (map_d) launch_export_upper_map<T>((const T*)p->L, p->lda, p->n, map_d, Ud, p->n, s);
  else launch_export_upper<T>((const T*)p->L, p->lda, p->n, Ud, p->n, s);
  if (ctx->memspace != AGP_MEM_DEVICE)
    CK(cudaMemcpyAsync(U_out, Ud, (size_t)p->n * p->n * sizeof(T), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return AGP_OK;
}

// ---- sequential conditioning: posterior(fx::FiniteGP{<:PosteriorGP}, y) via update_chol
// (/root/reference/src/exact_gpr_posterior.jl:46-56, /root/reference/src/util/common_covmat_ops.jl:38-42).
// L21 = C21 L11^-T by the same panel solves the factorisation uses (one per old column block), then
// ONE long-K trailing update C22 -= L21 L21', then a Cholesky of the new diagonal part.  The border
// rows carry [delta1; delta2]' through all of it, so v = L^-1 delta comes out of the same kernels.
template <typename T>
int post_extend_impl(agp_post* p, int layout, const void* X2, int64_t N2, const void* y2, const agp_mean* mean2,
                     const agp_noise* noise2, void* alpha_out, agp_post** post_out) {
  agp_ctx* ctx = p->ctx;
  { int rrc = post_replicate<T>(p); if (rrc) return rrc; }
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  if (N2 <= 0) { ctx->err = "N2 must be positive"; return AGP_ERR_DIM_MISMATCH; }
  if (!X2 || !y2) { ctx->err = "X2/y2 is NULL"; return AGP_ERR_INVALID; }
  static const agp_mean zero_mean{0, 0.0, nullptr};
  static const agp_noise default_noise{0, 1e-18, nullptr};
  if (!mean2) mean2 = &zero_mean;
  if (!noise2) noise2 = &default_noise;
  Scratch sc(ctx);
  const int64_t n1p = p->n_pad, n2p = round_up(N2, TILE), np = n1p + n2p, ldn = np + TILE;
  const int nblk1 = (int)(n1p / TILE), nblk2 = (int)(n2p / TILE), nblk = nblk1 + nblk2;
  const int D = p->D;
  T *mean_d = nullptr, *noise_d = nullptr, *y2d = nullptr, *X2t = nullptr;
  int rc;
  if (mean2->kind == 2) { rc = upload<T>(ctx, sc, mean2->v, N2, true, &mean_d); if (rc) return rc; }
  if (noise2->kind == 1) { rc = upload<T>(ctx, sc, noise2->v, N2, true, &noise_d); if (rc) return rc; }
  rc = upload<T>(ctx, sc, y2, N2, false, &y2d); if (rc) return rc;
  rc = prep_points<T>(ctx, sc, &p->k, (const T*)p->ard, layout, X2, N2, n2p, D, &X2t, false); if (rc) return rc;

  void *Ln = nullptr, *Dn = nullptr, *Xn = nullptr, *an = nullptr, *dn = nullptr, *vn = nullptr;
  CK(cudaMallocAsync(&Ln, (size_t)ldn * np * sizeof(T), s));
  CK(cudaMallocAsync(&Dn, (size_t)nblk * TILE * TILE * sizeof(T), s));
  CK(cudaMallocAsync(&Xn, (size_t)np * D * sizeof(T), s));
  CK(cudaMallocAsync(&an, (size_t)np * sizeof(T), s));
  CK(cudaMallocAsync(&dn, (size_t)np * sizeof(T), s));
  CK(cudaMallocAsync(&vn, (size_t)np, s));
  T* L = (T*)Ln; T* Dinv = (T*)Dn;
  unsigned char* valid = (unsigned char*)vn;
  // carry the old state over
  launch_copy2d<T>((const T*)p->L, p->lda, L, ldn, n1p, n1p, s);
  CK(cudaMemcpyAsync(Dinv, p->Dinv, (size_t)nblk1 * TILE * TILE * sizeof(T), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(Xn, p->Xt, (size_t)n1p * D * sizeof(T), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync((T*)Xn + n1p * D, X2t, (size_t)n2p * D * sizeof(T), cudaMemcpyDeviceToDevice, s));
  if (p->valid) CK(cudaMemcpyAsync(valid, p->valid, (size_t)n1p, cudaMemcpyDeviceToDevice, s));
  else { CK(cudaMemsetAsync(valid, 1, (size_t)p->n, s)); CK(cudaMemsetAsync(valid + p->n, 0, (size_t)(n1p - p->n), s)); }
  CK(cudaMemsetAsync(valid + n1p, 1, (size_t)N2, s));
  CK(cudaMemsetAsync(valid + n1p + N2, 0, (size_t)(n2p - N2), s));
  CK(cudaMemcpyAsync(dn, p->delta, (size_t)n1p * sizeof(T), cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync((T*)dn + n1p, 0, (size_t)n2p * sizeof(T), s));
  launch_sub_mean<T>(y2d, N2, mean2->kind, mean2->c, mean_d, (T*)dn + n1p, s);
  // C21 = K(x2, x1) and C22 = K(x2, x2) + Sigma_y2 straight into the new factor buffer
  GramParams g21{};
  fill_gram_params<T>(g21, &p->k, 0, 0, N2, p->n, nullptr, nullptr);
  g21.mask_b = valid;  // old rows (first n1p entries of the new mask)
  launch_gram<T>(X2t, (const T*)Xn, n2p, n1p, D, L + n1p, ldn, g21, s);
  GramParams g22{};
  fill_gram_params<T>(g22, &p->k, 1, 1, N2, N2, noise2, noise_d);
  launch_gram<T>(X2t, X2t, n2p, n2p, D, L + n1p + n1p * ldn, ldn, g22, s);
  launch_border_init<T>(L, ldn, np, np, (const T*)dn, np, 1, 0, 0.0, (const T*)nullptr, s);
  // panel solves of the new rows (+ border) against the old factor
  const int64_t Mr = n2p + TILE;
  for (int k = 0; k < nblk1; ++k) {
    T* Rk = L + n1p + (int64_t)k * TILE * ldn;
    GemmArgs t{};
    t.A = Rk; t.lda = ldn; t.B = Dinv + (int64_t)k * TILE * TILE; t.ldb = TILE;
    t.C = Rk; t.ldc = ldn; t.M = Mr; t.N = TILE; t.K = TILE;
    launch_gemm<T>(t, s);
    const int64_t cols_rest = n1p - (int64_t)(k + 1) * TILE;
    if (cols_rest <= 0) continue;
    GemmArgs u{};
    u.A = Rk; u.lda = ldn;
    u.B = L + (int64_t)(k + 1) * TILE + (int64_t)k * TILE * ldn; u.ldb = ldn;
    u.C = Rk + (int64_t)TILE * ldn; u.ldc = ldn; u.M = Mr; u.N = cols_rest; u.K = TILE; u.alpha_neg = 1; u.beta_one = 1;
    launch_gemm<T>(u, s);
  }
  {  // C22 -= L21 L21'  (one launch, K = n1p)
    GemmArgs u{};
    u.A = L + n1p; u.lda = ldn; u.B = L + n1p; u.ldb = ldn;
    u.C = L + n1p + n1p * ldn; u.ldc = ldn; u.M = Mr; u.N = n2p; u.K = n1p; u.alpha_neg = 1; u.beta_one = 1; u.lower_only = 1;
    launch_gemm<T>(u, s);
  }
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)(nblk2 + TILE + 2) * sizeof(double)));
  double* dscal = (double*)tmp;
  CK(sc.alloc(&tmp, sizeof(int)));
  int* dinfo = (int*)tmp;
  CK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  CK(sc.alloc(&tmp, (size_t)(nblk + 1) * sizeof(int)));
  int* dflags = (int*)tmp;
  CK(sc.alloc(&tmp, (size_t)np * sizeof(T)));
  T* rwork = (T*)tmp;
  CK(sc.alloc(&tmp, (size_t)TILE * sizeof(T)));
  T* lp_d = (T*)tmp;
  prof_begin(ctx);
  cholesky_inplace<T>(ctx, L + n1p + n1p * ldn, ldn, n2p, n2p + TILE, Dinv + (int64_t)nblk1 * TILE * TILE, dscal, dinfo);
  launch_extract_v<T>(L, ldn, np, 1, rwork, dscal + nblk2, s);
  launch_bwd_solve<T>(L, ldn, Dinv, nblk, rwork, dflags, s);
  CK(cudaMemcpyAsync(an, rwork, (size_t)np * sizeof(T), cudaMemcpyDeviceToDevice, s));
  launch_finalize_logpdf<T>(dscal, nblk2, dscal + nblk2, 1, N2, lp_d, dscal + nblk2 + TILE, s);
  int h_info = 0;
  double h_ld = 0.0;
  CK(cudaMemcpyAsync(&h_info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h_ld, dscal + nblk2 + TILE, sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (h_info != 0) {
    cudaFreeAsync(Ln, s); cudaFreeAsync(Dn, s); cudaFreeAsync(Xn, s); cudaFreeAsync(an, s); cudaFreeAsync(dn, s); cudaFreeAsync(vn, s);
    ctx->info = p->n + h_info;
    ctx->err = "extended covariance is not positive definite";
    return AGP_ERR_NOT_POSDEF;
  }
  if (post_out) {  // the reference's semantics: a NEW posterior, the conditioned-on one stays valid
    agp_post* q = new agp_post(*p);
    q->ard = nullptr;
    if (p->ard) {
      cudaMallocAsync(&q->ard, (size_t)D * sizeof(T), s);
      cudaMemcpyAsync(q->ard, p->ard, (size_t)D * sizeof(T), cudaMemcpyDeviceToDevice, s);
    }
    *post_out = q;
    p = q;
  } else {  // in place: the old state is released
    cudaFreeAsync(p->L, s); cudaFreeAsync(p->Dinv, s); cudaFreeAsync(p->Xt, s); cudaFreeAsync(p->alpha, s);
    cudaFreeAsync(p->delta, s);
    if (p->valid) cudaFreeAsync(p->valid, s);
  }
  p->L = Ln; p->Dinv = Dn; p->Xt = Xn; p->alpha = an; p->delta = dn; p->valid = valid;
  p->segs.push_back({n1p, N2});
  p->n += N2; p->n_pad = np; p->lda = ldn; p->logdet += h_ld;
  if (alpha_out) {
    int64_t off = 0;
    cudaMemcpyKind kout = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    for (auto& sg : p->segs) {
      CK(cudaMemcpyAsync((T*)alpha_out + off, (const T*)p->alpha + sg.first, (size_t)sg.second * sizeof(T), kout, s));
      off += sg.second;
    }
    CK(cudaStreamSynchronize(s));
  }
  return AGP_OK;
}

template <typename T>
int rand_impl(agp_ctx* ctx, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise, int layout,
              const void* X, int64_t N, int D, const void* Z, int S, void* out) {
  if (S <= 0) return AGP_OK;
  if (!Z || !out) { ctx->err = "Z/out is NULL"; return AGP_ERR_INVALID; }
  Scratch outer(ctx);
  T* L = nullptr;
  int64_t lda = 0;
  int rc = fit_impl<T>(ctx, k, mean, noise, layout, X, N, D, nullptr, 0, nullptr, nullptr, nullptr, &L, &lda, &outer);
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  const int64_t n_pad = round_up(N, TILE), s_pad = round_up(S, 4);
  void* tmp = nullptr;
  CK(outer.alloc(&tmp, (size_t)n_pad * s_pad * sizeof(T) * 2));
  T* Zd = (T*)tmp; T* Od = Zd + n_pad * s_pad;
  CK(cudaMemsetAsync(Zd, 0, (size_t)n_pad * s_pad * sizeof(T), s));
  cudaMemcpyKind kin = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CK(cudaMemcpy2DAsync(Zd, (size_t)n_pad * sizeof(T), Z, (size_t)N * sizeof(T), (size_t)N * sizeof(T), (size_t)S, kin, s));
  GemmArgs g{};  // out = L * Z (TRMM, lower)
  g.A = L; g.lda = lda; g.a_kmajor = 0;
  g.B = Zd; g.ldb = n_pad; g.b_kmajor = 1;
  g.C = Od; g.ldc = n_pad; g.M = n_pad; g.N = s_pad; g.K = n_pad; g.trmm_lower = 1;
  launch_gemm<T>(g, s);
  static const agp_mean zero_mean{0, 0.0, nullptr};
  if (!mean) mean = &zero_mean;
  T* mean_d = nullptr;
  if (mean->kind == 2) { rc = upload<T>(ctx, outer, mean->v, N, true, &mean_d); if (rc) return rc; }
  launch_add_mean_cols<T>(Od, n_pad, N, S, mean->kind, mean->c, mean_d, s);
  cudaMemcpyKind kout = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CK(cudaMemcpy2DAsync(out, (size_t)N * sizeof(T), Od, (size_t)n_pad * sizeof(T), (size_t)N * sizeof(T), (size_t)S, kout, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return AGP_OK;
}

template <typename T>
int gram_impl(agp_ctx* ctx, const agp_kernel* k, int layout, const void* X, int64_t N, int D, const void* Z,
              int64_t M, const agp_noise* noise, void* K_out) {
  int rc = check_kernel(ctx, k, D);
  if (rc) return rc;
  if (N <= 0 || (Z && M <= 0)) { ctx->err = "empty input"; return AGP_ERR_DIM_MISMATCH; }
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  const int64_t n_pad = round_up(N, 64);
  const int64_t cols = Z ? M : N, c_pad = round_up(cols, 64);
  T *ard_d = nullptr, *noise_d = nullptr, *Xt = nullptr, *Zt = nullptr;
  if (k->transform == AGP_T_ARD) { rc = upload<T>(ctx, sc, k->ard, D, true, &ard_d); if (rc) return rc; }
  if (noise && noise->kind == 1) { rc = upload<T>(ctx, sc, noise->v, N, true, &noise_d); if (rc) return rc; }
  rc = prep_points<T>(ctx, sc, k, ard_d, layout, X, N, n_pad, D, &Xt, false);
  if (rc) return rc;
  if (Z) { rc = prep_points<T>(ctx, sc, k, ard_d, layout, Z, M, c_pad, D, &Zt, false); if (rc) return rc; }
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)n_pad * c_pad * sizeof(T)));
  GramParams gp{};
  fill_gram_params<T>(gp, k, Z ? 0 : 1, 0, N, cols, Z ? nullptr : noise, noise_d);
  launch_gram<T>(Xt, Z ? Zt : Xt, n_pad, c_pad, D, (T*)tmp, n_pad, gp, s);
  cudaMemcpyKind kout = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CK(cudaMemcpy2DAsync(K_out, (size_t)N * sizeof(T), tmp, (size_t)n_pad * sizeof(T), (size_t)N * sizeof(T), (size_t)cols, kout, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return AGP_OK;
}




// ---- VFE (Titsias) : elbo / dtc / approximate posterior --------------------------------------------
// Follows /root/reference/src/sparse_approximations.jl:289-305 (_compute_intermediates), :248-254 (elbo),
// :58-75 (posterior), but STREAMS the data dimension: K_zx is generated chunk by chunk, scaled by
// Sigma_y^-1/2, solved against chol(K_zz) and folded into D = A A' (M x M), b = A delta and ||A||_F^2,
// so the M x N matrix A (32.8 GB at config C5) never exists.
template <typename T>
int vfe_core(agp_ctx* ctx, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise, int layout,
             const void* X, int64_t N, int D, const void* Zind, int64_t M, const agp_noise* jitter, const void* y,
             void* elbo_out, void* dtc_out, agp_vfe_post** post_out) {
  int rc = check_kernel(ctx, k, D);
  if (rc) return rc;
  if (N <= 0 || M <= 0) { ctx->err = "N and M must be positive"; return AGP_ERR_DIM_MISMATCH; }
  if (!X || !Zind || !y) { ctx->err = "X/Z/y is NULL"; return AGP_ERR_INVALID; }
  static const agp_mean zero_mean{0, 0.0, nullptr};
  static const agp_noise default_noise{0, 1e-18, nullptr};
  if (!mean) mean = &zero_mean;
  if (!noise) noise = &default_noise;
  if (!jitter) jitter = &default_noise;
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  const bool keep = post_out != nullptr;
  // + TILE: a rank's shard start is not tile-aligned, and the chunk Gram reads whole 128-point slabs from it
  const int64_t m_pad = round_up(M, TILE), lda = m_pad + TILE, n_padN = round_up(N, TILE) + TILE;
  const int nblk = (int)(m_pad / TILE);
  CK(cudaEventRecord(ctx->ev[0], s));
  T *ard_d = nullptr, *mean_d = nullptr, *noise_d = nullptr, *jit_d = nullptr, *yd = nullptr, *Zt = nullptr, *Xt = nullptr;
  if (k->transform == AGP_T_ARD) { rc = upload<T>(ctx, sc, k->ard, D, true, &ard_d); if (rc) return rc; }
  if (mean->kind == 2) { rc = upload<T>(ctx, sc, mean->v, N, true, &mean_d); if (rc) return rc; }
  if (noise->kind == 1) { rc = upload<T>(ctx, sc, noise->v, N, true, &noise_d); if (rc) return rc; }
  if (jitter->kind == 1) { rc = upload<T>(ctx, sc, jitter->v, M, true, &jit_d); if (rc) return rc; }
  rc = upload<T>(ctx, sc, y, N, false, &yd); if (rc) return rc;
  rc = prep_points<T>(ctx, sc, k, ard_d, layout, Zind, M, m_pad, D, &Zt, keep); if (rc) return rc;
  rc = prep_points<T>(ctx, sc, k, ard_d, layout, X, N, n_padN, D, &Xt, false); if (rc) return rc;
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)n_padN * 3 * sizeof(T)));
  T* kd = (T*)tmp; T* delta = kd + n_padN; T* isn = delta + n_padN;
  CK(cudaMemsetAsync(kd, 0, (size_t)n_padN * 3 * sizeof(T), s));
  CK(sc.alloc(&tmp, (size_t)(2 * nblk + TILE + 16) * sizeof(double)));
  double* dscal = (double*)tmp;  // [0..3] prep scalars + ||A||^2, [8..8+nblk) logdet Kzz, [8+nblk..) logdet Lam, then sq
  CK(cudaMemsetAsync(dscal, 0, (size_t)(2 * nblk + TILE + 16) * sizeof(double), s));
  double* ld_z = dscal + 8; double* ld_l = ld_z + nblk; double* sq = ld_l + nblk;
  CK(sc.alloc(&tmp, sizeof(int)));
  int* dinfo = (int*)tmp;
  CK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  // data shard of this rank (multi-GPU: the N dimension is partitioned, ONE all-reduce at the end -- SURVEY s8e)
  int64_t n_lo = 0, n_hi = N;
  if (ctx->nccl) {
    const int64_t per = (N + ctx->nranks - 1) / ctx->nranks;
    n_lo = (int64_t)ctx->rank * per; if (n_lo > N) n_lo = N;
    n_hi = n_lo + per; if (n_hi > N) n_hi = N;
  }
  launch_kdiag<T>(Xt, N, D, k->family, k->variance, k->linear_c, kd, s);
  launch_vfe_prep<T>(yd + n_lo, n_hi - n_lo, mean->kind, mean->c, mean_d ? mean_d + n_lo : nullptr, noise->kind, noise->s,
                     noise_d ? noise_d + n_lo : nullptr, kd + n_lo, delta + n_lo, isn + n_lo, dscal, s);

  // factor buffers
  void *Lzv = nullptr, *Dzv = nullptr, *Lmv = nullptr, *Dlv = nullptr, *mev = nullptr;
  CK(cudaMallocAsync(&Lzv, (size_t)lda * m_pad * sizeof(T), s));
  CK(cudaMallocAsync(&Dzv, (size_t)nblk * TILE * TILE * sizeof(T), s));
  CK(cudaMallocAsync(&Lmv, (size_t)lda * m_pad * sizeof(T), s));
  CK(cudaMallocAsync(&Dlv, (size_t)nblk * TILE * TILE * sizeof(T), s));
  CK(cudaMallocAsync(&mev, (size_t)m_pad * 2 * sizeof(T), s));
  if (!keep) { sc.ptrs.push_back(Lzv); sc.ptrs.push_back(Dzv); sc.ptrs.push_back(Lmv); sc.ptrs.push_back(Dlv); sc.ptrs.push_back(mev); }
  T* Lz = (T*)Lzv; T* Dz = (T*)Dzv; T* Lm = (T*)Lmv; T* Dl = (T*)Dlv; T* bvec = (T*)mev; T* rwork = bvec + m_pad;
  agp_vfe_post* vp = nullptr;
  if (keep) {
    vp = new agp_vfe_post();
    vp->ctx = ctx; vp->dtype = sizeof(T) == 8 ? AGP_F64 : AGP_F32; vp->m = M; vp->m_pad = m_pad; vp->lda = lda; vp->D = D;
    vp->U = Lzv; vp->Udinv = Dzv; vp->Lam = Lmv; vp->Ldinv = Dlv; vp->Zt = Zt; vp->m_e = mev;
    vp->k = *k; vp->k.ard = nullptr;
    if (ard_d) { sc.release(ard_d); vp->ard = ard_d; }
    vp->mean_kind = mean->kind == 2 ? 0 : mean->kind; vp->mean_c = mean->c;
  }
  // every early return (CK / CKN failures, non-PD) releases the half-built handle and its buffers
  struct VfeGuard { agp_vfe_post* p; ~VfeGuard() { if (p) agp_vfe_post_free(p); } } vguard{vp};
  auto fail = [&](int code) { return code; };

  // (1) chol(K_zz + jitter)
  GramParams gz{};
  fill_gram_params<T>(gz, k, 1, 1, M, M, jitter, jit_d);
  launch_gram<T>(Zt, Zt, m_pad, m_pad, D, Lz, lda, gz, s);
  launch_border_init<T>(Lz, lda, m_pad, m_pad, (const T*)nullptr, m_pad, 0, 0, 0.0, (const T*)nullptr, s);
  prof_begin(ctx);
  cholesky_inplace<T>(ctx, Lz, lda, m_pad, lda, Dz, ld_z, dinfo);
  CK(cudaEventRecord(ctx->ev[1], s));

  // (2) stream the data: D += A_c A_c', b += A_c delta_c, ||A||_F^2
  CK(cudaMemsetAsync(Lm, 0, (size_t)lda * m_pad * sizeof(T), s));
  CK(cudaMemsetAsync(bvec, 0, (size_t)m_pad * 2 * sizeof(T), s));
  int64_t cap = (int64_t)(2.0e9 / ((double)m_pad * sizeof(T)));
  cap = cap / TILE * TILE;
  if (cap < TILE) cap = TILE;
  if (cap > n_padN) cap = n_padN;
  // D += A_c A_c' on the tensor cores: the chunk is ONE long-K product (K = chunk width); the int32 accumulators of the
  // sliced kernel hold (d+1) * K * 64^2 < 2^31, so K <= 32768 keeps every slice count exact
  const bool syrk_tc = resolve_tensor_mode<T>(ctx, (int64_t)1 << 20) == 1 && m_pad >= 1024;
  if (syrk_tc && cap > 32768) cap = 32768;
  constexpr int is_f32 = std::is_same<T, double>::value ? 0 : 1;
  void* Bv = nullptr;
  CK(sc.alloc(&Bv, (size_t)m_pad * cap * sizeof(T)));
  T* B = (T*)Bv;
  for (int64_t c0 = n_lo; c0 < n_hi; c0 += cap) {
    const int64_t nc = (n_hi - c0 < cap) ? (n_hi - c0) : cap;
    const int64_t nc_pad = round_up(nc, TILE);
    GramParams gx{};
    fill_gram_params<T>(gx, k, 0, 0, M, nc, nullptr, nullptr);
    launch_gram<T>(Zt, Xt + c0 * D, m_pad, nc_pad, D, B, m_pad, gx, s);
    launch_scale_cols<T>(B, m_pad, m_pad, nc, isn + c0, s);
    forward_subst_multi<T>(ctx, Lz, lda, Dz, m_pad, B, m_pad, nc_pad);
    bool acc_done = false;
    if (syrk_tc && nc_pad >= 1024 && ensure_oz2(ctx, m_pad, (int)nc_pad, slices_of<T>(ctx), s) && ctx->oz2.bulk == 2) {
      ozaki_prepare_ex(ctx->oz2, B, is_f32, 0, m_pad, m_pad, 0, s);
      acc_done = ozaki_update_ex(ctx->oz2, Lm, is_f32, lda, m_pad, m_pad, 0, 1.0, 0, 0, 0, 0, s) == 0;
    }
    if (!acc_done) {
      GemmArgs g{};
      g.A = B; g.lda = m_pad; g.B = B; g.ldb = m_pad; g.C = Lm; g.ldc = lda;
      g.M = m_pad; g.N = m_pad; g.K = nc_pad; g.beta_one = 1; g.lower_only = 1;
      launch_gemm<T>(g, s);
    }
    launch_gemv_n_acc<T>(B, m_pad, m_pad, nc, delta + c0, bvec, s);
    launch_sumsq<T>(B, m_pad * nc_pad, dscal + 3, s);
  }
  if (ctx->nccl) {  // the one exchange step: D (M x M), b (M) and the three scalars
    CKN(ncclAllReduce(Lm, Lm, (size_t)lda * m_pad, NcclType<T>::v, ncclSum, ctx->nccl, s));
    CKN(ncclAllReduce(bvec, bvec, (size_t)m_pad, NcclType<T>::v, ncclSum, ctx->nccl, s));
    CKN(ncclAllReduce(dscal, dscal, 4, ncclDouble, ncclSum, ctx->nccl, s));
  }
  CK(cudaEventRecord(ctx->ev[2], s));

  // (3) Lambda = chol(D + I), with b riding in the border row
  launch_add_diag<T>(Lm, lda, m_pad, 1.0, s);
  launch_border_init<T>(Lm, lda, m_pad, m_pad, bvec, m_pad, 1, 0, 0.0, (const T*)nullptr, s);
  cholesky_inplace<T>(ctx, Lm, lda, m_pad, lda, Dl, ld_l, dinfo);
  launch_extract_v<T>(Lm, lda, m_pad, 1, rwork, sq, s);
  if (keep) {
    CK(sc.alloc(&tmp, (size_t)(nblk + 1) * sizeof(int)));
    int* dflags = (int*)tmp;
    launch_bwd_solve<T>(Lm, lda, Dl, nblk, rwork, dflags, s);       // m_e = Lambda^-1 b
    CK(cudaMemcpyAsync(bvec, rwork, (size_t)m_pad * sizeof(T), cudaMemcpyDeviceToDevice, s));  // m_e kept in slot 0
  }
  CK(cudaEventRecord(ctx->ev[3], s));
  std::vector<double> h((size_t)(2 * nblk + 16 + 1));
  int h_info = 0;
  CK(cudaMemcpyAsync(h.data(), dscal, (size_t)(2 * nblk + 16 + 1) * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&h_info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]); ctx->timings[0] = ms;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->timings[3] = ms;
  cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->timings[6] = ms;
  ctx->timings[7] = ctx->profile ? prof_total_ms(ctx) : 0.0;
  if (h_info != 0) {
    ctx->info = h_info;
    ctx->err = "VFE: K_zz + jitter or A A' + I is not positive definite";
    return fail(AGP_ERR_NOT_POSDEF);
  }
  double logdet_lam = 0.0;
  for (int b = 0; b < nblk; ++b) logdet_lam += h[8 + nblk + b];
  logdet_lam *= 2.0;
  const double log2pi = 1.8378770664093454835606594728112;
  const double sqv = h[8 + 2 * nblk];
  const double dtc = -0.5 * ((double)N * log2pi + h[0] + logdet_lam + h[1] - sqv);
  const double elbo = dtc - 0.5 * (h[2] - h[3]);
  T e = (T)elbo, dd = (T)dtc;
  if (elbo_out) memcpy(elbo_out, &e, sizeof(T));
  if (dtc_out) memcpy(dtc_out, &dd, sizeof(T));
  vguard.p = nullptr;
  if (keep) *post_out = vp;
  return AGP_OK;
}

template <typename T>
int vfe_mean_var_impl(agp_vfe_post* p, int layout, const void* Xs, int64_t Ms, void* mean_out, void* var_out) {
  agp_ctx* ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  if (Ms <= 0) return AGP_OK;
  int64_t cap = (int64_t)(2.0e9 / ((double)p->m_pad * sizeof(T)));
  { const int64_t c_env = env_int64("AGP_PREDICT_CHUNK", 0); if (c_env > 0) cap = c_env; }
  cap = cap / TILE * TILE;
  if (cap < TILE) cap = TILE;
  for (int64_t c0 = 0; c0 < Ms; c0 += cap) {
    const int64_t mc = (Ms - c0 < cap) ? (Ms - c0) : cap;
    const int64_t c_pad = round_up(mc, TILE);
    Scratch sc(ctx);
    const void* xs_chunk = (layout == AGP_POINT_MAJOR) ? (const void*)((const char*)Xs + (size_t)c0 * p->D * sizeof(T)) : Xs;
    const int saved_memspace = ctx->memspace;
    if (layout != AGP_POINT_MAJOR && !(c0 == 0 && mc == Ms)) {  // RowVecs test set larger than one chunk
      T* g = nullptr;
      int grc = gather_feature_major_chunk<T>(ctx, sc, Xs, Ms, p->D, c0, mc, &g);
      if (grc) return grc;
      xs_chunk = g;
      ctx->memspace = AGP_MEM_DEVICE;
    }
    T* Xst = nullptr;
    int rc = prep_points<T>(ctx, sc, &p->k, (const T*)p->ard, layout, xs_chunk, mc, c_pad, p->D, &Xst, false);
    ctx->memspace = saved_memspace;
    if (rc) return rc;
    void* tmp = nullptr;
    CK(sc.alloc(&tmp, (size_t)p->m_pad * c_pad * sizeof(T)));
    T* B = (T*)tmp;
    CK(sc.alloc(&tmp, (size_t)c_pad * 2 * sizeof(T)));
    T* mu = (T*)tmp; T* var = mu + c_pad;
    GramParams gp{};
    fill_gram_params<T>(gp, &p->k, 0, 0, p->m, mc, nullptr, nullptr);
    launch_gram<T>((const T*)p->Zt, Xst, p->m_pad, c_pad, p->D, B, p->m_pad, gp, s);
    forward_subst_multi<T>(ctx, (const T*)p->U, p->lda, (const T*)p->Udinv, p->m_pad, B, p->m_pad, c_pad);  // A*
    launch_gemv_t<T>(B, p->m_pad, p->m_pad, mc, (const T*)p->m_e, p->mean_kind, p->mean_c, (const T*)nullptr, mu, s);
    launch_kdiag<T>(Xst, mc, p->D, p->k.family, p->k.variance, p->k.linear_c, var, s);
    launch_colsumsq_acc<T>(B, p->m_pad, p->m_pad, mc, -1.0, var, s);
    forward_subst_multi<T>(ctx, (const T*)p->Lam, p->lda, (const T*)p->Ldinv, p->m_pad, B, p->m_pad, c_pad);
    launch_colsumsq_acc<T>(B, p->m_pad, p->m_pad, mc, 1.0, var, s);
    if (mean_out) { rc = download<T>(ctx, (T*)mean_out + c0, mu, mc, false); if (rc) return rc; }
    if (var_out) { rc = download<T>(ctx, (T*)var_out + c0, var, mc, false); if (rc) return rc; }
    CK(cudaStreamSynchronize(s));
  }
  CK(cudaGetLastError());
  return AGP_OK;
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU fit: one process per GPU, block-column-cyclic tiles on a 1 x Q process grid, NCCL panel
// broadcast over NVLink/NVSwitch, look-ahead so the broadcast of panel k+1 overlaps the bulk of the
// trailing update of panel k.  Column block j (128 columns, all rows + the border rows) lives on
// rank j mod Q as local block j div Q.  Every rank generates only its own Gram columns from the
// replicated points.  (grid_p > 1 is declared in the ABI but not built: on NVSwitch the panel
// broadcast is ~5 % of the factorisation at C4, so the 2-D row/column split buys nothing yet.)
// ------------------------------------------------------------------------------------------------
// ---- EXPERIMENTAL (composed of validated launches, not yet run on a device): full predictive covariance of the
// approximate posterior, and logpdf / rand of a FiniteGP over it.
//   mean_and_cov(::ApproxPosteriorGP, x*)  /root/reference/src/sparse_approximations.jl:205-210 (cov :187-190):
//   A = U' \ K(z, x*),  m* = m(x*) + A' m_e,  C* = K** - A'A + (Lam' \ A)'(Lam' \ A)
// mode "cov": C* (no noise) is returned.  mode "factor": K** + Sigma* is generated with the noise fused, C* + Sigma* is
// factored in place with (Y - m*)' in the border tile (logpdf, src/finite_gp_projection.jl:306-318) and / or multiplied
// into the caller's normals (rand, :233-240) -- the same tail as post_cond_impl.
template <typename T>
int vfe_cond_impl(agp_vfe_post* p, int layout, const void* Xs, int64_t M, const agp_noise* noise_s, void* mean_out,
                  void* cov_out, const void* Y, int S, void* logpdf_out, const void* Z, int Sz, void* rand_out) {
  agp_ctx* ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  if (M <= 0) { ctx->err = "M must be positive"; return AGP_ERR_DIM_MISMATCH; }
  if (!Xs) { ctx->err = "Xs is NULL"; return AGP_ERR_INVALID; }
  if (S < 0 || S > TILE) { ctx->err = "number of right-hand sides must be in [0,128]"; return AGP_ERR_UNSUPPORTED; }
  if (S > 0 && (!Y || !logpdf_out)) { ctx->err = "Y/logpdf_out is NULL"; return AGP_ERR_INVALID; }
  if (Sz > 0 && (!Z || !rand_out)) { ctx->err = "Z/out is NULL"; return AGP_ERR_INVALID; }
  const bool factor = (S > 0 || Sz > 0);
  static const agp_noise default_noise{0, 1e-18, nullptr};
  if (factor && !noise_s) noise_s = &default_noise;
  if (factor && noise_s->kind == 1 && !noise_s->v) { ctx->err = "noise vector is NULL"; return AGP_ERR_INVALID; }
  const int64_t c_pad = round_up(M, TILE), ldf = c_pad + TILE;
  const int nblk = (int)(c_pad / TILE);
  Scratch sc(ctx);
  T* Xst = nullptr;
  int rc = prep_points<T>(ctx, sc, &p->k, (const T*)p->ard, layout, Xs, M, c_pad, p->D, &Xst, false);
  if (rc) return rc;
  T *noise_d = nullptr, *Yd = nullptr;
  if (factor && noise_s->kind == 1) { rc = upload<T>(ctx, sc, noise_s->v, M, true, &noise_d); if (rc) return rc; }
  if (S > 0) { rc = upload<T>(ctx, sc, Y, (size_t)M * S, false, &Yd); if (rc) return rc; }
  void* tmp = nullptr;
  CK(sc.alloc(&tmp, (size_t)p->m_pad * c_pad * sizeof(T)));
  T* B = (T*)tmp;
  CK(sc.alloc(&tmp, (size_t)c_pad * sizeof(T)));
  T* mu = (T*)tmp;
  CK(cudaMemsetAsync(mu, 0, (size_t)c_pad * sizeof(T), s));
  CK(sc.alloc(&tmp, (size_t)ldf * c_pad * sizeof(T)));
  T* Lf = (T*)tmp;
  GramParams gp{};
  fill_gram_params<T>(gp, &p->k, 0, 0, p->m, M, nullptr, nullptr);
  launch_gram<T>((const T*)p->Zt, Xst, p->m_pad, c_pad, p->D, B, p->m_pad, gp, s);
  forward_subst_multi<T>(ctx, (const T*)p->U, p->lda, (const T*)p->Udinv, p->m_pad, B, p->m_pad, c_pad);  // A
  launch_gemv_t<T>(B, p->m_pad, p->m_pad, M, (const T*)p->m_e, p->mean_kind, p->mean_c, (const T*)nullptr, mu, s);
  GramParams gs{};  // K** (+ Sigma* when factoring), full square: the covariance is returned whole
  fill_gram_params<T>(gs, &p->k, 1, 0, M, M, factor ? noise_s : nullptr, noise_d);
  launch_gram<T>(Xst, Xst, c_pad, c_pad, p->D, Lf, ldf, gs, s);
  {
    GemmArgs g{};  // -= A'A
    g.A = B; g.lda = p->m_pad; g.a_kmajor = 1;
    g.B = B; g.ldb = p->m_pad; g.b_kmajor = 1;
    g.C = Lf; g.ldc = ldf; g.M = c_pad; g.N = c_pad; g.K = p->m_pad; g.alpha_neg = 1; g.beta_one = 1;
    launch_gemm<T>(g, s);
  }
  forward_subst_multi<T>(ctx, (const T*)p->Lam, p->lda, (const T*)p->Ldinv, p->m_pad, B, p->m_pad, c_pad);  // Lam' \ A
  {
    GemmArgs g{};  // += (Lam' \ A)'(Lam' \ A)
    g.A = B; g.lda = p->m_pad; g.a_kmajor = 1;
    g.B = B; g.ldb = p->m_pad; g.b_kmajor = 1;
    g.C = Lf; g.ldc = ldf; g.M = c_pad; g.N = c_pad; g.K = p->m_pad; g.beta_one = 1;
    launch_gemm<T>(g, s);
  }
  if (mean_out) { rc = download<T>(ctx, mean_out, mu, (size_t)M, false); if (rc) return rc; }
  if (!factor) {
    if (cov_out) {
      cudaMemcpyKind kind = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
      CK(cudaMemcpy2DAsync(cov_out, (size_t)M * sizeof(T), Lf, (size_t)ldf * sizeof(T), (size_t)M * sizeof(T), (size_t)M, kind, s));
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return AGP_OK;
  }
  CK(sc.alloc(&tmp, (size_t)nblk * TILE * TILE * sizeof(T)));
  T* Dinv = (T*)tmp;
  CK(sc.alloc(&tmp, (size_t)(nblk + TILE + 2) * sizeof(double)));
  double* dscal = (double*)tmp;
  CK(sc.alloc(&tmp, sizeof(int)));
  int* dinfo = (int*)tmp;
  CK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  CK(sc.alloc(&tmp, (size_t)TILE * sizeof(T)));
  T* lp_d = (T*)tmp;
  launch_border_init<T>(Lf, ldf, M, c_pad, Yd, M, S, 2, 0.0, (const T*)mu, s);  // border = (Y - m*)'
  prof_begin(ctx);
  cholesky_inplace<T>(ctx, Lf, ldf, c_pad, ldf, Dinv, dscal, dinfo);
  if (S > 0) {
    CK(sc.alloc(&tmp, (size_t)S * c_pad * sizeof(T)));
    launch_extract_v<T>(Lf, ldf, c_pad, S, (T*)tmp, dscal + nblk, s);
    launch_finalize_logpdf<T>(dscal, nblk, dscal + nblk, S, M, lp_d, dscal + nblk + TILE, s);
  }
  int h_info = 0;
  CK(cudaMemcpyAsync(&h_info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  if (h_info != 0) {
    ctx->info = h_info;
    char b[128];
    snprintf(b, sizeof(b), "approximate posterior covariance is not positive definite; Cholesky failed at pivot %d", h_info);
    ctx->err = b;
    return AGP_ERR_NOT_POSDEF;
  }
  if (S > 0) CK(cudaMemcpyAsync(logpdf_out, lp_d, (size_t)S * sizeof(T), cudaMemcpyDeviceToHost, s));
  if (Sz > 0) {
    const int64_t s_pad = round_up(Sz, 4);
    CK(sc.alloc(&tmp, (size_t)c_pad * s_pad * sizeof(T) * 2));
    T* Zd = (T*)tmp; T* Od = Zd + c_pad * s_pad;
    CK(cudaMemsetAsync(Zd, 0, (size_t)c_pad * s_pad * sizeof(T), s));
    cudaMemcpyKind kin = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    cudaMemcpyKind kout = ctx->memspace == AGP_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CK(cudaMemcpy2DAsync(Zd, (size_t)c_pad * sizeof(T), Z, (size_t)M * sizeof(T), (size_t)M * sizeof(T), (size_t)Sz, kin, s));
    GemmArgs g{};
    g.A = Lf; g.lda = ldf; g.a_kmajor = 0;
    g.B = Zd; g.ldb = c_pad; g.b_kmajor = 1;
    g.C = Od; g.ldc = c_pad; g.M = c_pad; g.N = s_pad; g.K = c_pad; g.trmm_lower = 1;
    launch_gemm<T>(g, s);
    launch_add_mean_cols<T>(Od, c_pad, M, Sz, 2, 0.0, (const T*)mu, s);
    CK(cudaMemcpy2DAsync(rand_out, (size_t)M * sizeof(T), Od, (size_t)c_pad * sizeof(T), (size_t)M * sizeof(T), (size_t)Sz, kout, s));
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  return AGP_OK;
}

template <typename T>
int fit_dist_impl(agp_ctx* ctx, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise, int layout,
                  const void* X, int64_t N, int D, const void* Y, int S, void* logpdf_out, void* alpha_out,
                  agp_post** post_out) {
  int rc = check_kernel(ctx, k, D);
  if (rc) return rc;
  if (N <= 0) { ctx->err = "N must be positive"; return AGP_ERR_DIM_MISMATCH; }
  if (S < 1 || S > TILE || !Y) { ctx->err = "distributed fit needs 1..128 right-hand sides"; return AGP_ERR_UNSUPPORTED; }
  const bool keep = post_out != nullptr;
  static const agp_mean zero_mean{0, 0.0, nullptr};
  static const agp_noise default_noise{0, 1e-18, nullptr};
  if (!mean) mean = &zero_mean;
  if (!noise) noise = &default_noise;
  cudaStream_t s = ctx->stream, s2 = ctx->stream2;
  CK(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  const int R = ctx->nranks, me = ctx->rank;
  // distribution block = one OUTER panel of G inner 128-blocks (W columns): the owner factors it locally,
  // one NCCL broadcast per outer panel, rank-W trailing updates
  const int G = resolve_G(ctx, round_up(N, TILE));
  const int64_t W = (int64_t)G * TILE;
  const int64_t n_pad = round_up(N, W), lda = n_pad + TILE;
  const int nto = (int)(n_pad / W), nt = (int)(n_pad / TILE);
  const int nloc = (nto - me + R - 1) / R;  // local outer blocks: global jo = ljo * R + me
  const int fp64_mode = resolve_fp64_mode(ctx, n_pad);
  prof_begin(ctx);
  CK(cudaEventRecord(ctx->ev[0], s));
  // big buffers FIRST (the local block columns, then the packed panels): the pool returns the blocks the previous fit freed
  // before the small staging buffers can split them (see fit_impl)
  void *Lv = nullptr, *Dv = nullptr, *Pv = nullptr;
  CK(cudaMallocAsync(&Lv, (size_t)lda * (nloc > 0 ? nloc : 1) * W * sizeof(T), s));
  struct BufGuard { void* p; cudaStream_t s; ~BufGuard() { if (p) cudaFreeAsync(p, s); } } lguard{Lv, s};
  CK(sc.alloc(&Pv, (size_t)3 * lda * W * sizeof(T)));
  T *ard_d = nullptr, *mean_d = nullptr, *noise_d = nullptr, *Yd = nullptr, *Xt = nullptr;
  if (k->transform == AGP_T_ARD) { rc = upload<T>(ctx, sc, k->ard, D, true, &ard_d); if (rc) return rc; }
  if (mean->kind == 2) { rc = upload<T>(ctx, sc, mean->v, N, true, &mean_d); if (rc) return rc; }
  if (noise->kind == 1) { rc = upload<T>(ctx, sc, noise->v, N, true, &noise_d); if (rc) return rc; }
  rc = upload<T>(ctx, sc, Y, (size_t)N * S, false, &Yd); if (rc) return rc;
  rc = prep_points<T>(ctx, sc, k, ard_d, layout, X, N, n_pad, D, &Xt, keep); if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[1], s));
  void* tmp = nullptr;
  // the factor and its inverse diagonal blocks outlive the call when a posterior handle is requested
  agp_post* post = nullptr;
  lguard.p = nullptr;  // ownership passes to the handle / scratch list below
  CK(cudaMallocAsync(&Dv, (size_t)(nloc > 0 ? nloc : 1) * G * TILE * TILE * sizeof(T), s));
  if (keep) {
    post = new agp_post();
    post->ctx = ctx; post->dtype = sizeof(T) == 8 ? AGP_F64 : AGP_F32;
    post->n = N; post->n_pad = n_pad; post->lda = lda; post->D = D;
    post->Lloc = Lv; post->Dinv_loc = Dv; post->Xt = Xt;
    post->dist_R = R; post->dist_me = me; post->dist_G = G; post->dist_nloc = nloc; post->dist_W = W;
    post->k = *k; post->k.ard = nullptr;
    post->mean_kind = mean->kind == 2 ? 0 : mean->kind; post->mean_c = mean->c;
    post->segs.push_back({0, N});
    if (ard_d) { sc.release(ard_d); post->ard = ard_d; }
  } else {
    sc.ptrs.push_back(Lv); sc.ptrs.push_back(Dv);
  }
  struct PostGuard { agp_post* p; ~PostGuard() { if (p) agp_post_free(p); } } guard{post};  // freed on every error return
  T* L = (T*)Lv;
  T* Dinv = (T*)Dv;  // inverse diagonal blocks of the LOCAL 128-blocks
  T* P[3] = {(T*)Pv, (T*)Pv + lda * W, (T*)Pv + 2 * lda * W};  // packed panels (rows_below x W, ld = rows_below): two in
                                                                   // flight in the default schedule, three in the pipelined one
  CK(sc.alloc(&tmp, (size_t)(nt + TILE + 4) * sizeof(double)));
  double* dscal = (double*)tmp;  // [0..nt) logdet parts, [nt..nt+TILE) sqmahal, [nt+TILE] logdet
  CK(cudaMemsetAsync(dscal, 0, (size_t)(nt + TILE + 4) * sizeof(double), s));
  CK(sc.alloc(&tmp, sizeof(int)));
  int* dinfo = (int*)tmp;
  CK(cudaMemsetAsync(dinfo, 0, sizeof(int), s));
  CK(sc.alloc(&tmp, (size_t)(S + 1) * n_pad * sizeof(T)));
  T* rwork = (T*)tmp; T* alpha = rwork + (size_t)S * n_pad;
  CK(cudaMemsetAsync(rwork, 0, (size_t)(S + 1) * n_pad * sizeof(T), s));
  CK(sc.alloc(&tmp, (size_t)TILE * sizeof(T)));
  T* lp_d = (T*)tmp;
  const OzakiWs* oz = nullptr;
  if constexpr (std::is_same<T, double>::value) {
    if (fp64_mode == 1 && nto > 2 && W % 64 == 0) {
      if (!ctx->oz.SL || ctx->oz.K != W || ctx->oz_rows < lda || ctx->oz.S != ctx->oz_S) {
        if (ctx->oz.SL) ozaki_ws_destroy(&ctx->oz, s);
        if (ozaki_ws_create(&ctx->oz, lda, (int)W, ctx->oz_S, s) == 0) ctx->oz_rows = lda;
        else { memset(&ctx->oz, 0, sizeof(ctx->oz)); ctx->oz_rows = 0; }
      }
      if (ctx->oz.SL) oz = &ctx->oz;
    }
  }
  // schedule: 2 = pipelined (default for R > 1; see the loop below), 1 = round-1 owner-first experiment, 0 = plain look-ahead
  int dist_sched = (R > 1) ? 2 : 0;
  { const char* e1 = getenv("AGP_DIST_SCHED"); if (e1) dist_sched = atoi(e1); }
  const OzakiWs* oz_b = nullptr;  // second slice buffer (pipelined schedule only)
  if constexpr (std::is_same<T, double>::value) {
    if (dist_sched == 2 && oz) {
      if (!ctx->oz2.SL || ctx->oz2.K != W || ctx->oz2_rows < lda || ctx->oz2.S != ctx->oz_S) {
        if (ctx->oz2.SL) ozaki_ws_destroy(&ctx->oz2, s);
        if (ozaki_ws_create(&ctx->oz2, lda, (int)W, ctx->oz_S, s) == 0) ctx->oz2_rows = lda;
        else { memset(&ctx->oz2, 0, sizeof(ctx->oz2)); ctx->oz2_rows = 0; }
      }
      oz_b = ctx->oz2.SL ? &ctx->oz2 : nullptr;
      if (!oz_b) dist_sched = 0;
    }
  }
  if (dist_sched == 2 && !ctx->stream_comm) {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    CK(cudaStreamCreateWithPriority(&ctx->stream_comm, cudaStreamNonBlocking, prio_hi));
  }

  // ---- Gram: only the local outer blocks, lower part, + border rows
  for (int lj = 0; lj < nloc; ++lj) {
    const int64_t jo = (int64_t)lj * R + me;
    GramParams gp{};
    fill_gram_params<T>(gp, k, 1, 1, N, N, noise, noise_d);
    gp.diag_off = jo * W;
    T* col = L + (int64_t)lj * W * lda;
    launch_gram<T>(Xt, Xt + jo * W * D, n_pad, W, D, col, lda, gp, s);
    launch_border_init_cols<T>(col, lda, n_pad, jo * W, W, N, Yd, N, S, mean->kind, mean->c, mean_d, s);
  }
  CK(cudaEventRecord(ctx->ev[2], s));

  // ---- distributed right-looking Cholesky with look-ahead
  auto local_first_after = [&](int kk) {  // first local outer block whose global index > kk
    int lj = (kk + 1 - me + R - 1) / R;
    if (lj < 0) lj = 0;
    while ((int64_t)lj * R + me <= kk) ++lj;
    return lj;
  };
  const OzakiWs* oz_cur = oz;  // slice buffer the `trailing` launches read (the pipelined schedule alternates two)
  auto trailing = [&](int kk, T* Pk, int lj_lo, int lj_hi, bool use_oz, cudaStream_t st) {  // local outer blocks [lj_lo, lj_hi)
    if (lj_lo >= lj_hi) return;
    const int64_t rows_below = lda - (int64_t)(kk + 1) * W;
    const int64_t j0 = (int64_t)lj_lo * R + me;
    T* C = L + (int64_t)(kk + 1) * W + (int64_t)lj_lo * W * lda;
    const int64_t Ncols = (int64_t)(lj_hi - lj_lo) * W, b_off = (j0 - (kk + 1)) * W;
    if (ctx->profile) cudaEventRecord(prof_event(ctx), st);
    bool done = false;
    if constexpr (std::is_same<T, double>::value) {
      if (use_oz) { ozaki_syrk(*oz_cur, C, lda, rows_below, Ncols, 1, (int64_t)R * W, W, b_off, 0, st); done = true; }
    }
    if (!done) {
      GemmArgs u{};
      u.A = Pk; u.lda = rows_below; u.B = Pk; u.ldb = rows_below; u.C = C; u.ldc = lda;
      u.M = rows_below; u.N = Ncols; u.K = W; u.alpha_neg = 1; u.beta_one = 1; u.lower_only = 1;
      u.b_tile_stride = (int64_t)R * W; u.b_tile_width = W; u.b_off = b_off;
      launch_gemm<T>(u, st);
    }
    if (ctx->profile) cudaEventRecord(prof_event(ctx), st);
  };
  bool rest_pending = false;
  size_t ev_idx = 0, last_rest = 0;
  // EXPERIMENTAL schedule (AGP_DIST_SCHED=1, not yet validated on a multi-GPU box; default off).  The persistent
  // update kernel holds every SM until it ends, so in the default order the next panel's owner factors it only
  // AFTER its own rest update and every peer's broadcast kernel starts only after theirs: the step costs
  // rest + factor + broadcast.  Here (1) the owner of panel kk+1 defers its rest update of step kk until panel kk+1
  // is factored and packed (the broadcast is enqueued right behind), and (2) rest updates run on nsm - reserve
  // CTAs so that the NCCL kernels of the next broadcast find SMs while the update is still running.
  bool sched2 = false;
  int reserve_sms = 16;
  if (R > 1) {
    sched2 = dist_sched == 1;
    const char* e2 = getenv("AGP_DIST_RESERVE_SMS");
    if (e2) reserve_sms = atoi(e2);
    if (reserve_sms < 0 || reserve_sms > 64) reserve_sms = 16;
  }
  // pipelined schedule: 1 = the owner of the next panel runs its rest update AFTER that panel's factorisation; 0 = at once
  // (bounded CTAs + stream priorities let the factorisation through)
  const bool dist_defer = env_int64("AGP_DIST_DEFER", 1) != 0;
  int nsm_dev = 148;
  cudaDeviceGetAttribute(&nsm_dev, cudaDevAttrMultiProcessorCount, ctx->device);
  struct DeferredRest { bool on; int kk; T* Pk; int lo, hi; bool use_oz; cudaEvent_t e_rest; };
  DeferredRest def{false, 0, nullptr, 0, 0, false, nullptr};
  auto rest_update = [&](int kk, T* Pk, int lo, int hi, bool use_oz, cudaEvent_t e_rest) {
    if (oz && sched2) oz->max_ctas = nsm_dev - reserve_sms;
    trailing(kk, Pk, lo, hi, use_oz, s2);
    if (oz) oz->max_ctas = 0;
    cudaEventRecord(e_rest, s2);
  };
  if (dist_sched == 2) {
    // ---- PIPELINED schedule.  The critical path of a 1 x R factorisation is the owner chain
    //   [panel k-1 received] -> update of block column k -> factor panel k -> broadcast panel k,
    // so (1) the owner factors its panel BEFORE its own rest update of the previous step (deferred to stream2 behind the
    // factorisation); (2) the panel is broadcast in G column pieces on a communication stream, each as soon as its inner
    // block is final, so only the last piece is exposed; (3) rest updates leave `reserve_sms` SMs to the NCCL kernels;
    // (4) panels rotate through three buffers and slices through two, so receiving / slicing panel k+1 never waits for
    // the rest update of step k.  Every rank issues the same collectives in the same order on stream_comm; the
    // collectives that follow the factorisation are issued after the streams are joined.
    cudaStream_t scm = ctx->stream_comm;
    std::vector<cudaEvent_t> e_rest((size_t)nto, nullptr), e_first((size_t)nto, nullptr);
    auto ev = [&]() { return dep_event(ctx, ev_idx++); };
    cudaEvent_t e_start = ev();
    cudaEventRecord(e_start, s);
    cudaStreamWaitEvent(scm, e_start, 0);  // the Gram (and the previous call) precede the first broadcast
    cudaStreamWaitEvent(s2, e_start, 0);
    struct Deferred { bool on; int kk; T* Pk; int lo, hi; bool use_oz; const OzakiWs* ws; cudaEvent_t e_prep; } dfr{false, 0, nullptr, 0, 0, false, nullptr, nullptr};
    auto issue_rest = [&](int kk, T* Pk, int lo, int hi, bool use_oz, const OzakiWs* ws) {
      e_rest[(size_t)kk] = ev();
      if (ws) { ws->max_ctas = nsm_dev - reserve_sms; ws->chunk_tiles = ctx->oz_chunk; }
      oz_cur = ws;
      // The rank that owns block column kk+2 will update it on the MAIN stream at step kk+1 (next-panel update) and then
      // factor it: this rest update must have finished with that column before.  It is the first local block of the
      // range (the rank does not own kk+1), so it goes first, alone, with its own event -- the chain of step kk+1 waits
      // for this piece only, not for the whole rest update.
      if (kk + 2 < nto && (kk + 2) % R == me && hi > lo) {
        e_first[(size_t)kk] = ev();
        trailing(kk, Pk, lo, lo + 1, use_oz, s2);
        cudaEventRecord(e_first[(size_t)kk], s2);
        trailing(kk, Pk, lo + 1, hi, use_oz, s2);
      } else {
        trailing(kk, Pk, lo, hi, use_oz, s2);
      }
      if (ws) { ws->max_ctas = 0; ws->chunk_tiles = 0; }
      cudaEventRecord(e_rest[(size_t)kk], s2);
    };
    for (int kk = 0; kk < nto; ++kk) {
      const int owner = kk % R;
      const int64_t rows_below = lda - (int64_t)(kk + 1) * W;
      T* Pk = P[kk % 3];
      const OzakiWs* ws_k = (kk & 1) ? oz_b : oz;
      if (kk >= 3 && e_rest[(size_t)kk - 3]) cudaStreamWaitEvent(scm, e_rest[(size_t)kk - 3], 0);  // P[kk % 3] is free again
      if (owner == me) {
        const int lk = kk / R;
        T* Lp = L + (int64_t)kk * W + (int64_t)lk * W * lda;
        if (kk >= 3 && e_rest[(size_t)kk - 3]) cudaStreamWaitEvent(s, e_rest[(size_t)kk - 3], 0);  // the pack below writes P[kk % 3]
        for (int g = 0; g < G; ++g) {
          // inner block g of the panel: rows from the panel's diagonal block down to the border
          factor_panel_step<T>(ctx, Lp, lda, G, g, lda - (int64_t)kk * W, Dinv + (int64_t)lk * G * TILE * TILE, dscal, kk * G, dinfo, s);
          launch_copy2d<T>(Lp + W + (int64_t)g * TILE * lda, lda, Pk + (int64_t)g * TILE * rows_below, rows_below, rows_below, TILE, s);
          cudaEvent_t e_col = ev();
          cudaEventRecord(e_col, s);
          cudaStreamWaitEvent(scm, e_col, 0);
          if (rows_below > 0)
            CKN(ncclBroadcast(Pk + (int64_t)g * TILE * rows_below, Pk + (int64_t)g * TILE * rows_below, (size_t)rows_below * TILE,
                              NcclType<T>::v, owner, ctx->nccl, scm));
        }
        if (dfr.on) {  // this rank's rest update of step kk-1, behind the factorisation it must not delay
          cudaEvent_t e_fact = ev();
          cudaEventRecord(e_fact, s);
          cudaStreamWaitEvent(s2, e_fact, 0);
          issue_rest(dfr.kk, dfr.Pk, dfr.lo, dfr.hi, dfr.use_oz, dfr.ws);
          dfr.on = false;
        }
      } else if (rows_below > 0) {
        for (int g = 0; g < G; ++g)
          CKN(ncclBroadcast(Pk + (int64_t)g * TILE * rows_below, Pk + (int64_t)g * TILE * rows_below, (size_t)rows_below * TILE,
                            NcclType<T>::v, owner, ctx->nccl, scm));
      }
      if (kk == nto - 1) break;
      cudaEvent_t e_recv = ev();
      cudaEventRecord(e_recv, scm);
      cudaStreamWaitEvent(s, e_recv, 0);
      if (kk >= 2 && e_rest[(size_t)kk - 2]) cudaStreamWaitEvent(s, e_rest[(size_t)kk - 2], 0);  // slice buffer kk & 1 is free again
      bool use_oz = false;
      if constexpr (std::is_same<T, double>::value) {
        if (ws_k && (int64_t)(nto - 1 - kk) * W >= 2 * TILE) {
          ozaki_prepare(*ws_k, (const double*)Pk, rows_below, rows_below, s);
          use_oz = true;
        }
      }
      cudaEvent_t e_prep = ev();
      cudaEventRecord(e_prep, s);
      const int lj_first = local_first_after(kk);
      int lj_bulk = lj_first;
      const bool next_is_mine = (kk + 1) % R == me;
      if (next_is_mine) {  // the next panel's block column first, on the main stream; its factorisation follows
        if (kk >= 1 && e_first[(size_t)kk - 1]) cudaStreamWaitEvent(s, e_first[(size_t)kk - 1], 0);  // see issue_rest
        else if (kk >= 1 && e_rest[(size_t)kk - 1]) cudaStreamWaitEvent(s, e_rest[(size_t)kk - 1], 0);
        oz_cur = ws_k;
        trailing(kk, Pk, lj_first, lj_first + 1, use_oz, s);
        lj_bulk = lj_first + 1;
      }
      if (next_is_mine && dist_defer) {
        dfr = Deferred{true, kk, Pk, lj_bulk, nloc, use_oz, use_oz ? ws_k : nullptr, e_prep};
      } else {
        cudaStreamWaitEvent(s2, e_prep, 0);
        issue_rest(kk, Pk, lj_bulk, nloc, use_oz, use_oz ? ws_k : nullptr);
      }
    }
    if (dfr.on) {  // cannot happen (the last step has no trailing work), kept for safety
      cudaStreamWaitEvent(s2, dfr.e_prep, 0);
      issue_rest(dfr.kk, dfr.Pk, dfr.lo, dfr.hi, dfr.use_oz, dfr.ws);
    }
    cudaEvent_t e_s2 = ev(), e_cm = ev();
    cudaEventRecord(e_s2, s2);
    cudaEventRecord(e_cm, scm);
    cudaStreamWaitEvent(s, e_s2, 0);
    cudaStreamWaitEvent(s, e_cm, 0);
    oz_cur = oz;
  } else
  for (int kk = 0; kk < nto; ++kk) {
    const int owner = kk % R;
    const int64_t rows_below = lda - (int64_t)(kk + 1) * W;
    T* Pk = P[kk & 1];
    if (owner == me) {
      const int lk = kk / R;
      T* Lp = L + (int64_t)kk * W + (int64_t)lk * W * lda;
      factor_panel<T>(ctx, Lp, lda, G, lda - (int64_t)kk * W, Dinv + (int64_t)lk * G * TILE * TILE, dscal, kk * G, dinfo, s);
      launch_copy2d<T>(Lp + W, lda, Pk, rows_below, rows_below, W, s);
    }
    if (def.on) {  // this rank owns panel kk and still owes step kk-1's rest update: it goes behind the factorisation
      cudaEvent_t e_fact = dep_event(ctx, ev_idx++);
      cudaEventRecord(e_fact, s);
      cudaStreamWaitEvent(s2, e_fact, 0);
      rest_update(def.kk, def.Pk, def.lo, def.hi, def.use_oz, def.e_rest);
      def.on = false;
    }
    if (R > 1) CKN(ncclBroadcast(Pk, Pk, (size_t)rows_below * W, NcclType<T>::v, owner, ctx->nccl, s));
    if (kk == nto - 1) break;
    if (rest_pending) cudaStreamWaitEvent(s, dep_event(ctx, last_rest), 0);  // frees P[(kk+1)&1] and the slice buffer
    bool use_oz = false;
    if constexpr (std::is_same<T, double>::value) {
      if (oz && (int64_t)(nto - 1 - kk) * W >= 2 * TILE) {
        ozaki_prepare(*oz, (const double*)Pk, rows_below, rows_below, s);
        use_oz = true;
      }
    }
    cudaEvent_t e_panel = dep_event(ctx, ev_idx++), e_rest = dep_event(ctx, ev_idx++);
    cudaEventRecord(e_panel, s);
    const int lj_first = local_first_after(kk);
    int lj_bulk = lj_first;
    if ((kk + 1) % R == me) {  // next panel first (only its owner has it), on the main stream
      trailing(kk, Pk, lj_first, lj_first + 1, use_oz, s);
      lj_bulk = lj_first + 1;
    }
    rest_pending = true;
    last_rest = ev_idx - 1;  // index of e_rest
    if (sched2 && (kk + 1) % R == me) {
      def = DeferredRest{true, kk, Pk, lj_bulk, nloc, use_oz, e_rest};  // launched at the top of the next iteration
    } else {
      cudaStreamWaitEvent(s2, e_panel, 0);
      rest_update(kk, Pk, lj_bulk, nloc, use_oz, e_rest);
    }
  }
  if (rest_pending) cudaStreamWaitEvent(s, dep_event(ctx, last_rest), 0);
  join_inverses(ctx);
  CK(cudaEventRecord(ctx->ev[3], s));

  // ---- v = border rows (distributed by column), sqmahal and logdet via all-reduce
  for (int lj = 0; lj < nloc; ++lj) {
    const int64_t jo = (int64_t)lj * R + me;
    for (int sI = 0; sI < S; ++sI)
      launch_copy2d<T>(L + n_pad + sI + (int64_t)lj * W * lda, lda, rwork + (size_t)sI * n_pad + jo * W, 1, 1, W, s);
  }
  for (int sI = 0; sI < S; ++sI) launch_sumsq<T>(rwork + (size_t)sI * n_pad, n_pad, dscal + nt + sI, s);
  if (R > 1) CKN(ncclAllReduce(dscal, dscal, (size_t)(nt + TILE), ncclDouble, ncclSum, ctx->nccl, s));
  // ---- distributed backward substitution for column 0: alpha = L^-T v.  The owner of an outer block holds its G diagonal
  // blocks and every tile between them, so it resolves the whole block in ONE single-CTA kernel; ONE broadcast per outer
  // block ships G*128 values (128 collectives at C4).  The critical chain per block is
  //   [alpha of block io+1 arrives] -> update of block io only (G CTAs) -> block solve -> broadcast;
  // the bulk update of the other local columns with block io+1's alpha runs behind the broadcast on the owner and before it
  // on the other ranks (who would otherwise idle in the collective).
  {
    int pending = -1;  // outer block whose alpha has been received but not yet applied to (all of) the local columns
    for (int io = nto - 1; io >= 0; --io) {
      const int owner = io % R, i_lo = io * G;
      const int64_t p_lo = (int64_t)(pending >= 0 ? pending : 0) * G;
      T* a_pend = alpha + p_lo * TILE;
      if (owner == me) {
        if (pending >= 0)  // own block first
          launch_bwd_update_local_multi<T>(L, lda, (int)p_lo, G, a_pend, rwork, nloc * G, me, R, G, i_lo, (int64_t)i_lo + G, s);
        launch_bwd_block_solve<T>(L + (int64_t)io * W + (int64_t)(io / R) * W * lda, lda, Dinv + (int64_t)(io / R) * G * TILE * TILE,
                                  rwork + (int64_t)i_lo * TILE, alpha + (int64_t)i_lo * TILE, G, s);
      } else if (pending >= 0) {
        launch_bwd_update_local_multi<T>(L, lda, (int)p_lo, G, a_pend, rwork, nloc * G, me, R, G, 0, p_lo, s);
      }
      // alpha is zero on every rank but the owner (the buffer starts zeroed and only owners write their blocks), so an
      // all-reduce(sum) IS the broadcast -- and its small-message latency (tree / NVLS) does not grow with the ring length
      // like ncclBroadcast's (measured 0.17 ms per block at 8 ranks with the broadcast)
      if (R > 1) CKN(ncclAllReduce(alpha + (int64_t)i_lo * TILE, alpha + (int64_t)i_lo * TILE, (size_t)G * TILE, NcclType<T>::v, ncclSum, ctx->nccl, s));
      if (owner == me && pending >= 0)
        launch_bwd_update_local_multi<T>(L, lda, (int)p_lo, G, a_pend, rwork, nloc * G, me, R, G, 0, i_lo, s);
      pending = io;
    }
  }
  launch_finalize_logpdf<T>(dscal, nt, dscal + nt, S, N, lp_d, dscal + nt + TILE, s);
  CK(cudaEventRecord(ctx->ev[4], s));
  int h_info = 0;
  if (R > 1) CKN(ncclAllReduce(dinfo, dinfo, 1, ncclInt, ncclMax, ctx->nccl, s));
  CK(cudaMemcpyAsync(&h_info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (logpdf_out) CK(cudaMemcpyAsync(logpdf_out, lp_d, (size_t)S * sizeof(T), cudaMemcpyDeviceToHost, s));
  if (alpha_out) { rc = download<T>(ctx, alpha_out, alpha, (size_t)N, false); if (rc) return rc; }
  CK(cudaEventRecord(ctx->ev[5], s));
  CK(cudaStreamSynchronize(s));
  CK(cudaStreamSynchronize(s2));
  CK(cudaGetLastError());
  float ms = 0;
  auto el = [&](int a, int b) { cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]); return (double)ms; };
  ctx->timings[0] = el(0, 5); ctx->timings[1] = el(0, 1); ctx->timings[2] = el(1, 2); ctx->timings[3] = el(2, 3);
  ctx->timings[4] = el(3, 4); ctx->timings[5] = el(4, 5); ctx->timings[6] = 0.0;
  ctx->timings[7] = ctx->profile ? prof_total_ms(ctx) : 0.0;
  if (h_info != 0) {
    ctx->info = h_info;
    ctx->err = "matrix is not positive definite (distributed Cholesky)";
    return AGP_ERR_NOT_POSDEF;
  }
  if (keep) {  // alpha, delta = y - m (first column) and log det move into the handle
    CK(cudaMallocAsync(&post->alpha, (size_t)n_pad * sizeof(T), s));
    CK(cudaMallocAsync(&post->delta, (size_t)n_pad * sizeof(T), s));
    CK(cudaMemcpyAsync(post->alpha, alpha, (size_t)n_pad * sizeof(T), cudaMemcpyDeviceToDevice, s));
    CK(cudaMemsetAsync(post->delta, 0, (size_t)n_pad * sizeof(T), s));
    launch_sub_mean<T>(Yd, N, mean->kind, mean->c, mean_d, (T*)post->delta, s);
    double h_logdet = 0.0;
    CK(cudaMemcpyAsync(&h_logdet, dscal + nt + TILE, sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    post->logdet = h_logdet;
    guard.p = nullptr;
    *post_out = post;
  }
  return AGP_OK;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// extern "C" ABI
// ------------------------------------------------------------------------------------------------
#define DISPATCH(dtype, call_f32, call_f64)                                   \
  ((dtype) == AGP_F32 ? (call_f32) : ((dtype) == AGP_F64 ? (call_f64) : (int)AGP_ERR_UNSUPPORTED))

extern "C" {

const char* agp_version(void) { return "agp-blackwell 0.1 (sm_100a)"; }

int32_t agp_init(agp_ctx** out, int32_t device, const agp_config* cfg) {
  if (!out) return AGP_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return AGP_ERR_CUDA;  // no CPU fallback, by design
  if (device < 0 || device >= ndev) return AGP_ERR_INVALID;
  agp_ctx* ctx = new agp_ctx();
  ctx->device = device;
  if (cfg) ctx->cfg = *cfg;
  ctx->cfg.tile_nb = env_int("AGP_NB", (cfg && cfg->tile_nb > 0) ? cfg->tile_nb : 0);  // 0 = auto
  if (ctx->cfg.tile_nb % TILE) ctx->cfg.tile_nb = 0;
  ctx->cfg.fp64_mode = env_int("AGP_FP64_MODE", cfg ? ctx->cfg.fp64_mode : -1);            // -1 = auto
  ctx->cfg.fp32_mode = env_int("AGP_FP32_MODE", cfg ? ctx->cfg.fp32_mode : -1);            // -1 = auto
  ctx->cfg.lookahead = env_int("AGP_LOOKAHEAD", cfg ? ctx->cfg.lookahead : 2);  // 2: depth-2 look-ahead on the DMMA path (validated in round 2: C2 3.49 -> 3.27 ms)
  ctx->cfg.use_graph = env_int("AGP_GRAPH", ctx->cfg.use_graph);
  ctx->profile = env_int("AGP_PROFILE", cfg ? cfg->profile_kernels : 0);
  ctx->oz_S = env_int("AGP_OZAKI_S", (cfg && cfg->ozaki_slices) ? cfg->ozaki_slices : 7);
  if (ctx->oz_S < 5 || ctx->oz_S > 8) ctx->oz_S = 7;
  ctx->oz_S32 = env_int("AGP_OZAKI_S32", 4);
  if (ctx->oz_S32 < 3 || ctx->oz_S32 > 5) ctx->oz_S32 = 4;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return AGP_ERR_CUDA; }
  // priorities: the panel chain (main stream), the strip inverses and the panel broadcasts go first; the bulk trailing
  // updates (stream2, bounded CTAs) fill whatever SMs are left -- block scheduling honours stream priority
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete ctx; return AGP_ERR_CUDA; }
  if (cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete ctx; return AGP_ERR_CUDA; }
  if (cudaStreamCreateWithPriority(&ctx->stream3, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete ctx; return AGP_ERR_CUDA; }
  ctx->oz_chunk = env_int("AGP_OZAKI_CHUNK", 16);
  if (ctx->oz_chunk < 0 || ctx->oz_chunk > 4096) ctx->oz_chunk = 16;
  cudaEventCreateWithFlags(&ctx->ev_s3, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_fac, cudaEventDisableTiming);
  for (int i = 0; i < 8; ++i) cudaEventCreate(&ctx->ev[i]);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;  // keep freed blocks: repeated fits of the same size never hit the OS
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = ctx;
  return AGP_OK;
}

int32_t agp_destroy(agp_ctx* ctx) {
  if (!ctx) return AGP_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 8; ++i) cudaEventDestroy(ctx->ev[i]);
  for (auto e : ctx->prof_ev) cudaEventDestroy(e);
  for (auto e : ctx->dep_ev) cudaEventDestroy(e);
  if (ctx->oz.SL) ozaki_ws_destroy(&ctx->oz, ctx->stream);
  if (ctx->oz2.SL) ozaki_ws_destroy(&ctx->oz2, ctx->stream);
  if (ctx->stream_comm) { cudaStreamSynchronize(ctx->stream_comm); cudaStreamDestroy(ctx->stream_comm); }
  if (ctx->nccl) ncclCommDestroy(ctx->nccl);
  cudaEventDestroy(ctx->ev_s3);
  cudaEventDestroy(ctx->ev_fac);
  cudaStreamDestroy(ctx->stream3);
  cudaStreamDestroy(ctx->stream2);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return AGP_OK;
}

const char* agp_last_error(const agp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
int64_t agp_last_info(const agp_ctx* ctx) { return ctx ? ctx->info : 0; }
int32_t agp_set_memspace(agp_ctx* ctx, int32_t m) {
  if (!ctx || (m != AGP_MEM_HOST && m != AGP_MEM_DEVICE)) return AGP_ERR_INVALID;
  ctx->memspace = m;
  return AGP_OK;
}
int32_t agp_set_config(agp_ctx* ctx, const agp_config* cfg) {
  if (!ctx || !cfg) return AGP_ERR_INVALID;
  if (cfg->tile_nb < 0 || cfg->tile_nb % TILE) return AGP_ERR_INVALID;
  ctx->cfg.tile_nb = cfg->tile_nb;
  ctx->cfg.fp64_mode = cfg->fp64_mode;
  ctx->cfg.fp32_mode = cfg->fp32_mode;
  ctx->cfg.lookahead = cfg->lookahead;
  ctx->profile = cfg->profile_kernels;
  if (cfg->ozaki_slices >= 5 && cfg->ozaki_slices <= 8) { ctx->cfg.ozaki_slices = cfg->ozaki_slices; ctx->oz_S = cfg->ozaki_slices; }
  return AGP_OK;
}
int32_t agp_get_config(const agp_ctx* ctx, agp_config* out) {
  if (!ctx || !out) return AGP_ERR_INVALID;
  *out = ctx->cfg;
  out->ozaki_slices = ctx->oz_S;
  out->profile_kernels = ctx->profile;
  return AGP_OK;
}
int32_t agp_last_timings(const agp_ctx* ctx, double* out, int32_t n) {
  if (!ctx || !out) return 0;
  int c = n < 8 ? n : 8;
  for (int i = 0; i < c; ++i) out[i] = ctx->timings[i];
  return c;
}
int64_t agp_launch_count(const agp_ctx*) { return agp_kernel_launches(); }

int32_t agp_gram(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, int32_t layout, const void* X, int64_t N,
                 int32_t D, const void* Z, int64_t M, const agp_noise* noise, void* K_out) {
  if (!ctx) return AGP_ERR_INVALID;
  return DISPATCH(dtype, gram_impl<float>(ctx, k, layout, X, N, D, Z, M, noise, K_out),
                  gram_impl<double>(ctx, k, layout, X, N, D, Z, M, noise, K_out));
}

int32_t agp_fit(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise,
                int32_t layout, const void* X, int64_t N, int32_t D, const void* Y, int32_t S, void* logpdf_out,
                void* alpha_out, agp_post** post_out) {
  if (!ctx) return AGP_ERR_INVALID;
  if (post_out) *post_out = nullptr;
  if (ctx->nccl) {  // distributed context: every rank calls with the same (replicated) inputs
    return DISPATCH(dtype, fit_dist_impl<float>(ctx, k, mean, noise, layout, X, N, D, Y, S, logpdf_out, alpha_out, post_out),
                    fit_dist_impl<double>(ctx, k, mean, noise, layout, X, N, D, Y, S, logpdf_out, alpha_out, post_out));
  }
  return DISPATCH(dtype, fit_many_impl<float>(ctx, k, mean, noise, layout, X, N, D, Y, S, logpdf_out, alpha_out, post_out),
                  fit_many_impl<double>(ctx, k, mean, noise, layout, X, N, D, Y, S, logpdf_out, alpha_out, post_out));
}

int32_t agp_post_mean_var(agp_post* p, int32_t layout, const void* Xs, int64_t M, const agp_mean* mean_s,
                          const agp_noise* noise_s, void* mean_out, void* var_out) {
  if (!p) return AGP_ERR_INVALID;
  if (p->dist_R > 1)
    return DISPATCH(p->dtype, post_mean_var_dist<float>(p, layout, Xs, M, mean_s, noise_s, mean_out, var_out),
                    post_mean_var_dist<double>(p, layout, Xs, M, mean_s, noise_s, mean_out, var_out));
  return DISPATCH(p->dtype, post_mean_var_impl<float>(p, layout, Xs, M, mean_s, noise_s, mean_out, var_out),
                  post_mean_var_impl<double>(p, layout, Xs, M, mean_s, noise_s, mean_out, var_out));
}

int32_t agp_post_mean_cov(agp_post* p, int32_t layout, const void* Xs, int64_t M, const agp_mean* mean_s,
                          void* mean_out, void* cov_out) {
  if (!p) return AGP_ERR_INVALID;
  return DISPATCH(p->dtype, post_mean_cov_impl<float>(p, layout, Xs, M, mean_s, mean_out, cov_out),
                  post_mean_cov_impl<double>(p, layout, Xs, M, mean_s, mean_out, cov_out));
}

int32_t agp_post_logpdf(agp_post* p, int32_t layout, const void* Xs, int64_t M, const agp_mean* mean_s,
                        const agp_noise* noise_s, const void* Y, int32_t S, void* logpdf_out) {
  if (!p) return AGP_ERR_INVALID;
  if (S <= 0) { p->ctx->err = "S must be positive"; return AGP_ERR_INVALID; }
  return DISPATCH(p->dtype, post_cond_impl<float>(p, layout, Xs, M, mean_s, noise_s, Y, S, logpdf_out, nullptr, 0, nullptr),
                  post_cond_impl<double>(p, layout, Xs, M, mean_s, noise_s, Y, S, logpdf_out, nullptr, 0, nullptr));
}

int32_t agp_post_rand(agp_post* p, int32_t layout, const void* Xs, int64_t M, const agp_mean* mean_s,
                      const agp_noise* noise_s, const void* Z, int32_t S, void* out) {
  if (!p) return AGP_ERR_INVALID;
  if (S <= 0) return AGP_OK;
  return DISPATCH(p->dtype, post_cond_impl<float>(p, layout, Xs, M, mean_s, noise_s, nullptr, 0, nullptr, Z, S, out),
                  post_cond_impl<double>(p, layout, Xs, M, mean_s, noise_s, nullptr, 0, nullptr, Z, S, out));
}

int32_t agp_post_logpdf_grad(agp_post* p, double* grad_out, void* noise_diag_out) {
  if (!p || !grad_out) return AGP_ERR_INVALID;
  return DISPATCH(p->dtype, post_logpdf_grad_impl<float>(p, grad_out, noise_diag_out),
                  post_logpdf_grad_impl<double>(p, grad_out, noise_diag_out));
}

int32_t agp_post_solve_lower(agp_post* p, const void* B, int64_t nrhs, void* V_out) {
  if (!p || !B || !V_out) return AGP_ERR_INVALID;
  return DISPATCH(p->dtype, post_solve_lower_impl<float>(p, B, nrhs, V_out), post_solve_lower_impl<double>(p, B, nrhs, V_out));
}

int32_t agp_post_factor_export(agp_post* p, void* U_out) {
  if (!p || !U_out) return AGP_ERR_INVALID;
  return DISPATCH(p->dtype, post_export_impl<float>(p, U_out), post_export_impl<double>(p, U_out));
}

int32_t agp_post_logdet(agp_post* p, double* out) {
  if (!p || !out) return AGP_ERR_INVALID;
  *out = p->logdet;
  return AGP_OK;
}
int64_t agp_post_n(const agp_post* p) { return p ? p->n : 0; }

int32_t agp_post_free(agp_post* p) {
  if (!p) return AGP_OK;
  cudaSetDevice(p->ctx->device);
  cudaStream_t s = p->ctx->stream;
  if (p->L) cudaFreeAsync(p->L, s);
  if (p->Dinv) cudaFreeAsync(p->Dinv, s);
  if (p->Xt) cudaFreeAsync(p->Xt, s);
  if (p->alpha) cudaFreeAsync(p->alpha, s);
  if (p->ard) cudaFreeAsync(p->ard, s);
  if (p->delta) cudaFreeAsync(p->delta, s);
  if (p->valid) cudaFreeAsync(p->valid, s);
  if (p->Lloc) cudaFreeAsync(p->Lloc, s);
  if (p->Dinv_loc) cudaFreeAsync(p->Dinv_loc, s);
  delete p;
  return AGP_OK;
}

int32_t agp_rand(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise,
                 int32_t layout, const void* X, int64_t N, int32_t D, const void* Z, int32_t S, void* out) {
  if (!ctx) return AGP_ERR_INVALID;
  return DISPATCH(dtype, rand_impl<float>(ctx, k, mean, noise, layout, X, N, D, Z, S, out),
                  rand_impl<double>(ctx, k, mean, noise, layout, X, N, D, Z, S, out));
}

int32_t agp_debug_ozaki_syrk(agp_ctx* ctx, void* C_dev, int64_t ldc, const void* P_dev, int64_t lda, int64_t M, int64_t N,
                             int32_t K, int32_t S, int32_t lower_only) {
  if (!ctx || !C_dev || !P_dev) return AGP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  OzakiWs ws;
  int rc = ozaki_ws_create(&ws, M, K, S, ctx->stream);
  if (rc) { ctx->err = "ozaki_ws_create failed (code " + std::to_string(rc) + ")"; return rc == 1 ? AGP_ERR_INVALID : AGP_ERR_CUDA; }
  if (!(lower_only && N % 128 == 0 && N >= 128)) ws.bulk = 0;  // the non-persistent kernel reads the row-major slice layout
  ozaki_prepare(ws, (const double*)P_dev, lda, M, ctx->stream);
  ozaki_syrk(ws, (double*)C_dev, ldc, M, N, lower_only, 0, 0, 0, 0, ctx->stream);
  ozaki_ws_destroy(&ws, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("ozaki syrk: ") + cudaGetErrorString(e); return AGP_ERR_CUDA; }
  return AGP_OK;
}

// general product on the tcgen05 path: C (M x N, fp32 or fp64) += sign * A B', A = M x K, B = N x K (B_dev == NULL: B = A,
// lower tiles only).  Each operand is fp32 or fp64, row-contiguous (element (r, k) at [r + k*ld]) or k-major ([k + r*ld]).
// Both operands are sliced into one workspace (A rows first, B rows from the next multiple of 128).
int32_t agp_debug_ozaki_gemm(agp_ctx* ctx, void* C_dev, int32_t c_is_float, int64_t ldc, const void* A_dev, int32_t a_is_float,
                             int32_t a_kmajor, int64_t lda, int64_t M, const void* B_dev, int32_t b_is_float, int32_t b_kmajor,
                             int64_t ldb, int64_t N, int32_t K, int32_t S, double sign) {
  if (!ctx || !C_dev || !A_dev) return AGP_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const int64_t m_pad = (M + 127) / 128 * 128, n_rows = B_dev ? (N + 127) / 128 * 128 : 0;
  OzakiWs ws;
  int rc = ozaki_ws_create(&ws, m_pad + n_rows, K, S, ctx->stream);
  if (rc) { ctx->err = "ozaki_ws_create failed (code " + std::to_string(rc) + ")"; return rc == 1 ? AGP_ERR_INVALID : AGP_ERR_CUDA; }
  ozaki_prepare_ex(ws, A_dev, a_is_float, a_kmajor, lda, M, 0, ctx->stream);
  if (B_dev) ozaki_prepare_ex(ws, B_dev, b_is_float, b_kmajor, ldb, N, m_pad, ctx->stream);
  const int urc = ozaki_update_ex(ws, C_dev, c_is_float, ldc, M, N, B_dev ? 1 : 0, sign, 0, 0, B_dev ? m_pad : 0, 0, ctx->stream);
  ozaki_ws_destroy(&ws, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (urc) { ctx->err = "ozaki_update_ex: unsupported shape / slice count (N % 128, S)"; return AGP_ERR_UNSUPPORTED; }
  if (e != cudaSuccess) { ctx->err = std::string("ozaki gemm: ") + cudaGetErrorString(e); return AGP_ERR_CUDA; }
  return AGP_OK;
}

// block-cyclic column map of the distributed trailing update on ONE device: P has m_panel rows; row r of C pairs with
// panel row r + a_off, local column n with panel row (n / b_tile_width) * b_tile_stride + n % b_tile_width + b_off.
// This drives the strip-table enumeration of the persistent kernel exactly as fit_dist_impl does.
int32_t agp_debug_ozaki_syrk_map(agp_ctx* ctx, void* C_dev, int64_t ldc, const void* P_dev, int64_t lda, int64_t m_panel,
                                 int64_t M, int64_t N, int32_t K, int32_t S, int64_t b_tile_stride, int64_t b_tile_width,
                                 int64_t b_off, int64_t a_off) {
  if (!ctx || !C_dev || !P_dev) return AGP_ERR_INVALID;
  if (N % 128 != 0 || N < 128 || M <= 0 || m_panel <= 0) { ctx->err = "N must be a positive multiple of 128"; return AGP_ERR_INVALID; }
  cudaSetDevice(ctx->device);
  OzakiWs ws;
  int rc = ozaki_ws_create(&ws, m_panel, K, S, ctx->stream);
  if (rc) { ctx->err = "ozaki_ws_create failed (code " + std::to_string(rc) + ")"; return rc == 1 ? AGP_ERR_INVALID : AGP_ERR_CUDA; }
  ozaki_prepare(ws, (const double*)P_dev, lda, m_panel, ctx->stream);
  ozaki_syrk(ws, (double*)C_dev, ldc, M, N, 1, b_tile_stride, b_tile_width, b_off, a_off, ctx->stream);
  ozaki_ws_destroy(&ws, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("ozaki syrk (map): ") + cudaGetErrorString(e); return AGP_ERR_CUDA; }
  return AGP_OK;
}

int32_t agp_bc_owner(int32_t ti, int32_t tj, int32_t P, int32_t Q) { return (ti % P) * Q + (tj % Q); }
int64_t agp_bc_local_tiles(int32_t nt, int32_t rank, int32_t P, int32_t Q) {
  int64_t c = 0;
  for (int i = 0; i < nt; ++i)
    for (int j = 0; j <= i; ++j)
      if (agp_bc_owner(i, j, P, Q) == rank) ++c;
  return c;
}

// ---- multi-GPU entry points are provided by dist.cu ------------------------------------------------
int32_t agp_post_extend(agp_post* p, int32_t layout, const void* X2, int64_t N2, const void* y2, const agp_mean* mean2,
                        const agp_noise* noise2, void* alpha_out, agp_post** post_out) {
  if (!p) return AGP_ERR_INVALID;
  if (post_out) *post_out = nullptr;
  return DISPATCH(p->dtype, post_extend_impl<float>(p, layout, X2, N2, y2, mean2, noise2, alpha_out, post_out),
                  post_extend_impl<double>(p, layout, X2, N2, y2, mean2, noise2, alpha_out, post_out));
}
int32_t agp_vfe_elbo(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise,
                     int32_t layout, const void* X, int64_t N, int32_t D, const void* Zind, int64_t M,
                     const agp_noise* jitter, const void* y, void* elbo_out, void* dtc_out) {
  if (!ctx) return AGP_ERR_INVALID;
  return DISPATCH(dtype, vfe_core<float>(ctx, k, mean, noise, layout, X, N, D, Zind, M, jitter, y, elbo_out, dtc_out, nullptr),
                  vfe_core<double>(ctx, k, mean, noise, layout, X, N, D, Zind, M, jitter, y, elbo_out, dtc_out, nullptr));
}
int32_t agp_vfe_fit(agp_ctx* ctx, int32_t dtype, const agp_kernel* k, const agp_mean* mean, const agp_noise* noise,
                    int32_t layout, const void* X, int64_t N, int32_t D, const void* Zind, int64_t M,
                    const agp_noise* jitter, const void* y, agp_vfe_post** out) {
  if (!ctx || !out) return AGP_ERR_INVALID;
  *out = nullptr;
  return DISPATCH(dtype, vfe_core<float>(ctx, k, mean, noise, layout, X, N, D, Zind, M, jitter, y, nullptr, nullptr, out),
                  vfe_core<double>(ctx, k, mean, noise, layout, X, N, D, Zind, M, jitter, y, nullptr, nullptr, out));
}
int32_t agp_vfe_mean_var(agp_vfe_post* p, int32_t layout, const void* Xs, int64_t Ms, void* mean_out, void* var_out) {
  if (!p) return AGP_ERR_INVALID;
  return DISPATCH(p->dtype, vfe_mean_var_impl<float>(p, layout, Xs, Ms, mean_out, var_out),
                  vfe_mean_var_impl<double>(p, layout, Xs, Ms, mean_out, var_out));
}
int32_t agp_vfe_mean_cov(agp_vfe_post* p, int32_t layout, const void* Xs, int64_t M, void* mean_out, void* cov_out) {
  if (!p) return AGP_ERR_INVALID;
  return DISPATCH(p->dtype, vfe_cond_impl<float>(p, layout, Xs, M, nullptr, mean_out, cov_out, nullptr, 0, nullptr, nullptr, 0, nullptr),
                  vfe_cond_impl<double>(p, layout, Xs, M, nullptr, mean_out, cov_out, nullptr, 0, nullptr, nullptr, 0, nullptr));
}
int32_t agp_vfe_post_logpdf(agp_vfe_post* p, int32_t layout, const void* Xs, int64_t M, const agp_noise* noise_s,
                            const void* Y, int32_t S, void* logpdf_out) {
  if (!p) return AGP_ERR_INVALID;
  if (S <= 0) { p->ctx->err = "S must be positive"; return AGP_ERR_INVALID; }
  return DISPATCH(p->dtype, vfe_cond_impl<float>(p, layout, Xs, M, noise_s, nullptr, nullptr, Y, S, logpdf_out, nullptr, 0, nullptr),
                  vfe_cond_impl<double>(p, layout, Xs, M, noise_s, nullptr, nullptr, Y, S, logpdf_out, nullptr, 0, nullptr));
}
int32_t agp_vfe_post_rand(agp_vfe_post* p, int32_t layout, const void* Xs, int64_t M, const agp_noise* noise_s,
                          const void* Z, int32_t S, void* out) {
  if (!p) return AGP_ERR_INVALID;
  if (S <= 0) return AGP_OK;
  return DISPATCH(p->dtype, vfe_cond_impl<float>(p, layout, Xs, M, noise_s, nullptr, nullptr, nullptr, 0, nullptr, Z, S, out),
                  vfe_cond_impl<double>(p, layout, Xs, M, noise_s, nullptr, nullptr, nullptr, 0, nullptr, Z, S, out));
}
int32_t agp_vfe_post_free(agp_vfe_post* p) {
  if (!p) return AGP_OK;
  cudaSetDevice(p->ctx->device);
  cudaStream_t s = p->ctx->stream;
  void* ptrs[] = {p->U, p->Udinv, p->Lam, p->Ldinv, p->Zt, p->m_e, p->ard};
  for (void* q : ptrs) if (q) cudaFreeAsync(q, s);
  delete p;
  return AGP_OK;
}
int32_t agp_nccl_unique_id(void* out128) {
  if (!out128) return AGP_ERR_INVALID;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return AGP_ERR_NCCL;
  memcpy(out128, &id, 128);
  return AGP_OK;
}
int32_t agp_init_dist(agp_ctx** out, int32_t device, int32_t rank, int32_t nranks, int32_t grid_p, int32_t grid_q,
                      const void* id128, const agp_config* cfg) {
  if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return AGP_ERR_INVALID;
  if (grid_p != 1 || grid_q != nranks) return AGP_ERR_UNSUPPORTED;  // 1 x Q block-column-cyclic grid in this build
  int32_t rc = agp_init(out, device, cfg);
  if (rc != AGP_OK) return rc;
  agp_ctx* ctx = *out;
  ctx->rank = rank; ctx->nranks = nranks; ctx->grid_p = grid_p; ctx->grid_q = grid_q;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  // the panel broadcasts run beside the persistent trailing-update kernel, which leaves a few SMs free: keep NCCL's
  // kernels within that reserve
  ncclConfig_t ncfg = NCCL_CONFIG_INITIALIZER;
  ncfg.maxCTAs = env_int("AGP_NCCL_MAX_CTAS", 16);
  if (ncclCommInitRankConfig(&ctx->nccl, nranks, id, rank, &ncfg) != ncclSuccess) {
    agp_destroy(ctx);
    *out = nullptr;
    return AGP_ERR_NCCL;
  }
  return AGP_OK;
}

}  // extern "C"

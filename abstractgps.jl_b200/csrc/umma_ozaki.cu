// umma_ozaki.cu -- K5 on the 5th-generation tensor cores: the trailing update and every other large product of the path
//     C (m x n, fp64 or fp32)  +=  sign * A B'      (trailing update: B = A = the factored outer panel, lower tiles, sign -1)
// executed as EXACT int8 x int8 -> int32 products on tcgen05.mma.kind::i8 with TMEM accumulators, operands staged by TMA,
// recombined exactly and rounded once in the epilogue (Ozaki-style error-free splitting).  Replaces the LAPACK potrf
// trailing update inside cholesky(_symmetric(C)) (/root/reference/src/finite_gp_projection.jl:308,
// /root/reference/src/exact_gpr_posterior.jl:31), the rank-512 updates of `C.U' \ X`
// (/root/reference/src/util/common_covmat_ops.jl:54,90) and the A A' accumulation of the VFE bound
// (/root/reference/src/sparse_approximations.jl:296-299).
//
// Splitting.  Row i of an operand is scaled by 2^-e_i (e_i: exponent of the row maximum) to |x| < 1 and cut into
// S signed 7-bit slices  x = sum_s q_s 2^-(7s-1),  q_s in [-64, 64]  -- every step exact in fp64.
// Then  a_i . b_j = 2^(e_i+e_j) sum_d 2^(-7d-5) ACC_d[i,j],  ACC_d = sum_{s+t=d+1} q_s . q_t  (int32, exact);
// diagonals d > S are dropped (relative 2^(-7S)).  fp64: S = 7 (~2^-49 of the row scale; 5..8 selectable);
// fp32: S = 4 (28 bits >= the 24-bit significand).
//
// Two kernels:
//  * umma_ozaki_syrk_kernel (v1): one CTA per 128 x 64 output tile, row-major slices behind a 2-D tensor map (64-byte
//    swizzle), Horner drain.  Generic shapes (ragged N, full rectangles of the fp64 debug entry); not on the hot path.
//  * umma_ozaki_syrk_v3_kernel: the production kernel -- persistent or bounded CTAs, warp-specialised (producer / MMA
//    issuer / 4 or 8 epilogue warps), slices in the blocked UMMA layout fetched with bulk copies; see its own header
//    further down.  (The round-1 persistent kernel "v2" was removed in round 2; the tile-walk helpers keep its prefix.)
//   warp 1 issues, for A slice s, one MMA against the STACK of B slices 1..S+1-s (consecutive in smem, so N = (S+1-s)*64
//   and the products land in consecutive TMEM column blocks d = s..S): S+3 MMAs per 32-byte K chunk instead of S(S+1)/2.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "kernels.h"
#include "umma_ozaki.h"

namespace {

constexpr int OZ_BM = 128, OZ_BN = 64, OZ_KB = 64, OZ_STAGES = 2;

// ---------------------------------------------------------------------------------------------
// pre-pass 1: row exponents
// ---------------------------------------------------------------------------------------------
template <typename Tin>
__global__ void __launch_bounds__(256) ozaki_rowscale_kernel(const Tin* __restrict__ P, int64_t lda, int64_t m, int K, int kmajor,
                                                             double* __restrict__ rscale, double* __restrict__ rinv) {
  // 32 rows x 8 column groups per block: every load instruction of a warp covers 32 consecutive rows of one column
  // (row-contiguous operand) or 32 consecutive k of one row (k-major operand); the row maxima meet in shared memory
  __shared__ double part[8][33];
  double mx = 0.0;
  int64_t row;
  if (!kmajor) {
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    row = blockIdx.x * 32ll + x;
    if (row < m) {
      const Tin* p = P + row + (int64_t)y * lda;
#pragma unroll 4
      for (int k = y; k < K; k += 8, p += 8 * lda) mx = fmax(mx, fabs((double)*p));
    }
    part[y][x] = mx;
    __syncthreads();
    if (y != 0) return;
#pragma unroll
    for (int j = 1; j < 8; ++j) mx = fmax(mx, part[j][x]);
  } else {  // element (row, k) at P[k + row * lda]: one warp per row, lanes along k
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int rr = w; rr < 32; rr += 8) {
      const int64_t r2 = blockIdx.x * 32ll + rr;
      double v = 0.0;
      if (r2 < m)
        for (int k = lane; k < K; k += 32) v = fmax(v, fabs((double)P[k + r2 * lda]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0) part[0][rr] = v;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    row = blockIdx.x * 32ll + threadIdx.x;
    mx = part[0][threadIdx.x];
  }
  if (row < m) {
    int e = 0;
    if (mx > 0.0 && isfinite(mx)) frexp(mx, &e);  // mx = f * 2^e, f in [0.5, 1)
    rscale[row] = ldexp(1.0, e);
    rinv[row] = ldexp(1.0, -e);
  }
}

// pre-pass 2: error-free slicing, 16 consecutive k per thread -> one 16-byte store per slice.  Rows land at
// [dst_row0, dst_row0 + m_fill) of the slice buffer (dst_row0 a multiple of 128; rscale / rinv are already offset).
template <int S, typename Tin>
__global__ void ozaki_slice_kernel(const Tin* __restrict__ P, int64_t lda, int64_t m, int64_t m_fill, int64_t m_alloc,
                                   int K, int kmajor, int64_t dst_row0, const double* __restrict__ rinv, int8_t* __restrict__ SL,
                                   int bulk) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int k0 = blockIdx.y * 16;
  if (row >= m_fill) return;
  double r[16];
  const double inv = (row < m) ? rinv[row] : 0.0;
  if (!kmajor) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = (row < m) ? (double)P[row + (int64_t)(k0 + i) * lda] * inv : 0.0;
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = (row < m) ? (double)P[(k0 + i) + row * lda] * inv : 0.0;
  }
  double up = 64.0, dn = 1.0 / 64.0;  // 2^(7s-1), 2^-(7s-1)
  const int64_t drow = row + dst_row0;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    union { int8_t b[16]; uint4 v; } pk;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double q = rint(r[i] * up);
      r[i] = fma(-q, dn, r[i]);
      pk.b[i] = (int8_t)(int)q;
    }
    if (bulk) {
      // UMMA "interleaved" (no-swizzle) K-major layout, written directly: chunk (slice, 128-row block, 32-byte
      // k-block) = 4096 contiguous bytes = [16 row-groups][2 k-halves][8 rows][16 B]  (SBO 256 B, LBO 128 B)
      const int64_t rb = drow >> 7, g = (drow & 127) >> 3, r8 = drow & 7;
      const int kb = k0 >> 5, h = (k0 >> 4) & 1;
      // bulk 1: [slice][row block][k block]; bulk 2 (v3 kernel): [row block][k block][slice] -- the S chunks one
      // (128-row tile, k block) needs are ONE contiguous S*4096-byte run
      const int64_t chunk = (bulk == 2) ? (rb * (K >> 5) + kb) * S + s : ((int64_t)s * (m_alloc >> 7) + rb) * (K >> 5) + kb;
      *reinterpret_cast<uint4*>(SL + chunk * 4096 + ((g * 2 + h) * 8 + r8) * 16) = pk.v;
    } else {
      *reinterpret_cast<uint4*>(SL + ((int64_t)s * m_alloc + drow) * K + k0) = pk.v;
    }
    up *= 128.0;
    dn *= (1.0 / 128.0);
  }
}

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 28)) __trap();  // turn a protocol bug into an error instead of a hang
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand tile, 64-byte swizzle: rows at 64 B pitch, 8-row groups at SBO = 512 B
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                         // LBO (ignored for swizzled K-major) [16,30)
  d |= (uint64_t)(512 >> 4) << 32;                // SBO = 512 B  [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                         // layout type: SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct OzTileArgs {
  void* C; int64_t ldc;   // fp64 (v1, v2, v3) or fp32 (v3<..., float>)
  double sign;            // v3: C += sign * (P P'); -1 for the trailing updates, +1 for accumulations
  int64_t M, N;           // extent of C (rows of P used, columns updated)
  int64_t m_alloc;        // row stride between slices in the slice buffer
  int K;                  // bytes (= elements) per slice row
  const double* rscale;   // 2^e per P row
  int64_t b_tile_stride, b_off;  // column n of C <-> P row  (n / 128) * b_tile_stride + n % 128 + b_off  (0 stride = identity + b_off)
  int64_t a_off;                 // row r of C <-> P row r + a_off
  int64_t b_tile_width;          // distribution block width in columns (0 -> 128)
  const int64_t* strip_start;    // v2, block-cyclic: first tile index of every 64-column strip (nbj + 1 entries)
  const int32_t* strip_bimin;    // v2, block-cyclic: first valid 128-row tile of every strip
  const int8_t* SLb;             // v2 bulk mode: slices in the blocked UMMA layout (nullptr -> tensor-map path)
  int lower_only;
  int gs_shift;                  // v2 <.., GE = 1> only: the tables describe GROUPS of 2^gs_shift strips, tiles row-major inside
  int epi;                       // v2 epilogue variant: 0 Horner over S fp64 terms, 1 int32 pair pre-combination (K <= 512); >=2 PROBE ONLY
};

template <int S>
__global__ void __launch_bounds__(192, 1) umma_ozaki_syrk_kernel(const __grid_constant__ CUtensorMap tmap, OzTileArgs a) {
  constexpr int A_BYTES = OZ_BM * OZ_KB, B_BYTES = OZ_BN * OZ_KB;
  constexpr int STAGE_BYTES = S * (A_BYTES + B_BYTES);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[OZ_STAGES], empty_bar[OZ_STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * OZ_BM, n0 = (int64_t)blockIdx.y * OZ_BN;
  const int64_t bw = a.b_tile_width ? a.b_tile_width : 128;
  const int64_t n_src0 = (a.b_tile_stride ? (n0 / bw) * a.b_tile_stride + (n0 % bw) : n0) + a.b_off;
  const int64_t m_src0 = m0 + a.a_off;
  if (a.lower_only && n_src0 >= m_src0 + OZ_BM) return;  // tile entirely above the diagonal (uniform exit)

  // 1024-byte aligned carve-up (dynamic smem base alignment is only guaranteed to 16 B)
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  auto a_tile = [&](int st, int sl) { return base + st * STAGE_BYTES + sl * A_BYTES; };
  auto b_tile = [&](int st, int sl) { return base + st * STAGE_BYTES + S * A_BYTES + sl * B_BYTES; };

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < OZ_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int num_kb = a.K / OZ_KB;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % OZ_STAGES;
        const uint32_t ph = (kb / OZ_STAGES) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        mbar_expect_tx(&full_bar[st], STAGE_BYTES);
#pragma unroll 1
        for (int sl = 0; sl < S; ++sl) {
          const int rbase = (int)(sl * a.m_alloc);
          tma_load_2d(a_tile(st, sl), &tmap, kb * OZ_KB, rbase + (int)m_src0, &full_bar[st]);
          tma_load_2d(a_tile(st, sl) + 64 * OZ_KB, &tmap, kb * OZ_KB, rbase + (int)m_src0 + 64, &full_bar[st]);
          tma_load_2d(b_tile(st, sl), &tmap, kb * OZ_KB, rbase + (int)n_src0, &full_bar[st]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = S32 (c_format 2 @ bit 4), A/B = INT8 signed (1 @ bits 7, 10), K-major both,
      // N>>3 @ bit 17, M>>4 @ bit 24
      const uint32_t idesc_base = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % OZ_STAGES;
        const uint32_t ph = (kb / OZ_STAGES) & 1;
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(a_tile(st, 0)), b0 = smem_u32(b_tile(st, 0));
#pragma unroll
        for (int kk = 0; kk < OZ_KB / 32; ++kk) {
#pragma unroll 1
          for (int sl = 0; sl < S; ++sl) {
            const int Ns = (S - sl) * OZ_BN;  // B slices 1..S-sl stacked: columns d = sl .. S-1
            const uint64_t adesc = umma_desc_sw64(a0 + sl * A_BYTES + kk * 32);
            for (int c = 0; c < Ns; c += 256) {
              const int nchunk = (Ns - c < 256) ? (Ns - c) : 256;
              const uint64_t bdesc = umma_desc_sw64(b0 + c * OZ_KB + kk * 32);
              const uint32_t idesc = idesc_base | ((uint32_t)(nchunk >> 3) << 17);
              const uint32_t accum = (kb == 0 && kk == 0 && sl == 0) ? 0u : 1u;
              umma_i8(tmem_base + (uint32_t)(sl * OZ_BN + c), adesc, bdesc, idesc, accum);
            }
          }
        }
        umma_commit(&empty_bar[st]);  // smem slot free once these MMAs have read it
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    // ---- epilogue: warps 2..5 own TMEM lanes 32*(warp&3) .. +31
    const int quarter = warp & 3;
    const int64_t row = m0 + 32 * quarter + lane;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const bool row_ok = row < a.M;
    const double rs = row_ok ? a.rscale[row + a.a_off] * (1.0 / 4096.0) : 0.0;  // 2^e_i * 2^-12
#pragma unroll 1
    for (int c = 0; c < OZ_BN; c += 16) {
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.0;
#pragma unroll 1
      for (int d = S - 1; d >= 0; --d) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(d * OZ_BN + c), r);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fma(v[i], 1.0 / 128.0, (double)(int)r[i]);
      }
      if (row_ok) {
        double cv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int64_t col = n0 + c + i;
          cv[i] = (col < a.N) ? ((double*)a.C)[row + col * a.ldc] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int64_t col = n0 + c + i;
          if (col < a.N) ((double*)a.C)[row + col * a.ldc] = fma(-v[i] * rs, a.rscale[n_src0 + c + i], cv[i]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// v2: PERSISTENT, fully overlapped variant for the diagonal-anchored lower-triangular update
// (a_off == b_off, identity column map).  One CTA per SM walks the tile list (row-major over the lower
// triangle, closed-form index -> (bi,bj)); the TMA producer streams 32-byte k-blocks (32 B swizzle,
// 4-5 stages) continuously ACROSS tiles, so the loads of tile t+1 run under the epilogue of tile t; the
// MMA warp re-arms as soon as the epilogue has drained TMEM (tmem_empty barrier).  The epilogue issues all
// S accumulator loads of a 16-column chunk before one wait and streams C with .cs loads/stores so the
// int8 slices stay L2-resident.
// ---------------------------------------------------------------------------------------------
constexpr int V2_KB = 32;
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;  // SBO = 8 rows x 32 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;           // SWIZZLE_32B
  return d;
}
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;  // LBO: the two 16-byte K halves of a core-matrix pair are 128 B apart
  d |= (uint64_t)(256 >> 4) << 32;  // SBO: 8-row groups are 256 B apart
  d |= (uint64_t)1 << 46;           // descriptor version (Blackwell); layout type 0 = interleaved / no swizzle
  return d;
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// cluster variants (AGP_OZAKI_CLUSTER=2): one copy feeds the same smem offset of every CTA in ctaMask and signals
// complete_tx on each destination CTA's own mbarrier at the same offset; the MMA warp's commit arrives on the
// "stage free" barrier of every CTA, because a peer's copy may overwrite this CTA's stage.
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ double ld_cs(const double* p) {
  double v;
  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_cs(double* p, double v) { asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_cs(const float* p) {
  float v;
  asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return (double)v;
}
__device__ __forceinline__ void st_cs(float* p, double v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"((float)v) : "memory"); }

// slot index -> (bi, bj) for the diagonal-anchored lower triangle, L2-BLOCKED: the tile grid is cut into
// super-blocks of SB row tiles x 2*SB column tiles (2048 x 2048 elements for SB = 16); slots walk one
// super-block at a time, so the 148 CTAs share ~15 MB of slices at any moment instead of cycling through the
// whole slice buffer (which is about the size of L2).  Row tile bi owns column tiles 0 .. min(nbj, 2*bi+2)-1;
// slots outside the triangle (in diagonal / edge super-blocks) are reported invalid and skipped by all roles.
constexpr int V2_SB = 16;
__device__ __forceinline__ bool v2_tile(int64_t t, int nbi, int nbj, int& bi, int& bj) {
  constexpr int64_t PER = (int64_t)V2_SB * 2 * V2_SB;
  const int64_t sb = t / PER;
  const int w = (int)(t - sb * PER);
  const int64_t nJ = (nbj + 2 * V2_SB - 1) / (2 * V2_SB);  // super-block columns
  const int64_t t_full = nJ * (nJ + 1) / 2;                // super-block rows I < nJ hold I+1 super-blocks
  int64_t I, Jc;
  if (sb < t_full) {
    I = (int64_t)((sqrt(8.0 * (double)sb + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= sb) ++I;
    while (I * (I + 1) / 2 > sb) --I;
    Jc = sb - I * (I + 1) / 2;
  } else {
    const int64_t r = sb - t_full;
    I = nJ + r / nJ;
    Jc = r % nJ;
  }
  bi = (int)(I * V2_SB + w / (2 * V2_SB));
  bj = (int)(Jc * 2 * V2_SB + w % (2 * V2_SB));
  return bi < nbi && bj < nbj && bj < 2 * bi + 2;
}

// table-driven variant for the block-cyclic (multi-GPU) trailing update: strips are 64 columns wide,
// strip j owns tiles [start[j], start[j+1]) = row tiles bimin[j] ...
__device__ __forceinline__ void v2_tile_tab(int64_t t, int nbj, const int64_t* __restrict__ start,
                                            const int32_t* __restrict__ bimin, int& bi, int& bj) {
  int lo = 0, hi = nbj;  // find the last j with start[j] <= t
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (start[mid] <= t) lo = mid; else hi = mid;
  }
  bj = lo;
  bi = bimin[lo] + (int)(t - start[lo]);
}
// GE = 1 (EXPERIMENTAL, block-cyclic path): the tables hold one entry per GROUP of 2^gs_shift strips (one local
// distribution block); inside a group tiles run row-major -- 2^gs_shift consecutive slots share one 128-row A tile and
// the group's B strips (~2 MB) stay in L2, instead of every slot of a wave asking for a different A tile.
template <int GE>
__device__ __forceinline__ bool v2_decode(const OzTileArgs& a, int64_t t, int nbi, int nbj, int& bi, int& bj, int64_t& brow) {
  if (a.strip_start) {
    if constexpr (GE == 1) {
      const int ng = nbj >> a.gs_shift;
      int lo = 0, hi = ng;  // last group with start[g] <= t
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.strip_start[mid] <= t) lo = mid; else hi = mid;
      }
      const int64_t tl = t - a.strip_start[lo];
      bi = a.strip_bimin[lo] + (int)(tl >> a.gs_shift);
      bj = (lo << a.gs_shift) + (int)(tl & ((1 << a.gs_shift) - 1));
    } else {
      v2_tile_tab(t, nbj, a.strip_start, a.strip_bimin, bi, bj);
    }
    const int64_t n0 = (int64_t)bj * OZ_BN, bw = a.b_tile_width ? a.b_tile_width : 128;
    brow = (a.b_tile_stride ? (n0 / bw) * a.b_tile_stride + (n0 % bw) : n0) + a.b_off;  // stride 0 = identity column map
    return true;
  }
  const bool ok = v2_tile(t, nbi, nbj, bi, bj);
  brow = (int64_t)bj * OZ_BN + a.b_off;
  return ok;
}

// ---------------------------------------------------------------------------------------------
// v3: same algorithm and tile walk as v2, restructured after the round-2 probe (profiles/r02_call1_*): the v2 main loop
// ran at the SAME speed with its operand traffic switched off -- it was bound by the single issuing thread (~38 SASS
// instructions with 7 R2UR.BROADCAST + an ELECT loop per tcgen05.mma because the issue sat in a divergent `lane == 0`
// region), and so was the producer (14 bulk copies per K chunk, each behind its own address arithmetic).  Here
//  * producer and MMA warps run warp-uniform loops and elect one lane only around the instruction itself, so
//    descriptors / addresses live in uniform registers; the S x (S+3)/... MMA sequence is fully unrolled with
//    compile-time instruction descriptors and descriptors formed by ONE add on a per-stage base;
//  * the slices are stored [row block][k block][slice]: the S A-chunks of a stage are one 28 KB bulk copy
//    (CL = 2: two multicast halves), the B chunks S copies of 2 KB -- 8 copies per stage instead of 14;
//  * the drain combines the S int32 accumulators exactly in int64 (two words: 4 + (S-4) terms), converts them with the
//    2^52 magic-number add instead of I2F.F64 and rounds ONCE (fma) -- 5 fp64-pipe operations per element, was 10.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ double i64_to_f64_exact(long long x) {  // |x| < 2^51: exact, one integer add + one DADD
  return __longlong_as_double(x + 0x4338000000000000LL) - 6755399441055744.0;
}
// exact value of sum_d acc_d 128^(3-d) as two int64 words (hi: d = 0..3, lo: d = 4..S-1, scaled by 128^(S-4)), then ONE
// rounding.  PAIR32: adjacent accumulators are first combined in int32 (valid for K <= 512, see v2).
template <int S, bool PAIR32>
__device__ __forceinline__ double oz_combine(const uint32_t (&r)[S][8], int i) {
  long long h = 0, l = 0;
  if constexpr (S <= 4) {  // fp32 operands (3 or 4 slices): everything fits one word, the conversion is the only rounding
    h = (int)r[0][i];
#pragma unroll
    for (int d = 1; d < S; ++d) h = h * 128 + (int)r[d][i];
#pragma unroll
    for (int d = S; d < 4; ++d) h *= 128;  // same 128^3 scaling as the longer splits
    return i64_to_f64_exact(h);
  } else if constexpr (PAIR32) {
    const int t01 = (int)r[0][i] * 128 + (int)r[1][i], t23 = (int)r[2][i] * 128 + (int)r[3][i];
    h = (long long)t01 * 16384 + t23;
    if constexpr (S == 5) l = (int)r[4][i];
    else if constexpr (S == 6) l = (int)r[4][i] * 128 + (int)r[5][i];
    else if constexpr (S == 7) l = (long long)((int)r[4][i] * 128 + (int)r[5][i]) * 128 + (int)r[6][i];
    else l = (long long)((int)r[4][i] * 128 + (int)r[5][i]) * 16384 + ((int)r[6][i] * 128 + (int)r[7][i]);
  } else {
    h = (((long long)(int)r[0][i] * 128 + (int)r[1][i]) * 128 + (int)r[2][i]) * 128 + (int)r[3][i];
    l = (int)r[4][i];
#pragma unroll
    for (int d = 5; d < S; ++d) l = l * 128 + (int)r[d][i];
  }
  constexpr double LO_SCALE = (S <= 5) ? 1.0 / 128.0 : (S == 6) ? 1.0 / 16384.0 : (S == 7) ? 1.0 / 2097152.0 : 1.0 / 268435456.0;
  return fma(i64_to_f64_exact(l), LO_SCALE, i64_to_f64_exact(h));
}

template <int S, int CL, int NEPI, int GE, typename CT>
__global__ void __launch_bounds__(64 + 32 * NEPI, 1) umma_ozaki_syrk_v3_kernel(OzTileArgs a, int64_t ntiles, int nbi, int nbj,
                                                                                int tpc) {
  // tpc = 0: persistent, CTA b walks slots b, b + grid, ...   tpc > 0: BOUNDED CTAs -- CTA b owns the tpc consecutive slots
  // [b * tpc, (b + 1) * tpc) and exits; the grid is ceil(ntiles / tpc).  Bounded CTAs hand their SM back every ~0.1 ms, so
  // kernels of a higher-priority stream (the panel chain, the NCCL broadcast) are scheduled between them instead of
  // waiting for the whole update.
  const int64_t t_begin = tpc ? (int64_t)blockIdx.x * tpc : (int64_t)blockIdx.x;
  const int64_t t_end = tpc ? ((t_begin + tpc < ntiles) ? t_begin + tpc : ntiles) : ntiles;
  const int64_t t_step = tpc ? 1 : (int64_t)gridDim.x;
  constexpr int A_BYTES = OZ_BM * V2_KB, B_BYTES = OZ_BN * V2_KB;
  constexpr int STAGE_BYTES = S * (A_BYTES + B_BYTES);
  constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CL); }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar, NEPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (CL > 1) cluster_sync_all();
  const uint32_t tmem_base = tmem_base_s;
  const int num_kb = a.K / V2_KB;
  constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1);

  if (warp == 0) {
    // ---- producer: all 32 lanes walk the loop (uniform control flow), one elected lane issues the copies
    uint32_t it = 0;
    const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
    const int64_t rb_bytes = (int64_t)num_kb * S * 4096;  // bytes of one 128-row block (all k blocks, all slices)
    for (int64_t t = t_begin; t < t_end; t += t_step) {
      int bi, bj;
      int64_t brow64;
      if (!v2_decode<GE>(a, t, nbi, nbj, bi, bj, brow64)) continue;
      const int64_t arow = (int64_t)bi * OZ_BM + a.a_off;
      const int8_t* asrc = a.SLb + (arow >> 7) * rb_bytes;
      const int8_t* bsrc = a.SLb + (brow64 >> 7) * rb_bytes + (brow64 & 64) * 32;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        if (a.epi == 5) {  // PROBE 5: MMA-only loop, no operand traffic
          if (elect_one()) mbar_arrive(&full_bar[st]);
        } else if (elect_one()) {
          uint8_t* dst = base + st * STAGE_BYTES;
          mbar_expect_tx(&full_bar[st], STAGE_BYTES);
          if constexpr (CL > 1) {  // each CTA fetches 1/CL of the contiguous A run and multicasts it to the cluster
            constexpr int PART = S * A_BYTES / CL;
            bulk_load_mc(dst + crank * PART, asrc + crank * PART, PART, &full_bar[st], CL_MASK);
          } else {
            bulk_load(dst, asrc, S * A_BYTES, &full_bar[st]);
          }
#pragma unroll
          for (int sl = 0; sl < S; ++sl) bulk_load(dst + S * A_BYTES + sl * B_BYTES, bsrc + sl * 4096, B_BYTES, &full_bar[st]);
        }
        __syncwarp();
        asrc += S * 4096;
        bsrc += S * 4096;
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: uniform loop, descriptors = per-stage base + compile-time offsets
    constexpr uint32_t IDESC_BASE = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
    constexpr uint64_t DESC_HI = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);  // LBO, SBO, version
    const uint32_t stage0_lo = (smem_u32(base) & 0x3FFFF) >> 4;
    uint32_t it = 0, lt = 0;
    for (int64_t t = t_begin; t < t_end; t += t_step) {
      {
        int bi_, bj_;
        int64_t br_;
        if (!v2_decode<GE>(a, t, nbi, nbj, bi_, bj_, br_)) continue;
      }
      mbar_wait(&tmem_empty_bar, (lt & 1) ^ 1);  // epilogue has drained the previous tile's accumulators
      tc_fence_after();
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = stage0_lo + st * (uint32_t)(STAGE_BYTES >> 4);
          const uint32_t b_lo = a_lo + (uint32_t)((S * A_BYTES) >> 4);
          const uint32_t acc0 = (kb != 0) ? 1u : 0u;
          if (a.epi != 6) {  // PROBE 6: operand traffic only
#pragma unroll
            for (int sl = 0; sl < S; ++sl) {
#pragma unroll
              for (int c = 0; c < (S - sl) * OZ_BN; c += 256) {
                const int nchunk = ((S - sl) * OZ_BN - c < 256) ? ((S - sl) * OZ_BN - c) : 256;
                umma_i8(tmem_base + (uint32_t)(sl * OZ_BN + c), DESC_HI | (uint64_t)(a_lo + (uint32_t)(sl * (A_BYTES >> 4))),
                        DESC_HI | (uint64_t)(b_lo + (uint32_t)((c * V2_KB) >> 4)), IDESC_BASE | ((uint32_t)(nchunk >> 3) << 17),
                        (sl == 0) ? acc0 : 1u);
              }
            }
          }
          if constexpr (CL > 1) umma_commit_mc(&empty_bar[st], CL_MASK);
          else umma_commit(&empty_bar[st]);
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar);
        }
        __syncwarp();
      }
      ++lt;
    }
  } else {
    const int quarter = warp & 3;
    constexpr int CB = OZ_BN * 4 / NEPI;                          // columns drained by one epilogue warp
    const int c0 = (NEPI == 4) ? 0 : ((warp - 2) >> 2) * CB;      // warps 2-5: columns [0, CB), warps 6-9: [CB, 2 CB)
    const bool pair32 = (a.epi == 1);
    uint32_t lt = 0;
    for (int64_t t = t_begin; t < t_end; t += t_step) {
      int bi, bj;
      int64_t brow64;
      if (!v2_decode<GE>(a, t, nbi, nbj, bi, bj, brow64)) continue;
      const int64_t m0 = (int64_t)bi * OZ_BM, n0 = (int64_t)bj * OZ_BN + c0;
      brow64 += c0;
      const int64_t row = m0 + 32 * quarter + lane;
      const bool row_ok = row < a.M;
      const double rs = row_ok ? a.sign * a.rscale[row + a.a_off] * (1.0 / 8589934592.0) : 0.0;  // +-2^e_i * 2^-12 * 128^-3
      CT* crow = (CT*)a.C + row;
      mbar_wait(&tmem_full_bar, lt & 1);
      tc_fence_after();
      double v[CB];
      if (a.epi >= 3 && a.epi != 7) {  // PROBE: no drain (3, 4, 5, 6) -- results are garbage, timing only
#pragma unroll
        for (int i = 0; i < CB; ++i) v[i] = 0.0;
      } else {
#pragma unroll
        for (int c8 = 0; c8 < CB; c8 += 8) {
          uint32_t r[S][8];
#pragma unroll
          for (int d = 0; d < S; ++d)
            tmem_ld8_nowait(tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(d * OZ_BN + c0 + c8), r[d]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (a.epi == 7) {  // PROBE 7: TMEM reads only
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t x = r[0][i];
#pragma unroll
              for (int d = 1; d < S; ++d) x ^= r[d][i];
              v[c8 + i] = __hiloint2double(0x43300000, (int)x);
            }
          } else if (pair32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[c8 + i] = oz_combine<S, true>(r, i);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[c8 + i] = oz_combine<S, false>(r, i);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar);
      // C += sign * (2^e_i 2^e_j 2^-33) * v, streamed (.cs) so the int8 slices stay resident in L2
      if (row_ok && a.epi != 2 && a.epi != 3 && a.epi != 5 && a.epi != 6) {
#pragma unroll
        for (int c = 0; c < CB; c += 16) {
          double cv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int64_t col = n0 + c + i;
            cv[i] = (col < a.N) ? ld_cs(crow + col * a.ldc) : 0.0;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int64_t col = n0 + c + i;
            if (col < a.N) st_cs(crow + col * a.ldc, fma(v[c + i], rs * a.rscale[brow64 + c + i], cv[i]));
          }
        }
      }
      ++lt;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

template <int S, int CL, int NEPI, int GE, typename CT = double>
void launch_v3_variant(const OzTileArgs& a, int64_t ntiles, int nbi, int nbj, int cap, size_t smem, cudaStream_t s, int tpc = 0) {
  static int max_clusters[64] = {0};  // per device; 0 = not queried yet
  static uint64_t configured = 0;
  constexpr unsigned THREADS = 64 + 32 * NEPI;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaLaunchConfig_t lc{};
  lc.blockDim = dim3(THREADS); lc.dynamicSmemBytes = smem; lc.stream = s;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeClusterDimension;
  la[0].val.clusterDim.x = CL; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
  lc.attrs = la; lc.numAttrs = 1;
  if (agp_first_use_on_device(&configured)) {
    cudaFuncSetAttribute(umma_ozaki_syrk_v3_kernel<S, CL, NEPI, GE, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int mc = 0;
    if (CL > 1) {
      lc.gridDim = dim3((unsigned)(cap / CL * CL));
      if (cudaOccupancyMaxActiveClusters(&mc, umma_ozaki_syrk_v3_kernel<S, CL, NEPI, GE, CT>, &lc) != cudaSuccess) { mc = 0; cudaGetLastError(); }
    } else {
      mc = 1 << 20;
    }
    max_clusters[dev & 63] = mc;
  }
  int64_t grid = (int64_t)max_clusters[dev & 63] * CL;
  if (grid > cap) grid = cap / CL * CL;
  if (grid > ntiles) grid = ntiles / CL * CL;  // the slot count is even when CL = 2 is selected
  if (CL > 1) tpc = 0;  // CTA pairs walk the slots in lockstep: persistent form only
  if (tpc > 0) grid = (ntiles + tpc - 1) / tpc;
  if (grid <= 0) return;
  lc.gridDim = dim3((unsigned)grid);
  cudaLaunchKernelEx(&lc, umma_ozaki_syrk_v3_kernel<S, CL, NEPI, GE, CT>, a, ntiles, nbi, nbj, tpc);
}

template <int S, typename CT = double>
void launch_syrk_v2_S(const OzakiWs& ws, void* C, int64_t ldc, int64_t M, int64_t N, int64_t b_tile_stride,
                      int64_t b_tile_width, int64_t b_off, int64_t a_off, cudaStream_t s, int full = 0, double sign = -1.0) {
  constexpr int STAGE_BYTES = S * (OZ_BM * V2_KB + OZ_BN * V2_KB);
  constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  static uint64_t configured = 0;  // per-device bit: the attribute is per device (one ctx per GPU in one process)
  static int nsm = 148;
  if (agp_first_use_on_device(&configured)) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  // EXPERIMENTAL switches (the variants compile, none has run on a device yet; the default <S, 1, 4, 0> kernel is the
  // validated one): AGP_OZAKI_CLUSTER=2 -> A-multicast CTA pairs, AGP_OZAKI_EPIWARPS=8 -> two epilogue warps per quarter,
  // AGP_OZAKI_GROUPED=1 -> block-cyclic (multi-GPU) tile order grouped by distribution block, row-major inside
  int want_cl = 1, want_ew = 4, want_ge = 0;
  {
    const char* e = getenv("AGP_OZAKI_CLUSTER");
    want_cl = (e && atoi(e) == 2) ? 2 : 1;
    const char* f = getenv("AGP_OZAKI_EPIWARPS");
    want_ew = (f && atoi(f) == 8) ? 8 : 4;
    const char* g = getenv("AGP_OZAKI_GROUPED");
    want_ge = (g && atoi(g) == 1) ? 1 : 0;
  }
  OzTileArgs a{};
  a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.m_alloc = ws.m_alloc; a.K = ws.K; a.rscale = ws.rscale;
  a.b_tile_stride = b_tile_stride; a.b_tile_width = b_tile_width; a.b_off = b_off; a.a_off = a_off; a.lower_only = 1;
  a.SLb = ws.bulk ? ws.SL : nullptr;
  a.sign = sign;
  if (full) want_ge = 0;
  {
    // default: the int32 pair pre-combination where it was measured and validated on the device (S = 7, K <= 512:
    // +4.4% kernel throughput, sampled output bit-identical to variant 0 -- profiles/r01_ozaki_probe.json);
    // AGP_OZAKI_EPI=0 restores the plain Horner drain, >= 2 are the timing-only variants of tools/ozaki_probe.py.
    const char* e = getenv("AGP_OZAKI_EPI");
    a.epi = e ? atoi(e) : 1;
    if (a.epi == 1 && ws.K > 512) a.epi = 0;  // the int32 pair bound needs K <= 512
  }
  const int nbi = (int)((M + OZ_BM - 1) / OZ_BM), nbj = (int)(N / OZ_BN);
  int64_t ntiles = 0;
  if (!full && b_tile_stride == 0 && a_off == b_off) {  // diagonal-anchored: closed-form, L2-blocked slot enumeration
    const int64_t nJ = (nbj + 2 * V2_SB - 1) / (2 * V2_SB), nI = (nbi + V2_SB - 1) / V2_SB;
    const int64_t nsb = (nI <= nJ) ? nI * (nI + 1) / 2 : nJ * (nJ + 1) / 2 + (nI - nJ) * nJ;
    ntiles = nsb * (int64_t)V2_SB * 2 * V2_SB;  // slots, including the skipped ones of diagonal / edge super-blocks
  } else {  // block-cyclic column map: per-strip table (host -> device, a few KB)
    std::vector<int64_t> start((size_t)nbj + 1);
    std::vector<int32_t> bimin((size_t)nbj);
    const int64_t bw = b_tile_width ? b_tile_width : 128;
    // grouped order: one table entry per distribution block (bw / 64 strips, a power of two), its first strip's bimin for
    // all of them -- inside the block on the diagonal that computes a few tiles above the diagonal (harmless: they land
    // in the unused upper triangle of the local block) in exchange for a uniform row-major walk
    int gs_shift = 0;
    if (want_ge) {
      const int64_t gs = bw / OZ_BN;
      while ((1ll << gs_shift) < gs) ++gs_shift;
      if ((1ll << gs_shift) != gs || gs < 2 || nbj % gs != 0 || !b_tile_stride) want_ge = 0;
    }
    if (want_ge) {
      const int gs = 1 << gs_shift, ng = nbj / gs;
      for (int g = 0; g < ng; ++g) {
        const int64_t n0 = (int64_t)g * gs * OZ_BN;
        const int64_t nsrc = (n0 / bw) * b_tile_stride + (n0 % bw) + b_off;
        int64_t bm = (nsrc - a_off) >= 0 ? (nsrc - a_off) / OZ_BM : 0;
        if (bm > nbi) bm = nbi;
        bimin[g] = (int32_t)bm;
        start[g] = ntiles;
        ntiles += (nbi - bm) * gs;
      }
      start[ng] = ntiles;
      a.gs_shift = gs_shift;
    } else
    for (int j = 0; j < nbj; ++j) {
      const int64_t n0 = (int64_t)j * OZ_BN;
      const int64_t nsrc = (b_tile_stride ? (n0 / bw) * b_tile_stride + (n0 % bw) : n0) + b_off;
      int64_t bm = (nsrc - a_off) >= 0 ? (nsrc - a_off) / OZ_BM : 0;  // first row tile with nsrc < a_off + bi*128 + 128
      if (full) bm = 0;  // rectangular product: every row tile of every strip
      if (bm > nbi) bm = nbi;
      bimin[j] = (int32_t)bm;
      start[j] = ntiles;
      ntiles += nbi - bm;
    }
    start[nbj] = ntiles;
    const int slot = (ws.tab_slot++) & 1;  // the main- and side-stream updates of one step are in flight together
    int64_t* d_start = ws.tab_start + (size_t)slot * (ws.tab_cap + 1);
    int32_t* d_bimin = ws.tab_bimin + (size_t)slot * (ws.tab_cap + 1);
    cudaMemcpyAsync(d_start, start.data(), ((size_t)nbj + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_bimin, bimin.data(), (size_t)nbj * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    a.strip_start = d_start;
    a.strip_bimin = d_bimin;
  }
  if (ntiles <= 0) return;
  const int cap = (ws.max_ctas > 0 && ws.max_ctas < nsm) ? ws.max_ctas : nsm;
  const bool ge = want_ge && a.strip_start;
  {  // 8 epilogue warps unless AGP_OZAKI_EPIWARPS=4; CTA pairs with AGP_OZAKI_CLUSTER=2
    const char* f = getenv("AGP_OZAKI_EPIWARPS");
    const int ew = (f && atoi(f) == 4) ? 4 : 8;
    const bool cl2 = want_cl == 2 && (!a.strip_start || ge) && ntiles >= 2;
    if constexpr (std::is_same<CT, double>::value && S >= 5) {
      if (ge && cl2 && ew == 8) launch_v3_variant<S, 2, 8, 1>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
      else if (ge && cl2) launch_v3_variant<S, 2, 4, 1>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
      else if (ge && ew == 8) launch_v3_variant<S, 1, 8, 1>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
      else if (ge) launch_v3_variant<S, 1, 4, 1>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
      else if (cl2 && ew == 8) launch_v3_variant<S, 2, 8, 0>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
      else if (cl2) launch_v3_variant<S, 2, 4, 0>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
      else if (ew == 8) launch_v3_variant<S, 1, 8, 0>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
      else launch_v3_variant<S, 1, 4, 0>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
    } else {  // fp32 output and / or short (3-, 4-slice) splits: the two main variants only
      if (cl2) launch_v3_variant<S, 2, 8, 0, CT>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
      else launch_v3_variant<S, 1, 8, 0, CT>(a, ntiles, nbi, nbj, cap, smem, s, ws.chunk_tiles);
    }
    agp_count_launch();
  }
}

template <int S>
void launch_syrk_S(const OzakiWs& ws, double* C, int64_t ldc, int64_t M, int64_t N, int lower_only, int64_t b_tile_stride,
                   int64_t b_tile_width, int64_t b_off, int64_t a_off, cudaStream_t s) {
  if (ws.bulk == 2 && lower_only && N % 128 == 0 && N >= 128 && N / OZ_BN <= ws.tab_cap) {
    launch_syrk_v2_S<S>(ws, C, ldc, M, N, b_tile_stride, b_tile_width, b_off, a_off, s);
    return;
  }
  const size_t smem = (size_t)OZ_STAGES * S * (OZ_BM * OZ_KB + OZ_BN * OZ_KB) + 1024;
  static uint64_t configured = 0;  // per-device bit: the attribute is per device (one ctx per GPU in one process)
  if (agp_first_use_on_device(&configured)) {
    cudaFuncSetAttribute(umma_ozaki_syrk_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  OzTileArgs a{};
  a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.m_alloc = ws.m_alloc; a.K = ws.K; a.rscale = ws.rscale;
  a.b_tile_stride = b_tile_stride; a.b_tile_width = b_tile_width; a.b_off = b_off; a.a_off = a_off; a.lower_only = lower_only;
  dim3 grid((unsigned)((M + OZ_BM - 1) / OZ_BM), (unsigned)((N + OZ_BN - 1) / OZ_BN));
  umma_ozaki_syrk_kernel<S><<<grid, 192, smem, s>>>(ws.tmap, a);
  agp_count_launch();
}

}  // namespace

int ozaki_ws_create(OzakiWs* ws, int64_t max_rows, int K, int S, cudaStream_t s) {
  memset(ws, 0, sizeof(*ws));
  if (S < 3 || S > 8 || K % OZ_KB != 0) return 1;
  EncodeTiledFn enc = get_encode();
  if (!enc) return 2;
  ws->m_alloc = (max_rows + 127) / 128 * 128;
  ws->K = K; ws->S = S;
  if (cudaMallocAsync((void**)&ws->SL, (size_t)S * ws->m_alloc * K, s) != cudaSuccess) return 3;
  if (cudaMallocAsync((void**)&ws->rscale, (size_t)ws->m_alloc * 2 * sizeof(double), s) != cudaSuccess) return 3;
  ws->rinv = ws->rscale + ws->m_alloc;
  ws->tab_cap = (int)(ws->m_alloc / OZ_BN) + 2;
  if (cudaMallocAsync((void**)&ws->tab_start, (size_t)2 * (ws->tab_cap + 1) * sizeof(int64_t), s) != cudaSuccess) return 3;
  if (cudaMallocAsync((void**)&ws->tab_bimin, (size_t)2 * (ws->tab_cap + 1) * sizeof(int32_t), s) != cudaSuccess) return 3;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)((int64_t)S * ws->m_alloc)};
  cuuint64_t gstr[1] = {(cuuint64_t)K};
  cuuint32_t box[2] = {(cuuint32_t)OZ_KB, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&ws->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ws->SL, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 4;
  ws->bulk = 2;  // [row block][k block][slice] layout of the persistent kernel; the debug entry clears it for generic shapes
  { const char* ct = getenv("AGP_OZAKI_CHUNK_TEST"); ws->chunk_tiles = ct ? atoi(ct) : 0; }  // tests: bounded CTAs everywhere
  return 0;
}

void ozaki_ws_destroy(OzakiWs* ws, cudaStream_t s) {
  if (ws->SL) cudaFreeAsync(ws->SL, s);
  if (ws->rscale) cudaFreeAsync(ws->rscale, s);
  if (ws->tab_start) cudaFreeAsync(ws->tab_start, s);
  if (ws->tab_bimin) cudaFreeAsync(ws->tab_bimin, s);
  memset(ws, 0, sizeof(*ws));
}

template <typename Tin>
static void prepare_t(const OzakiWs& ws, const Tin* P, int kmajor, int64_t lda, int64_t m, int64_t dst_row0, cudaStream_t s) {
  double* rs = ws.rscale + dst_row0;
  double* ri = ws.rinv + dst_row0;
  ozaki_rowscale_kernel<Tin><<<(unsigned)((m + 31) / 32), 256, 0, s>>>(P, lda, m, ws.K, kmajor, rs, ri);
  agp_count_launch();
  const int64_t m_used = (m + 127) / 128 * 128;  // zero-fill up to the tile edge
  dim3 grid((unsigned)((m_used + 127) / 128), (unsigned)(ws.K / 16));
#define AGP_SLICE(SS) ozaki_slice_kernel<SS, Tin><<<grid, 128, 0, s>>>(P, lda, m, m_used, ws.m_alloc, ws.K, kmajor, dst_row0, ri, ws.SL, ws.bulk)
  switch (ws.S) {
    case 3: AGP_SLICE(3); break;
    case 4: AGP_SLICE(4); break;
    case 5: AGP_SLICE(5); break;
    case 6: AGP_SLICE(6); break;
    case 7: AGP_SLICE(7); break;
    default: AGP_SLICE(8); break;
  }
#undef AGP_SLICE
  agp_count_launch();
}

void ozaki_prepare(const OzakiWs& ws, const double* P, int64_t lda, int64_t m, cudaStream_t s) {
  if (m <= 0) return;
  prepare_t<double>(ws, P, 0, lda, m, 0, s);
}

void ozaki_prepare_ex(const OzakiWs& ws, const void* P, int p_is_float, int kmajor, int64_t lda, int64_t m, int64_t dst_row0,
                      cudaStream_t s) {
  if (m <= 0) return;
  if (p_is_float) prepare_t<float>(ws, (const float*)P, kmajor, lda, m, dst_row0, s);
  else prepare_t<double>(ws, (const double*)P, kmajor, lda, m, dst_row0, s);
}

int ozaki_update_ex(const OzakiWs& ws, void* C, int c_is_float, int64_t ldc, int64_t M, int64_t N, int full, double sign,
                    int64_t b_tile_stride, int64_t b_tile_width, int64_t b_off, int64_t a_off, cudaStream_t s) {
  if (M <= 0 || N <= 0) return 0;
  if (ws.bulk != 2 || N % 128 != 0 || N / OZ_BN > ws.tab_cap) return 1;  // v3 kernel + interleaved slice layout only
  if (ws.K > 32768) return 1;  // int32 accumulators ((d+1) K 64^2 < 2^31) and the 2^51 range of the exact int64 -> fp64 drain
#define AGP_UPD(SS, CTT) launch_syrk_v2_S<SS, CTT>(ws, C, ldc, M, N, b_tile_stride, b_tile_width, b_off, a_off, s, full, sign)
  if (c_is_float) {
    switch (ws.S) {
      case 3: AGP_UPD(3, float); break;
      case 4: AGP_UPD(4, float); break;
      case 5: AGP_UPD(5, float); break;
      default: return 1;
    }
  } else {
    switch (ws.S) {
      case 4: AGP_UPD(4, double); break;
      case 5: AGP_UPD(5, double); break;
      case 6: AGP_UPD(6, double); break;
      case 7: AGP_UPD(7, double); break;
      case 8: AGP_UPD(8, double); break;
      default: return 1;
    }
  }
#undef AGP_UPD
  return 0;
}

void ozaki_syrk(const OzakiWs& ws, double* C, int64_t ldc, int64_t M, int64_t N, int lower_only, int64_t b_tile_stride,
                int64_t b_tile_width, int64_t b_off, int64_t a_off, cudaStream_t s) {
  if (M <= 0 || N <= 0) return;
  switch (ws.S) {
    case 5: launch_syrk_S<5>(ws, C, ldc, M, N, lower_only, b_tile_stride, b_tile_width, b_off, a_off, s); break;
    case 6: launch_syrk_S<6>(ws, C, ldc, M, N, lower_only, b_tile_stride, b_tile_width, b_off, a_off, s); break;
    case 7: launch_syrk_S<7>(ws, C, ldc, M, N, lower_only, b_tile_stride, b_tile_width, b_off, a_off, s); break;
    default: launch_syrk_S<8>(ws, C, ldc, M, N, lower_only, b_tile_stride, b_tile_width, b_off, a_off, s); break;
  }
}

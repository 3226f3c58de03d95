// potrf.cu -- K3: Cholesky of one 128x128 diagonal block, entirely in shared memory, fused with the
// inverse of the factor (used by the TRSM-as-GEMM panel solve and the blocked triangular solves).
// Replaces the diagonal-block part of LAPACK potrf inside cholesky(_symmetric(C))
// (/root/reference/src/finite_gp_projection.jl:308, /root/reference/src/exact_gpr_posterior.jl:31).
//
// Scheme: the lower triangle of the smem array holds A -> L; the STRICT UPPER triangle holds the rows
// of E = I * L^-T that a bordered elimination produces for free (E(r,c) lives at position (c,r)), its
// diagonal (1/L_jj) in dinv[].  16 micro-panels of 8 columns: (1) the 8 lanes owning the diagonal
// 8x8 rows factor it in registers with width-8 warp shuffles (no block barrier on the column chain),
// (2) every other active row (below, and the E rows) does an 8-step substitution against it,
// (3) rank-8 trailing update, one thread per (row, column-parity).  3 block barriers per micro-panel.
#include "kernels.h"
#include "agp.h"

namespace {
constexpr int PB = AGP_TILE;  // 128
constexpr int PLD = PB + 1;   // odd leading dimension -> conflict-free column/row access

template <typename T> __device__ __forceinline__ T dev_rsqrt_refined(T p);
template <> __device__ __forceinline__ double dev_rsqrt_refined<double>(double p) {
  double r = rsqrt(p);
  return fma(r * 0.5, fma(-p * r, r, 1.0), r);  // one Newton step -> < 1 ulp
}
template <> __device__ __forceinline__ float dev_rsqrt_refined<float>(float p) {
  float r = rsqrtf(p);
  return fmaf(r * 0.5f, fmaf(-p * r, r, 1.0f), r);
}

template <typename T> __device__ __forceinline__ void load8(const T* p, T* o);
template <> __device__ __forceinline__ void load8<double>(const double* p, double* o) {
  const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { double2 v = q[i]; o[2 * i] = v.x; o[2 * i + 1] = v.y; }
}
template <> __device__ __forceinline__ void load8<float>(const float* p, float* o) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 2; ++i) { float4 v = q[i]; o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w; }
}

template <typename T>
__global__ void __launch_bounds__(512, 1)
potrf_diag_kernel(T* __restrict__ A, int64_t lda, T* __restrict__ Dinv, double* __restrict__ logdet_part,
                  int blk, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* arr = reinterpret_cast<T*>(smem_raw);  // arr[c*PLD + i]
  T* P8 = arr + PB * PLD + ((PB * PLD) & 1);  // keep 16B alignment for T=double (PB*PLD even anyway)
  T* Ld = P8 + PB * 8;                        // Ld[row*8 + col]
  T* dinv = Ld + 64;
  __shared__ double red[4];
  const int tid = threadIdx.x;
  const int slot = tid & 255, half = tid >> 8;

  for (int idx = tid; idx < PB * PB; idx += 512) {
    int c = idx >> 7, i = idx & 127;
    arr[c * PLD + i] = (i >= c) ? A[i + (int64_t)c * lda] : (T)0;
  }
  __syncthreads();

  for (int j0 = 0; j0 < PB; j0 += 8) {
    // ---- (1) diagonal 8x8 in registers (whole warp executes; only the owning 8-lane group is real)
    if (tid < PB && (tid >> 5) == (j0 >> 5)) {
      const int i = tid, lane8 = i & 7;
      const bool mine = ((i & ~7) == j0);
      T a[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) a[c] = arr[(j0 + c) * PLD + i];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        T piv = __shfl_sync(0xffffffffu, a[c], c, 8);
        const bool bad = !(piv > (T)0);
        if (bad) {
          if (mine && lane8 == c) atomicCAS(info, 0, blk * PB + j0 + c + 1);
          piv = (T)1;
        }
        const T r = dev_rsqrt_refined<T>(piv);
        a[c] *= r;
        if (mine && lane8 == c) { dinv[j0 + c] = r; a[c] = piv * r; }
#pragma unroll
        for (int c2 = c + 1; c2 < 8; ++c2) {
          T l = __shfl_sync(0xffffffffu, a[c], c2, 8);
          if (lane8 >= c2) a[c2] -= a[c] * l;
        }
      }
      if (mine) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c <= lane8) { arr[(j0 + c) * PLD + i] = a[c]; Ld[lane8 * 8 + c] = a[c]; }
      }
    }
    __syncthreads();
    // ---- (2) substitution of every other active row against the 8x8 factor
    const bool is_s = slot < PB;
    const int row = is_s ? slot : slot - PB;
    const bool active = is_s ? (row >= j0 + 8) : (row < j0 + 8);
    T x[8];
    if (active) {
      T a[8];
      if (is_s || row < j0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = arr[(j0 + c) * PLD + row];
      } else {
        const int cr = row - j0;
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = (c == cr) ? (T)1 : (T)0;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        T s = a[c];
#pragma unroll
        for (int c2 = 0; c2 < c; ++c2) s -= x[c2] * Ld[c * 8 + c2];
        x[c] = s * dinv[j0 + c];
      }
      if (half == 0) {
        if (is_s) {
#pragma unroll
          for (int c = 0; c < 8; ++c) { arr[(j0 + c) * PLD + row] = x[c]; P8[row * 8 + c] = x[c]; }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (j0 + c > row) arr[(j0 + c) * PLD + row] = x[c];
        }
      }
    }
    __syncthreads();
    // ---- (3) rank-8 trailing update
    if (active) {
      const int kend = is_s ? row : (PB - 1);
      for (int k = j0 + 8 + half; k <= kend; k += 2) {
        T lk[8];
        load8<T>(P8 + k * 8, lk);
        T v = arr[k * PLD + row];
#pragma unroll
        for (int c = 0; c < 8; ++c) v -= x[c] * lk[c];
        arr[k * PLD + row] = v;
      }
    }
    __syncthreads();
  }

  // ---- write back L (upper zeroed) and Dinv = inv(L) (lower, col-major)
  for (int idx = tid; idx < PB * PB; idx += 512) {
    int c = idx >> 7, i = idx & 127;  // (row i, col c)
    A[i + (int64_t)c * lda] = (i >= c) ? arr[c * PLD + i] : (T)0;
  }
  for (int idx = tid; idx < PB * PB; idx += 512) {
    int r = idx >> 7, c = idx & 127;  // Dinv(c, r) = E(r, c)
    T v = (c > r) ? arr[c * PLD + r] : ((c == r) ? dinv[r] : (T)0);
    Dinv[c + r * PB] = v;
  }
  // ---- logdet partial: sum_j log L_jj = - sum_j log dinv_j
  double part = 0.0;
  if (tid < PB) part = -log((double)dinv[tid]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (tid < PB && (tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) logdet_part[blk] = red[0] + red[1] + red[2] + red[3];
}
}  // namespace

template <typename T>
void launch_potrf_diag(T* Ablk, int64_t lda, T* Dinv, double* logdet_part, int blk, int* info, cudaStream_t s) {
  const size_t smem = (size_t)(PB * PLD + 1 + PB * 8 + 64 + PB) * sizeof(T);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(potrf_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  potrf_diag_kernel<T><<<1, 512, smem, s>>>(Ablk, lda, Dinv, logdet_part, blk, info);
  agp_count_launch();
}
template void launch_potrf_diag<float>(float*, int64_t, float*, double*, int, int*, cudaStream_t);
template void launch_potrf_diag<double>(double*, int64_t, double*, double*, int, int*, cudaStream_t);

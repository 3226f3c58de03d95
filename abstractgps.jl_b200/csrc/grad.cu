// grad.cu -- EXPERIMENTAL (compiles, NOT yet run on a device): the fused reduction of the logpdf gradient,
//   dL/dtheta = 1/2 sum_ij W_ij dC_ij/dtheta,   W = alpha alpha' - C^-1,
// what Zygote produces through the reference for `logpdf(f(x, s2), y)` (/root/reference/test/finite_gp_projection.jl:152-178,
// /root/reference/examples/1-mauna-loa/script.jl:200-242).
//
// One CTA per 64x64 tile of the LOWER triangle (same tiling and the same direct-difference distances as gram_kernel):
// the tile's kappa and kappa'(r) r are recomputed from the transformed points, W_ij is read once from the C^-1 buffer
// (N^2/2 elements -- the kernel is bound by that read at small D and by the fp64 pipe at D = 64), off-diagonal elements
// count twice.  All sums are accumulated in fp64 whatever T is.  Scalars leave the CTA through one atomicAdd each;
// the ARD pass re-stages the point slabs and reduces one feature at a time.
//
// sums[0] = sum w kappa                (-> d/d variance  = 1/2 sums[0])
// sums[1] = sum w kappa'(r) r  |  sum w <tx, tx'>  (linear)   (-> d/d scale)
// sums[2] = sum w                      (linear only: d/d c)
// sums[3] = sum_i W_ii                 (-> d/d sigma^2 = 1/2 sums[3]; per-point: noise_diag[i] = 1/2 W_ii)
// sums[4] = sum_i alpha_i              (-> d/d mean constant)
// sums[5 + d] = sum w q tdiff_d^2  |  sum w tx_d tx'_d  (linear)        (-> d/d ard_d)
#include "kernels.h"
#include "agp.h"

namespace {

constexpr int RT = 64;   // tile
constexpr int RDC = 32;  // feature chunk

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// kappa(d2) and kappa'(r) r without the variance, fp64
__device__ __forceinline__ void kappa_pair(int family, double d2, double& kap, double& kr) {
  switch (family) {
    case AGP_SE: {
      const double e = exp(-0.5 * d2);
      kap = e; kr = -d2 * e;
      break;
    }
    case AGP_MATERN12: {
      const double d = sqrt(d2), e = exp(-d);
      kap = e; kr = -d * e;
      break;
    }
    case AGP_MATERN32: {
      const double s = 1.7320508075688772935 * sqrt(d2), e = exp(-s);
      kap = (1.0 + s) * e; kr = -3.0 * d2 * e;
      break;
    }
    default: {  // AGP_MATERN52
      const double s = 2.2360679774997896964 * sqrt(d2), e = exp(-s);
      kap = (1.0 + s + s * s * (1.0 / 3.0)) * e; kr = -(5.0 / 3.0) * d2 * (1.0 + s) * e;
      break;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
grad_reduce_kernel(const T* __restrict__ Xt, int D, int64_t n, const T* __restrict__ Cinv, int64_t ldc,
                   const T* __restrict__ alpha, int family, double linear_c, int want_ard,
                   double* __restrict__ sums, T* __restrict__ noise_diag) {
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (tj > ti) return;
  __shared__ T sa[RDC][RT + 1];
  __shared__ T sb[RDC][RT + 1];
  __shared__ double red[8][5];
  __shared__ double sard[RDC];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)ti * RT, col0 = (int64_t)tj * RT;
  const bool linear = (family == AGP_LINEAR);
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  for (int d0 = 0; d0 < D; d0 += RDC) {
    const int dc = min(RDC, D - d0);
    for (int idx = tid; idx < RT * RDC; idx += 256) {
      const int i = idx / RDC, d = idx - i * RDC;
      T va = 0, vb = 0;
      if (d < dc) {
        va = Xt[(row0 + i) * D + d0 + d];
        vb = Xt[(col0 + i) * D + d0 + d];
      }
      sa[d][i] = va;
      sb[d][i] = vb;
    }
    __syncthreads();
#pragma unroll 4
    for (int d = 0; d < RDC; ++d) {
      double a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = (double)sa[d][tx + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = (double)sb[d][ty + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (linear) acc[r][c] += a[r] * b[c];
          else { const double df = a[r] - b[c]; acc[r][c] += df * df; }
        }
    }
    __syncthreads();
  }
  // per-element weights; wq = weight of the element in the ARD sums
  double wq[4][4];
  double s_var = 0.0, s_scale = 0.0, s_c = 0.0, s_noise = 0.0, s_alpha = 0.0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int64_t gj = col0 + ty + 16 * c;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t gi = row0 + tx + 16 * r;
      wq[r][c] = 0.0;
      if (gi >= n || gj >= n || gj > gi) continue;
      const double ai = (double)alpha[gi], aj = (double)alpha[gj];
      const double w = ai * aj - (double)Cinv[gi + gj * ldc];
      const double mult = (gi == gj) ? 1.0 : 2.0;
      if (gi == gj) {
        s_noise += w;
        s_alpha += ai;
        if (noise_diag) noise_diag[gi] = (T)(0.5 * w);
      }
      if (linear) {
        s_var += mult * w * (acc[r][c] + linear_c);
        s_scale += mult * w * acc[r][c];
        s_c += mult * w;
        wq[r][c] = mult * w;
      } else {
        const double d2 = (gi == gj) ? 0.0 : acc[r][c];
        double kap, kr;
        kappa_pair(family, d2, kap, kr);
        s_var += mult * w * kap;
        s_scale += mult * w * kr;
        wq[r][c] = (d2 > 0.0) ? mult * w * kr / d2 : 0.0;
      }
    }
  }
  // block reduction of the five scalars
  {
    double v[5] = {s_var, s_scale, s_c, s_noise, s_alpha};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double t = warp_sum_d(v[q]);
      if (lane == 0) red[wid][q] = t;
    }
    __syncthreads();
    if (tid < 5) {
      double t = 0.0;
      for (int w8 = 0; w8 < 8; ++w8) t += red[w8][tid];
      atomicAdd(&sums[tid], t);
    }
  }
  if (!want_ard) return;
  // ARD pass: one feature at a time, sum_ij wq_ij * (tdiff_d)^2   (linear: wq_ij * tx_d * tx'_d)
  for (int d0 = 0; d0 < D; d0 += RDC) {
    const int dc = min(RDC, D - d0);
    __syncthreads();
    for (int idx = tid; idx < RT * RDC; idx += 256) {
      const int i = idx / RDC, d = idx - i * RDC;
      T va = 0, vb = 0;
      if (d < dc) {
        va = Xt[(row0 + i) * D + d0 + d];
        vb = Xt[(col0 + i) * D + d0 + d];
      }
      sa[d][i] = va;
      sb[d][i] = vb;
    }
    if (tid < RDC) sard[tid] = 0.0;
    __syncthreads();
    for (int d = 0; d < dc; ++d) {
      double a[4], b[4], part = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = (double)sa[d][tx + 16 * r];
#pragma unroll
      for (int c = 0; c < 4; ++c) b[c] = (double)sb[d][ty + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (linear) part += wq[r][c] * a[r] * b[c];
          else { const double df = a[r] - b[c]; part += wq[r][c] * df * df; }
        }
      part = warp_sum_d(part);
      if (lane == 0) atomicAdd(&sard[d], part);
    }
    __syncthreads();
    if (tid < dc) atomicAdd(&sums[5 + d0 + tid], sard[tid]);
  }
}

}  // namespace

template <typename T>
void launch_grad_reduce(const T* Xt, int D, int64_t n, int64_t n_pad, const T* Cinv, int64_t ldc, const T* alpha,
                        int family, double linear_c, int want_ard, double* sums, T* noise_diag, cudaStream_t s) {
  if (n_pad <= 0) return;
  const unsigned nt = (unsigned)(n_pad / RT);
  dim3 grid(nt, nt);
  grad_reduce_kernel<T><<<grid, 256, 0, s>>>(Xt, D, n, Cinv, ldc, alpha, family, linear_c, want_ard, sums, noise_diag);
  agp_count_launch();
}
template void launch_grad_reduce<float>(const float*, int, int64_t, int64_t, const float*, int64_t, const float*, int, double,
                                        int, double*, float*, cudaStream_t);
template void launch_grad_reduce<double>(const double*, int, int64_t, int64_t, const double*, int64_t, const double*, int,
                                         double, int, double*, double*, cudaStream_t);

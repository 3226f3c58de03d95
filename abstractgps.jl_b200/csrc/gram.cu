// gram.cu -- K1/K2: Gram construction fused with the diagonal-noise add.
// Replaces kernelmatrix(k,x[,z]) (KernelFunctions; call sites /root/reference/src/base_gp.jl:70,74)
// and `C + f.Sigma_y` (/root/reference/src/finite_gp_projection.jl:135).
//
// Layout: points are pre-transformed once (ScaleTransform / ARDTransform) into a point-major
// array Xt[n_pad][D].  One CTA produces a 64x64 output tile; both 64 x Dc point slabs are staged
// in shared memory and each thread keeps a 4x4 register block of squared distances computed by
// DIRECT differences (no ||x||^2+||y||^2-2xy cancellation).  Stores are column-major, 16
// consecutive rows per half-warp.  Bound: HBM write of N^2/2 elements at small D, fp64 pipe at
// large D.
#include <atomic>
#include "kernels.h"
#include "agp.h"

static std::atomic<int64_t> g_launches{0};
int64_t agp_kernel_launches() { return g_launches.load(); }
void agp_count_launch() { g_launches.fetch_add(1); }

template <typename T>
__global__ void prep_points_kernel(const T* __restrict__ X, int layout, int64_t n, int64_t n_pad, int D,
                                   int transform, T scale, const T* __restrict__ ard, T* __restrict__ Xt) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t total = n_pad * D;
  if (idx >= total) return;
  int64_t i = idx / D;
  int d = (int)(idx - i * D);
  T v = 0;
  if (i < n) {
    v = (layout == AGP_POINT_MAJOR) ? X[i * D + d] : X[(int64_t)d * n + i];
    if (transform == AGP_T_SCALE) v *= scale;
    else if (transform == AGP_T_ARD) v *= ard[d];
  }
  Xt[idx] = v;
}

template <typename T>
void launch_prep_points(const T* X, int layout, int64_t n, int64_t n_pad, int D, int transform, double scale,
                        const T* ard, T* Xt, cudaStream_t s) {
  int64_t total = n_pad * D;
  if (total == 0) return;
  int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  prep_points_kernel<T><<<(unsigned)blocks, threads, 0, s>>>(X, layout, n, n_pad, D, transform, (T)scale, ard, Xt);
  agp_count_launch();
}
template void launch_prep_points<float>(const float*, int, int64_t, int64_t, int, int, double, const float*, float*, cudaStream_t);
template void launch_prep_points<double>(const double*, int, int64_t, int64_t, int, int, double, const double*, double*, cudaStream_t);

template <typename T> __device__ __forceinline__ T dev_exp(T x);
template <> __device__ __forceinline__ float dev_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double dev_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T dev_sqrt(T x);
template <> __device__ __forceinline__ float dev_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double dev_sqrt<double>(double x) { return sqrt(x); }

template <typename T>
__device__ __forceinline__ T kappa(int family, T acc, T variance, T linear_c) {
  // acc = squared distance (stationary families) or dot product (linear)
  switch (family) {
    case AGP_SE: return variance * dev_exp<T>(-acc * (T)0.5);
    case AGP_MATERN12: return variance * dev_exp<T>(-dev_sqrt<T>(acc));
    case AGP_MATERN32: {
      T s = (T)1.7320508075688772935 * dev_sqrt<T>(acc);
      return variance * ((T)1 + s) * dev_exp<T>(-s);
    }
    case AGP_MATERN52: {
      T s = (T)2.2360679774997896964 * dev_sqrt<T>(acc);
      return variance * ((T)1 + s + s * s * (T)(1.0 / 3.0)) * dev_exp<T>(-s);
    }
    default: return variance * (acc + linear_c);
  }
}

constexpr int GT = 64;  // gram tile
constexpr int GDC = 32; // feature chunk

template <typename T>
__global__ void __launch_bounds__(256)
gram_kernel(const T* __restrict__ Xa, const T* __restrict__ Xb, int D, T* __restrict__ K, int64_t ldk,
            GramParams p) {
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (p.lower_only && (int64_t)tj * GT + p.diag_off > (int64_t)ti * GT + (GT - 1)) return;
  __shared__ T sa[GDC][GT + 1];
  __shared__ T sb[GDC][GT + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)ti * GT, col0 = (int64_t)tj * GT;
  T acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0;
  const bool linear = (p.family == AGP_LINEAR);
  for (int d0 = 0; d0 < D; d0 += GDC) {
    const int dc = min(GDC, D - d0);
    for (int idx = tid; idx < GT * GDC; idx += 256) {
      int i = idx / GDC, d = idx - i * GDC;
      T va = 0, vb = 0;
      if (d < dc) {
        va = Xa[(row0 + i) * D + d0 + d];
        vb = Xb[(col0 + i) * D + d0 + d];
      }
      sa[d][i] = va;
      sb[d][i] = vb;
    }
    __syncthreads();
    if (linear) {
#pragma unroll 4
      for (int d = 0; d < GDC; ++d) {
        T a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = sa[d][tx + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = sb[d][ty + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] += a[r] * b[c];
      }
    } else {
#pragma unroll 4
      for (int d = 0; d < GDC; ++d) {
        T a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = sa[d][tx + 16 * r];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = sb[d][ty + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            T df = a[r] - b[c];
            acc[r][c] += df * df;
          }
      }
    }
    __syncthreads();
  }
  const T variance = (T)p.variance, lc = (T)p.linear_c;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int64_t gj = col0 + ty + 16 * c;
    const int64_t gjg = gj + p.diag_off;  // global column index
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t gi = row0 + tx + 16 * r;
      T v;
      const bool pad_a = p.mask_a ? (p.mask_a[gi] == 0) : (gi >= p.valid_a);
      const bool pad_b = p.mask_b ? (p.mask_b[gj] == 0) : (gjg >= p.valid_b);
      if (pad_a || pad_b) {
        v = (p.symmetric && gi == gjg) ? (T)1 : (T)0;  // identity padding
      } else {
        T a = acc[r][c];
        if (p.symmetric && gi == gjg && !linear) a = 0;  // exactly-zero self distance
        v = kappa<T>(p.family, a, variance, lc);
        if (p.symmetric && gi == gjg && p.noise_kind >= 0)
          v += (p.noise_kind == 0) ? (T)p.noise_s : ((const T*)p.noise_v)[gi - p.noise_off];
      }
      K[gi + gj * ldk] = v;
    }
  }
}

template <typename T>
void launch_gram(const T* Xa, const T* Xb, int64_t na_pad, int64_t nb_pad, int D, T* K, int64_t ldk,
                 const GramParams& p, cudaStream_t s) {
  if (na_pad == 0 || nb_pad == 0) return;
  dim3 grid((unsigned)(na_pad / GT), (unsigned)(nb_pad / GT));
  gram_kernel<T><<<grid, 256, 0, s>>>(Xa, Xb, D, K, ldk, p);
  agp_count_launch();
}
template void launch_gram<float>(const float*, const float*, int64_t, int64_t, int, float*, int64_t, const GramParams&, cudaStream_t);
template void launch_gram<double>(const double*, const double*, int64_t, int64_t, int, double*, int64_t, const GramParams&, cudaStream_t);

// kernelmatrix_diag (/root/reference/src/base_gp.jl:72)
template <typename T>
__global__ void kdiag_kernel(const T* __restrict__ Xt, int64_t n, int D, int family, T variance, T linear_c,
                             T* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (family != AGP_LINEAR) { out[i] = variance; return; }
  T acc = 0;
  for (int d = 0; d < D; ++d) { T v = Xt[i * D + d]; acc += v * v; }
  out[i] = variance * (acc + linear_c);
}
template <typename T>
void launch_kdiag(const T* Xt, int64_t n, int D, int family, double variance, double linear_c, T* out,
                  cudaStream_t s) {
  if (n == 0) return;
  kdiag_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(Xt, n, D, family, (T)variance, (T)linear_c, out);
  agp_count_launch();
}
template void launch_kdiag<float>(const float*, int64_t, int, int, double, double, float*, cudaStream_t);
template void launch_kdiag<double>(const double*, int64_t, int, int, double, double, double*, cudaStream_t);

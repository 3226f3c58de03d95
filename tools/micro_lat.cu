// micro_lat.cu -- dependent-issue latencies that bound the in-SM potrf column chain (fp64 FMA, MUFU.RSQ64H,
// broadcast LDS, bar.sync) and DFMA throughput, measured with clock64 on one SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, int nwarps_active) {
  __shared__ double sm[256];
  sm[threadIdx.x] = threadIdx.x * 1e-3 + 1.0;
  __syncthreads();
  double x = out[0] + 1.0000001, y = 0.999999, z;
  long long t0, t1;
  // 1) dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  // 2) dependent DMUL chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = x * y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  // 3) MUFU.RSQ64H chain (approx rsqrt dependent)
  z = fabs(x) + 1.5;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(z)); z = z + 1.5; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  // 4) dependent LDS chain (pointer chasing through smem with fp64)
  int idx = threadIdx.x & 255;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { double v = sm[idx]; idx = ((int)v + i) & 255; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  // 5) bar.sync x 64
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  // 6) independent DFMA throughput: 16 chains
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = x + i;
  t0 = clock64();
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], y, 1e-9);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  // 7) CUDA rsqrt(double) chain
  z = fabs(z) + 1.5;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { z = rsqrt(z) + 1.5; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  double s = x + z + idx;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 4096);
  for (int threads : {32, 128, 256, 512}) {
    lat<<<1, threads>>>(out, cyc, 0);
    lat<<<1, threads>>>(out, cyc, 0);
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("threads=%3d: DFMA dep %.1f clk | DMUL dep %.1f | MUFU.RSQ64H+DADD %.1f | LDS dep(+cvt) %.1f | bar.sync %.1f | DFMA 16-indep: %.2f clk/warp-instr | rsqrt()+DADD %.1f\n",
           threads, h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / (64.0 * 16), h[6] / 64.0);
  }
  return 0;
}

// micro_dmma.cu -- throughput of the fp64 math paths on this GPU (roofline denominators for the
// fp64 kernels): DFMA, mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 (.f64).
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void __launch_bounds__(256) k_mma(double* out, int iters) {
  double a[8], b[4], c[8][4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i]), "d"(b[0]));
      } else if (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[i]), "d"(a[(i+1)&7]), "d"(b[0]));
      } else if (SHAPE == 2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[i]), "d"(a[(i+1)&7]), "d"(a[(i+2)&7]), "d"(a[(i+3)&7]), "d"(b[0]), "d"(b[1]));
      } else if (SHAPE == 3) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};" : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = fma(a[i], b[j], c[i][j]);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double fma_per_warp_instr, int instr_per_iter) {
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 4 * 256);
  for (int wps = 1; wps <= 8; wps *= 2) {   // CTAs per SM worth of warps: blocks = sms * wps/ (8 warps per block)
    int blocks = sms * wps;  // each block 8 warps
    int iters = 20000;
    k_mma<SHAPE><<<blocks, 256>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_mma<SHAPE><<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)blocks * 8 * iters * instr_per_iter * fma_per_warp_instr;
    printf("%-14s blocks/SM=%d  %.2f TFLOP/s  (%.1f FMA/clk/SM @1.965GHz)\n", name, wps, 2 * fmas / ms * 1e-9, fmas / (ms * 1e-3) / sms / 1.965e9);
  }
  cudaFree(out);
}

int main() {
  run<0>("mma.m8n8k4", 256, 8);
  run<1>("mma.m16n8k4", 512, 8);
  run<2>("mma.m16n8k8", 1024, 8);
  run<3>("mma.m16n8k16", 2048, 8);
  run<4>("dfma", 32, 32);
  return 0;
}

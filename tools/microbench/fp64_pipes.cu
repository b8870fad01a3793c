// Microbenchmark: FP64 throughput of DMMA.8x8x4 alone, DFMA alone, and both issued concurrently
// by different warps of the same SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: all warps DMMA; 1: all warps DFMA; 2: warps with (warp/4)%2==0 DMMA, others DFMA
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, double a0, double b0, int dfma_per_dmma) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = MODE == 0 || (MODE == 2 && ((warp >> 2) & 1) == 0);
  double acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  double a = a0 + threadIdx.x * 1e-12, b = b0;
  if (do_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma(acc[2 * i], acc[2 * i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters * dfma_per_dmma; ++it) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(int threads, int iters, int dfma_per_dmma, double* out, int sms, const char* name) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, threads>>>(out, iters / 10, 1.0000001, 1e-9, dfma_per_dmma);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sms, threads>>>(out, iters, 1.0000001, 1e-9, dfma_per_dmma);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const int warps = threads / 32;
  int mma_w = MODE == 0 ? warps : MODE == 1 ? 0 : 0, fma_w = 0;
  if (MODE == 2) { for (int w = 0; w < warps; ++w) if (((w >> 2) & 1) == 0) ++mma_w; else ++fma_w; }
  if (MODE == 1) fma_w = warps;
  const double mma_flops = (double)sms * mma_w * iters * 16.0 * 512.0;
  const double fma_flops = (double)sms * fma_w * (double)iters * dfma_per_dmma * 32.0 * 32.0 * 2.0;
  printf("%-28s threads=%4d  %.3f ms  DMMA %.2f TF  DFMA %.2f TF  total %.2f TF\n", name, threads, ms,
         mma_flops / ms / 1e9, fma_flops / ms / 1e9, (mma_flops + fma_flops) / ms / 1e9);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  const int iters = 20000;
  for (int threads : {128, 256, 512}) {
    run<0>(threads, iters, 1, out, sms, "DMMA only");
    run<1>(threads, iters, 8, out, sms, "DFMA only");
  }
  for (int r : {2, 4, 6, 8, 12, 16}) {
    char nm[64]; snprintf(nm, sizeof nm, "mixed (DFMA/DMMA work %d)", r);
    run<2>(256, iters, r, out, sms, nm);
    run<2>(512, iters, r, out, sms, nm);
  }
  printf("clock %d kHz, SMs %d\n", p.clockRate, sms);
  return 0;
}

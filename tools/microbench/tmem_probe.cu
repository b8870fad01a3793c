// Probe: tensor memory (TMEM, tcgen05.alloc / st / ld) used as a per-thread scratchpad by ordinary warps.
// 12 compute warps store 14 x 4 doubles each (112 32-bit columns per warp slot, 3 warps per lane quarter), a
// "helper" warpgroup allocates / frees; every value is read back and compared.  Also times N round trips.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NCW = 12, KR = 14, NB = 4;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const double (&v)[4]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, double (&v)[4]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

__global__ void __launch_bounds__(512, 1) probe(double* out, int* bad, long long* cycles, int reps) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == NCW) {   // one warp allocates all 512 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tbase_s;
  int nbad = 0;
  if (warp < NCW) {
    // lane quarter = warp % 4, column slot = warp / 4
    const uint32_t taddr0 = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * (KR * NB * 2));
    for (int k = 0; k < KR; ++k) {
      double v[4];
      for (int m = 0; m < NB; ++m) v[m] = 1000.0 * threadIdx.x + 10.0 * k + m + 0.25;
      tmem_st8(taddr0 + k * 8, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t0 = clock64();
    double acc = 0.0;
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < KR; ++k) {
        double v[4];
        tmem_ld8(taddr0 + k * 8, v);
        for (int m = 0; m < NB; ++m) {
          if (r == 0 && v[m] != 1000.0 * threadIdx.x + 10.0 * k + m + 0.25) ++nbad;
          acc += v[m];
        }
      }
    const long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (lane == 0) cycles[warp] = t1 - t0;
  }
  if (nbad) atomicAdd(bad, nbad);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == NCW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

int main() {
  double* out;
  int* bad;
  long long* cyc;
  cudaMalloc(&out, 512 * sizeof(double));
  cudaMalloc(&bad, sizeof(int));
  cudaMalloc(&cyc, 16 * sizeof(long long));
  cudaMemset(bad, 0, sizeof(int));
  const int reps = 200;
  probe<<<148, 512>>>(out, bad, cyc, reps);
  cudaError_t e = cudaDeviceSynchronize();
  int hbad = -1;
  long long hc[16];
  cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  printf("status=%s mismatches=%d cycles per LDTM.x8 + wait (warp 0)=%.1f (warp 11)=%.1f\n", cudaGetErrorString(e), hbad,
         (double)hc[0] / (reps * KR), (double)hc[11] / (reps * KR));
  return e == cudaSuccess && hbad == 0 ? 0 : 1;
}

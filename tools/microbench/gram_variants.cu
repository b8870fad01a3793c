// Standalone variants of the Gram kernel to locate its bottleneck (see profiles/r01_notes.md).
//   MODE 0: production pipeline      MODE 1: consumers never wait for data (compute + LDS ceiling)
//   MODE 2: producer re-reads k-block 0 (L2-resident source)      MODE 3: BK=16 but only A panel loaded (half the copies)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../bgge_b200/csrc -o gram_variants gram_variants.cu
#include "common.cuh"
#include <vector>
#include <cstdio>
using namespace bgge;
void bgge::set_error(const char*, ...) {}
constexpr int TM = 128, BK = 16, SLD = 132, NCONS = 8, NTHREADS = 384;

template <int MODE, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
gram_var(const double* __restrict__ X, int64_t ldx, int kblocks, const int2* __restrict__ tiles, int ntiles,
         double* __restrict__ S, int64_t lds, int64_t L) {
  constexpr int STAGE_DOUBLES = 2 * BK * SLD;
  constexpr uint32_t TX = (MODE == 3 ? 1u : 2u) * BK * TM * sizeof(double);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)STAGES * STAGE_DOUBLES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS); }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp >= NCONS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp != NCONS) return;
    const int row = lane & 15, op = lane >> 4;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int2 tc = tiles[t];
      const double* src = X + (int64_t)(op ? tc.y : tc.x) * TM + (int64_t)row * ldx;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int stage = it % STAGES;
        mbar_wait(&empty[stage], ((it / STAGES) & 1u) ^ 1u);
        if (lane == 0) mbar_expect_tx(&full[stage], TX);
        __syncwarp();
        double* dst = sm + (size_t)stage * STAGE_DOUBLES + (size_t)op * BK * SLD + (size_t)row * SLD;
        const int kbs = MODE == 2 ? 0 : kb;
        if (MODE != 3 || op == 0) bulk_g2s(dst, src + (int64_t)kbs * BK * ldx, TM * sizeof(double), &full[stage]);
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int wm = warp & 1, wn = warp >> 1, fr = lane >> 2, fk = lane & 3;
  uint32_t it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int2 tc = tiles[t];
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kb = 0; kb < kblocks; ++kb, ++it) {
      const int stage = it % STAGES;
      if (MODE != 1) mbar_wait(&full[stage], (it / STAGES) & 1u);
      const double* As = sm + (size_t)stage * STAGE_DOUBLES + wm * 64 + fr;
      const double* Bs = sm + (size_t)stage * STAGE_DOUBLES + BK * SLD + wn * 32 + fr;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        const int k = kk * 4 + fk;
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = As[k * SLD + i * 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k * SLD + j * 8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
    const int64_t gi0 = (int64_t)tc.x * TM + wm * 64 + fr, gj0 = (int64_t)tc.y * TM + wn * 32 + 2 * fk;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t gi = gi0 + i * 8, gj = gj0 + j * 8 + h;
          if (gi < L && gj < L) S[gi + gj * lds] = acc[i][j][h];
        }
  }
}

template <int MODE, int STAGES>
void run(const char* name, const double* X, int64_t ldx, int kblocks, const int2* tiles, int ntiles, double* S, int64_t L,
         int grid) {
  const size_t smem = (size_t)STAGES * 2 * BK * SLD * 8 + 2 * STAGES * 8;
  cudaFuncSetAttribute(gram_var<MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gram_var<MODE, STAGES><<<grid, NTHREADS, smem>>>(X, ldx, kblocks / 8, tiles, ntiles, S, L, L);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  gram_var<MODE, STAGES><<<grid, NTHREADS, smem>>>(X, ldx, kblocks, tiles, ntiles, S, L, L);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fl = (double)ntiles * 128.0 * 128.0 * 2.0 * kblocks * BK;
  printf("%-44s stages=%d grid=%d  %.2f ms  %.2f TF executed  (%s)\n", name, STAGES, grid, ms, fl / ms / 1e9, cudaGetErrorString(err));
}

__global__ void fill(double* x, size_t n, int mode) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    x[i] = mode == 1 ? (u - 0.5) * 3.4 : (double)((int)(u * 3.0)) - 1.0;   // 1: dense random, 2: {-1,0,1}
  }
}

int main(int argc, char** argv) {
  const int64_t L = 10000, ldx = 10112, p = argc > 2 ? atoll(argv[2]) : 32768;
  const int fmode = argc > 1 ? atoi(argv[1]) : 0;
  double* X; cudaMalloc(&X, ldx * p * 8); cudaMemset(X, 0, ldx * p * 8);
  if (fmode) fill<<<1184, 256>>>(X, (size_t)ldx * p, fmode);
  printf("fill mode %d, p=%lld\n", fmode, (long long)p);
  double* S; cudaMalloc(&S, L * L * 8);
  std::vector<int2> t;
  const int T = 79;
  for (int ti = T - 1; ti >= 0; --ti) for (int tj = 0; tj <= ti; ++tj) t.push_back(make_int2(ti, tj));
  int2* dt; cudaMalloc(&dt, t.size() * sizeof(int2)); cudaMemcpy(dt, t.data(), t.size() * sizeof(int2), cudaMemcpyHostToDevice);
  const int nt = (int)t.size(), kb = (int)(p / BK);
  run<0, 4>("production", X, ldx, kb, dt, nt, S, L, 148);
  run<0, 6>("production", X, ldx, kb, dt, nt, S, L, 148);
  run<1, 4>("no wait on data (compute+LDS ceiling)", X, ldx, kb, dt, nt, S, L, 148);
  run<2, 4>("L2-resident source", X, ldx, kb, dt, nt, S, L, 148);
  run<3, 4>("half the copies (A panel only)", X, ldx, kb, dt, nt, S, L, 148);
  run<0, 4>("production, 2 tiles (1 per CTA, no contention)", X, ldx, kb, dt, 2, S, L, 2);
  run<0, 4>("production, 74 CTAs", X, ldx, kb, dt, 74 * 4, S, L, 74);
  return 0;
}

// Microbenchmark for the genotype expansion (R/getK.R:138): store-path variants at C5 size (n = 40000, L = 10000).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o expand_probe expand_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// RPT rows per thread (even), CPB columns per block, MODE 0 (G) / 1 (GE mask); jord: column order (nullptr: natural)
template <int RPT, int CPB, int MODE>
__global__ void expand_v(const double* __restrict__ K0, long long ldk, const int* __restrict__ gid, const int* __restrict__ env,
                         long long n, double* __restrict__ out, long long ldo, const int* __restrict__ jord) {
  const long long i = RPT * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
  if (i >= n) return;
  int g[RPT], e[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    g[r] = gid[min(i + r, n - 1)];
    e[r] = MODE ? env[min(i + r, n - 1)] : 0;
  }
  const long long j0 = (long long)blockIdx.y * CPB;
#pragma unroll 4
  for (int jj = 0; jj < CPB; ++jj) {
    if (j0 + jj >= n) break;
    const long long j = jord ? jord[j0 + jj] : j0 + jj;
    const double* col = K0 + (long long)gid[j] * ldk;
    const int ej = MODE ? env[j] : 0;
    double v[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) v[r] = (!MODE || e[r] == ej) ? __ldg(col + g[r]) : 0.0;
    double* dst = out + i + j * ldo;
    if (i + RPT <= n) {
#pragma unroll
      for (int r = 0; r < RPT; r += 2) __stcs(reinterpret_cast<double2*>(dst + r), make_double2(v[r], v[r + 1]));
    } else {
      for (int r = 0; r < RPT && i + r < n; ++r) dst[r] = v[r];
    }
  }
}

template <int RPT, int CPB, int MODE>
float run(const double* K0, long long L, const int* gid, const int* env, long long n, double* out, const int* jord) {
  dim3 grid((unsigned)((n + 256 * RPT - 1) / (256 * RPT)), (unsigned)((n + CPB - 1) / CPB));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    expand_v<RPT, CPB, MODE><<<grid, 256>>>(K0, L, gid, env, n, out, n, jord);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (rep) best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const long long L = 10000, E = 4, n = L * E;
  double *K0, *out;
  int *gid, *env, *jord;
  CK(cudaMalloc(&K0, L * L * 8));
  CK(cudaMalloc(&out, n * n * 8));
  CK(cudaMalloc(&gid, n * 4));
  CK(cudaMalloc(&env, n * 4));
  CK(cudaMalloc(&jord, n * 4));
  std::vector<int> hg(n), he(n), ho(n);
  for (long long i = 0; i < n; ++i) { hg[i] = (int)(i % L); he[i] = (int)(i / L); }
  std::iota(ho.begin(), ho.end(), 0);
  std::stable_sort(ho.begin(), ho.end(), [&](int a, int b) { return hg[a] < hg[b]; });
  CK(cudaMemcpy(gid, hg.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(env, he.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(jord, ho.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(K0, 0, L * L * 8));
  const double gb = 8.0 * n * n / 1e9;
#define R(RPT, CPB, MODE, ORD) { float ms = run<RPT, CPB, MODE>(K0, L, gid, env, n, out, ORD ? jord : nullptr); \
    printf("rpt %d cpb %2d mode %d %s  %7.3f ms  %7.1f GB/s\n", RPT, CPB, MODE, ORD ? "sorted " : "natural", ms, gb / (ms * 1e-3)); }
  R(2, 16, 0, 0) R(2, 16, 0, 1) R(4, 16, 0, 0) R(4, 16, 0, 1) R(4, 8, 0, 1) R(2, 32, 0, 1) R(4, 32, 0, 1) R(8, 16, 0, 1)
  R(2, 16, 1, 0) R(2, 16, 1, 1) R(4, 16, 1, 0) R(4, 16, 1, 1) R(4, 8, 1, 1) R(4, 32, 1, 1) R(8, 16, 1, 1)
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a));
  CK(cudaMemsetAsync(out, 0, n * n * 8));
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("memset  %7.3f ms  %7.1f GB/s\n", ms, gb / (ms * 1e-3));
  return 0;
}

// Which cuSOLVER symmetric eigensolver entry is fastest for the setDEC problem (n = 10000, FP64, vectors)?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o syevd_probe syevd_probe.cu -lcusolver -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <chrono>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
#define CS(x) do { cusolverStatus_t s = (x); if (s != CUSOLVER_STATUS_SUCCESS) { printf("%s: cusolver status %d\n", #x, (int)s); return -1.0; } } while (0)

__global__ void fill(double* A, long long n, unsigned long long seed) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  unsigned long long x = idx * 6364136223846793005ull + seed;
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33;
  A[idx] = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 10000;
  double *X, *A, *W;
  int* info;
  CK(cudaMalloc(&X, n * n * 8));
  CK(cudaMalloc(&A, n * n * 8));
  CK(cudaMalloc(&W, n * 8));
  CK(cudaMalloc(&info, 4));
  fill<<<(unsigned)((n * n + 255) / 256), 256>>>(X, n, 12345);
  cublasHandle_t cb;
  cublasCreate(&cb);
  cusolverDnHandle_t h;
  cusolverDnCreate(&h);
  cusolverDnParams_t params;
  cusolverDnCreateParams(&params);
  const double one = 1.0 / n, zero = 0.0;
  auto make = [&]() { cublasDsyrk(cb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, (int)n, (int)n, &one, X, (int)n, &zero, A, (int)n); CK(cudaDeviceSynchronize()); };

  auto xsyevd = [&]() -> double {
    size_t wd = 0, wh = 0;
    CS(cusolverDnXsyevd_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, &wd, &wh));
    void* dw; CK(cudaMalloc(&dw, wd)); std::vector<char> hw(wh);
    make();
    double t0 = now();
    CS(cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, dw, wd, hw.data(), wh, info));
    CK(cudaDeviceSynchronize());
    double t = now() - t0;
    cudaFree(dw);
    return t;
  };
  auto dsyevd = [&]() -> double {
    int lw = 0;
    CS(cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, A, (int)n, W, &lw));
    double* dw; CK(cudaMalloc(&dw, (size_t)lw * 8));
    make();
    double t0 = now();
    CS(cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, A, (int)n, W, dw, lw, info));
    CK(cudaDeviceSynchronize());
    double t = now() - t0;
    cudaFree(dw);
    return t;
  };
  auto xbatched = [&]() -> double {
    size_t wd = 0, wh = 0;
    CS(cusolverDnXsyevBatched_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, &wd, &wh, 1));
    void* dw; CK(cudaMalloc(&dw, wd)); std::vector<char> hw(wh);
    make();
    double t0 = now();
    CS(cusolverDnXsyevBatched(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, dw, wd, hw.data(), wh, info, 1));
    CK(cudaDeviceSynchronize());
    double t = now() - t0;
    cudaFree(dw);
    return t;
  };
  auto xsyevdx = [&]() -> double {
    size_t wd = 0, wh = 0;
    double vl = 0, vu = 0; int64_t meig = 0;
    CS(cusolverDnXsyevdx_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_ALL, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, &vl, &vu, 0, 0, &meig, CUDA_R_64F, W, CUDA_R_64F, &wd, &wh));
    void* dw; CK(cudaMalloc(&dw, wd)); std::vector<char> hw(wh);
    make();
    double t0 = now();
    CS(cusolverDnXsyevdx(h, params, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_ALL, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, &vl, &vu, 0, 0, &meig, CUDA_R_64F, W, CUDA_R_64F, dw, wd, hw.data(), wh, info));
    CK(cudaDeviceSynchronize());
    double t = now() - t0;
    cudaFree(dw);
    return t;
  };
  for (int rep = 0; rep < 2; ++rep) {
    printf("n=%lld rep %d: Xsyevd %.3f s", n, rep, xsyevd()); fflush(stdout);
    printf("  Dsyevd %.3f s", dsyevd()); fflush(stdout);
    printf("  Xsyevdx(all) %.3f s", xsyevdx()); fflush(stdout);
    printf("  XsyevBatched(1) %.3f s\n", xbatched()); fflush(stdout);
  }
  // values only (how much of the time is the vectors?)
  {
    size_t wd = 0, wh = 0;
    cusolverDnXsyevd_bufferSize(h, params, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, &wd, &wh);
    void* dw; CK(cudaMalloc(&dw, wd)); std::vector<char> hw(wh);
    make();
    double t0 = now();
    cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W, CUDA_R_64F, dw, wd, hw.data(), wh, info);
    CK(cudaDeviceSynchronize());
    printf("Xsyevd values only %.3f s\n", now() - t0);
  }
  return 0;
}

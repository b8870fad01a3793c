// Per-kernel spectral decomposition with tol truncation: eig() / setDEC() of R/BGGE.R:112-182.
// cuSOLVER Xsyevd (lower triangle, like LAPACK dsyevr behind R's eigen()), then a device pass
// that reverses to descending order, keeps values > tol, compacts U and (optionally) fixes the
// sign of each eigenvector so that its largest-|.| component is positive.
#include "common.cuh"
#include "kernels.h"
#include <cusolverDn.h>
#include <vector>
#include <mutex>

namespace bgge {
namespace {

std::mutex g_mu;
cusolverDnHandle_t g_handle[64] = {nullptr};

int get_handle(cusolverDnHandle_t* out) {
  int dev = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  if (dev < 0 || dev >= 64) {
    set_error("eig: device index %d out of range", dev);
    return ERR_INVALID;
  }
  if (!g_handle[dev]) {
    cusolverStatus_t s = cusolverDnCreate(&g_handle[dev]);
    if (s != CUSOLVER_STATUS_SUCCESS) {
      set_error("cusolverDnCreate failed: %d", (int)s);
      return ERR_SOLVER;
    }
  }
  *out = g_handle[dev];
  return OK;
}

// one CTA per kept eigenpair c: Uout[:, c] = sign * V[:, m-1-c], sout[c] = W[m-1-c]
__global__ void compact_desc_kernel(const double* __restrict__ V, int64_t ldv, const double* __restrict__ W,
                                    int64_t m, int64_t nr, int canonical, double* __restrict__ Uout,
                                    int64_t ldu, double* __restrict__ sout) {
  const int64_t c = blockIdx.x;
  if (c >= nr) return;
  const double* src = V + (m - 1 - c) * ldv;
  double sgn = 1.0;
  if (canonical) {
    __shared__ double sa[32];
    __shared__ long long si[32];
    double best = -1.0;
    long long bi = 0;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
      const double a = fabs(src[i]);
      if (a > best) { best = a; bi = i; }   // strided scan keeps the smallest index per thread on ties
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int nw = blockDim.x >> 5;
      best = threadIdx.x < nw ? sa[threadIdx.x] : -1.0;
      bi = threadIdx.x < nw ? si[threadIdx.x] : 0;
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (threadIdx.x == 0) si[0] = bi;
    }
    __syncthreads();
    sgn = src[si[0]] < 0.0 ? -1.0 : 1.0;
  }
  double* dst = Uout + c * ldu;
  for (int64_t i = threadIdx.x; i < ldu; i += blockDim.x) dst[i] = i < m ? sgn * src[i] : 0.0;
  if (threadIdx.x == 0) sout[c] = W[m - 1 - c];
}

}  // namespace

// Step 1: syevd in place (dA lower triangle -> eigenvectors, ascending), eigenvalues to d_W,
// number of values > tol to *nr_out (which(values > tol), R/BGGE.R:114).  Synchronises.
int eig_decompose(double* dA, int64_t m, int64_t lda, double tol, double* d_W, int64_t* nr_out, cudaStream_t st) {
  BGGE_REQUIRE(m > 0 && lda >= m, "eig: bad dimensions m=%lld lda=%lld", (long long)m, (long long)lda);
  cusolverDnHandle_t h;
  BGGE_TRY(get_handle(&h));
  cusolverDnSetStream(h, st);
  cusolverDnParams_t params;
  cusolverDnCreateParams(&params);
  size_t wdev = 0, whost = 0;
  DevBuf<int> info;
  DevBuf<unsigned char> work;
  BGGE_CUDA(info.alloc(1));
  cusolverStatus_t cs = cusolverDnXsyevd_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m,
                                                    CUDA_R_64F, dA, lda, CUDA_R_64F, d_W, CUDA_R_64F, &wdev, &whost);
  if (cs != CUSOLVER_STATUS_SUCCESS) {
    cusolverDnDestroyParams(params);
    set_error("cusolverDnXsyevd_bufferSize failed: %d", (int)cs);
    return ERR_SOLVER;
  }
  BGGE_CUDA(work.alloc(wdev));
  std::vector<unsigned char> hwork(whost);
  cs = cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m, CUDA_R_64F, dA, lda,
                        CUDA_R_64F, d_W, CUDA_R_64F, work.p, wdev, hwork.data(), whost, info.p);
  cusolverDnDestroyParams(params);
  if (cs != CUSOLVER_STATUS_SUCCESS) {
    set_error("cusolverDnXsyevd failed: %d", (int)cs);
    return ERR_SOLVER;
  }
  int hinfo = 0;
  std::vector<double> hW((size_t)m);
  BGGE_CUDA(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  BGGE_CUDA(cudaMemcpyAsync(hW.data(), d_W, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
  BGGE_CUDA(cudaStreamSynchronize(st));
  if (hinfo != 0) {
    set_error("syevd did not converge (info=%d)", hinfo);
    return ERR_SOLVER;
  }
  int64_t nr = 0;   // ascending order: the kept values are the trailing ones
  for (int64_t i = m - 1; i >= 0; --i) {
    if (hW[(size_t)i] > tol) ++nr; else break;
  }
  *nr_out = nr;
  return OK;
}

// Step 2: descending order, first nr pairs, optional sign fix.  d_s: nr doubles, dU: ldu x nr.
int eig_compact(const double* dV, int64_t ldv, const double* d_W, int64_t m, int64_t nr, int canonical, double* d_s,
                double* dU, int64_t ldu, cudaStream_t st) {
  if (nr <= 0) return OK;
  BGGE_REQUIRE(ldu >= m, "eig: ldu=%lld < m=%lld", (long long)ldu, (long long)m);
  compact_desc_kernel<<<(unsigned)nr, 256, 0, st>>>(dV, ldv, d_W, m, nr, canonical, dU, ldu, d_s);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

}  // namespace bgge

// Per-kernel spectral decomposition with tol truncation: eig() / setDEC() of R/BGGE.R:112-182.
// cuSOLVER Xsyevd (lower triangle, like LAPACK dsyevr behind R's eigen()), then a device pass
// that reverses to descending order, keeps values > tol, compacts U and (optionally) fixes the
// sign of each eigenvector so that its largest-|.| component is positive.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cusolverDn.h>
#include <memory>
#include <new>
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

// M[a,b] = sqrt(c_a c_b) * K0[T_a, T_b]  (t x t), or sqrt(c_a c_b) * [a == b] when K0 == nullptr (Gi)
__global__ void gather_scale_kernel(const double* __restrict__ K0, int64_t ldk, const int* __restrict__ T,
                                    const double* __restrict__ sqc, int64_t t, double* __restrict__ M) {
  const int64_t b = blockIdx.y;
  const int64_t gb = T[b];
  const double sb = sqc[b];
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < t; a += (int64_t)gridDim.x * blockDim.x) {
    const double k = K0 ? K0[T[a] + gb * ldk] : (a == b ? 1.0 : 0.0);
    M[a + b * t] = sqc[a] * sb * k;
  }
}

// Two independent 64-bit hashes of the bit patterns of x[0..count): position-weighted sums mod 2^64 (integer
// addition is associative, so the result does not depend on the launch shape).  Identifies identical
// matrices across kernels of one model (the environment kernels of MDe are the same L x L matrix).
__global__ void hash_bits_kernel(const double* __restrict__ x, long long count, unsigned long long* __restrict__ out) {
  unsigned long long h1 = 0, h2 = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x[i]);
    h1 += b * (2ull * (unsigned long long)i + 1ull);
    unsigned long long m = b ^ (b >> 29);
    m *= 0x9E3779B97F4A7C15ull;
    m ^= m >> 32;
    h2 += m * (0xD6E8FEB86659FD93ull * (unsigned long long)(i + 1) | 1ull);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
    h2 += __shfl_xor_sync(0xffffffffu, h2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, h1);
    atomicAdd(out + 1, h2);
  }
}

// U[i, c] = W[loc[i], c] * isq[i]   (rows x nr into the padded layout, pad rows zero)
__global__ void expand_rows_kernel(const double* __restrict__ W, int64_t ldw, const int* __restrict__ loc,
                                   const double* __restrict__ isq, int64_t rows, int64_t nr, double* __restrict__ U,
                                   int64_t ldu) {
  const int64_t c = blockIdx.y;
  if (c >= nr) return;
  const double* w = W + c * ldw;
  double* u = U + c * ldu;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ldu; i += (int64_t)gridDim.x * blockDim.x)
    u[i] = i < rows ? w[loc[i]] * isq[i] : 0.0;
}

// dst[:, c] = +-src[:, c] so that the largest-|.| entry (first on ties) is positive; pad rows zeroed
__global__ void sign_fix_kernel(const double* __restrict__ src, int64_t ld, int64_t rows, const double* __restrict__ wts,
                                double* __restrict__ dst) {
  const int64_t c = blockIdx.x;
  const double* v = src + c * ld;
  __shared__ double sa[32];
  __shared__ long long si[32];
  double best = -1.0;
  long long bi = 0;
  for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) {
    const double a = fabs(v[i]) * (wts ? wts[i] : 1.0);   // weights: |U| of the expanded vector is |W| / sqrt(count)
    if (a > best) { best = a; bi = i; }
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
  const double sgn = v[si[0]] < 0.0 ? -1.0 : 1.0;
  double* d = dst + c * ld;
  for (int64_t i = threadIdx.x; i < ld; i += blockDim.x) d[i] = i < rows ? sgn * v[i] : 0.0;
}

}  // namespace

int sign_fix_launch(const double* dSrc, int64_t ld, int64_t rows, int64_t nr, const double* d_wts, double* dDst,
                    cudaStream_t st) {
  if (nr <= 0) return OK;
  sign_fix_kernel<<<(unsigned)nr, 256, 0, st>>>(dSrc, ld, rows, d_wts, dDst);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int gather_scale_launch(const double* dK0, int64_t ldk, const int* dT, const double* d_sqc, int64_t t, double* dM,
                        cudaStream_t st) {
  if (t <= 0) return OK;
  unsigned gx = (unsigned)((t + 255) / 256);
  if (gx > 32) gx = 32;
  BGGE_REQUIRE(t <= 65535, "expanded eigen: %lld distinct genotypes in one block exceed 65535", (long long)t);
  gather_scale_kernel<<<dim3(gx, (unsigned)t), 256, 0, st>>>(dK0, ldk, dT, d_sqc, t, dM);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int hash_bits_launch(const double* dX, int64_t count, unsigned long long* d_out2, cudaStream_t st) {
  BGGE_CUDA(cudaMemsetAsync(d_out2, 0, 2 * sizeof(unsigned long long), st));
  unsigned grid = (unsigned)std::min<int64_t>(2048, (count + 255) / 256);
  hash_bits_kernel<<<grid ? grid : 1, 256, 0, st>>>(dX, (long long)count, d_out2);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int expand_rows_launch(const double* dW, int64_t ldw, const int* d_loc, const double* d_isq, int64_t rows, int64_t nr,
                       double* dU, int64_t ldu, cudaStream_t st) {
  if (nr <= 0) return OK;
  unsigned gx = (unsigned)((ldu + 255) / 256);
  if (gx > 64) gx = 64;
  BGGE_REQUIRE(nr <= 65535, "expanded eigen: %lld eigenpairs in one block exceed 65535", (long long)nr);
  expand_rows_kernel<<<dim3(gx, (unsigned)nr), 256, 0, st>>>(dW, ldw, d_loc, d_isq, rows, nr, dU, ldu);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

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
  const bool timing = getenv("BGGE_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[bgge timing]   eig m=%lld %-22s %8.3f ms (host workspace %zu B, device %zu B)\n", (long long)m, what,
            std::chrono::duration<double, std::milli>(t1 - t0).count(), whost, wdev);
    t0 = t1;
  };
  BGGE_CUDA(work.alloc(wdev));
  lap("device workspace");
  // host workspace: one buffer per thread, kept across calls and never value-initialised (a fresh, zero-filled
  // std::vector of this size is page-faulted in on every call)
  static thread_local std::unique_ptr<unsigned char[]> hbuf;
  static thread_local size_t hcap = 0;
  if (whost > hcap) {
    hbuf.reset(new (std::nothrow) unsigned char[whost]);
    BGGE_REQUIRE(hbuf != nullptr, "eig: out of host memory (%zu bytes of cuSOLVER workspace)", whost);
    hcap = whost;
  }
  lap("host workspace");
  cs = cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m, CUDA_R_64F, dA, lda,
                        CUDA_R_64F, d_W, CUDA_R_64F, work.p, wdev, hbuf.get(), whost, info.p);
  lap("Xsyevd");
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

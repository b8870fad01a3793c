// getK device kernels other than the contraction (gram.cu):
//   * symmetrize            lower triangle -> upper (after row-panel builds)
//   * gk_distance           D = (sqrt(g_ii + g_jj - 2 g_ij))^2, exact-zero diagonal   (R/getK.R:142)
//   * quantile7             exact type-7 quantile of all L^2 entries of D              (R/getK.R:146)
//   * gk_exp                K0 = exp((-bw * D) / q)                                    (R/getK.R:146)
//   * expand                G / GE (MDs) / env-k (MDe) / Gi expansions as gathers      (R/getK.R:138,148,173,199-201,207-218,232)
// All of these are HBM-bound elementwise/gather passes over L^2 or n^2 doubles.
#include "common.cuh"
#include "kernels.h"

namespace bgge {
namespace {

// ------------------------------------------------------------------ symmetrize
__global__ void symmetrize_kernel(double* __restrict__ S, int64_t lds, int64_t L, int from_upper) {
  // 32x32 tile transpose through shared memory; only tiles strictly below the diagonal are
  // read, diagonal tiles mirror in place.  from_upper: the same with the roles of the triangles swapped
  // (block (bi, bj) then is the source tile in the UPPER triangle: rows i0.., columns j0.., bj >= bi).
  __shared__ double t[32][33];
  const int bi = from_upper ? blockIdx.y : blockIdx.x, bj = from_upper ? blockIdx.x : blockIdx.y;
  if (from_upper ? (bj < bi) : (bj > bi)) return;
  const int64_t i0 = (int64_t)bi * 32, j0 = (int64_t)bj * 32;
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    const int64_t i = i0 + threadIdx.x, j = j0 + c;
    t[c][threadIdx.x] = (i < L && j < L) ? S[i + j * lds] : 0.0;
  }
  __syncthreads();
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    // write S[j0 + x, i0 + c] = S[i0 + c, j0 + x] = t[x][c]
    const int64_t r = j0 + threadIdx.x, q = i0 + c;
    if (r < L && q < L && (bi != bj || (from_upper ? r > q : r < q))) S[r + q * lds] = t[threadIdx.x][c];
  }
}

// ------------------------------------------------------------------ GK distance
__global__ void take_diag_kernel(const double* __restrict__ S, int64_t lds, int64_t L, double* __restrict__ diag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < L) diag[i] = S[i + i * lds];
}

__global__ void gk_distance_kernel(double* __restrict__ S, int64_t lds, int64_t L, const double* __restrict__ diag) {
  for (int64_t j = blockIdx.y; j < L; j += gridDim.y) {
    const double gjj = diag[j];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
      double d = 0.0;
      if (i != j) {
        d = (diag[i] + gjj) - 2.0 * S[i + j * lds];
        d = d > 0.0 ? d : 0.0;
        const double r = sqrt(d);   // dist() returns the root, getK squares it again
        d = r * r;
      }
      S[i + j * lds] = d;
    }
  }
}

// ------------------------------------------------------------------ radix select (8 x 8-bit passes, MSB first)
struct SelectState {
  unsigned long long prefix;   // high bits fixed so far
  unsigned long long rank;     // remaining 0-based rank inside the prefix bucket
  unsigned long long hist[256];
};

__global__ void select_hist_kernel(const double* __restrict__ D, int64_t ldd, int64_t L, int pass,
                                   SelectState* __restrict__ st) {
  __shared__ unsigned int h[256];
  for (int b = threadIdx.x; b < 256; b += blockDim.x) h[b] = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = st->prefix;
  const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
  for (int64_t j = blockIdx.y; j < L; j += gridDim.y) {
    const double* col = D + j * ldd;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(col[i]);
      if ((key & himask) == prefix) {
        const unsigned int digit = (unsigned int)(key >> shift) & 0xffu;
        const unsigned int peers = __match_any_sync(__activemask(), digit);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[digit], __popc(peers));
      }
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 256; b += blockDim.x)
    if (h[b]) atomicAdd(&st->hist[b], (unsigned long long)h[b]);
}

__global__ void select_step_kernel(SelectState* st, int pass) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int shift = 56 - 8 * pass;
  unsigned long long r = st->rank;
  int b = 0;
  for (; b < 255; ++b) {
    const unsigned long long c = st->hist[b];
    if (r < c) break;
    r -= c;
  }
  st->rank = r;
  st->prefix |= (unsigned long long)b << shift;
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

__global__ void select_init_kernel(SelectState* st, unsigned long long rank) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->prefix = 0;
  st->rank = rank;
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

// type-7 combination (stats::quantile.default): qs = x[lo]; if (index > lo && x[hi] != qs)
// qs = (1 - h) * qs + h * x[hi]
__global__ void quantile_combine_kernel(const SelectState* lo, const SelectState* hi, double h, double* q_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double xlo = __longlong_as_double((long long)lo->prefix);
  const double xhi = __longlong_as_double((long long)hi->prefix);
  double qs = xlo;
  if (h > 0.0 && xhi != qs) qs = (1.0 - h) * qs + h * xhi;
  *q_out = qs;
}

// ------------------------------------------------------------------ GK exp map
__global__ void gk_exp_kernel(const double* __restrict__ D, int64_t ldd, int64_t L, double bw,
                              const double* __restrict__ q, double* __restrict__ K0, int64_t ldk) {
  const double qq = *q;
  const double nbw = -bw;
  for (int64_t j = blockIdx.y; j < L; j += gridDim.y)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x)
      K0[i + j * ldk] = exp((nbw * D[i + j * ldd]) / qq);
}

// ------------------------------------------------------------------ expansions
// mode 0: G[i,j]  = K0[gid i, gid j]
// mode 1: GE[i,j] = K0[gid i, gid j] * [env i == env j]                  (MDs)
// mode 2: Kk[i,j] = K0[gid i, gid j] * [env i == k] * [env j == k]       (MDe)
// mode 3: Gi[i,j] = [gid i == gid j]                                     (intercept.random)
template <int MODE>
__global__ void expand_kernel(const double* __restrict__ K0, int64_t ldk, const int* __restrict__ gid,
                              const int* __restrict__ env, int64_t n, int kenv, double* __restrict__ out,
                              int64_t ldo) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int gi = gid[i];
  const int ei = (MODE == 1 || MODE == 2) ? env[i] : 0;
  const int64_t j0 = (int64_t)blockIdx.y * 16;
#pragma unroll 4
  for (int jj = 0; jj < 16; ++jj) {
    const int64_t j = j0 + jj;
    if (j >= n) break;
    const int gj = gid[j];
    double v;
    if (MODE == 0) {
      v = K0[gi + (int64_t)gj * ldk];
    } else if (MODE == 1) {
      v = (ei == env[j]) ? K0[gi + (int64_t)gj * ldk] : 0.0;
    } else if (MODE == 2) {
      v = (ei == kenv && env[j] == kenv) ? K0[gi + (int64_t)gj * ldk] : 0.0;
    } else {
      v = (gi == gj) ? 1.0 : 0.0;
    }
    out[i + j * ldo] = v;
  }
}

// Two rows per thread, one 16-byte streaming store per element pair (ldo even, 16-byte aligned output): the pass is
// write-only -- K0 and the row codes stay in L2 -- so it is bound by the store path; evict-first stores keep the
// 12.8 GB output of a C5 kernel from flushing K0 out of L2.  Same values as expand_kernel (a gather and a mask).
// jord (optional): the columns in genotype order (expand_order_*): the CTAs that run together then need the same few
// columns of K0 -- K0 (0.8 GB at C5) is read from HBM once instead of once per environment -- and write E distant
// regions of the output at a time: 4285 -> 6105 GB/s at n = 40000 (memset of the same buffer: 7325 GB/s,
// profiles/r02_expand_probe.log).  The order only decides WHEN an element is written, never its value.
template <int MODE>
__global__ void expand2_kernel(const double* __restrict__ K0, int64_t ldk, const int* __restrict__ gid,
                               const int* __restrict__ env, int64_t n, int kenv, double* __restrict__ out,
                               int64_t ldo, const int* __restrict__ jord) {
  const int64_t i = 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
  if (i >= n) return;
  const bool two = i + 1 < n;
  const int ga = gid[i], gb = two ? gid[i + 1] : ga;
  const int ea = (MODE == 1 || MODE == 2) ? env[i] : 0;
  const int eb = (MODE == 1 || MODE == 2) ? (two ? env[i + 1] : ea) : 0;
  const int64_t j0 = (int64_t)blockIdx.y * 16;
#pragma unroll 4
  for (int jj = 0; jj < 16; ++jj) {
    if (j0 + jj >= n) break;
    const int64_t j = jord ? (int64_t)jord[j0 + jj] : j0 + jj;
    const int gj = gid[j];
    const double* col = K0 + (int64_t)gj * ldk;
    double2 v;
    if (MODE == 0) {
      v.x = __ldg(col + ga);
      v.y = __ldg(col + gb);
    } else if (MODE == 1) {
      const int ej = env[j];
      v.x = (ea == ej) ? __ldg(col + ga) : 0.0;
      v.y = (eb == ej) ? __ldg(col + gb) : 0.0;
    } else if (MODE == 2) {
      const bool cj = env[j] == kenv;
      v.x = (cj && ea == kenv) ? __ldg(col + ga) : 0.0;
      v.y = (cj && eb == kenv) ? __ldg(col + gb) : 0.0;
    } else {
      v.x = (ga == gj) ? 1.0 : 0.0;
      v.y = (gb == gj) ? 1.0 : 0.0;
    }
    double* dst = out + i + j * ldo;
    if (two) __stcs(reinterpret_cast<double2*>(dst), v);
    else __stcs(dst, v.x);
  }
}

// Counting sort of the columns by genotype code (codes in [0, L); anything else goes to bucket L): count -> exclusive
// scan -> scatter.  The order inside a bucket depends on the atomics and is irrelevant (see expand2_kernel).
__global__ void expand_order_count_kernel(const int* __restrict__ gid, int64_t n, int L, int* __restrict__ cnt) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int g = gid[j];
  atomicAdd(&cnt[(unsigned)g < (unsigned)L ? g : L], 1);
}

__global__ void __launch_bounds__(1024) expand_order_scan_kernel(int* __restrict__ cnt, int nb) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  const int per = (nb + 1023) / 1024;
  const int lo = min(nb, t * per), hi = min(nb, lo + per);
  int s = 0;
  for (int k = lo; k < hi; ++k) s += cnt[k];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {   // inclusive scan of the per-thread sums
    const int v = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = t > 0 ? part[t - 1] : 0;
  for (int k = lo; k < hi; ++k) {
    const int c = cnt[k];
    cnt[k] = run;
    run += c;
  }
}

__global__ void expand_order_fill_kernel(const int* __restrict__ gid, int64_t n, int L, int* __restrict__ cursor,
                                         int* __restrict__ jord) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int g = gid[j];
  jord[atomicAdd(&cursor[(unsigned)g < (unsigned)L ? g : L], 1)] = (int)j;
}

}  // namespace

// ====================================================================== launchers
int symmetrize_launch(double* dS, int64_t lds, int64_t L, cudaStream_t st, int from_upper) {
  if (L <= 0) return OK;
  const unsigned nb = (unsigned)((L + 31) / 32);
  symmetrize_kernel<<<dim3(nb, nb), dim3(32, 8), 0, st>>>(dS, lds, L, from_upper);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int gk_distance_launch(double* dS, int64_t lds, int64_t L, double* d_diag_scratch, cudaStream_t st) {
  if (L <= 0) return OK;
  take_diag_kernel<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(dS, lds, L, d_diag_scratch);
  BGGE_CUDA(cudaGetLastError());
  unsigned gx = (unsigned)((L + 255) / 256);
  if (gx > 64) gx = 64;
  gk_distance_kernel<<<dim3(gx, (unsigned)(L < 65535 ? L : 65535)), 256, 0, st>>>(dS, lds, L, d_diag_scratch);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

size_t quantile_scratch_bytes() { return 2 * sizeof(SelectState); }

int quantile7_launch(const double* dD, int64_t ldd, int64_t L, double prob, void* d_scratch, double* d_q,
                     cudaStream_t st) {
  BGGE_REQUIRE(L > 0, "quantile7: empty matrix");
  BGGE_REQUIRE(prob >= 0.0 && prob <= 1.0, "quantile7: 'probs' outside [0,1]");
  SelectState* states = reinterpret_cast<SelectState*>(d_scratch);
  const double nn = (double)L * (double)L;
  const double index = (nn - 1.0) * prob;   // 0-based
  const double lo = floor(index), hi = ceil(index);
  const double h = index - lo;
  const unsigned long long ranks[2] = {(unsigned long long)lo, (unsigned long long)hi};
  unsigned gy = (unsigned)(L < 1184 ? L : 1184);
  unsigned gx = (unsigned)((L + 1023) / 1024);
  if (gx > 8) gx = 8;
  for (int w = 0; w < 2; ++w) {
    select_init_kernel<<<1, 32, 0, st>>>(states + w, ranks[w]);
    for (int pass = 0; pass < 8; ++pass) {
      select_hist_kernel<<<dim3(gx, gy), 256, 0, st>>>(dD, ldd, L, pass, states + w);
      select_step_kernel<<<1, 32, 0, st>>>(states + w, pass);
    }
  }
  quantile_combine_kernel<<<1, 32, 0, st>>>(states, states + 1, h, d_q);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int gk_exp_launch(const double* dD, int64_t ldd, int64_t L, double bw, const double* d_q, double* dK0, int64_t ldk,
                  cudaStream_t st) {
  if (L <= 0) return OK;
  unsigned gx = (unsigned)((L + 255) / 256);
  if (gx > 64) gx = 64;
  gk_exp_kernel<<<dim3(gx, (unsigned)(L < 65535 ? L : 65535)), 256, 0, st>>>(dD, ldd, L, bw, d_q, dK0, ldk);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int expand_launch(const double* dK0, int64_t ldk, int64_t L, const int* d_gid, const int* d_env, int64_t n, int mode,
                  int kenv, double* dOut, int64_t ldo, cudaStream_t st) {
  if (n <= 0) return OK;
  BGGE_REQUIRE(mode >= 0 && mode <= 3, "expand: unknown mode %d", mode);
  const unsigned gy = (unsigned)((n + 15) / 16);
  BGGE_REQUIRE(gy <= 65535, "expand: n=%lld exceeds grid limit", (long long)n);
  if (ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(dOut) & 15u) == 0 && !getenv("BGGE_EXPAND_SCALAR")) {
    dim3 grid2((unsigned)((n + 511) / 512), gy);
    // large outputs: columns in genotype order (stream-ordered scratch: L + 1 bucket cursors, n column indices)
    int64_t sort_min = 2048;
    if (const char* ev = getenv("BGGE_EXPAND_SORT_MIN")) sort_min = atoll(ev);   // test / tuning hook
    int* scratch = nullptr;
    int* jord = nullptr;
    if (L > 0 && L < (1ll << 30) && n >= sort_min && n < (1ll << 31)) {
      {   // keep the stream-ordered pool's memory across synchronisations (default: released at every sync)
        static bool tuned[64] = {};
        int dev = 0;
        BGGE_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !tuned[dev]) {
          cudaMemPool_t mp;
          if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
            unsigned long long keep = 64ull << 20;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
          }
          cudaGetLastError();
          tuned[dev] = true;
        }
      }
      BGGE_CUDA(cudaMallocAsync((void**)&scratch, (size_t)(L + 1 + n) * sizeof(int), st));
      jord = scratch + (L + 1);
      const unsigned gb = (unsigned)((n + 255) / 256);
      BGGE_CUDA(cudaMemsetAsync(scratch, 0, (size_t)(L + 1) * sizeof(int), st));
      expand_order_count_kernel<<<gb, 256, 0, st>>>(d_gid, n, (int)L, scratch);
      expand_order_scan_kernel<<<1, 1024, 0, st>>>(scratch, (int)L + 1);
      expand_order_fill_kernel<<<gb, 256, 0, st>>>(d_gid, n, (int)L, scratch, jord);
    }
    switch (mode) {
      case 0: expand2_kernel<0><<<grid2, 256, 0, st>>>(dK0, ldk, d_gid, d_env, n, kenv, dOut, ldo, jord); break;
      case 1: expand2_kernel<1><<<grid2, 256, 0, st>>>(dK0, ldk, d_gid, d_env, n, kenv, dOut, ldo, jord); break;
      case 2: expand2_kernel<2><<<grid2, 256, 0, st>>>(dK0, ldk, d_gid, d_env, n, kenv, dOut, ldo, jord); break;
      default: expand2_kernel<3><<<grid2, 256, 0, st>>>(dK0, ldk, d_gid, d_env, n, kenv, dOut, ldo, jord); break;
    }
    const cudaError_t le = cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, st);
    BGGE_CUDA(le);
    return OK;
  }
  dim3 grid((unsigned)((n + 255) / 256), gy);
  switch (mode) {
    case 0: expand_kernel<0><<<grid, 256, 0, st>>>(dK0, ldk, d_gid, d_env, n, kenv, dOut, ldo); break;
    case 1: expand_kernel<1><<<grid, 256, 0, st>>>(dK0, ldk, d_gid, d_env, n, kenv, dOut, ldo); break;
    case 2: expand_kernel<2><<<grid, 256, 0, st>>>(dK0, ldk, d_gid, d_env, n, kenv, dOut, ldo); break;
    default: expand_kernel<3><<<grid, 256, 0, st>>>(dK0, ldk, d_gid, d_env, n, kenv, dOut, ldo); break;
  }
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

}  // namespace bgge

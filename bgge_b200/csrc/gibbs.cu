// Device-resident Gibbs sweep of BGGE() (R/BGGE.R:274-403).  One iteration is the launch sequence
//   [xf_reduce, xf_draw, xf_apply]   fixed effects (only when XF is given)      R/BGGE.R:284-290
//   for each kernel j:  proj_kernel   d = U'(e + u_j),  b ~ N(tau v d, v)        R/BGGE.R:297-305 / 314-330
//                       back_kernel   u_j = U b,  e -= u_j,  Welford(u_j)        R/BGGE.R:306-307 / 332-339, 392
//   scalar_kernel       sigb_j, sigsq, tau, NA imputation, chain record, next mu R/BGGE.R:360-395, 278-281
// captured once into a CUDA graph and replayed `ite` times; nothing returns to the host in between.
//
// Residual bookkeeping: the device keeps r = temp + (mu - mu0), i.e. the residual without the
// *current* intercept, plus the scalar shift = mu - mu0, so the intercept step (R/BGGE.R:278-281)
// needs no vector pass: e_true = r - shift.
//
// Both U passes stream column-major U with threads along rows (coalesced 16-byte loads that
// bypass L1).  proj: a CTA owns CT whole eigen-columns, so d (and therefore b) is final inside the
// CTA and the draw is fused.  back: a thread-block cluster splits the columns of one 256-row tile
// CS ways and the leader CTA reduces the partial sums through distributed shared memory in a
// fixed order (deterministic, no atomics).
#include "gibbs.cuh"
#include "kernels.h"
#include "rng.cuh"
#include <cooperative_groups.h>
#include <cmath>
#include <algorithm>

namespace cg = cooperative_groups;

namespace bgge {
namespace {

constexpr int A_THREADS = 256;
constexpr int B_THREADS = 128;              // 2 rows per thread -> 256-row tiles
constexpr int B_ROWS = 2 * B_THREADS;
constexpr int S_THREADS = 1024;

__device__ __forceinline__ long long z_pos(const DrawCfg& dc, long long it, long long off) {
  return dc.z_init + (it - 1) * dc.z_per_it + off;
}

// ------------------------------------------------------------------ init
// u_j = z / (2n) (R/BGGE.R:254); r = ((y - mu0) - XF Bet) - (u_1 + u_2 + ...) (R/BGGE.R:260-267)
__global__ void init_state_kernel(long long n, int nk, int q, const double* __restrict__ y,
                                  const double* __restrict__ xf, const double* __restrict__ bet,
                                  double* __restrict__ u, double* __restrict__ r, double mu0, DrawCfg dc) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sd = 1.0 / (2.0 * (double)n);
  double usum = 0.0;
  for (int j = 0; j < nk; ++j) {
    const long long idx = (long long)j * n + i;
    const double z = draw_normal(dc, 0, STREAM_INIT_U, (unsigned long long)idx, idx);
    const double uj = z * sd;
    u[idx] = uj;
    usum = (j == 0) ? uj : usum + uj;
  }
  double t = y[i] - mu0;
  if (q > 0) {
    double xb = 0.0;
    for (int k = 0; k < q; ++k) xb += xf[i + (long long)k * n] * bet[k];
    t = t - xb;
  }
  r[i] = t - usum;
}

// ------------------------------------------------------------------ projection + coefficient draw
template <int CT>
__global__ void __launch_bounds__(A_THREADS)
proj_kernel(const ATask* __restrict__ tasks, const double* __restrict__ r, const double* __restrict__ u,
            const double* __restrict__ s, double* __restrict__ b, double* __restrict__ bpart,
            const Scal* __restrict__ sc, int j, DrawCfg dc, long long z_off) {
  const ATask t = tasks[blockIdx.x];
  const double shift = sc->shift;
  double acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = 0.0;
  const double* rp = r + t.pos0;
  const double* up = u + t.pos0;
  const double* Ucol = t.U;
  for (int row = 2 * threadIdx.x; row < t.rows; row += 2 * A_THREADS) {
    const bool two = row + 1 < t.rows;
    double2 x[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c)
      x[c] = (c < t.ncols) ? ldg_stream2(Ucol + (long long)c * t.ldu + row) : make_double2(0.0, 0.0);
    const double e0 = (rp[row] - shift) + up[row];
    const double e1 = two ? (rp[row + 1] - shift) + up[row + 1] : 0.0;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      acc[c] = fma(x[c].x, e0, acc[c]);
      if (two) acc[c] = fma(x[c].y, e1, acc[c]);
    }
  }
  __shared__ double red[A_THREADS / 32][CT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    const double v = warp_sum(acc[c]);
    if (lane == 0) red[warp][c] = v;
  }
  __syncthreads();
  if (warp == 0) {
    double zq = 0.0;
    if (lane < t.ncols) {
      double d = 0.0;
#pragma unroll
      for (int w = 0; w < A_THREADS / 32; ++w) d += red[w][lane];
      const long long col = (long long)t.col_off + lane;
      const double sv = s[col];
      const double lam = sc->sigb[j];
      const double tau = sc->tau;
      const double vari = sv * lam / (1.0 + sv * lam * tau);        // R/BGGE.R:302
      const double media = tau * vari * d;                          // R/BGGE.R:303
      const long long it = sc->it;
      const double z = draw_normal(dc, it, (uint32_t)j, (unsigned long long)col, z_pos(dc, it, z_off + col));
      const double bv = media + sqrt(vari) * z;                     // R/BGGE.R:89-92
      b[col] = bv;
      zq = bv * bv * (1.0 / sv);                                    // R/BGGE.R:96
    }
    zq = warp_sum(zq);
    if (lane == 0) bpart[blockIdx.x] = zq;
  }
}

// ------------------------------------------------------------------ back-transform + residual update
__global__ void __launch_bounds__(B_THREADS)
back_kernel(const BTask* __restrict__ tasks, const double* __restrict__ b, double* __restrict__ r,
            double* __restrict__ u, double* __restrict__ umean, double* __restrict__ um2,
            const Scal* __restrict__ sc, long long thin, long long burn, int cs) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double2 part[B_THREADS];
  const BTask t = tasks[blockIdx.x];
  const int rank = blockIdx.y;
  const int row = 2 * threadIdx.x;
  double a0 = 0.0, a1 = 0.0;
  if (t.nr > 0 && row < t.rows) {
    int chunk = (t.nr + cs - 1) / cs;
    chunk = (chunk + 7) & ~7;
    const int c0 = rank * chunk;
    const int c1 = min(t.nr, c0 + chunk);
    const double* Up = t.U + t.urow0 + row;
    const double* bp = b + t.col_off;
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
      double2 x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = ldg_stream2(Up + (long long)(c + k) * t.ldu);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double bv = bp[c + k];
        a0 = fma(x[k].x, bv, a0);
        a1 = fma(x[k].y, bv, a1);
      }
    }
    for (; c < c1; ++c) {
      const double2 x = ldg_stream2(Up + (long long)c * t.ldu);
      const double bv = bp[c];
      a0 = fma(x.x, bv, a0);
      a1 = fma(x.y, bv, a1);
    }
  }
  part[threadIdx.x] = make_double2(a0, a1);
  cluster.sync();
  if (rank == 0 && row < t.rows) {
    double s0 = a0, s1 = a1;
    for (int p = 1; p < cs; ++p) {
      const double2* rp = cluster.map_shared_rank(part, p);
      const double2 v = rp[threadIdx.x];
      s0 += v.x;
      s1 += v.y;
    }
    const long long it = sc->it;
    const bool keep = (it % thin == 0) && (it > burn);
    const double kcount = keep ? (double)(it / thin - burn / thin) : 1.0;
    const long long gi = (long long)t.row0 + row;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (row + h >= t.rows) break;
      const double un = h ? s1 : s0;
      const long long g = gi + h;
      r[g] = (r[g] + u[g]) - un;                                    // R/BGGE.R:297,307
      u[g] = un;
      if (keep) {                                                   // running mean / M2 of kept draws
        const double m = umean[g];
        const double dlt = un - m;
        const double mn = m + dlt / kcount;
        umean[g] = mn;
        um2[g] += dlt * (un - mn);
      }
    }
  }
  cluster.sync();   // peers' shared memory must stay alive until the leader has read it
}

// ------------------------------------------------------------------ fixed effects
__global__ void xf_reduce_kernel(long long n, int q, const double* __restrict__ xf, const double* __restrict__ r,
                                 const double* __restrict__ bet, const Scal* __restrict__ sc,
                                 double* __restrict__ part) {
  __shared__ double red[8];
  const double shift = sc->shift;
  for (int k = 0; k < q; ++k) {
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      double xb = 0.0;
      for (int k2 = 0; k2 < q; ++k2) xb += xf[i + (long long)k2 * n] * bet[k2];
      const double t = (r[i] - shift) + xb;                         // temp + XF Bet  (R/BGGE.R:285)
      acc = fma(xf[i + (long long)k * n], t, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
      part[(long long)blockIdx.x * q + k] = v;
    }
    __syncthreads();
  }
}

// media = tXX XF'temp; Bet = media + chol(sigsq tXX)' z   (R/BGGE.R:286-288, 106-109)
__global__ void xf_draw_kernel(int q, int nblocks, const double* __restrict__ part, const double* __restrict__ txx,
                               const double* __restrict__ lchol, double* __restrict__ bet, double* __restrict__ bet_old,
                               const Scal* __restrict__ sc, DrawCfg dc) {
  extern __shared__ double sh[];   // v[q], z[q]
  double* v = sh;
  double* z = sh + q;
  const long long it = sc->it;
  for (int k = threadIdx.x; k < q; k += blockDim.x) {
    double a = 0.0;
    for (int bk = 0; bk < nblocks; ++bk) a += part[(long long)bk * q + k];
    v[k] = a;
    z[k] = draw_normal(dc, it, STREAM_XF, (unsigned long long)k, z_pos(dc, it, dc.z_off_xf + k));
  }
  __syncthreads();
  const double sd = sqrt(sc->sigsq);
  for (int k = threadIdx.x; k < q; k += blockDim.x) {
    double media = 0.0, lz = 0.0;
    for (int k2 = 0; k2 < q; ++k2) media += txx[k + (long long)k2 * q] * v[k2];
    for (int k2 = 0; k2 <= k; ++k2) lz += lchol[k + (long long)k2 * q] * z[k2];
    bet_old[k] = bet[k];
    bet[k] = media + sd * lz;
  }
}

__global__ void xf_apply_kernel(long long n, int q, const double* __restrict__ xf, double* __restrict__ r,
                                const double* __restrict__ bet, const double* __restrict__ bet_old) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xo = 0.0, xn = 0.0;
  for (int k = 0; k < q; ++k) {
    const double x = xf[i + (long long)k * n];
    xo += x * bet_old[k];
    xn += x * bet[k];
  }
  r[i] = (r[i] + xo) - xn;                                          // R/BGGE.R:285,289
}

// ------------------------------------------------------------------ end-of-iteration scalars
struct ScalarArgs {
  long long n;
  int nk, q;
  long long n_na, thin, ncum;
  double* r;
  double* u;
  double* y;
  const double* xf;
  const double* bet;
  const int* na_idx;
  Scal* sc;
  double* chain_mu;
  double* chain_varE;
  double* chain_varU;
  double* chain_bet;
  const double* bpart[MAXK];
  int n_bpart[MAXK];
  long long nr[MAXK];
};

__global__ void __launch_bounds__(S_THREADS)
scalar_kernel(const ScalarArgs a, DrawCfg dc) {
  // Latency matters here (one CTA, every iteration): all draws are made up front by separate warps (they
  // do not depend on the data), the b^2/s partials are summed one kernel per warp, and the two vector
  // reductions (sum e^2 for sigsq, sum r for the next intercept) share one pass and one block reduction.
  __shared__ double sh_a[33], sh_b[33];
  __shared__ double sh_chi[MAXK + 1];
  __shared__ double sh_zmu, sh_sigsq;
  // No early pdl_trigger() here: the launches of the next iteration start only when this kernel has completed.
  // That is what makes the reads below safe before pdl_wait(): it / shift / mu were written by the previous
  // scalar_kernel, and none of the launches since then (hence this one) could start before that kernel was done.
  // (With an early trigger a single-kernel model could cascade scalar -> gather -> sweep -> finalize -> scalar
  // while the first scalar kernel is still running.)
  Scal* sc = a.sc;
  const long long it = sc->it;        // iteration that just finished (0 right after init)
  const double shift = sc->shift;
  const double mu = sc->mu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;

  // ---- draws (warp w, lane 0): chi-squares of this iteration, normal of the next intercept
  if (lane == 0) {
    if (it > 0) {
      for (int j = warp; j <= a.nk; j += nwarps) {
        if (j < a.nk)
          sh_chi[j] = draw_chisq(dc, it, STREAM_CHI_K + (uint32_t)j, (double)a.nr[j] + NU, (it - 1) * dc.c_per_it + j);
        else
          sh_chi[a.nk] = draw_chisq(dc, it, STREAM_CHI_E, (double)a.n + NU, (it - 1) * dc.c_per_it + a.nk);
      }
    }
    if (warp == nwarps - 1) sh_zmu = draw_normal(dc, it + 1, STREAM_MU, 0ull, z_pos(dc, it + 1, 0));
  }
  pdl_wait();   // the draws above overlap the tail of the previous launch
  // ---- one pass over r: sum e^2 and sum r
  double sse = 0.0, sr = 0.0;
  for (long long i = threadIdx.x; i < a.n; i += blockDim.x) {
    const double rv = a.r[i];
    const double e = rv - shift;
    sse = fma(e, e, sse);
    sr += rv;
  }
  sse = warp_sum(sse);
  sr = warp_sum(sr);
  if (lane == 0) {
    sh_a[warp] = sse;
    sh_b[warp] = sr;
  }
  __syncthreads();   // also publishes sh_chi / sh_zmu
  if (it > 0) {
    // variance of each random effect (R/BGGE.R:95-98, 360): warp j sums kernel j's partials
    for (int j = warp; j < a.nk; j += nwarps) {
      double zq = 0.0;
      for (int k = lane; k < a.n_bpart[j]; k += 32) zq += a.bpart[j][k];
      zq = warp_sum(zq);
      if (lane == 0) sc->sigb[j] = (zq + NU * sc->Sc[j]) / sh_chi[j];
    }
  }
  if (warp == 0) {
    double ta = lane < nwarps ? sh_a[lane] : 0.0;
    double tb = lane < nwarps ? sh_b[lane] : 0.0;
    ta = warp_sum(ta);
    tb = warp_sum(tb);
    if (lane == 0) {
      sh_b[32] = tb;
      if (it > 0) {   // residual variance (R/BGGE.R:101-103, 364-367)
        const double s2 = (ta + sc->Sce) / sh_chi[a.nk];
        sc->sigsq = s2;
        sc->tau = 1.0 / s2;
        sh_sigsq = s2;
      } else {
        sh_sigsq = sc->sigsq;
      }
    }
  }
  __syncthreads();
  double sr_tot = sh_b[32];
  if (it > 0 && a.n_na > 0) {
    // missing phenotypes (R/BGGE.R:370-381); the change of sum r is reduced separately
    const double sd = sqrt(sh_sigsq);
    double dsum = 0.0;
    for (long long k = threadIdx.x; k < a.n_na; k += blockDim.x) {
      const long long i = a.na_idx[k];
      double uhat = 0.0;
      for (int j = 0; j < a.nk; ++j) uhat = (j == 0) ? a.u[i] : uhat + a.u[(long long)j * a.n + i];
      double aux = 0.0;
      for (int k2 = 0; k2 < a.q; ++k2) aux += a.xf[i + (long long)k2 * a.n] * a.bet[k2];
      const double z = draw_normal(dc, it, STREAM_NA, (unsigned long long)k, z_pos(dc, it, dc.z_off_na + k));
      const double yv = ((aux + mu) + uhat) + sd * z;
      a.y[i] = yv;
      const double rn = (((yv - uhat) - aux) - mu) + shift;
      dsum += rn - a.r[i];
      a.r[i] = rn;
    }
    dsum = warp_sum(dsum);
    if (lane == 0) sh_a[warp] = dsum;
    __syncthreads();
    if (warp == 0) {
      double t = lane < nwarps ? sh_a[lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) sh_a[32] = t;
    }
    __syncthreads();
    sr_tot += sh_a[32];
  }
  if (threadIdx.x == 0) {
    if (it > 0 && it % a.thin == 0) {   // thinned chain (R/BGGE.R:384-395)
      const long long slot = it / a.thin - 1;
      if (slot < a.ncum) {
        a.chain_varE[slot] = sh_sigsq;
        a.chain_mu[slot] = mu;
        for (int j = 0; j < a.nk; ++j) a.chain_varU[(long long)j * a.ncum + slot] = sc->sigb[j];
        for (int k2 = 0; k2 < a.q; ++k2) a.chain_bet[slot * a.q + k2] = a.bet[k2];
      }
    }
    // intercept of the next iteration (R/BGGE.R:278-281): mean(temp + mu) = mean(r) + mu0
    const double nmu = (sr_tot / (double)a.n + sc->mu0) + sqrt(sh_sigsq / (double)a.n) * sh_zmu;
    sc->mu_done = mu;
    sc->mu = nmu;
    sc->shift = nmu - sc->mu0;
    sc->it = it + 1;
  }
}

}  // namespace

// ------------------------------------------------------------------ small host linear algebra (q x q; shared with batch.cu)
bool solve_spd_inverse(int q, std::vector<double> A /*col-major*/, std::vector<double>& inv) {
  // Gauss-Jordan with partial pivoting (same family as LAPACK dgesv behind R's solve())
  inv.assign((size_t)q * q, 0.0);
  for (int i = 0; i < q; ++i) inv[(size_t)i + (size_t)i * q] = 1.0;
  for (int c = 0; c < q; ++c) {
    int piv = c;
    for (int i = c + 1; i < q; ++i)
      if (std::fabs(A[(size_t)i + (size_t)c * q]) > std::fabs(A[(size_t)piv + (size_t)c * q])) piv = i;
    if (A[(size_t)piv + (size_t)c * q] == 0.0) return false;
    if (piv != c)
      for (int k = 0; k < q; ++k) {
        std::swap(A[(size_t)c + (size_t)k * q], A[(size_t)piv + (size_t)k * q]);
        std::swap(inv[(size_t)c + (size_t)k * q], inv[(size_t)piv + (size_t)k * q]);
      }
    const double d = A[(size_t)c + (size_t)c * q];
    for (int k = 0; k < q; ++k) {
      A[(size_t)c + (size_t)k * q] /= d;
      inv[(size_t)c + (size_t)k * q] /= d;
    }
    for (int i = 0; i < q; ++i) {
      if (i == c) continue;
      const double f = A[(size_t)i + (size_t)c * q];
      if (f == 0.0) continue;
      for (int k = 0; k < q; ++k) {
        A[(size_t)i + (size_t)k * q] -= f * A[(size_t)c + (size_t)k * q];
        inv[(size_t)i + (size_t)k * q] -= f * inv[(size_t)c + (size_t)k * q];
      }
    }
  }
  return true;
}

bool chol_lower(int q, const std::vector<double>& A, std::vector<double>& Lw) {
  Lw.assign((size_t)q * q, 0.0);
  for (int j = 0; j < q; ++j) {
    double d = A[(size_t)j + (size_t)j * q];
    for (int k = 0; k < j; ++k) d -= Lw[(size_t)j + (size_t)k * q] * Lw[(size_t)j + (size_t)k * q];
    if (!(d > 0.0)) return false;
    const double dj = std::sqrt(d);
    Lw[(size_t)j + (size_t)j * q] = dj;
    for (int i = j + 1; i < q; ++i) {
      double v = A[(size_t)i + (size_t)j * q];
      for (int k = 0; k < j; ++k) v -= Lw[(size_t)i + (size_t)k * q] * Lw[(size_t)j + (size_t)k * q];
      Lw[(size_t)i + (size_t)j * q] = v / dj;
    }
  }
  return true;
}

namespace {

template <typename T>
int upload(DevBuf<T>& buf, const std::vector<T>& v, cudaStream_t st) {
  BGGE_CUDA(buf.alloc(v.size()));
  if (!v.empty()) BGGE_CUDA(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return OK;
}

int launch_back(Gibbs* g, HostKernel& hk) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)hk.n_btasks, (unsigned)hk.cs, 1);
  cfg.blockDim = dim3(B_THREADS, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = g->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)hk.cs;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int j = (int)(&hk - g->kernels.data());
  BGGE_CUDA(cudaLaunchKernelEx(&cfg, back_kernel, (const BTask*)hk.btasks.p, (const double*)hk.b.p, g->r.p,
                               g->u.p + (size_t)j * g->n, g->umean.p + (size_t)j * g->n,
                               g->um2.p + (size_t)j * g->n, (const Scal*)g->scal.p, (long long)g->thin,
                               (long long)g->burn, hk.cs));
  return OK;
}

int launch_proj(Gibbs* g, HostKernel& hk, int j) {
  const double* up = g->u.p + (size_t)j * g->n;
#define BGGE_PROJ(CTV)                                                                                   \
  proj_kernel<CTV><<<hk.n_atasks, A_THREADS, 0, g->stream>>>(hk.atasks.p, g->r.p, up, hk.s.p, hk.b.p,   \
                                                             hk.bpart.p, g->scal.p, j, g->draw, hk.z_off)
  switch (hk.ct) {
    case 16: BGGE_PROJ(16); break;
    case 8: BGGE_PROJ(8); break;
    case 4: BGGE_PROJ(4); break;
    default: BGGE_PROJ(2); break;
  }
#undef BGGE_PROJ
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int launch_scalar(Gibbs* g, bool merged = false) {
  ScalarArgs a{};
  a.n = g->n;
  a.nk = g->nk;
  a.q = g->q;
  a.n_na = g->n_na;
  a.thin = g->thin;
  a.ncum = g->ncum;
  a.r = g->r.p;
  a.u = g->u.p;
  a.y = g->y.p;
  a.xf = g->xf.p;
  a.bet = g->bet.p;
  a.na_idx = g->na_idx.p;
  a.sc = g->scal.p;
  a.chain_mu = g->chain_mu.p;
  a.chain_varE = g->chain_varE.p;
  a.chain_varU = g->chain_varU.p;
  a.chain_bet = g->chain_bet.p;
  for (int j = 0; j < g->nk; ++j) {
    a.bpart[j] = g->kernels[j].bpart.p;
    a.n_bpart[j] = g->kernels[j].fused ? g->kernels[j].fused_clusters : g->kernels[j].n_atasks;
    if (merged && g->unit_of[(size_t)j] >= 0) a.n_bpart[j] = g->units[(size_t)g->unit_of[(size_t)j]].fused_clusters;
    if (g->shard_world > 1) {   // the all-reduced sum of b^2/s sits behind the reduced vector
      a.bpart[j] = g->kernels[j].xbuf.p + g->n + g->kernels[j].vtotal;
      a.n_bpart[j] = 1;
    }
    a.nr[j] = g->kernels[j].nr_total;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1, 1, 1);
  cfg.blockDim = dim3(S_THREADS, 1, 1);
  cfg.stream = g->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // the early reads of it / shift / mu need a launch between two scalar kernels that waits itself (fused steps)
  bool chain = pdl_enabled() && g->q == 0 && g->initialised;
  for (int j = 0; j < g->nk; ++j) chain = chain && g->kernels[j].fused;
  cfg.numAttrs = chain ? 1 : 0;
  BGGE_CUDA(cudaLaunchKernelEx(&cfg, scalar_kernel, a, g->draw));
  return OK;
}

}  // namespace

int launch_scalar_sharded(Gibbs* g) { return launch_scalar(g, false); }

namespace {

int launch_iteration(Gibbs* g, cudaEvent_t* ev = nullptr, bool merged = false) {
  int ne = 0;
  auto mark = [&]() {
    if (ev) cudaEventRecord(ev[ne++], g->stream);
  };
  mark();
  if (g->q > 0) {
    xf_reduce_kernel<<<g->xf_blocks, 256, 0, g->stream>>>(g->n, g->q, g->xf.p, g->r.p, g->bet.p, g->scal.p,
                                                          g->xf_part.p);
    xf_draw_kernel<<<1, 64, 2 * g->q * sizeof(double), g->stream>>>(g->q, g->xf_blocks, g->xf_part.p, g->txx.p,
                                                                    g->lchol.p, g->bet.p, g->bet.p + g->q, g->scal.p,
                                                                    g->draw);
    xf_apply_kernel<<<(unsigned)((g->n + 255) / 256), 256, 0, g->stream>>>(g->n, g->q, g->xf.p, g->r.p, g->bet.p,
                                                                           g->bet.p + g->q);
    BGGE_CUDA(cudaGetLastError());
    mark();
  }
  for (int j = 0; j < g->nk; ++j) {
    if (merged && g->unit_of[(size_t)j] >= 0) {   // one launch for the whole unit, at its first kernel
      HostKernel& uk = g->units[(size_t)g->unit_of[(size_t)j]];
      if (uk.members[0] == j) {
        BGGE_TRY(launch_fused(g, uk, -1));
        BGGE_TRY(launch_fused_finalize(g, uk, -1));
      }
      continue;
    }
    HostKernel& hk = g->kernels[j];
    if (hk.fused) {
      BGGE_TRY(launch_fused(g, hk, j));   // proj + back in one pass over U (sweep_fused.cu)
      mark();
      BGGE_TRY(launch_fused_finalize(g, hk, j));
      mark();
    } else {
      if (hk.n_atasks > 0) BGGE_TRY(launch_proj(g, hk, j));
      mark();
      BGGE_TRY(launch_back(g, hk));
      mark();
    }
  }
  BGGE_TRY(launch_scalar(g, merged));
  mark();
  return OK;
}

}  // namespace

int ensure_dense(HostBlock& bl, cudaStream_t st) {
  if (bl.vrows == 0 || bl.dU_dense) return OK;
  bl.ldu_dense = round_up(bl.rows, 16);
  BGGE_CUDA(pool_alloc((void**)&bl.dU_dense, (size_t)bl.ldu_dense * bl.nr * sizeof(double)));
  bl.owns_dense = true;
  return expand_rows_launch(bl.dU, bl.ldu, bl.d_loc, bl.d_isq, bl.rows, bl.nr, bl.dU_dense, bl.ldu_dense, st);
}

Gibbs::~Gibbs() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);   // pooled blocks may be reused as soon as they are released
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (graph) cudaGraphDestroy(graph);
  for (auto& k : kernels)
    for (auto& bl : k.blocks)
    {
      if (bl.owns && bl.dU) pool_free(bl.dU);
      if (bl.owns_map) {
        if (bl.d_loc) pool_free(bl.d_loc);
        if (bl.d_isq) pool_free(bl.d_isq);
        if (bl.d_rowptr) pool_free(bl.d_rowptr);
        if (bl.d_rowidx) pool_free(bl.d_rowidx);
      }
      if (bl.owns_dense && bl.dU_dense) pool_free(bl.dU_dense);
    }
  if (own_stream && stream) cudaStreamDestroy(stream);
}

// Build task lists, upload state, run the init step (R/BGGE.R:191-267).
int gibbs_init(Gibbs* g) {
  BGGE_REQUIRE(g->n > 0 && g->nk > 0 && g->nk <= MAXK, "gibbs: need 1..%d kernels and n > 0", MAXK);
  BGGE_REQUIRE((int64_t)g->y_host.size() == g->n, "gibbs: y not set");
  BGGE_REQUIRE(g->ite > 0 && g->thin > 0 && g->burn >= 0, "gibbs: bad ite/burn/thin");
  BGGE_REQUIRE(g->n < (1ll << 31) - 1024, "gibbs: n too large for 32-bit row indices");
  cudaStream_t st = g->stream;
  const int64_t n = g->n;
  const int sms = sm_count();

  // ---- phenotypes: NA -> mean, var(y) on the imputed vector (R/BGGE.R:191-200, 246-256)
  std::vector<double> y = g->y_host;
  std::vector<int> na_idx;
  long double acc = 0.0L;
  int64_t nobs = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (g->na_host.size() == (size_t)n ? g->na_host[(size_t)i] != 0 : std::isnan(y[(size_t)i])) {
      na_idx.push_back((int)i);
    } else {
      acc += y[(size_t)i];
      ++nobs;
    }
  }
  BGGE_REQUIRE(nobs >= 2, "gibbs: need at least two observed phenotypes");
  const double mu0 = (double)(acc / (long double)nobs);
  for (int i : na_idx) y[(size_t)i] = mu0;
  g->n_na = (int64_t)na_idx.size();
  long double m1 = 0.0L;
  for (int64_t i = 0; i < n; ++i) m1 += y[(size_t)i];
  const long double ybar = m1 / (long double)n;
  long double ss = 0.0L;
  for (int64_t i = 0; i < n; ++i) {
    const long double d = (long double)y[(size_t)i] - ybar;
    ss += d * d;
  }
  const double vy = (double)(ss / (long double)(n - 1));

  // ---- fixed effects initial values (R/BGGE.R:208-212)
  const int q = g->q;
  std::vector<double> bet((size_t)2 * q, 0.0), txx, lch;
  if (q > 0) {
    BGGE_REQUIRE((int64_t)g->xf_host.size() == n * q, "gibbs: XF not set");
    std::vector<double> xtx((size_t)q * q, 0.0), xty((size_t)q, 0.0);
    for (int a = 0; a < q; ++a) {
      for (int c = a; c < q; ++c) {
        long double v = 0.0L;
        for (int64_t i = 0; i < n; ++i) v += (long double)g->xf_host[(size_t)(i + a * n)] * g->xf_host[(size_t)(i + c * n)];
        xtx[(size_t)a + (size_t)c * q] = xtx[(size_t)c + (size_t)a * q] = (double)v;
      }
      long double v = 0.0L;
      for (int64_t i = 0; i < n; ++i) v += (long double)g->xf_host[(size_t)(i + a * n)] * y[(size_t)i];
      xty[(size_t)a] = (double)v;
    }
    if (!solve_spd_inverse(q, xtx, txx)) {
      set_error("gibbs: crossprod(XF) is singular");
      return ERR_INVALID;
    }
    for (int a = 0; a < q; ++a) {
      double v = 0.0;
      for (int c = 0; c < q; ++c) v += txx[(size_t)a + (size_t)c * q] * xty[(size_t)c];
      bet[(size_t)a] = v;
    }
    if (!chol_lower(q, txx, lch)) {
      set_error("gibbs: solve(crossprod(XF)) is not positive definite");
      return ERR_INVALID;
    }
  }

  // ---- per-kernel task lists
  long long z_off = 1 + q;
  long long sum_nr = 0;
  int max_clusters = 0;
  int64_t max_vpart = 0;
  for (int j = 0; j < g->nk; ++j) {
    HostKernel& hk = g->kernels[j];
    BGGE_REQUIRE(!hk.blocks.empty(), "gibbs: kernel %d has no eigen blocks (all eigenvalues <= tol)", j);
    {   // blocks in row order, eigenvalues re-concatenated to match (col_off = insertion offset until here)
      int64_t tot = 0;
      for (auto& bl : hk.blocks) tot += bl.nr;
      BGGE_REQUIRE((int64_t)hk.s_host.size() == tot, "gibbs: kernel %d eigenvalue count mismatch", j);
      std::sort(hk.blocks.begin(), hk.blocks.end(), [](const HostBlock& a, const HostBlock& b) { return a.pos0 < b.pos0; });
      std::vector<double> ss;
      ss.reserve((size_t)tot);
      for (auto& bl : hk.blocks) ss.insert(ss.end(), hk.s_host.begin() + bl.col_off, hk.s_host.begin() + bl.col_off + bl.nr);
      hk.s_host.swap(ss);
    }
    hk.nr_total = 0;
    int64_t cursor = 0;
    int64_t max_nr = 0;
    for (auto& bl : hk.blocks) {
      BGGE_REQUIRE(bl.pos0 >= cursor && bl.pos0 + bl.rows <= n, "gibbs: kernel %d has overlapping / out-of-range blocks", j);
      BGGE_REQUIRE(bl.nr > 0 && bl.ldu % 2 == 0 && bl.ldu >= bl.stream_rows(), "gibbs: kernel %d block layout invalid", j);
      bl.col_off = hk.nr_total;
      hk.nr_total += bl.nr;
      max_nr = std::max(max_nr, bl.nr);
      cursor = bl.pos0 + bl.rows;
    }
    BGGE_REQUIRE((int64_t)hk.s_host.size() == hk.nr_total, "gibbs: kernel %d eigenvalue count mismatch", j);
    BGGE_TRY(upload(hk.s, hk.s_host, st));
    BGGE_CUDA(hk.b.alloc((size_t)hk.nr_total));
    hk.z_off = z_off;                                        // the fused panels carry these
    BGGE_CUDA(hk.bpart.alloc((size_t)std::max(1, sms)));
    BGGE_CUDA(cudaMemsetAsync(hk.bpart.p, 0, (size_t)std::max(1, sms) * sizeof(double), st));
    if (g->mode >= 1) BGGE_TRY(plan_fused(g, hk, j, st));
    hk.n_atasks = hk.n_btasks = 0;
    BGGE_REQUIRE(g->shard_world == 1 || hk.fused, "column-sharded chain: kernel %d does not fit the fused sweep", j);
    if (!hk.fused) {
      // two-pass kernels work on dense U: materialise factorised blocks
      for (auto& bl : hk.blocks) BGGE_TRY(ensure_dense(bl, st));
      std::vector<BTask> bt;
      auto add_rows = [&](const HostBlock* bl, int64_t pos0, int64_t rows) {
        for (int64_t r0 = 0; r0 < rows; r0 += B_ROWS) {
          BTask t{};
          t.U = bl ? (bl->vrows > 0 ? bl->dU_dense : bl->dU) : nullptr;
          t.ldu = bl ? (bl->vrows > 0 ? bl->ldu_dense : bl->ldu) : 0;
          t.rows = (int)std::min<int64_t>(B_ROWS, rows - r0);
          t.row0 = (int)(pos0 + r0);
          t.urow0 = (int)r0;
          t.nr = bl ? (int)bl->nr : 0;
          t.col_off = bl ? (int)bl->col_off : 0;
          bt.push_back(t);
        }
      };
      int64_t cur2 = 0;
      for (auto& bl : hk.blocks) {
        if (bl.pos0 > cur2) add_rows(nullptr, cur2, bl.pos0 - cur2);   // rows no block covers
        add_rows(&bl, bl.pos0, bl.rows);
        cur2 = bl.pos0 + bl.rows;
      }
      if (cur2 < n) add_rows(nullptr, cur2, n - cur2);
      // columns per proj CTA: largest CT that still gives >= 2 CTAs per SM
      hk.ct = 2;
      for (int ct : {16, 8, 4}) {
        int64_t tasks = 0;
        for (auto& bl : hk.blocks) tasks += (bl.nr + ct - 1) / ct;
        if (tasks >= 2 * sms) { hk.ct = ct; break; }
      }
      std::vector<ATask> at;
      for (auto& bl : hk.blocks) {
        const double* Ud = bl.vrows > 0 ? bl.dU_dense : bl.dU;
        const int64_t ld = bl.vrows > 0 ? bl.ldu_dense : bl.ldu;
        for (int64_t c = 0; c < bl.nr; c += hk.ct) {
          ATask t{};
          t.U = Ud + c * ld;
          t.ldu = ld;
          t.rows = (int)bl.rows;
          t.pos0 = (int)bl.pos0;
          t.ncols = (int)std::min<int64_t>(hk.ct, bl.nr - c);
          t.col_off = (int)(bl.col_off + c);
          at.push_back(t);
        }
      }
      // column split of the back-transform: enough CTAs to fill the machine, >= 32 columns each
      int cs = 1;
      while (cs < 8 && (int64_t)bt.size() * cs < 4 * sms && max_nr / (cs * 2) >= 32) cs *= 2;
      hk.cs = cs;
      hk.n_atasks = (int)at.size();
      hk.n_btasks = (int)bt.size();
      BGGE_TRY(upload(hk.atasks, at, st));
      BGGE_TRY(upload(hk.btasks, bt, st));
    }
    max_clusters = std::max(max_clusters, hk.fused ? hk.fused_clusters : 0);
    max_vpart = std::max<int64_t>(max_vpart, hk.fused ? hk.vtotal * hk.fused_clusters : 0);
    if (std::max(hk.n_atasks, hk.fused_clusters) > std::max(1, sms))
      BGGE_CUDA(hk.bpart.alloc((size_t)std::max(hk.n_atasks, hk.fused_clusters)));   // two-pass task lists only
    z_off += hk.nr_total;
    sum_nr += hk.nr_total;
  }

  // ---- merged units: consecutive single-block kernels sharing one matrix on disjoint rows (the environment kernels
  // of MDe) are conditionally independent and are swept by one multi-right-hand-side launch from iteration 2 on
  g->units.clear();
  g->unit_of.assign((size_t)g->nk, -1);
  if (g->mode >= 1 && g->shard_world == 1 && !getenv("BGGE_NO_MERGED_STEPS")) {
    for (int j = 0; j < g->nk;) {
      std::vector<int> set{j};
      const HostKernel& h0 = g->kernels[(size_t)j];
      if (h0.fused && h0.blocks.size() == 1) {
        const HostBlock& b0 = h0.blocks[0];
        while ((int)set.size() < 4) {
          const int jn = j + (int)set.size();
          if (jn >= g->nk || !g->kernels[(size_t)jn].fused || g->kernels[(size_t)jn].blocks.size() != 1) break;
          const HostBlock& bn = g->kernels[(size_t)jn].blocks[0];
          if (bn.dU != b0.dU || bn.nr != b0.nr || bn.rows != b0.rows || bn.vrows != b0.vrows || bn.ldu != b0.ldu) break;
          if (g->kernels[(size_t)jn].s_host != h0.s_host) break;   // same matrix, same eigenvalues
          bool disjoint = true;
          for (int kk : set) {
            const HostBlock& bk = g->kernels[(size_t)kk].blocks[0];
            if (bn.pos0 < bk.pos0 + bk.rows && bk.pos0 < bn.pos0 + bn.rows) disjoint = false;
          }
          if (!disjoint) break;
          set.push_back(jn);
        }
      }
      if (set.size() > 1) {
        HostKernel uk;
        uk.merged = true;
        uk.members = set;
        for (int kk : set) {
          HostBlock bl = g->kernels[(size_t)kk].blocks[0];
          bl.owns = bl.owns_map = bl.owns_dense = false;
          bl.dU_dense = nullptr;
          bl.kid = kk;
          uk.blocks.push_back(bl);
        }
        std::sort(uk.blocks.begin(), uk.blocks.end(), [](const HostBlock& a, const HostBlock& b) { return a.pos0 < b.pos0; });
        BGGE_TRY(plan_fused(g, uk, -1, st));
        if (uk.fused && uk.fused_clusters <= std::max(1, sms)) {
          for (int kk : set) g->unit_of[(size_t)kk] = (int)g->units.size();
          max_clusters = std::max(max_clusters, uk.fused_clusters);
          max_vpart = std::max<int64_t>(max_vpart, uk.vtotal * uk.fused_clusters);
          g->units.push_back(std::move(uk));
        }
      }
      j += (int)set.size();
    }
  }

  // ---- draw configuration / tape bookkeeping
  g->draw.z_init = (long long)g->nk * n;
  g->draw.z_per_it = 1 + q + sum_nr + g->n_na;
  g->draw.z_off_xf = 1;
  g->draw.z_off_na = 1 + q + sum_nr;
  g->draw.c_per_it = g->nk + 1;
  g->draw.zt = g->ztape.p;
  g->draw.ct = g->ctape.p;
  if (g->draw.mode == 1) {
    // the intercept of iteration ite+1 is drawn ahead, hence the +1
    const long long need_z = g->draw.z_init + g->ite * g->draw.z_per_it + 1;
    const long long need_c = g->ite * g->draw.c_per_it;
    if (g->zt_len < need_z - 1 || g->ct_len < need_c) {
      set_error("gibbs: injected tape too short (normals %lld < %lld or chisq %lld < %lld)", (long long)g->zt_len,
                need_z - 1, (long long)g->ct_len, need_c);
      return ERR_TAPE;
    }
  }

  // ---- device state
  g->ncum = g->ite / g->thin;                                       // R/BGGE.R:229
  BGGE_CUDA(g->r.alloc((size_t)n));
  if (max_clusters > 0) BGGE_CUDA(g->upart.alloc((size_t)n * max_clusters));
  if (max_vpart > 0) BGGE_CUDA(g->vupart.alloc((size_t)max_vpart));
  if (g->shard_world > 1) {
    BGGE_REQUIRE(q == 0, "column-sharded chain: fixed effects are not supported");
    for (int j = 0; j < g->nk; ++j) BGGE_TRY(plan_shard(g, g->kernels[(size_t)j], st));
  }
  BGGE_CUDA(g->u.alloc((size_t)n * g->nk));
  BGGE_CUDA(g->umean.alloc((size_t)n * g->nk));
  BGGE_CUDA(g->um2.alloc((size_t)n * g->nk));
  BGGE_CUDA(cudaMemsetAsync(g->umean.p, 0, (size_t)n * g->nk * sizeof(double), st));
  BGGE_CUDA(cudaMemsetAsync(g->um2.p, 0, (size_t)n * g->nk * sizeof(double), st));
  BGGE_TRY(upload(g->y, y, st));
  BGGE_TRY(upload(g->na_idx, na_idx, st));
  const size_t nc = (size_t)std::max<int64_t>(1, g->ncum);
  BGGE_CUDA(g->chain_mu.alloc(nc));
  BGGE_CUDA(g->chain_varE.alloc(nc));
  BGGE_CUDA(g->chain_varU.alloc(nc * g->nk));
  BGGE_CUDA(g->chain_bet.alloc(nc * std::max(1, q)));
  BGGE_CUDA(cudaMemsetAsync(g->chain_mu.p, 0, nc * sizeof(double), st));
  BGGE_CUDA(cudaMemsetAsync(g->chain_varE.p, 0, nc * sizeof(double), st));
  BGGE_CUDA(cudaMemsetAsync(g->chain_varU.p, 0, nc * g->nk * sizeof(double), st));
  BGGE_CUDA(cudaMemsetAsync(g->chain_bet.p, 0, nc * std::max(1, q) * sizeof(double), st));
  if (q > 0) {
    BGGE_TRY(upload(g->xf, g->xf_host, st));
    BGGE_TRY(upload(g->bet, bet, st));
    BGGE_TRY(upload(g->txx, txx, st));
    BGGE_TRY(upload(g->lchol, lch, st));
    g->xf_blocks = (int)std::min<int64_t>(2 * sms, (n + 255) / 256);
    BGGE_CUDA(g->xf_part.alloc((size_t)g->xf_blocks * q));
  }
  Scal sc{};
  sc.mu = mu0;
  sc.mu_done = mu0;
  sc.mu0 = mu0;
  sc.shift = 0.0;
  sc.sigsq = vy;                                                    // R/BGGE.R:256
  sc.tau = 0.01;                                                    // R/BGGE.R:251
  sc.Sce = (NU + 2.0) * (1.0 - g->R2) * vy;                         // R/BGGE.R:246
  for (int j = 0; j < g->nk; ++j) {
    sc.sigb[j] = 0.2;                                               // R/BGGE.R:258
    sc.Sc[j] = (NU + 2.0) * g->R2 * vy / g->kernels[j].mean_diag;   // R/BGGE.R:249
  }
  sc.it = 0;
  BGGE_CUDA(g->scal.alloc(1));
  BGGE_CUDA(cudaMemcpyAsync(g->scal.p, &sc, sizeof(Scal), cudaMemcpyHostToDevice, st));
  init_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, g->nk, q, g->y.p, g->xf.p, g->bet.p, g->u.p,
                                                                g->r.p, mu0, g->draw);
  BGGE_CUDA(cudaGetLastError());
  BGGE_TRY(launch_scalar(g));   // it = 0: draws the intercept of iteration 1
  BGGE_CUDA(cudaStreamSynchronize(st));
  BGGE_TRY(plan_resident(g));
  g->done = 0;
  g->initialised = true;
  return OK;
}

int gibbs_run(Gibbs* g, int64_t iters) {
  if (!g->initialised) {
    set_error("gibbs: run before init");
    return ERR_STATE;
  }
  BGGE_REQUIRE(iters >= 0 && g->done + iters <= g->ite, "gibbs: asked for %lld iterations, %lld of %lld already done",
               (long long)iters, (long long)g->done, (long long)g->ite);
  BGGE_REQUIRE(g->shard_world == 1, "gibbs: a column-sharded chain is driven step by step (bgge_gibbs_shard_begin / _end)");
  if (g->resident.enabled) {   // the whole run is one cooperative launch
    if (g->resident.has_merged && g->done == 0 && iters > 0) {
      // merged steps rely on u_k = 0 outside kernel k's block, true once every kernel has been swept once
      BGGE_TRY(launch_iteration(g));
      g->done += 1;
      iters -= 1;
    }
    if (iters > 0) {
      const int rs = launch_resident(g, iters);
      if (rs != ERR_RETRY_STREAMING) {
        if (rs != OK) return rs;
        g->done += iters;
        return OK;
      }
      // fall through to the streaming path with the remaining iterations
    } else {
      return OK;
    }
  }
  const bool merged = !g->units.empty();
  if (merged && g->done == 0 && iters > 0) {
    // merged units rely on u_k = 0 outside kernel k's block, true once every kernel has been swept once
    BGGE_TRY(launch_iteration(g));
    g->done += 1;
    iters -= 1;
  }
  if (g->use_graph && !g->graph_exec) {
    BGGE_CUDA(cudaStreamBeginCapture(g->stream, cudaStreamCaptureModeThreadLocal));
    int s = launch_iteration(g, nullptr, merged);
    cudaError_t ce = cudaStreamEndCapture(g->stream, &g->graph);
    if (s != OK) return s;
    BGGE_CUDA(ce);
    BGGE_CUDA(cudaGraphInstantiate(&g->graph_exec, g->graph, 0));
  }
  for (int64_t i = 0; i < iters; ++i) {
    if (g->use_graph) {
      BGGE_CUDA(cudaGraphLaunch(g->graph_exec, g->stream));
    } else {
      BGGE_TRY(launch_iteration(g, nullptr, merged));
    }
  }
  g->done += iters;
  return OK;
}

// Instrumented (non-graph) sweeps: CUDA events around every launch.  ms layout:
// [proj_0, back_0, proj_1, back_1, ..., scalar, xf]; each entry is the summed device time.
int gibbs_profile(Gibbs* g, int64_t iters, double* ms) {
  if (!g->initialised) {
    set_error("gibbs: profile before init");
    return ERR_STATE;
  }
  BGGE_REQUIRE(iters >= 0 && g->done + iters <= g->ite, "gibbs: asked for %lld iterations, %lld of %lld already done",
               (long long)iters, (long long)g->done, (long long)g->ite);
  BGGE_REQUIRE(g->shard_world == 1, "gibbs: profile is not available for a column-sharded chain");
  const int nslots = 2 * g->nk + 2;
  for (int i = 0; i < nslots; ++i) ms[i] = 0.0;
  const int nev = 2 * g->nk + 3 + (g->q > 0 ? 1 : 0);
  std::vector<cudaEvent_t> ev((size_t)nev);
  for (auto& e : ev) BGGE_CUDA(cudaEventCreate(&e));
  int status = OK;
  for (int64_t i = 0; i < iters && status == OK; ++i) {
    status = launch_iteration(g, ev.data());
    if (status != OK) break;
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) {
      set_error("gibbs_profile: stream synchronize failed");
      status = ERR_CUDA;
      break;
    }
    int k = 0;
    float t = 0.f;
    if (g->q > 0) {
      cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
      ms[2 * g->nk + 1] += t;
      ++k;
    }
    for (int j = 0; j < g->nk; ++j) {
      cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
      ms[2 * j] += t;
      ++k;
      cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
      ms[2 * j + 1] += t;
      ++k;
    }
    cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
    ms[2 * g->nk] += t;
    g->done += 1;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return status;
}

}  // namespace bgge

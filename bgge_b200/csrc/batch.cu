// Batched Gibbs sampler: B independent chains / cross-validation folds that share the eigen blocks (s, U)
// of every kernel.  Each chain runs exactly the single-chain algorithm of
// gibbs.cu (same counter-based draws: key (seed, chain0 + b)), but the two U passes become skinny GEMMs
//     D = U'E (nr x B)   and   Unew = U Bm (rows x B)
// on the FP64 tensor cores (dgemm.cu), with deterministic split-K.  All per-chain vectors are stored as
// row-major n x Bp matrices (chain index fastest) so they are GEMM operands as they are.
#include "batch.cuh"
#include "kernels.h"
#include "rng.cuh"
#include <algorithm>
#include <cmath>

namespace bgge {
namespace {

constexpr int BT = 256;

__device__ __forceinline__ DrawCfg chain_cfg(const DrawCfg& dc, int b) {
  DrawCfg c = dc;
  c.chain = dc.chain + (unsigned)b;
  return c;
}

// u_j = z/(2n), R = (y - mu0) - sum_j u_j          (R/BGGE.R:254, 260-267)
__global__ void b_init_kernel(long long n, int nk, int B, int Bp, const double* __restrict__ Y,
                              const double* __restrict__ mu0, double* __restrict__ U, double* __restrict__ R, DrawCfg dc) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n * Bp) return;
  const long long i = idx / Bp;
  const int b = (int)(idx % Bp);
  if (b >= B) {
    R[idx] = 0.0;
    for (int j = 0; j < nk; ++j) U[(long long)j * n * Bp + idx] = 0.0;
    return;
  }
  const DrawCfg c = chain_cfg(dc, b);
  const double sd = 1.0 / (2.0 * (double)n);
  double usum = 0.0;
  for (int j = 0; j < nk; ++j) {
    const long long k = (long long)j * n + i;
    const double uj = draw_normal(c, 0, STREAM_INIT_U, (unsigned long long)k, 0) * sd;
    U[(long long)j * n * Bp + idx] = uj;
    usum = (j == 0) ? uj : usum + uj;
  }
  R[idx] = (Y[idx] - mu0[b]) - usum;
}

// E' = (R - shift_b) + U_j
__global__ void b_eprime_kernel(long long total, int Bp, const double* __restrict__ R, const double* __restrict__ Uj,
                                const double* __restrict__ shift, double* __restrict__ Ep) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  Ep[idx] = (R[idx] - shift[idx % Bp]) + Uj[idx];
}

// d = sum of K-split partials; b = tau v d + sqrt(v) z; partial sums of b^2/s        (R/BGGE.R:302-305, 96)
// Small kernels of the batched sampler are tiled as (32 chains) x (8 row / column lanes): with few chains per GPU
// (40 at 8 GPUs) a thread-per-chain layout leaves the chip idle behind long serial loops (r02: b_draw 79 us at B = 40).
constexpr int TX = 32, TY = 8;

__global__ void __launch_bounds__(TX * TY)
b_draw_kernel(int nr, int B, int Bp, int nsplit, const double* __restrict__ Dpart, long long dstride,
              const double* __restrict__ s, const double* __restrict__ sigb, const double* __restrict__ tau,
              double* __restrict__ Bm, double* __restrict__ zqpart, const long long* __restrict__ itp, int j, DrawCfg dc) {
  __shared__ double zs[TY][TX];
  const int b = blockIdx.y * TX + threadIdx.x, cy = threadIdx.y;
  const int c0 = blockIdx.x * 32, c1 = min(nr, c0 + 32);
  const long long it = *itp;
  double zq = 0.0;
  if (b < B) {
    const DrawCfg c = chain_cfg(dc, b);
    const double lam = sigb[b], tb = tau[b];
    for (int cc = c0 + cy; cc < c1; cc += TY) {
      double d = 0.0;
      for (int sp = 0; sp < nsplit; ++sp) d += Dpart[(long long)sp * dstride + (long long)cc * Bp + b];
      const double sv = s[cc];
      const double vari = sv * lam / (1.0 + sv * lam * tb);
      const double media = tb * vari * d;
      const double z = draw_normal(c, it, (uint32_t)j, (unsigned long long)cc, 0);
      const double bv = media + sqrt(vari) * z;
      Bm[(long long)cc * Bp + b] = bv;
      zq += bv * bv * (1.0 / sv);
    }
  } else if (b < Bp) {
    for (int cc = c0 + cy; cc < c1; cc += TY) Bm[(long long)cc * Bp + b] = 0.0;
  }
  zs[cy][threadIdx.x] = zq;
  __syncthreads();
  if (cy == 0 && b < Bp) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < TY; ++k) t += zs[k][threadIdx.x];   // fixed order
    zqpart[(long long)blockIdx.x * Bp + b] = t;
  }
}

// u_new = sum of K-split partials; R = (R + u_old) - u_new; Welford of kept draws     (R/BGGE.R:306-307, 392)
__global__ void b_update_kernel(long long total, int Bp, int nsplit, const double* __restrict__ Upart, long long ustride,
                                const unsigned char* __restrict__ cover, double* __restrict__ R, double* __restrict__ Uj, double* __restrict__ Um,
                                double* __restrict__ Um2, const long long* __restrict__ itp, long long thin, long long burn) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  double un = 0.0;   // rows no block of this kernel covers: u = 0 (R/BGGE.R:334); Upart is shared between kernels
  if (cover[idx / Bp])
    for (int sp = 0; sp < nsplit; ++sp) un += Upart[(long long)sp * ustride + idx];
  R[idx] = (R[idx] + Uj[idx]) - un;
  Uj[idx] = un;
  const long long it = *itp;
  if ((it % thin == 0) && (it > burn)) {
    const double k = (double)(it / thin - burn / thin);
    const double m = Um[idx];
    const double dlt = un - m;
    const double mn = m + dlt / k;
    Um[idx] = mn;
    Um2[idx] += dlt * (un - mn);
  }
}

// partial[chunk][b] = sum over a row chunk of (R - shift)^2   (mode 0)  or of R (mode 1)
__global__ void __launch_bounds__(TX * TY)
b_colsum_kernel(long long n, int Bp, int rows_per_chunk, const double* __restrict__ R, const double* __restrict__ shift,
                int mode, double* __restrict__ part) {
  __shared__ double zs[TY][TX];
  const int b = blockIdx.y * TX + threadIdx.x, ry = threadIdx.y;
  const long long i0 = (long long)blockIdx.x * rows_per_chunk;
  const long long i1 = min(n, i0 + rows_per_chunk);
  double a = 0.0;
  if (b < Bp) {
    const double sh = mode == 0 ? shift[b] : 0.0;
    for (long long i = i0 + ry; i < i1; i += TY) {
      const double v = R[i * Bp + b] - sh;
      a = mode == 0 ? fma(v, v, a) : a + v;
    }
  }
  zs[ry][threadIdx.x] = a;
  __syncthreads();
  if (ry == 0 && b < Bp) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < TY; ++k) t += zs[k][threadIdx.x];
    part[(long long)blockIdx.x * Bp + b] = t;
  }
}

// sum of count values p[k * stride] by one warp in a fixed order: lane-strided partial sums, then a xor tree
__device__ __forceinline__ double warp_sum_strided(const double* __restrict__ p, int count, long long stride, int lane) {
  double a = 0.0;
  for (int k = lane; k < count; k += 32) a += p[(long long)k * stride];
  return warp_sum(a);
}

struct ChainArgs {
  long long n;
  int nk, B, Bp, nchunks;
  long long thin, ncum;
  double* mu;
  double* mu0;
  double* shift;
  double* sigsq;
  double* tau;
  const double* Sce;
  double* sigb;        // [nk][Bp]
  const double* Sc;    // [nk][Bp]
  long long* it;
  const double* part;  // [nchunks][Bp]
  const double* zqpart[MAXK];
  int n_zq[MAXK];
  long long nr[MAXK];
  double* chain_mu;
  double* chain_varE;
  double* chain_varU;
};

// per chain (one warp each): sigb_j, sigsq, tau        (R/BGGE.R:95-103, 360-367)
// the sums are formed by the whole warp; then lane j draws the chi-square of kernel j and lane nk the residual one, so
// the Marsaglia-Tsang rejection loops run side by side instead of one after the other
__global__ void __launch_bounds__(128) b_chain_var_kernel(const ChainArgs a, DrawCfg dc) {
  __shared__ double ss[4][MAXK + 1];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + w;
  if (b >= a.B) return;
  const long long it = *a.it;
  const DrawCfg c = chain_cfg(dc, b);
  for (int j = 0; j < a.nk; ++j) {
    const double zq = warp_sum_strided(a.zqpart[j] + b, a.n_zq[j], a.Bp, lane);   // sum b^2/s
    if (lane == 0) ss[w][j] = zq;
  }
  const double sse = warp_sum_strided(a.part + b, a.nchunks, a.Bp, lane);         // e'e
  if (lane == 0) ss[w][a.nk] = sse;
  __syncwarp();
  for (int j = lane; j <= a.nk; j += 32) {
    if (j < a.nk) {
      const double chi = draw_chisq(c, it, STREAM_CHI_K + (uint32_t)j, (double)a.nr[j] + NU, 0);
      a.sigb[(long long)j * a.Bp + b] = (ss[w][j] + NU * a.Sc[(long long)j * a.Bp + b]) / chi;
    } else {
      const double chi = draw_chisq(c, it, STREAM_CHI_E, (double)a.n + NU, 0);
      const double s2 = (ss[w][j] + a.Sce[b]) / chi;
      a.sigsq[b] = s2;
      a.tau[b] = 1.0 / s2;
    }
  }
}

// NA imputation (R/BGGE.R:370-381) fused with the column sums of r for the next intercept
__global__ void __launch_bounds__(TX * TY)
b_na_sumr_kernel(long long n, int nk, int B, int Bp, int rows_per_chunk, double* __restrict__ R,
                 const double* __restrict__ U, double* __restrict__ Y, const int* __restrict__ narank,
                 const double* __restrict__ mu, const double* __restrict__ shift, const double* __restrict__ sigsq,
                 const long long* __restrict__ itp, int do_na, double* __restrict__ part, DrawCfg dc) {
  __shared__ double zs[TY][TX];
  const int b = blockIdx.y * TX + threadIdx.x, ry = threadIdx.y;
  const long long i0 = (long long)blockIdx.x * rows_per_chunk;
  const long long i1 = min(n, i0 + rows_per_chunk);
  const long long it = *itp;
  double a = 0.0;
  if (b < B) {
    const DrawCfg c = chain_cfg(dc, b);
    const double mub = mu[b], shb = shift[b], sd = sqrt(sigsq[b]);
    for (long long i = i0 + ry; i < i1; i += TY) {
      const long long idx = i * Bp + b;
      double rv = R[idx];
      if (do_na) {
        const int k = narank[idx];
        if (k >= 0) {
          double uhat = 0.0;
          for (int j = 0; j < nk; ++j) uhat = (j == 0) ? U[idx] : uhat + U[(long long)j * n * Bp + idx];
          const double z = draw_normal(c, it, STREAM_NA, (unsigned long long)k, 0);
          const double yv = ((0.0 + mub) + uhat) + sd * z;
          Y[idx] = yv;
          rv = (((yv - uhat) - 0.0) - mub) + shb;
          R[idx] = rv;
        }
      }
      a += rv;
    }
  }
  zs[ry][threadIdx.x] = a;
  __syncthreads();
  if (ry == 0 && b < Bp) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < TY; ++k) t += zs[k][threadIdx.x];
    part[(long long)blockIdx.x * Bp + b] = t;
  }
}

// per chain: record the thinned chain (R/BGGE.R:384-395), draw the next intercept (R/BGGE.R:278-281)
__global__ void b_advance_kernel(long long* it) { *it += 1; }

__global__ void __launch_bounds__(128) b_chain_mu_kernel(const ChainArgs a, DrawCfg dc) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= a.B) return;
  const long long it = *a.it;
  const double sr = warp_sum_strided(a.part + b, a.nchunks, a.Bp, lane);
  if (lane != 0) return;
  const DrawCfg c = chain_cfg(dc, b);
  if (it > 0 && it % a.thin == 0) {
    const long long slot = it / a.thin - 1;
    if (slot < a.ncum) {
      a.chain_mu[slot * a.Bp + b] = a.mu[b];
      a.chain_varE[slot * a.Bp + b] = a.sigsq[b];
      for (int j = 0; j < a.nk; ++j) a.chain_varU[((long long)j * a.ncum + slot) * a.Bp + b] = a.sigb[(long long)j * a.Bp + b];
    }
  }
  const double z = draw_normal(c, it + 1, STREAM_MU, 0ull, 0);
  const double nmu = (sr / (double)a.n + a.mu0[b]) + sqrt(a.sigsq[b] / (double)a.n) * z;
  a.mu[b] = nmu;
  a.shift[b] = nmu - a.mu0[b];
}

template <typename T>
int up(DevBuf<T>& buf, const std::vector<T>& v, cudaStream_t st) {
  BGGE_CUDA(buf.alloc(v.size()));
  if (!v.empty()) BGGE_CUDA(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return OK;
}

ChainArgs chain_args(Batch* g, const double* part) {
  ChainArgs a{};
  a.n = g->n;
  a.nk = g->nk;
  a.B = g->B;
  a.Bp = g->Bp;
  a.nchunks = g->nchunks;
  a.thin = g->thin;
  a.ncum = g->ncum;
  a.mu = g->mu.p;
  a.mu0 = g->mu0.p;
  a.shift = g->shift.p;
  a.sigsq = g->sigsq.p;
  a.tau = g->tau.p;
  a.Sce = g->Sce.p;
  a.sigb = g->sigb.p;
  a.Sc = g->Sc.p;
  a.it = g->it.p;
  a.part = part;
  for (int j = 0; j < g->nk; ++j) {
    a.zqpart[j] = g->kernels[j].zqpart.p;
    a.n_zq[j] = g->kernels[j].n_ctiles;
    a.nr[j] = g->kernels[j].nr_total;
  }
  a.chain_mu = g->chain_mu.p;
  a.chain_varE = g->chain_varE.p;
  a.chain_varU = g->chain_varU.p;
  return a;
}

int launch_tail(Batch* g, bool init) {
  cudaStream_t st = g->stream;
  const dim3 cgrid((unsigned)g->nchunks, (unsigned)((g->Bp + TX - 1) / TX));
  const dim3 cblock(TX, TY);
  const unsigned cb = (unsigned)((g->B + 3) / 4);   // one warp per chain
  if (!init) {
    b_colsum_kernel<<<cgrid, cblock, 0, st>>>(g->n, g->Bp, g->rows_per_chunk, g->R.p, g->shift.p, 0, g->part.p);
    b_chain_var_kernel<<<cb, 128, 0, st>>>(chain_args(g, g->part.p), g->draw);
  }
  b_na_sumr_kernel<<<cgrid, cblock, 0, st>>>(g->n, g->nk, g->B, g->Bp, g->rows_per_chunk, g->R.p, g->U.p, g->Y.p, g->narank.p,
                                        g->mu.p, g->shift.p, g->sigsq.p, g->it.p, (!init && g->any_na) ? 1 : 0, g->part2.p,
                                        g->draw);
  b_chain_mu_kernel<<<cb, 128, 0, st>>>(chain_args(g, g->part2.p), g->draw);
  b_advance_kernel<<<1, 1, 0, st>>>(g->it.p);   // separate launch: other CTAs of the kernel above still read it
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int launch_iteration(Batch* g) {
  cudaStream_t st = g->stream;
  const long long total = g->n * g->Bp;
  const unsigned eg = (unsigned)((total + BT - 1) / BT);
  for (int j = 0; j < g->nk; ++j) {
    BatchKernel& bk = g->kernels[j];
    double* Uj = g->U.p + (size_t)j * total;
    b_eprime_kernel<<<eg, BT, 0, st>>>(total, g->Bp, g->R.p, Uj, g->shift.p, g->Ep.p);
    BGGE_TRY(dgemm_launch(bk.proj_items.p, bk.n_proj, g->tn, st));
    b_draw_kernel<<<dim3((unsigned)bk.n_ctiles, (unsigned)((g->Bp + TX - 1) / TX)), dim3(TX, TY), 0, st>>>(
        (int)bk.nr_total, g->B, g->Bp, bk.ksplit_proj, bk.Dpart.p, bk.nr_pad * g->Bp, bk.s.p,
        g->sigb.p + (size_t)j * g->Bp, g->tau.p, bk.Bm.p, bk.zqpart.p, g->it.p, j, g->draw);
    BGGE_TRY(dgemm_launch(bk.back_items.p, bk.n_back, g->tn, st));
    b_update_kernel<<<eg, BT, 0, st>>>(total, g->Bp, bk.ksplit_back, g->Upart.p, total, bk.cover.p, g->R.p, Uj,
                                      g->Um.p + (size_t)j * total, g->Um2.p + (size_t)j * total, g->it.p, g->thin, g->burn);
    BGGE_CUDA(cudaGetLastError());
  }
  return launch_tail(g, false);
}

}  // namespace

Batch::~Batch() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (graph) cudaGraphDestroy(graph);
  for (auto& k : kernels)
    for (auto& bl : k.blocks) {
      if (bl.owns && bl.dU) pool_free(bl.dU);
      if (bl.dTb) pool_free(bl.dTb);
      if (bl.dUb) pool_free(bl.dUb);
    }
  if (own_stream && stream) cudaStreamDestroy(stream);
}

int batch_init(Batch* g) {
  BGGE_REQUIRE(g->n > 1 && g->nk >= 1 && g->nk <= MAXK && g->B >= 1, "batch: bad sizes");
  BGGE_REQUIRE((int64_t)g->y_host.size() == g->n * g->B, "batch: y not set");
  BGGE_REQUIRE(g->ite > 0 && g->thin > 0 && g->burn >= 0, "batch: bad ite/burn/thin");
  cudaStream_t st = g->stream;
  const int64_t n = g->n;
  const int B = g->B;
  // chains are padded to a multiple of 8 (one DMMA fragment) only; the N tile width is the instantiated width that
  // covers Bp with the fewest padded columns, wider tiles first on ties (80 chains -> one 80-wide tile, 320 -> 4 x 80)
  g->Bq = (int)round_up(B, 8);
  g->Bp = g->Bq + 4;
  const int Bp = g->Bp, Bq = g->Bq;
  {
    int best_tn = 128;
    long best_cols = 1L << 40;
    for (int k = kGemmTileNCount - 1; k >= 0; --k) {
      const int tn = kGemmTileN[k];
      const long cols = (long)((Bq + tn - 1) / tn) * tn;
      if (cols < best_cols) {
        best_cols = cols;
        best_tn = tn;
      }
    }
    if (const char* ev = getenv("BGGE_BATCH_TN")) {   // tuning hook
      const int v = atoi(ev);
      for (int k = 0; k < kGemmTileNCount; ++k)
        if (kGemmTileN[k] == v) best_tn = v;
    }
    g->tn = best_tn;
  }
  const int sms = sm_count();
  g->ncum = g->ite / g->thin;

  // ---- per-chain phenotype statistics (R/BGGE.R:191-200, 245-258), matrices in [i][b] layout
  std::vector<double> Y((size_t)n * Bp, 0.0), mu0((size_t)Bp, 0.0), vy((size_t)Bp, 1.0);
  std::vector<int> narank((size_t)n * Bp, -1);
  g->any_na = false;
  for (int b = 0; b < B; ++b) {
    const double* yb = g->y_host.data() + (size_t)b * n;
    const unsigned char* nb = g->na_host.empty() ? nullptr : g->na_host.data() + (size_t)b * n;
    long double acc = 0.0L;
    int64_t nobs = 0;
    for (int64_t i = 0; i < n; ++i) {
      const bool na = nb ? nb[i] != 0 : std::isnan(yb[i]);
      if (!na) {
        acc += yb[i];
        ++nobs;
      }
    }
    BGGE_REQUIRE(nobs >= 2, "batch: chain %d has fewer than two observed phenotypes", b);
    const double m0 = (double)(acc / (long double)nobs);
    int rank = 0;
    long double m1 = 0.0L;
    for (int64_t i = 0; i < n; ++i) {
      const bool na = nb ? nb[i] != 0 : std::isnan(yb[i]);
      const double v = na ? m0 : yb[i];
      Y[(size_t)i * Bp + b] = v;
      if (na) {
        narank[(size_t)i * Bp + b] = rank++;
        g->any_na = true;
      }
      m1 += v;
    }
    const long double ybar = m1 / (long double)n;
    long double ss = 0.0L;
    for (int64_t i = 0; i < n; ++i) {
      const long double d = (long double)Y[(size_t)i * Bp + b] - ybar;
      ss += d * d;
    }
    mu0[(size_t)b] = m0;
    vy[(size_t)b] = (double)(ss / (long double)(n - 1));
  }
  std::vector<double> Sce((size_t)Bp, 1.0), Sc((size_t)Bp * g->nk, 1.0), sigsq((size_t)Bp, 1.0), tau((size_t)Bp, 0.01),
      sigb((size_t)Bp * g->nk, 0.2), zero((size_t)Bp, 0.0);
  for (int b = 0; b < B; ++b) {
    Sce[(size_t)b] = (NU + 2.0) * (1.0 - g->R2) * vy[(size_t)b];
    sigsq[(size_t)b] = vy[(size_t)b];
    for (int j = 0; j < g->nk; ++j) Sc[(size_t)j * Bp + b] = (NU + 2.0) * g->R2 * vy[(size_t)b] / g->kernels[j].mean_diag;
  }

  // ---- state
  const size_t total = (size_t)n * Bp;
  BGGE_TRY(up(g->Y, Y, st));
  BGGE_TRY(up(g->narank, narank, st));
  BGGE_TRY(up(g->mu0, mu0, st));
  BGGE_TRY(up(g->mu, mu0, st));
  BGGE_TRY(up(g->shift, zero, st));
  BGGE_TRY(up(g->sigsq, sigsq, st));
  BGGE_TRY(up(g->tau, tau, st));
  BGGE_TRY(up(g->Sce, Sce, st));
  BGGE_TRY(up(g->Sc, Sc, st));
  BGGE_TRY(up(g->sigb, sigb, st));
  BGGE_CUDA(g->R.alloc(total));
  BGGE_CUDA(g->Ep.alloc(total + 64));
  BGGE_CUDA(g->U.alloc(total * g->nk));
  BGGE_CUDA(g->Um.alloc(total * g->nk));
  BGGE_CUDA(g->Um2.alloc(total * g->nk));
  BGGE_CUDA(cudaMemsetAsync(g->Um.p, 0, total * g->nk * sizeof(double), st));
  BGGE_CUDA(cudaMemsetAsync(g->Um2.p, 0, total * g->nk * sizeof(double), st));
  BGGE_CUDA(cudaMemsetAsync(g->Ep.p, 0, (total + 64) * sizeof(double), st));
  BGGE_CUDA(g->it.alloc(1));
  BGGE_CUDA(cudaMemsetAsync(g->it.p, 0, sizeof(long long), st));
  // row chunks of the column sums / the imputation pass: >= 4 blocks per SM and chain tile, a multiple of the 8 row lanes
  g->rows_per_chunk = (int)std::max<int64_t>(16, round_up((n + 4 * sms - 1) / (4 * sms), TY));
  g->nchunks = (int)((n + g->rows_per_chunk - 1) / g->rows_per_chunk);
  BGGE_CUDA(g->part.alloc((size_t)g->nchunks * Bp));
  BGGE_CUDA(g->part2.alloc((size_t)g->nchunks * Bp));
  const size_t nc = (size_t)std::max<int64_t>(1, g->ncum);
  BGGE_CUDA(g->chain_mu.alloc(nc * Bp));
  BGGE_CUDA(g->chain_varE.alloc(nc * Bp));
  BGGE_CUDA(g->chain_varU.alloc(nc * Bp * g->nk));

  // ---- per kernel: transposed copies, GEMM work items (split-K so that every launch fills the machine)
  int max_ksplit_back = 1;
  for (int j = 0; j < g->nk; ++j) {
    BatchKernel& bk = g->kernels[j];
    BGGE_REQUIRE(!bk.blocks.empty(), "batch: kernel %d has no eigen blocks", j);
    {   // blocks in row order, eigenvalues re-concatenated to match (col_off = insertion offset until here)
      int64_t tot = 0;
      for (auto& bl : bk.blocks) tot += bl.nr;
      BGGE_REQUIRE((int64_t)bk.s_host.size() == tot, "batch: kernel %d eigenvalue count mismatch", j);
      std::sort(bk.blocks.begin(), bk.blocks.end(), [](const BatchBlock& a, const BatchBlock& b) { return a.pos0 < b.pos0; });
      std::vector<double> ss;
      for (auto& bl : bk.blocks) ss.insert(ss.end(), bk.s_host.begin() + bl.col_off, bk.s_host.begin() + bl.col_off + bl.nr);
      bk.s_host.swap(ss);
    }
    bk.nr_total = 0;
    int64_t cursor = 0, proj_tiles = 0, back_tiles = 0, max_rows = 0, max_nr = 0;
    for (auto& bl : bk.blocks) {
      BGGE_REQUIRE(bl.pos0 >= cursor && bl.pos0 + bl.rows <= n && bl.nr > 0 && bl.ldu % 2 == 0, "batch: kernel %d block layout invalid", j);
      cursor = bl.pos0 + bl.rows;
      bl.col_off = bk.nr_total;
      bk.nr_total += bl.nr;
      proj_tiles += (bl.nr + 127) / 128;
      back_tiles += (bl.rows + 127) / 128;
      max_rows = std::max(max_rows, bl.rows);
      max_nr = std::max(max_nr, bl.nr);
      // blocked copies of t(U) and U: element (m, k) of t(U) is U[k + m ldu], of U it is U[m + k ldu]
      const int64_t mtT = (bl.nr + 127) / 128, kbT = (bl.rows + 15) / 16, mtU = (bl.rows + 127) / 128, kbU = (bl.nr + 15) / 16;
      BGGE_CUDA(pool_alloc((void**)&bl.dTb, (size_t)(mtT * kbT * GEMM_ABLK_DOUBLES) * sizeof(double)));
      BGGE_CUDA(pool_alloc((void**)&bl.dUb, (size_t)(mtU * kbU * GEMM_ABLK_DOUBLES) * sizeof(double)));
      BGGE_TRY(gemm_pack_launch(bl.dU, bl.ldu, 1, bl.nr, bl.rows, bl.dTb, st));
      BGGE_TRY(gemm_pack_launch(bl.dU, 1, bl.ldu, bl.rows, bl.nr, bl.dUb, st));
    }
    BGGE_REQUIRE((int64_t)bk.s_host.size() == bk.nr_total, "batch: kernel %d eigenvalue count mismatch", j);
    std::vector<unsigned char> cover((size_t)n, 0);
    for (auto& bl : bk.blocks) std::fill(cover.begin() + bl.pos0, cover.begin() + bl.pos0 + bl.rows, (unsigned char)1);
    BGGE_TRY(up(bk.cover, cover, st));
    const int ntile_n = (Bq + g->tn - 1) / g->tn;
    BGGE_REQUIRE(ntile_n > 1 || Bp <= g->tn + 4, "batch: internal error (B slab pitch %d exceeds the stage)", Bp);
    // K splits: a cost model in microseconds -- waves of persistent CTAs x (k-blocks per item at the FP64 tensor
    // rate + ring fill / epilogue) + the traffic of the per-split partial outputs.  (r01's model charged the partial
    // sums as if the output were as large as at 320 chains; at 40 chains per GPU it picked too few splits and left
    // a quarter of the SMs idle.)
    auto pick_split = [&](int64_t tiles, int64_t kmax, int64_t mtot) {
      const int64_t kblocks = (kmax + 15) / 16;
      const double t_kb = 32.0 * g->tn / 1.9e3;                       // 128 x tn x 16 FMAs at 64 FMA / clk / SM, ~1.9 GHz
      const double t_fix = 3.0;
      const double t_part = 2.0 * 8.0 * (double)mtot * Bp / 3.0e6;    // write + re-read of one split's partial C at ~3 TB/s
      int best = 1;
      double best_cost = 1e300;
      for (int sp = 1; sp <= 32 && (sp == 1 || kblocks / sp >= 12); ++sp) {
        const int64_t items = tiles * ntile_n * sp;
        const double waves = (double)((items + sms - 1) / sms);
        const double cost = waves * ((double)((kblocks + sp - 1) / sp) * t_kb + t_fix) + sp * t_part;
        if (cost < best_cost) { best_cost = cost; best = sp; }
      }
      return best;
    };
    int64_t rows_tot = 0;
    for (auto& bl : bk.blocks) rows_tot += bl.rows;
    bk.ksplit_proj = pick_split(proj_tiles, max_rows, bk.nr_total);
    bk.ksplit_back = pick_split(back_tiles, max_nr, rows_tot);
    if (const char* ev = getenv("BGGE_BATCH_KSPLIT")) {   // tuning hook: "proj,back"
      int a1 = 0, a2 = 0;
      if (sscanf(ev, "%d,%d", &a1, &a2) == 2 && a1 >= 1 && a1 <= 32 && a2 >= 1 && a2 <= 32) {
        bk.ksplit_proj = a1;
        bk.ksplit_back = a2;
      }
    }
    max_ksplit_back = std::max(max_ksplit_back, bk.ksplit_back);
    bk.nr_pad = round_up(bk.nr_total, 32);
    bk.n_ctiles = (int)((bk.nr_total + 31) / 32);
    BGGE_CUDA(bk.Dpart.alloc((size_t)bk.ksplit_proj * bk.nr_pad * Bp));
    BGGE_CUDA(bk.Bm.alloc((size_t)bk.nr_pad * Bp));
    BGGE_CUDA(cudaMemsetAsync(bk.Bm.p, 0, (size_t)bk.nr_pad * Bp * sizeof(double), st));
    BGGE_CUDA(bk.zqpart.alloc((size_t)bk.n_ctiles * Bp));
    BGGE_TRY(up(bk.s, bk.s_host, st));
  }
  BGGE_CUDA(g->Upart.alloc(total * max_ksplit_back));
  BGGE_CUDA(cudaMemsetAsync(g->Upart.p, 0, total * max_ksplit_back * sizeof(double), st));
  for (int j = 0; j < g->nk; ++j) {
    BatchKernel& bk = g->kernels[j];
    std::vector<GemmItem> pi, bi;
    for (auto& bl : bk.blocks) {
      const int kbp = (int)((bl.rows + 15) / 16), kbb = (int)((bl.nr + 15) / 16);
      for (int sp = 0; sp < bk.ksplit_proj; ++sp) {
        const int k0 = (int)((int64_t)kbp * sp / bk.ksplit_proj), k1 = (int)((int64_t)kbp * (sp + 1) / bk.ksplit_proj);
        for (int64_t m0 = 0; m0 < bl.nr; m0 += 128)
          for (int n0 = 0; n0 < Bq; n0 += g->tn) {
            GemmItem t{};
            t.Ablk = bl.dTb;
            t.ablk_kblocks = kbp;
            t.bcontig = Bq <= g->tn ? 1 : 0;
            t.B = g->Ep.p + (size_t)bl.pos0 * Bp;
            t.ldb = Bp;
            t.C = bk.Dpart.p + (size_t)sp * bk.nr_pad * Bp + (size_t)bl.col_off * Bp;
            t.ldc = Bp;
            t.M = (int)bl.nr;
            t.N = Bq;
            t.K = (int)bl.rows;
            t.m0 = (int)m0;
            t.n0 = n0;
            t.kb0 = k0;
            t.kb1 = k1;
            pi.push_back(t);
          }
      }
      for (int sp = 0; sp < bk.ksplit_back; ++sp) {
        const int k0 = (int)((int64_t)kbb * sp / bk.ksplit_back), k1 = (int)((int64_t)kbb * (sp + 1) / bk.ksplit_back);
        for (int64_t m0 = 0; m0 < bl.rows; m0 += 128)
          for (int n0 = 0; n0 < Bq; n0 += g->tn) {
            GemmItem t{};
            t.Ablk = bl.dUb;
            t.ablk_kblocks = kbb;
            t.bcontig = Bq <= g->tn ? 1 : 0;
            t.B = bk.Bm.p + (size_t)bl.col_off * Bp;
            t.ldb = Bp;
            t.C = g->Upart.p + (size_t)sp * total + (size_t)bl.pos0 * Bp;
            t.ldc = Bp;
            t.M = (int)bl.rows;
            t.N = Bq;
            t.K = (int)bl.nr;
            t.m0 = (int)m0;
            t.n0 = n0;
            t.kb0 = k0;
            t.kb1 = k1;
            bi.push_back(t);
          }
      }
    }
    bk.n_proj = (int)pi.size();
    bk.n_back = (int)bi.size();
    BGGE_TRY(up(bk.proj_items, pi, st));
    BGGE_TRY(up(bk.back_items, bi, st));
  }
  // a kernel with fewer K splits than Upart holds must not read stale splits: every kernel uses its own count
  const unsigned ig = (unsigned)((total + BT - 1) / BT);
  b_init_kernel<<<ig, BT, 0, st>>>(n, g->nk, B, Bp, g->Y.p, g->mu0.p, g->U.p, g->R.p, g->draw);
  BGGE_CUDA(cudaGetLastError());
  BGGE_TRY(launch_tail(g, true));
  BGGE_CUDA(cudaStreamSynchronize(st));
  g->done = 0;
  g->initialised = true;
  return OK;
}

int batch_run(Batch* g, int64_t iters) {
  if (!g->initialised) {
    set_error("batch: run before init");
    return ERR_STATE;
  }
  BGGE_REQUIRE(iters >= 0 && g->done + iters <= g->ite, "batch: asked for %lld iterations, %lld of %lld already done",
               (long long)iters, (long long)g->done, (long long)g->ite);
  if (!g->graph_exec) {
    BGGE_CUDA(cudaStreamBeginCapture(g->stream, cudaStreamCaptureModeThreadLocal));
    int s = launch_iteration(g);
    cudaError_t ce = cudaStreamEndCapture(g->stream, &g->graph);
    if (s != OK) return s;
    BGGE_CUDA(ce);
    BGGE_CUDA(cudaGraphInstantiate(&g->graph_exec, g->graph, 0));
  }
  for (int64_t i = 0; i < iters; ++i) BGGE_CUDA(cudaGraphLaunch(g->graph_exec, g->stream));
  g->done += iters;
  return OK;
}

}  // namespace bgge

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
constexpr int TXY = 256;                                       // threads of the (chains x lanes) tiled kernels, see b_draw_kernel
inline int tile_x(int Bp) { return Bp <= 48 ? 16 : 32; }

__device__ __forceinline__ DrawCfg chain_cfg(const DrawCfg& dc, int b) {
  DrawCfg c = dc;
  c.chain = dc.chain + (unsigned)b;
  return c;
}

// sum of count values p[k * stride] by one warp in a fixed order: lane-strided partial sums, then a xor tree
__device__ __forceinline__ double warp_sum_strided(const double* __restrict__ p, int count, long long stride, int lane) {
  double a = 0.0;
  for (int k = lane; k < count; k += 32) a += p[(long long)k * stride];
  return warp_sum(a);
}

// u_j = z/(2n), R = ((y - mu0) - XF Bet) - sum_j u_j          (R/BGGE.R:254, 260-267)
// (E0 = R + U_0 is the B operand of the first GEMM of an iteration -- see the note at b_draw_kernel -- and is kept up
// to date by whichever kernel changes R or the next U: here, b_update_kernel, b_na_sumr_kernel, b_xf_apply_kernel)
__global__ void b_init_kernel(long long n, int nk, int B, int Bp, const double* __restrict__ Y,
                              const double* __restrict__ mu0, double* __restrict__ U, double* __restrict__ R, DrawCfg dc,
                              int q, const double* __restrict__ xf, const double* __restrict__ Bet, double* __restrict__ Ep) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n * Bp) return;
  const long long i = idx / Bp;
  const int b = (int)(idx % Bp);
  if (b >= B) {
    R[idx] = 0.0;
    Ep[idx] = 0.0;
    for (int j = 0; j < nk; ++j) U[(long long)j * n * Bp + idx] = 0.0;
    return;
  }
  const DrawCfg c = chain_cfg(dc, b);
  const double sd = 1.0 / (2.0 * (double)n);
  double usum = 0.0, u0 = 0.0;
  for (int j = 0; j < nk; ++j) {
    const long long k = (long long)j * n + i;
    const double uj = draw_normal(c, 0, STREAM_INIT_U, (unsigned long long)k, 0) * sd;
    U[(long long)j * n * Bp + idx] = uj;
    usum = (j == 0) ? uj : usum + uj;
    if (j == 0) u0 = uj;
  }
  double rv = Y[idx] - mu0[b];
  if (q > 0) {                                                      // R/BGGE.R:264
    double xb = 0.0;
    for (int k = 0; k < q; ++k) xb += xf[i + (long long)k * n] * Bet[(long long)k * Bp + b];
    rv -= xb;
  }
  R[idx] = rv - usum;
  Ep[idx] = (rv - usum) + u0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Fixed effects shared by all chains (R/BGGE.R:284-290, 106-109): three small kernels at the head of an iteration.
// XF is n x q column-major (every chain of a tile reads the same x: broadcast loads), Bet / BetOld are [q][Bp].
// ---------------------------------------------------------------------------------------------------------------------
// partial sums over a row chunk of XF'(temp + XF Bet), temp = R - shift     -> part[chunk][k][Bp]
__global__ void __launch_bounds__(TXY)
b_xf_reduce_kernel(long long n, int q, int Bp, int rows_per_chunk, const double* __restrict__ xf,
                   const double* __restrict__ R, const double* __restrict__ Bet, const double* __restrict__ shift,
                   double* __restrict__ part, long long* __restrict__ advance_it) {
  pdl_wait();   // programmatic dependent launch: resident early, runs once the predecessor has completed
  pdl_trigger();
  __shared__ double zs[TXY];
  const int tx = blockDim.x, ty = blockDim.y;
  const int b = blockIdx.y * tx + threadIdx.x, ry = threadIdx.y;
  if (advance_it != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && ry == 0) *advance_it += 1;
  const long long i0 = (long long)blockIdx.x * rows_per_chunk;
  const long long i1 = min(n, i0 + rows_per_chunk);
  double acc[BATCH_MAXQ], bet[BATCH_MAXQ];
#pragma unroll
  for (int k = 0; k < BATCH_MAXQ; ++k) {
    acc[k] = 0.0;
    bet[k] = (k < q && b < Bp) ? Bet[(long long)k * Bp + b] : 0.0;
  }
  if (b < Bp) {
    const double sh = shift[b];
    for (long long i = i0 + ry; i < i1; i += ty) {
      double x[BATCH_MAXQ];
      double xb = 0.0;
#pragma unroll
      for (int k = 0; k < BATCH_MAXQ; ++k) {
        x[k] = k < q ? xf[i + (long long)k * n] : 0.0;
        if (k < q) xb += x[k] * bet[k];
      }
      const double t = (R[i * Bp + b] - sh) + xb;                   // temp + XF Bet   (R/BGGE.R:285)
#pragma unroll
      for (int k = 0; k < BATCH_MAXQ; ++k)
        if (k < q) acc[k] = fma(x[k], t, acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < BATCH_MAXQ; ++k) {
    if (k >= q) break;
    zs[ry * tx + threadIdx.x] = acc[k];
    __syncthreads();
    if (ry == 0 && b < Bp) {
      double t = 0.0;
      for (int l = 0; l < ty; ++l) t += zs[l * tx + threadIdx.x];   // fixed order
      part[((long long)blockIdx.x * q + k) * Bp + b] = t;
    }
    __syncthreads();
  }
}

// one warp per chain: media = tXX XF'temp; Bet = media + chol(sigsq tXX)' z; lane k owns coefficient k
__global__ void __launch_bounds__(128)
b_xf_draw_kernel(int q, int B, int Bp, int nchunks, const double* __restrict__ part, const double* __restrict__ txx,
                 const double* __restrict__ lchol, const double* __restrict__ sigsq, double* __restrict__ Bet,
                 double* __restrict__ BetOld, const long long* __restrict__ itp, DrawCfg dc) {
  pdl_wait();   // programmatic dependent launch: resident early, runs once the predecessor has completed
  pdl_trigger();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const long long it = *itp;
  const DrawCfg c = chain_cfg(dc, b);
  double v[BATCH_MAXQ];
#pragma unroll
  for (int k = 0; k < BATCH_MAXQ; ++k)
    v[k] = k < q ? warp_sum_strided(part + (long long)k * Bp + b, nchunks, (long long)q * Bp, lane) : 0.0;
  const double z = lane < q ? draw_normal(c, it, STREAM_XF, (unsigned long long)lane, 0) : 0.0;
  double media = 0.0, lz = 0.0;
#pragma unroll
  for (int k2 = 0; k2 < BATCH_MAXQ; ++k2) {
    const double zk = __shfl_sync(0xffffffffu, z, k2);
    if (k2 < q && lane < q) {
      media += txx[lane + (long long)k2 * q] * v[k2];
      if (k2 <= lane) lz += lchol[lane + (long long)k2 * q] * zk;
    }
  }
  if (lane < q) {
    const long long o = (long long)lane * Bp + b;
    BetOld[o] = Bet[o];
    Bet[o] = media + sqrt(sigsq[b]) * lz;
  }
}

// R = (R + XF BetOld) - XF Bet                                                  (R/BGGE.R:285, 289)
__global__ void b_xf_apply_kernel(long long n, int q, int B, int Bp, const double* __restrict__ xf,
                                  const double* __restrict__ Bet, const double* __restrict__ BetOld, double* __restrict__ R,
                                  const double* __restrict__ U0, double* __restrict__ Ep) {
  pdl_wait();   // programmatic dependent launch: resident early, runs once the predecessor has completed
  pdl_trigger();
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n * Bp) return;
  const long long i = idx / Bp;
  const int b = (int)(idx % Bp);
  if (b >= B) return;
  double xo = 0.0, xn = 0.0;
  for (int k = 0; k < q; ++k) {
    const double x = xf[i + (long long)k * n];
    xo += x * BetOld[(long long)k * Bp + b];
    xn += x * Bet[(long long)k * Bp + b];
  }
  const double rn = (R[idx] + xo) - xn;
  R[idx] = rn;
  Ep[idx] = rn + U0[idx];
}

// The projection d = U'((R - shift_b) + U_j) is computed as U'E0 - shift_b U'1 with E0 = R + U_j: E0 does not depend on
// the intercept, so the pass that last changes R (the previous kernel's update, the imputation, the fixed-effects
// apply) writes it on the side and no separate E' pass over the n x B matrices is needed; U'1 (column sums of U over the
// block's rows) is computed once at init.  The first GEMM of an iteration advances the iteration counter.
// d = sum of K-split partials; b = tau v d + sqrt(v) z; partial sums of b^2/s        (R/BGGE.R:302-305, 96)
// Small kernels of the batched sampler are tiled as (32 chains) x (8 row / column lanes): with few chains per GPU
// (40 at 8 GPUs) a thread-per-chain layout leaves the chip idle behind long serial loops (r02: b_draw 79 us at B = 40).
// Block = tx chains x ty lanes with tx * ty = 256; tx = 16 for <= 48 chain columns (44 columns = 16 + 16 + 12 lanes
// busy instead of 32 + 12), else 32.

// Blocks [0, n_ctiles): one coefficient draw per thread (ty columns x tx chains per block), partial sums of b^2/s per
// block.  Blocks >= n_ctiles prepare, off the critical path, the draws later kernels of the iteration would otherwise
// wait for: block n_ctiles the chi-squares of kernel j (and, for the last kernel, the residual chi-square and the
// next intercept's normal) per chain, the blocks behind it the imputation normals (last kernel only).  The draws are
// counter based, so they are the same numbers the consumers used to draw themselves.
struct PreDraws {
  double* chiK;     // [nk][Bp]
  double* chiE;     // [Bp]
  double* zmu;      // [Bp]
  double* zna;      // [max_na][Bp]
  int max_na, last; // last: this launch belongs to the last kernel of the iteration
  const int* n_na;  // [Bp] missing phenotypes per chain
  long long n;
  long long nr_j;
};
__global__ void __launch_bounds__(TXY)
b_draw_kernel(int nr, int n_ctiles, int B, int Bp, int nsplit, const double* __restrict__ Dpart, long long dstride,
              const double* __restrict__ s, const double* __restrict__ sigb, const double* __restrict__ tau,
              double* __restrict__ Bm, double* __restrict__ zqpart, const long long* __restrict__ itp, int j, DrawCfg dc,
              PreDraws pd, const double* __restrict__ shift, const double* __restrict__ ut1) {
  pdl_wait();   // programmatic dependent launch: resident early, runs once the predecessor has completed
  pdl_trigger();
  __shared__ double zs[TXY];
  const int tx = blockDim.x, ty = blockDim.y;
  const int b = blockIdx.y * tx + threadIdx.x, cy = threadIdx.y;
  const long long it = *itp;
  if ((int)blockIdx.x >= n_ctiles) {
    const int extra = (int)blockIdx.x - n_ctiles;
    if (b >= B) return;
    const DrawCfg c = chain_cfg(dc, b);
    if (extra == 0) {
      if (cy == 0) pd.chiK[(long long)j * Bp + b] = draw_chisq(c, it, STREAM_CHI_K + (uint32_t)j, (double)pd.nr_j + NU, 0);
      if (pd.last && cy == 1) pd.chiE[b] = draw_chisq(c, it, STREAM_CHI_E, (double)pd.n + NU, 0);
      if (pd.last && cy == 2) pd.zmu[b] = draw_normal(c, it + 1, STREAM_MU, 0ull, 0);
    } else {
      const int k = (extra - 1) * ty + cy;
      if (k < pd.n_na[b]) pd.zna[(long long)k * Bp + b] = draw_normal(c, it, STREAM_NA, (unsigned long long)k, 0);
    }
    return;
  }
  const int cc = blockIdx.x * ty + cy;
  double zq = 0.0;
  if (cc < nr) {
    if (b < B) {
      const DrawCfg c = chain_cfg(dc, b);
      const double lam = sigb[b], tb = tau[b];
      double d = 0.0;
      for (int sp = 0; sp < nsplit; ++sp) d += Dpart[(long long)sp * dstride + (long long)cc * Bp + b];
      d -= shift[b] * ut1[cc];              // U'(E0 - shift 1)
      const double sv = s[cc];
      const double vari = sv * lam / (1.0 + sv * lam * tb);
      const double media = tb * vari * d;
      const double z = draw_normal(c, it, (uint32_t)j, (unsigned long long)cc, 0);
      const double bv = media + sqrt(vari) * z;
      Bm[(long long)cc * Bp + b] = bv;
      zq = bv * bv * (1.0 / sv);
    } else if (b < Bp) {
      Bm[(long long)cc * Bp + b] = 0.0;
    }
  }
  zs[cy * tx + threadIdx.x] = zq;
  __syncthreads();
  if (cy == 0 && b < Bp) {
    double t = 0.0;
    for (int k = 0; k < ty; ++k) t += zs[k * tx + threadIdx.x];   // fixed order
    zqpart[(long long)blockIdx.x * Bp + b] = t;
  }
}

// u_new = sum of the partial planes; R = (R + u_old) - u_new; Welford of kept draws     (R/BGGE.R:306-307, 392)
// After the last kernel of the iteration R is final for the residual variance: the same pass then leaves the per-chunk
// sums of (R - shift)^2 in `part` (want_sse), which saves the separate column-sum launch.
__global__ void __launch_bounds__(TXY)
b_update_kernel(long long n, int Bp, int rows_per_chunk, int nsplit, const double* __restrict__ Upart, long long ustride,
                const unsigned char* __restrict__ cover, double* __restrict__ R, double* __restrict__ Uj,
                double* __restrict__ Um, double* __restrict__ Um2, const long long* __restrict__ itp, long long thin,
                long long burn, const double* __restrict__ shift, double* __restrict__ part, int want_sse,
                const double* Unext, int next_is_self, double* __restrict__ Ep) {
  pdl_wait();   // programmatic dependent launch: resident early, runs once the predecessor has completed
  pdl_trigger();
  __shared__ double zs[TXY];
  const int tx = blockDim.x, ty = blockDim.y;
  const int b = blockIdx.y * tx + threadIdx.x, ry = threadIdx.y;
  const long long i0 = (long long)blockIdx.x * rows_per_chunk;
  const long long i1 = min(n, i0 + rows_per_chunk);
  const long long it = *itp;
  const bool keep = (it % thin == 0) && (it > burn);
  const double kcount = keep ? (double)(it / thin - burn / thin) : 1.0;
  double a = 0.0;
  if (b < Bp) {
    const double sh = want_sse ? shift[b] : 0.0;
    for (long long i = i0 + ry; i < i1; i += ty) {
      const long long idx = i * Bp + b;
      double un = 0.0;   // rows no block of this kernel covers: u = 0 (R/BGGE.R:334)
      if (cover[i])
        for (int sp = 0; sp < nsplit; ++sp) un += Upart[(long long)sp * ustride + idx];
      const double rn = (R[idx] + Uj[idx]) - un;
      R[idx] = rn;
      Uj[idx] = un;
      Ep[idx] = rn + (next_is_self ? un : Unext[idx]);   // E0 of the next kernel step (single-kernel models: this kernel)
      if (keep) {
        const double m = Um[idx];
        const double dlt = un - m;
        const double mn = m + dlt / kcount;
        Um[idx] = mn;
        Um2[idx] += dlt * (un - mn);
      }
      const double v = rn - sh;
      a = fma(v, v, a);
    }
  }
  if (want_sse) {
    zs[ry * tx + threadIdx.x] = a;
    __syncthreads();
    if (ry == 0 && b < Bp) {
      double t = 0.0;
      for (int k = 0; k < ty; ++k) t += zs[k * tx + threadIdx.x];
      part[(long long)blockIdx.x * Bp + b] = t;
    }
  }
}

// ut1[c] = sum over the block's rows of U[row, c]  (one block per eigen-column, fixed order)
__global__ void __launch_bounds__(128) b_colsum_kernel(const double* __restrict__ U, long long ldu, long long rows,
                                                        double* __restrict__ out) {
  __shared__ double red[4];
  const double* col = U + (long long)blockIdx.x * ldu;
  double a = 0.0;
  for (long long i = threadIdx.x; i < rows; i += 128) a += col[i];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = ((red[0] + red[1]) + red[2]) + red[3];
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
  int q;
  const double* Bet;    // [q][Bp]
  double* chain_bet;    // [ncum][q][Bp]
  const double* chiK;   // prepared draws (b_draw_kernel's extra blocks); zmu == nullptr: draw the intercept normal here
  const double* chiE;
  const double* zmu;
};

// per chain (one warp each): sigb_j, sigsq, tau        (R/BGGE.R:95-103, 360-367)
// the sums are formed by the whole warp; then lane j draws the chi-square of kernel j and lane nk the residual one, so
// the Marsaglia-Tsang rejection loops run side by side instead of one after the other
// sum of count values p[k * stride] by a block of 128 threads in a fixed order (thread-strided partial sums, xor tree
// per warp, the four warp sums added in warp order); every thread gets the total.  One block per chain: with one WARP
// per chain each lane walked ~10 dependent-latency loads of the 300+ per-chunk partials (r02: 7.9 us for 40 chains).
__device__ __forceinline__ double block_sum_strided(const double* __restrict__ p, int count, long long stride, double* red) {
  double a = 0.0;
  for (int k = threadIdx.x; k < count; k += 128) a += p[(long long)k * stride];
  a = warp_sum(a);
  __syncthreads();                       // red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  return ((red[0] + red[1]) + red[2]) + red[3];
}

__global__ void __launch_bounds__(128) b_chain_var_kernel(const ChainArgs a, DrawCfg dc) {
  pdl_wait();   // programmatic dependent launch: resident early, runs once the predecessor has completed
  pdl_trigger();
  __shared__ double red[4];
  __shared__ double ss[MAXK + 1];
  const int b = blockIdx.x;              // one block per chain
  (void)dc;
  for (int j = 0; j < a.nk; ++j) {
    const double zq = block_sum_strided(a.zqpart[j] + b, a.n_zq[j], a.Bp, red);   // sum b^2/s
    if (threadIdx.x == 0) ss[j] = zq;
  }
  const double sse = block_sum_strided(a.part + b, a.nchunks, a.Bp, red);         // e'e
  if (threadIdx.x == 0) ss[a.nk] = sse;
  __syncthreads();
  for (int j = threadIdx.x; j <= a.nk; j += 128) {
    if (j < a.nk) {
      const double chi = a.chiK[(long long)j * a.Bp + b];    // drawn by the extra block of kernel j's b_draw launch
      a.sigb[(long long)j * a.Bp + b] = (ss[j] + NU * a.Sc[(long long)j * a.Bp + b]) / chi;
    } else {
      const double chi = a.chiE[b];
      const double s2 = (ss[j] + a.Sce[b]) / chi;
      a.sigsq[b] = s2;
      a.tau[b] = 1.0 / s2;
    }
  }
}

// NA imputation (R/BGGE.R:370-381) fused with the column sums of r for the next intercept
__global__ void __launch_bounds__(TXY)
b_na_sumr_kernel(long long n, int nk, int B, int Bp, int rows_per_chunk, double* __restrict__ R,
                 const double* __restrict__ U, double* __restrict__ Y, const int* __restrict__ narank,
                 const double* __restrict__ mu, const double* __restrict__ shift, const double* __restrict__ sigsq,
                 const long long* __restrict__ itp, int do_na, double* __restrict__ part, const double* __restrict__ zna,
                 int q, const double* __restrict__ xf, const double* __restrict__ Bet, double* __restrict__ Ep) {
  pdl_wait();   // programmatic dependent launch: resident early, runs once the predecessor has completed
  pdl_trigger();
  __shared__ double zs[TXY];
  const int tx = blockDim.x, ty = blockDim.y;
  const int b = blockIdx.y * tx + threadIdx.x, ry = threadIdx.y;
  const long long i0 = (long long)blockIdx.x * rows_per_chunk;
  const long long i1 = min(n, i0 + rows_per_chunk);
  (void)itp;
  double a = 0.0;
  if (b < B) {
    const double mub = mu[b], shb = shift[b], sd = sqrt(sigsq[b]);
    for (long long i = i0 + ry; i < i1; i += ty) {
      const long long idx = i * Bp + b;
      double rv = R[idx];
      if (do_na) {
        const int k = narank[idx];
        if (k >= 0) {
          double uhat = 0.0;
          for (int j = 0; j < nk; ++j) uhat = (j == 0) ? U[idx] : uhat + U[(long long)j * n * Bp + idx];
          const double z = zna[(long long)k * Bp + b];   // drawn by the extra blocks of the last b_draw launch
          double aux = 0.0;                              // XF[yNA, ] Bet (R/BGGE.R:375)
          for (int c = 0; c < q; ++c) aux += xf[i + (long long)c * n] * Bet[(long long)c * Bp + b];
          const double yv = ((aux + mub) + uhat) + sd * z;
          Y[idx] = yv;
          rv = (((yv - uhat) - aux) - mub) + shb;
          R[idx] = rv;
          Ep[idx] = rv + U[idx];                  // E0 of the next iteration's first kernel step
        }
      }
      a += rv;
    }
  }
  zs[ry * tx + threadIdx.x] = a;
  __syncthreads();
  if (ry == 0 && b < Bp) {
    double t = 0.0;
    for (int k = 0; k < ty; ++k) t += zs[k * tx + threadIdx.x];
    part[(long long)blockIdx.x * Bp + b] = t;
  }
}

// per chain: record the thinned chain (R/BGGE.R:384-395), draw the next intercept (R/BGGE.R:278-281)
__global__ void __launch_bounds__(128) b_chain_mu_kernel(const ChainArgs a, DrawCfg dc) {
  pdl_wait();   // programmatic dependent launch: resident early, runs once the predecessor has completed
  pdl_trigger();
  __shared__ double red[4];
  const int b = blockIdx.x;              // one block per chain
  const long long it = *a.it;
  const double sr = block_sum_strided(a.part + b, a.nchunks, a.Bp, red);
  if (threadIdx.x != 0) return;
  const DrawCfg c = chain_cfg(dc, b);
  if (it > 0 && it % a.thin == 0) {
    const long long slot = it / a.thin - 1;
    if (slot < a.ncum) {
      a.chain_mu[slot * a.Bp + b] = a.mu[b];
      a.chain_varE[slot * a.Bp + b] = a.sigsq[b];
      for (int j = 0; j < a.nk; ++j) a.chain_varU[((long long)j * a.ncum + slot) * a.Bp + b] = a.sigb[(long long)j * a.Bp + b];
      for (int k = 0; k < a.q; ++k) a.chain_bet[(slot * a.q + k) * a.Bp + b] = a.Bet[(long long)k * a.Bp + b];
    }
  }
  const double z = a.zmu ? a.zmu[b] : draw_normal(c, it + 1, STREAM_MU, 0ull, 0);
  const double nmu = (sr / (double)a.n + a.mu0[b]) + sqrt(a.sigsq[b] / (double)a.n) * z;
  a.mu[b] = nmu;
  a.shift[b] = nmu - a.mu0[b];
}

// Launch with the programmatic-serialization attribute (captured into the iteration graph as a programmatic edge):
// the kernels of an iteration are short and strictly dependent, so the launch latency between them is a visible
// share of the iteration at few chains per GPU.  Every kernel starts with pdl_wait(); pdl_trigger().
bool batch_pdl() {
  static const bool on = pdl_enabled() && !(getenv("BGGE_BATCH_PDL") && atoi(getenv("BGGE_BATCH_PDL")) == 0);
  return on;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = batch_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
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
  a.q = g->q;
  a.Bet = g->Bet.p;
  a.chain_bet = g->chain_bet.p;
  a.chiK = g->chiK.p;
  a.chiE = g->chiE.p;
  a.zmu = g->zmu.p;
  return a;
}

int launch_tail(Batch* g, bool init) {
  cudaStream_t st = g->stream;
  const int tx = tile_x(g->Bp), ty = TXY / tx;
  const dim3 cgrid((unsigned)g->nchunks, (unsigned)((g->Bp + tx - 1) / tx));
  const dim3 cblock(tx, ty);
  const unsigned cb = (unsigned)g->B;               // one block per chain
  if (!init)   // the per-chunk sums of (R - shift)^2 were left in `part` by the last kernel's update pass
    BGGE_CUDA(launch_dep(b_chain_var_kernel, dim3(cb), dim3(128), st, chain_args(g, g->part.p), g->draw));
  BGGE_CUDA(launch_dep(b_na_sumr_kernel, cgrid, cblock, st, g->n, g->nk, g->B, g->Bp, g->rows_per_chunk, g->R.p, g->U.p, g->Y.p,
                       g->narank.p, g->mu.p, g->shift.p, g->sigsq.p, g->it.p, (!init && g->any_na) ? 1 : 0, g->part2.p, g->zna.p,
                       g->q, g->xf.p, g->Bet.p, g->Ep.p));
  ChainArgs ca = chain_args(g, g->part2.p);
  if (init) ca.zmu = nullptr;   // no b_draw launch has run yet: the first intercept's normal is drawn in place
  BGGE_CUDA(launch_dep(b_chain_mu_kernel, dim3(cb), dim3(128), st, ca, g->draw));
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int launch_iteration(Batch* g) {
  cudaStream_t st = g->stream;
  const long long total = g->n * g->Bp;
  const unsigned eg = (unsigned)((total + BT - 1) / BT);
  const int tx = tile_x(g->Bp), ty = TXY / tx;
  if (g->q > 0) {   // fixed effects, between the intercept (drawn by the previous tail) and the kernels; advances `it`
    BGGE_CUDA(launch_dep(b_xf_reduce_kernel, dim3((unsigned)g->nchunks, (unsigned)((g->Bp + tx - 1) / tx)), dim3(tx, ty), st, g->n,
                         g->q, g->Bp, g->rows_per_chunk, g->xf.p, g->R.p, g->Bet.p, g->shift.p, g->xf_part.p, g->it.p));
    BGGE_CUDA(launch_dep(b_xf_draw_kernel, dim3((unsigned)((g->B + 3) / 4)), dim3(128), st, g->q, g->B, g->Bp, g->nchunks,
                         g->xf_part.p, g->txx.p, g->lchol.p, g->sigsq.p, g->Bet.p, g->BetOld.p, g->it.p, g->draw));
    BGGE_CUDA(launch_dep(b_xf_apply_kernel, dim3(eg), dim3(BT), st, g->n, g->q, g->B, g->Bp, g->xf.p, g->Bet.p, g->BetOld.p, g->R.p,
                         g->U.p, g->Ep.p));
  }
  for (int j = 0; j < g->nk; ++j) {
    BatchKernel& bk = g->kernels[j];
    double* Uj = g->U.p + (size_t)j * total;
    // (E0 = R + U_j is already in place; the first launch of the iteration advances the iteration counter)
    BGGE_TRY(dgemm_launch(bk.proj_items.p, bk.n_proj, g->tn, st, bk.grid_proj, batch_pdl(), (j == 0 && g->q == 0) ? g->it.p : nullptr));
    const bool last = j == g->nk - 1;
    PreDraws pd{g->chiK.p, g->chiE.p, g->zmu.p, g->zna.p, g->max_na, last ? 1 : 0, g->n_na.p, (long long)g->n,
                (long long)bk.nr_total};
    const int extra = 1 + (last && g->any_na ? (g->max_na + ty - 1) / ty : 0);   // prepared draws, see b_draw_kernel
    BGGE_CUDA(launch_dep(b_draw_kernel, dim3((unsigned)(bk.n_ctiles + extra), (unsigned)((g->Bp + tx - 1) / tx)), dim3(tx, ty), st,
                         (int)bk.nr_total, bk.n_ctiles, g->B, g->Bp, bk.ksplit_proj, bk.Dpart.p, bk.nr_pad * g->Bp, bk.s.p,
                         g->sigb.p + (size_t)j * g->Bp, g->tau.p, bk.Bm.p, bk.zqpart.p, g->it.p, j, g->draw, pd, g->shift.p,
                         bk.ut1.p));
    BGGE_TRY(dgemm_launch(bk.back_items.p, bk.n_back, g->tn, st, bk.grid_back, batch_pdl()));
    BGGE_CUDA(launch_dep(b_update_kernel, dim3((unsigned)g->nchunks, (unsigned)((g->Bp + tx - 1) / tx)), dim3(tx, ty), st, g->n,
                         g->Bp, g->rows_per_chunk, bk.ksplit_back, bk.Upart.p, total, bk.cover.p, g->R.p, Uj,
                         g->Um.p + (size_t)j * total, g->Um2.p + (size_t)j * total, g->it.p, g->thin, g->burn, g->shift.p,
                         g->part.p, j == g->nk - 1 ? 1 : 0, g->U.p + (size_t)((j + 1) % g->nk) * total, g->nk == 1 ? 1 : 0,
                         g->Ep.p));
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
  std::vector<int> narank((size_t)n * Bp, -1), nna_host((size_t)Bp, 0);
  g->max_na = 0;
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
    nna_host[(size_t)b] = rank;
    g->max_na = std::max(g->max_na, rank);
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

  // ---- fixed effects (R/BGGE.R:208-212): one XF for all chains, Bet_b = solve(crossprod(XF), crossprod(XF, y_b))
  const int q = g->q;
  BGGE_REQUIRE(q >= 0 && q <= BATCH_MAXQ, "batch: at most %d fixed effects in the batched sampler (q = %d)", BATCH_MAXQ, q);
  std::vector<double> bet0((size_t)std::max(1, q) * Bp, 0.0), txx, lch;
  if (q > 0) {
    BGGE_REQUIRE((int64_t)g->xf_host.size() == n * q, "batch: XF not set");
    std::vector<double> xtx((size_t)q * q, 0.0);
    for (int a = 0; a < q; ++a)
      for (int c = a; c < q; ++c) {
        long double v = 0.0L;
        for (int64_t i = 0; i < n; ++i) v += (long double)g->xf_host[(size_t)(i + a * n)] * g->xf_host[(size_t)(i + c * n)];
        xtx[(size_t)a + (size_t)c * q] = xtx[(size_t)c + (size_t)a * q] = (double)v;
      }
    if (!solve_spd_inverse(q, xtx, txx)) {
      set_error("batch: crossprod(XF) is singular");
      return ERR_INVALID;
    }
    if (!chol_lower(q, txx, lch)) {
      set_error("batch: solve(crossprod(XF)) is not positive definite");
      return ERR_INVALID;
    }
    std::vector<double> xty((size_t)q);
    for (int b = 0; b < B; ++b) {
      for (int a = 0; a < q; ++a) {
        long double v = 0.0L;
        for (int64_t i = 0; i < n; ++i) v += (long double)g->xf_host[(size_t)(i + a * n)] * Y[(size_t)i * Bp + b];
        xty[(size_t)a] = (double)v;
      }
      for (int a = 0; a < q; ++a) {
        double v = 0.0;
        for (int c = 0; c < q; ++c) v += txx[(size_t)a + (size_t)c * q] * xty[(size_t)c];
        bet0[(size_t)a * Bp + b] = v;
      }
    }
  }

  // ---- state
  const size_t total = (size_t)n * Bp;
  BGGE_TRY(up(g->Bet, bet0, st));
  BGGE_TRY(up(g->BetOld, bet0, st));
  if (q > 0) {
    BGGE_TRY(up(g->xf, g->xf_host, st));
    BGGE_TRY(up(g->txx, txx, st));
    BGGE_TRY(up(g->lchol, lch, st));
  }
  BGGE_TRY(up(g->Y, Y, st));
  BGGE_TRY(up(g->narank, narank, st));
  BGGE_TRY(up(g->n_na, nna_host, st));
  BGGE_CUDA(g->chiK.alloc((size_t)g->nk * Bp));
  BGGE_CUDA(g->chiE.alloc((size_t)Bp));
  BGGE_CUDA(g->zmu.alloc((size_t)Bp));
  BGGE_CUDA(g->zna.alloc((size_t)std::max(1, g->max_na) * Bp));
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
  g->rows_per_chunk = (int)std::max<int64_t>(16, round_up((n + 4 * sms - 1) / (4 * sms), TXY / tile_x(Bp)));
  g->nchunks = (int)((n + g->rows_per_chunk - 1) / g->rows_per_chunk);
  BGGE_CUDA(g->part.alloc((size_t)g->nchunks * Bp));
  BGGE_CUDA(g->part2.alloc((size_t)g->nchunks * Bp));
  const size_t nc = (size_t)std::max<int64_t>(1, g->ncum);
  BGGE_CUDA(g->chain_mu.alloc(nc * Bp));
  BGGE_CUDA(g->chain_varE.alloc(nc * Bp));
  BGGE_CUDA(g->chain_varU.alloc(nc * Bp * g->nk));
  BGGE_CUDA(g->chain_bet.alloc(nc * Bp * (size_t)std::max(1, q)));
  BGGE_CUDA(g->xf_part.alloc((size_t)g->nchunks * (size_t)std::max(1, q) * Bp));

  // ---- per kernel: transposed copies, GEMM work items (split-K so that every launch fills the machine)
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
      const int64_t gm = gemm_tile_m(g->tn), ablk = gemm_ablk_doubles(g->tn);
      const int64_t mtT = (bl.nr + gm - 1) / gm, kbT = (bl.rows + 15) / 16, mtU = (bl.rows + gm - 1) / gm, kbU = (bl.nr + 15) / 16;
      BGGE_CUDA(pool_alloc((void**)&bl.dTb, (size_t)(mtT * kbT * ablk) * sizeof(double)));
      BGGE_CUDA(pool_alloc((void**)&bl.dUb, (size_t)(mtU * kbU * ablk) * sizeof(double)));
      BGGE_TRY(gemm_pack_launch(bl.dU, bl.ldu, 1, bl.nr, bl.rows, g->tn, bl.dTb, st));
      BGGE_TRY(gemm_pack_launch(bl.dU, 1, bl.ldu, bl.rows, bl.nr, g->tn, bl.dUb, st));
    }
    BGGE_REQUIRE((int64_t)bk.s_host.size() == bk.nr_total, "batch: kernel %d eigenvalue count mismatch", j);
    std::vector<unsigned char> cover((size_t)n, 0);
    for (auto& bl : bk.blocks) std::fill(cover.begin() + bl.pos0, cover.begin() + bl.pos0 + bl.rows, (unsigned char)1);
    BGGE_TRY(up(bk.cover, cover, st));
    const int ntile_n = (Bq + g->tn - 1) / g->tn;
    BGGE_REQUIRE(ntile_n > 1 || Bp <= g->tn + 4, "batch: internal error (B slab pitch %d exceeds the stage)", Bp);
    bk.nr_pad = round_up(bk.nr_total, 32);
    bk.n_ctiles = (int)((bk.nr_total + (TXY / tile_x(Bp)) - 1) / (TXY / tile_x(Bp)));   // one block per ty eigen-columns
    BGGE_CUDA(bk.Bm.alloc((size_t)bk.nr_pad * Bp));
    BGGE_CUDA(cudaMemsetAsync(bk.Bm.p, 0, (size_t)bk.nr_pad * Bp * sizeof(double), st));
    BGGE_CUDA(bk.zqpart.alloc((size_t)bk.n_ctiles * Bp));
    BGGE_TRY(up(bk.s, bk.s_host, st));
    BGGE_CUDA(bk.ut1.alloc((size_t)bk.nr_pad));
    BGGE_CUDA(cudaMemsetAsync(bk.ut1.p, 0, (size_t)bk.nr_pad * sizeof(double), st));
    for (auto& bl : bk.blocks)
      b_colsum_kernel<<<(unsigned)bl.nr, 128, 0, st>>>(bl.dU, bl.ldu, bl.rows, bk.ut1.p + bl.col_off);
    BGGE_CUDA(cudaGetLastError());
  }
  // ---- GEMM work items, stream-K style: the (tile, k-block) space of a GEMM is cut into one contiguous range per
  // persistent CTA (equal k-block counts), a range that crosses a tile boundary becomes two items.  Every item writes
  // its partial C into the plane = its position among the pieces of its tile; the consumer kernels add the planes in
  // order (planes a tile does not use stay zero from initialisation).  Deterministic, no atomics, no wave
  // quantisation: the 79 row tiles of u = U b at C4 no longer run as 2 waves on 148 SMs.
  struct TileSpec {
    GemmItem base;      // everything but kb0 / kb1 / C
    int kblocks;
    long long coff;     // offset of the tile's block inside a plane of the output
  };
  // (called twice per GEMM: once without an output to learn the plane count, then with the allocated planes)
  auto streamk = [&](const std::vector<TileSpec>& tiles, double* cbase, long long plane_stride, int* nplanes, int* grid) {
    std::vector<GemmItem> out;
    long long total_kb = 0;
    for (auto& t : tiles) total_kb += t.kblocks;
    const int min_piece = 6;
    int P = (int)std::max<long long>(1, std::min<long long>(sms, total_kb / (2 * min_piece)));
    if (const char* ev = getenv("BGGE_BATCH_CTAS")) P = std::max(1, std::min(sms, atoi(ev)));   // tuning hook
    std::vector<std::vector<GemmItem>> per_cta((size_t)P);
    std::vector<int> used(tiles.size(), 0);
    size_t ti = 0;
    long long tile_start = 0, pos = 0;     // linear k-block position of the current tile / of the cursor
    for (int c = 0; c < P; ++c) {
      long long end = c == P - 1 ? total_kb : total_kb * (c + 1) / P;
      // snap the cut to a tile edge when it would leave a sliver
      {
        size_t tj = ti;
        long long ts = tile_start;
        while (tj < tiles.size() && ts + tiles[tj].kblocks <= end) ts += tiles[tj++].kblocks;
        if (tj < tiles.size()) {
          if (end - ts > 0 && end - ts < min_piece) end = ts;
          else if (ts + tiles[tj].kblocks - end < min_piece) end = ts + tiles[tj].kblocks;
        }
      }
      while (pos < end && ti < tiles.size()) {
        const long long tile_end = tile_start + tiles[ti].kblocks;
        const long long piece_end = std::min(end, tile_end);
        GemmItem it = tiles[ti].base;
        it.kb0 = (int)(pos - tile_start);
        it.kb1 = (int)(piece_end - tile_start);
        it.C = cbase ? cbase + (long long)used[ti] * plane_stride + tiles[ti].coff : nullptr;
        ++used[ti];
        per_cta[(size_t)c].push_back(it);
        pos = piece_end;
        if (pos == tile_end) {
          tile_start = tile_end;
          ++ti;
        }
      }
    }
    size_t rounds = 0;
    for (auto& v : per_cta) rounds = std::max(rounds, v.size());
    GemmItem idle{};   // M = 0, no k-blocks: nothing loaded, nothing stored
    for (size_t r = 0; r < rounds; ++r)
      for (int c = 0; c < P; ++c) out.push_back(r < per_cta[(size_t)c].size() ? per_cta[(size_t)c][r] : idle);
    *nplanes = 1;
    for (int u : used) *nplanes = std::max(*nplanes, u);
    *grid = P;
    return out;
  };
  for (int j = 0; j < g->nk; ++j) {
    BatchKernel& bk = g->kernels[j];
    std::vector<TileSpec> pt, bt;
    for (auto& bl : bk.blocks) {
      const int kbp = (int)((bl.rows + 15) / 16), kbb = (int)((bl.nr + 15) / 16);
      const int64_t gm = gemm_tile_m(g->tn);
      for (int64_t m0 = 0; m0 < bl.nr; m0 += gm)
        for (int n0 = 0; n0 < Bq; n0 += g->tn) {
          GemmItem t{};
          t.Ablk = bl.dTb;
          t.ablk_kblocks = kbp;
          t.bcontig = Bq <= g->tn ? 1 : 0;
          t.B = g->Ep.p + (size_t)bl.pos0 * Bp;
          t.ldb = Bp;
          t.ldc = Bp;
          t.M = (int)bl.nr;
          t.N = Bq;
          t.K = (int)bl.rows;
          t.m0 = (int)m0;
          t.n0 = n0;
          pt.push_back(TileSpec{t, kbp, (long long)bl.col_off * Bp});
        }
      for (int64_t m0 = 0; m0 < bl.rows; m0 += gm)
        for (int n0 = 0; n0 < Bq; n0 += g->tn) {
          GemmItem t{};
          t.Ablk = bl.dUb;
          t.ablk_kblocks = kbb;
          t.bcontig = Bq <= g->tn ? 1 : 0;
          t.B = bk.Bm.p + (size_t)bl.col_off * Bp;
          t.ldb = Bp;
          t.ldc = Bp;
          t.M = (int)bl.rows;
          t.N = Bq;
          t.K = (int)bl.nr;
          t.m0 = (int)m0;
          t.n0 = n0;
          bt.push_back(TileSpec{t, kbb, (long long)bl.pos0 * Bp});
        }
    }
    streamk(pt, nullptr, 0, &bk.ksplit_proj, &bk.grid_proj);
    streamk(bt, nullptr, 0, &bk.ksplit_back, &bk.grid_back);
    BGGE_CUDA(bk.Dpart.alloc((size_t)bk.ksplit_proj * bk.nr_pad * Bp));
    BGGE_CUDA(cudaMemsetAsync(bk.Dpart.p, 0, (size_t)bk.ksplit_proj * bk.nr_pad * Bp * sizeof(double), st));
    BGGE_CUDA(bk.Upart.alloc(total * bk.ksplit_back));
    BGGE_CUDA(cudaMemsetAsync(bk.Upart.p, 0, total * bk.ksplit_back * sizeof(double), st));
    std::vector<GemmItem> pi = streamk(pt, bk.Dpart.p, bk.nr_pad * Bp, &bk.ksplit_proj, &bk.grid_proj);
    std::vector<GemmItem> bi = streamk(bt, bk.Upart.p, (long long)total, &bk.ksplit_back, &bk.grid_back);
    bk.n_proj = (int)pi.size();
    bk.n_back = (int)bi.size();
    BGGE_TRY(up(bk.proj_items, pi, st));
    BGGE_TRY(up(bk.back_items, bi, st));
  }
  const unsigned ig = (unsigned)((total + BT - 1) / BT);
  b_init_kernel<<<ig, BT, 0, st>>>(n, g->nk, B, Bp, g->Y.p, g->mu0.p, g->U.p, g->R.p, g->draw, q, g->xf.p, g->Bet.p, g->Ep.p);
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

// Resident sweep for small models (wheat-sized configs C1/C2: the whole eigen basis is a few MB).
// One cooperative persistent kernel runs `iters` Gibbs iterations per launch: every CTA keeps its column
// slices of every basis matrix in shared memory for the whole run, the phases of an iteration are separated
// by grid-wide barriers instead of kernel launches (the streaming path needs 3 launches per kernel step and
// is launch-latency bound at this size), and the chain scalars are computed redundantly (bit-identically)
// by every CTA so they need no broadcast.
//
// A kernel step j works on groups of blocks that share one matrix W (t x nr; a dense block is a group of its
// own with the identity row map).  For its own columns a CTA does
//   P1  e~_m = etil0_j[m] - shift * wsum_j[m]          (etil0 = C^-1/2 Z'(r + u_j), produced by the previous P5)
//   P2  d = W[:,c]' e~_m,  b ~ N(tau v d, v)           one warp per (column, member block)    R/BGGE.R:298-305
//   P3  partial (W b)_m over its columns                -> part[cta][m][t]
//   -- grid barrier --
//   P5  rows of y, distributed by the virtual rows of the NEXT kernel step: u_i = isq_i * sum_cta part[..][loc_i],
//       residual / u / Welford update (R/BGGE.R:306-307), and the next step's etil0 in the same pass
//   -- grid barrier --
// and after the last kernel step every CTA computes sigb_j, sigsq, tau and the next intercept from per-CTA
// partial sums (R/BGGE.R:360-367, 278-281); missing phenotypes (R/BGGE.R:370-381) cost two more barriers.
// All sums run in a fixed order: results do not depend on how a run is split into launches.
//
// Merged steps: consecutive kernels that consist of one block each, share one basis matrix and cover disjoint
// rows (the environment kernels of MDe in a balanced trial) are conditionally independent given the rest, so up
// to 4 of them are swept as ONE step (one group, one member per kernel, each with its own sigb / draws / u).
// This needs u_k = 0 outside kernel k's block, which holds after the first iteration: gibbs_run runs iteration
// 1 on the streaming path when the plan contains a merged step.
#include "gibbs.cuh"
#include "kernels.h"
#include "rng.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;

namespace bgge {
namespace {

constexpr int R_THREADS = 512;
constexpr int R_WARPS = R_THREADS / 32;
constexpr int R_SUB = 8;                      // lanes that share one row in the row phase
constexpr int R_NPL = 19;                     // partials per lane: supports up to R_SUB * R_NPL = 152 CTAs
constexpr int R_MAXCOLS = 64;                 // columns of one matrix per CTA (bsm capacity)
constexpr int R_MAXG = 32;                    // basis matrices (groups) per model
constexpr int R_MAXKR = 16;                   // kernels per model
constexpr int R_MAXQ = 8;                     // fixed-effect columns

struct RParams {
  const RGroup* groups;
  const RMember* members;
  const RStep* steps;
  int nk, nsteps, ncta, ngroups;
  long long n;
  long long nr_k[R_MAXKR];     // kept eigenvalues per kernel (chi-square degrees of freedom)
  double* r;
  double* u;
  double* y;
  double* umean;
  double* um2;
  Scal* sc;
  double* part;               // [ncta][pstride]
  long long pstride;
  double* zqpart;             // [2][nk][ncta], double-buffered by iteration parity
  double* sspart;             // [3][ncta]: sum e^2, sum r, NA change of sum r
  const int* na_idx;
  long long n_na;
  double* chain_mu;
  double* chain_varE;
  double* chain_varU;
  long long thin, burn, ncum;
  long long iters;
  int etp;                    // padded length of one e~ vector in shared memory
  long long smem_w;           // doubles of basis slices per CTA
  int max_entries, max_units, ndraw, ncolc, nwsum;
  // fixed effects (R/BGGE.R:284-290): the kernel keeps r WITHOUT the fixed-effect part (r_s = r + XF b), so the
  // XF step needs no pass over the rows: XF'(e + XF b_old) = XF' r_s - shift * XF'1
  int q;
  const double* xf;           // n x q, column-major
  const double* txx;          // (XF'XF)^-1, q x q
  const double* lchol;        // lower Cholesky factor of txx
  const double* x1;           // XF'1
  const double* mx0;          // [q][ps0]: C^-1/2 Z' XF of step 0 (its residual is formed before the new b is known)
  long long ps0;
  double* bet;                // [q] current coefficients (state shared with the streaming path)
  double* chain_bet;
  double* xrpart;             // [q][ncta]: XF' r_s pieces
  unsigned long long* bar;    // grid barrier counter (zeroed before the launch)
  unsigned long long* timing; // debug (BGGE_RESIDENT_TIMING): ns per phase, accumulated by CTA 0
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");
  return t;
}
#define R_TICK(slot)                                         \
  if (TIMED && timed) {                                      \
    const unsigned long long now_ = gtime();                 \
    tacc[slot] += now_ - tlast;                              \
    tlast = now_;                                            \
  }

struct RShared {
  RGroup groups[R_MAXG];
  RMember members[4 * R_MAXG];
  RStep steps[R_MAXKR];
  int nent[R_MAXKR], nun[R_MAXKR];
  int c0[R_MAXG], ncol[R_MAXG];        // my columns of every basis matrix
  double red[2 + R_MAXQ][R_WARPS];
  double tot[R_MAXKR + 2 + R_MAXQ];
  double bet[R_MAXQ], betn[R_MAXQ], zxf[2][R_MAXQ], x1[R_MAXQ];
  double txx[R_MAXQ * R_MAXQ], lch[R_MAXQ * R_MAXQ];
  double bsm[4 * R_MAXCOLS];
  double dpart[4 * R_MAXCOLS];
  double chi[2][R_MAXKR + 1];
  double zmu[2];
  double sigb[R_MAXKR];
  double Sc[R_MAXKR];
  long long nrk[R_MAXKR];
};

// Split grid barrier (all CTAs are co-resident: cooperative launch).  Work placed between arrive and wait
// overlaps the barrier latency; it must be CTA-local.
struct GridBar {
  unsigned long long* ctr;
  unsigned long long target;
  unsigned ncta;
};
__device__ __forceinline__ void bar_arrive(GridBar& b) {
  __syncthreads();
  b.target += b.ncta;
  if (threadIdx.x == 0)   // release: orders the CTA's writes (made visible to this thread by the bar.sync) before the count
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(b.ctr), "l"(1ull) : "memory");
}
__device__ __forceinline__ void bar_wait(const GridBar& b) {
  if (threadIdx.x == 0) {
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(b.ctr) : "memory");
    } while (v < b.target);
  }
  __syncthreads();
}

struct RSmem {                 // dynamic shared memory carve-up
  double* wsm;                 // basis slices
  double* etil;                // [4][etp]
  double* vals;                // [max_entries]
  double* rvals;               // [max_entries] updated residuals of my rows (XF' r_s pieces)
  double* zbuf;                // [2][ndraw] coefficient normals of my columns, by iteration parity
  double* svs;                 // [ncolc] eigenvalues of my columns
  double* isv;                 // 1 / eigenvalue
  double* tv;                  // [ndraw] tau * v per (member, column)   (R/BGGE.R:302-303)
  double* sd;                  // [ndraw] sqrt(v)
  double* wsum;                // [nwsum] row-weight sums per virtual row, all kernels (offset RStep::woff)
  REntry* sent;                // [nk][max_entries] my rows of every row phase
  RUnit* sunit;                // [nk][max_units]
};

// Draws of iteration itn that do not depend on the data: the normals of my coefficient columns, the chi-squares
// and the normal of the intercept that follows itn.  Stored under parity itn & 1.  (One copy of this code: the
// kernel is instruction-fetch sensitive, every phase runs once per iteration.)
__device__ __noinline__ void draws_for(DrawCfg dc, RShared* shp, double* zbuf, int ndraw, int ngroups, int nk, int q, long long n,
                                       long long itn, bool normals, bool scalars) {
  RShared& sh = *shp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par = (int)(itn & 1);
  if (normals)
    for (int idx = tid; idx < ndraw; idx += R_THREADS) {
      int g = 0;
      while (g + 1 < ngroups && idx >= sh.groups[g + 1].doff) ++g;
      const RGroup& gr = sh.groups[g];
      const int f = idx - gr.doff, mm = f / gr.maxcols, cc = f % gr.maxcols;
      if (cc < sh.ncol[g]) {
        const RMember& mbr = sh.members[gr.mem0 + mm];
        const long long col = (long long)mbr.col_off + sh.c0[g] + cc;
        zbuf[(size_t)par * ndraw + idx] = draw_normal(dc, itn, (uint32_t)mbr.kidx, (unsigned long long)col,
                                                      dc.z_init + (itn - 1) * dc.z_per_it + mbr.z_off + col);
      }
    }
  // scalar draws, one per warp from the last warp down: the kernels' chi-squares go with the normals (first
  // barrier of an iteration), the residual chi-square and the intercept normal with the second
  if (lane == 0) {
    for (int d = R_WARPS - 1 - warp; d <= nk + 1 + q; d += R_WARPS) {
      if (!(d < nk ? normals : scalars)) continue;
      if (d > nk + 1) {   // fixed-effect normals of the iteration that follows itn (R/BGGE.R:107)
        const int k = d - nk - 2;
        sh.zxf[par][k] = draw_normal(dc, itn + 1, STREAM_XF, (unsigned long long)k, dc.z_init + itn * dc.z_per_it + dc.z_off_xf + k);
        continue;
      }
      if (d < nk)
        sh.chi[par][d] = draw_chisq(dc, itn, STREAM_CHI_K + (uint32_t)d, (double)sh.nrk[d] + NU, (itn - 1) * dc.c_per_it + d);
      else if (d == nk)
        sh.chi[par][nk] = draw_chisq(dc, itn, STREAM_CHI_E, (double)n + NU, (itn - 1) * dc.c_per_it + nk);
      else
        sh.zmu[par] = draw_normal(dc, itn + 1, STREAM_MU, 0ull, dc.z_init + itn * dc.z_per_it);
    }
  }
}

// per-(member, column) constants of the coefficient draw for the current sigb / tau (R/BGGE.R:302-303); same
// layout as the normals (group offset doff, member-major)
__device__ __forceinline__ void column_constants(RShared* shp, const double* svs, double* tv, double* sd, int ndraw, int ngroups,
                                                 double tau) {
  RShared& sh = *shp;
  for (int idx = R_THREADS - 1 - (int)threadIdx.x; idx < ndraw; idx += R_THREADS) {
    int g = 0;
    while (g + 1 < ngroups && idx >= sh.groups[g + 1].doff) ++g;
    const RGroup& gr = sh.groups[g];
    const int f = idx - gr.doff, mm = f / gr.maxcols, cc = f - mm * gr.maxcols;
    const double sv = svs[gr.coff + cc];
    const double lam = sh.sigb[sh.members[gr.mem0 + mm].kidx];
    const double vari = sv * lam / (1.0 + sv * lam * tau);
    tv[idx] = tau * vari;
    sd[idx] = sqrt(vari);
  }
}

// b = tXX (XF' r_s - shift XF'1) + sqrt(sigsq) L z  (R/BGGE.R:286-288, 106-109); every thread calls, result in sh.bet
__device__ __forceinline__ void xf_draw(RShared& sh, int q, const double* xr, double shift, double sigsq, const double* z) {
  const int k = threadIdx.x;
  double bn = 0.0;
  if (k < q) {
    double media = 0.0, lz = 0.0;
    for (int k2 = 0; k2 < q; ++k2) media += sh.txx[k + k2 * q] * (xr[k2] - shift * sh.x1[k2]);
    for (int k2 = 0; k2 <= k; ++k2) lz += sh.lch[k + k2 * q] * z[k2];
    bn = media + sqrt(sigsq) * lz;
  }
  __syncthreads();
  if (k < q) sh.bet[k] = bn;
  __syncthreads();
}

// Row phase of step j.  mode 0: update (u_i from the partial sums, r, Welford) and the next step's etil0;
// mode 1: only the gather of the next step's etil0 (after the imputation of missing phenotypes);
// mode 2: launch prologue: as 1 and r -> r_s = r + XF b;  mode 3: launch epilogue: r_s -> r = r_s - XF b.
// last (mode 0 on the last step, always for modes 1-2): the etil0 produced is for step 0 of the NEXT iteration, whose
// fixed effects are not known yet, so it carries no XF term (P1 of step 0 subtracts it) and the per-CTA pieces of
// XF' r_s are formed; other steps get etil0 with the current XF b removed.
template <bool TIMED>
__device__ __forceinline__ void rows_phase(const RParams& p, RShared& sh, const RSmem& m, int j, int mode, bool last,
                                           double shift, bool keep, double kcount, unsigned long long* tacc = nullptr,
                                           unsigned long long* tlast_p = nullptr) {
  const bool timed = TIMED && tacc != nullptr;
  unsigned long long tlast = timed ? *tlast_p : 0ull;
  const bool update = mode == 0;
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int sub = tid / R_SUB, sl = tid % R_SUB;
  const int ne = sh.nent[j], nu = sh.nun[j];
  const REntry* ent = m.sent + (size_t)j * p.max_entries;
  const RUnit* units = m.sunit + (size_t)j * p.max_units;
  const int jn = sh.steps[j].next;
  const int q = p.q;
  double sse = 0.0, sr = 0.0;
  for (int k0 = 0; k0 < ne; k0 += R_THREADS / R_SUB) {
    const int k = k0 + sub;
    const bool valid = k < ne;
    REntry E{};
    E.kj = E.kn = -1;
    if (valid) E = ent[k];
    // kj: the kernel of this step that owns the row (-1: none, a merged step away from its blocks);
    // kn: the kernel whose u enters the next step's residual for this row (-1: none)
    const bool upd = update && E.kj >= 0;
    double* ujp = p.u + (size_t)(upd ? E.kj : 0) * p.n + E.row;
    double rold = 0.0, uold = 0.0, ujn = 0.0, mm = 0.0, m2 = 0.0, xfb = 0.0;
    if (valid && sl == 0) {
      rold = __ldcg(p.r + E.row);
      if (upd) uold = __ldcg(ujp);
      if (E.kn >= 0 && !(upd && E.kn == E.kj)) ujn = __ldcg(p.u + (size_t)E.kn * p.n + E.row);
      if (upd && keep) {
        mm = __ldcg(p.umean + (size_t)E.kj * p.n + E.row);
        m2 = __ldcg(p.um2 + (size_t)E.kj * p.n + E.row);
      }
      if (q > 0 && mode != 1)
        for (int kk = 0; kk < q; ++kk) xfb = fma(__ldg(p.xf + (size_t)kk * p.n + E.row), sh.bet[kk], xfb);
    }
    double v;
    if (update) {
      double x[R_NPL];
      const bool live = valid && upd && E.poff >= 0;
      const double* pp = p.part + (live ? E.poff : 0) + (size_t)sl * p.pstride;
      const size_t pstep = (size_t)R_SUB * p.pstride;
      const int nq = live ? (p.ncta - sl + R_SUB - 1) / R_SUB : 0;   // partials this lane reads
#pragma unroll
      for (int qq = 0; qq < R_NPL; ++qq) {
        x[qq] = qq < nq ? __ldcg(pp) : 0.0;
        pp += pstep;
      }
      R_TICK(14)
      double un = ((((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]))) +
                   (((x[8] + x[9]) + (x[10] + x[11])) + ((x[12] + x[13]) + (x[14] + x[15])))) +
                  ((x[16] + x[17]) + x[18]);
      un += __shfl_xor_sync(0xffffffffu, un, 4);
      un += __shfl_xor_sync(0xffffffffu, un, 2);
      un += __shfl_xor_sync(0xffffffffu, un, 1);
      un *= E.isq;
      if (TIMED && timed && un == 123.456) tacc[0] += 1;   // force the sum to be complete at the tick
      R_TICK(15)
      const double rn = upd ? (rold + uold) - un : rold;
      if (valid && sl == 0) {
        if (upd) {
          p.r[E.row] = rn;
          *ujp = un;
          if (keep) {
            const double dlt = un - mm;
            const double mn = mm + dlt / kcount;
            p.umean[(size_t)E.kj * p.n + E.row] = mn;
            p.um2[(size_t)E.kj * p.n + E.row] = m2 + dlt * (un - mn);
          }
        }
        if (last) {
          const double e = (rn - shift) - xfb;
          sse = fma(e, e, sse);
          sr += rn;
          if (q > 0) m.rvals[E.vpos] = rn;
        }
      }
      v = E.w * ((last ? rn : rn - xfb) + (upd && E.kn == E.kj ? un : ujn));
    } else if (mode == 3) {
      if (valid && sl == 0) p.r[E.row] = rold - xfb;
      v = 0.0;
    } else {
      const double rs = mode == 2 ? rold + xfb : rold;
      if (valid && sl == 0 && q > 0) {
        if (mode == 2) p.r[E.row] = rs;
        m.rvals[E.vpos] = rs;
      }
      v = E.w * (rs + ujn);
    }
    if (valid && sl == 0) m.vals[E.vpos] = v;
  }
  R_TICK(11)
  __syncthreads();
  R_TICK(12)
  if (mode == 3) return;
  double* etil0 = sh.steps[jn].etil0;
  for (int un = tid; un < nu; un += R_THREADS) {
    const RUnit U = units[un];
    if (U.eoff < 0) continue;
    double acc = 0.0;
    for (int k = U.e0; k < U.e1; ++k) acc += m.vals[k];
    etil0[U.eoff] = acc;
  }
  R_TICK(13)
  if (last && update) {   // per-CTA pieces of sum e^2 and sum r (fixed order)
    sse = warp_sum(sse);
    sr = warp_sum(sr);
    if ((tid & 31) == 0) {
      sh.red[0][tid >> 5] = sse;
      sh.red[1][tid >> 5] = sr;
    }
  }
  if (last && q > 0) {    // per-CTA pieces of XF' r_s: 64 threads per column, rows in storage order
    const int kk = tid >> 6, t = tid & 63;
    double acc = 0.0;
    if (kk < q)
      for (int e = t; e < ne; e += 64) acc = fma(__ldg(p.xf + (size_t)kk * p.n + ent[e].row), m.rvals[ent[e].vpos], acc);
    acc = warp_sum(acc);
    if ((tid & 31) == 0) sh.red[2 + (tid >> 6)][(tid >> 5) & 1] = acc;
  }
  if (last) {
    __syncthreads();
    if (update && tid < 2) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < R_WARPS; ++w) t += sh.red[tid][w];
      p.sspart[(size_t)tid * p.ncta + cta] = t;
    }
    if (tid >= 32 && tid < 32 + q) {
      const int kk = tid - 32;
      p.xrpart[(size_t)kk * p.ncta + cta] = sh.red[2 + kk][0] + sh.red[2 + kk][1];
    }
  }
  if (timed) *tlast_p = tlast;
}

template <bool TIMED>
__global__ void __launch_bounds__(R_THREADS, 1)
resident_sweep_kernel(const RParams p, DrawCfg dc) {
  extern __shared__ __align__(16) unsigned char rsm[];
  __shared__ RShared sh;
  const int cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  RSmem m;
  m.wsm = reinterpret_cast<double*>(rsm);
  m.etil = m.wsm + p.smem_w;
  m.vals = m.etil + 4 * p.etp;
  m.rvals = m.vals + p.max_entries;
  m.zbuf = m.rvals + p.max_entries;
  m.svs = m.zbuf + 2 * p.ndraw;
  m.isv = m.svs + p.ncolc;
  m.tv = m.isv + p.ncolc;
  m.sd = m.tv + p.ndraw;
  m.wsum = m.sd + p.ndraw;
  m.sent = reinterpret_cast<REntry*>(m.wsum + p.nwsum);
  m.sunit = reinterpret_cast<RUnit*>(m.sent + (size_t)p.nsteps * p.max_entries);
  GridBar gb{p.bar, 0ull, (unsigned)p.ncta};

  // ---- descriptors into shared memory (every later phase starts from them: no dependent global loads)
  {
    auto copy8 = [&](void* dst, const void* src, size_t bytes) {
      unsigned long long* d = reinterpret_cast<unsigned long long*>(dst);
      const unsigned long long* s_ = reinterpret_cast<const unsigned long long*>(src);
      for (size_t i = tid; i < bytes / 8; i += R_THREADS) d[i] = s_[i];
    };
    copy8(sh.groups, p.groups, (size_t)p.ngroups * sizeof(RGroup));
    copy8(sh.members, p.members, (size_t)p.ngroups * 4 * sizeof(RMember));
    copy8(sh.steps, p.steps, (size_t)p.nsteps * sizeof(RStep));
  }
  __syncthreads();
  if (tid < p.ngroups) {
    const int nr = sh.groups[tid].nr;
    const int c0 = (int)((long long)nr * cta / p.ncta), c1 = (int)((long long)nr * (cta + 1) / p.ncta);
    sh.c0[tid] = c0;
    sh.ncol[tid] = c1 - c0;
  }
  __syncthreads();
  // ---- my column slices of every basis matrix, their eigenvalues, row weights, my rows of every row phase
  for (int g = 0; g < p.ngroups; ++g) {
    const RGroup& gr = sh.groups[g];
    const int c0 = sh.c0[g], c1 = c0 + sh.ncol[g];
    double* dst = m.wsm + gr.woff;
    for (int c = c0; c < c1; ++c)
      for (int a = tid; a < gr.t; a += R_THREADS) dst[(size_t)(c - c0) * gr.t_pad + a] = gr.W[(long long)c * gr.ldw + a];
    for (int cc = tid; cc < gr.maxcols; cc += R_THREADS) {
      const double sv = cc < c1 - c0 ? gr.s[c0 + cc] : 1.0;
      m.svs[gr.coff + cc] = sv;
      m.isv[gr.coff + cc] = 1.0 / sv;
    }
  }
  for (int j = 0; j < p.nsteps; ++j) {
    const RStep& S = sh.steps[j];
    for (int a = tid; a < S.nws; a += R_THREADS) m.wsum[S.woff + a] = S.wsum[a];
    const int u0 = (int)((long long)S.nunits * cta / p.ncta), u1 = (int)((long long)S.nunits * (cta + 1) / p.ncta);
    const int eb = u0 < u1 ? S.units[u0].e0 : 0, ee = u0 < u1 ? S.units[u1 - 1].e1 : 0;
    for (int k = eb + tid; k < ee; k += R_THREADS) m.sent[(size_t)j * p.max_entries + (k - eb)] = S.entries[k];
    for (int un = u0 + tid; un < u1; un += R_THREADS) {
      RUnit U = S.units[un];
      U.e0 -= eb;
      U.e1 -= eb;
      m.sunit[(size_t)j * p.max_units + (un - u0)] = U;
    }
    if (tid == 0) {
      sh.nent[j] = ee - eb;
      sh.nun[j] = u1 - u0;
    }
  }
  // chain scalars: every CTA holds (and updates identically) its own copy
  Scal* sc = p.sc;
  double mu = sc->mu, shift = sc->shift, sigsq = sc->sigsq, tau = sc->tau, mu_done = sc->mu_done;
  const double mu0 = sc->mu0, Sce = sc->Sce;
  long long it = sc->it;
  for (int j = tid; j < p.nk; j += R_THREADS) {
    sh.sigb[j] = sc->sigb[j];
    sh.Sc[j] = sc->Sc[j];
    sh.nrk[j] = p.nr_k[j];
  }
  for (int k = tid; k < p.q; k += R_THREADS) {
    sh.bet[k] = p.bet[k];
    sh.x1[k] = p.x1[k];
  }
  for (int k = tid; k < p.q * p.q; k += R_THREADS) {
    sh.txx[k] = p.txx[k];
    sh.lch[k] = p.lchol[k];
  }
  __syncthreads();
  column_constants(&sh, m.svs, m.tv, m.sd, p.ndraw, p.ngroups, tau);
  draws_for(dc, &sh, m.zbuf, p.ndraw, p.ngroups, p.nk, p.q, p.n, it, true, true);
  // first kernel's etil0 from the state in memory
  rows_phase<false>(p, sh, m, p.nsteps - 1, p.q > 0 ? 2 : 1, true, 0.0, false, 1.0);
  bar_arrive(gb);
  bar_wait(gb);
  if (p.q > 0) {   // fixed effects of the first iteration of this launch (later ones are drawn at the end of the one before)
    for (int v = warp; v < p.q; v += R_WARPS) {
      double x = 0.0;
      for (int c = lane; c < p.ncta; c += 32) x += __ldcg(p.xrpart + (size_t)v * p.ncta + c);
      x = warp_sum(x);
      if (lane == 0) sh.tot[v] = x;
    }
    if (tid < p.q)
      sh.betn[tid] = draw_normal(dc, it, STREAM_XF, (unsigned long long)tid, dc.z_init + (it - 1) * dc.z_per_it + dc.z_off_xf + tid);
    __syncthreads();
    xf_draw(sh, p.q, sh.tot, shift, sigsq, sh.betn);
  }

  const bool timed = p.timing && cta == 0 && tid == 0;
  unsigned long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long tlast = gtime();
  long long tmod = it % p.thin, qit = it / p.thin;   // it % thin and it / thin, kept incrementally
  const long long qburn = p.burn / p.thin;
  for (long long step = 0; step < p.iters; ++step, ++it) {
    const bool keep = (tmod == 0) && (it > p.burn);
    const double kcount = keep ? (double)(qit - qburn) : 1.0;
    const int par = (int)(it & 1);
    for (int j = 0; j < p.nsteps; ++j) {
      const RStep& S = sh.steps[j];
      double zq = 0.0;   // b^2 / s of the (column, member) this thread draws; member slot = tid % nb of the group
      for (int g = S.g0; g < S.g1; ++g) {
        const RGroup& gr = sh.groups[g];
        const int ncol = sh.ncol[g], nb = gr.nb;
        const double* W = m.wsm + gr.woff;
        // P1: residual in genotype space for every member block (the members' virtual rows are contiguous:
        // one flat vector of nb * t entries); loads are issued four deep
        const int tg = gr.t, tot = nb * tg;
        {
          const long long po = sh.members[gr.mem0].part_off;
          const double* e0 = S.etil0 + po;
          const double* ws = m.wsum + S.woff + po;
          double* et = m.etil;
          for (int i0 = tid; i0 < tot; i0 += 4 * R_THREADS) {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int idx = i0 + q * R_THREADS;
              v[q] = idx < tot ? __ldcg(e0 + idx) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int idx = i0 + q * R_THREADS;
              if (idx < tot) {
                double e = v[q] - shift * ws[idx];
                if (j == 0 && p.q > 0)   // step 0: etil0 was formed before this iteration's fixed effects were drawn
                  for (int kk = 0; kk < p.q; ++kk) e -= __ldg(p.mx0 + (size_t)kk * p.ps0 + po + idx) * sh.bet[kk];
                et[idx] = e;
              }
            }
          }
        }
        __syncthreads();
        R_TICK(5)
        // P2: W[:,c]' e~ for all members at once; a column is split over nch warps
        int nch = ncol > 0 ? max(1, R_WARPS / ncol) : 1;
        nch = min(nch, (gr.t + 63) >> 6);
        const int clen = (((gr.t + nch - 1) / nch) + 31) & ~31;
        for (int w = warp; w < ncol * nch; w += R_WARPS) {
          const int cc = w / nch, ch = w - cc * nch;
          const int a0 = ch * clen, a1 = min(gr.t, a0 + clen);
          const double* wc = W + (size_t)cc * gr.t_pad;
          double d[4] = {0.0, 0.0, 0.0, 0.0};
          const double* et = m.etil;
          for (int a = a0 + lane; a < a1; a += 32) {
            const double wv = wc[a];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
              if (mb < nb) d[mb] = fma(wv, et[mb * tg + a], d[mb]);
          }
#pragma unroll
          for (int mb = 0; mb < 4; ++mb)
            if (mb < nb) {
              const double t = warp_sum(d[mb]);
              if (lane == 0) sh.dpart[w * 4 + mb] = t;
            }
        }
        __syncthreads();
        R_TICK(6)
        // coefficient draw: one thread per (column, member)   R/BGGE.R:298-305
        if (tid < ncol * nb) {
          const int cc = tid / nb, mb = tid % nb;
          double d = 0.0;
          for (int ch = 0; ch < nch; ++ch) d += sh.dpart[(cc * nch + ch) * 4 + mb];
          const int di = gr.doff + mb * gr.maxcols + cc;
          const double z = m.zbuf[(size_t)par * p.ndraw + di];
          const double bv = m.tv[di] * d + m.sd[di] * z;
          sh.bsm[mb * R_MAXCOLS + cc] = bv;
          zq += bv * bv * m.isv[gr.coff + cc];                              // R/BGGE.R:96
        }
        __syncthreads();
        R_TICK(7)
        // P3: my columns' contribution to W b, all members of the group per row
        {
          double* out = p.part + (size_t)cta * p.pstride + sh.members[gr.mem0].part_off;
          const int tp = gr.t_pad;
          for (int a = tid; a < tg; a += R_THREADS) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int cc = 0; cc < ncol; ++cc) {
              const double wv = W[(size_t)cc * tp + a];
#pragma unroll
              for (int mb = 0; mb < 4; ++mb)
                if (mb < nb) acc[mb] = fma(wv, sh.bsm[mb * R_MAXCOLS + cc], acc[mb]);
            }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
              if (mb < nb) out[mb * tg + a] = acc[mb];
          }
        }
        if (g + 1 < S.g1) __syncthreads();
        R_TICK(8)
      }
      // sum of b^2/s of my columns: one kernel (every member slot adds up), or one kernel per member slot (merged step)
      {
        const int slot = S.merged ? tid % sh.groups[S.g0].nb : 0;   // the drawing threads are tid = cc * nb + mb
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q == 0 || S.merged) {
            const double t = warp_sum(slot == q ? zq : 0.0);
            if (lane == 0) sh.red[q][warp] = t;
          }
        }
        __syncthreads();
        if (tid < (S.merged ? sh.groups[S.g0].nb : 1)) {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < R_WARPS; ++w) t += sh.red[tid][w];
          p.zqpart[((size_t)par * p.nk + S.zk[tid]) * p.ncta + cta] = t;
        }
      }
      R_TICK(0)
      bar_arrive(gb);
      R_TICK(9)
      if (j == 0 && step + 1 < p.iters) draws_for(dc, &sh, m.zbuf, p.ndraw, p.ngroups, p.nk, p.q, p.n, it + 1, true, false);   // hidden behind the barrier
      R_TICK(10)
      bar_wait(gb);
      R_TICK(1)
      rows_phase<TIMED>(p, sh, m, j, 0, j == p.nsteps - 1, shift, keep, kcount, timed ? tacc : nullptr, &tlast);
      R_TICK(2)
      bar_arrive(gb);
      if (j == 0 && step + 1 < p.iters) draws_for(dc, &sh, m.zbuf, p.ndraw, p.ngroups, p.nk, p.q, p.n, it + 1, false, true);
      bar_wait(gb);
      R_TICK(3)
    }
    // ---- scalars, identically in every CTA: sum the per-CTA pieces (R/BGGE.R:360-367)
    for (int v = warp; v < p.nk + 2 + p.q; v += R_WARPS) {
      const double* src = v < p.nk       ? p.zqpart + ((size_t)par * p.nk + v) * p.ncta
                          : v < p.nk + 2 ? p.sspart + (size_t)(v - p.nk) * p.ncta
                                         : p.xrpart + (size_t)(v - p.nk - 2) * p.ncta;
      double x = 0.0;
      for (int c = lane; c < p.ncta; c += 32) x += __ldcg(src + c);
      x = warp_sum(x);
      if (lane == 0) sh.tot[v] = x;
    }
    __syncthreads();
    double sr = sh.tot[p.nk + 1];   // sum of r_s; the fixed-effect part comes off below
    sigsq = (sh.tot[p.nk] + Sce) / sh.chi[par][p.nk];
    tau = 1.0 / sigsq;
    if (tid < p.nk) sh.sigb[tid] = (sh.tot[tid] + NU * sh.Sc[tid]) / sh.chi[par][tid];
    __syncthreads();
    column_constants(&sh, m.svs, m.tv, m.sd, p.ndraw, p.ngroups, tau);
    if (p.n_na > 0) {   // missing phenotypes (R/BGGE.R:370-381): rows split over the CTAs
      const double sd = sqrt(sigsq);
      double dsum = 0.0;
      for (long long k = (long long)cta * R_THREADS + tid; k < p.n_na; k += (long long)p.ncta * R_THREADS) {
        const long long i = p.na_idx[k];
        double uhat = 0.0;
        for (int j = 0; j < p.nk; ++j) uhat = (j == 0) ? __ldcg(p.u + i) : uhat + __ldcg(p.u + (size_t)j * p.n + i);
        double aux = 0.0;
        for (int kk = 0; kk < p.q; ++kk) aux += __ldg(p.xf + (size_t)kk * p.n + i) * sh.bet[kk];
        const double z = draw_normal(dc, it, STREAM_NA, (unsigned long long)k, dc.z_init + (it - 1) * dc.z_per_it + dc.z_off_na + k);
        const double yv = ((aux + mu) + uhat) + sd * z;
        p.y[i] = yv;
        const double rn = ((((yv - uhat) - aux) - mu) + shift) + aux;   // r_s keeps the fixed-effect part
        dsum += rn - __ldcg(p.r + i);
        p.r[i] = rn;
      }
      dsum = warp_sum(dsum);
      if (lane == 0) sh.red[0][warp] = dsum;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < R_WARPS; ++w) t += sh.red[0][w];
        p.sspart[(size_t)2 * p.ncta + cta] = t;
      }
      bar_arrive(gb);
      bar_wait(gb);
      rows_phase<false>(p, sh, m, p.nsteps - 1, 1, true, 0.0, false, 1.0);   // also renews the XF' r_s pieces
      bar_arrive(gb);
      bar_wait(gb);
      double x = 0.0;
      for (int c = lane; c < p.ncta; c += 32) x += __ldcg(p.sspart + (size_t)2 * p.ncta + c);
      sr += warp_sum(x);   // every warp computes the same sum
      for (int v = warp; v < p.q; v += R_WARPS) {
        double xx = 0.0;
        for (int c = lane; c < p.ncta; c += 32) xx += __ldcg(p.xrpart + (size_t)v * p.ncta + c);
        xx = warp_sum(xx);
        if (lane == 0) sh.tot[p.nk + 2 + v] = xx;
      }
      __syncthreads();
    }
    for (int kk = 0; kk < p.q; ++kk) sr -= sh.x1[kk] * sh.bet[kk];   // sum r = sum r_s - 1'XF b
    if (cta == 0 && tid == 0 && tmod == 0) {   // thinned chain (R/BGGE.R:384-395)
      const long long slot = qit - 1;
      if (slot < p.ncum) {
        p.chain_varE[slot] = sigsq;
        p.chain_mu[slot] = mu;
        for (int j = 0; j < p.nk; ++j) p.chain_varU[(long long)j * p.ncum + slot] = sh.sigb[j];
        for (int kk = 0; kk < p.q; ++kk) p.chain_bet[slot * p.q + kk] = sh.bet[kk];
      }
    }
    // next intercept (R/BGGE.R:278-281): mean(temp + mu) = mean(r) + mu0
    mu_done = mu;
    mu = (sr / (double)p.n + mu0) + sqrt(sigsq / (double)p.n) * sh.zmu[par];
    shift = mu - mu0;
    if (p.q > 0 && step + 1 < p.iters) {   // fixed effects of the next iteration (R/BGGE.R:284-290)
      __syncthreads();                     // sh.bet was read above (chain record, sum r)
      xf_draw(sh, p.q, sh.tot + p.nk + 2, shift, sigsq, sh.zxf[par]);
    }
    if (++tmod == p.thin) {
      tmod = 0;
      ++qit;
    }
    R_TICK(4)
  }
  if (TIMED && timed) {
#pragma unroll
    for (int k = 0; k < 16; ++k) p.timing[k] = tacc[k];
  }
  if (p.q > 0) {   // back to the streaming path's convention: r without the fixed-effect part
    __syncthreads();
    rows_phase<false>(p, sh, m, p.nsteps - 1, 3, false, 0.0, false, 1.0);
    if (cta == 0 && tid < p.q) p.bet[tid] = sh.bet[tid];
  }
  if (cta == 0 && tid == 0) {   // hand the chain state back (fetch functions, streaming path)
    sc->mu = mu;
    sc->mu_done = mu_done;
    sc->shift = shift;
    sc->sigsq = sigsq;
    sc->tau = tau;
    sc->it = it;
    for (int j = 0; j < p.nk; ++j) sc->sigb[j] = sh.sigb[j];
  }
}

template <class T>
int upload_vec(DevBuf<T>& d, const std::vector<T>& h, cudaStream_t st) {
  BGGE_CUDA(d.alloc(std::max<size_t>(1, h.size())));
  if (!h.empty()) BGGE_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return OK;
}

}  // namespace

// Decide whether the whole model fits the resident kernel; build its tables.  Called at the end of gibbs_init.
int plan_resident(Gibbs* g) {
  ResidentPlan& rp = g->resident;
  rp.enabled = false;
  rp.has_merged = false;
  if (g->mode != 1 || g->shard_world > 1 || getenv("BGGE_NO_RESIDENT")) return OK;
  if (g->n > 32768 || g->nk > R_MAXKR || g->q > R_MAXQ) return OK;
  cudaStream_t st = g->stream;
  const int sms = sm_count();
  const int64_t n = g->n;
  const int nk = g->nk;
  const bool allow_merge = !getenv("BGGE_NO_MERGED_STEPS");

  // ---- steps: one kernel, or up to 4 consecutive single-block kernels sharing one matrix on disjoint rows
  struct StepDef {
    std::vector<int> kernels;
    bool merged;
  };
  std::vector<StepDef> sdef;
  for (int j = 0; j < nk;) {
    StepDef sd{{j}, false};
    const HostKernel& h0 = g->kernels[(size_t)j];
    if (allow_merge && h0.blocks.size() == 1) {
      const HostBlock& b0 = h0.blocks[0];
      while ((int)sd.kernels.size() < 4) {
        const int jn = j + (int)sd.kernels.size();
        if (jn >= nk || g->kernels[(size_t)jn].blocks.size() != 1) break;
        const HostBlock& bn = g->kernels[(size_t)jn].blocks[0];
        if (bn.dU != b0.dU || bn.nr != b0.nr || bn.stream_rows() != b0.stream_rows() || (bn.vrows > 0) != (b0.vrows > 0)) break;
        bool disjoint = true;
        for (int kk : sd.kernels) {
          const HostBlock& bk = g->kernels[(size_t)kk].blocks[0];
          if (bn.pos0 < bk.pos0 + bk.rows && bk.pos0 < bn.pos0 + bn.rows) disjoint = false;
        }
        if (!disjoint) break;
        sd.kernels.push_back(jn);
      }
      sd.merged = sd.kernels.size() > 1;
    }
    j += (int)sd.kernels.size();
    sdef.push_back(sd);
  }
  const int nsteps = (int)sdef.size();

  // ---- groups (blocks streaming one matrix, up to 4 members) and, per step and row, where the row's
  // virtual row sits in a part row, its weight and the kernel that owns it
  std::vector<RGroup> groups;
  std::vector<RMember> members;
  std::vector<int> sg0;
  std::vector<std::vector<int>> row_poff((size_t)nsteps, std::vector<int>((size_t)n, -1));
  std::vector<std::vector<int>> row_k((size_t)nsteps, std::vector<int>((size_t)n, -1));
  std::vector<std::vector<double>> row_isq((size_t)nsteps, std::vector<double>((size_t)n, 0.0));
  std::vector<long long> sstride((size_t)nsteps, 0);
  int max_t = 0;
  long long pstride = 0;
  for (int si = 0; si < nsteps; ++si) {
    const StepDef& sd = sdef[(size_t)si];
    sg0.push_back((int)groups.size());
    struct Ref {
      int k, b;
    };
    std::vector<std::vector<Ref>> gmem;   // members of the groups of this step
    for (int k : sd.kernels) {
      HostKernel& hk = g->kernels[(size_t)k];
      for (size_t bi = 0; bi < hk.blocks.size(); ++bi) {
        HostBlock& bl = hk.blocks[bi];
        int gi = -1;
        if (sd.merged) {
          gi = gmem.empty() ? -1 : 0;
        } else if (bl.vrows > 0) {
          for (size_t q = 0; q < gmem.size(); ++q) {
            const HostBlock& b0 = g->kernels[(size_t)gmem[q][0].k].blocks[(size_t)gmem[q][0].b];
            if (b0.vrows > 0 && b0.dU == bl.dU && b0.vrows == bl.vrows && b0.nr == bl.nr && gmem[q].size() < 4) {
              gi = (int)q;
              break;
            }
          }
        }
        if (gi < 0) {
          RGroup gr{};
          gr.W = bl.dU;
          gr.ldw = bl.ldu;
          gr.t = (int)bl.stream_rows();
          gr.t_pad = (gr.t + 1) | 1;
          gr.nr = (int)bl.nr;
          gr.nb = 0;
          gr.mem0 = (int)members.size();   // members of a group are contiguous: reserve 4 slots
          gr.s = hk.s.p + bl.col_off;
          members.resize(members.size() + 4);
          groups.push_back(gr);
          gmem.push_back({});
          gi = (int)gmem.size() - 1;
          max_t = std::max(max_t, gr.t);
        }
        gmem[(size_t)gi].push_back(Ref{k, (int)bi});
      }
    }
    // virtual rows of the members of one group are contiguous (the kernel treats them as one vector of nb * t)
    long long poff = 0;
    for (size_t q = 0; q < gmem.size(); ++q) {
      RGroup& gr = groups[(size_t)sg0.back() + q];
      for (const Ref& rf : gmem[q]) {
        HostKernel& hk = g->kernels[(size_t)rf.k];
        HostBlock& bl = hk.blocks[(size_t)rf.b];
        RMember mb{};
        mb.pos0 = (int)bl.pos0;
        mb.rows = (int)bl.rows;
        mb.col_off = (int)bl.col_off;
        mb.kidx = rf.k;
        mb.z_off = hk.z_off;
        mb.part_off = poff;
        members[(size_t)(gr.mem0 + gr.nb)] = mb;
        ++gr.nb;
        std::vector<int> loc;
        std::vector<double> isq;
        if (bl.vrows > 0) {
          loc.resize((size_t)bl.rows);
          isq.resize((size_t)bl.rows);
          BGGE_CUDA(cudaMemcpyAsync(loc.data(), bl.d_loc, loc.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
          BGGE_CUDA(cudaMemcpyAsync(isq.data(), bl.d_isq, isq.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
          BGGE_CUDA(cudaStreamSynchronize(st));
        }
        for (int64_t i = 0; i < bl.rows; ++i) {
          const size_t row = (size_t)(bl.pos0 + i);
          row_poff[(size_t)si][row] = (int)(poff + (bl.vrows > 0 ? loc[(size_t)i] : i));
          row_isq[(size_t)si][row] = bl.vrows > 0 ? isq[(size_t)i] : 1.0;
          row_k[(size_t)si][row] = rf.k;
        }
        poff += gr.t;
      }
    }
    if (!sd.merged)   // rows no block of the kernel covers still belong to it: u_j = 0 there (R/BGGE.R:334)
      for (int64_t i = 0; i < n; ++i) row_k[(size_t)si][(size_t)i] = sd.kernels[0];
    sstride[(size_t)si] = poff;
    pstride = std::max(pstride, poff);
    if (sd.merged) rp.has_merged = true;
  }
  sg0.push_back((int)groups.size());
  if (groups.empty() || (int)groups.size() > R_MAXG || pstride >= (1ll << 30)) return OK;
  // grid size: as many CTAs as useful (>= 4 columns of the widest matrix each), at most one per SM
  int max_nr = 0;
  for (auto& gr : groups) max_nr = std::max(max_nr, gr.nr);
  const int cta_cap = std::min(sms, R_SUB * R_NPL);   // the row phase reads R_NPL partials per lane
  int ncta = std::max(1, std::min(cta_cap, max_nr / 4));
  if (const char* e = getenv("BGGE_RESIDENT_CTAS")) ncta = std::max(1, std::min(cta_cap, atoi(e)));
  const int etp = (max_t + 3) & ~1;

  // ---- per step: the rows ordered by the units (virtual rows with rows of y, then uncovered rows) of the next step
  struct HostStep {
    std::vector<REntry> entries;
    std::vector<RUnit> units;
    std::vector<double> wsum;
  };
  std::vector<HostStep> hs((size_t)nsteps);
  int max_entries = 1;
  int max_units = 1, ndraw = 0, ncolc = 0, nwsum = 0;
  for (int si = 0; si < nsteps; ++si) nwsum += (int)std::max<long long>(1, sstride[(size_t)si]);
  auto entries_of = [&](int ncta_) {   // most rows / units one CTA gets in any row phase
    int worst = 1;
    max_units = 1;
    for (int si = 0; si < nsteps; ++si) {
      const auto& un = hs[(size_t)si].units;
      const int nu = (int)un.size();
      for (int c = 0; c < ncta_; ++c) {
        const int u0 = (int)((long long)nu * c / ncta_), u1 = (int)((long long)nu * (c + 1) / ncta_);
        if (u0 < u1) worst = std::max(worst, un[(size_t)u1 - 1].e1 - un[(size_t)u0].e0);
        max_units = std::max(max_units, u1 - u0);
      }
    }
    return worst;
  };
  for (int si = 0; si < nsteps; ++si) {
    const int sn = (si + 1) % nsteps;
    HostStep& h = hs[(size_t)si];
    std::vector<std::vector<int>> by_off((size_t)sstride[(size_t)sn]);
    std::vector<int> orphans;
    for (int64_t i = 0; i < n; ++i) {
      const int po = row_poff[(size_t)sn][(size_t)i];
      if (po >= 0) by_off[(size_t)po].push_back((int)i);
      else orphans.push_back((int)i);
    }
    auto push_entry = [&](int i, double w) {
      REntry e{};
      e.row = i;
      e.poff = row_poff[(size_t)si][(size_t)i];
      e.kj = row_k[(size_t)si][(size_t)i];
      e.kn = row_poff[(size_t)sn][(size_t)i] >= 0 ? row_k[(size_t)sn][(size_t)i] : -1;
      e.isq = row_isq[(size_t)si][(size_t)i];
      e.w = w;
      h.entries.push_back(e);
    };
    for (size_t po = 0; po < by_off.size(); ++po) {
      if (by_off[po].empty()) continue;
      RUnit u{};
      u.e0 = (int)h.entries.size();
      for (int i : by_off[po]) push_entry(i, row_isq[(size_t)sn][(size_t)i]);
      u.e1 = (int)h.entries.size();
      u.eoff = (int)po;
      h.units.push_back(u);
    }
    for (int i : orphans) {
      RUnit u{};
      u.e0 = (int)h.entries.size();
      push_entry(i, 0.0);
      u.e1 = (int)h.entries.size();
      u.eoff = -1;
      h.units.push_back(u);
    }
    // wsum of THIS step: sum of the row weights per virtual row
    h.wsum.assign((size_t)std::max<long long>(1, sstride[(size_t)si]), 0.0);
    for (int64_t i = 0; i < n; ++i)
      if (row_poff[(size_t)si][(size_t)i] >= 0) h.wsum[(size_t)row_poff[(size_t)si][(size_t)i]] += row_isq[(size_t)si][(size_t)i];
  }
  // ---- shared memory: slices of every group, e~ (4 x etp), staged row values, draws, column constants, row tables
  size_t smem = 0;
  long long smem_w = 0;
  for (;;) {
    long long off = 0;
    bool ok = true;
    ndraw = ncolc = 0;
    for (auto& gr : groups) {
      gr.maxcols = (gr.nr + ncta - 1) / ncta;
      if (gr.maxcols > R_MAXCOLS) ok = false;
      gr.woff = off;
      gr.doff = ndraw;
      gr.coff = ncolc;
      ndraw += gr.maxcols * gr.nb;
      ncolc += gr.maxcols;
      off += (long long)gr.maxcols * gr.t_pad;
    }
    off = (off + 1) & ~1ll;
    smem_w = off;
    max_entries = entries_of(ncta);
    smem = ((size_t)off + 4 * (size_t)etp + 2 * (size_t)max_entries + 4 * (size_t)ndraw + 2 * (size_t)ncolc + (size_t)nwsum) *
               sizeof(double) +
           (size_t)nsteps * ((size_t)max_entries * sizeof(REntry) + (size_t)max_units * sizeof(RUnit));
    if (ok && smem <= 200 * 1024) break;
    if (ncta >= cta_cap) return OK;      // does not fit even with one CTA per SM: streaming path
    ncta = std::min(cta_cap, ncta * 2);
  }
  // inside every CTA's range the rows are processed in part-offset order (neighbouring lanes then read the same
  // 32-byte sectors of the partial sums); vpos keeps the unit-major slot the unit sums read
  for (int si = 0; si < nsteps; ++si) {
    HostStep& h = hs[(size_t)si];
    const int nu = (int)h.units.size();
    for (int c = 0; c < ncta; ++c) {
      const int u0 = (int)((long long)nu * c / ncta), u1 = (int)((long long)nu * (c + 1) / ncta);
      if (u0 >= u1) continue;
      const int eb = h.units[(size_t)u0].e0, ee = h.units[(size_t)u1 - 1].e1;
      for (int k = eb; k < ee; ++k) h.entries[(size_t)k].vpos = k - eb;
      std::stable_sort(h.entries.begin() + eb, h.entries.begin() + ee, [](const REntry& a, const REntry& b) {
        return (unsigned)a.poff < (unsigned)b.poff;   // -1 (no block) last
      });
    }
  }
  int dev = 0, coop = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  BGGE_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return OK;
  BGGE_CUDA(cudaFuncSetAttribute(resident_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  BGGE_CUDA(cudaFuncSetAttribute(resident_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int per_sm = 0;
  BGGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resident_sweep_kernel<true>, R_THREADS, smem));
  if (per_sm < 1 || ncta > per_sm * sms) return OK;

  rp.ncta = ncta;
  rp.smem = smem;
  rp.smem_w = smem_w;
  rp.pstride = pstride;
  rp.etp = etp;
  rp.max_entries = max_entries;
  rp.max_units = max_units;
  rp.ndraw = ndraw;
  rp.ncolc = ncolc;
  rp.nwsum = nwsum;
  rp.nsteps = nsteps;
  rp.ngroups = (int)groups.size();
  BGGE_CUDA(rp.bar.alloc(1));
  rp.entries.clear();
  rp.units.clear();
  rp.wsum.clear();
  rp.etil0.clear();
  rp.entries.resize((size_t)nsteps);
  rp.units.resize((size_t)nsteps);
  rp.wsum.resize((size_t)nsteps);
  rp.etil0.resize((size_t)nsteps);
  std::vector<RStep> steps((size_t)nsteps);
  for (int si = 0; si < nsteps; ++si) {
    BGGE_TRY(upload_vec(rp.entries[(size_t)si], hs[(size_t)si].entries, st));
    BGGE_TRY(upload_vec(rp.units[(size_t)si], hs[(size_t)si].units, st));
    BGGE_TRY(upload_vec(rp.wsum[(size_t)si], hs[(size_t)si].wsum, st));
    const size_t ne0 = (size_t)std::max<long long>(1, sstride[(size_t)si]);
    BGGE_CUDA(rp.etil0[(size_t)si].alloc(ne0));
    BGGE_CUDA(cudaMemsetAsync(rp.etil0[(size_t)si].p, 0, ne0 * sizeof(double), st));
    RStep& S = steps[(size_t)si];
    S.entries = rp.entries[(size_t)si].p;
    S.units = rp.units[(size_t)si].p;
    S.nunits = (int)hs[(size_t)si].units.size();
    S.g0 = sg0[(size_t)si];
    S.g1 = sg0[(size_t)si + 1];
    S.next = (si + 1) % nsteps;
    S.merged = sdef[(size_t)si].merged ? 1 : 0;
    for (int q = 0; q < 4; ++q)
      S.zk[q] = sdef[(size_t)si].kernels[std::min<size_t>((size_t)q, sdef[(size_t)si].kernels.size() - 1)];
    S.wsum = rp.wsum[(size_t)si].p;
    S.nws = (int)hs[(size_t)si].wsum.size();
    S.woff = si == 0 ? 0 : steps[(size_t)si - 1].woff + steps[(size_t)si - 1].nws;
    S.etil0 = rp.etil0[(size_t)si].p;
  }
  BGGE_TRY(upload_vec(rp.steps, steps, st));
  BGGE_TRY(upload_vec(rp.groups, groups, st));
  BGGE_TRY(upload_vec(rp.members, members, st));
  BGGE_CUDA(rp.part.alloc((size_t)ncta * pstride));
  BGGE_CUDA(rp.zqpart.alloc((size_t)2 * nk * ncta));
  BGGE_CUDA(rp.sspart.alloc((size_t)3 * ncta));
  if (g->q > 0) {   // XF'1 and, for step 0, C^-1/2 Z' XF per virtual row
    const int q = g->q;
    std::vector<double> x1((size_t)q, 0.0), mx0((size_t)q * (size_t)std::max<long long>(1, sstride[0]), 0.0);
    for (int k = 0; k < q; ++k) {
      long double a1 = 0.0L;
      for (int64_t i = 0; i < n; ++i) {
        const double x = g->xf_host[(size_t)(i + (int64_t)k * n)];
        a1 += x;
        const int po = row_poff[0][(size_t)i];
        if (po >= 0) mx0[(size_t)k * (size_t)sstride[0] + (size_t)po] += row_isq[0][(size_t)i] * x;
      }
      x1[(size_t)k] = (double)a1;
    }
    BGGE_TRY(upload_vec(rp.x1, x1, st));
    BGGE_TRY(upload_vec(rp.mx0, mx0, st));
    BGGE_CUDA(rp.xrpart.alloc((size_t)q * ncta));
    rp.ps0 = sstride[0];
  }
  BGGE_CUDA(cudaStreamSynchronize(st));
  rp.enabled = true;
  if (getenv("BGGE_DEBUG"))
    fprintf(stderr, "[bgge] resident plan: ncta=%d smem=%zu steps=%d (merged: %s) groups=%zu pstride=%lld max_t=%d max_entries=%d\n",
            ncta, smem, nsteps, rp.has_merged ? "yes" : "no", groups.size(), (long long)pstride, max_t, max_entries);
  return OK;
}

int launch_resident(Gibbs* g, int64_t iters) {
  ResidentPlan& rp = g->resident;
  RParams p{};
  p.groups = rp.groups.p;
  p.members = rp.members.p;
  p.steps = rp.steps.p;
  p.nk = g->nk;
  p.nsteps = rp.nsteps;
  for (int j = 0; j < g->nk; ++j) p.nr_k[j] = g->kernels[(size_t)j].nr_total;
  p.ncta = rp.ncta;
  p.n = g->n;
  p.r = g->r.p;
  p.u = g->u.p;
  p.y = g->y.p;
  p.umean = g->umean.p;
  p.um2 = g->um2.p;
  p.sc = g->scal.p;
  p.part = rp.part.p;
  p.pstride = rp.pstride;
  p.zqpart = rp.zqpart.p;
  p.sspart = rp.sspart.p;
  p.na_idx = g->na_idx.p;
  p.n_na = g->n_na;
  p.chain_mu = g->chain_mu.p;
  p.chain_varE = g->chain_varE.p;
  p.chain_varU = g->chain_varU.p;
  p.thin = g->thin;
  p.burn = g->burn;
  p.ncum = g->ncum;
  p.iters = iters;
  p.etp = rp.etp;
  p.smem_w = rp.smem_w;
  p.max_entries = rp.max_entries;
  p.max_units = rp.max_units;
  p.ndraw = rp.ndraw;
  p.ncolc = rp.ncolc;
  p.nwsum = rp.nwsum;
  p.ngroups = rp.ngroups;
  p.q = g->q;
  p.xf = g->xf.p;
  p.txx = g->txx.p;
  p.lchol = g->lchol.p;
  p.x1 = rp.x1.p;
  p.mx0 = rp.mx0.p;
  p.ps0 = rp.ps0;
  p.bet = g->bet.p;
  p.chain_bet = g->chain_bet.p;
  p.xrpart = rp.xrpart.p;
  p.bar = rp.bar.p;
  BGGE_CUDA(cudaMemsetAsync(rp.bar.p, 0, sizeof(unsigned long long), g->stream));
  DrawCfg dc = g->draw;
  DevBuf<unsigned long long> timing;
  if (getenv("BGGE_RESIDENT_TIMING")) {
    BGGE_CUDA(timing.alloc(16));
    BGGE_CUDA(cudaMemsetAsync(timing.p, 0, 16 * sizeof(unsigned long long), g->stream));
    p.timing = timing.p;
  }
  void* args[] = {&p, &dc};
  const cudaError_t le = cudaLaunchCooperativeKernel(p.timing ? (void*)resident_sweep_kernel<true> : (void*)resident_sweep_kernel<false>,
                                                     dim3((unsigned)rp.ncta), dim3(R_THREADS), args, rp.smem, g->stream);
  if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) {
    // the grid cannot be co-resident right now (GPU shared with other work): the streaming kernels take over
    (void)cudaGetLastError();
    rp.enabled = false;
    return ERR_RETRY_STREAMING;
  }
  BGGE_CUDA(le);
  if (p.timing) {
    unsigned long long t[16];
    BGGE_CUDA(cudaMemcpyAsync(t, timing.p, sizeof(t), cudaMemcpyDeviceToHost, g->stream));
    BGGE_CUDA(cudaStreamSynchronize(g->stream));
    const double k = 1e-3 / (double)std::max<int64_t>(1, iters);
    fprintf(stderr,
            "[bgge] resident us/iter (CTA 0): P1 %.2f  P2 %.2f  draw %.2f  P3 %.2f  zq %.2f | arriveA %.2f  slack %.2f  waitA %.2f | rows: "
            "issue %.2f  arrive %.2f  update %.2f  sync %.2f  units %.2f  sums %.2f | barrierB %.2f | scalars %.2f\n",
            t[5] * k, t[6] * k, t[7] * k, t[8] * k, t[0] * k, t[9] * k, t[10] * k, t[1] * k, t[14] * k, t[15] * k, t[11] * k, t[12] * k, t[13] * k, t[2] * k,
            t[3] * k, t[4] * k);
  }
  return OK;
}

}  // namespace bgge

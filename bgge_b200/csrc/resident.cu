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
constexpr int R_SUB = 8;                      // lanes that share one row in P5
constexpr int R_MAXCOLS = 64;                 // columns of one matrix per CTA (bsm capacity)
constexpr int R_MAXV = MAXK + 2;

struct RParams {
  const RGroup* groups;
  const RMember* members;
  const RStep* steps;
  int nk, ncta;
  long long n;
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
  int max_entries;
};

struct RShared {
  double red[R_MAXV][R_WARPS];
  double bsm[4 * R_MAXCOLS];
  double chi[MAXK + 1];
  double sigb[MAXK];
  double Sc[MAXK];
  double zmu;
};

// P5 of step j (update = true), or only the gather of the next step's etil0 (update = false: prologue of a
// launch and after the imputation of missing phenotypes).
__device__ __forceinline__ void rows_phase(const RParams& p, const RStep& S, int j, bool update, bool last, double shift,
                                           bool keep, double kcount, double* vals, RShared& sh) {
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int sub = tid / R_SUB, sl = tid % R_SUB;
  const int u0 = (int)((long long)S.nunits * cta / p.ncta), u1 = (int)((long long)S.nunits * (cta + 1) / p.ncta);
  const int eb = u0 < u1 ? S.units[u0].e0 : 0, ee = u0 < u1 ? S.units[u1 - 1].e1 : 0;
  const int jn = S.next;
  double* uj = p.u + (size_t)j * p.n;
  const double* ujn_p = p.u + (size_t)jn * p.n;
  double sse = 0.0, sr = 0.0;
  for (int k0 = eb; k0 < ee; k0 += R_THREADS / R_SUB) {
    const int k = k0 + sub;
    const bool valid = k < ee;
    REntry E{};
    if (valid) E = S.entries[k];
    double rold = 0.0, uold = 0.0, ujn = 0.0, mm = 0.0, m2 = 0.0;
    if (valid && sl == 0) {
      rold = p.r[E.row];
      if (update) uold = uj[E.row];
      if (!update || jn != j) ujn = ujn_p[E.row];
      if (update && keep) {
        mm = p.umean[(size_t)j * p.n + E.row];
        m2 = p.um2[(size_t)j * p.n + E.row];
      }
    }
    double v;
    if (update) {
      double un = 0.0;
      if (valid && E.poff >= 0) {
        const double* pp = p.part + E.poff;
        for (int c = sl; c < p.ncta; c += R_SUB) un += pp[(size_t)c * p.pstride];
      }
      un += __shfl_xor_sync(0xffffffffu, un, 4);
      un += __shfl_xor_sync(0xffffffffu, un, 2);
      un += __shfl_xor_sync(0xffffffffu, un, 1);
      un *= E.isq;
      const double rn = (rold + uold) - un;
      if (valid && sl == 0) {
        p.r[E.row] = rn;
        uj[E.row] = un;
        if (keep) {
          const double dlt = un - mm;
          const double mn = mm + dlt / kcount;
          p.umean[(size_t)j * p.n + E.row] = mn;
          p.um2[(size_t)j * p.n + E.row] = m2 + dlt * (un - mn);
        }
        if (last) {
          const double e = rn - shift;
          sse = fma(e, e, sse);
          sr += rn;
        }
      }
      v = E.w * (rn + (jn == j ? un : ujn));
    } else {
      v = E.w * (rold + ujn);
    }
    if (valid && sl == 0) vals[k - eb] = v;
  }
  __syncthreads();
  double* etil0 = p.steps[jn].etil0;
  for (int un = u0 + tid; un < u1; un += R_THREADS) {
    const RUnit U = S.units[un];
    if (U.eoff < 0) continue;
    double acc = 0.0;
    for (int k = U.e0; k < U.e1; ++k) acc += vals[k - eb];
    etil0[U.eoff] = acc;
  }
  if (last) {   // per-CTA pieces of sum e^2 and sum r (fixed order)
    sse = warp_sum(sse);
    sr = warp_sum(sr);
    if ((tid & 31) == 0) {
      sh.red[0][tid >> 5] = sse;
      sh.red[1][tid >> 5] = sr;
    }
    __syncthreads();
    if (tid < 2) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < R_WARPS; ++w) t += sh.red[tid][w];
      p.sspart[(size_t)tid * p.ncta + cta] = t;
    }
  }
}

__global__ void __launch_bounds__(R_THREADS, 1)
resident_sweep_kernel(const RParams p, DrawCfg dc) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char rsm[];
  __shared__ RShared sh;
  double* wsm = reinterpret_cast<double*>(rsm);                 // per group: slice [maxcols][t_pad]
  double* etil = wsm + p.smem_w;                                // [4][etp]
  double* vals = etil + 4 * p.etp;                              // [max_entries]
  const int cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ngroups = p.steps[p.nk - 1].g1;

  // ---- load my column slices of every basis matrix once
  for (int g = 0; g < ngroups; ++g) {
    const RGroup& gr = p.groups[g];
    const int c0 = (int)((long long)gr.nr * cta / p.ncta), c1 = (int)((long long)gr.nr * (cta + 1) / p.ncta);
    double* dst = wsm + gr.woff;
    for (int c = c0; c < c1; ++c)
      for (int a = tid; a < gr.t; a += R_THREADS) dst[(size_t)(c - c0) * gr.t_pad + a] = gr.W[(long long)c * gr.ldw + a];
  }
  // chain scalars: every CTA holds (and updates identically) its own copy
  Scal* sc = p.sc;
  double mu = sc->mu, shift = sc->shift, sigsq = sc->sigsq, tau = sc->tau, mu_done = sc->mu_done;
  const double mu0 = sc->mu0, Sce = sc->Sce;
  long long it = sc->it;
  for (int j = tid; j < p.nk; j += R_THREADS) {
    sh.sigb[j] = sc->sigb[j];
    sh.Sc[j] = sc->Sc[j];
  }
  // first kernel's etil0 from the state in memory
  rows_phase(p, p.steps[p.nk - 1], p.nk - 1, false, false, 0.0, false, 1.0, vals, sh);
  grid.sync();

  for (long long step = 0; step < p.iters; ++step, ++it) {
    const bool keep = (it % p.thin == 0) && (it > p.burn);
    const double kcount = keep ? (double)(it / p.thin - p.burn / p.thin) : 1.0;
    // draws that do not depend on the data: made up front by the warps P2 uses last
    if (lane == 0) {
      for (int d = R_WARPS - 1 - warp; d <= p.nk + 1; d += R_WARPS) {
        if (d < p.nk)
          sh.chi[d] = draw_chisq(dc, it, STREAM_CHI_K + (uint32_t)d, (double)p.steps[d].nr_total + NU, (it - 1) * dc.c_per_it + d);
        else if (d == p.nk)
          sh.chi[p.nk] = draw_chisq(dc, it, STREAM_CHI_E, (double)p.n + NU, (it - 1) * dc.c_per_it + p.nk);
        else
          sh.zmu = draw_normal(dc, it + 1, STREAM_MU, 0ull, dc.z_init + it * dc.z_per_it);
      }
    }
    for (int j = 0; j < p.nk; ++j) {
      const RStep& S = p.steps[j];
      double zq = 0.0;
      const double lam = sh.sigb[j];
      for (int g = S.g0; g < S.g1; ++g) {
        const RGroup& gr = p.groups[g];
        const int c0 = (int)((long long)gr.nr * cta / p.ncta), c1 = (int)((long long)gr.nr * (cta + 1) / p.ncta);
        const int ncol = c1 - c0;
        const double* W = wsm + gr.woff;
        // P1: residual in genotype space for every member block
        for (int m = 0; m < gr.nb; ++m) {
          const long long po = p.members[gr.mem0 + m].part_off;
          for (int a = tid; a < gr.t; a += R_THREADS) etil[m * p.etp + a] = S.etil0[po + a] - shift * S.wsum[po + a];
        }
        __syncthreads();
        // P2: one warp per (column, member): dot, draw
        for (int w = warp; w < ncol * gr.nb; w += R_WARPS) {
          const int cc = w / gr.nb, m = w % gr.nb;
          const double* wc = W + (size_t)cc * gr.t_pad;
          const double* em = etil + m * p.etp;
          double d0 = 0.0, d1 = 0.0;
          int a = lane;
          for (; a + 32 < gr.t; a += 64) {
            d0 = fma(wc[a], em[a], d0);
            d1 = fma(wc[a + 32], em[a + 32], d1);
          }
          if (a < gr.t) d0 = fma(wc[a], em[a], d0);
          const double d = warp_sum(d0 + d1);
          if (lane == 0) {
            const long long col = (long long)p.members[gr.mem0 + m].col_off + c0 + cc;
            const double sv = gr.s[c0 + cc];
            const double vari = sv * lam / (1.0 + sv * lam * tau);          // R/BGGE.R:302
            const double media = tau * vari * d;                            // R/BGGE.R:303
            const double z = draw_normal(dc, it, (uint32_t)j, (unsigned long long)col,
                                         dc.z_init + (it - 1) * dc.z_per_it + gr.z_off + col);
            const double bv = media + sqrt(vari) * z;
            sh.bsm[m * R_MAXCOLS + cc] = bv;
            zq += bv * bv * (1.0 / sv);                                     // R/BGGE.R:96
          }
        }
        __syncthreads();
        // P3: my columns' contribution to W b, per member
        for (int m = 0; m < gr.nb; ++m) {
          double* out = p.part + (size_t)cta * p.pstride + p.members[gr.mem0 + m].part_off;
          const double* bm = sh.bsm + m * R_MAXCOLS;
          for (int a = tid; a < gr.t; a += R_THREADS) {
            double acc = 0.0;
            for (int cc = 0; cc < ncol; ++cc) acc = fma(W[(size_t)cc * gr.t_pad + a], bm[cc], acc);
            out[a] = acc;
          }
        }
        __syncthreads();
      }
      // sum of b^2/s of my columns (lanes 0 hold the pieces)
      if (lane == 0) sh.red[0][warp] = zq;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < R_WARPS; ++w) t += sh.red[0][w];
        p.zqpart[((size_t)(it & 1) * p.nk + j) * p.ncta + cta] = t;
      }
      grid.sync();
      rows_phase(p, S, j, true, j == p.nk - 1, shift, keep, kcount, vals, sh);
      grid.sync();
    }
    // ---- scalars, identically in every CTA: sum the per-CTA pieces (R/BGGE.R:360-367)
    const int nv = p.nk + 2;
    for (int v = 0; v < nv; ++v) {
      double x = 0.0;
      if (tid < p.ncta)
        x = v < p.nk ? p.zqpart[((size_t)(it & 1) * p.nk + v) * p.ncta + tid] : p.sspart[(size_t)(v - p.nk) * p.ncta + tid];
      x = warp_sum(x);
      if (lane == 0) sh.red[v][warp] = x;
    }
    __syncthreads();
    double sse = 0.0, sr = 0.0;
#pragma unroll
    for (int w = 0; w < R_WARPS; ++w) {
      sse += sh.red[p.nk][w];
      sr += sh.red[p.nk + 1][w];
    }
    if (tid < p.nk) {
      double zq = 0.0;
#pragma unroll
      for (int w = 0; w < R_WARPS; ++w) zq += sh.red[tid][w];
      sh.sigb[tid] = (zq + NU * sh.Sc[tid]) / sh.chi[tid];
    }
    sigsq = (sse + Sce) / sh.chi[p.nk];
    tau = 1.0 / sigsq;
    __syncthreads();
    if (p.n_na > 0) {   // missing phenotypes (R/BGGE.R:370-381): rows split over the CTAs
      const double sd = sqrt(sigsq);
      double dsum = 0.0;
      for (long long k = (long long)cta * R_THREADS + tid; k < p.n_na; k += (long long)p.ncta * R_THREADS) {
        const long long i = p.na_idx[k];
        double uhat = 0.0;
        for (int j = 0; j < p.nk; ++j) uhat = (j == 0) ? p.u[i] : uhat + p.u[(size_t)j * p.n + i];
        const double z = draw_normal(dc, it, STREAM_NA, (unsigned long long)k, dc.z_init + (it - 1) * dc.z_per_it + dc.z_off_na + k);
        const double yv = ((0.0 + mu) + uhat) + sd * z;
        p.y[i] = yv;
        const double rn = (((yv - uhat) - 0.0) - mu) + shift;
        dsum += rn - p.r[i];
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
      grid.sync();
      rows_phase(p, p.steps[p.nk - 1], p.nk - 1, false, false, 0.0, false, 1.0, vals, sh);
      grid.sync();
      double x = tid < p.ncta ? p.sspart[(size_t)2 * p.ncta + tid] : 0.0;
      x = warp_sum(x);
      if (lane == 0) sh.red[0][warp] = x;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < R_WARPS; ++w) sr += sh.red[0][w];
      __syncthreads();
    }
    if (cta == 0 && tid == 0 && it % p.thin == 0) {   // thinned chain (R/BGGE.R:384-395)
      const long long slot = it / p.thin - 1;
      if (slot < p.ncum) {
        p.chain_varE[slot] = sigsq;
        p.chain_mu[slot] = mu;
        for (int j = 0; j < p.nk; ++j) p.chain_varU[(long long)j * p.ncum + slot] = sh.sigb[j];
      }
    }
    // next intercept (R/BGGE.R:278-281): mean(temp + mu) = mean(r) + mu0
    mu_done = mu;
    mu = (sr / (double)p.n + mu0) + sqrt(sigsq / (double)p.n) * sh.zmu;
    shift = mu - mu0;
    __syncthreads();   // sh.zmu / sh.chi are redrawn at the top of the next iteration
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
  if (g->q > 0 || g->mode != 1 || getenv("BGGE_NO_RESIDENT")) return OK;
  if (g->n > 32768) return OK;
  cudaStream_t st = g->stream;
  const int sms = sm_count();
  const int64_t n = g->n;
  const int nk = g->nk;
  // groups: blocks sharing one matrix (same rule as the fused plan), up to 4 members
  std::vector<RGroup> groups;
  std::vector<RMember> members;
  std::vector<int> kg0;
  // per kernel, per row: offset into a part row (-1: no block covers the row) and the row weight
  std::vector<std::vector<int>> row_poff((size_t)nk, std::vector<int>((size_t)n, -1));
  std::vector<std::vector<double>> row_isq((size_t)nk, std::vector<double>((size_t)n, 0.0));
  std::vector<long long> kstride((size_t)nk, 0);
  int max_t = 0;
  long long pstride = 0;
  for (int j = 0; j < nk; ++j) {
    HostKernel& hk = g->kernels[(size_t)j];
    kg0.push_back((int)groups.size());
    long long poff = 0;
    std::vector<int> first_of_group;   // block index of the first member of each group of this kernel
    for (size_t bi = 0; bi < hk.blocks.size(); ++bi) {
      HostBlock& bl = hk.blocks[bi];
      int gi = -1;
      if (bl.vrows > 0)
        for (size_t k = 0; k < first_of_group.size(); ++k) {
          const HostBlock& b0 = hk.blocks[(size_t)first_of_group[k]];
          const RGroup& gr = groups[(size_t)kg0.back() + k];
          if (b0.vrows > 0 && b0.dU == bl.dU && b0.vrows == bl.vrows && b0.nr == bl.nr && gr.nb < 4) {
            gi = (int)k;
            break;
          }
        }
      if (gi < 0) {
        RGroup gr{};
        gr.W = bl.dU;
        gr.ldw = bl.ldu;
        gr.t = (int)bl.stream_rows();
        gr.t_pad = (gr.t + 1) | 1;       // odd stride: P3 rows and P2 columns both conflict-free enough
        gr.nr = (int)bl.nr;
        gr.nb = 0;
        gr.mem0 = (int)members.size();   // members of a group are contiguous: reserve 4 slots
        gr.s = hk.s.p + bl.col_off;
        gr.z_off = hk.z_off;
        members.resize(members.size() + 4);
        groups.push_back(gr);
        first_of_group.push_back((int)bi);
        gi = (int)first_of_group.size() - 1;
      }
      RGroup& gr = groups[(size_t)kg0.back() + gi];
      RMember mb{};
      mb.pos0 = (int)bl.pos0;
      mb.rows = (int)bl.rows;
      mb.col_off = (int)bl.col_off;
      mb.part_off = poff;
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
        row_poff[(size_t)j][(size_t)(bl.pos0 + i)] = (int)(poff + (bl.vrows > 0 ? loc[(size_t)i] : i));
        row_isq[(size_t)j][(size_t)(bl.pos0 + i)] = bl.vrows > 0 ? isq[(size_t)i] : 1.0;
      }
      poff += gr.t;
      members[(size_t)(gr.mem0 + gr.nb)] = mb;
      ++gr.nb;
      max_t = std::max(max_t, gr.t);
    }
    kstride[(size_t)j] = poff;
    pstride = std::max(pstride, poff);
  }
  kg0.push_back((int)groups.size());
  if (groups.empty() || pstride >= (1ll << 30)) return OK;
  // grid size: as many CTAs as useful (>= 4 columns of the widest matrix each), at most one per SM
  int max_nr = 0;
  for (auto& gr : groups) max_nr = std::max(max_nr, gr.nr);
  int ncta = std::max(1, std::min(sms, max_nr / 4));
  if (const char* e = getenv("BGGE_RESIDENT_CTAS")) ncta = std::max(1, std::min(sms, atoi(e)));
  if (ncta > R_THREADS) return OK;
  const int etp = (max_t + 3) & ~1;

  // units of every kernel (virtual rows that have rows of y, then rows no block covers) and, per step j, the rows
  // ordered by the units of the next kernel
  struct HostStep {
    std::vector<REntry> entries;
    std::vector<RUnit> units;
    std::vector<double> wsum;
  };
  std::vector<HostStep> hs((size_t)nk);
  int max_entries = 1;
  auto entries_of = [&](int ncta_) {
    int worst = 1;
    for (int j = 0; j < nk; ++j) {
      const auto& un = hs[(size_t)j].units;
      const int nu = (int)un.size();
      for (int c = 0; c < ncta_; ++c) {
        const int u0 = (int)((long long)nu * c / ncta_), u1 = (int)((long long)nu * (c + 1) / ncta_);
        if (u0 < u1) worst = std::max(worst, un[(size_t)u1 - 1].e1 - un[(size_t)u0].e0);
      }
    }
    return worst;
  };
  for (int j = 0; j < nk; ++j) {
    const int jn = (j + 1) % nk;
    HostStep& h = hs[(size_t)j];
    // rows of the next kernel grouped by part offset (= virtual row), ascending rows inside a unit
    std::vector<std::vector<int>> by_off((size_t)kstride[(size_t)jn]);
    std::vector<int> orphans;
    for (int64_t i = 0; i < n; ++i) {
      const int po = row_poff[(size_t)jn][(size_t)i];
      if (po >= 0) by_off[(size_t)po].push_back((int)i);
      else orphans.push_back((int)i);
    }
    auto push_entry = [&](int i, double w) {
      REntry e{};
      e.row = i;
      e.poff = row_poff[(size_t)j][(size_t)i];
      e.isq = row_isq[(size_t)j][(size_t)i];
      e.w = w;
      h.entries.push_back(e);
    };
    for (size_t po = 0; po < by_off.size(); ++po) {
      if (by_off[po].empty()) continue;
      RUnit u{};
      u.e0 = (int)h.entries.size();
      for (int i : by_off[po]) push_entry(i, row_isq[(size_t)jn][(size_t)i]);
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
    // wsum of THIS kernel j: sum of the row weights per virtual row
    h.wsum.assign((size_t)std::max<long long>(1, kstride[(size_t)j]), 0.0);
    for (int64_t i = 0; i < n; ++i)
      if (row_poff[(size_t)j][(size_t)i] >= 0) h.wsum[(size_t)row_poff[(size_t)j][(size_t)i]] += row_isq[(size_t)j][(size_t)i];
  }
  // shared memory: slices of every group + e~ (4 x etp) + the staged row values
  size_t smem = 0;
  long long smem_w = 0;
  for (;;) {
    long long off = 0;
    bool ok = true;
    for (auto& gr : groups) {
      gr.maxcols = (gr.nr + ncta - 1) / ncta;
      if (gr.maxcols > R_MAXCOLS) ok = false;
      gr.woff = off;
      off += (long long)gr.maxcols * gr.t_pad;
    }
    off = (off + 1) & ~1ll;
    smem_w = off;
    max_entries = entries_of(ncta);
    smem = ((size_t)off + 4 * (size_t)etp + (size_t)max_entries + 2) * sizeof(double);
    if (ok && smem <= 200 * 1024) break;
    if (ncta >= sms) return OK;          // does not fit even with one CTA per SM: streaming path
    ncta = std::min(sms, ncta * 2);
  }
  int dev = 0, coop = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  BGGE_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return OK;
  BGGE_CUDA(cudaFuncSetAttribute(resident_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int per_sm = 0;
  BGGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resident_sweep_kernel, R_THREADS, smem));
  if (per_sm < 1 || ncta > per_sm * sms) return OK;

  rp.ncta = ncta;
  rp.smem = smem;
  rp.smem_w = smem_w;
  rp.pstride = pstride;
  rp.etp = etp;
  rp.max_entries = max_entries;
  rp.entries.resize((size_t)nk);
  rp.units.resize((size_t)nk);
  rp.wsum.resize((size_t)nk);
  rp.etil0.resize((size_t)nk);
  std::vector<RStep> steps((size_t)nk);
  for (int j = 0; j < nk; ++j) {
    BGGE_TRY(upload_vec(rp.entries[(size_t)j], hs[(size_t)j].entries, st));
    BGGE_TRY(upload_vec(rp.units[(size_t)j], hs[(size_t)j].units, st));
    BGGE_TRY(upload_vec(rp.wsum[(size_t)j], hs[(size_t)j].wsum, st));
    BGGE_CUDA(rp.etil0[(size_t)j].alloc((size_t)std::max<long long>(1, kstride[(size_t)j])));
    BGGE_CUDA(cudaMemsetAsync(rp.etil0[(size_t)j].p, 0, (size_t)std::max<long long>(1, kstride[(size_t)j]) * sizeof(double), st));
    RStep& S = steps[(size_t)j];
    S.entries = rp.entries[(size_t)j].p;
    S.units = rp.units[(size_t)j].p;
    S.nunits = (int)hs[(size_t)j].units.size();
    S.g0 = kg0[(size_t)j];
    S.g1 = kg0[(size_t)j + 1];
    S.nr_total = g->kernels[(size_t)j].nr_total;
    S.wsum = rp.wsum[(size_t)j].p;
    S.etil0 = rp.etil0[(size_t)j].p;
    S.next = (j + 1) % nk;
  }
  BGGE_TRY(upload_vec(rp.steps, steps, st));
  BGGE_TRY(upload_vec(rp.groups, groups, st));
  BGGE_TRY(upload_vec(rp.members, members, st));
  BGGE_CUDA(rp.part.alloc((size_t)ncta * pstride));
  BGGE_CUDA(rp.zqpart.alloc((size_t)2 * nk * ncta));
  BGGE_CUDA(rp.sspart.alloc((size_t)3 * ncta));
  BGGE_CUDA(cudaStreamSynchronize(st));
  rp.enabled = true;
  if (getenv("BGGE_DEBUG"))
    fprintf(stderr, "[bgge] resident plan: ncta=%d smem=%zu groups=%zu pstride=%lld max_t=%d max_entries=%d\n", ncta, smem,
            groups.size(), (long long)pstride, max_t, max_entries);
  return OK;
}

int launch_resident(Gibbs* g, int64_t iters) {
  ResidentPlan& rp = g->resident;
  RParams p{};
  p.groups = rp.groups.p;
  p.members = rp.members.p;
  p.steps = rp.steps.p;
  p.nk = g->nk;
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
  DrawCfg dc = g->draw;
  void* args[] = {&p, &dc};
  BGGE_CUDA(cudaLaunchCooperativeKernel((void*)resident_sweep_kernel, dim3((unsigned)rp.ncta), dim3(R_THREADS), args, rp.smem,
                                        g->stream));
  return OK;
}

}  // namespace bgge

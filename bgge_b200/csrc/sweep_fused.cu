// Single-pass kernel step of the Gibbs sweep: for one random effect j it computes
//     d = U'(e + u_j),   b ~ N(tau v d, v),   u_j = U b                       (R/BGGE.R:297-306 / 314-338)
// while reading U from HBM ONCE.  (proj_kernel + back_kernel in gibbs.cu read it twice.)
//
// A thread-block cluster of CS CTAs owns a panel of eigen-columns of one block; CTA `rank` owns the
// row slice [rank*Rs, rank*Rs + Rs).  Per group of G columns:
//   1. the TMA unit (cp.async.bulk + mbarrier, SASS UBLKCP) lands the CTA's G column slices in a
//      shared-memory ring (NS groups deep, issued by thread 0);
//   2. every thread moves its rows of the group to registers (the ring slot is refilled at once),
//      forms its part of the G dot products against the register-resident residual slice, block-reduces;
//   3. the G partial dots go to every CTA of the cluster with st.async over distributed shared memory,
//      completing a transaction barrier on the receiver (no cluster-wide barrier in the loop);
//   4. one group later (software pipelined, so the exchange latency hides behind step 2 of the next
//      group) each CTA sums the CS partials in rank order, draws the same b (counter-based Philox or
//      the injected tape -> identical in all CTAs) and applies u_slice += U_slice b from registers.
// Partial u per cluster goes to upart[cluster][n]; fused_finalize_kernel sums the clusters in fixed
// order (deterministic) and does the residual / Welford update of back_kernel.
#include "gibbs.cuh"
#include "kernels.h"
#include "rng.cuh"
#include "sweep_mrhs.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

namespace bgge {
namespace {

constexpr int F_ROLE = 256;                      // threads per compute role (dot warps / axpy warps)
constexpr int F_THREADS = 2 * F_ROLE + 32;       // + one producer warp (one elected lane issues TMA)
constexpr int F_ROWS_PER_STEP = 2 * F_ROLE;      // each role thread owns row pairs (16-byte shared loads)
constexpr int F_MAX_NS = 10;                     // data ring depth limit
constexpr int NPX = 32;                          // max partial-exchange ring depth; runtime depth is 2 ns + 4
constexpr int F_MAXCS = 8;                       // largest cluster (portable limit; 16 brought 7 clusters and no gain)

__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}
// waiting warps share issue slots with working ones: back off between polls
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  int polls = 0;
  while (!mbar_try_wait(bar, parity))
    if (++polls > 8) __nanosleep(32);   // short waits (small, latency-bound problems) never sleep
}
__device__ __forceinline__ void role_barrier(int id) {   // named barrier over one 256-thread role
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(F_ROLE) : "memory");
}

// Sums P values (P a power of two <= 32) over the 32 lanes with P - 1 + (5 - log2 P) 64-bit shuffles instead of
// 5 P: at every step a lane keeps half of its values and hands the other half to its partner.  On return
// a[0] of the lanes with lane % (32/P) == 0 ... (every lane) holds the total of value index
// idx = (lane * P) >> 5  (the top log2 P lane bits); all lanes of one idx agree.
template <int P>
__device__ __forceinline__ double warp_multi_reduce(double (&a)[P], int lane, int& idx) {
  int n = P;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    if (n > 1) {
      const int half = n >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        if (i < half) {
          const double send = upper ? a[i] : a[i + half];
          const double keep = upper ? a[i + half] : a[i];
          a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      n = half;
    } else {
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], off);
    }
  }
  idx = (lane * P) >> 5;
  return a[0];
}

// Warp-specialised CTA:
//   producer (warp 16, one lane)  waits empty[slot], issues the bulk copies of the next group
//   dot warps 0..7                wait full[slot]; partial dots against the register-resident residual
//                                 slice; role-local reduction; lanes 0..CS-1 of warp 0 push the partials
//                                 to every CTA of the cluster (st.async + remote mbarrier complete_tx).
//                                 They never wait for the exchange and run ahead as far as the ring allows.
//   axpy warps 8..15              wait pbar[group] (all CS partials), every thread forms d and b itself
//                                 (partials summed in rank order; draw terms precomputed per 256-column
//                                 batch), u_slice += U_slice b from the same shared-memory copy, then
//                                 release the slot (empty[slot]).
// Protocol note: a remote CTA can send group h only after its axpy finished h - ns (slot reuse), which
// needs my partials of h - ns; my axpy can lag my dots by ns groups, so up to 2 ns + 1 partial entries
// are live; the exchange ring holds 2 ns + 4 entries (<= NPX).
template <int RPT, int G, int NB>
__global__ void __launch_bounds__(F_THREADS, 1)
sweep_fused_kernel(const FPanel* __restrict__ panels, const int2* __restrict__ cluster_panels,
                   const double* __restrict__ r, const double* __restrict__ u_all,
                   double* __restrict__ upart, long long n,
                   const double* __restrict__ ve, double* __restrict__ vupart, long long vtotal,
                   const Scal* __restrict__ sc, int merged, DrawCfg dc, int ns, int npx) {
  pdl_trigger();
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const uint32_t cs = cluster.num_blocks();
  const int cid = blockIdx.x / cs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ZB = G * (F_ROLE / G);           // columns per draw batch

  extern __shared__ __align__(128) unsigned char fsm[];
  // carve: ring[ns] | parts[NPX][F_MAXCS][NB*G] | red[2][8][NB*G] | bsm[2][NB*G] | tvb,isb[2][ZB] | zsb[NB][2][ZB] |
  //        full[F_MAX_NS] | empty[F_MAX_NS] | pbar[NPX]
  constexpr int NG = NB * G;
  const int2 cp = cluster_panels[cid];
  int max_rs = 0;
  for (int pi = 0; pi < cp.y; ++pi) max_rs = max(max_rs, panels[cp.x + pi].rs);
  const size_t slot_doubles = (size_t)G * max_rs;
  double* ring = reinterpret_cast<double*>(fsm);
  double* parts = ring + (size_t)ns * slot_doubles;
  double* red = parts + (size_t)npx * F_MAXCS * NG;
  double* bsm = red + 2 * 8 * NG;                // NB > 1: coefficients of the current group (double buffered)
  const int tvm = merged ? NB : 1;               // merged unit: every member has its own sigb, hence its own tau * v
  double* tvb = bsm + 2 * NG;                    // tau * v       per column of a draw batch, [tvm][2][ZB]
  double* isb = tvb + (size_t)tvm * 2 * ZB;      // 1 / s         (members share the matrix, hence s)
  double* zsb = isb + 2 * ZB;                    // sqrt(v) * z,  [NB][2][ZB]
  uint64_t* full = reinterpret_cast<uint64_t*>(zsb + (size_t)NB * 2 * ZB);
  uint64_t* empty = full + F_MAX_NS;
  uint64_t* pbar = empty + F_MAX_NS;
  BGGE_DEVASSERT(reinterpret_cast<unsigned char*>(pbar + NPX) <= fsm + dynamic_smem_bytes() && ns >= 3 && ns <= F_MAX_NS &&
                 cs <= F_MAXCS);

  // Rows of a ring slot beyond this CTA's slice are never written by the TMA copies; zero the ring once so
  // they stay finite: the hot loops then need no per-row predicates (such rows meet e = 0 in the dots and
  // their u accumulators are never written back).
  // (a read of the last row pairs may run past a slot into the next slot or into the small arrays behind the
  // ring, so everything up to the barriers is zeroed, not only the ring)
  for (size_t i = threadIdx.x; i < (size_t)(reinterpret_cast<double*>(full) - ring); i += F_THREADS) ring[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < ns; ++k) {
      mbar_init(&full[k], 1);
      mbar_init(&empty[k], F_ROLE / 32);
    }
    for (int k = 0; k < npx; ++k) mbar_init(&pbar[k], 1);
    mbar_fence_init();
  }
  cluster.sync();   // barriers of every CTA exist before anybody sends

  // The producer only reads the eigen basis (constant) and starts filling the ring while the previous launch
  // (gather / finalize / scalars of the step before) is still finishing; everybody else needs its results.
  if (warp != 2 * (F_ROLE / 32)) pdl_wait();
  const double shift = sc->shift;
  const double tau = sc->tau;
  const long long it = sc->it;

  if (warp == 2 * (F_ROLE / 32)) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      int pslot = 0;
      uint32_t pphase = 0;
      for (int pi = 0; pi < cp.y; ++pi) {
        const FPanel& pn = panels[cp.x + pi];   // stays in global memory: members are indexed dynamically
        const int row0 = (int)rank * pn.rs;
        const int my_rows = max(0, min(pn.rs, pn.rows - row0));
        const uint32_t slice_bytes = (uint32_t)((my_rows + 1) & ~1) * (uint32_t)sizeof(double);
        const double* Ubase = pn.U + row0;
        for (int c0 = pn.c0; c0 < pn.c1; c0 += G) {
          const int slot = pslot;
          const int gc = min(G, pn.c1 - c0);
          BGGE_DEVASSERT(slot < ns && slice_bytes % 16 == 0 && my_rows <= max_rs && gc * max_rs <= (int)slot_doubles);
          mbar_wait_relaxed(&empty[slot], pphase ^ 1u);
          if (++pslot == ns) {
            pslot = 0;
            pphase ^= 1u;
          }
          mbar_expect_tx(&full[slot], slice_bytes * (uint32_t)gc);
          if (slice_bytes > 0) {
            double* dst = ring + (size_t)slot * slot_doubles;
            for (int g = 0; g < gc; ++g)
              bulk_g2s(dst + (size_t)g * max_rs, Ubase + (long long)(c0 + g) * pn.ldu, slice_bytes, &full[slot]);
          }
        }
      }
    }
  } else if (warp < F_ROLE / 32) {
    // ------------------------------------------------------------------ dot warps
    const int tid = threadIdx.x;
    uint32_t cc = 0;
    int dslot = 0, dpg = 0;
    uint32_t dphase = 0;
    for (int pi = 0; pi < cp.y; ++pi) {
      const FPanel& pn = panels[cp.x + pi];   // stays in global memory: members are indexed dynamically
      const int row0 = (int)rank * pn.rs;
      const int my_rows = max(0, min(pn.rs, pn.rows - row0));
      // residual slices e = (r - shift) + u_old (one per member block), register resident for the whole panel
      double2 e[NB][RPT];
#pragma unroll
      for (int m = 0; m < NB; ++m) {
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int row = k * F_ROWS_PER_STEP + 2 * tid;
          e[m][k] = make_double2(0.0, 0.0);
          if (m < pn.nb) {
            if (pn.virt && !pn.direct) {   // factorised block: the residual was already gathered to genotype space
              const long long gi = pn.voffs[m] + row0 + row;
              e[m][k].x = row < my_rows ? ve[gi] : 0.0;
              e[m][k].y = row + 1 < my_rows ? ve[gi + 1] : 0.0;
            } else {
              const long long gi = (long long)pn.pos0s[m] + row0 + row;
              const double* u_old = u_all + (size_t)pn.kids[m] * n;
              e[m][k].x = row < my_rows ? (r[gi] - shift) + u_old[gi] : 0.0;
              e[m][k].y = row + 1 < my_rows ? (r[gi + 1] - shift) + u_old[gi + 1] : 0.0;
            }
          }
        }
      }
      for (int c0 = pn.c0; c0 < pn.c1; c0 += G, ++cc) {
        const int slot = dslot;
        const int gc = min(G, pn.c1 - c0);
        BGGE_DEVASSERT(slot >= 0 && slot < ns && dpg >= 0 && dpg < npx && npx <= NPX);
        mbar_wait_relaxed(&full[slot], dphase);
        if (++dslot == ns) {
          dslot = 0;
          dphase ^= 1u;
        }
        const double* src = ring + (size_t)slot * slot_doubles + 2 * tid;
        double* rd = red + (cc & 1u) * 8 * NG;
        constexpr int PV = NG <= 1 ? 1 : NG <= 2 ? 2 : NG <= 4 ? 4 : NG <= 8 ? 8 : NG <= 16 ? 16 : 32;
        double a[PV];
#pragma unroll
        for (int v = 0; v < PV; ++v) a[v] = 0.0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g < gc) {
            const double* sg = src + (size_t)g * max_rs;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
              // no row predicate: rows past the slice are finite (zeroed ring) and meet e = 0
              const double2 x = *reinterpret_cast<const double2*>(sg + k * F_ROWS_PER_STEP);
#pragma unroll
              for (int m = 0; m < NB; ++m) {
                a[m * G + g] = fma(x.x, e[m][k].x, a[m * G + g]);
                a[m * G + g] = fma(x.y, e[m][k].y, a[m * G + g]);
              }
            }
          }
        }
        {
          int vi;
          const double tot = warp_multi_reduce<PV>(a, lane, vi);
          if ((lane & (32 / PV - 1)) == 0 && vi < NG) rd[warp * NG + vi] = tot;
        }
        role_barrier(1);
        const uint32_t pg = (uint32_t)dpg;
        if (++dpg == npx) dpg = 0;
        if (warp == 0) {
          if (lane == 0) mbar_expect_tx(&pbar[pg], cs * (uint32_t)(gc * pn.nb) * (uint32_t)sizeof(double));
          __syncwarp();
          // lane l owns value v = l % 16 (member m = v / G, column g = v % G): it adds the 8 warp partials in
          // warp order (deterministic) and pushes the sum to half of the cluster (l / 16 selects the half)
          const int v = NG <= 16 ? (lane & 15) : lane;
          const uint32_t q0 = NG <= 16 ? (uint32_t)(lane >> 4) : 0u, qstep = NG <= 16 ? 2u : 1u;
          const int m = v / G, gq = v % G;
          if (v < NG && m < pn.nb && gq < gc) {
            double p = 0.0;
#pragma unroll
            for (int w = 0; w < F_ROLE / 32; ++w) p += rd[w * NG + v];
            const uint32_t dst = smem_u32(parts + (size_t)pg * F_MAXCS * NG + (size_t)rank * NG + v);
            const uint32_t bar_addr = smem_u32(&pbar[pg]);
            for (uint32_t q = q0; q < cs; q += qstep)
              st_async_f64(map_to_rank(dst, q), p, map_to_rank(bar_addr, q));
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ axpy warps
    const int tid = threadIdx.x - F_ROLE;
    uint32_t cc = 0;
    int aslot = 0, apg = 0;
    uint32_t aphase = 0, pphase2 = 0;
    double zq = 0.0;   // thread 0 of the role: sum of b^2 / s (R/BGGE.R:96)
    for (int pi = 0; pi < cp.y; ++pi) {
      const FPanel& pn = panels[cp.x + pi];   // stays in global memory: members are indexed dynamically
      const int row0 = (int)rank * pn.rs;
      const int my_rows = max(0, min(pn.rs, pn.rows - row0));
      const int ncols = pn.c1 - pn.c0;
      double2 uacc[NB][RPT];
#pragma unroll
      for (int m = 0; m < NB; ++m)
#pragma unroll
        for (int k = 0; k < RPT; ++k) uacc[m][k] = make_double2(0.0, 0.0);
      for (int c0 = pn.c0; c0 < pn.c1; c0 += G, ++cc) {
        const int slot = aslot;
        const uint32_t sphase = aphase;
        if (++aslot == ns) {
          aslot = 0;
          aphase ^= 1u;
        }
        const int gc = min(G, pn.c1 - c0);
        const int rel0 = c0 - pn.c0;
        if (rel0 % ZB == 0) {
          // draw batch for panel columns [rel0, rel0 + ZB): the normal draw does not depend on d
          const int rel = rel0 + tid;
          if (tid < ZB && rel < ncols) {
            const double sv = pn.ss[0][pn.c0 + rel];                               // members share s (same matrix)
            const int zb = ((rel0 / ZB) & 1) * ZB + tid;
            isb[zb] = 1.0 / sv;
            double sq = 0.0;
            for (int m = 0; m < pn.nb; ++m) {
              if (m == 0 || merged) {                                              // one sigb per kernel
                const double lam = sc->sigb[pn.kids[m]];
                const double vari = sv * lam / (1.0 + sv * lam * tau);             // R/BGGE.R:302
                tvb[(size_t)m * 2 * ZB + zb] = tau * vari;
                sq = sqrt(vari);
              }
              const long long cm = (long long)pn.col_offs[m] + pn.c0 + rel;
              const double z = draw_normal(dc, it, (uint32_t)pn.kids[m], (unsigned long long)cm,
                                           dc.z_init + (it - 1) * dc.z_per_it + pn.zoffs[m] + cm);
              zsb[(size_t)m * 2 * ZB + zb] = sq * z;
            }
          }
          role_barrier(2);
        }
        const uint32_t pg = (uint32_t)apg;
        mbar_wait_relaxed(&pbar[pg], pphase2);
        if (++apg == npx) {
          apg = 0;
          pphase2 ^= 1u;
        }
        double bv[NB == 1 ? G : 1];
        const double* bcur = bsm + (cc & 1u) * NG;
        if constexpr (NB == 1) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            bv[g] = 0.0;
            if (g < gc) {
              const double* pp = parts + (size_t)pg * F_MAXCS * NG + g;
              double d = 0.0;
              for (uint32_t q = 0; q < cs; ++q) d += pp[(size_t)q * NG];          // rank order
              const int zb = (((rel0 + g) / ZB) & 1) * ZB + (rel0 + g) % ZB;
              const double media = tvb[zb] * d;                                     // (tau vari) d   R/BGGE.R:303
              bv[g] = media + zsb[zb];                                              // + sqrt(vari) z R/BGGE.R:89-92
              if (tid == 0) zq += bv[g] * bv[g] * isb[zb];
            }
          }
        } else {
          // one lane per (member, column) forms the coefficient once; everybody reads it after the barrier
          if (tid < pn.nb * G) {
            const int m = tid / G, g = tid % G;
            double bcoef = 0.0;
            if (g < gc) {
              const double* pp = parts + (size_t)pg * F_MAXCS * NG + m * G + g;
              double d = 0.0;
              for (uint32_t q = 0; q < cs; ++q) d += pp[(size_t)q * NG];
              const int zb = (((rel0 + g) / ZB) & 1) * ZB + (rel0 + g) % ZB;
              bcoef = tvb[(size_t)(merged ? m : 0) * 2 * ZB + zb] * d + zsb[(size_t)m * 2 * ZB + zb];
              zq += bcoef * bcoef * isb[zb];
            }
            bsm[(cc & 1u) * NG + m * G + g] = bcoef;
          }
          role_barrier(2);
        }
        mbar_wait_relaxed(&full[slot], sphase);   // acquire the TMA-written slot for this role too
        const double* src = ring + (size_t)slot * slot_doubles + 2 * tid;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g < gc) {
            double2 x[RPT];
            const double* sg = src + (size_t)g * max_rs;
#pragma unroll
            for (int k = 0; k < RPT; ++k) x[k] = *reinterpret_cast<const double2*>(sg + k * F_ROWS_PER_STEP);
#pragma unroll
            for (int m = 0; m < NB; ++m) {
              const double bb = NB == 1 ? bv[NB == 1 ? g : 0] : bcur[m * G + g];
#pragma unroll
              for (int k = 0; k < RPT; ++k) {
                uacc[m][k].x = fma(x[k].x, bb, uacc[m][k].x);
                uacc[m][k].y = fma(x[k].y, bb, uacc[m][k].y);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
      // this cluster's contribution to u over the block's rows
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        if (m < pn.nb) {
          double* up = pn.virt ? vupart + (size_t)pn.slot * vtotal + pn.voffs[m] + row0
                               : upart + (size_t)pn.slot * n + pn.pos0s[m] + row0;
#pragma unroll
          for (int k = 0; k < RPT; ++k) {
            const int row = k * F_ROWS_PER_STEP + 2 * tid;
            if (row < my_rows) up[row] = uacc[m][k].x;
            if (row + 1 < my_rows) up[row + 1] = uacc[m][k].y;
          }
        }
      }
      role_barrier(2);   // batch buffers are reused by the next panel
    }
    // all panels of a launch belong to one kernel, or to one merged unit (one group): a cluster without panels
    // finds the destination of its (zero) sum in panel 0
    const FPanel& p0 = panels[cp.y > 0 ? cp.x : 0];
    if constexpr (NB > 1) {   // the (member, column) lanes of warp 0 of this role hold the pieces of sum b^2/s
      if (merged) {           // one sum per member kernel
        for (int m = 0; m < p0.nb; ++m) {
          const double zm = warp_sum(tid < 32 && tid / G == m ? zq : 0.0);
          if (rank == 0 && tid == 0) p0.bps[m][cid] = zm;
        }
      } else {
        zq = warp_sum(tid < 32 ? zq : 0.0);
      }
    }
    if (!merged && rank == 0 && tid == 0) p0.bps[0][cid] = zq;
  }
  cluster.sync();      // nobody leaves while a peer may still write into its shared memory
}

// u_j = sum over clusters of upart; e <- (e + u_old) - u_new; Welford  (R/BGGE.R:306-307, 392).
// A tile is 256 rows; 8 warps take 32-row groups in turn and the 8 "q-lanes" of a row (threadIdx.y) sum
// interleaved cluster slots, then combine in a fixed tree -> deterministic, and 8x more loads in flight
// than one thread walking up to 148 partials.
constexpr int FIN_ROWS = 32, FIN_Q = 8;
__global__ void __launch_bounds__(FIN_ROWS * FIN_Q)
fused_finalize_kernel(const FTile* __restrict__ tiles, const double* __restrict__ upart, long long n,
                      const double* __restrict__ vupart, long long vtotal, double* __restrict__ r, double* __restrict__ u_all,
                      double* __restrict__ umean_all, double* __restrict__ um2_all, const Scal* __restrict__ sc, long long thin,
                      long long burn) {
  __shared__ double part[FIN_Q][FIN_ROWS];
  pdl_trigger();
  const FTile t = tiles[blockIdx.x];
  if (t.kid < 0) return;   // merged unit: rows none of its kernels covers (their u is zero there already)
  double* u = u_all + (size_t)t.kid * n;
  double* umean = umean_all + (size_t)t.kid * n;
  double* um2 = um2_all + (size_t)t.kid * n;
  const int lane = threadIdx.x, ql = threadIdx.y;
  pdl_wait();
  const long long it = sc->it;
  const bool keep = (it % thin == 0) && (it > burn);
  const double kcount = keep ? (double)(it / thin - burn / thin) : 1.0;
  for (int row0 = 0; row0 < t.rows; row0 += FIN_ROWS) {
    const int row = row0 + lane;
    const long long g = (long long)t.row0 + row;
    double a = 0.0;
    if (row < t.rows) {
      if (t.loc) {   // factorised block: u_i = isq_i * (W b)[loc_i]
        const long long v = t.voff + t.loc[row];
        for (int q = t.qlo + ql; q <= t.qhi; q += FIN_Q) a += vupart[(size_t)q * vtotal + v];
      } else {
        for (int q = t.qlo + ql; q <= t.qhi; q += FIN_Q) a += upart[(size_t)q * n + g];
      }
    }
    part[ql][lane] = a;
    __syncthreads();
    if (ql == 0 && row < t.rows) {
      double un = ((part[0][lane] + part[1][lane]) + (part[2][lane] + part[3][lane])) +
                  ((part[4][lane] + part[5][lane]) + (part[6][lane] + part[7][lane]));
      if (t.loc) un *= t.isq[row];
      r[g] = (r[g] + u[g]) - un;
      u[g] = un;
      if (keep) {
        const double m = umean[g];
        const double dlt = un - m;
        const double mn = m + dlt / kcount;
        umean[g] = mn;
        um2[g] += dlt * (un - mn);
      }
    }
    __syncthreads();
  }
}

// factorised blocks: ve[voff + a] = sum over the rows i of genotype a of isq_i * ((r_i - shift) + u_old_i), in
// CSR order (deterministic); one launch covers every factorised block of the kernel
__global__ void __launch_bounds__(256)
virt_gather_kernel(const VDesc* __restrict__ desc, int ndesc, const double* __restrict__ r,
                   const double* __restrict__ u_all, long long n, const Scal* __restrict__ sc, double* __restrict__ ve) {
  pdl_trigger();
  int d = 0;
  while (d + 1 < ndesc && (int)blockIdx.x >= desc[d + 1].cta0) ++d;
  const VDesc vd = desc[d];
  const int a = ((int)blockIdx.x - vd.cta0) * blockDim.x + threadIdx.x;
  const int k0 = a < vd.vrows ? vd.rowptr[a] : 0, k1 = a < vd.vrows ? vd.rowptr[a + 1] : 0;   // constant tables
  pdl_wait();
  if (a >= vd.vrows) return;
  const double shift = sc->shift;
  const double* rp = r + vd.pos0;
  const double* up = u_all + (size_t)vd.kid * n + vd.pos0;
  double acc = 0.0;
  for (int k = k0; k < k1; ++k) {
    const int i = vd.rowidx[k];
    acc = fma(vd.isq[i], (rp[i] - shift) + up[i], acc);
  }
  ve[vd.voff + a] = acc;
}

// Column-sharded chain: the cluster partials of this process, summed in fixed order, -> one vector the host
// all-reduces.  Segments are <= 1024 entries; the extra last CTA adds up the clusters' sum b^2/s.
__global__ void __launch_bounds__(256)
shard_sum_kernel(const SSeg* __restrict__ segs, int nsegs, const double* __restrict__ bpart, int nclusters,
                 double* __restrict__ xbuf, long long zq_index) {
  pdl_trigger();
  pdl_wait();
  if ((int)blockIdx.x == nsegs) {
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int c = 0; c < nclusters; ++c) t += bpart[c];
      xbuf[zq_index] = t;
    }
    return;
  }
  const SSeg sg = segs[blockIdx.x];
  for (int i = threadIdx.x; i < sg.len; i += blockDim.x) {
    double a = 0.0;
    for (int q = sg.qlo; q <= sg.qhi; ++q) a += sg.src[(size_t)q * sg.stride + sg.off + i];
    xbuf[sg.doff + i] = a;
  }
}

template <int RPT, int G, int NB>
int launch_fused_t(Gibbs* g, HostKernel& hk, int j) {
  auto kfn = sweep_fused_kernel<RPT, G, NB>;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(hk.fused_clusters * hk.fused_cs), 1, 1);
  cfg.blockDim = dim3(F_THREADS, 1, 1);
  cfg.dynamicSmemBytes = hk.fused_smem;
  cfg.stream = g->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)hk.fused_cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)j;
  BGGE_CUDA(cudaLaunchKernelEx(&cfg, kfn, (const FPanel*)hk.fpanels.p, (const int2*)hk.fcluster.p,
                               (const double*)g->r.p, (const double*)g->u.p, g->upart.p, (long long)g->n,
                               (const double*)hk.ve.p, g->vupart.p, (long long)hk.vtotal,
                               (const Scal*)g->scal.p, hk.merged ? 1 : 0, g->draw, hk.fused_ns, 2 * hk.fused_ns + 4));
  return OK;
}

struct Combo { int rpt, gcols; };
constexpr Combo kCombos[] = {{1, 8}, {2, 6}, {3, 4}, {4, 3}, {6, 2}, {8, 1}, {10, 1}, {12, 1}};   // slot <= 48 KB
// register budget: NB * RPT <= 12 row pairs per thread -> NB = 2 up to RPT 6, NB = 4 up to RPT 3

template <int RPT, int G, int NB>
int occupancy_t(int cs, size_t smem, int* nclusters) {
  auto kfn = sweep_fused_kernel<RPT, G, NB>;
  // one template instantiation serves several model kernels with different ring sizes: opt in to the
  // device maximum once instead of per launch
  int dev = 0, optin = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  BGGE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  (void)dev;
  BGGE_REQUIRE((size_t)optin >= smem, "fused plan: %zu bytes of shared memory exceed the device limit %d", smem, optin);
  BGGE_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  if (cs > 8) BGGE_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(cs * 64), 1, 1);
  cfg.blockDim = dim3(F_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BGGE_CUDA(cudaOccupancyMaxActiveClusters(nclusters, kfn, &cfg));
  return OK;
}

#define BGGE_FUSED_CASE(FN, R, GG, ...)                                  \
  case R:                                                                \
    if (hk.fused_nb == 1) return FN<R, GG, 1>(__VA_ARGS__);              \
    if constexpr (R <= 6) {                                              \
      if (hk.fused_nb == 2) return FN<R, GG, 2>(__VA_ARGS__);            \
    }                                                                    \
    if constexpr (R <= 3) {                                              \
      if (hk.fused_nb == 4) return FN<R, GG, 4>(__VA_ARGS__);            \
    }                                                                    \
    break;

#define BGGE_FUSED_DISPATCH(FN, ...)                       \
  switch (hk.fused_rpt) {                                  \
    BGGE_FUSED_CASE(FN, 1, 8, __VA_ARGS__)                 \
    BGGE_FUSED_CASE(FN, 2, 6, __VA_ARGS__)                 \
    BGGE_FUSED_CASE(FN, 3, 4, __VA_ARGS__)                 \
    BGGE_FUSED_CASE(FN, 4, 3, __VA_ARGS__)                 \
    BGGE_FUSED_CASE(FN, 6, 2, __VA_ARGS__)                 \
    BGGE_FUSED_CASE(FN, 8, 1, __VA_ARGS__)                 \
    BGGE_FUSED_CASE(FN, 10, 1, __VA_ARGS__)                \
    BGGE_FUSED_CASE(FN, 12, 1, __VA_ARGS__)                \
    default: break;                                        \
  }                                                        \
  set_error("fused sweep: no kernel instantiation for rpt=%d nb=%d", hk.fused_rpt, hk.fused_nb); \
  return ERR_INVALID;

int occupancy_dispatch(HostKernel& hk, int cs, size_t smem, int* ncl) { BGGE_FUSED_DISPATCH(occupancy_t, cs, smem, ncl) }

}  // namespace

int launch_fused(Gibbs* g, HostKernel& hk, int j) {
  if (hk.n_vdesc > 0) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)hk.vgrid, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = g->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    BGGE_CUDA(cudaLaunchKernelEx(&cfg, virt_gather_kernel, (const VDesc*)hk.vdesc.p, hk.n_vdesc, (const double*)g->r.p,
                                 (const double*)g->u.p, (long long)g->n, (const Scal*)g->scal.p, hk.ve.p));
  }
  if (hk.fused_engine == 1) return mrhs_launch(g, hk);
  auto run = [&]() -> int { BGGE_FUSED_DISPATCH(launch_fused_t, g, hk, j) };
  return run();
}

int launch_fused_finalize(Gibbs* g, HostKernel& hk, int j) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)hk.n_ftiles, 1, 1);
  cfg.blockDim = dim3(FIN_ROWS, FIN_Q, 1);
  cfg.stream = g->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)j;
  BGGE_CUDA(cudaLaunchKernelEx(&cfg, fused_finalize_kernel, (const FTile*)hk.ftiles.p, (const double*)g->upart.p, (long long)g->n,
                               (const double*)g->vupart.p, (long long)hk.vtotal, g->r.p, g->u.p, g->umean.p, g->um2.p,
                               (const Scal*)g->scal.p, (long long)g->thin, (long long)g->burn));
  return OK;
}

namespace {
int launch_pdl(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr) {
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return OK;
}
}  // namespace

int launch_shard_sum(Gibbs* g, HostKernel& hk) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)hk.n_ssegs + 1, 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = g->stream;
  cudaLaunchAttribute attr[1];
  launch_pdl(cfg, attr);
  BGGE_CUDA(cudaLaunchKernelEx(&cfg, shard_sum_kernel, (const SSeg*)hk.ssegs.p, hk.n_ssegs, (const double*)hk.bpart.p,
                               hk.fused_clusters, hk.xbuf.p, (long long)(g->n + hk.vtotal)));
  return OK;
}

// u, r, Welford from the all-reduced vector: the finalize kernel with one "cluster" whose partials are the sums
int launch_shard_apply(Gibbs* g, HostKernel& hk) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)hk.n_ftiles, 1, 1);
  cfg.blockDim = dim3(FIN_ROWS, FIN_Q, 1);
  cfg.stream = g->stream;
  cudaLaunchAttribute attr[1];
  launch_pdl(cfg, attr);
  BGGE_CUDA(cudaLaunchKernelEx(&cfg, fused_finalize_kernel, (const FTile*)hk.atiles.p, (const double*)hk.xbuf.p, (long long)g->n,
                               (const double*)(hk.xbuf.p + g->n), (long long)hk.vtotal, g->r.p, g->u.p, g->umean.p, g->um2.p,
                               (const Scal*)g->scal.p, (long long)g->thin, (long long)g->burn));
  return OK;
}

// Decide whether kernel j can use the fused path and build its panel / tile tables.
// Returns OK with hk.fused == false when the shape does not fit (caller keeps the two-pass kernels).
int plan_fused(Gibbs* g, HostKernel& hk, int j, cudaStream_t st) {
  hk.fused = false;
  // groups of blocks that stream the same matrix: a dense / unshared block alone, or up to 4 factorised
  // blocks sharing one W (identical diagonal blocks) swept as one multi-RHS panel
  struct Group {
    std::vector<int> members;
    int64_t rows, nr;
    int64_t c_lo = 0, c_hi = 0;   // columns this process sweeps (all of them unless the chain is column-sharded)
  };
  std::vector<Group> groups;
  hk.vtotal = 0;
  int64_t max_rows = 0;
  int nbmax = 1;
  for (size_t bi = 0; bi < hk.blocks.size(); ++bi) {
    HostBlock& bl = hk.blocks[bi];
    if (bl.vrows > 0) {
      bl.voff = hk.vtotal;
      hk.vtotal += round_up(bl.vrows, 16);
    }
    max_rows = std::max(max_rows, bl.stream_rows());
    bool placed = false;
    if (bl.vrows > 0 || hk.merged)   // a merged unit also shares a matrix between dense blocks (U = W)
      for (auto& gr : groups) {
        const HostBlock& b0 = hk.blocks[(size_t)gr.members[0]];
        if ((b0.vrows > 0) == (bl.vrows > 0) && b0.dU == bl.dU && b0.vrows == bl.vrows && b0.rows == bl.rows &&
            b0.nr == bl.nr && gr.members.size() < 4) {
          gr.members.push_back((int)bi);
          nbmax = std::max(nbmax, (int)gr.members.size());
          placed = true;
          break;
        }
      }
    if (!placed) groups.push_back(Group{{(int)bi}, bl.stream_rows(), bl.nr});
  }
  if (max_rows <= 0) return OK;
  const int nb_t = nbmax > 2 ? 4 : nbmax;            // instantiated right-hand-side counts: 1, 2, 4
  int cs = 1, gcols = 1, ncl = 0;
  hk.fused_engine = 0;
  // multi-RHS panels go to sweep_mrhs_kernel (register-resident e and u at 4-CTA clusters, helper warps for the
  // exchange); BGGE_MRHS=0 keeps them on sweep_fused_kernel
  if (nb_t >= 2 && !(getenv("BGGE_MRHS") && atoi(getenv("BGGE_MRHS")) == 0) &&
      (max_rows >= mrhs_min_rows() || (getenv("BGGE_MRHS") && atoi(getenv("BGGE_MRHS")) == 2))) {
    MrhsCfg mc;
    BGGE_TRY(mrhs_configure(max_rows, nb_t, hk.merged, &mc));
    if (mc.ok) {
      hk.fused_engine = 1;
      cs = mc.cs;
      gcols = mc.g;
      hk.fused_tmem = mc.tmem;
      ncl = mc.clusters;
      hk.fused_cs = cs;
      hk.fused_rpt = 0;
      hk.fused_g = gcols;
      hk.fused_nb = nb_t;
      hk.fused_ns = mc.ns;
      hk.fused_kr = mc.kr;
      hk.fused_lag = mc.lag;
      hk.fused_smem = mc.smem;
    }
  }
  if (hk.fused_engine == 0) {
    const int rpt_cap = 12 / nb_t;                     // register budget: NB * RPT <= 12 row pairs per thread
    // smallest cluster whose row slice fits the register budget: long contiguous column slices (~40 KB)
    // stream best and small clusters leave no SM idle
    while (cs < 8 && max_rows > (int64_t)cs * rpt_cap * F_ROWS_PER_STEP) cs *= 2;
    if (const char* ev = getenv("BGGE_FUSED_CS")) {   // tuning hook
      const int v = atoi(ev);
      if (v == 1 || v == 2 || v == 4 || v == 8) cs = std::max(cs, v);
    }
    const int64_t rs_max = round_up((max_rows + cs - 1) / cs, 16);
    const int rpt_need = (int)((rs_max + F_ROWS_PER_STEP - 1) / F_ROWS_PER_STEP);
    const Combo* combo = nullptr;
    for (const Combo& c : kCombos)
      if (c.rpt >= rpt_need && c.rpt <= rpt_cap) { combo = &c; break; }
    if (!combo) return OK;   // slices too long for the register budget: two-pass path
    gcols = combo->gcols;
    hk.fused_cs = cs;
    hk.fused_rpt = combo->rpt;
    hk.fused_g = combo->gcols;
    hk.fused_nb = nb_t;
    const size_t slot_bytes = (size_t)combo->gcols * rs_max * sizeof(double);
    const size_t zb = (size_t)combo->gcols * (F_ROLE / combo->gcols);
    const size_t ng = (size_t)nb_t * combo->gcols;
    int dev = 0, optin = 0;
    BGGE_CUDA(cudaGetDevice(&dev));
    BGGE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t tvm = hk.merged ? (size_t)nb_t : 1;   // tau * v per member kernel in a merged unit
    auto fixed_bytes = [&](int nsv) {   // exchange ring depth 2 ns + 4 (protocol note) -> grows with the data ring
      return ((size_t)(2 * nsv + 4) * F_MAXCS * ng + 2 * 8 * ng + 2 * ng + (2 * tvm + 2) * zb + (size_t)nb_t * 2 * zb) * sizeof(double) +
             (2 * F_MAX_NS + NPX) * sizeof(uint64_t) + 128;
    };
    int ns = 0;
    for (int c = F_MAX_NS; c >= 3; --c)
      if ((size_t)c * slot_bytes + fixed_bytes(c) + 1024 <= (size_t)optin) { ns = c; break; }
    if (ns < 3) return OK;
    const size_t fixed = fixed_bytes(ns);
    hk.fused_ns = ns;
    hk.fused_smem = ns * slot_bytes + fixed;
    BGGE_TRY(occupancy_dispatch(hk, cs, hk.fused_smem, &ncl));
  }
  if (ncl < 1) return OK;
  if (const char* ev = getenv("BGGE_FUSED_MAX_CLUSTERS")) ncl = std::max(1, std::min(ncl, atoi(ev)));   // test / tuning hook
  // column shard (bgge_gibbs_set_shard): rank r of w sweeps columns [nr r / w, nr (r + 1) / w) of every matrix,
  // boundaries on multiples of the column-group size
  for (auto& gr : groups) {
    const int64_t gc = gcols;
    gr.c_lo = g->shard_world > 1 ? gr.nr * g->shard_rank / g->shard_world / gc * gc : 0;
    gr.c_hi = g->shard_world > 1 && g->shard_rank + 1 < g->shard_world ? gr.nr * (g->shard_rank + 1) / g->shard_world / gc * gc
                                                                      : gr.nr;
  }
  // work-balanced contiguous split of the (group, column) space over the clusters
  double total = 0.0;
  int64_t total_cols = 0;
  for (auto& gr : groups) {
    total += (double)gr.rows * (double)(gr.c_hi - gr.c_lo);
    total_cols += gr.c_hi - gr.c_lo;
  }
  ncl = (int)std::min<int64_t>(ncl, std::max<int64_t>(1, total_cols / std::max(1, gcols)));
  std::vector<FPanel> panels;
  std::vector<int2> cpan((size_t)ncl, make_int2(0, 0));
  std::vector<std::pair<int, int>> block_q(hk.blocks.size(), {1 << 30, -1});
  {
    const double quota = total / (double)ncl;
    size_t gi = 0;
    while (total_cols > 0 && gi < groups.size() && groups[gi].c_lo >= groups[gi].c_hi) ++gi;
    int64_t col = gi < groups.size() ? groups[gi].c_lo : 0;   // next unassigned column of group gi
    for (int q = 0; q < ncl; ++q) {
      const bool last = q == ncl - 1;
      double need = quota;
      cpan[(size_t)q].x = (int)panels.size();
      while (gi < groups.size() && (last || need > 0.02 * quota)) {
        Group& gr = groups[gi];
        const int64_t avail = gr.c_hi - col;
        int64_t take = avail;
        if (!last) {
          const int64_t want = (int64_t)std::ceil(need / (double)gr.rows);
          take = std::min<int64_t>(avail, (want + gcols - 1) / gcols * gcols);
        }
        if (take <= 0 && total_cols > 0) break;   // (a shard without any column gets one empty panel, below)
        const HostBlock& b0 = hk.blocks[(size_t)gr.members[0]];
        FPanel p{};
        p.U = b0.dU;
        p.ldu = b0.ldu;
        p.rows = (int)gr.rows;
        p.pos0 = (int)b0.pos0;
        p.c0 = (int)col;
        p.c1 = (int)(col + take);
        p.col_off = (int)b0.col_off;
        p.rs = (int)round_up((gr.rows + cs - 1) / cs, 16);
        p.slot = q;
        p.virt = b0.vrows > 0 ? 1 : 0;
        p.voff = b0.voff;
        p.direct = 1;
        for (size_t m = 0; m < gr.members.size(); ++m) p.direct = p.direct && hk.blocks[(size_t)gr.members[m]].identity ? 1 : 0;
        // opt-in: reading r and u in place saves the gather launch, but measured no faster (C5 3050 vs 3119 it/s, C3
        // 9642 vs 9806, profiles/r02_notes.md) -- the gather overlaps the previous finalize under PDL anyway
        if (!getenv("BGGE_DIRECT_RESIDUAL")) p.direct = 0;
        p.nb = (int)gr.members.size();
        for (size_t m = 0; m < gr.members.size(); ++m) {
          const HostBlock& bm = hk.blocks[(size_t)gr.members[m]];
          const int kid = bm.kid >= 0 ? bm.kid : j;
          HostKernel& owner = g->kernels[(size_t)kid];
          p.voffs[m] = bm.voff;
          p.col_offs[m] = (int)bm.col_off;
          p.kids[m] = kid;
          p.pos0s[m] = (int)bm.pos0;
          p.zoffs[m] = owner.z_off;
          p.ss[m] = owner.s.p + bm.col_off;
          p.bps[m] = owner.bpart.p;
          block_q[(size_t)gr.members[m]].first = std::min(block_q[(size_t)gr.members[m]].first, q);
          block_q[(size_t)gr.members[m]].second = std::max(block_q[(size_t)gr.members[m]].second, q);
        }
        panels.push_back(p);
        need -= (double)gr.rows * (double)take;
        col += take;
        if (total_cols == 0) {   // empty shard: the panel has no columns, the launch only zeroes its partial sums
          gi = groups.size();
          break;
        }
        if (col >= gr.c_hi) {
          ++gi;
          while (gi < groups.size() && groups[gi].c_lo >= groups[gi].c_hi) ++gi;
          col = gi < groups.size() ? groups[gi].c_lo : 0;
        }
      }
      cpan[(size_t)q].y = (int)panels.size() - cpan[(size_t)q].x;
    }
    BGGE_REQUIRE(gi == groups.size(), "fused plan: internal error (columns left unassigned)");
  }
  // gather descriptors of the factorised blocks
  {
    std::vector<VDesc> vd;
    int cta = 0;
    // (blocks whose row map is the identity are read in place by the sweep -- see FPanel::direct -- unless another
    // member of their panel needs the gathered form)
    std::vector<char> need_gather(hk.blocks.size(), 0);
    for (auto& p : panels)
      if (p.virt && !p.direct)
        for (size_t bi = 0; bi < hk.blocks.size(); ++bi)
          for (int m = 0; m < p.nb; ++m)
            if (hk.blocks[bi].vrows > 0 && hk.blocks[bi].voff == p.voffs[m]) need_gather[bi] = 1;
    for (size_t bi = 0; bi < hk.blocks.size(); ++bi) {
      auto& bl = hk.blocks[bi];
      if (bl.vrows > 0 && need_gather[bi]) {
        VDesc d{};
        d.vrows = (int)bl.vrows;
        d.pos0 = (int)bl.pos0;
        d.cta0 = cta;
        d.voff = bl.voff;
        d.rowptr = bl.d_rowptr;
        d.rowidx = bl.d_rowidx;
        d.isq = bl.d_isq;
        d.kid = bl.kid >= 0 ? bl.kid : j;
        vd.push_back(d);
        cta += (int)((bl.vrows + 255) / 256);
      }
    }
    hk.n_vdesc = (int)vd.size();
    hk.vgrid = cta;
    if (!vd.empty()) {
      BGGE_CUDA(hk.vdesc.alloc(vd.size()));
      BGGE_CUDA(cudaMemcpyAsync(hk.vdesc.p, vd.data(), vd.size() * sizeof(VDesc), cudaMemcpyHostToDevice, st));
      BGGE_CUDA(cudaStreamSynchronize(st));
    }
  }
  // finalize tiles: every row of y, gaps included (qlo > qhi -> zero)
  std::vector<FTile> tiles;
  int64_t cursor = 0;
  auto add_rows = [&](int64_t pos0, int64_t rows, int qlo, int qhi, int kid, const HostBlock* vb = nullptr) {
    for (int64_t r0 = 0; r0 < rows; r0 += 64) {
      FTile t{};
      t.row0 = (int)(pos0 + r0);
      t.rows = (int)std::min<int64_t>(64, rows - r0);
      t.qlo = qlo;
      t.qhi = qhi;
      t.kid = kid;
      if (vb) {
        t.loc = vb->d_loc + r0;
        t.isq = vb->d_isq + r0;
        t.voff = vb->voff;
      }
      tiles.push_back(t);
    }
  };
  for (size_t bi = 0; bi < hk.blocks.size(); ++bi) {
    HostBlock& bl = hk.blocks[bi];
    // rows no block covers: u_j = 0 there (R/BGGE.R:334); a merged unit leaves them alone (kid -1)
    if (bl.pos0 > cursor) add_rows(cursor, bl.pos0 - cursor, 1, 0, hk.merged ? -1 : j);
    add_rows(bl.pos0, bl.rows, block_q[bi].first, block_q[bi].second, bl.kid >= 0 ? bl.kid : j, bl.vrows > 0 ? &bl : nullptr);
    cursor = bl.pos0 + bl.rows;
  }
  if (cursor < g->n) add_rows(cursor, g->n - cursor, 1, 0, hk.merged ? -1 : j);
  hk.fused_clusters = ncl;
  hk.n_ftiles = (int)tiles.size();
  if (hk.vtotal > 0) {
    BGGE_CUDA(hk.ve.alloc((size_t)hk.vtotal));
    BGGE_CUDA(cudaMemsetAsync(hk.ve.p, 0, (size_t)hk.vtotal * sizeof(double), st));
  }
  BGGE_CUDA(hk.fpanels.alloc(panels.size()));
  BGGE_CUDA(cudaMemcpyAsync(hk.fpanels.p, panels.data(), panels.size() * sizeof(FPanel), cudaMemcpyHostToDevice, st));
  BGGE_CUDA(hk.fcluster.alloc(cpan.size()));
  BGGE_CUDA(cudaMemcpyAsync(hk.fcluster.p, cpan.data(), cpan.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
  BGGE_CUDA(hk.ftiles.alloc(tiles.size()));
  BGGE_CUDA(cudaMemcpyAsync(hk.ftiles.p, tiles.data(), tiles.size() * sizeof(FTile), cudaMemcpyHostToDevice, st));
  for (size_t bi = 0; bi < hk.blocks.size(); ++bi) {
    hk.blocks[bi].fq_lo = block_q[bi].first;
    hk.blocks[bi].fq_hi = block_q[bi].second;
  }
  BGGE_CUDA(cudaStreamSynchronize(st));
  hk.fused = true;
  if (getenv("BGGE_DEBUG"))
    fprintf(stderr, "[bgge] fused plan: engine=%d cs=%d clusters=%d rpt=%d kr=%d g=%d nb=%d ns=%d smem=%zu groups=%zu panels=%zu vtotal=%lld\n",
            hk.fused_engine, hk.fused_cs, hk.fused_clusters, hk.fused_rpt, hk.fused_kr, hk.fused_g, hk.fused_nb, hk.fused_ns, hk.fused_smem,
            groups.size(), panels.size(), (long long)hk.vtotal);
  return OK;
}

// Column-sharded chain: tables of shard_sum_kernel and of the apply pass.  Called from gibbs_init once the partial-sum
// buffers exist.
int plan_shard(Gibbs* g, HostKernel& hk, cudaStream_t st) {
  std::vector<SSeg> segs;
  auto add_seg = [&](const double* src, long long stride, long long off, long long doff, int64_t len, int qlo, int qhi) {
    if (qlo > qhi) return;   // no cluster of this process touches the block: its part of the vector stays zero
    for (int64_t i0 = 0; i0 < len; i0 += 1024) {
      SSeg sg{};
      sg.src = src;
      sg.stride = stride;
      sg.off = off + i0;
      sg.doff = doff + i0;
      sg.len = (int)std::min<int64_t>(1024, len - i0);
      sg.qlo = qlo;
      sg.qhi = qhi;
      segs.push_back(sg);
    }
  };
  BGGE_CUDA(hk.xbuf.alloc((size_t)(g->n + hk.vtotal + 1)));
  BGGE_CUDA(cudaMemsetAsync(hk.xbuf.p, 0, (size_t)(g->n + hk.vtotal + 1) * sizeof(double), st));
  for (const HostBlock& bl : hk.blocks) {
    if (bl.vrows > 0)
      add_seg(g->vupart.p, hk.vtotal, bl.voff, g->n + bl.voff, bl.vrows, bl.fq_lo, bl.fq_hi);
    else
      add_seg(g->upart.p, g->n, bl.pos0, bl.pos0, bl.rows, bl.fq_lo, bl.fq_hi);
  }
  hk.n_ssegs = (int)segs.size();
  BGGE_CUDA(hk.ssegs.alloc(std::max<size_t>(1, segs.size())));
  if (!segs.empty()) BGGE_CUDA(cudaMemcpyAsync(hk.ssegs.p, segs.data(), segs.size() * sizeof(SSeg), cudaMemcpyHostToDevice, st));
  // every block tile reads slot 0 of the reduced vector (a block without local columns still receives the other
  // processes' contributions); tiles of rows no block covers keep qlo > qhi (u = 0 there)
  std::vector<FTile> at((size_t)hk.n_ftiles);
  BGGE_CUDA(cudaMemcpyAsync(at.data(), hk.ftiles.p, at.size() * sizeof(FTile), cudaMemcpyDeviceToHost, st));
  BGGE_CUDA(cudaStreamSynchronize(st));
  for (auto& t : at) {
    bool in_block = false;
    for (auto& bl : hk.blocks) in_block = in_block || (t.row0 >= bl.pos0 && t.row0 < bl.pos0 + bl.rows);
    if (in_block) t.qlo = t.qhi = 0;
  }
  BGGE_CUDA(hk.atiles.alloc(std::max<size_t>(1, at.size())));
  BGGE_CUDA(cudaMemcpyAsync(hk.atiles.p, at.data(), at.size() * sizeof(FTile), cudaMemcpyHostToDevice, st));
  BGGE_CUDA(cudaStreamSynchronize(st));
  return OK;
}

}  // namespace bgge

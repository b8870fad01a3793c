// Multi-right-hand-side kernel step of the Gibbs sweep (round 2): the same computation as sweep_fused_kernel
//     d_m = W'(e_m),   b_m ~ N(tau v d_m, v),   u_m = W b_m      for the NB blocks m that share one basis W
// (identical diagonal blocks of a BD kernel, R/BGGE.R:310-339; merged environment kernels of MDe), W read from HBM
// ONCE, but organised so that no compute warp ever waits at a CTA-wide barrier or does serial work:
//
//   12 compute warps (384 threads, one row per thread and row step; the residual slices e_m and the u accumulators
//      of their rows live in registers, setmaxnreg 152): for column group h they wait full[h], form the 2 x NB partial
//      dots, reduce inside the warp, drop the warp's partials in shared memory and arrive on dotdone[h] -- then, `lag`
//      groups behind, wait bready[h - lag], apply u += W[:, h - lag] b from the same shared-memory copy and free the
//      ring slot.  The exchange latency hides behind the dots of the next `lag` groups.
//   producer warp (one lane): TMA bulk copies (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) of the CTA's row
//      slice of the next group's columns into the ring; never waits for anything but a free slot.
//   pusher warp: waits dotdone[h], adds the 12 warp partials in warp order and sends the CTA's partial dots to every
//      CTA of the cluster (st.async over distributed shared memory + remote mbarrier complete_tx).
//   coefficient warp: waits for the cluster's partials of group h, adds them in rank order, forms b (R/BGGE.R:300-305)
//      and the pieces of sum b^2/s (R/BGGE.R:96), publishes b through bready[h].
//   draw warp: prepares tau*v, 1/s and sqrt(v)*z for the next 64 columns (Philox or the injected tape), double
//      buffered, off everybody's critical path.
//
// A cluster of CS CTAs owns a panel of eigen-columns; CTA `rank` owns the row slice [rank*rs, rank*rs + rs).  With
// NB = 4 and 10000 rows CS = 4 fits (7 rows x 4 members per thread for e and for u), so 36-37 clusters cover the
// chip where the 8-CTA clusters of sweep_fused_kernel covered 120 of 148 SMs.
// Determinism: every sum has a fixed order (rows ascending per thread, fixed shuffle tree, warps in order, ranks in
// order, clusters in order in fused_finalize_kernel).
#include "gibbs.cuh"
#include "kernels.h"
#include "rng.cuh"
#include "sweep_mrhs.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace bgge {
namespace {

constexpr int M_NCW = 12;                  // compute warps
constexpr int M_T = M_NCW * 32;            // compute threads
constexpr int M_THREADS = M_T + 128;       // + helper warpgroup (producer, pusher, coefficient, draw)
constexpr int M_NP = 4;                    // warp-partial ring (>= lag + 1)
constexpr int M_NPX = 16;                  // exchange / coefficient ring (>= ns + lag + 1)
constexpr int M_ZB = 64;                   // columns per draw batch
constexpr int M_MAXNS = MRHS_MAX_NS;       // TMA ring depth limit (groups)
constexpr int M_MAXCS = 8;

__device__ __forceinline__ uint32_t map_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void wait_relaxed(uint64_t* bar, uint32_t parity) {
  int polls = 0;
  while (!mbar_try_wait(bar, parity))
    if (++polls > 4) __nanosleep(40);
}
// pusher / coefficient warps sit on the exchange's critical path: mbarrier.try_wait already suspends the warp in
// hardware for a while, no extra sleep
__device__ __forceinline__ void wait_tight(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// see sweep_fused.cu: P values summed over the warp with P - 1 + (5 - log2 P) shuffles; value idx = (lane * P) >> 5
template <int P>
__device__ __forceinline__ double multi_reduce(double (&a)[P], int lane, int& idx) {
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

struct MSmem {            // carve of the dynamic shared memory (all offsets in doubles from the base)
  double *ring, *wp, *parts, *bsm, *tvb, *isb, *zsb;
  uint64_t *full, *empty, *dotdone, *pbar, *bready, *zready, *zfree;
  uint32_t* tmem;   // tensor-memory base address (TM variant)
};

template <int NB, int G>
__device__ __forceinline__ MSmem carve(unsigned char* base, int ns, size_t slot_doubles, int tvm) {
  constexpr int NG = NB * G;
  MSmem s;
  s.ring = reinterpret_cast<double*>(base);
  s.wp = s.ring + (size_t)ns * slot_doubles;
  s.parts = s.wp + M_NP * M_NCW * NG;
  s.bsm = s.parts + M_NPX * M_MAXCS * NG;
  s.tvb = s.bsm + M_NPX * NG;                      // [tvm][2][ZB]
  s.isb = s.tvb + (size_t)tvm * 2 * M_ZB;          // [2][ZB]
  s.zsb = s.isb + 2 * M_ZB;                        // [NB][2][ZB]
  s.full = reinterpret_cast<uint64_t*>(s.zsb + (size_t)NB * 2 * M_ZB);
  s.empty = s.full + M_MAXNS;
  s.dotdone = s.empty + M_MAXNS;
  s.pbar = s.dotdone + M_NP;
  s.bready = s.pbar + M_NPX;
  s.zready = s.bready + M_NPX;
  s.zfree = s.zready + 2;
  s.tmem = reinterpret_cast<uint32_t*>(s.zfree + 2);
  return s;
}

// ---- tensor memory (TMEM, 256 KB per SM) as a per-thread scratchpad: tcgen05.st / tcgen05.ld in the 32x32b shape move
// N consecutive 32-bit columns of the calling thread's own lane (lane quarter = warp % 4), SASS STTM / LDTM.
template <int NB>
__device__ __forceinline__ void tmem_store(uint32_t taddr, const double (&v)[NB]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  if constexpr (NB == 4) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  } else {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3])
                 : "memory");
  }
}
template <int NB>
__device__ __forceinline__ void tmem_load(uint32_t taddr, uint32_t (&r)[2 * NB]) {
  if constexpr (NB == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ double u2d(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

// G = eigen-columns per exchange group (= per ring slot).  TM = false: e_m and u in registers (KR * NB <= 28).
// TM = true: e_m lives in tensor memory (written once per panel with tcgen05.st, read back one row step at a time with
// tcgen05.ld), only u stays in registers, so twice the rows fit a CTA: 2-CTA clusters hold a 10000-row, 4-member
// panel and 74 clusters cover all 148 SMs.
template <int KR, int NB, int G, bool TM>
__global__ void __launch_bounds__(M_THREADS, 1)
sweep_mrhs_kernel(const FPanel* __restrict__ panels, const int2* __restrict__ cluster_panels,
                  const double* __restrict__ r, const double* __restrict__ u_all, double* __restrict__ upart,
                  long long n, const double* __restrict__ ve, double* __restrict__ vupart, long long vtotal,
                  const Scal* __restrict__ sc, int merged, DrawCfg dc, int ns, int lag) {
  pdl_trigger();
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const uint32_t cs = cluster.num_blocks();
  const int cid = blockIdx.x / cs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NG = NB * G;
  static_assert(G == 1 || G == 2, "one or two columns per group");

  extern __shared__ __align__(128) unsigned char msm[];
  const int2 cp = cluster_panels[cid];
  int max_rs = 0;
  for (int pi = 0; pi < cp.y; ++pi) max_rs = max(max_rs, panels[cp.x + pi].rs);
  const size_t slot_doubles = (size_t)G * max_rs;
  const int tvm = merged ? NB : 1;
  const MSmem s = carve<NB, G>(msm, ns, slot_doubles, tvm);
  BGGE_DEVASSERT(reinterpret_cast<unsigned char*>(s.tmem + 4) <= msm + dynamic_smem_bytes());
  BGGE_DEVASSERT(ns >= 3 && ns <= M_MAXNS && lag >= 1 && lag + 1 <= M_NP && ns + lag + 1 <= M_NPX && cs <= M_MAXCS);
  // unpredicated row steps: the furthest read of the last ring slot stays in front of the mbarriers
  // (last slot, last of its G columns, row step KR - 1, thread M_T - 1)
  BGGE_DEVASSERT(s.ring + (size_t)(ns - 1) * slot_doubles + (size_t)(G - 1) * max_rs + (size_t)KR * M_T <= reinterpret_cast<double*>(s.full) ||
                 cp.y == 0);

  // rows of a slot beyond the CTA's slice are never written by the bulk copies: zero everything once so they are
  // finite (they meet e = 0 in the dots and their u accumulators are never stored)
  for (size_t i = threadIdx.x; i < (size_t)(reinterpret_cast<double*>(s.full) - s.ring); i += M_THREADS) s.ring[i] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the bulk copies (async proxy) overwrite these zeros
  if (threadIdx.x == 0) {
    for (int k = 0; k < ns; ++k) {
      mbar_init(&s.full[k], 1);
      mbar_init(&s.empty[k], M_NCW);
    }
    for (int k = 0; k < M_NP; ++k) mbar_init(&s.dotdone[k], M_NCW);
    for (int k = 0; k < M_NPX; ++k) {
      mbar_init(&s.pbar[k], 1);
      mbar_init(&s.bready[k], 1);
    }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&s.zready[k], 1);
      mbar_init(&s.zfree[k], 1);
    }
    mbar_fence_init();
  }
  if constexpr (TM) {
    if (warp == M_NCW) {   // the producer warp owns the allocation (all 512 columns: one CTA per SM)
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s.tmem)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  cluster.sync();   // every CTA's barriers exist (and its shared memory is zeroed) before anybody sends
  uint32_t tbase = 0;
  if constexpr (TM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tbase = *s.tmem;
  }

  if (warp < M_NCW) {
    // ================================================================== compute warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    pdl_wait();
    const double shift = sc->shift;
    const int tid = threadIdx.x;
    int dslot = 0, aslot = 0;          // ring slot of the next group to dot / to apply
    uint32_t dphase = 0;
    uint32_t dcnt = 0, acnt = 0;       // running group counters (exchange-ring indices derive from them)
    for (int pi = 0; pi < cp.y; ++pi) {
      const FPanel& pn = panels[cp.x + pi];
      const int row0 = (int)rank * pn.rs;
      const int my_rows = max(0, min(pn.rs, pn.rows - row0));
      // this thread's tensor-memory window: lane quarter warp % 4, 2 * KR * NB columns per warp slot (3 slots)
      const uint32_t taddr0 = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * (2 * KR * NB));
      double e[TM ? 1 : KR][NB], uacc[KR][NB];
#pragma unroll
      for (int k = 0; k < KR; ++k) {
        const int row = k * M_T + tid;
        double ek[NB];
#pragma unroll
        for (int m = 0; m < NB; ++m) {
          uacc[k][m] = 0.0;
          double v = 0.0;
          if (m < pn.nb && row < my_rows) {
            if (pn.virt && !pn.direct) {
              v = ve[pn.voffs[m] + row0 + row];
            } else {
              const long long gi = (long long)pn.pos0s[m] + row0 + row;
              v = (r[gi] - shift) + u_all[(size_t)pn.kids[m] * n + gi];
            }
          }
          ek[m] = v;
        }
        if constexpr (TM) {
          tmem_store<NB>(taddr0 + (uint32_t)(k * 2 * NB), ek);
        } else {
#pragma unroll
          for (int m = 0; m < NB; ++m) e[k][m] = ek[m];
        }
      }
      if constexpr (TM) tmem_wait_st();
      // Row steps that reach past the CTA's slice (KR * 384 >= rs) read on into the next column / the next ring slot /
      // the small arrays behind the ring: always finite doubles (everything up to the barriers is zeroed at start and
      // only ever receives W, partial sums and coefficients), and they meet e = 0 in the dots while their u
      // accumulators are never stored.  So the hot loops carry no row predicates.  mrhs_configure() checks that
      // the overshoot stays inside those arrays.
      const int ngroups = (pn.c1 - pn.c0 + G - 1) / G;
      for (int h = 0; h < ngroups + lag; ++h) {
        if (h < ngroups) {
          // ---------------------------------------------------------- dots of group h
          BGGE_DEVASSERT(dslot >= 0 && dslot < ns && (int)(dcnt - acnt) <= lag + 1);
          wait_relaxed(&s.full[dslot], dphase);
          const double* src0 = s.ring + (size_t)dslot * slot_doubles + tid;
          const double* src1 = src0 + (G == 2 ? max_rs : 0);
          if (++dslot == ns) {
            dslot = 0;
            dphase ^= 1u;
          }
          double a[NG];
#pragma unroll
          for (int v = 0; v < NG; ++v) a[v] = 0.0;
          if constexpr (TM) {
            // e comes back from tensor memory two row steps at a time (one wait per pair)
#pragma unroll
            for (int k = 0; k < KR; k += 2) {
              uint32_t w0[2 * NB], w1[2 * NB];
              tmem_load<NB>(taddr0 + (uint32_t)(k * 2 * NB), w0);
              if (k + 1 < KR) tmem_load<NB>(taddr0 + (uint32_t)((k + 1) * 2 * NB), w1);
              const double x00 = src0[k * M_T], x01 = G == 2 ? src1[k * M_T] : 0.0;
              const double x10 = k + 1 < KR ? src0[(k + 1) * M_T] : 0.0, x11 = (G == 2 && k + 1 < KR) ? src1[(k + 1) * M_T] : 0.0;
              tmem_wait_ld();
#pragma unroll
              for (int m = 0; m < NB; ++m) {
                const double e0 = u2d(w0[2 * m], w0[2 * m + 1]);
                a[m * G] = fma(x00, e0, a[m * G]);
                if constexpr (G == 2) a[m * G + 1] = fma(x01, e0, a[m * G + 1]);
              }
              if (k + 1 < KR) {
#pragma unroll
                for (int m = 0; m < NB; ++m) {
                  const double e1 = u2d(w1[2 * m], w1[2 * m + 1]);
                  a[m * G] = fma(x10, e1, a[m * G]);
                  if constexpr (G == 2) a[m * G + 1] = fma(x11, e1, a[m * G + 1]);
                }
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < KR; ++k) {
              const double x0 = src0[k * M_T], x1 = G == 2 ? src1[k * M_T] : 0.0;
#pragma unroll
              for (int m = 0; m < NB; ++m) {
                a[m * G] = fma(x0, e[k][m], a[m * G]);
                if constexpr (G == 2) a[m * G + 1] = fma(x1, e[k][m], a[m * G + 1]);
              }
            }
          }
          int vi;
          const double tot = multi_reduce<NG>(a, lane, vi);
          double* wpd = s.wp + ((size_t)(dcnt & (M_NP - 1)) * M_NCW + warp) * NG;
          if ((lane & (32 / NG - 1)) == 0) wpd[vi] = tot;
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.dotdone[dcnt & (M_NP - 1)]);
          ++dcnt;
        }
        if (h >= lag) {
          // ---------------------------------------------------------- u += W[:, group h - lag] b
          const uint32_t px = acnt & (M_NPX - 1);
          BGGE_DEVASSERT(aslot >= 0 && aslot < ns && acnt < dcnt + (h >= ngroups ? lag : 0) + 1);
          wait_relaxed(&s.bready[px], (acnt / M_NPX) & 1u);
          double b[NG];
          const double* bs = s.bsm + (size_t)px * NG;
#pragma unroll
          for (int v = 0; v < NG; v += 2) {
            const double2 t = *reinterpret_cast<const double2*>(bs + v);
            b[v] = t.x;
            b[v + 1] = t.y;
          }
          const double* src0 = s.ring + (size_t)aslot * slot_doubles + tid;
          const double* src1 = src0 + (G == 2 ? max_rs : 0);
#pragma unroll
          for (int k = 0; k < KR; ++k) {
            const double x0 = src0[k * M_T];
            if constexpr (G == 2) {
              const double x1 = src1[k * M_T];
#pragma unroll
              for (int m = 0; m < NB; ++m) uacc[k][m] = fma(x1, b[m * G + 1], fma(x0, b[m * G], uacc[k][m]));
            } else {
#pragma unroll
              for (int m = 0; m < NB; ++m) uacc[k][m] = fma(x0, b[m], uacc[k][m]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.empty[aslot]);
          if (++aslot == ns) aslot = 0;
          ++acnt;
        }
      }
      // this cluster's contribution to u over the block's rows
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        if (m < pn.nb) {
          double* up = pn.virt ? vupart + (size_t)pn.slot * vtotal + pn.voffs[m] + row0
                               : upart + (size_t)pn.slot * n + pn.pos0s[m] + row0;
#pragma unroll
          for (int k = 0; k < KR; ++k) {
            const int row = k * M_T + tid;
            if (row < my_rows) up[row] = uacc[k][m];
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int hw = warp - M_NCW;
    if (hw == 0) {
      // ================================================================ producer
      if (lane == 0) {
        int pslot = 0;
        uint32_t pphase = 0;
        for (int pi = 0; pi < cp.y; ++pi) {
          const FPanel& pn = panels[cp.x + pi];
          const int row0 = (int)rank * pn.rs;
          const int my_rows = max(0, min(pn.rs, pn.rows - row0));
          const uint32_t slice_bytes = (uint32_t)((my_rows + 1) & ~1) * (uint32_t)sizeof(double);
          const double* Ubase = pn.U + row0;
          for (int c0 = pn.c0; c0 < pn.c1; c0 += G) {
            const int gc = min(G, pn.c1 - c0);
            BGGE_DEVASSERT(pslot < ns && slice_bytes % 16 == 0 && (size_t)my_rows <= (size_t)max_rs &&
                           ((size_t)(Ubase + (long long)c0 * pn.ldu) & 15) == 0);
            wait_relaxed(&s.empty[pslot], pphase ^ 1u);
            mbar_expect_tx(&s.full[pslot], slice_bytes * (uint32_t)gc);
            if (slice_bytes > 0) {
              double* dst = s.ring + (size_t)pslot * slot_doubles;
              for (int g = 0; g < gc; ++g)
                bulk_g2s(dst + (size_t)g * max_rs, Ubase + (long long)(c0 + g) * pn.ldu, slice_bytes, &s.full[pslot]);
            }
            if (++pslot == ns) {
              pslot = 0;
              pphase ^= 1u;
            }
          }
        }
      }
    } else if (hw == 1) {
      // ================================================================ pusher
      uint32_t cnt = 0;
      const int v = lane % NG;
      const uint32_t q0 = (uint32_t)(lane / NG), qstep = 32u / NG;
      const int m = v / G, gq = v % G;
      for (int pi = 0; pi < cp.y; ++pi) {
        const FPanel& pn = panels[cp.x + pi];
        for (int c0 = pn.c0; c0 < pn.c1; c0 += G, ++cnt) {
          const int gc = min(G, pn.c1 - c0);
          const uint32_t pw = cnt & (M_NP - 1), px = cnt & (M_NPX - 1);
          wait_tight(&s.dotdone[pw], (cnt / M_NP) & 1u);
          if (lane == 0) mbar_expect_tx(&s.pbar[px], cs * (uint32_t)(gc * pn.nb) * (uint32_t)sizeof(double));
          __syncwarp();
          if (m < pn.nb && gq < gc) {
            const double* wps = s.wp + (size_t)pw * M_NCW * NG + v;
            double p = 0.0;
#pragma unroll
            for (int w = 0; w < M_NCW; ++w) p += wps[w * NG];                      // warp order
            const uint32_t dst = smem_u32(s.parts + ((size_t)px * M_MAXCS + rank) * NG + v);
            const uint32_t bar = smem_u32(&s.pbar[px]);
            for (uint32_t q = q0; q < cs; q += qstep) st_async64(map_rank(dst, q), p, map_rank(bar, q));
          }
        }
      }
    } else if (hw == 2) {
      // ================================================================ coefficients
      pdl_wait();   // (its only global write, the sum of b^2/s, must not overtake the previous launch's reads)
      uint32_t cnt = 0, batch = 0;
      const int v = lane, m = v / G, gq = v % G;
      double zq = 0.0;
      for (int pi = 0; pi < cp.y; ++pi) {
        const FPanel& pn = panels[cp.x + pi];
        const int ncols = pn.c1 - pn.c0;
        for (int rel0 = 0; rel0 < ncols; rel0 += G, ++cnt) {
          const int gc = min(G, ncols - rel0);
          if (rel0 % M_ZB == 0) wait_relaxed(&s.zready[batch & 1u], (batch >> 1) & 1u);
          const uint32_t px = cnt & (M_NPX - 1);
          wait_tight(&s.pbar[px], (cnt / M_NPX) & 1u);
          if (v < NG) {
            double bcoef = 0.0;
            if (m < pn.nb && gq < gc) {
              const double* pp = s.parts + (size_t)px * M_MAXCS * NG + v;
              double d = 0.0;
              for (uint32_t q = 0; q < cs; ++q) d += pp[(size_t)q * NG];          // rank order
              const int zb = (int)(batch & 1u) * M_ZB + (rel0 + gq) % M_ZB;
              BGGE_DEVASSERT(zb >= 0 && zb < 2 * M_ZB && isfinite(d));
              bcoef = s.tvb[(size_t)(merged ? m : 0) * 2 * M_ZB + zb] * d + s.zsb[(size_t)m * 2 * M_ZB + zb];   // R/BGGE.R:303, 89-92
              zq = fma(bcoef * bcoef, s.isb[zb], zq);                              // R/BGGE.R:96
            }
            s.bsm[(size_t)px * NG + v] = bcoef;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.bready[px]);
          if ((rel0 + G) % M_ZB == 0 || rel0 + G >= ncols) {   // last group of the draw batch
            if (lane == 0) mbar_arrive(&s.zfree[batch & 1u]);
            ++batch;
          }
        }
      }
      // all panels of a launch belong to one kernel, or to one merged unit: a cluster without panels finds the
      // destination of its (zero) sum in panel 0
      const FPanel& p0 = panels[cp.y > 0 ? cp.x : 0];
      if (merged) {
        for (int mm = 0; mm < p0.nb; ++mm) {
          const double zm = warp_sum(v < NG && m == mm ? zq : 0.0);
          if (rank == 0 && lane == 0) p0.bps[mm][cid] = zm;
        }
      } else {
        const double zt = warp_sum(v < NG ? zq : 0.0);
        if (rank == 0 && lane == 0) p0.bps[0][cid] = zt;
      }
    } else {
      // ================================================================ draw tables
      pdl_wait();
      const double tau = sc->tau;
      const long long it = sc->it;
      uint32_t batch = 0;
      for (int pi = 0; pi < cp.y; ++pi) {
        const FPanel& pn = panels[cp.x + pi];
        const int ncols = pn.c1 - pn.c0;
        for (int rel0 = 0; rel0 < ncols; rel0 += M_ZB, ++batch) {
          wait_relaxed(&s.zfree[batch & 1u], ((batch >> 1) & 1u) ^ 1u);
          for (int t = lane; t < M_ZB; t += 32) {
            const int rel = rel0 + t;
            if (rel < ncols) {
              const int zb = (int)(batch & 1u) * M_ZB + t;
              const double sv = pn.ss[0][pn.c0 + rel];                             // members share the matrix, hence s
              s.isb[zb] = 1.0 / sv;
              double sq = 0.0;
              for (int m = 0; m < pn.nb; ++m) {
                if (m == 0 || merged) {                                            // one sigb per kernel
                  const double lam = sc->sigb[pn.kids[m]];
                  const double vari = sv * lam / (1.0 + sv * lam * tau);           // R/BGGE.R:302
                  s.tvb[(size_t)m * 2 * M_ZB + zb] = tau * vari;
                  sq = sqrt(vari);
                }
                const long long cm = (long long)pn.col_offs[m] + pn.c0 + rel;
                const double z = draw_normal(dc, it, (uint32_t)pn.kids[m], (unsigned long long)cm,
                                             dc.z_init + (it - 1) * dc.z_per_it + pn.zoffs[m] + cm);
                s.zsb[(size_t)m * 2 * M_ZB + zb] = sq * z;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.zready[batch & 1u]);
        }
      }
    }
  }
  if constexpr (TM) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();      // nobody leaves while a peer may still write into its shared memory
  if constexpr (TM) {
    if (warp == M_NCW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
  }
}

size_t fixed_bytes(int nb, int g, bool merged) {
  const size_t ng = (size_t)nb * g;
  const size_t tvm = merged ? (size_t)nb : 1;
  return ((size_t)M_NP * M_NCW * ng + (size_t)M_NPX * M_MAXCS * ng + (size_t)M_NPX * ng + (2 * tvm + 2 + 2 * (size_t)nb) * M_ZB) *
             sizeof(double) +
         (2 * M_MAXNS + M_NP + 2 * M_NPX + 4) * sizeof(uint64_t) + 16 + 128;
}
size_t barrier_bytes() { return (2 * M_MAXNS + M_NP + 2 * M_NPX + 4) * sizeof(uint64_t) + 16 + 128; }

typedef void (*MrhsFn)(const FPanel*, const int2*, const double*, const double*, double*, long long, const double*, double*,
                       long long, const Scal*, int, DrawCfg, int, int);

struct Inst {
  int nb, kr, g, tm;
  MrhsFn fn;
};
const Inst kInst[] = {
    // e and u in registers, two columns per group
    {4, 2, 2, 0, sweep_mrhs_kernel<2, 4, 2, false>},   {4, 3, 2, 0, sweep_mrhs_kernel<3, 4, 2, false>},
    {4, 4, 2, 0, sweep_mrhs_kernel<4, 4, 2, false>},   {4, 5, 2, 0, sweep_mrhs_kernel<5, 4, 2, false>},
    {4, 6, 2, 0, sweep_mrhs_kernel<6, 4, 2, false>},   {4, 7, 2, 0, sweep_mrhs_kernel<7, 4, 2, false>},
    {2, 4, 2, 0, sweep_mrhs_kernel<4, 2, 2, false>},   {2, 6, 2, 0, sweep_mrhs_kernel<6, 2, 2, false>},
    {2, 8, 2, 0, sweep_mrhs_kernel<8, 2, 2, false>},   {2, 10, 2, 0, sweep_mrhs_kernel<10, 2, 2, false>},
    {2, 12, 2, 0, sweep_mrhs_kernel<12, 2, 2, false>}, {2, 14, 2, 0, sweep_mrhs_kernel<14, 2, 2, false>},
    // e in tensor memory, one column per group: four members at twice the rows per CTA (2-CTA clusters at 10000 rows)
    {4, 8, 1, 1, sweep_mrhs_kernel<8, 4, 1, true>},    {4, 10, 1, 1, sweep_mrhs_kernel<10, 4, 1, true>},
    {4, 12, 1, 1, sweep_mrhs_kernel<12, 4, 1, true>},  {4, 14, 1, 1, sweep_mrhs_kernel<14, 4, 1, true>},
};

const Inst* find_inst(int nb, int kr_need, int tm) {
  for (const Inst& i : kInst)
    if (i.nb == nb && i.tm == tm && i.kr >= kr_need) return &i;
  return nullptr;
}

int tmem_mode() {   // BGGE_MRHS_TMEM: 0 never, 1 when it lets a smaller cluster hold the panel (default), 2 whenever it fits
  if (const char* ev = getenv("BGGE_MRHS_TMEM")) return atoi(ev);
  return MRHS_TMEM_DEFAULT;
}

}  // namespace

int64_t mrhs_min_rows() {
  if (const char* ev = getenv("BGGE_MRHS_MIN_ROWS")) return atoll(ev);
  return 2048;
}

// Picks cluster size, rows per thread, ring depth and the number of co-resident clusters for a multi-RHS step
// whose longest streamed block has max_rows rows.  cfg->ok stays false when the shape does not fit.
int mrhs_configure(int64_t max_rows, int nb, bool merged, MrhsCfg* cfg) {
  cfg->ok = false;
  if (nb != 2 && nb != 4) return OK;
  int dev = 0, optin = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  BGGE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int cs_min = 1;
  if (const char* ev = getenv("BGGE_MRHS_CS")) {   // tuning hook
    const int v = atoi(ev);
    if (v == 1 || v == 2 || v == 4 || v == 8) cs_min = v;
  }
  const int tmode = tmem_mode();
  // Candidate order.  Default (mode 1): the register variant while it fits 1- or 2-CTA clusters (C3: 3000 rows on
  // cs = 2 measured faster than the tensor-memory variant on cs = 1), then the tensor-memory variant at its smallest
  // cluster (C5: 10000 rows x 4 members on cs = 2 instead of cs = 4), then the register variant on larger clusters.
  // Mode 0: register variant only.  Mode 2: tensor-memory variant first.
  struct Cand { int tm, cs_lo, cs_hi; };
  const Cand only_reg[3] = {{0, 1, M_MAXCS}, {0, 1, 0}, {0, 1, 0}};
  const Cand tm_first[3] = {{1, 1, M_MAXCS}, {0, 1, M_MAXCS}, {0, 1, 0}};
  const Cand mixed[3] = {{0, 1, 2}, {1, 1, M_MAXCS}, {0, 4, M_MAXCS}};
  const Cand* cands = tmode == 0 ? only_reg : tmode == 2 ? tm_first : mixed;
  for (int ci = 0; ci < 3; ++ci) {
    const Cand& cd = cands[ci];
    for (int cs = std::max(cs_min, cd.cs_lo); cs <= cd.cs_hi; cs *= 2) {
      const int64_t rs = round_up((max_rows + cs - 1) / cs, 16);
      const int kr_need = (int)((rs + M_T - 1) / M_T);
      // more than 7 row steps per column on one CTA: two CTAs with half the steps each measured faster
      // (C3 unit of two members, 3000 rows: 9730 -> 10060 it/s), so leave cs = 1 to the shorter blocks
      if (!cd.tm && cs == 1 && kr_need > 7 && cd.cs_hi >= 2 && !getenv("BGGE_MRHS_CS")) continue;
      const Inst* inst = find_inst(nb, kr_need, cd.tm);
      if (!inst) continue;
      const size_t slot_bytes = (size_t)inst->g * rs * sizeof(double);
      const size_t fixed = fixed_bytes(nb, inst->g, merged);
      // unpredicated row steps may read (kr * 384 - rs) doubles past the last ring slot: they must stay in the zeroed
      // arrays in front of the mbarriers
      const size_t tail = fixed - barrier_bytes();
      if ((size_t)(inst->kr * (int64_t)M_T - rs) * sizeof(double) > tail) continue;
      int ns = (int)std::min<size_t>(M_MAXNS, ((size_t)optin - fixed - 1024) / slot_bytes);
      if (ns < 3) continue;
      int lag = 2;     // groups between the dots and the axpy of a group
      if (const char* ev = getenv("BGGE_MRHS_LAG")) lag = std::max(1, std::min(3, atoi(ev)));
      lag = std::min(lag, ns - 2);
      cfg->cs = cs;
      cfg->kr = inst->kr;
      cfg->g = inst->g;
      cfg->tmem = inst->tm;
      cfg->ns = ns;
      cfg->lag = lag;
      cfg->smem = (size_t)ns * slot_bytes + fixed;
      BGGE_CUDA(cudaFuncSetAttribute(inst->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
      cudaLaunchConfig_t lc{};
      lc.gridDim = dim3((unsigned)(cs * 64), 1, 1);
      lc.blockDim = dim3(M_THREADS, 1, 1);
      lc.dynamicSmemBytes = cfg->smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)cs;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      lc.attrs = attr;
      lc.numAttrs = 1;
      int ncl = 0;
      BGGE_CUDA(cudaOccupancyMaxActiveClusters(&ncl, inst->fn, &lc));
      if (ncl < 1) continue;
      cfg->clusters = ncl;
      cfg->ok = true;
      return OK;
    }
  }
  return OK;
}

int mrhs_launch(Gibbs* g, HostKernel& hk) {
  const Inst* inst = find_inst(hk.fused_nb, hk.fused_kr, hk.fused_tmem);
  BGGE_REQUIRE(inst && inst->kr == hk.fused_kr && inst->g == hk.fused_g, "multi-RHS sweep: no kernel instantiation for kr=%d nb=%d", hk.fused_kr,
               hk.fused_nb);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(hk.fused_clusters * hk.fused_cs), 1, 1);
  cfg.blockDim = dim3(M_THREADS, 1, 1);
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
  BGGE_CUDA(cudaLaunchKernelEx(&cfg, inst->fn, (const FPanel*)hk.fpanels.p, (const int2*)hk.fcluster.p, (const double*)g->r.p,
                               (const double*)g->u.p, g->upart.p, (long long)g->n, (const double*)hk.ve.p, g->vupart.p,
                               (long long)hk.vtotal, (const Scal*)g->scal.p, hk.merged ? 1 : 0, g->draw, hk.fused_ns,
                               hk.fused_lag));
  return OK;
}

}  // namespace bgge

// FP64 tensor-core Gram contraction  S = X X^T  (the contraction under getK's GB and GK
// kernels; replaces tcrossprod(X) at R/getK.R:136 and the O(L^2 p) part of dist(X) at :142).
//
// X is the padded device copy of the marker matrix: column-major, leading dimension ldx
// (multiple of 128, rows >= L are zero), p_pad columns (multiple of 16, columns >= p zero).
// Work unit: one 128x128 tile of the lower triangle of S, full depth p.  Persistent CTAs walk
// a host-built tile list.  Per CTA: 16-marker slabs of the two row panels are streamed into a
// 4-stage shared-memory ring by the TMA unit (cp.async.bulk + mbarrier complete_tx; SASS
// UBLKCP) issued by a dedicated producer warp whose warpgroup gives its registers to the consumers
// (setmaxnreg 40 / 232); 8 consumer warps each own a 64x32 sub-tile and issue mma.sync.m8n8k4.f64
// (SASS DMMA.8x8x4; tcgen05 has no f64 kind).  [r01: a consumer warp doubling as producer spent
// ~40 % of its time in the 32-copy issue loop and stalled the whole CTA -- profiles/r01_gram_*.]  Shared rows are padded to
// 132 doubles so the (k = lane%4, m = lane/4) fragment pattern hits 16 distinct 8-byte banks
// per half-warp.
#include "common.cuh"
#include "kernels.h"

namespace bgge {
namespace {

constexpr int TM = 128;          // tile rows
constexpr int TN = 128;          // tile cols
constexpr int BK = 16;           // markers per stage
constexpr int SLD = 132;         // padded smem leading dimension (doubles); 132 % 16 == 4
constexpr int STAGES = 4;
constexpr int NCONS = 8;         // consumer warps (2 along M x 4 along N)
constexpr int NTHREADS = 384;    // 2 consumer warpgroups + 1 producer warpgroup (setmaxnreg needs groups of 4 warps)
constexpr int STAGE_DOUBLES = 2 * BK * SLD;
constexpr uint32_t STAGE_TX_BYTES = 2u * BK * TM * sizeof(double);
constexpr size_t GRAM_SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double) + 2 * STAGES * sizeof(uint64_t);

__global__ void __launch_bounds__(NTHREADS, 1)
gram_dmma_kernel(const double* __restrict__ X, int64_t ldx, int kblocks, const int2* __restrict__ tiles,
                 int ntiles, double* __restrict__ S, int64_t lds, int64_t L, int64_t row_off,
                 double divisor, int mirror) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)STAGES * STAGE_DOUBLES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCONS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= NCONS) {
    // ------------------------------------------------------------ producer warpgroup (warps 8..11)
    // Hands its registers to the consumers; only warp 8 works: lane l streams marker row (l & 15)
    // of panel (l >> 4) for every k-block of every tile of this CTA, STAGES blocks ahead.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp != NCONS) return;
    const int row = lane & 15;
    const int op = lane >> 4;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int2 tc = tiles[t];
      const double* src = X + (int64_t)(op ? tc.y : tc.x) * TM + (int64_t)row * ldx;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int stage = it % STAGES;
        mbar_wait(&empty[stage], ((it / STAGES) & 1u) ^ 1u);
        if (lane == 0) mbar_expect_tx(&full[stage], STAGE_TX_BYTES);
        __syncwarp();
        double* dst = sm + (size_t)stage * STAGE_DOUBLES + (size_t)op * BK * SLD + (size_t)row * SLD;
        bulk_g2s(dst, src + (int64_t)kb * BK * ldx, TM * sizeof(double), &full[stage]);
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumer warps 0..7
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int wm = warp & 1;     // 64-row half
  const int wn = warp >> 1;    // 32-col quarter
  const int fr = lane >> 2;    // fragment row (m or n) 0..7
  const int fk = lane & 3;     // fragment k 0..3
  uint32_t it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int2 tc = tiles[t];
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kb = 0; kb < kblocks; ++kb, ++it) {
      const int stage = it % STAGES;
      mbar_wait(&full[stage], (it / STAGES) & 1u);
      const double* As = sm + (size_t)stage * STAGE_DOUBLES + wm * 64 + fr;
      const double* Bs = sm + (size_t)stage * STAGE_DOUBLES + BK * SLD + wn * 32 + fr;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        const int k = kk * 4 + fk;
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = As[k * SLD + i * 8];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k * SLD + j * 8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }

    // epilogue: C fragment (row = lane/4, cols 2*(lane%4)+{0,1}); scale and store
    const int64_t gi0 = (int64_t)tc.x * TM + wm * 64 + fr;
    const int64_t gj0 = (int64_t)tc.y * TN + wn * 32 + 2 * fk;
    const bool offdiag = mirror && (tc.x != tc.y);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t gi = gi0 + i * 8;
      if (gi >= L) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t gj = gj0 + j * 8 + h;
          if (gj >= L) continue;
          const double v = acc[i][j][h] / divisor;
          S[(gi - row_off) + gj * lds] = v;
          if (offdiag) S[gj + gi * lds] = v;
        }
      }
    }
  }
}

}  // namespace

int gram_launch(const double* dX, int64_t ldx, int64_t L, int64_t p_pad, const int2* d_tiles, int ntiles,
                double* dS, int64_t lds, int64_t row_off, double divisor, int mirror, cudaStream_t st) {
  BGGE_REQUIRE(ldx % TM == 0 && ldx >= round_up(L, TM), "gram: ldx=%lld must be a multiple of 128 >= L", (long long)ldx);
  BGGE_REQUIRE(p_pad % BK == 0 && p_pad > 0, "gram: padded marker count %lld must be a positive multiple of 16", (long long)p_pad);
  BGGE_REQUIRE(!(mirror && row_off != 0), "gram: mirror store needs the full square output");
  if (ntiles == 0) return OK;
  static bool attr_set = false;
  if (!attr_set) {
    BGGE_CUDA(cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM));
    attr_set = true;
  }
  int grid = ntiles < sm_count() ? ntiles : sm_count();
  gram_dmma_kernel<<<grid, NTHREADS, GRAM_SMEM, st>>>(dX, ldx, (int)(p_pad / BK), d_tiles, ntiles, dS, lds, L,
                                                             row_off, divisor, mirror);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

}  // namespace bgge

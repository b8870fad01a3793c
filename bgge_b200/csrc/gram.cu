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
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace bgge {
namespace {

struct GramItem {                // one work item: tile (ti, tj) of the lower triangle, k-blocks [kb0, kb1)
  int ti, tj, kb0, kb1, split;
};

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
gram_dmma_kernel(const double* __restrict__ X, int64_t ldx, const GramItem* __restrict__ tiles,
                 int ntiles, double* __restrict__ S, double* __restrict__ Spart, int64_t part_stride, int64_t pld,
                 int64_t prow_off, int64_t prow_thr, int64_t prow_off2, int64_t lds, int64_t L, int64_t row_off,
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
      const GramItem tc = tiles[t];
      const double* src = X + (int64_t)(op ? tc.tj : tc.ti) * TM + (int64_t)row * ldx;
      for (int kb = tc.kb0; kb < tc.kb1; ++kb, ++it) {
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
    const GramItem tc = tiles[t];
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kb = tc.kb0; kb < tc.kb1; ++kb, ++it) {
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

    // epilogue: C fragment (row = lane/4, cols 2*(lane%4)+{0,1}).  Unsplit tiles are scaled and stored
    // (with the mirror); K-split tiles store their raw partial, gram_reduce_kernel finishes them.
    const int64_t gi0 = (int64_t)tc.ti * TM + wm * 64 + fr;
    const int64_t gj0 = (int64_t)tc.tj * TN + wn * 32 + 2 * fk;
    // mirror: 0 = the tile itself (lower triangle), 1 = tile + transposed tile, 2 = transposed tile only (row-panel
    // builds store S[:, panel rows], which is contiguous in column-major memory: one in-place exchange, no staging)
    const bool direct = Spart == nullptr;
    const bool lower = !direct || mirror != 2;
    const bool upper = direct && (mirror == 2 || (mirror == 1 && tc.ti != tc.tj));
    double* out = direct ? S : Spart + (int64_t)tc.split * part_stride;
    // K-split partials live in a compact workspace over the rows of this launch's (one or two) tile-row panels
    const int64_t old = direct ? lds : pld, oro = direct ? row_off : ((int64_t)tc.ti * TM >= prow_thr ? prow_off2 : prow_off);
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
          const double v = direct ? acc[i][j][h] / divisor : acc[i][j][h];
          if (lower) out[(gi - oro) + gj * old] = v;
          if (upper) S[gj + gi * lds] = v;
        }
      }
    }
  }
}

// sums the K-split partials of every tile in fixed order, scales, stores (+ mirror)
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const GramItem* __restrict__ tiles, int nsplit, const double* __restrict__ Spart,
                   int64_t part_stride, int64_t pld, int64_t prow_off, int64_t prow_thr, int64_t prow_off2,
                   double* __restrict__ S, int64_t lds, int64_t L, int64_t row_off, double divisor, int mirror) {
  const GramItem tc = tiles[blockIdx.x];   // the split-0 item of each tile
  const int64_t i = (int64_t)tc.ti * TM + (threadIdx.x & 127);
  if (i >= L) return;
  for (int c = threadIdx.x >> 7; c < TN; c += 2) {
    const int64_t jx = (int64_t)tc.tj * TN + c;
    if (jx >= L) break;
    const int64_t poff = (i - (i >= prow_thr ? prow_off2 : prow_off)) + jx * pld;
    double v = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) v += Spart[(int64_t)sp * part_stride + poff];
    v /= divisor;
    if (mirror != 2) S[(i - row_off) + jx * lds] = v;
    if (mirror == 2 || (mirror == 1 && tc.ti != tc.tj)) S[jx + i * lds] = v;
  }
}

}  // namespace

// Lower-triangle tiles of tile rows [tr0, tr1).  When the tile count leaves a poor last wave on the
// persistent grid (mid-size L), the marker dimension is split so that the items spread evenly; partials go
// to a workspace and are summed in fixed order by gram_reduce_kernel (deterministic).
int gram_launch(const double* dX, int64_t ldx, int64_t L, int64_t p_pad, int64_t tr0, int64_t tr1, double* dS,
                int64_t lds, int64_t row_off, double divisor, int mirror, cudaStream_t st, int split_on_panel,
                int64_t tr0b, int64_t tr1b) {
  BGGE_REQUIRE(ldx % TM == 0 && ldx >= round_up(L, TM), "gram: ldx=%lld must be a multiple of 128 >= L", (long long)ldx);
  BGGE_REQUIRE(p_pad % BK == 0 && p_pad > 0, "gram: padded marker count %lld must be a positive multiple of 16", (long long)p_pad);
  BGGE_REQUIRE(!(mirror && row_off != 0), "gram: mirror store needs the full square output");
  BGGE_REQUIRE(mirror >= 0 && mirror <= 2, "gram: store mode %d", mirror);
  std::vector<int2> tl;
  // (a second tile-row range [tr0b, tr1b) lets a rank build both of its row panels with one launch)
  for (int64_t ti = tr1b - 1; ti >= tr0b; --ti)
    for (int64_t tj = 0; tj <= ti; ++tj) tl.push_back(make_int2((int)ti, (int)tj));
  for (int64_t ti = tr1 - 1; ti >= tr0; --ti)   // longest tile rows first
    for (int64_t tj = 0; tj <= ti; ++tj) tl.push_back(make_int2((int)ti, (int)tj));
  const int64_t ntiles = (int64_t)tl.size();
  if (ntiles == 0) return OK;
  const int sms = sm_count();
  const int kblocks = (int)(p_pad / BK);
  int nsplit = 1;
  {
    // the split is a function of (L, p) only -- computed on the full triangle, not on this panel -- so that
    // a row-panel build (multi-GPU) stays bit-identical to the one-shot build
    // (split_on_panel: the collective row-panel build balances its own tile list instead)
    const int64_t Tfull = (L + TM - 1) / TM;
    const int64_t nfull = split_on_panel ? ntiles : Tfull * (Tfull + 1) / 2;
    auto cost = [&](int sp) {
      const double waves = (double)((nfull * sp + sms - 1) / sms);
      return waves * ((double)kblocks / sp + 8.0) + (sp > 1 ? 0.03 * sp * kblocks / std::max<int64_t>(1, nfull / sms + 1) : 0.0);
    };
    double best = cost(1);
    for (int sp = 2; sp <= 16 && kblocks / sp >= 64; ++sp)
      if (cost(sp) < (split_on_panel ? 0.98 : 0.95) * best && cost(sp) < cost(nsplit)) nsplit = sp;
  }
  if (const char* ev = getenv("BGGE_GRAM_SPLIT")) {
    const int v = atoi(ev);
    if (v >= 1 && v <= 16) nsplit = v;
  }
  std::vector<GramItem> items;
  for (int sp = 0; sp < nsplit; ++sp)
    for (auto& t : tl) {
      GramItem g{};
      g.ti = t.x;
      g.tj = t.y;
      g.kb0 = (int)((int64_t)kblocks * sp / nsplit);
      g.kb1 = (int)((int64_t)kblocks * (sp + 1) / nsplit);
      g.split = sp;
      items.push_back(g);
    }
  DevBuf<GramItem> d_items;
  DevBuf<double> part;
  BGGE_CUDA(d_items.alloc(items.size()));
  BGGE_CUDA(cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(GramItem), cudaMemcpyHostToDevice, st));
  // K-split partials: compact workspace over the rows of this launch's panels -- panel A = the lower tile-row range,
  // panel B (if any) right behind it: workspace row = i - prow_off for i < prow_thr, else i - prow_off2
  int64_t a0 = tr0, a1 = tr1, b0 = tr0b, b1 = tr1b;
  if (b1 > b0 && b0 < a0) {
    std::swap(a0, b0);
    std::swap(a1, b1);
  }
  const bool two = b1 > b0;
  BGGE_REQUIRE(!two || a1 <= b0, "gram: the two tile-row panels overlap");
  const int64_t rowsA = (a1 - a0) * TM, rowsB = two ? (b1 - b0) * TM : 0;
  const int64_t prow_off = a0 * TM, prow_thr = two ? b0 * TM : (int64_t)1 << 60, prow_off2 = two ? b0 * TM - rowsA : 0;
  const int64_t pld = rowsA + rowsB;
  const int64_t part_stride = pld * std::min<int64_t>(L, (two ? b1 : a1) * TM);
  if (nsplit > 1) BGGE_CUDA(part.alloc((size_t)part_stride * nsplit));
  static bool attr_set[64] = {false};   // function attributes are per device
  int dev = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    BGGE_CUDA(cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int nitems = (int)items.size();
  const int grid = nitems < sms ? nitems : sms;
  gram_dmma_kernel<<<grid, NTHREADS, GRAM_SMEM, st>>>(dX, ldx, d_items.p, nitems, dS, nsplit > 1 ? part.p : nullptr,
                                                       part_stride, pld, prow_off, prow_thr, prow_off2, lds, L, row_off,
                                                       divisor, mirror);
  BGGE_CUDA(cudaGetLastError());
  if (nsplit > 1) {
    gram_reduce_kernel<<<(unsigned)ntiles, 256, 0, st>>>(d_items.p, nsplit, part.p, part_stride, pld, prow_off, prow_thr,
                                                        prow_off2, dS, lds, L, row_off, divisor, mirror);
    BGGE_CUDA(cudaGetLastError());
  }
  BGGE_CUDA(cudaStreamSynchronize(st));   // the item list and the workspace die here
  return OK;
}

}  // namespace bgge

// Skinny FP64 tensor-core GEMM for batched chains / folds (north_star (3): "multiple chains or
// cross-validation folds batch the GEMVs into skinny GEMMs"):
//     C[m][n] = sum_{k in [k0,k1)} A[m + k*lda] * B[n + k*ldb]        (C row-major, ldc)
// i.e. A is "M-fast" (column-major M x K), B is "N-fast" (row-major K x N), the same operand form as the
// Gram kernel (gram.cu), so   d = U'E  uses A = t(U)  and  u = U b  uses A = U  unchanged.
// Work items = (M tile of 128, N tile of TN, K split); every item writes its own partial C (no atomics,
// the consumers sum the K splits in fixed order).  Per CTA: TMA producer warp (cp.async.bulk rows into a
// padded 4-stage ring), 8 consumer warps x 16 rows x TN columns of mma.sync.m8n8k4.f64 (DMMA.8x8x4).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace bgge {
namespace {

constexpr int GK = 16;           // K slab
constexpr int GSTAGES = 4;
constexpr int GCONS = 8;
constexpr int GTHREADS = 384;    // 2 consumer warpgroups + producer warpgroup

// MI = 8-row fragments per consumer warp: 2 -> 128-row CTA tiles, 4 -> 256-row tiles.  Narrow N tiles (<= 48 chains)
// use MI = 4: 20 DMMA per 9 fragment loads instead of 10 per 7 (r02 ncu at TN = 40, MI = 2: DMMA pipe 76 %, the rest
// fixed-latency waits between DMMAs).
template <int TN, int MI>
struct GemmSmem {
  static constexpr int GM = 64 * MI;       // M tile
  static constexpr int SLD_A = GM + 4;     // padded pitch (== 4 mod 16: conflict-free DMMA fragments)
  static constexpr int SLD_B = TN + 4;     // pitch of a B slab row (a contiguous slab brings its own pitch ldb <= TN + 4)
  static constexpr int STAGE_DOUBLES = GK * SLD_A + GK * SLD_B;
  static constexpr size_t BYTES = (size_t)GSTAGES * STAGE_DOUBLES * sizeof(double) + 2 * GSTAGES * sizeof(uint64_t);
};

template <int TN, int MI>
__global__ void __launch_bounds__(GTHREADS, 1)
dgemm_dmma_kernel(const GemmItem* __restrict__ items, int nitems, long long* __restrict__ advance) {
  using SM = GemmSmem<TN, MI>;
  constexpr int GM = SM::GM, GSLD_A = SM::SLD_A;
  constexpr long long ABLK = (long long)GK * GSLD_A;   // doubles of one blocked A chunk == one A ring stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)GSTAGES * SM::STAGE_DOUBLES);
  uint64_t* empty = full + GSTAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // stale shared memory must be finite: clipped tail rows are never loaded and multiply zero operands
  for (int i = threadIdx.x; i < GSTAGES * SM::STAGE_DOUBLES; i += GTHREADS) sm[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GCONS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  // programmatic dependent launch: everything above (shared-memory clearing, barrier set-up) overlaps the tail of the
  // previous kernel; the operands are only touched from here on
  pdl_wait();
  pdl_trigger();
  // (the batched sampler lets the first launch of an iteration advance its iteration counter: nothing in this kernel
  // reads it, everything launched after it sees the new value)
  if (advance != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *advance += 1;

  if (warp >= GCONS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp != GCONS) return;
    // producer: lane l < 16 copies A slab row l, lane >= 16 copies B slab row l-16 -- or, with a blocked A operand /
    // a contiguous B slab, lane 0 / lane 16 issue ONE copy for the whole slab
    const int row = lane & 15;
    const bool isB = lane >= 16;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
      const GemmItem gi = items[w];
      const int mrows = min(GM, gi.M - gi.m0);
      const int ncols = min(TN, gi.N - gi.n0);
      const uint32_t bytesA = (uint32_t)((mrows + 1) & ~1) * 8u;
      const uint32_t bytesB = (uint32_t)((ncols + 1) & ~1) * 8u;
      const bool ablk = gi.Ablk != nullptr, bcon = gi.bcontig != 0;
      const double* ablk_base = ablk ? gi.Ablk + (long long)(gi.m0 / GM) * gi.ablk_kblocks * ABLK : nullptr;
      for (int kb = gi.kb0; kb < gi.kb1; ++kb, ++it) {
        const int stage = it % GSTAGES;
        mbar_wait(&empty[stage], ((it / GSTAGES) & 1u) ^ 1u);
        const int k = kb * GK + row;
        const int klive = min(GK, gi.K - kb * GK);
        const bool live = k < gi.K;
        // every lane announces its own bytes; lane 0 arrives
        uint32_t mine;
        if (isB) mine = bcon ? (row == 0 ? (uint32_t)(klive * gi.ldb) * 8u : 0u) : (live ? bytesB : 0u);
        else mine = ablk ? (row == 0 ? (uint32_t)(ABLK * 8) : 0u) : (live ? bytesA : 0u);
        uint32_t tot = mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) mbar_expect_tx(&full[stage], tot);
        __syncwarp();
        if (mine > 0) {
          double* stage_base = sm + (size_t)stage * SM::STAGE_DOUBLES;
          if (isB) {
            double* dst = stage_base + GK * GSLD_A + (bcon ? 0 : row * SM::SLD_B);
            const double* src = bcon ? gi.B + (long long)kb * GK * gi.ldb : gi.B + gi.n0 + (long long)k * gi.ldb;
            bulk_g2s(dst, src, mine, &full[stage]);
          } else {
            double* dst = stage_base + (ablk ? 0 : row * GSLD_A);
            const double* src = ablk ? ablk_base + (long long)kb * ABLK : gi.A + gi.m0 + (long long)k * gi.lda;
            bulk_g2s(dst, src, mine, &full[stage]);
          }
        }
      }
    }
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  constexpr int NT8 = TN / 8;
  const int fr = lane >> 2, fk = lane & 3;
  const int m_w = warp * 8 * MI;
  uint32_t it = 0;
  for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
    const GemmItem gi = items[w];
    double acc[MI][NT8][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NT8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kb = gi.kb0; kb < gi.kb1; ++kb, ++it) {
      const int stage = it % GSTAGES;
      mbar_wait(&full[stage], (it / GSTAGES) & 1u);
      const double* As = sm + (size_t)stage * SM::STAGE_DOUBLES + m_w + fr;
      const double* Bs = sm + (size_t)stage * SM::STAGE_DOUBLES + GK * GSLD_A + fr;
      const int sldb = gi.bcontig ? (int)gi.ldb : SM::SLD_B;
      const int klive = min(GK, gi.K - kb * GK);   // rows >= klive hold stale (finite) data: skip them
#pragma unroll
      for (int kk = 0; kk < GK / 4; ++kk) {
        if (kk * 4 < klive) {
          const int k = kk * 4 + fk;
          const bool kv = k < klive;
          double a[MI], b[NT8];
#pragma unroll
          for (int i = 0; i < MI; ++i) a[i] = kv ? As[k * GSLD_A + i * 8] : 0.0;
#pragma unroll
          for (int j = 0; j < NT8; ++j) b[j] = kv ? Bs[k * sldb + j * 8] : 0.0;
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NT8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
    // epilogue: C fragment (row = lane/4, cols 2*(lane%4) + {0,1}) -> this item's partial C
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int m = gi.m0 + m_w + i * 8 + fr;
      if (m >= gi.M) continue;
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        const int nn = gi.n0 + j * 8 + 2 * fk;
        double* c = gi.C + (long long)m * gi.ldc + nn;
        if (nn + 1 < gi.N) {
          *reinterpret_cast<double2*>(c) = make_double2(acc[i][j][0], acc[i][j][1]);
        } else if (nn < gi.N) {
          c[0] = acc[i][j][0];
        }
      }
    }
  }
}

}  // namespace

namespace {
// dst[((mt * KB + kb) * 16 + kl) * (gm + 4) + ml] = A(mt * gm + ml, kb * 16 + kl), zero outside M x K and in the pad columns
__global__ void __launch_bounds__(256)
gemm_pack_kernel(const double* __restrict__ src, long long sm_, long long sk_, long long M, long long K, long long KB,
                 int gm, double* __restrict__ dst) {
  const long long chunk = blockIdx.x;             // (mt, kb)
  const long long mt = chunk / KB, kb = chunk % KB;
  const int pitch = gm + 4, ablk = GK * pitch;
  double* d = dst + chunk * ablk;
  for (int e = threadIdx.x; e < ablk; e += blockDim.x) {
    // consecutive threads walk the source's fast dimension: m when stride_m == 1, k otherwise
    int kl, ml;
    if (sm_ == 1) {
      kl = e / pitch;
      ml = e % pitch;
    } else {
      ml = e / GK;
      kl = e % GK;
    }
    const long long m = mt * gm + ml, k = kb * GK + kl;
    const double v = (ml < gm && m < M && k < K) ? src[m * sm_ + k * sk_] : 0.0;
    d[kl * pitch + ml] = v;
  }
}

template <int TN, int MI>
int dgemm_launch_t(const GemmItem* d_items, int nitems, cudaStream_t st, int grid_hint, bool pdl, long long* advance) {
  static bool attr[64] = {false};   // function attributes are per device
  int dev = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  dev = dev >= 0 && dev < 64 ? dev : 0;
  if (!attr[dev]) {
    BGGE_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<TN, MI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmSmem<TN, MI>::BYTES));
    attr[dev] = true;
  }
  const int grid = grid_hint > 0 ? std::min(grid_hint, nitems) : (nitems < sm_count() ? nitems : sm_count());
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GTHREADS);
  cfg.dynamicSmemBytes = GemmSmem<TN, MI>::BYTES;
  cfg.stream = st;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = la;
  cfg.numAttrs = pdl ? 1 : 0;
  BGGE_CUDA(cudaLaunchKernelEx(&cfg, dgemm_dmma_kernel<TN, MI>, d_items, nitems, advance));
  return OK;
}
}  // namespace

// N tile widths the kernel is instantiated for: the batched sampler picks the one that covers its chain count with
// the fewest padded columns (80 chains per GPU run as one 80-wide tile, not as two 64-wide ones)
const int kGemmTileN[] = {16, 32, 40, 48, 64, 80, 96, 128};
const int kGemmTileNCount = 8;

int gemm_tile_m(int tn) { return tn <= 48 ? 256 : 128; }
long long gemm_ablk_doubles(int tn) { return (long long)GK * (gemm_tile_m(tn) + 4); }

int gemm_pack_launch(const double* src, long long stride_m, long long stride_k, long long M, long long K, int tn, double* dst,
                     cudaStream_t st) {
  const int gm = gemm_tile_m(tn);
  const long long MT = (M + gm - 1) / gm, KB = (K + GK - 1) / GK;
  if (MT * KB <= 0) return OK;
  BGGE_REQUIRE(MT * KB < (1ll << 31), "gemm_pack: too many chunks");
  gemm_pack_kernel<<<(unsigned)(MT * KB), 256, 0, st>>>(src, stride_m, stride_k, M, K, KB, gm, dst);
  BGGE_CUDA(cudaGetLastError());
  return OK;
}

int dgemm_launch(const GemmItem* d_items, int nitems, int tn, cudaStream_t st, int grid, bool pdl, long long* advance) {
  if (nitems <= 0) return OK;
  switch (tn) {
    case 16: return dgemm_launch_t<16, 4>(d_items, nitems, st, grid, pdl, advance);
    case 32: return dgemm_launch_t<32, 4>(d_items, nitems, st, grid, pdl, advance);
    case 40: return dgemm_launch_t<40, 4>(d_items, nitems, st, grid, pdl, advance);
    case 48: return dgemm_launch_t<48, 4>(d_items, nitems, st, grid, pdl, advance);
    case 64: return dgemm_launch_t<64, 2>(d_items, nitems, st, grid, pdl, advance);
    case 80: return dgemm_launch_t<80, 2>(d_items, nitems, st, grid, pdl, advance);
    case 96: return dgemm_launch_t<96, 2>(d_items, nitems, st, grid, pdl, advance);
    case 128: return dgemm_launch_t<128, 2>(d_items, nitems, st, grid, pdl, advance);
    default: break;
  }
  set_error("dgemm: unsupported N tile %d", tn);
  return ERR_INVALID;
}

}  // namespace bgge

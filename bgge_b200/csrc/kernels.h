// Internal (non-ABI) launch interfaces between the translation units of libbgge_b200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace bgge {

// gram.cu
// mirror: 0 lower-triangle tiles only, 1 tiles + transposes (full symmetric S), 2 transposes only (S[:, panel rows]).
// split_on_panel: choose the K split from this launch's tile count (collective row-panel builds) instead of from the
// full triangle (which keeps a panel build bit-identical to the one-shot build).  [tr0b, tr1b): optional second panel.
int gram_launch(const double* dX, int64_t ldx, int64_t L, int64_t p_pad, int64_t tile_row0, int64_t tile_row1,
                double* dS, int64_t lds, int64_t row_off, double divisor, int mirror, cudaStream_t st,
                int split_on_panel = 0, int64_t tr0b = 0, int64_t tr1b = 0);

// getk.cu
int symmetrize_launch(double* dS, int64_t lds, int64_t L, cudaStream_t st, int from_upper = 0);   // lower -> upper (or the reverse)
int gk_distance_launch(double* dS, int64_t lds, int64_t L, double* d_diag_scratch, cudaStream_t st);
size_t quantile_scratch_bytes();
int quantile7_launch(const double* dD, int64_t ldd, int64_t L, double prob, void* d_scratch, double* d_q,
                     cudaStream_t st);
int gk_exp_launch(const double* dD, int64_t ldd, int64_t L, double bw, const double* d_q, double* dK0, int64_t ldk,
                  cudaStream_t st);
// L = order of K0 (genotype codes lie in [0, L)); L <= 0: unknown, columns are expanded in their natural order
int expand_launch(const double* dK0, int64_t ldk, int64_t L, const int* d_gid, const int* d_env, int64_t n, int mode,
                  int kenv, double* dOut, int64_t ldo, cudaStream_t st);

// eig.cu
int eig_decompose(double* dA, int64_t m, int64_t lda, double tol, double* d_W, int64_t* nr_out, cudaStream_t st);
int eig_compact(const double* dV, int64_t ldv, const double* d_W, int64_t m, int64_t nr, int canonical, double* d_s,
                double* dU, int64_t ldu, cudaStream_t st);

int sign_fix_launch(const double* dSrc, int64_t ld, int64_t rows, int64_t nr, const double* d_wts, double* dDst,
                    cudaStream_t st);
int gather_scale_launch(const double* dK0, int64_t ldk, const int* dT, const double* d_sqc, int64_t t, double* dM,
                        cudaStream_t st);
int hash_bits_launch(const double* dX, int64_t count, unsigned long long* d_out2, cudaStream_t st);
int expand_rows_launch(const double* dW, int64_t ldw, const int* d_loc, const double* d_isq, int64_t rows, int64_t nr,
                       double* dU, int64_t ldu, cudaStream_t st);

// dgemm.cu -- one work item of the skinny DMMA GEMM: C[m][n] (row-major, ldc) = sum_k A[m + k lda] B[n + k ldb]
struct GemmItem {
  const double* A;      // plain operand (M-fast), or nullptr when Ablk is set
  const double* B;
  double* C;
  // Blocked A operand (gemm_pack_launch): for M tile mt and k-block kb one contiguous [16][132] chunk -- exactly a
  // ring stage of the kernel -- at Ablk + (mt * kblocks_total + kb) * gemm_ablk_doubles(tn), so the producer issues ONE
  // bulk copy per stage instead of 16 (the TMA unit is request-rate bound on 1 KB copies; r02 ncu: consumers waited
  // on `full` for 30 % of the launch at 40 chains).  bcontig: the B slab of a k-block is contiguous in memory
  // (single N tile, n0 == 0, ldb == the shared-memory pitch): one more single copy.
  const double* Ablk;
  long long ablk_kblocks;
  int bcontig, pad0;
  long long lda, ldb, ldc;
  int M, N, K;          // full extents (for edge clipping)
  int m0, n0;           // tile origin
  int kb0, kb1;         // 16-deep k-blocks [kb0, kb1)
};
// grid > 0: number of persistent CTAs the item list was laid out for (item w runs on CTA w % grid)
// pdl: launch with the programmatic-serialization attribute (the kernel waits for its predecessor itself)
// advance: optional device counter incremented once by the launch (the kernel itself never reads it)
int dgemm_launch(const GemmItem* d_items, int nitems, int tn, cudaStream_t st, int grid = 0, bool pdl = false,
                 long long* advance = nullptr);
// M tile of the GEMM instantiation for N tile width tn (256 rows for tn <= 48, else 128) and the size of one blocked
// A chunk ([16][tile_m + 4] doubles)
int gemm_tile_m(int tn);
long long gemm_ablk_doubles(int tn);
// packs A (element (m, k) at src[m * stride_m + k * stride_k], M x K) into the blocked layout of N tile width tn; dst
// holds ceil(M / tile_m) * ceil(K / 16) * gemm_ablk_doubles(tn) doubles, padding zero-filled
int gemm_pack_launch(const double* src, long long stride_m, long long stride_k, long long M, long long K, int tn, double* dst,
                     cudaStream_t st);
extern const int kGemmTileN[];     // instantiated N tile widths, ascending
extern const int kGemmTileNCount;

// staging.cu: synchronous host <-> device copies; large pageable buffers go through a threaded pinned-staging pipeline
int host_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind,
                 bool sync_device = true);
int host_copy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);

}  // namespace bgge

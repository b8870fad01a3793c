// Batched (multi-chain / multi-fold) Gibbs sampler state.  Internal header.
#pragma once
#include "gibbs.cuh"
#include "kernels.h"

namespace bgge {

constexpr int BATCH_MAXQ = 16;   // fixed effects the batched sampler keeps in registers per thread

struct BatchBlock {
  int64_t pos0 = 0, rows = 0, nr = 0, ldu = 0, col_off = 0;
  double* dU = nullptr;   // rows x nr, column-major (borrowed from the handle that decomposed, or owned)
  // the A operands of the two skinny GEMMs in the blocked layout of dgemm.cu (one bulk copy per ring stage), owned:
  double* dTb = nullptr;  // t(U): M = nr, K = rows   (d = U'e)
  double* dUb = nullptr;  // U:    M = rows, K = nr   (u = U b)
  bool owns = false;
};

struct BatchKernel {
  std::vector<BatchBlock> blocks;
  std::vector<double> s_host;
  double mean_diag = 1.0;
  int64_t nr_total = 0, nr_pad = 0;
  int ksplit_proj = 1, ksplit_back = 1, n_proj = 0, n_back = 0, n_ctiles = 0;   // ksplit_* = planes of partial outputs
  int grid_proj = 0, grid_back = 0;      // persistent CTAs the stream-K item lists were laid out for
  DevBuf<double> s, Dpart, Upart, Bm, zqpart;
  DevBuf<double> ut1;                    // U'1 per eigen-column (column sums over the block's rows): the GEMM runs on
                                         // E0 = R + U_j, the draw subtracts shift * U'1 (see b_draw_kernel)
  DevBuf<GemmItem> proj_items, back_items;
  DevBuf<unsigned char> cover;           // 1 for rows some block of this kernel covers
};

struct Batch {
  int64_t n = 0;
  int nk = 0, B = 0, Bp = 0, device = 0;   // Bp = Bq + 4: row pitch of the n x Bp matrices (== 4 mod 8: a B slab that is
  int Bq = 0;                              // one contiguous copy stays bank-conflict free); Bq = B rounded up to 8 = GEMM N
  int tn = 64;                   // N tile width of the skinny GEMMs (dgemm.cu instantiation)
  cudaStream_t stream = nullptr;
  bool own_stream = false, initialised = false, any_na = false;
  std::vector<BatchKernel> kernels;
  int q = 0;                             // fixed effects shared by all chains (0: none), q <= BATCH_MAXQ
  std::vector<double> xf_host;           // n x q, column-major
  DevBuf<double> xf, txx, lchol;         // XF, solve(crossprod(XF)), its lower Cholesky factor
  DevBuf<double> Bet, BetOld;            // [q][Bp]
  DevBuf<double> xf_part;                // [nchunks][q][Bp] partial sums of XF'(temp + XF Bet)
  DevBuf<double> chain_bet;              // [ncum][q][Bp]
  std::vector<double> y_host;            // chain-major: y of chain b at [b*n, (b+1)*n)
  std::vector<unsigned char> na_host;
  int64_t ite = 0, burn = 0, thin = 1, ncum = 0, done = 0;
  double R2 = 0.5;
  DrawCfg draw{};
  int rows_per_chunk = 0, nchunks = 0;
  // all matrices are row-major n x Bp (chain index fastest)
  DevBuf<double> Y, R, Ep, U, Um, Um2, part, part2;
  DevBuf<int> narank, n_na;              // rank of a missing entry inside its chain (-1: observed); count per chain
  int max_na = 0;
  DevBuf<double> chiK, chiE, zmu, zna;   // draws prepared by b_draw_kernel's extra blocks for the tail kernels
  DevBuf<double> mu, mu0, shift, sigsq, tau, Sce, Sc, sigb;
  DevBuf<long long> it;
  DevBuf<double> chain_mu, chain_varE, chain_varU;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  ~Batch();
};

int batch_init(Batch* g);
int batch_run(Batch* g, int64_t iters);

}  // namespace bgge

/*
 * libbgge_b200 -- C ABI of the B200-native getK / BGGE hot paths.
 *
 * The reference (italo-granato/BGGE v0.6.6) is pure R and has no FFI of its own: its boundary is
 * the two exported closures getK() (R/getK.R:95-96) and BGGE() (R/BGGE.R:84).  These entry points
 * are what a `.Call` shim under those two closures binds (INTEGRATION.md shows the shim); the same
 * symbols are driven from Python/ctypes by bgge_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 on success, otherwise a bgge_status code; bgge_last_error() gives
 *     the message of the last failure on the calling thread.  Nothing throws, nothing exits.
 *   - matrices are column-major IEEE doubles (R's layout); index vectors are 0-based int32.
 *   - `bgge_*`      take HOST buffers and move data themselves.
 *   - `bgge_dev_*`  take DEVICE pointers on the current CUDA device plus a cudaStream_t (as
 *                   void*; NULL = default stream) and move nothing; callers own the memory.
 *   - there is no CPU fallback: without an sm_100 device every compute call fails with
 *     BGGE_ERR_NODEVICE.
 */
#ifndef BGGE_B200_H
#define BGGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

enum bgge_status {
  BGGE_OK = 0,
  BGGE_ERR_INVALID = 1,
  BGGE_ERR_CUDA = 2,
  BGGE_ERR_SOLVER = 3,
  BGGE_ERR_STATE = 4,
  BGGE_ERR_NODEVICE = 5,
  BGGE_ERR_TAPE = 6,
  BGGE_ERR_UNSUPPORTED = 7
};

/* kernel = switch(kernel) of R/getK.R:133-153 */
enum bgge_kernel_kind { BGGE_KERNEL_GB = 0, BGGE_KERNEL_GK = 1 };
/* expansion modes: R/getK.R:138,148,173 (G) / :199-201 (MDs GE) / :207-218 (MDe env k) / :232 (Gi) */
enum bgge_expand_mode { BGGE_EXPAND_G = 0, BGGE_EXPAND_GE = 1, BGGE_EXPAND_ENV = 2, BGGE_EXPAND_GI = 3 };
/* K[[i]]$Type of R/BGGE.R:122,144,156 */
enum bgge_kernel_type { BGGE_TYPE_D = 0, BGGE_TYPE_BD = 1 };
enum bgge_rng_mode { BGGE_RNG_PHILOX = 0, BGGE_RNG_TAPE = 1 };

/* ------------------------------------------------------------------ library / device */
int bgge_version(void);
const char* bgge_last_error(void);
int bgge_device_count(int* count);
int bgge_set_device(int device);
/* the library keeps freed device scratch blocks (<= 4 GB in total) for reuse; this returns them to the driver */
int bgge_release_cache(void);
/* free bytes / total bytes of the current device */
int bgge_device_memory(uint64_t* free_bytes, uint64_t* total_bytes);

/* ------------------------------------------------------------------ getK, host buffers */
/*
 * Base kernels over the L genotypes (R/getK.R:131-149).
 *   X          L x p, X[subjects,] already subset/reordered by the caller (R/getK.R:131)
 *   kind       GB: K0 = X X' / p (R/getK.R:136).  GK: D = dist(X)^2, q = quantile(D, quantil),
 *              K0_i = exp(-bandwidth[i] * D / q) (R/getK.R:142-146)
 *   out_K0     nb matrices L x L, back to back (nb = 1 for GB)
 *   out_q      optional (may be NULL): the GK quantile
 */
int bgge_getk_base(const double* X, int64_t L, int64_t p, int kind, const double* bandwidth, int nb,
                   double quantil, double* out_K0, double* out_q);

/*
 * Genotype / environment expansion of one L x L base kernel to n x n (R/getK.R:138,148,173,
 * 199-201, 207-218, 232).  gid[i] / env[i] = 0-based level codes of Y[,2] / Y[,1] (R/getK.R:101-104).
 */
int bgge_getk_expand(const double* K0, int64_t L, const int32_t* gid, const int32_t* env, int64_t n,
                     int mode, int env_k, double* out);

/*
 * Does the host matrix K (n x n, leading dimension ldk) still equal the expansion of K0 (L x L, leading dimension L;
 * NULL for the intercept mode) that bgge_getk_expand would produce?  Every entry is compared (==, so -0 equals +0 and
 * a NaN never matches); *matches = 1 / 0.  Host-only (no device needed), multi-threaded.  BGGE() uses it to decide
 * whether a kernel entry may take the factorised route or must be decomposed as the dense matrix it was given
 * (the reference always decomposes the matrix it is given, R/BGGE.R:144,160).
 */
int bgge_factor_matches(const double* K, int64_t n, int64_t ldk, const double* K0, int64_t L, const int32_t* gid,
                        const int32_t* env, int mode, int env_k, int* matches);

/*
 * Whole getK numeric body in one call: one upload of X, every requested kernel built on the
 * device and copied to out[k] (n x n each), in the order getK() returns them:
 *   G_1..G_nb, then for MDs GE_1..GE_nb, for MDe env_0..env_{E-1} per base kernel, then Gi.
 * model: 0 = SM/MM, 1 = MDs, 2 = MDe (R/getK.R:188-229).  n_out must equal the kernel count.
 */
int bgge_getk(const double* X, int64_t L, int64_t p, int kind, const double* bandwidth, int nb,
              double quantil, const int32_t* gid, const int32_t* env, int64_t n, int n_env, int model,
              int intercept_random, double* const* out, int n_out, double* out_q);

/* ------------------------------------------------------------------ getK, device pointers */
/* Required layout of the device marker matrix: leading dimension ldx = multiple of 128 >= L with
 * rows L..ldx-1 zero, p_pad = multiple of 16 >= p with columns p..p_pad-1 zero. */
int bgge_marker_ld(int64_t L, int64_t* ldx);
int bgge_marker_cols(int64_t p, int64_t* p_pad);
/*
 * Gram panel: tile rows [tile_row0, tile_row1) (128 rows each) of the LOWER triangle of
 * S = X X' / divisor, written to dS[(i - row_off) + j*lds].  mirror != 0 also writes the
 * transposed entries (needs row_off == 0 and a full L x L dS).
 */
int bgge_dev_gram(const double* dX, int64_t ldx, int64_t L, int64_t p_pad, int64_t tile_row0,
                  int64_t tile_row1, double* dS, int64_t lds, int64_t row_off, double divisor, int mirror,
                  void* stream);
int bgge_dev_symmetrize(double* dS, int64_t lds, int64_t L, void* stream);
/* in place: Gram matrix -> squared Euclidean distances with exact-zero diagonal (R/getK.R:142) */
int bgge_dev_gk_distance(double* dS, int64_t lds, int64_t L, void* stream);
/* exact stats::quantile(type = 7) over all L*L entries; result to device scalar d_q and, if
 * non-NULL, host scalar h_q (synchronises the stream) */
int bgge_dev_quantile7(const double* dD, int64_t ldd, int64_t L, double prob, double* d_q, double* h_q,
                       void* stream);
int bgge_dev_gk_exp(const double* dD, int64_t ldd, int64_t L, double bw, const double* d_q, double* dK0,
                    int64_t ldk, void* stream);
int bgge_dev_expand(const double* dK0, int64_t ldk, int64_t L, const int32_t* d_gid, const int32_t* d_env,
                    int64_t n, int mode, int env_k, double* dOut, int64_t ldo, void* stream);

/* ------------------------------------------------------------------ eigen (eig(), R/BGGE.R:112-116) */
/*
 * Symmetric eigendecomposition of K (m x m, leading dimension ldk, lower triangle read), values
 * descending, keep values > tol.  out_s: m doubles; out_U: m x m doubles (first nr columns valid,
 * leading dimension m); either may be NULL.  canonical_sign != 0 flips each vector so that its
 * largest-|.| entry is positive (R leaves the sign to LAPACK).
 */
int bgge_eig(const double* K, int64_t m, int64_t ldk, double tol, int canonical_sign, double* out_s,
             double* out_U, int64_t* out_nr);
/* dA (m x m, lda) is destroyed.  d_s: m doubles.  dU: ldu x max_cols, ldu >= m (rows m..ldu-1 are
 * zero-filled).  Synchronises the stream (the kept count comes back to the host). */
int bgge_dev_eig(double* dA, int64_t m, int64_t lda, double tol, int canonical_sign, double* d_s, double* dU,
                 int64_t ldu, int64_t max_cols, int64_t* out_nr, void* stream);

/* ------------------------------------------------------------------ Gibbs sampler (R/BGGE.R:189-443) */
typedef struct bgge_gibbs bgge_gibbs;

/* n = length(y), nk = length(K), q = ncol(XF) or 0.  stream NULL: the handle creates its own. */
int bgge_gibbs_create(int64_t n, int nk, int q, void* stream, bgge_gibbs** out);
void bgge_gibbs_destroy(bgge_gibbs* h);
/* y with NA flagged either by na_mask[i] != 0 or, when na_mask is NULL, by NaN (R/BGGE.R:191-200) */
int bgge_gibbs_set_y(bgge_gibbs* h, const double* y, const uint8_t* na_mask);
/* XF n x q column-major (R/BGGE.R:208-212) */
int bgge_gibbs_set_xf(bgge_gibbs* h, const double* XF);
/* mean(diag(K[[j]]$Kernel)) for the prior scale Sc[j] (R/BGGE.R:249) */
int bgge_gibbs_set_kernel(bgge_gibbs* h, int j, int type, double mean_diag);
/*
 * One eigen block of kernel j (setDEC(), R/BGGE.R:144-154 / 156-171): rows [pos0, pos0+rows) of y,
 * nr kept eigenpairs, s descending.  on_device == 0: s/U are host pointers (U rows x nr, leading
 * dimension ldu) and are copied.  on_device != 0: U is a device pointer that the handle borrows
 * (ldu even, base 16-byte aligned; caller keeps it alive); s is still a host pointer.
 */
int bgge_gibbs_add_block(bgge_gibbs* h, int j, int64_t pos0, int64_t rows, int64_t nr, const double* s,
                         const double* U, int64_t ldu, int on_device);
/*
 * Convenience: decompose a kernel matrix on the device exactly as setDEC() does and add its blocks.
 * K n x n host (on_device == 0) or device pointer (on_device != 0, left intact); type D: one block
 * over all rows; type BD: one block per ne[] entry, blocks without an eigenvalue > tol are skipped
 * (R/BGGE.R:162).  Also sets mean(diag(K)).
 */
int bgge_gibbs_add_kernel_matrix(bgge_gibbs* h, int j, int type, const double* K, int64_t ldk, int on_device,
                                 const int64_t* ne, int n_ne, double tol, int canonical_sign);
/*
 * setDEC() of an expanded kernel WITHOUT materialising the n x n matrix: the kernel is described by its
 * L x L base kernel K0 (host or device pointer, leading dimension ldk; ignored for mode GI), the 0-based
 * genotype / environment codes of the n rows (host pointers) and the expansion mode of bgge_getk_expand.
 * Each block is decomposed through the |T| x |T| matrix C^{1/2} K0[T,T] C^{1/2} (T = genotypes present,
 * C = their counts), which has the same eigenpairs > tol as the block itself; blocks whose mask is not
 * constant fall back to the direct m x m decomposition.  Sets mean(diag(K)) as well.  n_codes = length of the gid /
 * env arrays; it must equal the handle's n (the reference fails with "non-conformable arguments" otherwise).
 * When genotypes repeat inside a block (|T| <= 2/3 of its rows, e.g. the G kernel of a multi-environment
 * trial) the block stays factorised: the sweep streams W (|T| x nr) and applies Z C^{-1/2} on the fly, so
 * that kernel step moves |T|/rows of the bytes.  bgge_gibbs_fetch_block / block_device expand it on demand.
 */
int bgge_gibbs_add_expanded_kernel(bgge_gibbs* h, int j, int type, const double* K0, int64_t L, int64_t ldk,
                                   int on_device, const int32_t* gid, const int32_t* env, int64_t n_codes, int mode,
                                   int env_k, const int64_t* ne, int n_ne, double tol, int canonical_sign);
/* eigen blocks currently attached to kernel j (valid before bgge_gibbs_init; U is rows x nr, leading dimension ldu) */
int bgge_gibbs_block_count(bgge_gibbs* h, int j, int* n_blocks);
int bgge_gibbs_fetch_block(bgge_gibbs* h, int j, int b, int64_t* pos0, int64_t* rows, int64_t* nr, double* s, double* U,
                           int64_t ldu);
/* device pointer / leading dimension of a block's U (owned by handle h; other handles may borrow it with
 * bgge_gibbs_add_block(..., on_device = 1) while h is alive: one decomposition, many chains) */
int bgge_gibbs_block_device(bgge_gibbs* h, int j, int b, const double** dU, int64_t* ldu);
/* Column-sharded single chain (the sweep of R/BGGE.R:296-339 spread over processes / GPUs).
 * bgge_gibbs_set_shard (before init): this handle sweeps eigen-columns [nr rank / world, nr (rank + 1) / world) of every
 * basis matrix; the state is replicated.  Then, instead of bgge_gibbs_run, per iteration and per kernel j:
 *   bgge_gibbs_shard_begin(h, j, &vec, &count)   launches the sweep; vec (device, count doubles) holds this process'
 *                                                partial W b and sum b^2/s
 *   <all-reduce (sum) vec over the processes, on the handle's stream>
 *   bgge_gibbs_shard_end(h, j)                   u_j, residual, running means from the reduced vector
 * and after the last kernel bgge_gibbs_shard_iteration_end(h) (variances, imputation, next intercept).  Draws are
 * keyed by the global column: every process holds the same chain.  No fixed effects in this mode. */
int bgge_gibbs_set_shard(bgge_gibbs* h, int rank, int world);
int bgge_gibbs_shard_begin(bgge_gibbs* h, int j, double** d_vector, int64_t* count);
int bgge_gibbs_shard_end(bgge_gibbs* h, int j);
int bgge_gibbs_shard_iteration_end(bgge_gibbs* h);
/* adds `iterations` (may be negative) to the handle's count of finished iterations: for callers that capture one
 * sharded iteration, collective included, into a CUDA graph and replay it (captured calls do not execute, replays
 * do not pass through the library) */
int bgge_gibbs_shard_advance(bgge_gibbs* h, int64_t iterations);
/* sweep implementation: 1 (default) = resident kernel when the whole eigen basis fits the chip's shared memory
 * (small models: one cooperative launch runs all iterations), else the fused single-pass kernel step where
 * the block shapes fit (U read once per kernel per iteration); 2 = as 1 but never resident; 0 = two-pass
 * proj/back kernels only.  Call before bgge_gibbs_init. */
int bgge_gibbs_set_mode(bgge_gibbs* h, int mode);
/* after init: whether bgge_gibbs_run uses the resident kernel, its grid size and shared memory per CTA */
int bgge_gibbs_resident_info(bgge_gibbs* h, int* enabled, int* ctas, int64_t* smem_bytes);
/* after init: merged units of the streaming path -- consecutive single-block kernels that share one decomposed
 * matrix on disjoint rows (the environment kernels of MDe, R/getK.R:207-218) are swept by one launch from the second
 * iteration on; how many units there are and how many kernels they contain */
int bgge_gibbs_merged_info(bgge_gibbs* h, int* units, int* kernels_in_units);
/* after init: which path kernel j uses (*fused: 0 two-pass proj/back, 1 single-pass sweep_fused_kernel, 2 single-pass
 * multi-right-hand-side sweep_mrhs_kernel), its cluster size / cluster (or tile) count, launches per step */
int bgge_gibbs_kernel_info(bgge_gibbs* h, int j, int* fused, int* cluster_size, int* clusters, int* launches);
/* after init: bytes of eigen basis one step of kernel j streams (W, not U, for blocks kept factorised; twice
 * for two-pass steps) and the number of its blocks kept factorised */
int bgge_gibbs_kernel_bytes(bgge_gibbs* h, int j, int64_t* basis_bytes, int* factored_blocks);
/* kernel j_dst of dst borrows all eigen blocks (dense or factorised), eigenvalues and mean(diag) of kernel
 * j_src of src; src must stay alive.  One decomposition, many chains. */
int bgge_gibbs_share_kernel(bgge_gibbs* dst, int j_dst, bgge_gibbs* src, int j_src);
/* injected draws in the reference's call order; copied to the device */
int bgge_gibbs_set_tape(bgge_gibbs* h, const double* normals, int64_t n_normals, const double* chisq,
                        int64_t n_chisq);
/* ite/burn/thin/R2 of R/BGGE.R:84; rng_mode PHILOX uses (seed, chain) */
int bgge_gibbs_init(bgge_gibbs* h, int64_t ite, int64_t burn, int64_t thin, double R2, int rng_mode,
                    uint64_t seed, uint32_t chain);
/* run the next `iters` sweeps (asynchronous on the handle's stream) */
int bgge_gibbs_run(bgge_gibbs* h, int64_t iters);
int bgge_gibbs_sync(bgge_gibbs* h);
/* run `iters` sweeps without the CUDA graph, CUDA events around every launch; ms[2*nk + 2] receives
 * summed device milliseconds: proj_0, back_0, ..., proj_{nk-1}, back_{nk-1}, scalar step, XF step */
int bgge_gibbs_profile(bgge_gibbs* h, int64_t iters, double* ms);
/* sizes: nr[j] = total kept eigenvalues of kernel j; n_cum = floor(ite/thin); normals / chisq per call */
int bgge_gibbs_sizes(bgge_gibbs* h, int64_t* nr, int64_t* n_cum, int64_t* n_na, int64_t* tape_normals,
                     int64_t* tape_chisq);
/* current state after the iterations run so far (any pointer may be NULL):
 * mu, sigsq scalars; sigb[nk]; u[nk*n] kernel-major; e[n] = R's `temp`; bet[q] */
int bgge_gibbs_state(bgge_gibbs* h, double* mu, double* sigsq, double* sigb, double* u, double* e, double* bet);
/* thinned chains (all n_cum recorded values, burn-in included, as in out$chain, R/BGGE.R:435):
 * mu[n_cum], varE[n_cum], varU[nk*n_cum] kernel-major, bet[n_cum*q] row-major by draw */
int bgge_gibbs_chain(bgge_gibbs* h, double* mu, double* varE, double* varU, double* bet);
/*
 * Posterior summaries over the kept draws (R/BGGE.R:408-433): yHat[n], varE, varE_sd, per kernel
 * u[nk*n], u_sd[nk*n], varu[nk], varu_sd[nk], y[n] with the last imputations (R/BGGE.R:441).
 * Any pointer may be NULL.
 */
int bgge_gibbs_summary(bgge_gibbs* h, double* yHat, double* varE, double* varE_sd, double* u, double* u_sd,
                       double* varu, double* varu_sd, double* y);

/* ------------------------------------------------------------------ batched chains / CV folds */
/*
 * B independent BGGE() chains (different y / NA masks = folds, different RNG streams) that share the eigen
 * blocks of every kernel: the two U passes of each kernel step run as skinny FP64 tensor-core GEMMs
 * (nr x B and n x B).  Chain b draws with key (seed, chain0 + b) and reproduces a single-chain handle
 * initialised with bgge_gibbs_init(..., seed, chain0 + b).  Device RNG only, no XF.
 * All per-chain inputs / outputs are chain-major (chain b's vector contiguous).
 */
typedef struct bgge_batch bgge_batch;
int bgge_batch_create(int64_t n, int nk, int n_chains, void* stream, bgge_batch** out);
void bgge_batch_destroy(bgge_batch* h);
int bgge_batch_set_y(bgge_batch* h, const double* Y, const uint8_t* na_mask);              /* B x n */
/* fixed effects shared by all chains: XF n x q column-major, q <= 16 (R/BGGE.R:208-212, 284-290); before init */
int bgge_batch_set_xf(bgge_batch* h, const double* XF, int q);
int bgge_batch_set_kernel(bgge_batch* h, int j, int type, double mean_diag);
int bgge_batch_add_block(bgge_batch* h, int j, int64_t pos0, int64_t rows, int64_t nr, const double* s, const double* U,
                         int64_t ldu, int on_device);
int bgge_batch_init(bgge_batch* h, int64_t ite, int64_t burn, int64_t thin, double R2, uint64_t seed, uint32_t chain0);
int bgge_batch_run(bgge_batch* h, int64_t iters);
int bgge_batch_sync(bgge_batch* h);
/* padded_chains: chain columns the GEMM tiles compute (>= n_chains); gemm_flops_per_iter: 4 rows nr per real chain */
int bgge_batch_info(bgge_batch* h, int* padded_chains, int64_t* gemm_flops_per_iter, int* launches_per_iter);
/* mu[B x n_cum], varE[B x n_cum], varU[nk x B x n_cum] */
int bgge_batch_chain(bgge_batch* h, double* mu, double* varE, double* varU);
/* thinned chain of the fixed effects, bet[B x n_cum x q] (the reference keeps B.mcmc internally, R/BGGE.R:392) */
int bgge_batch_chain_bet(bgge_batch* h, double* bet);
/* yHat[B x n], varE[B], varE_sd[B], u[nk x B x n], u_sd[nk x B x n], varu[nk x B], varu_sd[nk x B], y[B x n] */
int bgge_batch_summary(bgge_batch* h, double* yHat, double* varE, double* varE_sd, double* u, double* u_sd, double* varu,
                       double* varu_sd, double* y);

/* ------------------------------------------------------------------ one-shot BGGE() */
/*
 * What the .Call shim under BGGE() invokes: eigen of every kernel (setDEC), `ite` sweeps and the
 * summaries, in one call.  K[j]: n x n host matrices; types[j]; ne/n_ne as in BGGE(ne=).  XF NULL
 * when q == 0.  normals/chisq NULL unless rng_mode == BGGE_RNG_TAPE.  Outputs as in
 * bgge_gibbs_summary + bgge_gibbs_chain (any may be NULL).
 */
int bgge_fit(const double* y, const uint8_t* na_mask, int64_t n, const double* const* K, const int* types, int nk,
             const double* XF, int q, const int64_t* ne, int n_ne, int64_t ite, int64_t burn, int64_t thin,
             double tol, double R2, int rng_mode, uint64_t seed, const double* normals, int64_t n_normals,
             const double* chisq, int64_t n_chisq, double* yHat, double* varE, double* varE_sd, double* u,
             double* u_sd, double* varu, double* varu_sd, double* y_out, double* chain_mu, double* chain_varE,
             double* chain_varU);

/* ------------------------------------------------------------------ test hooks */
/* Philox4x32-10 block function on the device (known-answer tests) */
int bgge_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* n standard normals / chi-squares from the device generator: stream = kernel index semantics */
int bgge_test_draws(uint64_t seed, uint32_t chain, int64_t n, double* normals, double df, double* chisq);

/*
 * The same with the K list in the form getK() builds it (R/getK.R:138,148,173,201,216,232): every kernel is either an
 * expansion of an L x L base kernel through the row codes gid / env (mode = a bgge_expand_mode; the n x n matrix is
 * never formed, setDEC runs through bgge_gibbs_add_expanded_kernel and the sweep streams the factorised basis) or a
 * dense n x n matrix (mode = BGGE_KERNEL_DENSE: hand-built K lists, R/BGGE.R:55-58; kernels a caller edited).
 * Base kernels with the same K0 pointer are uploaded once.  This is the entry the .Call shim under BGGE() binds.
 *   chunk     > 0: the sweeps run in pieces of `chunk` iterations (BGGE(verbose = chunk), R/BGGE.R:398-401) and
 *             progress(done, seconds_per_iteration, user) is called after each piece on the calling thread -- the R
 *             shim prints the "Iter: ... time: ..." line and calls R_CheckUserInterrupt() there; a non-zero return
 *             stops the run (BGGE_ERR_STATE, message "interrupted").  chunk <= 0 or progress == NULL: one piece.
 *   chain_bet optional recorded fixed effects, n_cum rows of max(q, 1) values (row-major, as bgge_gibbs_chain).
 */
#define BGGE_KERNEL_DENSE (-1)
typedef struct bgge_kernel_desc {
  int type;              /* BGGE_TYPE_D / BGGE_TYPE_BD */
  int mode;              /* bgge_expand_mode, or BGGE_KERNEL_DENSE */
  int env_k;             /* environment of an MDe kernel (mode BGGE_EXPAND_ENV) */
  int reserved;
  int64_t L;             /* order of K0 */
  const double* K0;      /* L x L, column-major, host; NULL for BGGE_EXPAND_GI */
  const double* dense;   /* n x n, column-major, host; mode BGGE_KERNEL_DENSE only */
} bgge_kernel_desc;
typedef int (*bgge_progress_fn)(int64_t done, double seconds_per_iteration, void* user);
int bgge_fit_kernels(const double* y, const uint8_t* na_mask, int64_t n, const bgge_kernel_desc* kernels, int nk,
                     const int32_t* gid, const int32_t* env, const double* XF, int q, const int64_t* ne, int n_ne,
                     int64_t ite, int64_t burn, int64_t thin, double tol, double R2, int rng_mode, uint64_t seed,
                     const double* normals, int64_t n_normals, const double* chisq, int64_t n_chisq, int64_t chunk,
                     bgge_progress_fn progress, void* user, double* yHat, double* varE, double* varE_sd, double* u,
                     double* u_sd, double* varu, double* varu_sd, double* y_out, double* chain_mu, double* chain_varE,
                     double* chain_varU, double* chain_bet, int* factored_kernels);

/* ------------------------------------------------------------------ multi-GPU inside the library
 * SURVEY.md 8b "Threading" / 8e: the reference is one R process, so the fan-out over the GPUs of a box has to live
 * below the .Call boundary.  Communicators wrap NCCL (bound at run time with dlopen, an already loaded libnccl is
 * reused); no torch, MPI or other launcher is needed for the single-process forms.
 */
typedef struct bgge_comm bgge_comm;
#define BGGE_COMM_ID_BYTES 128
/* one communicator per GPU of this process (devices == NULL: 0..ndev-1); out receives ndev handles, rank d = out[d] */
int bgge_comm_init_all(int ndev, const int* devices, bgge_comm** out);
/* multi-process form: rank 0 creates an id, ships the 128 bytes to the other processes by any channel, then every
 * process calls init_rank with its rank on its current device */
int bgge_comm_unique_id(void* id);
int bgge_comm_init_rank(const void* id, int world, int rank, bgge_comm** out);
int bgge_comm_info(const bgge_comm* c, int* rank, int* world, int* device);
void bgge_comm_destroy(bgge_comm* c);
/* in-place sum of d_buf[0..count) over the ranks, ordered on `stream` */
int bgge_dev_allreduce(bgge_comm* c, double* d_buf, int64_t count, void* stream);
/*
 * Row-panel build of S = X X' / divisor (R/getK.R:136; the O(L^2 p) part of :142) over the ranks of c: every rank
 * holds the padded markers dX (bgge_marker_ld / _cols layout), builds the lower-triangle tiles of its two tile-row
 * panels on the FP64 tensor cores, stores them transposed (contiguous column panels of dS) and the panels are
 * exchanged in place; on return dS (L x L, leading dimension L) is complete and symmetric on every rank.
 * ms_compute / ms_exchange (optional): device time of the contraction and of exchange + mirror pass.
 */
int bgge_dev_gram_panels(bgge_comm* c, const double* dX, int64_t ldx, int64_t L, int64_t p_pad, double* dS,
                         double divisor, void* stream, float* ms_compute, float* ms_exchange);
/* bgge_getk_base (R/getK.R:131-149) over the ndev GPUs of this process, host buffers in and out: one host thread per
 * device, marker columns uploaded 1/ndev per device and all-gathered over NVLink, row-panel contraction, device 0
 * runs the L x L epilogue.  comms = the handles of bgge_comm_init_all. */
int bgge_getk_base_multi(bgge_comm* const* comms, int ndev, const double* X, int64_t L, int64_t p, int kind,
                         const double* bandwidth, int nb, double quantil, double* out_K0, double* out_q);
/* Column-sharded chain without a host-side collective: the handle becomes shard `rank of world` of c (call before
 * bgge_gibbs_init, instead of bgge_gibbs_set_shard) and bgge_gibbs_run_sharded sweeps `iters` iterations with the
 * per-step all-reduce issued on the handle's own stream.  Every rank calls it with the same arguments. */
int bgge_gibbs_set_comm(bgge_gibbs* h, bgge_comm* c);
int bgge_gibbs_run_sharded(bgge_gibbs* h, int64_t iters);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BGGE_B200_H */

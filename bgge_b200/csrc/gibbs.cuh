// Device-resident Gibbs sampler state for BGGE() (R/BGGE.R:229-403).  Internal header.
#pragma once
#include "common.cuh"
#include <vector>

namespace bgge {

constexpr int MAXK = 64;   // max number of kernels (random effects) in one model
constexpr double NU = 3.0; // R/BGGE.R:245

// RNG stream ids (Philox counter word 3); kernel j's coefficient draws use stream j.
enum : uint32_t {
  STREAM_MU = 0xFFFF0001u,
  STREAM_XF = 0xFFFF0002u,
  STREAM_NA = 0xFFFF0003u,
  STREAM_CHI_E = 0xFFFF0004u,
  STREAM_INIT_U = 0xFFFF0005u,
  STREAM_CHI_K = 0xFFFE0000u   // + kernel index
};

struct DrawCfg {
  int mode;                      // 0 = device Philox, 1 = injected tapes
  unsigned long long seed;
  unsigned int chain;
  const double* zt;              // normals tape (device)
  const double* ct;              // chi-square tape (device)
  long long z_init;              // nk * n  (init normals precede the loop)
  long long z_per_it;            // 1 + q + sum nr + nNa
  long long z_off_xf;            // 1
  long long z_off_na;            // 1 + q + sum nr
  int c_per_it;                  // nk + 1
};

// chain scalars living in device memory
struct Scal {
  double mu, mu0, shift, sigsq, tau, Sce;
  double mu_done;                // intercept of the last finished iteration (mu is one draw ahead)
  double sigb[MAXK];
  double Sc[MAXK];
  long long it;                  // iteration currently being computed (1-based); 0 before init
};

struct ATask {                   // CT eigen-columns of one block, all of its rows
  const double* U;
  long long ldu;
  int rows, pos0, ncols, col_off;
};

struct BTask {                   // one row tile of one block (or of a gap: nr == 0)
  const double* U;
  long long ldu;
  int rows;                      // valid rows in this tile
  int row0;                      // global row of the tile's first row
  int urow0;                     // row offset inside the block
  int nr, col_off;
};

struct FPanel {                  // fused path: columns [c0, c1) of one block, handled by one cluster
  const double* U;               // dense U (rows x nr) or, for a factorised block, W (vrows x nr)
  long long ldu;
  int rows, pos0, c0, c1, col_off;   // rows = rows the kernel streams (vrows for a factorised block)
  int rs;                        // rows per CTA of the cluster (multiple of 16)
  int slot;                      // upart slot (= cluster index)
  int virt;                      // 1: rows live in the kernel's virtual vectors at offset voff
  int direct;                    // virt only: 1 = identity row map, the residual is read from r / u at pos0s[] instead of
                                 // the gathered vector (the output still goes through the virtual partial sums)
  long long voff;
  // multi-right-hand-side panels: nb factorised blocks that share this W (identical diagonal blocks of a BD
  // kernel) are swept together; member k reads / writes the virtual vectors at voffs[k] and draws with the
  // column offset col_offs[k].  nb == 1: voffs[0] == voff, col_offs[0] == col_off.
  int nb;
  long long voffs[4];
  int col_offs[4];
  // per member: the kernel it belongs to (sigb, draw stream, u), first row of a dense member, the kernel's
  // coefficient-normal offset in a tape iteration, its eigenvalues (from the block's first column) and where the
  // cluster's sum of b^2/s goes.  All members belong to one kernel unless the launch is a merged unit
  // (kernels sharing one matrix on disjoint rows, e.g. the environment kernels of MDe).
  int kids[4];
  int pos0s[4];
  long long zoffs[4];
  const double* ss[4];
  double* bps[4];
};

struct FTile {                   // fused finalize: up to 64 rows, clusters [qlo, qhi] contributed (qlo > qhi: none)
  int row0, rows, qlo, qhi;
  const int* loc;                // factorised block: row -> virtual row (already offset to row0), else nullptr
  const double* isq;             // 1/sqrt(count) per row (offset to row0)
  long long voff;
  int kid, pad;                  // kernel whose u / Welford arrays the rows belong to (-1: none, merged units)
};

struct SSeg {                    // shard_sum_kernel: dst[doff + i] = sum_{q = qlo..qhi} src[q * stride + off + i], i < len
  const double* src;
  long long stride, off, doff;
  int len, qlo, qhi, pad;
};

struct VDesc {                   // one factorised block for virt_gather_kernel
  int vrows, pos0, cta0;         // CTAs [cta0, cta0 + ceil(vrows/256)) of the gather launch belong to it
  long long voff;
  const int* rowptr;
  const int* rowidx;
  const double* isq;
  int kid, pad;                  // kernel whose u enters the gathered residual
};

// An eigen block.  Dense: U (rows x nr).  Factorised (bgge_gibbs_add_expanded_kernel, repeated genotypes):
// U = Z C^{-1/2} W is kept as W (vrows x nr) + the row map, so the sweep streams W only:
// U'e = W'(C^{-1/2} Z'e),  U b = Z C^{-1/2} (W b).
struct HostBlock {
  int64_t pos0 = 0, rows = 0, nr = 0, ldu = 0, col_off = 0;
  double* dU = nullptr;          // dense U, or W when vrows > 0
  bool owns = false;
  int64_t vrows = 0, voff = 0;   // factorised: number of distinct genotypes, offset in the virtual vectors
  int* d_loc = nullptr;          // [rows]
  double* d_isq = nullptr;       // [rows]
  int* d_rowptr = nullptr;       // [vrows + 1] rows of each virtual row (CSR, ascending)
  int* d_rowidx = nullptr;       // [rows]
  double* dU_dense = nullptr;    // lazily materialised dense copy (two-pass path, fetch / borrow)
  int64_t ldu_dense = 0;
  bool owns_dense = false;
  bool owns_map = false;         // loc / isq / CSR arrays (dU may be shared with an identical block)
  bool identity = false;         // factorised block whose row map is the identity (every genotype once, in order, weight
                                 // 1): the sweep reads r / u of its rows directly, no gather launch is needed for it
  int kid = -1;                  // merged units only: the kernel this block belongs to
  int fq_lo = 1 << 30, fq_hi = -1;   // fused plan: clusters that sweep columns of this block (lo > hi: none)
  int64_t stream_rows() const { return vrows > 0 ? vrows : rows; }
};

struct HostKernel {
  std::vector<HostBlock> blocks;
  int64_t nr_total = 0;
  double mean_diag = 1.0;
  DevBuf<double> s;              // concatenated eigenvalues
  DevBuf<double> b;              // concatenated coefficients
  DevBuf<double> bpart;          // per A-task partial sums of b^2/s
  DevBuf<ATask> atasks;
  DevBuf<BTask> btasks;
  int n_atasks = 0, n_btasks = 0, ct = 8, cs = 1;
  std::vector<double> s_host;
  long long z_off = 0;           // offset of this kernel's normals inside an iteration
  // fused single-pass path (sweep_fused.cu)
  bool fused = false, fused_attr_set = false;
  int fused_cs = 1, fused_rpt = 1, fused_g = 1, fused_ns = 0, fused_clusters = 0, n_ftiles = 0;
  size_t fused_smem = 0;
  DevBuf<FPanel> fpanels;
  DevBuf<int2> fcluster;
  DevBuf<FTile> ftiles;
  int64_t vtotal = 0;            // total virtual rows of the factorised blocks
  DevBuf<double> ve;             // virtual residual  C^{-1/2} Z'(e + u_old), [vtotal]
  int fused_nb = 1;              // right-hand sides per panel the fused kernel is instantiated for (1, 2 or 4)
  int fused_engine = 0;          // 0: sweep_fused_kernel, 1: sweep_mrhs_kernel (multi-RHS panels, sweep_mrhs.cu)
  int fused_kr = 0, fused_lag = 0;   // engine 1: row steps per compute thread, groups between dots and axpy
  int fused_tmem = 0;            // engine 1: residual slices kept in tensor memory (2-CTA clusters for 4 members)
  DevBuf<VDesc> vdesc;           // factorised blocks of this kernel (one gather launch for all of them)
  int n_vdesc = 0, vgrid = 0;
  // column-sharded chain (bgge_gibbs_set_shard): this process' partial W b and sum b^2/s of the kernel step, in one
  // vector [ rows of y : n | virtual rows : vtotal | sum b^2/s : 1 ] that the host all-reduces between
  // bgge_gibbs_shard_begin and bgge_gibbs_shard_end
  DevBuf<double> xbuf;
  DevBuf<SSeg> ssegs;
  int n_ssegs = 0;
  DevBuf<FTile> atiles;          // the finalize tiles reading the reduced vector (one "cluster")
  // merged unit (streaming path): up to 4 single-block kernels sharing one matrix on disjoint rows, swept by one
  // launch; `members` lists them, the blocks above are copies tagged with their kernel
  bool merged = false;
  std::vector<int> members;
};

// Resident path (resident.cu): small models whose whole basis fits the shared memory of the chip.
struct RGroup {                  // one basis matrix W (t x nr) and the <= 4 blocks of one kernel that share it
  const double* W;
  long long ldw;
  const double* s;               // eigenvalues of the columns (identical for every member)
  int t, t_pad, nr, nb, mem0;    // members [mem0, mem0 + nb)
  int maxcols;                   // columns per CTA (upper bound)
  long long woff;                // offset (doubles) of this group's slice in a CTA's shared memory
  int doff, coff;                // offsets of its (member, column) draws / its column eigenvalues in shared memory
};
struct RMember {                 // one eigen block
  int pos0, rows, col_off;
  int kidx;                      // kernel (random effect) the block belongs to: its sigb, draw stream, u
  long long part_off;            // offset of its virtual rows inside a part row / the step's etil0 and wsum
  long long z_off;               // offset of the kernel's coefficient normals inside an iteration (tape mode)
};
struct REntry {                  // one row of y in the row phase of step j, ordered by the units of the next kernel
  int row;
  int poff;                      // kernel j: part offset of the row's virtual row, -1 if no block covers the row
  int vpos;                      // slot of its value in the CTA's staging array (entries are stored sorted by poff)
  int kj;                        // kernel of this step that owns the row (-1: none; merged steps only)
  int kn, pad;                   // kernel whose u enters the next step's residual for the row (-1: none)
  double isq;                    // kernel j: row weight (1/sqrt(count); 1 for a dense block)
  double w;                      // next kernel: row weight
};
struct RUnit {                   // one virtual row of the next kernel: entries [e0, e1), eoff < 0: rows it does not cover
  int e0, e1, eoff, pad;
};
struct RStep {                   // one kernel, or up to 4 kernels sharing a matrix on disjoint rows (merged)
  const REntry* entries;
  const RUnit* units;
  int nunits, g0, g1, next;      // groups [g0, g1) of this step; next = index of the following step
  int merged, zk[4];             // zk[m]: kernel that receives sum b^2/s of member slot m (all equal unless merged)
  int woff, nws, pad;            // this step's wsum in a CTA's shared memory: offset and length
  const double* wsum;            // sum of the row weights per virtual row
  double* etil0;                 // C^-1/2 Z'(r + u) per virtual row, written by the previous row phase
};
struct ResidentPlan {
  bool enabled = false, has_merged = false;
  int nsteps = 0;
  int ncta = 0, etp = 0, max_entries = 0, max_units = 0, ndraw = 0, ncolc = 0, ngroups = 0, nwsum = 0;
  DevBuf<unsigned long long> bar;
  size_t smem = 0;
  long long pstride = 0, smem_w = 0;
  DevBuf<RGroup> groups;
  DevBuf<RMember> members;
  DevBuf<RStep> steps;
  std::vector<DevBuf<REntry>> entries;
  std::vector<DevBuf<RUnit>> units;
  std::vector<DevBuf<double>> wsum, etil0;
  DevBuf<double> part, zqpart, sspart, x1, mx0, xrpart;
  long long ps0 = 0;
};

// One decomposed base matrix C^{1/2} K0[T,T] C^{1/2}, kept per handle: identical blocks of one kernel and identical
// blocks of different kernels (the environment kernels of MDe) share W and the eigenvalues.
struct SharedDecomp {
  unsigned long long h1, h2;     // content hash of the matrix that was decomposed
  double tol;                    // eigenvalue threshold it was truncated with
  int canonical;                 // sign convention applied
  std::vector<int> T, cnt;
  double* W;
  int64_t ldw, nr;
  std::vector<double> s;         // eigenvalues (of the UNSCALED K0[T,T] when ucount > 0)
  int ucount = 0;                // > 0: every genotype of T occurs equally often in the first user's block; the matrix
                                 // decomposed is K0[T,T] itself, users scale the eigenvalues by their own count
  double max_dropped = 0.0;      // largest eigenvalue that was truncated (-inf: none), ucount > 0 only
};

struct Gibbs {
  int64_t n = 0;
  int nk = 0, q = 0, n_chains = 1;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::vector<HostKernel> kernels;
  std::vector<SharedDecomp> decomps;
  std::vector<HostKernel> units;      // merged units of the streaming path
  std::vector<int> unit_of;           // per kernel: index into units, or -1
  // data
  std::vector<double> y_host;
  std::vector<unsigned char> na_host;
  std::vector<double> xf_host;
  int64_t n_na = 0;
  // device state
  DevBuf<double> r, u, y, xf, umean, um2, bet, xf_part, txx, lchol, betmean, upart, vupart;
  int shard_rank = 0, shard_world = 1;   // column shard of a single chain across processes / GPUs
  int shard_step = -1;           // kernel step between shard_begin and shard_end (-1: none)
  void* comm = nullptr;          // bgge_comm* when the library runs the all-reduce itself (bgge_gibbs_set_comm)
  int mode = 1;                  // 0: two-pass kernels only, 1: resident kernel if the model fits on chip, else
                                 // fused single-pass where the shape fits, 2: as 1 without the resident kernel
  ResidentPlan resident;
  DevBuf<int> na_idx;
  DevBuf<Scal> scal;
  DevBuf<double> chain_mu, chain_varE, chain_varU, chain_bet;
  DevBuf<double> ztape, ctape;
  int64_t zt_len = 0, ct_len = 0;
  // run config
  int64_t ite = 0, burn = 0, thin = 1, ncum = 0, done = 0;
  double R2 = 0.5;
  DrawCfg draw{};
  bool initialised = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  bool use_graph = true;
  int xf_blocks = 0;

  ~Gibbs();
};

int gibbs_init(Gibbs* g);
int gibbs_run(Gibbs* g, int64_t iters);
int gibbs_profile(Gibbs* g, int64_t iters, double* ms);
int plan_fused(Gibbs* g, HostKernel& hk, int j, cudaStream_t st);
int launch_fused(Gibbs* g, HostKernel& hk, int j);
int launch_fused_finalize(Gibbs* g, HostKernel& hk, int j);
int plan_shard(Gibbs* g, HostKernel& hk, cudaStream_t st);
int launch_shard_sum(Gibbs* g, HostKernel& hk);
int launch_shard_apply(Gibbs* g, HostKernel& hk);
int launch_scalar_sharded(Gibbs* g);
int plan_resident(Gibbs* g);
int launch_resident(Gibbs* g, int64_t iters);
int ensure_dense(HostBlock& bl, cudaStream_t st);   // materialise U = Z C^{-1/2} W of a factorised block

// q x q host linear algebra of the fixed-effects step (gibbs.cu): inverse of a symmetric positive definite matrix
// (column-major) and the lower Cholesky factor; false when singular / not positive definite.
bool solve_spd_inverse(int q, std::vector<double> A, std::vector<double>& inv);
bool chol_lower(int q, const std::vector<double>& A, std::vector<double>& Lw);

}  // namespace bgge

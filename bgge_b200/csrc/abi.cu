// extern "C" surface of libbgge_b200 (declared in include/bgge_b200.h) and the host-side
// orchestration behind the host-buffer entry points.
#include "../../include/bgge_b200.h"
#include "common.cuh"
#include "gibbs.cuh"
#include "batch.cuh"
#include "kernels.h"
#include "multi.cuh"
#include "rng.cuh"
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace bgge {

// ------------------------------------------------------------------ error / device plumbing
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static int g_sm_count[64] = {0};

// ------------------------------------------------------------------ caching device allocator
namespace {
struct PoolBlock {
  size_t bytes;
  int device;
};
std::mutex g_pool_mu;
std::map<void*, PoolBlock> g_pool_live;                          // blocks handed out
std::multimap<std::pair<int, size_t>, void*> g_pool_free;        // (device, bytes) -> cached block
size_t g_pool_cached = 0;
// Blocks above POOL_MAX_BLOCK go straight back to the driver; at most POOL_MAX_CACHED bytes stay cached (per
// process; bgge_release_cache() hands them back).  Sized for the C5 setDEC: one BGGE() call allocates the base kernel,
// the scaled L x L problem, the eigenvectors (0.8 GB each) and cuSOLVER's 3.2 GB syevd workspace -- with smaller limits
// every call went through cudaMalloc / cudaFree of those blocks, which r02 measured at 1.5 ms usually but 50-450 ms
// every few calls (BGGE_TIMING=1), i.e. the difference between 1.11 and 1.56 s per call.
constexpr size_t POOL_MAX_BLOCK = 4096ull << 20;
constexpr size_t POOL_MAX_CACHED = 10240ull << 20;

size_t pool_round(size_t b) {
  if (b <= (1u << 20)) {
    size_t r = 512;
    while (r < b) r <<= 1;
    return r;
  }
  return (b + (1u << 20) - 1) & ~((size_t)(1u << 20) - 1);
}
}  // namespace

cudaError_t pool_alloc(void** p, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const size_t rb = pool_round(bytes);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool_free.find({dev, rb});
    if (it != g_pool_free.end()) {
      *p = it->second;
      g_pool_free.erase(it);
      g_pool_cached -= rb;
      g_pool_live[*p] = {rb, dev};
      return cudaSuccess;
    }
  }
  e = cudaMalloc(p, rb);
  if (e != cudaSuccess) {   // give cached memory back and retry once
    cudaGetLastError();
    pool_release_all();
    e = cudaMalloc(p, rb);
    if (e != cudaSuccess) return e;
  }
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_pool_live[*p] = {rb, dev};
  return cudaSuccess;
}

void pool_free(void* p) {
  if (!p) return;
  PoolBlock b{0, 0};
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool_live.find(p);
    if (it == g_pool_live.end()) {
      cudaFree(p);
      return;
    }
    b = it->second;
    g_pool_live.erase(it);
    if (b.bytes <= POOL_MAX_BLOCK && g_pool_cached + b.bytes <= POOL_MAX_CACHED) {
      g_pool_free.insert({{b.device, b.bytes}, p});
      g_pool_cached += b.bytes;
      return;
    }
  }
  cudaFree(p);
}

void pool_release_all() {
  std::vector<std::pair<int, void*>> blocks;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& kv : g_pool_free) blocks.push_back({kv.first.first, kv.second});
    g_pool_free.clear();
    g_pool_cached = 0;
  }
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& b : blocks) {
    cudaSetDevice(b.first);
    cudaFree(b.second);
  }
  cudaSetDevice(cur);
}

int ensure_device() {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libbgge_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return ERR_NODEVICE;
  }
  int dev = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && g_sm_count[dev] > 0) return OK;
  cudaDeviceProp prop;
  BGGE_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("device %d (%s) is sm_%d%d; libbgge_b200 is built for sm_100a only", dev, prop.name, prop.major,
              prop.minor);
    return ERR_NODEVICE;
  }
  if (dev >= 0 && dev < 64) g_sm_count[dev] = prop.multiProcessorCount;
  return OK;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64 && g_sm_count[dev] > 0) ? g_sm_count[dev] : 148;
}

bool pdl_enabled() {
  static const bool on = getenv("BGGE_NO_PDL") == nullptr;
  return on;
}

namespace {

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// host X (L x p) -> padded device copy
int upload_markers(const double* X, int64_t L, int64_t p, DevBuf<double>& dX, int64_t* ldx, int64_t* p_pad) {
  *ldx = round_up(L, 128);
  *p_pad = round_up(p, 16);
  BGGE_CUDA(dX.alloc((size_t)(*ldx) * (size_t)(*p_pad)));
  BGGE_CUDA(cudaMemset(dX.p, 0, (size_t)(*ldx) * (size_t)(*p_pad) * sizeof(double)));
  BGGE_TRY(host_copy_2d(dX.p, (size_t)(*ldx) * sizeof(double), X, (size_t)L * sizeof(double), (size_t)L * sizeof(double),
                        (size_t)p, cudaMemcpyHostToDevice));
  return OK;
}

// S[i] = (P_0[i] + P_1[i] + ... + P_{k-1}[i]) / divisor, planes `stride` doubles apart, fixed order
__global__ void sum_planes_kernel(const double* __restrict__ P, long long stride, int k, long long count, double divisor,
                                  double* __restrict__ S) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    double a = P[i];
    for (int c = 1; c < k; ++c) a += P[(long long)c * stride + i];
    S[i] = a / divisor;
  }
}

// S = X X' / divisor for a HOST marker matrix (L x p, column-major, leading dimension L).  Small matrices: one upload,
// one contraction.  Large ones (>= 1 GB): the marker columns are cut into chunks that travel through two device
// buffers -- chunk c + 1 is uploaded (staging.cu) while the tensor cores contract chunk c into its own L x L plane --
// and the planes are added in chunk order at the end (deterministic; equal to the one-shot result to rounding).
// C5 (8 GB of markers, 4 chunks): bgge_getk_base 0.69 s (upload, then contraction) -> 0.39-0.43 s.
int gram_from_host(const double* X, int64_t L, int64_t p, double divisor, double* dS, cudaStream_t st) {
  const int64_t ldx = round_up(L, 128);
  const int64_t T = (L + 127) / 128;
  const size_t bytes = (size_t)L * (size_t)p * sizeof(double);
  int nch = 1;
  if (bytes >= (1ull << 30) && !getenv("BGGE_NO_CHUNKED_GRAM")) nch = (int)std::min<size_t>(8, std::max<size_t>(2, (bytes + (1ull << 30)) >> 31));   // ~2 GB per chunk
  if (const char* ev = getenv("BGGE_GRAM_HOST_CHUNKS")) nch = std::max(1, std::min(16, atoi(ev)));   // test / tuning hook
  nch = (int)std::min<int64_t>(nch, (p + 1023) / 1024);
  if (nch <= 1) {
    DevBuf<double> dX;
    int64_t l2, p_pad;
    BGGE_TRY(upload_markers(X, L, p, dX, &l2, &p_pad));
    BGGE_TRY(gram_launch(dX.p, l2, L, p_pad, 0, T, dS, L, 0, divisor, 1, st));
    BGGE_CUDA(cudaStreamSynchronize(st));
    return OK;
  }
  const int64_t pc = round_up((p + nch - 1) / nch, 16);   // marker columns per chunk
  int dev = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  DevBuf<double> buf[2], planes;
  cudaStream_t aux = nullptr, gst = nullptr;
  struct Cleanup {
    cudaStream_t *a, *b;
    ~Cleanup() {
      if (*a) cudaStreamDestroy(*a);
      if (*b) cudaStreamDestroy(*b);
    }
  } cleanup{&aux, &gst};
  for (int k = 0; k < 2; ++k) BGGE_CUDA(buf[k].alloc((size_t)ldx * (size_t)pc));
  BGGE_CUDA(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
  BGGE_CUDA(cudaStreamCreateWithFlags(&gst, cudaStreamNonBlocking));
  BGGE_CUDA(planes.alloc((size_t)nch * (size_t)L * (size_t)L));
  BGGE_CUDA(cudaStreamSynchronize(st));
  // zero the buffer (row and column padding), then bring the chunk's columns in; nothing else touches the buffer
  auto upload = [&](int c) -> int {
    const int b = c & 1;
    const int64_t c0 = (int64_t)c * pc, cols = std::max<int64_t>(0, std::min<int64_t>(p, c0 + pc) - c0);
    BGGE_CUDA(cudaMemsetAsync(buf[b].p, 0, (size_t)ldx * (size_t)pc * sizeof(double), aux));
    BGGE_CUDA(cudaStreamSynchronize(aux));
    if (cols > 0)
      BGGE_TRY(host_copy_2d(buf[b].p, (size_t)ldx * sizeof(double), X + (size_t)c0 * (size_t)L, (size_t)L * sizeof(double),
                            (size_t)L * sizeof(double), (size_t)cols, cudaMemcpyHostToDevice, false));
    return OK;
  };
  BGGE_TRY(upload(0));
  for (int c = 0; c < nch; ++c) {
    // gram_launch returns when its kernels have finished (its work lists die with it): it runs on a helper thread
    // while this thread uploads the next chunk into the other buffer
    int gstatus = OK;
    std::string gmsg;
    std::thread worker([&, c]() {
      if (cudaSetDevice(dev) != cudaSuccess) {
        gstatus = ERR_CUDA;
        gmsg = "cudaSetDevice failed on the contraction thread";
        return;
      }
      gstatus = gram_launch(buf[c & 1].p, ldx, L, pc, 0, T, planes.p + (size_t)c * (size_t)L * (size_t)L, L, 0, 1.0, 1, gst);
      if (gstatus != OK) gmsg = last_error();
    });
    const int ustatus = c + 1 < nch ? upload(c + 1) : OK;
    worker.join();
    if (gstatus != OK) {
      set_error("%s", gmsg.c_str());
      return gstatus;
    }
    BGGE_TRY(ustatus);
  }
  const long long count = (long long)L * L;
  sum_planes_kernel<<<(unsigned)std::min<long long>(4096, (count + 255) / 256), 256, 0, st>>>(planes.p, count, nch, count, divisor, dS);
  BGGE_CUDA(cudaGetLastError());
  BGGE_CUDA(cudaStreamSynchronize(st));
  return OK;
}

// device base kernels: K0[b] (L x L, ld L) for each bandwidth.  dX == nullptr: hostX (L x p, ld L) is contracted through
// gram_from_host instead.
int base_kernels_device(const double* dX, int64_t ldx, int64_t L, int64_t p, int64_t p_pad, int kind,
                        const double* bandwidth, int nb, double quantil, std::vector<DevBuf<double>>& K0,
                        double* out_q, cudaStream_t st, const double* hostX = nullptr) {
  BGGE_REQUIRE(kind == BGGE_KERNEL_GB || kind == BGGE_KERNEL_GK,
               "kernel selected is not available. Please choose one method available or make available other "
               "kernel through argument K");
  const int64_t T = (L + 127) / 128;
  auto gram = [&](double* dS, double divisor) -> int {
    if (dX) return gram_launch(dX, ldx, L, p_pad, 0, T, dS, L, 0, divisor, 1, st);
    return gram_from_host(hostX, L, p, divisor, dS, st);
  };
  if (kind == BGGE_KERNEL_GB) {
    K0.resize(1);
    BGGE_CUDA(K0[0].alloc((size_t)L * L));
    BGGE_TRY(gram(K0[0].p, (double)p));
    BGGE_CUDA(cudaStreamSynchronize(st));
    return OK;
  }
  BGGE_REQUIRE(nb >= 1 && bandwidth != nullptr, "GK kernel needs at least one bandwidth");
  DevBuf<double> D, diag, q;
  DevBuf<unsigned char> scratch;
  BGGE_CUDA(D.alloc((size_t)L * L));
  BGGE_CUDA(diag.alloc((size_t)L));
  BGGE_CUDA(q.alloc(1));
  BGGE_CUDA(scratch.alloc(quantile_scratch_bytes()));
  BGGE_TRY(gram(D.p, 1.0));
  BGGE_TRY(gk_distance_launch(D.p, L, L, diag.p, st));
  BGGE_TRY(quantile7_launch(D.p, L, L, quantil, scratch.p, q.p, st));
  K0.resize((size_t)nb);
  for (int b = 0; b < nb; ++b) {
    BGGE_CUDA(K0[(size_t)b].alloc((size_t)L * L));
    BGGE_TRY(gk_exp_launch(D.p, L, L, bandwidth[b], q.p, K0[(size_t)b].p, L, st));
  }
  if (out_q) BGGE_CUDA(cudaMemcpyAsync(out_q, q.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  BGGE_CUDA(cudaStreamSynchronize(st));
  return OK;
}

__global__ void philox_test_kernel(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}
__global__ void draws_test_kernel(DrawCfg dc, long long n, double* z, double df, double* c) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (z) z[i] = draw_normal(dc, 1, 0u, (unsigned long long)i, 0);
  if (c) c[i] = draw_chisq(dc, i + 1, STREAM_CHI_E, df, 0);
}

double mean_sd(const std::vector<double>& v, double* sd) {
  const size_t k = v.size();
  long double m = 0.0L;
  for (double x : v) m += x;
  m = k ? m / (long double)k : 0.0L;
  if (sd) {
    long double ss = 0.0L;
    for (double x : v) ss += ((long double)x - m) * ((long double)x - m);
    *sd = k > 1 ? (double)sqrtl(ss / (long double)(k - 1)) : NAN;
  }
  return k ? (double)m : NAN;
}

}  // namespace
}  // namespace bgge

using namespace bgge;

struct bgge_gibbs {
  Gibbs g;
};

struct bgge_batch {
  Batch g;
};

extern "C" {

int bgge_version(void) { return 100; }
const char* bgge_last_error(void) { return last_error(); }

int bgge_device_count(int* count) {
  BGGE_REQUIRE(count != nullptr, "device_count: NULL output");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  *count = c;
  return OK;
}

int bgge_set_device(int device) {
  BGGE_CUDA(cudaSetDevice(device));
  return ensure_device();
}

int bgge_release_cache(void) {
  pool_release_all();
  return OK;
}

int bgge_device_memory(uint64_t* free_bytes, uint64_t* total_bytes) {
  BGGE_TRY(ensure_device());
  size_t f = 0, t = 0;
  BGGE_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return OK;
}

// ====================================================================== getK
int bgge_marker_ld(int64_t L, int64_t* ldx) {
  BGGE_REQUIRE(L > 0 && ldx, "marker_ld: bad arguments");
  *ldx = round_up(L, 128);
  return OK;
}
int bgge_marker_cols(int64_t p, int64_t* p_pad) {
  BGGE_REQUIRE(p > 0 && p_pad, "marker_cols: bad arguments");
  *p_pad = round_up(p, 16);
  return OK;
}

int bgge_dev_gram(const double* dX, int64_t ldx, int64_t L, int64_t p_pad, int64_t tile_row0, int64_t tile_row1,
                  double* dS, int64_t lds, int64_t row_off, double divisor, int mirror, void* stream) {
  BGGE_TRY(ensure_device());
  const int64_t T = (L + 127) / 128;
  BGGE_REQUIRE(dX && dS && L > 0, "dev_gram: NULL pointer or empty matrix");
  BGGE_REQUIRE(0 <= tile_row0 && tile_row0 <= tile_row1 && tile_row1 <= T, "dev_gram: tile rows [%lld,%lld) outside [0,%lld)",
               (long long)tile_row0, (long long)tile_row1, (long long)T);
  BGGE_REQUIRE(divisor != 0.0, "dev_gram: zero divisor");
  return gram_launch(dX, ldx, L, p_pad, tile_row0, tile_row1, dS, lds, row_off, divisor, mirror, as_stream(stream));
}

int bgge_dev_symmetrize(double* dS, int64_t lds, int64_t L, void* stream) {
  BGGE_TRY(ensure_device());
  return symmetrize_launch(dS, lds, L, as_stream(stream));
}

int bgge_dev_gk_distance(double* dS, int64_t lds, int64_t L, void* stream) {
  BGGE_TRY(ensure_device());
  DevBuf<double> diag;
  BGGE_CUDA(diag.alloc((size_t)L));
  BGGE_TRY(gk_distance_launch(dS, lds, L, diag.p, as_stream(stream)));
  BGGE_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return OK;
}

int bgge_dev_quantile7(const double* dD, int64_t ldd, int64_t L, double prob, double* d_q, double* h_q, void* stream) {
  BGGE_TRY(ensure_device());
  DevBuf<unsigned char> scratch;
  DevBuf<double> qtmp;
  BGGE_CUDA(scratch.alloc(quantile_scratch_bytes()));
  if (!d_q) {
    BGGE_CUDA(qtmp.alloc(1));
    d_q = qtmp.p;
  }
  BGGE_TRY(quantile7_launch(dD, ldd, L, prob, scratch.p, d_q, as_stream(stream)));
  if (h_q) BGGE_CUDA(cudaMemcpyAsync(h_q, d_q, sizeof(double), cudaMemcpyDeviceToHost, as_stream(stream)));
  BGGE_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return OK;
}

int bgge_dev_gk_exp(const double* dD, int64_t ldd, int64_t L, double bw, const double* d_q, double* dK0, int64_t ldk,
                    void* stream) {
  BGGE_TRY(ensure_device());
  return gk_exp_launch(dD, ldd, L, bw, d_q, dK0, ldk, as_stream(stream));
}

int bgge_dev_expand(const double* dK0, int64_t ldk, int64_t L, const int32_t* d_gid, const int32_t* d_env, int64_t n,
                    int mode, int env_k, double* dOut, int64_t ldo, void* stream) {
  BGGE_TRY(ensure_device());
  return expand_launch(dK0, ldk, L, d_gid, d_env, n, mode, env_k, dOut, ldo, as_stream(stream));
}

int bgge_getk_base(const double* X, int64_t L, int64_t p, int kind, const double* bandwidth, int nb, double quantil,
                   double* out_K0, double* out_q) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(X && out_K0 && L > 0 && p > 0, "getk_base: NULL pointer or empty matrix");
  std::vector<DevBuf<double>> K0;
  BGGE_TRY(base_kernels_device(nullptr, 0, L, p, 0, kind, bandwidth, nb, quantil, K0, out_q, nullptr, X));
  for (size_t b = 0; b < K0.size(); ++b)
    BGGE_TRY(host_copy(out_K0 + b * (size_t)L * L, K0[b].p, (size_t)L * L * sizeof(double), cudaMemcpyDeviceToHost));
  return OK;
}

static int check_codes(const int32_t* v, int64_t n, int64_t levels, const char* what) {
  for (int64_t i = 0; i < n; ++i)
    BGGE_REQUIRE(v[i] >= 0 && v[i] < levels, "%s[%lld]=%d outside [0,%lld)", what, (long long)i, v[i], (long long)levels);
  return OK;
}

int bgge_getk_expand(const double* K0, int64_t L, const int32_t* gid, const int32_t* env, int64_t n, int mode,
                     int env_k, double* out) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(gid && out && L > 0 && n > 0, "getk_expand: NULL pointer or empty input");
  BGGE_REQUIRE(mode == BGGE_EXPAND_GI || K0 != nullptr, "getk_expand: K0 is NULL");
  BGGE_REQUIRE((mode != BGGE_EXPAND_GE && mode != BGGE_EXPAND_ENV) || env != nullptr, "getk_expand: env is NULL");
  BGGE_TRY(check_codes(gid, n, L, "gid"));
  DevBuf<double> dK0, dOut;
  DevBuf<int> dgid, denv;
  if (K0) {
    BGGE_CUDA(dK0.alloc((size_t)L * L));
    BGGE_TRY(host_copy(dK0.p, K0, (size_t)L * L * sizeof(double), cudaMemcpyHostToDevice));
  }
  BGGE_CUDA(dgid.alloc((size_t)n));
  BGGE_CUDA(cudaMemcpy(dgid.p, gid, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  if (env) {
    BGGE_CUDA(denv.alloc((size_t)n));
    BGGE_CUDA(cudaMemcpy(denv.p, env, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  }
  BGGE_CUDA(dOut.alloc((size_t)n * n));
  BGGE_TRY(expand_launch(dK0.p, L, L, dgid.p, denv.p, n, mode, env_k, dOut.p, n, nullptr));
  BGGE_TRY(host_copy(out, dOut.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
  return OK;
}

int bgge_factor_matches(const double* K, int64_t n, int64_t ldk, const double* K0, int64_t L, const int32_t* gid,
                        const int32_t* env, int mode, int env_k, int* matches) {
  BGGE_REQUIRE(K && gid && matches && n > 0 && ldk >= n && L > 0, "factor_matches: NULL pointer or empty input");
  BGGE_REQUIRE(mode >= BGGE_EXPAND_G && mode <= BGGE_EXPAND_GI, "factor_matches: unknown expansion mode %d", mode);
  BGGE_REQUIRE(mode == BGGE_EXPAND_GI || K0 != nullptr, "factor_matches: K0 is NULL");
  BGGE_REQUIRE(mode == BGGE_EXPAND_G || mode == BGGE_EXPAND_GI || env != nullptr, "factor_matches: env codes missing");
  BGGE_TRY(check_codes(gid, n, L, "gid"));
  std::atomic<bool> differs{false};
  int T = (int)std::thread::hardware_concurrency();
  if (const char* ev = getenv("BGGE_HOST_THREADS")) T = atoi(ev);
  T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(T, 32), n * n / (1 << 20) + 1));
  auto work = [&](int t) {
    for (int64_t j = t; j < n && !differs.load(std::memory_order_relaxed); j += T) {   // interleaved columns: early exit
      const double* col = K + j * ldk;
      const int gj = gid[j];
      bool bad = false;
      if (mode == BGGE_EXPAND_GI) {
        for (int64_t i = 0; i < n; ++i) bad |= !(col[i] == (gid[i] == gj ? 1.0 : 0.0));
      } else {
        const double* kc = K0 + (int64_t)gj * L;
        if (mode == BGGE_EXPAND_G) {
          for (int64_t i = 0; i < n; ++i) bad |= !(col[i] == kc[gid[i]]);
        } else if (mode == BGGE_EXPAND_GE) {
          const int ej = env[j];
          for (int64_t i = 0; i < n; ++i) bad |= !(col[i] == (env[i] == ej ? kc[gid[i]] : 0.0));
        } else {
          const bool cj = env[j] == env_k;
          for (int64_t i = 0; i < n; ++i) bad |= !(col[i] == (cj && env[i] == env_k ? kc[gid[i]] : 0.0));
        }
      }
      if (bad) differs.store(true, std::memory_order_relaxed);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  *matches = differs.load() ? 0 : 1;
  return OK;
}

int bgge_getk(const double* X, int64_t L, int64_t p, int kind, const double* bandwidth, int nb, double quantil,
              const int32_t* gid, const int32_t* env, int64_t n, int n_env, int model, int intercept_random,
              double* const* out, int n_out, double* out_q) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(X && gid && out && L > 0 && p > 0 && n > 0, "getk: NULL pointer or empty input");
  BGGE_REQUIRE(model >= 0 && model <= 2, "Model selected is not available ");
  BGGE_REQUIRE(model == 0 || (env != nullptr && n_env >= 1), "getk: env codes needed for MDs/MDe");
  BGGE_TRY(check_codes(gid, n, L, "gid"));
  if (env) BGGE_TRY(check_codes(env, n, n_env, "env"));
  const int nbase = kind == BGGE_KERNEL_GB ? 1 : nb;
  const int expect = nbase * (model == 0 ? 1 : model == 1 ? 2 : 1 + n_env) + (intercept_random ? 1 : 0);
  BGGE_REQUIRE(n_out == expect, "getk: %d output matrices supplied, model needs %d", n_out, expect);

  std::vector<DevBuf<double>> K0;
  {
    BGGE_TRY(base_kernels_device(nullptr, 0, L, p, 0, kind, bandwidth, nb, quantil, K0, out_q, nullptr, X));
  }   // markers released before the n x n expansions are allocated
  DevBuf<int> dgid, denv;
  DevBuf<double> dOut;
  BGGE_CUDA(dgid.alloc((size_t)n));
  BGGE_CUDA(cudaMemcpy(dgid.p, gid, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  if (env) {
    BGGE_CUDA(denv.alloc((size_t)n));
    BGGE_CUDA(cudaMemcpy(denv.p, env, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  }
  BGGE_CUDA(dOut.alloc((size_t)n * n));
  int o = 0;
  auto emit = [&](const double* k0, int mode, int k) -> int {
    BGGE_TRY(expand_launch(k0, L, L, dgid.p, denv.p, n, mode, k, dOut.p, n, nullptr));
    BGGE_TRY(host_copy(out[o++], dOut.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
    return OK;
  };
  for (int b = 0; b < nbase; ++b) BGGE_TRY(emit(K0[(size_t)b].p, BGGE_EXPAND_G, 0));
  if (model == 1)
    for (int b = 0; b < nbase; ++b) BGGE_TRY(emit(K0[(size_t)b].p, BGGE_EXPAND_GE, 0));
  if (model == 2)
    for (int b = 0; b < nbase; ++b)
      for (int k = 0; k < n_env; ++k) BGGE_TRY(emit(K0[(size_t)b].p, BGGE_EXPAND_ENV, k));
  if (intercept_random) BGGE_TRY(emit(nullptr, BGGE_EXPAND_GI, 0));
  return OK;
}

// ====================================================================== eigen
int bgge_dev_eig(double* dA, int64_t m, int64_t lda, double tol, int canonical_sign, double* d_s, double* dU,
                 int64_t ldu, int64_t max_cols, int64_t* out_nr, void* stream) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(dA && d_s && out_nr, "dev_eig: NULL pointer");
  DevBuf<double> W;
  BGGE_CUDA(W.alloc((size_t)m));
  int64_t nr = 0;
  BGGE_TRY(eig_decompose(dA, m, lda, tol, W.p, &nr, as_stream(stream)));
  *out_nr = nr;
  if (dU) {
    BGGE_REQUIRE(nr <= max_cols, "dev_eig: %lld eigenvalues exceed tol but output holds %lld columns", (long long)nr,
                 (long long)max_cols);
    BGGE_TRY(eig_compact(dA, lda, W.p, m, nr, canonical_sign, d_s, dU, ldu, as_stream(stream)));
  } else if (nr > 0) {
    DevBuf<double> tmp;   // eigenvalues only: compact into a 1-row dummy is wasteful; reverse on host
    std::vector<double> hW((size_t)m), desc((size_t)nr);
    BGGE_CUDA(cudaMemcpyAsync(hW.data(), W.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, as_stream(stream)));
    BGGE_CUDA(cudaStreamSynchronize(as_stream(stream)));
    for (int64_t c = 0; c < nr; ++c) desc[(size_t)c] = hW[(size_t)(m - 1 - c)];
    BGGE_CUDA(cudaMemcpy(d_s, desc.data(), (size_t)nr * sizeof(double), cudaMemcpyHostToDevice));
  }
  BGGE_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return OK;
}

int bgge_eig(const double* K, int64_t m, int64_t ldk, double tol, int canonical_sign, double* out_s, double* out_U,
             int64_t* out_nr) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(K && out_nr && m > 0 && ldk >= m, "eig: NULL pointer or bad dimensions");
  DevBuf<double> dA, dS, dU;
  BGGE_CUDA(dA.alloc((size_t)m * m));
  BGGE_CUDA(dS.alloc((size_t)m));
  BGGE_TRY(host_copy_2d(dA.p, (size_t)m * sizeof(double), K, (size_t)ldk * sizeof(double), (size_t)m * sizeof(double), (size_t)m,
                        cudaMemcpyHostToDevice));
  if (out_U) BGGE_CUDA(dU.alloc((size_t)m * m));
  int64_t nr = 0;
  BGGE_TRY(bgge_dev_eig(dA.p, m, m, tol, canonical_sign, dS.p, dU.p, m, m, &nr, nullptr));
  *out_nr = nr;
  if (out_s && nr > 0) BGGE_CUDA(cudaMemcpy(out_s, dS.p, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_U && nr > 0) BGGE_CUDA(cudaMemcpy(out_U, dU.p, (size_t)m * nr * sizeof(double), cudaMemcpyDeviceToHost));
  return OK;
}

// ====================================================================== Gibbs handle
int bgge_gibbs_create(int64_t n, int nk, int q, void* stream, bgge_gibbs** out) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(out != nullptr, "gibbs_create: NULL output");
  BGGE_REQUIRE(n > 1 && nk >= 1 && nk <= MAXK && q >= 0 && q <= 1024, "gibbs_create: bad sizes n=%lld nk=%d q=%d",
               (long long)n, nk, q);
  std::unique_ptr<bgge_gibbs> h(new (std::nothrow) bgge_gibbs());
  BGGE_REQUIRE(h != nullptr, "gibbs_create: out of host memory");
  Gibbs& g = h->g;
  g.n = n;
  g.nk = nk;
  g.q = q;
  g.kernels.resize((size_t)nk);
  BGGE_CUDA(cudaGetDevice(&g.device));
  if (stream) {
    g.stream = as_stream(stream);
  } else {
    BGGE_CUDA(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    g.own_stream = true;
  }
  *out = h.release();
  return OK;
}

void bgge_gibbs_destroy(bgge_gibbs* h) { delete h; }

#define BGGE_HANDLE(h)                                                    \
  BGGE_REQUIRE((h) != nullptr, "NULL gibbs handle");                     \
  BGGE_CUDA(cudaSetDevice((h)->g.device))

int bgge_gibbs_set_y(bgge_gibbs* h, const double* y, const uint8_t* na_mask) {
  BGGE_HANDLE(h);
  BGGE_REQUIRE(y != nullptr, "gibbs_set_y: NULL y");
  Gibbs& g = h->g;
  g.y_host.assign(y, y + g.n);
  if (na_mask) g.na_host.assign(na_mask, na_mask + g.n); else g.na_host.clear();
  return OK;
}

int bgge_gibbs_set_xf(bgge_gibbs* h, const double* XF) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.q > 0 && XF != nullptr, "gibbs_set_xf: handle was created with q=0 or XF is NULL");
  g.xf_host.assign(XF, XF + g.n * g.q);
  return OK;
}

int bgge_gibbs_set_kernel(bgge_gibbs* h, int j, int type, double mean_diag) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(j >= 0 && j < g.nk, "gibbs_set_kernel: kernel index %d outside [0,%d)", j, g.nk);
  BGGE_REQUIRE(type == BGGE_TYPE_D || type == BGGE_TYPE_BD, "Matrix should be of types BD or D");
  BGGE_REQUIRE(mean_diag > 0.0 && std::isfinite(mean_diag), "gibbs_set_kernel: mean(diag(K)) must be positive");
  g.kernels[(size_t)j].mean_diag = mean_diag;
  return OK;
}

int bgge_gibbs_add_block(bgge_gibbs* h, int j, int64_t pos0, int64_t rows, int64_t nr, const double* s, const double* U,
                         int64_t ldu, int on_device) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(!g.initialised, "gibbs_add_block: handle already initialised");
  BGGE_REQUIRE(j >= 0 && j < g.nk, "gibbs_add_block: kernel index %d outside [0,%d)", j, g.nk);
  BGGE_REQUIRE(s && U && rows > 0 && nr > 0 && pos0 >= 0 && pos0 + rows <= g.n && ldu >= rows,
               "gibbs_add_block: bad block pos0=%lld rows=%lld nr=%lld ldu=%lld", (long long)pos0, (long long)rows,
               (long long)nr, (long long)ldu);
  for (int64_t c = 0; c < nr; ++c) BGGE_REQUIRE(s[c] > 0.0, "gibbs_add_block: eigenvalue %lld is not positive", (long long)c);
  HostKernel& hk = g.kernels[(size_t)j];
  HostBlock bl;
  bl.pos0 = pos0;
  bl.rows = rows;
  bl.nr = nr;
  if (on_device) {
    BGGE_REQUIRE(ldu % 2 == 0 && (reinterpret_cast<uintptr_t>(U) & 15u) == 0,
                 "gibbs_add_block: device U needs an even leading dimension and a 16-byte aligned base");
    bl.dU = const_cast<double*>(U);
    bl.ldu = ldu;
    bl.owns = false;
  } else {
    bl.ldu = round_up(rows, 16);
    BGGE_CUDA(pool_alloc((void**)&bl.dU, (size_t)bl.ldu * nr * sizeof(double)));
    bl.owns = true;
    bl.col_off = (int64_t)hk.s_host.size();   // where this block's eigenvalues start (insertion order)
    hk.blocks.push_back(bl);   // owned from here on (freed by the handle even on failure below)
    BGGE_CUDA(cudaMemset(bl.dU, 0, (size_t)bl.ldu * nr * sizeof(double)));
    BGGE_CUDA(cudaMemcpy2D(bl.dU, (size_t)bl.ldu * sizeof(double), U, (size_t)ldu * sizeof(double),
                           (size_t)rows * sizeof(double), (size_t)nr, cudaMemcpyHostToDevice));
    hk.s_host.insert(hk.s_host.end(), s, s + nr);
    return OK;
  }
  bl.col_off = (int64_t)hk.s_host.size();   // where this block's eigenvalues start (insertion order)
    hk.blocks.push_back(bl);
  hk.s_host.insert(hk.s_host.end(), s, s + nr);
  return OK;
}

int bgge_gibbs_add_kernel_matrix(bgge_gibbs* h, int j, int type, const double* K, int64_t ldk, int on_device,
                                 const int64_t* ne, int n_ne, double tol, int canonical_sign) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(!g.initialised, "gibbs_add_kernel_matrix: handle already initialised");
  BGGE_REQUIRE(j >= 0 && j < g.nk && K != nullptr && ldk >= g.n, "gibbs_add_kernel_matrix: bad arguments");
  BGGE_REQUIRE(type == BGGE_TYPE_D || type == BGGE_TYPE_BD, "Matrix should be of types BD or D");
  const int64_t n = g.n;
  // mean(diag(K))  (R/BGGE.R:249)
  std::vector<double> diag((size_t)n);
  BGGE_CUDA(cudaMemcpy2D(diag.data(), sizeof(double), K, (size_t)(ldk + 1) * sizeof(double), sizeof(double), (size_t)n,
                         on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
  long double tr = 0.0L;
  for (double d : diag) tr += d;
  BGGE_TRY(bgge_gibbs_set_kernel(h, j, type, (double)(tr / (long double)n)));

  std::vector<std::pair<int64_t, int64_t>> spans;   // (pos0, rows)
  if (type == BGGE_TYPE_D) {
    spans.push_back({0, n});
  } else {
    BGGE_REQUIRE(ne != nullptr && n_ne > 1, "ne invalid. For type BD, number of subjects in each sub matrices should be provided");
    int64_t pos = 0;
    for (int k = 0; k < n_ne; ++k) {
      BGGE_REQUIRE(ne[k] > 0 && pos + ne[k] <= n, "gibbs_add_kernel_matrix: ne does not partition the rows of y");
      spans.push_back({pos, ne[k]});
      pos += ne[k];
    }
  }
  HostKernel& hk = g.kernels[(size_t)j];
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (auto& sp : spans) {
    const int64_t m = sp.second;
    DevBuf<double> dA, dW, dS;
    BGGE_CUDA(dA.alloc((size_t)m * m));
    BGGE_CUDA(dW.alloc((size_t)m));
    BGGE_TRY(host_copy_2d(dA.p, (size_t)m * sizeof(double), K + sp.first + sp.first * ldk, (size_t)ldk * sizeof(double),
                          (size_t)m * sizeof(double), (size_t)m, kind));   // (device-to-device goes straight to cudaMemcpy2D)
    int64_t nr = 0;
    BGGE_TRY(eig_decompose(dA.p, m, m, tol, dW.p, &nr, g.stream));
    if (nr == 0) continue;                                          // R/BGGE.R:162
    HostBlock bl;
    bl.pos0 = sp.first;
    bl.rows = m;
    bl.nr = nr;
    bl.ldu = round_up(m, 16);
    BGGE_CUDA(pool_alloc((void**)&bl.dU, (size_t)bl.ldu * nr * sizeof(double)));
    bl.owns = true;
    bl.col_off = (int64_t)hk.s_host.size();   // where this block's eigenvalues start (insertion order)
    hk.blocks.push_back(bl);
    BGGE_CUDA(dS.alloc((size_t)nr));
    BGGE_TRY(eig_compact(dA.p, m, dW.p, m, nr, canonical_sign, dS.p, bl.dU, bl.ldu, g.stream));
    std::vector<double> s((size_t)nr);
    BGGE_CUDA(cudaMemcpyAsync(s.data(), dS.p, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    BGGE_CUDA(cudaStreamSynchronize(g.stream));
    hk.s_host.insert(hk.s_host.end(), s.begin(), s.end());
  }
  BGGE_REQUIRE(!hk.blocks.empty(), "kernel %d: no eigenvalue exceeds tol in any block", j);
  return OK;
}

// setDEC() of an *expanded* kernel without materialising it.  For a block of rows S with
// genotype codes g(i) and an all-ones mask, K_S = Z K0[T,T] Z' with Z the |S| x |T| incidence matrix and
// C = Z'Z = diag(counts).  If C^{1/2} K0[T,T] C^{1/2} = W diag(s) W' then K_S = U diag(s) U' with
// U = Z C^{-1/2} W (orthonormal columns), so the |T| x |T| problem gives the same eigenpairs (> tol) as
// eigen(K_S) (R/BGGE.R:112-116, 144-171).  Blocks whose mask is not constant fall back to the direct route.
int bgge_gibbs_add_expanded_kernel(bgge_gibbs* h, int j, int type, const double* K0, int64_t L, int64_t ldk,
                                   int on_device, const int32_t* gid, const int32_t* env, int64_t n_codes, int mode,
                                   int env_k, const int64_t* ne, int n_ne, double tol, int canonical_sign) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(!g.initialised, "gibbs_add_expanded_kernel: handle already initialised");
  BGGE_REQUIRE(n_codes == g.n, "gibbs_add_expanded_kernel: %lld genotype / environment codes for a model with n = %lld rows "
               "(non-conformable arguments)", (long long)n_codes, (long long)g.n);
  BGGE_REQUIRE(j >= 0 && j < g.nk && gid != nullptr && L > 0, "gibbs_add_expanded_kernel: bad arguments");
  BGGE_REQUIRE(type == BGGE_TYPE_D || type == BGGE_TYPE_BD, "Matrix should be of types BD or D");
  BGGE_REQUIRE(mode >= BGGE_EXPAND_G && mode <= BGGE_EXPAND_GI, "gibbs_add_expanded_kernel: unknown expansion mode %d", mode);
  BGGE_REQUIRE(mode == BGGE_EXPAND_GI || (K0 != nullptr && ldk >= L), "gibbs_add_expanded_kernel: K0 missing");
  BGGE_REQUIRE((mode != BGGE_EXPAND_GE && mode != BGGE_EXPAND_ENV) || env != nullptr, "gibbs_add_expanded_kernel: env codes missing");
  const int64_t n = g.n;
  BGGE_TRY(check_codes(gid, n, L, "gid"));
  cudaStream_t st = g.stream;

  // base kernel on the device + its diagonal on the host
  DevBuf<double> k0buf;
  const double* dK0 = nullptr;
  int64_t ld0 = L;
  std::vector<double> diag((size_t)L, 1.0);
  if (mode != BGGE_EXPAND_GI) {
    if (on_device) {
      dK0 = K0;
      ld0 = ldk;
    } else {
      BGGE_CUDA(k0buf.alloc((size_t)L * L));
      BGGE_TRY(host_copy_2d(k0buf.p, (size_t)L * sizeof(double), K0, (size_t)ldk * sizeof(double), (size_t)L * sizeof(double),
                            (size_t)L, cudaMemcpyHostToDevice));
      dK0 = k0buf.p;
    }
    BGGE_CUDA(cudaMemcpy2D(diag.data(), sizeof(double), dK0, (size_t)(ld0 + 1) * sizeof(double), sizeof(double), (size_t)L,
                           cudaMemcpyDeviceToHost));
  }
  auto mask = [&](int64_t a, int64_t b) -> bool {
    if (mode == BGGE_EXPAND_GE) return env[a] == env[b];
    if (mode == BGGE_EXPAND_ENV) return env[a] == env_k && env[b] == env_k;
    return true;
  };
  long double tr = 0.0L;   // mean(diag(K)) over all n rows (R/BGGE.R:249)
  for (int64_t i = 0; i < n; ++i) tr += mask(i, i) ? (long double)diag[(size_t)gid[i]] : 0.0L;
  BGGE_REQUIRE(tr > 0.0L, "gibbs_add_expanded_kernel: kernel %d has an all-zero diagonal", j);
  BGGE_TRY(bgge_gibbs_set_kernel(h, j, type, (double)(tr / (long double)n)));

  std::vector<std::pair<int64_t, int64_t>> spans;
  if (type == BGGE_TYPE_D) {
    spans.push_back({0, n});
  } else {
    BGGE_REQUIRE(ne != nullptr && n_ne > 1, "ne invalid. For type BD, number of subjects in each sub matrices should be provided");
    int64_t pos = 0;
    for (int k = 0; k < n_ne; ++k) {
      BGGE_REQUIRE(ne[k] > 0 && pos + ne[k] <= n, "gibbs_add_expanded_kernel: ne does not partition the rows of y");
      spans.push_back({pos, ne[k]});
      pos += ne[k];
    }
  }
  HostKernel& hk = g.kernels[(size_t)j];
  DevBuf<int> dgid_all, denv_all;   // only for the fallback route
  // Identical blocks (same genotypes with the same counts, e.g. every environment of a balanced trial) are
  // the same matrix: decompose it once, let the blocks share W and sweep them as one multi-RHS panel.
  const bool share_blocks = !getenv("BGGE_NO_SHARED_BLOCKS") && !getenv("BGGE_NO_FACTORED_SWEEP");
  std::map<std::pair<std::vector<int>, std::vector<int>>, int> key_count;
  auto block_key = [&](int64_t pos, int64_t m) {
    std::vector<int> cnt((size_t)L, 0), T, c;
    for (int64_t i = 0; i < m; ++i) ++cnt[(size_t)gid[pos + i]];
    for (int64_t gidx = 0; gidx < L; ++gidx)
      if (cnt[(size_t)gidx] > 0) {
        T.push_back((int)gidx);
        c.push_back(cnt[(size_t)gidx]);
      }
    return std::make_pair(T, c);
  };
  auto block_is_live = [&](int64_t pos, int64_t m) {   // uniform all-ones mask: the block will be decomposed
    if (mode == BGGE_EXPAND_GE) {
      for (int64_t i = 1; i < m; ++i)
        if (env[pos + i] != env[pos]) return false;
      return true;
    }
    if (mode == BGGE_EXPAND_ENV) {
      for (int64_t i = 0; i < m; ++i)
        if (env[pos + i] != env_k) return false;
      return true;
    }
    return true;
  };
  if (share_blocks && spans.size() > 1)
    for (auto& sp : spans)
      if (block_is_live(sp.first, sp.second)) ++key_count[block_key(sp.first, sp.second)];
  for (auto& sp : spans) {
    const int64_t pos = sp.first, m = sp.second;
    // is the mask constant on this block?  (first row decides; every pair is checked through the row classes)
    bool uniform = true, ones = true;
    if (mode == BGGE_EXPAND_GE) {
      for (int64_t i = 1; i < m && uniform; ++i) uniform = env[pos + i] == env[pos];
    } else if (mode == BGGE_EXPAND_ENV) {
      int64_t in = 0;
      for (int64_t i = 0; i < m; ++i) in += env[pos + i] == env_k;
      uniform = in == 0 || in == m;
      ones = in == m;
    }
    if (uniform && !ones) continue;   // all-zero block: no eigenvalue exceeds tol (R/BGGE.R:162)
    HostBlock bl;
    bl.pos0 = pos;
    bl.rows = m;
    bl.ldu = round_up(m, 16);
    std::vector<double> s_host;
    if (uniform) {
      // ---- factorised route
      std::vector<int> cnt((size_t)L, 0), slot((size_t)L, -1), T, loc((size_t)m);
      for (int64_t i = 0; i < m; ++i) ++cnt[(size_t)gid[pos + i]];
      for (int64_t gidx = 0; gidx < L; ++gidx)
        if (cnt[(size_t)gidx] > 0) {
          slot[(size_t)gidx] = (int)T.size();
          T.push_back((int)gidx);
        }
      const int64_t t = (int64_t)T.size();
      std::vector<double> sqc((size_t)t), isq((size_t)m);
      for (int64_t a = 0; a < t; ++a) sqc[(size_t)a] = std::sqrt((double)cnt[(size_t)T[(size_t)a]]);
      for (int64_t i = 0; i < m; ++i) {
        loc[(size_t)i] = slot[(size_t)gid[pos + i]];
        isq[(size_t)i] = 1.0 / sqc[(size_t)loc[(size_t)i]];
      }
      // row map of the block (every factorised block owns one, even when W is shared)
      int* d_loc = nullptr;
      double* d_isq = nullptr;
      auto upload_map = [&]() -> int {
        BGGE_CUDA(pool_alloc((void**)&d_loc, (size_t)m * sizeof(int)));
        BGGE_CUDA(pool_alloc((void**)&d_isq, (size_t)m * sizeof(double)));
        BGGE_CUDA(cudaMemcpyAsync(d_loc, loc.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, st));
        BGGE_CUDA(cudaMemcpyAsync(d_isq, isq.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
        return OK;
      };
      bool identity_map = t == m;   // every genotype once, rows in genotype order (then every weight is 1)
      for (int64_t i = 0; i < m && identity_map; ++i) identity_map = loc[(size_t)i] == (int)i;
      auto make_virtual = [&](double* W, int64_t ldw, int64_t nr, bool owns_w) -> int {
        std::vector<int> rowptr((size_t)t + 1, 0), rowidx((size_t)m);
        for (int64_t i = 0; i < m; ++i) ++rowptr[(size_t)loc[(size_t)i] + 1];
        for (int64_t a = 0; a < t; ++a) rowptr[(size_t)a + 1] += rowptr[(size_t)a];
        std::vector<int> fill(rowptr.begin(), rowptr.end() - 1);
        for (int64_t i = 0; i < m; ++i) rowidx[(size_t)fill[(size_t)loc[(size_t)i]]++] = (int)i;
        BGGE_CUDA(pool_alloc((void**)&bl.d_rowptr, ((size_t)t + 1) * sizeof(int)));
        BGGE_CUDA(pool_alloc((void**)&bl.d_rowidx, (size_t)m * sizeof(int)));
        BGGE_CUDA(cudaMemcpyAsync(bl.d_rowptr, rowptr.data(), ((size_t)t + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        BGGE_CUDA(cudaMemcpyAsync(bl.d_rowidx, rowidx.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, st));
        bl.dU = W;
        bl.ldu = ldw;
        bl.nr = nr;
        bl.vrows = t;
        bl.d_loc = d_loc;
        bl.d_isq = d_isq;
        bl.owns = owns_w;
        bl.owns_map = true;
        bl.identity = identity_map;
        bl.col_off = (int64_t)hk.s_host.size();   // where this block's eigenvalues start (insertion order)
    hk.blocks.push_back(bl);
        BGGE_CUDA(cudaStreamSynchronize(st));   // host staging vectors die at the end of this scope
        return OK;
      };
      std::vector<int> cntT((size_t)t);
      for (int64_t a = 0; a < t; ++a) cntT[(size_t)a] = cnt[(size_t)T[(size_t)a]];
      bool identity = t == m;   // every genotype once, rows in genotype order: U = W
      for (int64_t i = 0; i < m && identity; ++i) identity = loc[(size_t)i] == (int)i;
      const bool env_mode = mode == BGGE_EXPAND_ENV;
      const bool duplicated = share_blocks && spans.size() > 1 && key_count[std::make_pair(T, cntT)] > 1;
      // the environment kernels of MDe are one block each and, in a balanced trial, the same matrix: shared through
      // the handle-level cache (bgge::SharedDecomp) like the identical blocks of one BD kernel
      // Uniform counts (every genotype of T occurs c times: any balanced design): C^{1/2} K0[T,T] C^{1/2} = c K0[T,T].
      // Decompose K0[T,T] itself with the threshold tol / c and scale the eigenvalues by c afterwards; blocks and
      // KERNELS that differ only in c -- the G kernel (c = E) and the GE blocks (c = 1) of a balanced MDs model --
      // then share ONE decomposition (C5: one 10000^2 syevd instead of two).
      int ucount = cntT.empty() ? 0 : cntT[0];
      for (int64_t a = 1; a < t && ucount; ++a)
        if (cntT[(size_t)a] != ucount) ucount = 0;
      if (!share_blocks || getenv("BGGE_NO_SCALED_SHARING")) ucount = 0;
      const double eff_tol = ucount ? tol / (double)ucount : tol;
      const bool shareable = share_blocks && (duplicated || env_mode || ucount > 0);
      auto add_dense_shared = [&](double* W, int64_t ldw, int64_t nr, bool owns_w) -> int {
        bl.dU = W;
        bl.ldu = ldw;
        bl.nr = nr;
        bl.owns = owns_w;
        bl.col_off = (int64_t)hk.s_host.size();
        hk.blocks.push_back(bl);
        return OK;
      };
      DevBuf<int> dT;
      DevBuf<double> dsqc, dM, dW, dS, dWc, dwts;
      DevBuf<unsigned long long> dhash;
      BGGE_CUDA(dT.alloc((size_t)t));
      BGGE_CUDA(dsqc.alloc((size_t)t));
      BGGE_CUDA(dM.alloc((size_t)t * t));
      BGGE_CUDA(dW.alloc((size_t)t));
      BGGE_CUDA(cudaMemcpyAsync(dT.p, T.data(), (size_t)t * sizeof(int), cudaMemcpyHostToDevice, st));
      std::vector<double> sqc_gather = ucount ? std::vector<double>((size_t)t, 1.0) : sqc;
      BGGE_CUDA(cudaMemcpyAsync(dsqc.p, sqc_gather.data(), (size_t)t * sizeof(double), cudaMemcpyHostToDevice, st));
      BGGE_TRY(gather_scale_launch(dK0, ld0, dT.p, dsqc.p, t, dM.p, st));
      unsigned long long hh[2] = {0, 0};
      if (shareable) {
        BGGE_CUDA(dhash.alloc(2));
        BGGE_TRY(hash_bits_launch(dM.p, t * t, dhash.p, st));
        BGGE_CUDA(cudaMemcpyAsync(hh, dhash.p, sizeof(hh), cudaMemcpyDeviceToHost, st));
        BGGE_CUDA(cudaStreamSynchronize(st));
        const SharedDecomp* hit = nullptr;
        for (auto& d : g.decomps) {
          if (d.h1 != hh[0] || d.h2 != hh[1] || d.canonical != canonical_sign || d.T != T) continue;
          if (ucount > 0 && d.ucount > 0) {
            // the cached pairs are those above d.tol: enough if this block's threshold is not lower, or if
            // nothing between the two thresholds was truncated
            if (eff_tol >= d.tol || d.max_dropped <= eff_tol) hit = &d;
          } else if (ucount == 0 && d.ucount == 0 && d.tol == tol && d.cnt == cntT) {
            hit = &d;
          }
        }
        if (hit) {   // this matrix was decomposed already: share its W and eigenvalues
          int64_t nr_use = hit->nr;
          if (ucount > 0) {
            nr_use = 0;
            while (nr_use < hit->nr && hit->s[(size_t)nr_use] > eff_tol) ++nr_use;   // descending: a prefix
          }
          if (nr_use == 0) continue;
          if (env_mode && identity && !duplicated) {
            BGGE_TRY(add_dense_shared(hit->W, hit->ldw, nr_use, false));
          } else {
            BGGE_TRY(upload_map());
            BGGE_TRY(make_virtual(hit->W, hit->ldw, nr_use, false));
          }
          for (int64_t c = 0; c < nr_use; ++c) hk.s_host.push_back(ucount > 0 ? hit->s[(size_t)c] * (double)ucount : hit->s[(size_t)c]);
          continue;
        }
      }
      int64_t nr = 0;
      BGGE_TRY(eig_decompose(dM.p, t, t, eff_tol, dW.p, &nr, st));
      double max_dropped = -INFINITY;                               // syevd order is ascending: the value below the kept ones
      if (ucount > 0 && nr < t)
        BGGE_CUDA(cudaMemcpy(&max_dropped, dW.p + (t - nr - 1), sizeof(double), cudaMemcpyDeviceToHost));
      auto remember = [&](double* W, int64_t ldw_, int64_t nr_, const std::vector<double>& s_unscaled) {
        g.decomps.push_back(SharedDecomp{hh[0], hh[1], eff_tol, canonical_sign, T, cntT, W, ldw_, nr_, s_unscaled});
        g.decomps.back().ucount = ucount;
        g.decomps.back().max_dropped = max_dropped;
      };
      if (nr == 0) {                                                // R/BGGE.R:162
        if (shareable) remember(nullptr, 0, 0, {});
        continue;
      }
      bl.nr = nr;
      const int64_t ldw = round_up(t, 16);
      BGGE_CUDA(dS.alloc((size_t)nr));
      BGGE_CUDA(dWc.alloc((size_t)ldw * nr));
      BGGE_TRY(eig_compact(dM.p, t, dW.p, t, nr, 0, dS.p, dWc.p, ldw, st));
      double* dWfinal = nullptr;
      BGGE_CUDA(pool_alloc((void**)&dWfinal, (size_t)ldw * nr * sizeof(double)));
      if (canonical_sign) {
        // same convention as the dense route: the largest-|.| entry of the EXPANDED vector is positive,
        // i.e. the largest |W[a,c]| / sqrt(count_a)
        std::vector<double> wts((size_t)t);
        for (int64_t a = 0; a < t; ++a) wts[(size_t)a] = 1.0 / sqc[(size_t)a];
        BGGE_CUDA(dwts.alloc((size_t)t));
        BGGE_CUDA(cudaMemcpyAsync(dwts.p, wts.data(), (size_t)t * sizeof(double), cudaMemcpyHostToDevice, st));
        BGGE_TRY(sign_fix_launch(dWc.p, ldw, t, nr, dwts.p, dWfinal, st));
      } else {
        BGGE_CUDA(cudaMemcpyAsync(dWfinal, dWc.p, (size_t)ldw * nr * sizeof(double), cudaMemcpyDeviceToDevice, st));
      }
      s_host.resize((size_t)nr);
      BGGE_CUDA(cudaMemcpyAsync(s_host.data(), dS.p, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost, st));
      BGGE_CUDA(cudaStreamSynchronize(st));
      const std::vector<double> s_unscaled = s_host;
      if (ucount > 1)
        for (auto& v : s_host) v *= (double)ucount;
      const bool no_fact = getenv("BGGE_NO_FACTORED_SWEEP") != nullptr;
      const bool keep_factored = (duplicated || 3 * t <= 2 * m) && !no_fact;
      if (keep_factored) {   // fewer rows to stream (repeated genotypes) or W shared with identical blocks
        BGGE_TRY(upload_map());
        BGGE_TRY(make_virtual(dWfinal, ldw, nr, true));
        if (shareable) remember(dWfinal, ldw, nr, s_unscaled);
      } else if (shareable && identity) {   // U = W: the block streams W itself, later identical kernels share it
        BGGE_TRY(add_dense_shared(dWfinal, ldw, nr, true));
        remember(dWfinal, ldw, nr, s_unscaled);
      } else {
        BGGE_TRY(upload_map());
        BGGE_CUDA(pool_alloc((void**)&bl.dU, (size_t)bl.ldu * nr * sizeof(double)));
        bl.owns = true;
        bl.col_off = (int64_t)hk.s_host.size();   // where this block's eigenvalues start (insertion order)
        hk.blocks.push_back(bl);
        BGGE_TRY(expand_rows_launch(dWfinal, ldw, d_loc, d_isq, m, nr, bl.dU, bl.ldu, st));
        BGGE_CUDA(cudaStreamSynchronize(st));
        pool_free(dWfinal);
        pool_free(d_loc);
        pool_free(d_isq);
      }
    } else {
      // ---- mixed mask inside the block: materialise the m x m block and decompose it directly
      if (!dgid_all.p) {
        BGGE_CUDA(dgid_all.alloc((size_t)n));
        BGGE_CUDA(cudaMemcpy(dgid_all.p, gid, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        BGGE_CUDA(denv_all.alloc((size_t)n));
        BGGE_CUDA(cudaMemcpy(denv_all.p, env, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
      }
      DevBuf<double> dA, dW, dS;
      BGGE_CUDA(dA.alloc((size_t)m * m));
      BGGE_CUDA(dW.alloc((size_t)m));
      BGGE_TRY(expand_launch(dK0, ld0, L, dgid_all.p + pos, denv_all.p + pos, m, mode, env_k, dA.p, m, st));
      int64_t nr = 0;
      BGGE_TRY(eig_decompose(dA.p, m, m, tol, dW.p, &nr, st));
      if (nr == 0) continue;
      bl.nr = nr;
      BGGE_CUDA(pool_alloc((void**)&bl.dU, (size_t)bl.ldu * nr * sizeof(double)));
      bl.owns = true;
      bl.col_off = (int64_t)hk.s_host.size();   // where this block's eigenvalues start (insertion order)
    hk.blocks.push_back(bl);
      BGGE_CUDA(dS.alloc((size_t)nr));
      BGGE_TRY(eig_compact(dA.p, m, dW.p, m, nr, canonical_sign, dS.p, bl.dU, bl.ldu, st));
      s_host.resize((size_t)nr);
      BGGE_CUDA(cudaMemcpyAsync(s_host.data(), dS.p, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost, st));
      BGGE_CUDA(cudaStreamSynchronize(st));
    }
    hk.s_host.insert(hk.s_host.end(), s_host.begin(), s_host.end());
  }
  BGGE_REQUIRE(!hk.blocks.empty(), "kernel %d: no eigenvalue exceeds tol in any block", j);
  return OK;
}

// eigen blocks held by the handle (tests, reuse of a decomposition)
int bgge_gibbs_block_count(bgge_gibbs* h, int j, int* n_blocks) {
  BGGE_HANDLE(h);
  BGGE_REQUIRE(j >= 0 && j < h->g.nk && n_blocks, "gibbs_block_count: bad arguments");
  *n_blocks = (int)h->g.kernels[(size_t)j].blocks.size();
  return OK;
}

int bgge_gibbs_block_device(bgge_gibbs* h, int j, int b, const double** dU, int64_t* ldu) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(j >= 0 && j < g.nk && dU && ldu, "gibbs_block_device: bad arguments");
  HostKernel& hk = g.kernels[(size_t)j];
  BGGE_REQUIRE(b >= 0 && b < (int)hk.blocks.size(), "gibbs_block_device: bad block index");
  HostBlock& bl = hk.blocks[(size_t)b];
  BGGE_TRY(ensure_dense(bl, g.stream));        // a factorised block is expanded on first request
  BGGE_CUDA(cudaStreamSynchronize(g.stream));
  *dU = bl.vrows > 0 ? bl.dU_dense : bl.dU;
  *ldu = bl.vrows > 0 ? bl.ldu_dense : bl.ldu;
  return OK;
}

int bgge_gibbs_fetch_block(bgge_gibbs* h, int j, int b, int64_t* pos0, int64_t* rows, int64_t* nr, double* s, double* U,
                           int64_t ldu) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(j >= 0 && j < g.nk, "gibbs_fetch_block: bad kernel index");
  HostKernel& hk = g.kernels[(size_t)j];
  BGGE_REQUIRE(b >= 0 && b < (int)hk.blocks.size(), "gibbs_fetch_block: bad block index");
  HostBlock& bl = hk.blocks[(size_t)b];
  if (pos0) *pos0 = bl.pos0;
  if (rows) *rows = bl.rows;
  if (nr) *nr = bl.nr;
  if (s) {
    std::memcpy(s, hk.s_host.data() + bl.col_off, (size_t)bl.nr * sizeof(double));   // col_off tracks s_host
  }
  if (U) {
    BGGE_REQUIRE(ldu >= bl.rows, "gibbs_fetch_block: ldu too small");
    BGGE_TRY(ensure_dense(bl, g.stream));
    BGGE_CUDA(cudaStreamSynchronize(g.stream));
    BGGE_CUDA(cudaMemcpy2D(U, (size_t)ldu * sizeof(double), bl.vrows > 0 ? bl.dU_dense : bl.dU,
                           (size_t)(bl.vrows > 0 ? bl.ldu_dense : bl.ldu) * sizeof(double),
                           (size_t)bl.rows * sizeof(double), (size_t)bl.nr, cudaMemcpyDeviceToHost));
  }
  return OK;
}

int bgge_gibbs_set_tape(bgge_gibbs* h, const double* normals, int64_t n_normals, const double* chisq, int64_t n_chisq) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(normals && chisq && n_normals > 0 && n_chisq > 0, "gibbs_set_tape: empty tape");
  // one spare slot: the intercept of iteration ite+1 is drawn ahead and never used
  BGGE_CUDA(g.ztape.alloc((size_t)n_normals + 1));
  BGGE_CUDA(g.ctape.alloc((size_t)n_chisq + 1));
  BGGE_CUDA(cudaMemset(g.ztape.p + n_normals, 0, sizeof(double)));
  BGGE_CUDA(cudaMemset(g.ctape.p + n_chisq, 0, sizeof(double)));
  BGGE_CUDA(cudaMemcpy(g.ztape.p, normals, (size_t)n_normals * sizeof(double), cudaMemcpyHostToDevice));
  BGGE_CUDA(cudaMemcpy(g.ctape.p, chisq, (size_t)n_chisq * sizeof(double), cudaMemcpyHostToDevice));
  g.zt_len = n_normals;
  g.ct_len = n_chisq;
  return OK;
}

int bgge_gibbs_init(bgge_gibbs* h, int64_t ite, int64_t burn, int64_t thin, double R2, int rng_mode, uint64_t seed,
                    uint32_t chain) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(!g.initialised, "gibbs_init: handle already initialised");
  BGGE_REQUIRE(rng_mode == BGGE_RNG_PHILOX || rng_mode == BGGE_RNG_TAPE, "gibbs_init: unknown rng mode %d", rng_mode);
  BGGE_REQUIRE(rng_mode != BGGE_RNG_TAPE || g.zt_len > 0, "gibbs_init: tape mode without a tape");
  g.ite = ite;
  g.burn = burn;
  g.thin = thin;
  g.R2 = R2;
  g.draw.mode = rng_mode;
  g.draw.seed = seed;
  g.draw.chain = chain;
  return gibbs_init(&g);
}

int bgge_gibbs_run(bgge_gibbs* h, int64_t iters) {
  BGGE_HANDLE(h);
  return gibbs_run(&h->g, iters);
}

// ---------------------------------------------------------------------- column-sharded single chain
// Every process holds the full state (r, u, scalars: a few n-vectors) and sweeps only its slice of the eigen-columns
// of every basis matrix; draws are keyed by the global column, so the chain does not depend on the number of
// processes except for the order of the sums.  Per kernel step the partial W b and sum b^2/s of all processes are
// added up by the caller (NCCL all-reduce in the Python mirror) between shard_begin and shard_end.
int bgge_gibbs_set_shard(bgge_gibbs* h, int rank, int world) {
  BGGE_HANDLE(h);
  BGGE_REQUIRE(!h->g.initialised, "gibbs_set_shard: handle already initialised");
  BGGE_REQUIRE(world >= 1 && rank >= 0 && rank < world, "gibbs_set_shard: rank %d of %d", rank, world);
  h->g.shard_rank = rank;
  h->g.shard_world = world;
  return OK;
}

int bgge_gibbs_shard_begin(bgge_gibbs* h, int j, double** d_vector, int64_t* count) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised && g.shard_world > 1, "gibbs_shard_begin: handle is not a column shard");
  BGGE_REQUIRE(g.shard_step == -1 && j >= 0 && j < g.nk, "gibbs_shard_begin: step %d out of order", j);
  BGGE_REQUIRE(g.done < g.ite, "gibbs_shard_begin: all %lld iterations are done", (long long)g.ite);
  HostKernel& hk = g.kernels[(size_t)j];
  BGGE_TRY(launch_fused(&g, hk, j));
  BGGE_TRY(launch_shard_sum(&g, hk));
  g.shard_step = j;
  // a kernel whose blocks are all factorised exchanges only the genotype-space part (+ the sum b^2/s behind it)
  bool all_virtual = true;
  for (auto& bl : hk.blocks) all_virtual = all_virtual && bl.vrows > 0;
  if (d_vector) *d_vector = hk.xbuf.p + (all_virtual ? g.n : 0);
  if (count) *count = (all_virtual ? 0 : g.n) + hk.vtotal + 1;
  return OK;
}

int bgge_gibbs_shard_end(bgge_gibbs* h, int j) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised && g.shard_step == j, "gibbs_shard_end: step %d was not begun", j);
  BGGE_TRY(launch_shard_apply(&g, g.kernels[(size_t)j]));
  g.shard_step = -1;
  return OK;
}

int bgge_gibbs_shard_iteration_end(bgge_gibbs* h) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised && g.shard_world > 1 && g.shard_step == -1, "gibbs_shard_iteration_end: a step is still open");
  BGGE_TRY(launch_scalar_sharded(&g));
  g.done += 1;
  return OK;
}

// Iteration bookkeeping for callers that capture one sharded iteration into a CUDA graph (NCCL collective
// included) and replay it: the captured calls are not executed, the replays are not seen by the library.
int bgge_gibbs_shard_advance(bgge_gibbs* h, int64_t iterations) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised && g.shard_world > 1 && g.shard_step == -1, "gibbs_shard_advance: not a column shard at an iteration boundary");
  BGGE_REQUIRE(g.done + iterations >= 0 && g.done + iterations <= g.ite, "gibbs_shard_advance: %lld + %lld outside [0, %lld]",
               (long long)g.done, (long long)iterations, (long long)g.ite);
  g.done += iterations;
  return OK;
}

// The same chain with the collective inside the library: NCCL all-reduce of the step vector on the handle's stream
// (stream-ordered between the shard_sum and the apply kernels, no host synchronisation anywhere in the loop).
int bgge_gibbs_set_comm(bgge_gibbs* h, bgge_comm* c) {
  BGGE_HANDLE(h);
  BGGE_REQUIRE(c != nullptr, "gibbs_set_comm: NULL communicator");
  BGGE_REQUIRE(!h->g.initialised, "gibbs_set_comm: handle already initialised");
  BGGE_REQUIRE(c->device == h->g.device, "gibbs_set_comm: communicator lives on device %d, the handle on %d", c->device,
               h->g.device);
  h->g.shard_rank = c->rank;
  h->g.shard_world = c->world;
  h->g.comm = c;
  return OK;
}

int bgge_gibbs_run_sharded(bgge_gibbs* h, int64_t iters) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised && g.comm != nullptr, "gibbs_run_sharded: call bgge_gibbs_set_comm before bgge_gibbs_init");
  BGGE_REQUIRE(iters >= 0 && g.done + iters <= g.ite, "gibbs_run_sharded: %lld + %lld iterations exceed ite = %lld",
               (long long)g.done, (long long)iters, (long long)g.ite);
  if (g.shard_world == 1) return gibbs_run(&g, iters);
  for (int64_t i = 0; i < iters; ++i) {
    for (int j = 0; j < g.nk; ++j) {
      double* vec = nullptr;
      int64_t cnt = 0;
      BGGE_TRY(bgge_gibbs_shard_begin(h, j, &vec, &cnt));
      BGGE_TRY(comm_allreduce_sum(static_cast<bgge_comm*>(g.comm), vec, cnt, g.stream));
      BGGE_TRY(bgge_gibbs_shard_end(h, j));
    }
    BGGE_TRY(bgge_gibbs_shard_iteration_end(h));
  }
  return OK;
}

int bgge_gibbs_set_mode(bgge_gibbs* h, int mode) {
  BGGE_HANDLE(h);
  BGGE_REQUIRE(!h->g.initialised, "gibbs_set_mode: handle already initialised");
  BGGE_REQUIRE(mode >= 0 && mode <= 2, "gibbs_set_mode: mode must be 0 (two-pass), 1 (resident / fused) or 2 (fused)");
  h->g.mode = mode;
  return OK;
}

// dst's kernel j_dst borrows every eigen block (dense or factorised) of src's kernel j_src: one
// decomposition, many chains.  src must outlive dst.
int bgge_gibbs_share_kernel(bgge_gibbs* dst, int j_dst, bgge_gibbs* src, int j_src) {
  BGGE_HANDLE(dst);
  BGGE_REQUIRE(src != nullptr, "gibbs_share_kernel: NULL source handle");
  Gibbs& d = dst->g;
  Gibbs& sg = src->g;
  BGGE_REQUIRE(!d.initialised, "gibbs_share_kernel: destination already initialised");
  BGGE_REQUIRE(j_dst >= 0 && j_dst < d.nk && j_src >= 0 && j_src < sg.nk && d.n == sg.n && d.device == sg.device,
               "gibbs_share_kernel: kernel index out of range or handles of different size / device");
  HostKernel& hd = d.kernels[(size_t)j_dst];
  const HostKernel& hs = sg.kernels[(size_t)j_src];
  BGGE_REQUIRE(hd.blocks.empty() && !hs.blocks.empty(), "gibbs_share_kernel: destination kernel not empty or source kernel empty");
  BGGE_CUDA(cudaStreamSynchronize(sg.stream));
  for (const HostBlock& b : hs.blocks) {
    HostBlock c = b;
    c.owns = false;
    c.owns_dense = false;
    c.owns_map = false;
    hd.blocks.push_back(c);
  }
  hd.s_host = hs.s_host;
  hd.mean_diag = hs.mean_diag;
  return OK;
}

int bgge_gibbs_kernel_info(bgge_gibbs* h, int j, int* fused, int* cluster_size, int* clusters, int* launches) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised && j >= 0 && j < g.nk, "gibbs_kernel_info: bad kernel index or handle not initialised");
  HostKernel& hk = g.kernels[(size_t)j];
  if (fused) *fused = hk.fused ? (hk.fused_engine == 1 ? 2 : 1) : 0;
  if (cluster_size) *cluster_size = hk.fused ? hk.fused_cs : hk.cs;
  if (clusters) *clusters = hk.fused ? hk.fused_clusters : hk.n_btasks;
  if (launches) {
    *launches = 2 + (hk.fused && hk.n_vdesc > 0 ? 1 : 0);   // [gather] + sweep + finalize (identity row maps need no gather)
    // a merged unit is launched once, accounted to its first kernel
    if (!g.units.empty() && g.unit_of[(size_t)j] >= 0 && g.units[(size_t)g.unit_of[(size_t)j]].members[0] != j) *launches = 0;
  }
  return OK;
}

// after init: does the model run in the resident (on-chip basis, one launch per run) kernel?
int bgge_gibbs_resident_info(bgge_gibbs* h, int* enabled, int* ctas, int64_t* smem_bytes) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised, "gibbs_resident_info: handle not initialised");
  if (enabled) *enabled = g.resident.enabled ? 1 : 0;
  if (ctas) *ctas = g.resident.enabled ? g.resident.ncta : 0;
  if (smem_bytes) *smem_bytes = g.resident.enabled ? (int64_t)g.resident.smem : 0;
  return OK;
}

// after init: merged units of the streaming path (kernels sharing one matrix on disjoint rows, one launch each)
int bgge_gibbs_merged_info(bgge_gibbs* h, int* units, int* kernels_in_units) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised, "gibbs_merged_info: handle not initialised");
  int nk = 0;
  for (auto& u : g.units) nk += (int)u.members.size();
  if (units) *units = (int)g.units.size();
  if (kernels_in_units) *kernels_in_units = nk;
  return OK;
}

// bytes of eigen basis one kernel step of kernel j streams from HBM (W for factorised blocks; two-pass
// steps read it twice) and how many of its blocks are kept factorised
int bgge_gibbs_kernel_bytes(bgge_gibbs* h, int j, int64_t* basis_bytes, int* factored_blocks) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised && j >= 0 && j < g.nk, "gibbs_kernel_bytes: bad kernel index or handle not initialised");
  HostKernel& hk = g.kernels[(size_t)j];
  int64_t bytes = 0;
  int nv = 0;
  std::vector<const double*> seen;   // fused: a W shared by identical blocks is streamed once per <= 4 of them
  std::vector<int> uses;
  for (auto& b : hk.blocks) {
    if (!hk.fused) {
      bytes += 16 * b.rows * b.nr;
      continue;
    }
    nv += b.vrows > 0 ? 1 : 0;
    bool counted = false;
    if (b.vrows > 0)
      for (size_t k = 0; k < seen.size(); ++k)
        if (seen[k] == b.dU && uses[k] < 4) {
          ++uses[k];
          counted = true;
          break;
        }
    if (!counted) {
      bytes += 8 * b.stream_rows() * b.nr;
      if (b.vrows > 0) {
        seen.push_back(b.dU);
        uses.push_back(1);
      }
    }
  }
  // member of a merged unit (kernels sharing one matrix on disjoint rows, swept by one launch): the matrix is
  // streamed once per unit, accounted to the unit's first kernel
  if (!g.units.empty() && g.unit_of[(size_t)j] >= 0 && g.units[(size_t)g.unit_of[(size_t)j]].members[0] != j) bytes = 0;
  if (basis_bytes) *basis_bytes = bytes;
  if (factored_blocks) *factored_blocks = nv;
  return OK;
}

int bgge_gibbs_profile(bgge_gibbs* h, int64_t iters, double* ms) {
  BGGE_HANDLE(h);
  BGGE_REQUIRE(ms != nullptr, "gibbs_profile: NULL output");
  return gibbs_profile(&h->g, iters, ms);
}

int bgge_gibbs_sync(bgge_gibbs* h) {
  BGGE_HANDLE(h);
  BGGE_CUDA(cudaStreamSynchronize(h->g.stream));
  return OK;
}

int bgge_gibbs_sizes(bgge_gibbs* h, int64_t* nr, int64_t* n_cum, int64_t* n_na, int64_t* tape_normals,
                     int64_t* tape_chisq) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  int64_t sum_nr = 0;
  for (int j = 0; j < g.nk; ++j) {
    int64_t t = 0;
    for (auto& b : g.kernels[(size_t)j].blocks) t += b.nr;
    if (nr) nr[j] = t;
    sum_nr += t;
  }
  int64_t nna = g.n_na;
  if (!g.initialised) {
    nna = 0;
    for (int64_t i = 0; i < (int64_t)g.y_host.size(); ++i)
      nna += g.na_host.size() == g.y_host.size() ? (g.na_host[(size_t)i] != 0) : std::isnan(g.y_host[(size_t)i]);
  }
  if (n_cum) *n_cum = g.initialised ? g.ncum : 0;
  if (n_na) *n_na = nna;
  const int64_t ite = g.ite;
  if (tape_normals) *tape_normals = (int64_t)g.nk * g.n + ite * (1 + g.q + sum_nr + nna);
  if (tape_chisq) *tape_chisq = ite * (g.nk + 1);
  return OK;
}

int bgge_gibbs_state(bgge_gibbs* h, double* mu, double* sigsq, double* sigb, double* u, double* e, double* bet) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised, "gibbs_state: handle not initialised");
  BGGE_CUDA(cudaStreamSynchronize(g.stream));
  Scal sc;
  BGGE_CUDA(cudaMemcpy(&sc, g.scal.p, sizeof(Scal), cudaMemcpyDeviceToHost));
  // The scalar step has already drawn the intercept of iteration done+1 (sc.mu / sc.shift); what R
  // holds after iteration `done` is mu_done and temp = r - (mu_done - mu0).
  if (sigsq) *sigsq = sc.sigsq;
  if (sigb) std::memcpy(sigb, sc.sigb, sizeof(double) * (size_t)g.nk);
  if (mu) *mu = sc.mu_done;
  if (u) BGGE_CUDA(cudaMemcpy(u, g.u.p, (size_t)g.n * g.nk * sizeof(double), cudaMemcpyDeviceToHost));
  if (e) {
    BGGE_CUDA(cudaMemcpy(e, g.r.p, (size_t)g.n * sizeof(double), cudaMemcpyDeviceToHost));
    const double sh = sc.mu_done - sc.mu0;
    for (int64_t i = 0; i < g.n; ++i) e[i] -= sh;
  }
  if (bet && g.q > 0) BGGE_CUDA(cudaMemcpy(bet, g.bet.p, (size_t)g.q * sizeof(double), cudaMemcpyDeviceToHost));
  return OK;
}

int bgge_gibbs_chain(bgge_gibbs* h, double* mu, double* varE, double* varU, double* bet) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised, "gibbs_chain: handle not initialised");
  BGGE_CUDA(cudaStreamSynchronize(g.stream));
  const size_t nc = (size_t)g.ncum;
  if (nc == 0) return OK;
  if (mu) BGGE_CUDA(cudaMemcpy(mu, g.chain_mu.p, nc * sizeof(double), cudaMemcpyDeviceToHost));
  if (varE) BGGE_CUDA(cudaMemcpy(varE, g.chain_varE.p, nc * sizeof(double), cudaMemcpyDeviceToHost));
  if (varU) BGGE_CUDA(cudaMemcpy(varU, g.chain_varU.p, nc * g.nk * sizeof(double), cudaMemcpyDeviceToHost));
  if (bet && g.q > 0) BGGE_CUDA(cudaMemcpy(bet, g.chain_bet.p, nc * g.q * sizeof(double), cudaMemcpyDeviceToHost));
  return OK;
}

int bgge_gibbs_summary(bgge_gibbs* h, double* yHat, double* varE, double* varE_sd, double* u, double* u_sd,
                       double* varu, double* varu_sd, double* y) {
  BGGE_HANDLE(h);
  Gibbs& g = h->g;
  BGGE_REQUIRE(g.initialised, "gibbs_summary: handle not initialised");
  BGGE_CUDA(cudaStreamSynchronize(g.stream));
  const int64_t n = g.n;
  const size_t nc = (size_t)g.ncum;
  std::vector<double> cmu(nc), cve(nc), cvu(nc * g.nk), cbet(nc * std::max(1, g.q));
  BGGE_TRY(bgge_gibbs_chain(h, cmu.data(), cve.data(), cvu.data(), cbet.data()));
  // draw <- seq(ite)[seq(ite) %% thin == 0] > burn   (R/BGGE.R:408), limited to what has been run
  std::vector<size_t> kept;
  for (size_t slot = 0; slot < nc; ++slot) {
    const int64_t it = (int64_t)(slot + 1) * g.thin;
    if (it > g.burn && it <= g.done) kept.push_back(slot);
  }
  auto pick = [&](const double* src) {
    std::vector<double> v;
    v.reserve(kept.size());
    for (size_t s : kept) v.push_back(src[s]);
    return v;
  };
  const size_t k = kept.size();
  const double mu_est = mean_sd(pick(cmu.data()), nullptr);
  double sd = NAN;
  const double ve = mean_sd(pick(cve.data()), &sd);
  if (varE) *varE = ve;
  if (varE_sd) *varE_sd = sd;
  for (int j = 0; j < g.nk; ++j) {
    double sdj = NAN;
    const double m = mean_sd(pick(cvu.data() + (size_t)j * nc), &sdj);
    if (varu) varu[j] = m;
    if (varu_sd) varu_sd[j] = sdj;
  }
  std::vector<double> um, m2;
  if (yHat || u || u_sd) {
    um.resize((size_t)n * g.nk);
    BGGE_CUDA(cudaMemcpy(um.data(), g.umean.p, um.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (k == 0) std::fill(um.begin(), um.end(), NAN);
    if (u) std::memcpy(u, um.data(), um.size() * sizeof(double));
  }
  if (u_sd) {
    m2.resize((size_t)n * g.nk);
    BGGE_CUDA(cudaMemcpy(m2.data(), g.um2.p, m2.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < m2.size(); ++i) u_sd[i] = k > 1 ? std::sqrt(m2[i] / (double)(k - 1)) : NAN;
  }
  if (yHat) {
    std::vector<double> bmean((size_t)std::max(1, g.q), 0.0);
    for (int c = 0; c < g.q; ++c) {
      long double a = 0.0L;
      for (size_t s : kept) a += cbet[s * (size_t)g.q + (size_t)c];
      bmean[(size_t)c] = k ? (double)(a / (long double)k) : NAN;
    }
    for (int64_t i = 0; i < n; ++i) {
      double v = mu_est;                                            // R/BGGE.R:410-411
      if (g.q > 0) {
        double xb = 0.0;
        for (int c = 0; c < g.q; ++c) xb += g.xf_host[(size_t)(i + (int64_t)c * n)] * bmean[(size_t)c];
        v += xb;                                                    // R/BGGE.R:414-415
      }
      double us = 0.0;
      for (int j = 0; j < g.nk; ++j) us += um[(size_t)j * n + (size_t)i];
      yHat[i] = v + us;                                             // R/BGGE.R:418-419
    }
  }
  if (y) BGGE_CUDA(cudaMemcpy(y, g.y.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return OK;
}

// ====================================================================== batched chains / folds
int bgge_batch_create(int64_t n, int nk, int n_chains, void* stream, bgge_batch** out) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(out != nullptr, "batch_create: NULL output");
  BGGE_REQUIRE(n > 1 && nk >= 1 && nk <= MAXK && n_chains >= 1 && n_chains <= 4096 && n * (int64_t)n_chains < (1ll << 40),
               "batch_create: bad sizes n=%lld nk=%d chains=%d", (long long)n, nk, n_chains);
  std::unique_ptr<bgge_batch> h(new (std::nothrow) bgge_batch());
  BGGE_REQUIRE(h != nullptr, "batch_create: out of host memory");
  Batch& g = h->g;
  g.n = n;
  g.nk = nk;
  g.B = n_chains;
  g.kernels.resize((size_t)nk);
  BGGE_CUDA(cudaGetDevice(&g.device));
  if (stream) {
    g.stream = as_stream(stream);
  } else {
    BGGE_CUDA(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    g.own_stream = true;
  }
  *out = h.release();
  return OK;
}

void bgge_batch_destroy(bgge_batch* h) { delete h; }

int bgge_batch_set_y(bgge_batch* h, const double* Y, const uint8_t* na_mask) {
  BGGE_HANDLE(h);
  BGGE_REQUIRE(Y != nullptr, "batch_set_y: NULL Y");
  Batch& g = h->g;
  g.y_host.assign(Y, Y + g.n * g.B);
  if (na_mask) g.na_host.assign(na_mask, na_mask + g.n * g.B); else g.na_host.clear();
  return OK;
}

// XF: n x q column-major, one design matrix for all chains (R/BGGE.R:208-212, 284-290); q = 0 removes it
int bgge_batch_set_xf(bgge_batch* h, const double* XF, int q) {
  BGGE_HANDLE(h);
  Batch& g = h->g;
  BGGE_REQUIRE(!g.initialised, "batch_set_xf: handle already initialised");
  BGGE_REQUIRE(q >= 0 && q <= BATCH_MAXQ, "batch_set_xf: q = %d outside [0, %d]", q, BATCH_MAXQ);
  BGGE_REQUIRE(q == 0 || XF != nullptr, "batch_set_xf: NULL XF");
  g.q = q;
  if (q > 0) g.xf_host.assign(XF, XF + g.n * (int64_t)q); else g.xf_host.clear();
  return OK;
}

int bgge_batch_set_kernel(bgge_batch* h, int j, int type, double mean_diag) {
  BGGE_HANDLE(h);
  Batch& g = h->g;
  BGGE_REQUIRE(j >= 0 && j < g.nk, "batch_set_kernel: kernel index %d outside [0,%d)", j, g.nk);
  BGGE_REQUIRE(type == BGGE_TYPE_D || type == BGGE_TYPE_BD, "Matrix should be of types BD or D");
  BGGE_REQUIRE(mean_diag > 0.0 && std::isfinite(mean_diag), "batch_set_kernel: mean(diag(K)) must be positive");
  g.kernels[(size_t)j].mean_diag = mean_diag;
  return OK;
}

int bgge_batch_add_block(bgge_batch* h, int j, int64_t pos0, int64_t rows, int64_t nr, const double* s, const double* U,
                         int64_t ldu, int on_device) {
  BGGE_HANDLE(h);
  Batch& g = h->g;
  BGGE_REQUIRE(!g.initialised, "batch_add_block: handle already initialised");
  BGGE_REQUIRE(j >= 0 && j < g.nk, "batch_add_block: kernel index %d outside [0,%d)", j, g.nk);
  BGGE_REQUIRE(s && U && rows > 0 && nr > 0 && pos0 >= 0 && pos0 + rows <= g.n && ldu >= rows,
               "batch_add_block: bad block pos0=%lld rows=%lld nr=%lld ldu=%lld", (long long)pos0, (long long)rows,
               (long long)nr, (long long)ldu);
  for (int64_t c = 0; c < nr; ++c) BGGE_REQUIRE(s[c] > 0.0, "batch_add_block: eigenvalue %lld is not positive", (long long)c);
  BatchKernel& bk = g.kernels[(size_t)j];
  BatchBlock bl;
  bl.pos0 = pos0;
  bl.rows = rows;
  bl.nr = nr;
  if (on_device) {
    BGGE_REQUIRE(ldu % 2 == 0 && (reinterpret_cast<uintptr_t>(U) & 15u) == 0,
                 "batch_add_block: device U needs an even leading dimension and a 16-byte aligned base");
    bl.dU = const_cast<double*>(U);
    bl.ldu = ldu;
  } else {
    bl.ldu = round_up(rows, 16);
    BGGE_CUDA(pool_alloc((void**)&bl.dU, (size_t)bl.ldu * nr * sizeof(double)));
    bl.owns = true;
    bl.col_off = (int64_t)bk.s_host.size();
    bk.blocks.push_back(bl);
    bk.s_host.insert(bk.s_host.end(), s, s + nr);
    BGGE_CUDA(cudaMemset(bl.dU, 0, (size_t)bl.ldu * nr * sizeof(double)));
    BGGE_CUDA(cudaMemcpy2D(bl.dU, (size_t)bl.ldu * sizeof(double), U, (size_t)ldu * sizeof(double),
                           (size_t)rows * sizeof(double), (size_t)nr, cudaMemcpyHostToDevice));
    return OK;
  }
  bl.col_off = (int64_t)bk.s_host.size();
  bk.blocks.push_back(bl);
  bk.s_host.insert(bk.s_host.end(), s, s + nr);
  return OK;
}

int bgge_batch_init(bgge_batch* h, int64_t ite, int64_t burn, int64_t thin, double R2, uint64_t seed, uint32_t chain0) {
  BGGE_HANDLE(h);
  Batch& g = h->g;
  BGGE_REQUIRE(!g.initialised, "batch_init: handle already initialised");
  g.ite = ite;
  g.burn = burn;
  g.thin = thin;
  g.R2 = R2;
  g.draw = DrawCfg{};
  g.draw.mode = 0;
  g.draw.seed = seed;
  g.draw.chain = chain0;
  return batch_init(&g);
}

int bgge_batch_run(bgge_batch* h, int64_t iters) {
  BGGE_HANDLE(h);
  return batch_run(&h->g, iters);
}

int bgge_batch_sync(bgge_batch* h) {
  BGGE_HANDLE(h);
  BGGE_CUDA(cudaStreamSynchronize(h->g.stream));
  return OK;
}

int bgge_batch_info(bgge_batch* h, int* padded_chains, int64_t* gemm_flops_per_iter, int* launches_per_iter) {
  BGGE_HANDLE(h);
  Batch& g = h->g;
  BGGE_REQUIRE(g.initialised, "batch_info: handle not initialised");
  if (padded_chains) *padded_chains = (g.Bq + g.tn - 1) / g.tn * g.tn;   // chain columns the GEMM tiles compute
  int64_t fl = 0;                                                        // useful flops: the B real chains
  for (auto& k : g.kernels)
    for (auto& b : k.blocks) fl += 4 * b.rows * b.nr * (int64_t)g.B;
  if (gemm_flops_per_iter) *gemm_flops_per_iter = fl;
  // per kernel: GEMM, draw, GEMM, update; tail: variances, imputation, intercept; fixed effects: reduce, draw, apply
  if (launches_per_iter) *launches_per_iter = 4 * g.nk + 3 + (g.q > 0 ? 3 : 0);
  return OK;
}

// chains: mu[B x ncum], varE[B x ncum] chain-major; varU[nk x B x ncum]
int bgge_batch_chain(bgge_batch* h, double* mu, double* varE, double* varU) {
  BGGE_HANDLE(h);
  Batch& g = h->g;
  BGGE_REQUIRE(g.initialised, "batch_chain: handle not initialised");
  BGGE_CUDA(cudaStreamSynchronize(g.stream));
  const size_t nc = (size_t)g.ncum, Bp = (size_t)g.Bp, B = (size_t)g.B;
  if (nc == 0) return OK;
  std::vector<double> t(nc * Bp);
  auto pull = [&](const double* src, double* dst) -> int {
    BGGE_CUDA(cudaMemcpy(t.data(), src, nc * Bp * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < B; ++b)
      for (size_t k = 0; k < nc; ++k) dst[b * nc + k] = t[k * Bp + b];
    return OK;
  };
  if (mu) BGGE_TRY(pull(g.chain_mu.p, mu));
  if (varE) BGGE_TRY(pull(g.chain_varE.p, varE));
  if (varU)
    for (int j = 0; j < g.nk; ++j) BGGE_TRY(pull(g.chain_varU.p + (size_t)j * nc * Bp, varU + (size_t)j * B * nc));
  return OK;
}

// thinned chain of the fixed effects: bet[B x ncum x q] (chain-major, then recorded draw, then coefficient)
int bgge_batch_chain_bet(bgge_batch* h, double* bet) {
  BGGE_HANDLE(h);
  Batch& g = h->g;
  BGGE_REQUIRE(g.initialised, "batch_chain_bet: handle not initialised");
  BGGE_REQUIRE(bet != nullptr, "batch_chain_bet: NULL output");
  BGGE_CUDA(cudaStreamSynchronize(g.stream));
  const size_t nc = (size_t)g.ncum, Bp = (size_t)g.Bp, B = (size_t)g.B, q = (size_t)g.q;
  if (nc == 0 || q == 0) return OK;
  std::vector<double> t(nc * q * Bp);
  BGGE_CUDA(cudaMemcpy(t.data(), g.chain_bet.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (size_t b = 0; b < B; ++b)
    for (size_t k = 0; k < nc; ++k)
      for (size_t c = 0; c < q; ++c) bet[(b * nc + k) * q + c] = t[(k * q + c) * Bp + b];
  return OK;
}

// per chain summaries over the kept draws (R/BGGE.R:408-433); all outputs chain-major:
// yHat[B x n], varE[B], varE_sd[B], u[nk x B x n], u_sd[nk x B x n], varu[nk x B], varu_sd[nk x B], y[B x n]
int bgge_batch_summary(bgge_batch* h, double* yHat, double* varE, double* varE_sd, double* u, double* u_sd, double* varu,
                       double* varu_sd, double* y) {
  BGGE_HANDLE(h);
  Batch& g = h->g;
  BGGE_REQUIRE(g.initialised, "batch_summary: handle not initialised");
  BGGE_CUDA(cudaStreamSynchronize(g.stream));
  const size_t nc = (size_t)g.ncum, B = (size_t)g.B, Bp = (size_t)g.Bp, n = (size_t)g.n;
  std::vector<double> cmu(B * nc), cve(B * nc), cvu((size_t)g.nk * B * nc);
  BGGE_TRY(bgge_batch_chain(h, cmu.data(), cve.data(), cvu.data()));
  std::vector<size_t> kept;
  for (size_t slot = 0; slot < nc; ++slot) {
    const int64_t it = (int64_t)(slot + 1) * g.thin;
    if (it > g.burn && it <= g.done) kept.push_back(slot);
  }
  const size_t k = kept.size();
  std::vector<double> mu_est(B, NAN);
  for (size_t b = 0; b < B; ++b) {
    std::vector<double> a, e;
    for (size_t sl : kept) {
      a.push_back(cmu[b * nc + sl]);
      e.push_back(cve[b * nc + sl]);
    }
    mu_est[b] = mean_sd(a, nullptr);
    double sd = NAN;
    const double m = mean_sd(e, &sd);
    if (varE) varE[b] = m;
    if (varE_sd) varE_sd[b] = sd;
    for (int j = 0; j < g.nk; ++j) {
      std::vector<double> v;
      for (size_t sl : kept) v.push_back(cvu[((size_t)j * B + b) * nc + sl]);
      double sdj = NAN;
      const double mj = mean_sd(v, &sdj);
      if (varu) varu[(size_t)j * B + b] = mj;
      if (varu_sd) varu_sd[(size_t)j * B + b] = sdj;
    }
  }
  std::vector<double> t(n * Bp);
  if (yHat) {
    for (size_t b = 0; b < B; ++b)
      for (size_t i = 0; i < n; ++i) yHat[b * n + i] = mu_est[b];
    if (g.q > 0) {                                                   // + XF colMeans(B_mcmc[draw, ])   (R/BGGE.R:414-415)
      const size_t q = (size_t)g.q;
      std::vector<double> cb(B * nc * q);
      BGGE_TRY(bgge_batch_chain_bet(h, cb.data()));
      std::vector<double> bmean(q);
      for (size_t b = 0; b < B; ++b) {
        for (size_t c = 0; c < q; ++c) {
          long double a = 0.0L;
          for (size_t sl : kept) a += cb[(b * nc + sl) * q + c];
          bmean[c] = k ? (double)(a / (long double)k) : NAN;
        }
        for (size_t i = 0; i < n; ++i) {
          double xb = 0.0;
          for (size_t c = 0; c < q; ++c) xb += g.xf_host[i + c * n] * bmean[c];
          yHat[b * n + i] += xb;
        }
      }
    }
  }
  for (int j = 0; j < g.nk; ++j) {
    if (yHat || u) {
      BGGE_CUDA(cudaMemcpy(t.data(), g.Um.p + (size_t)j * n * Bp, n * Bp * sizeof(double), cudaMemcpyDeviceToHost));
      for (size_t b = 0; b < B; ++b)
        for (size_t i = 0; i < n; ++i) {
          const double v = k ? t[i * Bp + b] : NAN;
          if (u) u[((size_t)j * B + b) * n + i] = v;
          if (yHat) yHat[b * n + i] += v;
        }
    }
    if (u_sd) {
      BGGE_CUDA(cudaMemcpy(t.data(), g.Um2.p + (size_t)j * n * Bp, n * Bp * sizeof(double), cudaMemcpyDeviceToHost));
      for (size_t b = 0; b < B; ++b)
        for (size_t i = 0; i < n; ++i) u_sd[((size_t)j * B + b) * n + i] = k > 1 ? std::sqrt(t[i * Bp + b] / (double)(k - 1)) : NAN;
    }
  }
  if (y) {
    BGGE_CUDA(cudaMemcpy(t.data(), g.Y.p, n * Bp * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < B; ++b)
      for (size_t i = 0; i < n; ++i) y[b * n + i] = t[i * Bp + b];
  }
  return OK;
}

// ====================================================================== one-shot BGGE()
int bgge_fit(const double* y, const uint8_t* na_mask, int64_t n, const double* const* K, const int* types, int nk,
             const double* XF, int q, const int64_t* ne, int n_ne, int64_t ite, int64_t burn, int64_t thin, double tol,
             double R2, int rng_mode, uint64_t seed, const double* normals, int64_t n_normals, const double* chisq,
             int64_t n_chisq, double* yHat, double* varE, double* varE_sd, double* u, double* u_sd, double* varu,
             double* varu_sd, double* y_out, double* chain_mu, double* chain_varE, double* chain_varU) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(y && K && types && nk >= 1, "fit: NULL pointer");
  bool any_bd = false;
  for (int j = 0; j < nk; ++j) {
    BGGE_REQUIRE(types[j] == BGGE_TYPE_D || types[j] == BGGE_TYPE_BD, "Matrix should be of types BD or D");   // R/BGGE.R:220
    any_bd |= types[j] == BGGE_TYPE_BD;
  }
  BGGE_REQUIRE(!(any_bd && n_ne <= 1), "Type BD should be used only for block diagonal matrices");            // R/BGGE.R:223
  bgge_gibbs* h = nullptr;
  BGGE_TRY(bgge_gibbs_create(n, nk, q, nullptr, &h));
  std::unique_ptr<bgge_gibbs, void (*)(bgge_gibbs*)> guard(h, bgge_gibbs_destroy);
  BGGE_TRY(bgge_gibbs_set_y(h, y, na_mask));
  if (q > 0) BGGE_TRY(bgge_gibbs_set_xf(h, XF));
  for (int j = 0; j < nk; ++j)
    BGGE_TRY(bgge_gibbs_add_kernel_matrix(h, j, types[j], K[j], n, 0, ne, n_ne, tol, 1));
  if (rng_mode == BGGE_RNG_TAPE) BGGE_TRY(bgge_gibbs_set_tape(h, normals, n_normals, chisq, n_chisq));
  BGGE_TRY(bgge_gibbs_init(h, ite, burn, thin, R2, rng_mode, seed, 0));
  BGGE_TRY(bgge_gibbs_run(h, ite));
  BGGE_TRY(bgge_gibbs_summary(h, yHat, varE, varE_sd, u, u_sd, varu, varu_sd, y_out));
  BGGE_TRY(bgge_gibbs_chain(h, chain_mu, chain_varE, chain_varU, nullptr));
  return OK;
}

// BGGE() with the K list as getK() builds it: factored descriptions go through the L x L setDEC route and stay
// factorised in the sweep, dense entries through the n x n route.  See include/bgge_b200.h.
int bgge_fit_kernels(const double* y, const uint8_t* na_mask, int64_t n, const bgge_kernel_desc* kernels, int nk,
                     const int32_t* gid, const int32_t* env, const double* XF, int q, const int64_t* ne, int n_ne,
                     int64_t ite, int64_t burn, int64_t thin, double tol, double R2, int rng_mode, uint64_t seed,
                     const double* normals, int64_t n_normals, const double* chisq, int64_t n_chisq, int64_t chunk,
                     bgge_progress_fn progress, void* user, double* yHat, double* varE, double* varE_sd, double* u,
                     double* u_sd, double* varu, double* varu_sd, double* y_out, double* chain_mu, double* chain_varE,
                     double* chain_varU, double* chain_bet, int* factored_kernels) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(y && kernels && nk >= 1 && n > 1, "fit_kernels: NULL pointer or empty model");
  BGGE_REQUIRE(ite >= 1 && thin >= 1 && burn >= 0, "fit_kernels: ite=%lld burn=%lld thin=%lld", (long long)ite, (long long)burn,
               (long long)thin);
  bool any_bd = false, any_fact = false;
  for (int j = 0; j < nk; ++j) {
    const bgge_kernel_desc& kd = kernels[j];
    BGGE_REQUIRE(kd.type == BGGE_TYPE_D || kd.type == BGGE_TYPE_BD, "Matrix should be of types BD or D");       // R/BGGE.R:220
    BGGE_REQUIRE(kd.mode == BGGE_KERNEL_DENSE || (kd.mode >= BGGE_EXPAND_G && kd.mode <= BGGE_EXPAND_GI),
                 "fit_kernels: kernel %d has unknown mode %d", j, kd.mode);
    BGGE_REQUIRE(kd.mode != BGGE_KERNEL_DENSE || kd.dense != nullptr, "fit_kernels: kernel %d: dense matrix missing", j);
    any_bd |= kd.type == BGGE_TYPE_BD;
    any_fact |= kd.mode != BGGE_KERNEL_DENSE;
  }
  BGGE_REQUIRE(!any_bd || ne != nullptr, "For type BD, number of subjects in each sub matrices should be provided");  // R/BGGE.R:129
  BGGE_REQUIRE(!(any_bd && n_ne <= 1), "Type BD should be used only for block diagonal matrices");              // R/BGGE.R:223
  BGGE_REQUIRE(!any_fact || gid != nullptr, "fit_kernels: genotype codes missing");
  // BGGE_TIMING=1: wall-clock of the stages of this call on stderr (device drained at every mark)
  const bool timing = getenv("BGGE_TIMING") != nullptr;
  auto tlast = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const auto tn = std::chrono::steady_clock::now();
    fprintf(stderr, "[bgge timing] fit_kernels %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(tn - tlast).count());
    tlast = tn;
  };
  bgge_gibbs* h = nullptr;
  BGGE_TRY(bgge_gibbs_create(n, nk, q, nullptr, &h));
  std::unique_ptr<bgge_gibbs, void (*)(bgge_gibbs*)> guard(h, bgge_gibbs_destroy);
  BGGE_TRY(bgge_gibbs_set_y(h, y, na_mask));
  if (q > 0) BGGE_TRY(bgge_gibbs_set_xf(h, XF));
  mark("create + y");
  {
    // every distinct base kernel is uploaded once and shared by the kernels that expand it (G and GE of MDs, G and
    // the E environment kernels of MDe)
    std::vector<std::pair<const double*, std::unique_ptr<DevBuf<double>>>> up;
    int nfact = 0;
    for (int j = 0; j < nk; ++j) {
      const bgge_kernel_desc& kd = kernels[j];
      if (kd.mode == BGGE_KERNEL_DENSE) {
        BGGE_TRY(bgge_gibbs_add_kernel_matrix(h, j, kd.type, kd.dense, n, 0, ne, n_ne, tol, 1));
        continue;
      }
      const double* dK0 = nullptr;
      if (kd.mode != BGGE_EXPAND_GI) {
        BGGE_REQUIRE(kd.K0 != nullptr && kd.L > 0, "fit_kernels: kernel %d: base kernel missing", j);
        for (auto& e : up)
          if (e.first == kd.K0 && (int64_t)e.second->n == kd.L * kd.L) dK0 = e.second->p;
        if (!dK0) {
          std::unique_ptr<DevBuf<double>> b(new DevBuf<double>());
          BGGE_CUDA(b->alloc((size_t)kd.L * kd.L));
          BGGE_TRY(host_copy(b->p, kd.K0, (size_t)kd.L * kd.L * sizeof(double), cudaMemcpyHostToDevice));
          dK0 = b->p;
          up.emplace_back(kd.K0, std::move(b));
          mark("base kernel upload");
        }
      }
      BGGE_TRY(bgge_gibbs_add_expanded_kernel(h, j, kd.type, dK0, kd.L, kd.L, 1, gid, env, n, kd.mode, kd.env_k, ne, n_ne, tol,
                                              1));
      ++nfact;
      mark("setDEC of one kernel");
    }
    if (factored_kernels) *factored_kernels = nfact;
    BGGE_CUDA(cudaStreamSynchronize(h->g.stream));   // the base kernels die here; the eigen blocks are what the chain keeps
  }
  if (rng_mode == BGGE_RNG_TAPE) BGGE_TRY(bgge_gibbs_set_tape(h, normals, n_normals, chisq, n_chisq));
  mark("release base kernels");
  BGGE_TRY(bgge_gibbs_init(h, ite, burn, thin, R2, rng_mode, seed, 0));
  mark("init (plans, state)");
  const int64_t piece = (chunk > 0 && progress) ? chunk : ite;
  for (int64_t done = 0; done < ite;) {                                                                          // R/BGGE.R:274
    const int64_t k = std::min(piece, ite - done);
    const auto t0 = std::chrono::steady_clock::now();
    BGGE_TRY(bgge_gibbs_run(h, k));
    done += k;
    if (progress) {
      BGGE_CUDA(cudaStreamSynchronize(h->g.stream));
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (progress(done, dt / (double)k, user) != 0) {                                                         // R/BGGE.R:398-401
        set_error("interrupted after %lld of %lld iterations", (long long)done, (long long)ite);
        return ERR_STATE;
      }
    }
  }
  mark("sweeps");
  BGGE_TRY(bgge_gibbs_summary(h, yHat, varE, varE_sd, u, u_sd, varu, varu_sd, y_out));
  BGGE_TRY(bgge_gibbs_chain(h, chain_mu, chain_varE, chain_varU, chain_bet));
  mark("summaries + chains");
  if (timing) {
    guard.reset();
    mark("destroy");
  }
  return OK;
}

// ====================================================================== test hooks
int bgge_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  BGGE_TRY(ensure_device());
  DevBuf<uint32_t> d;
  BGGE_CUDA(d.alloc(10));
  BGGE_CUDA(cudaMemcpy(d.p, ctr, 4 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  BGGE_CUDA(cudaMemcpy(d.p + 4, key, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  philox_test_kernel<<<1, 1>>>(d.p, d.p + 4, d.p + 6);
  BGGE_CUDA(cudaGetLastError());
  BGGE_CUDA(cudaMemcpy(out, d.p + 6, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return OK;
}

int bgge_test_draws(uint64_t seed, uint32_t chain, int64_t n, double* normals, double df, double* chisq) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(n > 0, "test_draws: n must be positive");
  DrawCfg dc{};
  dc.mode = 0;
  dc.seed = seed;
  dc.chain = chain;
  DevBuf<double> z, c;
  if (normals) BGGE_CUDA(z.alloc((size_t)n));
  if (chisq) BGGE_CUDA(c.alloc((size_t)n));
  draws_test_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dc, n, z.p, df, c.p);
  BGGE_CUDA(cudaGetLastError());
  if (normals) BGGE_CUDA(cudaMemcpy(normals, z.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  if (chisq) BGGE_CUDA(cudaMemcpy(chisq, c.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return OK;
}

}  // extern "C"

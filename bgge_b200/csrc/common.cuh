// Shared device/host helpers for libbgge_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>

namespace bgge {

// ---------------------------------------------------------------- error plumbing
enum Status : int {
  OK = 0,
  ERR_INVALID = 1,   // bad argument
  ERR_CUDA = 2,      // CUDA runtime / launch failure
  ERR_SOLVER = 3,    // cuSOLVER failure / non-convergence
  ERR_STATE = 4,     // call out of order (e.g. run before init)
  ERR_NODEVICE = 5,  // no usable sm_100 device
  ERR_TAPE = 6,      // injected tape exhausted
  ERR_UNSUPPORTED = 7,
  ERR_RETRY_STREAMING = 100   // internal: the resident kernel could not be launched, use the streaming kernels
};

void set_error(const char* fmt, ...);
const char* last_error();

#define BGGE_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::bgge::set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__,       \
                        (int)_e, cudaGetErrorString(_e), #call);                    \
      return ::bgge::ERR_CUDA;                                                      \
    }                                                                               \
  } while (0)

#define BGGE_REQUIRE(cond, ...)                                                     \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::bgge::set_error(__VA_ARGS__);                                               \
      return ::bgge::ERR_INVALID;                                                   \
    }                                                                               \
  } while (0)

#define BGGE_TRY(expr)                                                              \
  do {                                                                              \
    int _s = (expr);                                                                \
    if (_s != 0) return _s;                                                         \
  } while (0)

int ensure_device();   // checks that the current device is sm_100; caches SM count
int sm_count();
bool pdl_enabled();   // BGGE_NO_PDL=1 switches programmatic dependent launch off

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Device memory comes from a small per-device caching pool (abi.cu): repeated BGGE()/getK() calls
// reuse their scratch blocks instead of paying cudaMalloc/cudaFree (cudaFree alone cost up to 180 ms
// per handle teardown in a process that also holds tens of GB of other allocations).
cudaError_t pool_alloc(void** p, size_t bytes);
void pool_free(void* p);
void pool_release_all();

// RAII device buffer (host side).  Callers synchronise the stream that used it before it dies.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) {
    o.p = nullptr;
    o.n = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      n = o.n;
      o.p = nullptr;
      o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) pool_free(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return pool_alloc((void**)&p, count * sizeof(T));
  }
};

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// `make CHECKS=1` builds libbgge_b200_checks.so with device-side assertions in the hand-synchronised kernels (ring
// indices, shared-memory carve, row ranges): the stand-in for compute-sanitizer, which is closed on this pool
// (profiles/r02_sanitizer.md).  A failed assertion prints its location and traps (the launch fails loudly).
#ifdef BGGE_CHECKS
#define BGGE_DEVASSERT(cond)                                                                                        \
  do {                                                                                                              \
    if (!(cond)) {                                                                                                  \
      printf("[bgge] device assertion failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__,          \
             (int)blockIdx.x, (int)threadIdx.x);                                                                    \
      __trap();                                                                                                     \
    }                                                                                                               \
  } while (0)
#else
#define BGGE_DEVASSERT(cond) ((void)0)
#endif
__device__ __forceinline__ uint32_t dynamic_smem_bytes() {
  uint32_t v;
  asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(v));
  return v;
}

// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream still runs; pdl_wait() blocks until the predecessor has completed and
// its writes are visible (no-op otherwise).  Everything before it may only touch data no earlier launch of the
// same iteration writes (tables, the eigen basis).  pdl_trigger() lets the successor's CTAs be scheduled as soon
// as every CTA of this grid has started.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 16-byte load that does not allocate in L1 (U is read once per pass)
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double ldg_stream1(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier / bulk-copy (TMA unit, SASS UBLKCP) wrappers
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// D(8x8) += A(8x4,row) * B(4x8,col), FP64 tensor core (SASS DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

#endif  // __CUDACC__

}  // namespace bgge

// Multi-GPU inside the library (SURVEY.md 8b "Threading", 8e): NCCL communicators owned by libbgge_b200, the
// row-panel getK contraction with an in-place exchange, a host-buffer getK that fans out over the GPUs of one
// process (one host thread per device), and the all-reduce of the column-sharded chain on the handle's own stream.
// Callers need neither torch nor any other collective library.
//
// NCCL is bound at run time (dlopen of libnccl.so.2; an already loaded copy -- e.g. the one a Python process
// imported with torch -- is reused), so the single-GPU library has no NCCL dependency.
//
// Row-panel layout.  S = X X' is symmetric: rank r computes the tiles (ti, tj <= ti) of its tile rows but stores them
// TRANSPOSED, S[tj-rows, ti-columns].  In column-major memory the columns of a tile-row panel are one contiguous
// range, so every rank owns two contiguous column panels of S (chunks r and 2 world - 1 - r of 2 world equal chunks:
// equal tile counts) and the exchange is 2 world in-place broadcasts inside one NCCL group -- no staging buffer,
// no packing kernel, no zero fill.  A transposing pass (upper -> lower) completes the matrix on every rank.
#include "../../include/bgge_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "multi.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace bgge {
namespace {

struct Nccl {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  std::string why;
};

Nccl g_nccl;
std::once_flag g_nccl_once;

void nccl_bind() {
  Nccl& n = g_nccl;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.so = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // a copy the process already uses
    if (n.so) break;
  }
  if (!n.so)
    if (const char* ev = getenv("BGGE_NCCL_LIB")) n.so = dlopen(ev, RTLD_NOW | RTLD_GLOBAL);
  if (!n.so)
    for (const char* nm : names) {
      n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.so) break;
    }
  if (!n.so) {
    const char* de = dlerror();   // (a second dlerror() call would return NULL)
    n.why = std::string("libnccl.so.2 could not be loaded: ") + (de ? de : "?");
    return;
  }
  bool ok = true;
  auto sym = [&](const char* s) -> void* {
    void* p = dlsym(n.so, s);
    if (!p) {
      ok = false;
      n.why = std::string("NCCL symbol missing: ") + s;
    }
    return p;
  };
  n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
  n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
  n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
  n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
  n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
  n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
  n.Broadcast = reinterpret_cast<decltype(n.Broadcast)>(sym("ncclBroadcast"));
  n.AllGather = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
  n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
  n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
  if (!ok) {
    dlclose(n.so);
    n.so = nullptr;
  }
}

int nccl_ready() {
  std::call_once(g_nccl_once, nccl_bind);
  if (!g_nccl.so) {
    set_error("multi-GPU entry points need NCCL: %s", g_nccl.why.c_str());
    return ERR_UNSUPPORTED;
  }
  return OK;
}

#define BGGE_NCCL(call)                                                                             \
  do {                                                                                              \
    ncclResult_t _r = (call);                                                                       \
    if (_r != ncclSuccess) {                                                                        \
      ::bgge::set_error("%s:%d NCCL error %d (%s) in %s", __FILE__, __LINE__, (int)_r,              \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", #call);            \
      return ::bgge::ERR_CUDA;                                                                      \
    }                                                                                               \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace

int comm_allreduce_sum(bgge_comm* c, double* buf, int64_t count, cudaStream_t st) {
  BGGE_TRY(nccl_ready());
  BGGE_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)c->comm, st));
  return OK;
}

void panel_chunks(int64_t L, int world, int rank, int64_t t0[2], int64_t t1[2]) {
  const int64_t T = (L + 127) / 128;
  const int64_t ch = (T + 2 * world - 1) / (2 * world);
  const int64_t c[2] = {rank, 2 * world - 1 - rank};
  for (int k = 0; k < 2; ++k) {
    t0[k] = std::min(T, c[k] * ch);
    t1[k] = std::min(T, (c[k] + 1) * ch);
  }
}

// dS (L x L, leading dimension L) <- X X' / divisor on every rank of the communicator
int gram_panels_exchange(bgge_comm* c, const double* dX, int64_t ldx, int64_t L, int64_t p_pad, double* dS, double divisor,
                         cudaStream_t st, float* ms_compute, float* ms_exchange) {
  const int world = c->world;
  const int64_t T = (L + 127) / 128;
  if (world == 1) {
    BGGE_TRY(gram_launch(dX, ldx, L, p_pad, 0, T, dS, L, 0, divisor, 1, st));
    return OK;
  }
  BGGE_TRY(nccl_ready());
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  const bool timed = ms_compute || ms_exchange;
  if (timed)
    for (auto& e : ev) BGGE_CUDA(cudaEventCreate(&e));
  if (timed) BGGE_CUDA(cudaEventRecord(ev[0], st));
  int64_t t0[2], t1[2];
  panel_chunks(L, world, c->rank, t0, t1);
  // both row panels of this rank in one launch, stored transposed (column panels of S)
  BGGE_TRY(gram_launch(dX, ldx, L, p_pad, t0[0], t1[0], dS, L, 0, divisor, 2, st, 1, t0[1], t1[1]));
  if (timed) BGGE_CUDA(cudaEventRecord(ev[1], st));
  const int64_t ch = (T + 2 * world - 1) / (2 * world);
  BGGE_NCCL(g_nccl.GroupStart());
  for (int ck = 0; ck < 2 * world; ++ck) {
    const int64_t c0 = std::min(L, ck * ch * 128), c1 = std::min(L, (ck + 1) * ch * 128);
    if (c1 <= c0) continue;
    const int root = ck < world ? ck : 2 * world - 1 - ck;
    double* p = dS + c0 * L;
    BGGE_NCCL(g_nccl.Broadcast(p, p, (size_t)((c1 - c0) * L), ncclDouble, root, (ncclComm_t)c->comm, st));
  }
  BGGE_NCCL(g_nccl.GroupEnd());
  BGGE_TRY(symmetrize_launch(dS, L, L, st, 1));   // upper (what the panels hold) -> lower
  if (timed) {
    BGGE_CUDA(cudaEventRecord(ev[2], st));
    BGGE_CUDA(cudaEventSynchronize(ev[2]));
    float a = 0.f, b = 0.f;
    BGGE_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
    BGGE_CUDA(cudaEventElapsedTime(&b, ev[1], ev[2]));
    if (ms_compute) *ms_compute = a;
    if (ms_exchange) *ms_exchange = b;
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return OK;
}

}  // namespace bgge

using namespace bgge;

extern "C" {

int bgge_comm_unique_id(void* id) {
  BGGE_REQUIRE(id != nullptr, "comm_unique_id: NULL output");
  BGGE_TRY(nccl_ready());
  static_assert(sizeof(ncclUniqueId) == BGGE_COMM_ID_BYTES, "NCCL unique id size");
  ncclUniqueId u;
  BGGE_NCCL(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return OK;
}

int bgge_comm_init_rank(const void* id, int world, int rank, bgge_comm** out) {
  BGGE_TRY(ensure_device());
  BGGE_REQUIRE(id && out && world >= 1 && rank >= 0 && rank < world, "comm_init_rank: bad arguments (rank %d of %d)", rank, world);
  BGGE_TRY(nccl_ready());
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t cm;
  BGGE_NCCL(g_nccl.CommInitRank(&cm, world, u, rank));
  bgge_comm* c = new bgge_comm();
  c->comm = cm;
  c->rank = rank;
  c->world = world;
  BGGE_CUDA(cudaGetDevice(&c->device));
  *out = c;
  return OK;
}

int bgge_comm_init_all(int ndev, const int* devices, bgge_comm** out) {
  BGGE_REQUIRE(out && ndev >= 1 && ndev <= 64, "comm_init_all: bad arguments (ndev %d)", ndev);
  int have = 0;
  BGGE_CUDA(cudaGetDeviceCount(&have));
  std::vector<int> devs((size_t)ndev);
  for (int d = 0; d < ndev; ++d) {
    devs[(size_t)d] = devices ? devices[d] : d;
    BGGE_REQUIRE(devs[(size_t)d] >= 0 && devs[(size_t)d] < have, "comm_init_all: device %d of %d", devs[(size_t)d], have);
  }
  BGGE_TRY(nccl_ready());
  std::vector<ncclComm_t> cms((size_t)ndev);
  BGGE_NCCL(g_nccl.CommInitAll(cms.data(), ndev, devs.data()));
  for (int d = 0; d < ndev; ++d) {
    bgge_comm* c = new bgge_comm();
    c->comm = cms[(size_t)d];
    c->rank = d;
    c->world = ndev;
    c->device = devs[(size_t)d];
    out[d] = c;
  }
  return OK;
}

int bgge_comm_info(const bgge_comm* c, int* rank, int* world, int* device) {
  BGGE_REQUIRE(c != nullptr, "comm_info: NULL communicator");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (device) *device = c->device;
  return OK;
}

void bgge_comm_destroy(bgge_comm* c) {
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->comm);
  delete c;
}

int bgge_dev_allreduce(bgge_comm* c, double* d_buf, int64_t count, void* stream) {
  BGGE_REQUIRE(c && d_buf && count > 0, "dev_allreduce: bad arguments");
  BGGE_CUDA(cudaSetDevice(c->device));
  return comm_allreduce_sum(c, d_buf, count, as_stream(stream));
}

int bgge_dev_gram_panels(bgge_comm* c, const double* dX, int64_t ldx, int64_t L, int64_t p_pad, double* dS, double divisor,
                         void* stream, float* ms_compute, float* ms_exchange) {
  BGGE_REQUIRE(c && dX && dS && L > 0 && divisor != 0.0, "dev_gram_panels: bad arguments");
  BGGE_CUDA(cudaSetDevice(c->device));
  BGGE_TRY(ensure_device());
  return gram_panels_exchange(c, dX, ldx, L, p_pad, dS, divisor, as_stream(stream), ms_compute, ms_exchange);
}

// getK base kernels (R/getK.R:131-149) over the GPUs of this process.  One host thread per device: each uploads its
// slice of the marker columns (1/ndev of X over its own PCIe link), an in-place all-gather replicates X, the row
// panels of X X' are built and exchanged, device 0 runs the L x L epilogue (GK: distance, quantile, exp) and returns
// the kernels.  Results equal bgge_getk_base to rounding of the K-split sums (<= 1e-12 relative).
int bgge_getk_base_multi(bgge_comm* const* comms, int ndev, const double* X, int64_t L, int64_t p, int kind,
                         const double* bandwidth, int nb, double quantil, double* out_K0, double* out_q) {
  BGGE_REQUIRE(comms && ndev >= 1 && X && out_K0 && L > 0 && p > 0, "getk_base_multi: NULL pointer or empty matrix");
  BGGE_REQUIRE(kind == BGGE_KERNEL_GB || kind == BGGE_KERNEL_GK,
               "kernel selected is not available. Please choose one method available or make available other "
               "kernel through argument K");
  BGGE_REQUIRE(kind == BGGE_KERNEL_GB || (nb >= 1 && bandwidth != nullptr), "GK kernel needs at least one bandwidth");
  for (int d = 0; d < ndev; ++d)
    BGGE_REQUIRE(comms[d] && comms[d]->world == ndev && comms[d]->rank == d, "getk_base_multi: communicator %d is not rank %d of %d", d, d, ndev);
  if (ndev > 1) BGGE_TRY(nccl_ready());
  const int64_t ldx = round_up(L, 128);
  const int64_t pc = round_up((round_up(p, 16) + ndev - 1) / ndev, 16);   // marker columns per device (padded)
  const int64_t p_pad = pc * ndev;
  std::vector<int> status((size_t)ndev, OK);
  std::vector<std::string> msg((size_t)ndev);
  // Host rendezvous in front of the first collective: a device that failed to allocate or upload must not leave the
  // others waiting inside NCCL, so everybody learns about a failure before anybody enters the all-gather.
  std::mutex rv_mu;
  std::condition_variable rv_cv;
  int rv_arrived = 0;
  bool rv_failed = false;
  auto rendezvous = [&](bool ok) -> bool {   // returns true when every device got here without an error
    std::unique_lock<std::mutex> lk(rv_mu);
    if (!ok) rv_failed = true;
    if (++rv_arrived == ndev) rv_cv.notify_all();
    else rv_cv.wait(lk, [&] { return rv_arrived == ndev; });
    return !rv_failed;
  };
  auto worker = [&](int d) -> int {
    bgge_comm* c = comms[d];
    cudaStream_t st = nullptr;
    auto open_device = [&]() -> int {
      BGGE_CUDA(cudaSetDevice(c->device));
      BGGE_TRY(ensure_device());
      BGGE_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      return OK;
    };
    const int os = open_device();
    if (os != OK) {
      rendezvous(false);
      return os;
    }
    struct StreamGuard {
      cudaStream_t s;
      ~StreamGuard() { cudaStreamDestroy(s); }
    } sg{st};
    DevBuf<double> dX, S;
    double* mine = nullptr;
    // my marker columns [d pc, (d + 1) pc): zero the slice (row / column padding), then a strided copy of what exists
    auto prepare = [&]() -> int {
      BGGE_CUDA(dX.alloc((size_t)ldx * (size_t)p_pad));
      BGGE_CUDA(S.alloc((size_t)L * L));
      mine = dX.p + (size_t)d * pc * ldx;
      BGGE_CUDA(cudaMemsetAsync(mine, 0, (size_t)pc * ldx * sizeof(double), st));
      const int64_t cfirst = (int64_t)d * pc, ccount = std::max<int64_t>(0, std::min<int64_t>(p, cfirst + pc) - cfirst);
      if (ccount > 0)
        BGGE_CUDA(cudaMemcpy2DAsync(mine, (size_t)ldx * sizeof(double), X + (size_t)cfirst * L, (size_t)L * sizeof(double),
                                    (size_t)L * sizeof(double), (size_t)ccount, cudaMemcpyHostToDevice, st));
      return OK;
    };
    const int ps = prepare();
    if (!rendezvous(ps == OK)) {
      if (ps == OK) set_error("another device failed before the exchange");
      return ps == OK ? ERR_STATE : ps;
    }
    if (ndev > 1)
      BGGE_NCCL(g_nccl.AllGather(mine, dX.p, (size_t)pc * ldx, ncclDouble, (ncclComm_t)c->comm, st));
    const double divisor = kind == BGGE_KERNEL_GB ? (double)p : 1.0;
    BGGE_TRY(gram_panels_exchange(c, dX.p, ldx, L, p_pad, S.p, divisor, st, nullptr, nullptr));
    if (d == 0) {
      if (kind == BGGE_KERNEL_GB) {
        BGGE_CUDA(cudaMemcpyAsync(out_K0, S.p, (size_t)L * L * sizeof(double), cudaMemcpyDeviceToHost, st));
      } else {
        DevBuf<double> diag, q, K0;
        DevBuf<unsigned char> scratch;
        BGGE_CUDA(diag.alloc((size_t)L));
        BGGE_CUDA(q.alloc(1));
        BGGE_CUDA(K0.alloc((size_t)L * L));
        BGGE_CUDA(scratch.alloc(quantile_scratch_bytes()));
        BGGE_TRY(gk_distance_launch(S.p, L, L, diag.p, st));
        BGGE_TRY(quantile7_launch(S.p, L, L, quantil, scratch.p, q.p, st));
        for (int b = 0; b < nb; ++b) {
          BGGE_TRY(gk_exp_launch(S.p, L, L, bandwidth[b], q.p, K0.p, L, st));
          BGGE_CUDA(cudaMemcpyAsync(out_K0 + (size_t)b * L * L, K0.p, (size_t)L * L * sizeof(double), cudaMemcpyDeviceToHost, st));
          BGGE_CUDA(cudaStreamSynchronize(st));
        }
        if (out_q) BGGE_CUDA(cudaMemcpyAsync(out_q, q.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        BGGE_CUDA(cudaStreamSynchronize(st));
      }
    }
    BGGE_CUDA(cudaStreamSynchronize(st));
    return OK;
  };
  int cur = 0;
  cudaGetDevice(&cur);
  std::vector<std::thread> th;
  for (int d = 1; d < ndev; ++d)
    th.emplace_back([&, d]() {
      status[(size_t)d] = worker(d);
      if (status[(size_t)d] != OK) msg[(size_t)d] = last_error();
    });
  status[0] = worker(0);
  if (status[0] != OK) msg[0] = last_error();
  for (auto& t : th) t.join();
  cudaSetDevice(cur);
  for (int d = 0; d < ndev; ++d)
    if (status[(size_t)d] != OK) {
      set_error("getk_base_multi: device %d: %s", comms[d]->device, msg[(size_t)d].c_str());
      return status[(size_t)d];
    }
  return OK;
}

}  // extern "C"

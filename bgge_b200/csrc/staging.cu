// Host <-> device copies of large PAGEABLE buffers (what an R or ctypes caller hands over: the marker matrix, base
// kernels, the n x n kernels getK() returns).  cudaMemcpy stages pageable memory through one driver thread
// (measured r02: 9.6 GB/s up for the 8 GB marker matrix of C5, 3.9 GB/s down for a 12.8 GB kernel into freshly
// allocated pages); here T host threads each run their own double-buffered pipeline pageable <-> pinned <-> device
// over a contiguous range of rows, so the host-side memcpy (and the first-touch page faults of an output buffer) run
// in parallel and overlap the DMA.  Same semantics as a synchronous cudaMemcpy2D.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace bgge {
namespace {

constexpr size_t STAGE_BYTES = 16u << 20;     // one pinned buffer
constexpr size_t STAGED_MIN_BYTES = 64u << 20;   // below this a plain copy is as fast
constexpr int MAX_LANES = 8;

struct Lane {
  unsigned char* pin[2] = {nullptr, nullptr};
  cudaStream_t st = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool ready = false;
};

std::mutex g_copy_mu;                 // one staged copy at a time per process (the lanes are shared)
Lane g_lanes[64][MAX_LANES];          // [device][lane]

int lane_init(Lane& ln) {
  if (ln.ready) return OK;
  for (int b = 0; b < 2; ++b) {
    BGGE_CUDA(cudaHostAlloc((void**)&ln.pin[b], STAGE_BYTES, cudaHostAllocDefault));
    BGGE_CUDA(cudaEventCreateWithFlags(&ln.ev[b], cudaEventDisableTiming));
  }
  BGGE_CUDA(cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking));
  ln.ready = true;
  return OK;
}

int copy_threads() {
  int t = (int)std::thread::hardware_concurrency() / 2;
  if (const char* ev = getenv("BGGE_COPY_THREADS")) t = atoi(ev);
  return std::max(1, std::min(MAX_LANES, t));
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// rows [r0, r1) of the copy through lane ln; to_device: host (src) -> device (dst), else device (src) -> host (dst)
int lane_copy(Lane& ln, int dev, unsigned char* dst, size_t dpitch, const unsigned char* src, size_t spitch, size_t width,
              size_t r0, size_t r1, bool to_device) {
  BGGE_CUDA(cudaSetDevice(dev));
  const size_t rc = std::max<size_t>(1, STAGE_BYTES / width);   // rows per chunk
  auto host_rows = [&](unsigned char* hp, size_t hpitch, unsigned char* pp, size_t row, size_t rows, bool into_pinned) {
    if (hpitch == width) {
      if (into_pinned) memcpy(pp, hp + row * hpitch, rows * width);
      else memcpy(hp + row * hpitch, pp, rows * width);
      return;
    }
    for (size_t r = 0; r < rows; ++r) {
      if (into_pinned) memcpy(pp + r * width, hp + (row + r) * hpitch, width);
      else memcpy(hp + (row + r) * hpitch, pp + r * width, width);
    }
  };
  size_t c = 0;
  if (to_device) {
    for (size_t row = r0; row < r1; row += rc, ++c) {
      const int b = (int)(c & 1);
      const size_t rows = std::min(rc, r1 - row);
      if (c >= 2) BGGE_CUDA(cudaEventSynchronize(ln.ev[b]));   // the DMA out of this buffer has finished
      host_rows(const_cast<unsigned char*>(src), spitch, ln.pin[b], row, rows, true);
      BGGE_CUDA(cudaMemcpy2DAsync(dst + row * dpitch, dpitch, ln.pin[b], width, width, rows, cudaMemcpyHostToDevice, ln.st));
      BGGE_CUDA(cudaEventRecord(ln.ev[b], ln.st));
    }
  } else {
    size_t prev_row = 0, prev_rows = 0;
    for (size_t row = r0; row < r1; row += rc, ++c) {
      const int b = (int)(c & 1);
      const size_t rows = std::min(rc, r1 - row);
      BGGE_CUDA(cudaMemcpy2DAsync(ln.pin[b], width, src + row * spitch, spitch, width, rows, cudaMemcpyDeviceToHost, ln.st));
      BGGE_CUDA(cudaEventRecord(ln.ev[b], ln.st));
      if (prev_rows) {   // while that DMA runs, hand the previous chunk to the caller's buffer
        BGGE_CUDA(cudaEventSynchronize(ln.ev[b ^ 1]));
        host_rows(dst, dpitch, ln.pin[b ^ 1], prev_row, prev_rows, false);
      }
      prev_row = row;
      prev_rows = rows;
    }
    if (prev_rows) {
      const int b = (int)((c - 1) & 1);
      BGGE_CUDA(cudaEventSynchronize(ln.ev[b]));
      host_rows(dst, dpitch, ln.pin[b], prev_row, prev_rows, false);
    }
  }
  BGGE_CUDA(cudaStreamSynchronize(ln.st));
  return OK;
}

}  // namespace

// Synchronous 2-D copy between a host buffer and device memory (kind: cudaMemcpyHostToDevice / DeviceToHost).
// Large pageable buffers go through the threaded staging pipeline, everything else through cudaMemcpy2D.
// sync_device = false: the caller guarantees that no work in flight touches the device range (the copy may then
// overlap kernels that are still running on other buffers).
int host_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind,
                 bool sync_device) {
  if (width == 0 || height == 0) return OK;
  const bool to_device = kind == cudaMemcpyHostToDevice;
  const void* host_ptr = to_device ? src : dst;
  const int T = copy_threads();
  size_t min_bytes = STAGED_MIN_BYTES;
  if (const char* ev = getenv("BGGE_STAGED_MIN_BYTES")) min_bytes = (size_t)atoll(ev);   // test hook
  if ((kind != cudaMemcpyHostToDevice && kind != cudaMemcpyDeviceToHost) || width * height < min_bytes ||
      width > STAGE_BYTES || T < 2 || getenv("BGGE_NO_STAGED_COPY") || is_pinned(host_ptr)) {
    BGGE_CUDA(cudaMemcpy2D(dst, dpitch, src, spitch, width, height, kind));
    return OK;
  }
  int dev = 0;
  BGGE_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    BGGE_CUDA(cudaMemcpy2D(dst, dpitch, src, spitch, width, height, kind));
    return OK;
  }
  std::lock_guard<std::mutex> lk(g_copy_mu);
  // the lanes use their own streams: everything the caller enqueued before (on any stream) must be complete first,
  // as it would be for a cudaMemcpy that follows kernels on the default stream
  if (sync_device) BGGE_CUDA(cudaDeviceSynchronize());
  const int lanes = (int)std::min<size_t>((size_t)T, height);
  for (int t = 0; t < lanes; ++t) BGGE_TRY(lane_init(g_lanes[dev][t]));
  std::vector<int> status((size_t)lanes, OK);
  std::vector<std::string> msg((size_t)lanes);
  auto run = [&](int t) {
    const size_t r0 = height * (size_t)t / (size_t)lanes, r1 = height * (size_t)(t + 1) / (size_t)lanes;
    status[(size_t)t] = lane_copy(g_lanes[dev][t], dev, static_cast<unsigned char*>(dst), dpitch,
                                  static_cast<const unsigned char*>(src), spitch, width, r0, r1, to_device);
    if (status[(size_t)t] != OK) msg[(size_t)t] = last_error();
  };
  std::vector<std::thread> th;
  for (int t = 1; t < lanes; ++t) th.emplace_back(run, t);
  run(0);
  for (auto& x : th) x.join();
  BGGE_CUDA(cudaSetDevice(dev));
  for (int t = 0; t < lanes; ++t)
    if (status[(size_t)t] != OK) {
      set_error("staged copy: %s", msg[(size_t)t].c_str());
      return status[(size_t)t];
    }
  return OK;
}

int host_copy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  // a flat buffer as rows of 1 MB (+ a tail) so that the lanes get contiguous ranges
  const size_t w = 1u << 20;
  const size_t rows = bytes / w, tail = bytes - rows * w;
  if (rows > 0) BGGE_TRY(host_copy_2d(dst, w, src, w, w, rows, kind, true));
  if (tail > 0)
    BGGE_CUDA(cudaMemcpy(static_cast<unsigned char*>(dst) + rows * w, static_cast<const unsigned char*>(src) + rows * w, tail, kind));
  return OK;
}

}  // namespace bgge

// Internal interface of multi.cu (NCCL communicators owned by the library).
#pragma once
#include "common.cuh"

struct bgge_comm {
  void* comm = nullptr;   // ncclComm_t
  int rank = 0, world = 1, device = 0;
};

namespace bgge {

// in-place sum over the ranks of the communicator, ordered on `st`
int comm_allreduce_sum(bgge_comm* c, double* buf, int64_t count, cudaStream_t st);
// tile-row ranges [t0[k], t1[k]) of the two row panels rank owns (chunks rank and 2 world - 1 - rank)
void panel_chunks(int64_t L, int world, int rank, int64_t t0[2], int64_t t1[2]);
int gram_panels_exchange(bgge_comm* c, const double* dX, int64_t ldx, int64_t L, int64_t p_pad, double* dS, double divisor,
                         cudaStream_t st, float* ms_compute, float* ms_exchange);

}  // namespace bgge

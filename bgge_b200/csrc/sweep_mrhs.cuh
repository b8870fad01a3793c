// Multi-right-hand-side single-pass sweep (sweep_mrhs.cu): planning / launch interface used by sweep_fused.cu.
#pragma once
#include "gibbs.cuh"

namespace bgge {

constexpr int MRHS_MAX_NS = 8;   // ring depth limit (groups)
constexpr int MRHS_TMEM_DEFAULT = 1;   // BGGE_MRHS_TMEM default (see sweep_mrhs.cu: 0 never, 1 to shrink the cluster, 2 always)

struct MrhsCfg {
  bool ok = false;
  int cs = 1;          // CTAs per cluster (row split)
  int kr = 0;          // row steps per compute thread the kernel is instantiated for
  int g = 2;           // eigen-columns per exchange group (= per ring slot)
  int tmem = 0;        // 1: residual slices in tensor memory
  int ns = 0;          // ring depth (groups)
  int lag = 2;         // groups between the dots and the axpy of a group
  int clusters = 0;    // co-resident clusters
  size_t smem = 0;
};

int64_t mrhs_min_rows();   // blocks shorter than this stay on sweep_fused_kernel (BGGE_MRHS_MIN_ROWS)
int mrhs_configure(int64_t max_rows, int nb, bool merged, MrhsCfg* cfg);
int mrhs_launch(Gibbs* g, HostKernel& hk);

}  // namespace bgge

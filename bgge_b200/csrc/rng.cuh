// Counter-based draws for the Gibbs sweep: Philox4x32-10 keyed by (seed, chain), counter =
// (index, iteration, stream).  Normals by Box-Muller on 53-bit uniforms; chi-square as
// 2*Gamma(df/2) by Marsaglia-Tsang.  In injected mode (DrawCfg::mode == 1) both come from
// host-supplied tapes in the reference's call order instead.
#pragma once
#include "gibbs.cuh"

namespace bgge {

__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int rnd = 0; rnd < 10; ++rnd) {
    const uint64_t p0 = (uint64_t)M0 * c[0];
    const uint64_t p1 = (uint64_t)M1 * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}

#ifdef __CUDACC__
__device__ inline void philox_draw(const DrawCfg& dc, long long it, uint32_t stream, unsigned long long idx,
                                   uint32_t out[4]) {
  out[0] = (uint32_t)idx;
  out[1] = (uint32_t)(idx >> 32);
  out[2] = (uint32_t)it;
  out[3] = stream;
  philox4x32_10(out, (uint32_t)dc.seed, (uint32_t)(dc.seed >> 32) ^ dc.chain);
}

__device__ inline double u53(uint32_t hi, uint32_t lo) {   // uniform in (0,1)
  const unsigned long long a = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)a + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ inline double philox_normal(const DrawCfg& dc, long long it, uint32_t stream, unsigned long long idx) {
  uint32_t w[4];
  philox_draw(dc, it, stream, idx, w);
  const double u1 = u53(w[0], w[1]);
  const double u2 = u53(w[2], w[3]);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__device__ inline double draw_normal(const DrawCfg& dc, long long it, uint32_t stream, unsigned long long idx,
                                     long long tape_pos) {
  if (dc.mode == 1) return dc.zt[tape_pos];
  return philox_normal(dc, it, stream, idx);
}

// chi-square with df degrees of freedom (df >= 2 here: df = nr + 3 or n + 3)
__device__ inline double draw_chisq(const DrawCfg& dc, long long it, uint32_t stream, double df, long long tape_pos) {
  if (dc.mode == 1) return dc.ct[tape_pos];
  const double a = 0.5 * df;
  const double d = a - 1.0 / 3.0;
  const double c = 1.0 / sqrt(9.0 * d);
  for (unsigned long long attempt = 0;; ++attempt) {
    const double x = philox_normal(dc, it, stream, 2 * attempt);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    uint32_t w[4];
    philox_draw(dc, it, stream, 2 * attempt + 1, w);
    const double uu = u53(w[0], w[1]);
    if (log(uu) < 0.5 * x * x + d - d * v + d * log(v)) return 2.0 * d * v;
  }
}
#endif

}  // namespace bgge

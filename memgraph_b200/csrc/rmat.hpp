// memgraph_b200/csrc/rmat.hpp -- counter-based RMAT edge generator shared by host and device.
//
// Workload synthesis for BASELINE.json's configs: RMAT(a, b, c, d) with the
// reference generator's defaults a=0.57 b=0.19 c=0.19 (mage/cpp/cugraph_module/algorithms/
// graph_generator.cu:143-145), no clip-and-flip, no dedup, no vertex permutation.  The reference
// delegates the arithmetic to cugraph::generate_rmat_edgelist (third party, not available), so
// the bit stream is defined HERE: one splitmix64 hash per (seed, edge, level pair); its high and
// low 32 bits choose the quadrants of two consecutive levels, most significant vertex bit first:
//   u <  A       -> (src bit 0, dst bit 0)
//   u <  A+B     -> (0, 1)
//   u <  A+B+C   -> (1, 0)
//   else         -> (1, 1)
// Pure integer arithmetic, so host and device agree bit for bit.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define MGB_HD __host__ __device__ __forceinline__
#else
#define MGB_HD inline
#endif

namespace mgb200 {

struct RmatThresholds {
  uint32_t t1, t2, t3;  // A, A+B, A+B+C scaled to 2^32
};

inline RmatThresholds rmat_thresholds(double a, double b, double c) {
  const double two32 = 4294967296.0;
  auto clampu = [&](double x) -> uint32_t {
    if (x <= 0.0) return 0u;
    if (x >= two32 - 1.0) return 0xffffffffu;
    return static_cast<uint32_t>(x);
  };
  return RmatThresholds{clampu(a * two32), clampu((a + b) * two32), clampu((a + b + c) * two32)};
}

MGB_HD uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

MGB_HD void rmat_quadrant(uint32_t u, const RmatThresholds &t, uint32_t &src, uint32_t &dst) {
  const uint32_t sbit = u >= t.t2 ? 1u : 0u;
  const uint32_t dbit = (u >= t.t1 && u < t.t2) || u >= t.t3 ? 1u : 0u;
  src = (src << 1) | sbit;
  dst = (dst << 1) | dbit;
}

MGB_HD void rmat_edge(uint32_t scale, uint64_t edge, uint64_t seed, const RmatThresholds &t, uint32_t &src_out,
                      uint32_t &dst_out) {
  const uint64_t edge_key = splitmix64(seed ^ (edge * 0xD1342543DE82EF95ull));
  uint32_t src = 0, dst = 0;
  for (uint32_t level = 0; level < scale; level += 2) {
    const uint64_t h = splitmix64(edge_key + (level >> 1) * 0xA24BAED4963EE407ull);
    rmat_quadrant(static_cast<uint32_t>(h >> 32), t, src, dst);
    if (level + 1 < scale) rmat_quadrant(static_cast<uint32_t>(h), t, src, dst);
  }
  src_out = src;
  dst_out = dst;
}

}  // namespace mgb200

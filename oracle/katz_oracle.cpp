// oracle/katz_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Groundwork for the next path.
//
// CPU restatement of the reference's STATIC Katz centrality, katz_centrality.get(alpha = 0.2, epsilon = 1e-2):
//   mage/cpp/katz_centrality_module/algorithm/katz.cpp  SetKatz :389-410, KatzCentralityLoop :222-251,
//   Converged :163-211, context Init / AddIteration :33-68, MaxDegree :133-143, WrapResults :127-130
// on plain arrays (the reference keeps unordered_map<id, double> per iteration; with ids 0..n-1 a vector holds the same
// values and the arithmetic is performed in the same order: nodes ascending, in-neighbours in edge-insertion order).
// Pinned bit-exact against the reference's own katz.cpp compiled in place (oracle/_ref/libkatz_ref.so, oracle/Makefile,
// tests/test_katz_oracle.py).  C++ rather than C because the convergence test sorts with std::partial_sort and an
// UNSTABLE sort's tie order is part of the observable behaviour: the same libstdc++ call is made here (:185-189).
//
// What is computed:
//   omega_0 = 1, c_0 = 0;  deg_max = max out-degree;  gamma = deg_max / (1 - alpha^2 * deg_max)          (:399-400)
//   iteration i >= 1:  omega_i[v] = sum_{u -> v} omega_{i-1}[u]                                         (:233-236)
//                      c_i[v] = c_{i-1}[v] + alpha^i * omega_i[v]                                       (:237-239)
//                      lr[v] = c_i[v];  ur[v] = c_i[v] + alpha^(i+1) * omega_i[v] * gamma                (:242-245)
//   stop when, with the vertices sorted by c_i descending, no adjacent pair has ur[a_i] - eps >= lr[a_{i-1}]  (:203-209)
//   (k is forced to the vertex count :170, so the "deactivate" loop :192-198 never runs)
//   no edges -> every centrality is 0 (:395-397).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

extern "C" int oracle_katz(uint64_t n, uint64_t m, const uint64_t *from, const uint64_t *to, double alpha,
                           double epsilon, uint64_t max_iterations_guard, double *centrality_out,
                           uint64_t *iterations_out) {
  for (uint64_t e = 0; e < m; ++e)
    if (from[e] >= n || to[e] >= n) return 2;
  std::vector<double> c_prev(n, 0.0), c(n, 0.0), omega_prev(n, 1.0), omega(n, 0.0), lr(n, 0.0), ur(n, 0.0);
  uint64_t iteration = 0;
  if (m == 0) {
    for (uint64_t v = 0; v < n; ++v) centrality_out[v] = 0.0;
    if (iterations_out) *iterations_out = 0;
    return 0;
  }
  std::vector<std::vector<uint64_t>> in_neighbours(n);
  std::vector<uint64_t> out_degree(n, 0);
  for (uint64_t e = 0; e < m; ++e) {
    in_neighbours[to[e]].push_back(from[e]);
    out_degree[from[e]]++;
  }
  const uint64_t deg_max = *std::max_element(out_degree.begin(), out_degree.end());
  const double gamma = static_cast<double>(deg_max) / (1. - (alpha * alpha * static_cast<double>(deg_max)));
  bool converged = false;
  do {
    iteration++;
    for (uint64_t v = 0; v < n; ++v) {
      double w = 0.0;
      for (const uint64_t u : in_neighbours[v]) w += omega_prev[u];
      omega[v] = w;
      c[v] = c_prev[v] + pow(alpha, static_cast<double>(iteration)) * omega[v];
      lr[v] = c[v];
      ur[v] = c[v] + pow(alpha, static_cast<double>(iteration + 1)) * omega[v] * gamma;
    }
    // Converged(active_nodes, k, epsilon), :163-211, with k = number of vertices
    std::vector<std::pair<uint64_t, double>> active;
    active.reserve(n);
    for (uint64_t v = 0; v < n; ++v) active.emplace_back(v, c[v]);  // std::set order = ascending id
    std::partial_sort(active.begin(), active.end(), active.end(),
                      [](std::pair<uint64_t, double> a, std::pair<uint64_t, double> b) -> bool { return a.second > b.second; });
    converged = true;
    for (std::size_t i = 1; i < active.size(); i++) {
      if (ur[active[i].first] - epsilon >= lr[active[i - 1].first]) {
        converged = false;
        break;
      }
    }
    c_prev = c;
    omega_prev = omega;
    if (max_iterations_guard && iteration >= max_iterations_guard) break;  // test guard only; the reference has none
  } while (!converged);
  for (uint64_t v = 0; v < n; ++v) centrality_out[v] = c[v];
  if (iterations_out) *iterations_out = iteration;
  return converged ? 0 : 3;
}

// The order std::partial_sort(begin, end, end, comp) leaves n (id, key) pairs in, ids ascending on entry -- the exact
// call of Converged (:185-189).  Checker for the product's index-arithmetic restatement of that permutation
// (memgraph_b200/csrc/katz_heap.hpp, tests/test_katz_tie_order.py).
extern "C" int oracle_partial_sort_order(uint64_t n, const double *keys, uint32_t *order_out) {
  std::vector<std::pair<uint64_t, double>> active;
  active.reserve(n);
  for (uint64_t v = 0; v < n; ++v) active.emplace_back(v, keys[v]);
  std::partial_sort(active.begin(), active.end(), active.end(),
                    [](std::pair<uint64_t, double> a, std::pair<uint64_t, double> b) -> bool { return a.second > b.second; });
  for (uint64_t i = 0; i < n; ++i) order_out[i] = static_cast<uint32_t>(active[i].first);
  return 0;
}

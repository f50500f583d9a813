// oracle/katz_ref_shim.cpp -- TEST INFRASTRUCTURE.  C-ABI shim around the UNMODIFIED reference Katz centrality
// (mage/cpp/katz_centrality_module/algorithm/katz.cpp, compiled in place by oracle/Makefile into oracle/_ref/libkatz_ref.so).
// The static entry point katz_alg::SetKatz(GraphView, alpha, epsilon) only needs the header-only mg_graph::Graph;
// katz.cpp also contains the online variant over mgp::Graph, whose mgp_* imports are satisfied by abort() stubs
// (oracle/katz_ref_stubs.c) because nothing on the static path calls them.
#include <cstdint>
#include <exception>
#include <optional>

#include <mg_graph.hpp>

#include "katz.hpp"

namespace mgp::MemoryDispatcher {
thread_local std::optional<mgp_memory *> current_memory;  // defined by the host in a real deployment (mpg_support.cpp:16-19)
}

extern "C" int ref_katz(std::uint64_t n, std::uint64_t m, const std::uint64_t *from, const std::uint64_t *to, double alpha,
                        double epsilon, double *centrality_out) {
  try {
    mg_graph::Graph<> g;
    for (std::uint64_t v = 0; v < n; ++v) g.CreateNode(v);
    for (std::uint64_t e = 0; e < m; ++e) g.CreateEdge(from[e], to[e], mg_graph::GraphType::kDirectedGraph);
    const auto result = katz_alg::SetKatz(g, alpha, epsilon);
    for (const auto &[id, value] : result) centrality_out[id] = value;
    return static_cast<int>(result.size()) == static_cast<int>(n) ? 0 : 1;
  } catch (const std::exception &) {
    return 2;
  }
}

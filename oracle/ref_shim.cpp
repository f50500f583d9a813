// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C-ABI shim around the UNMODIFIED reference algorithm.  oracle/Makefile compiles this file
// together with /root/reference/mage/cpp/pagerank_module/algorithm/pagerank.cpp (read in place,
// never copied) into oracle/_ref/libpagerank_ref.so.  It exposes exactly the two reference entry
// points the unit test uses (pagerank_test.cpp:23-27):
//     pagerank_alg::PageRankGraph(n, m, edges)                         pagerank.cpp:165
//     pagerank_alg::ParallelIterativePageRank(graph, it, d, eps, T)    pagerank.cpp:193
// so Python (ctypes) can (a) pin oracle/pagerank_oracle.c against the real thing and (b) time the
// reference on the box's host cores for bench.py's `--impl reference` arm.
#include <cstdint>
#include <cstring>
#include <exception>
#include <new>
#include <utility>
#include <vector>

#include "pagerank.hpp"  // resolved with -I<reference>/mage/cpp/pagerank_module/algorithm

namespace {
void CopyMessage(char *dst, std::size_t cap, const char *msg) {
  if (dst == nullptr || cap == 0) return;
  std::strncpy(dst, msg, cap - 1);
  dst[cap - 1] = '\0';
}
}  // namespace

extern "C" {

// Returns an owning handle to a reference PageRankGraph, or nullptr (message in err).
void *ref_graph_create(std::uint64_t number_of_nodes, std::uint64_t number_of_edges, std::uint64_t n_input_edges,
                       const std::uint64_t *from, const std::uint64_t *to, char *err, std::size_t err_cap) {
  try {
    std::vector<pagerank_alg::EdgePair> edges;
    edges.reserve(n_input_edges);
    for (std::uint64_t e = 0; e < n_input_edges; ++e) edges.emplace_back(from[e], to[e]);
    return new pagerank_alg::PageRankGraph(number_of_nodes, number_of_edges, edges);
  } catch (const std::exception &e) {
    CopyMessage(err, err_cap, e.what());
    return nullptr;
  }
}

void ref_graph_destroy(void *graph) { delete static_cast<pagerank_alg::PageRankGraph *>(graph); }

// Runs the reference algorithm; rank_out must hold node_count doubles.  0 on success, 1 when the
// reference threw (message in err) -- the module surfaces that text through
// mgp_result_set_error_msg (pagerank_module.cpp:108-111).
int ref_pagerank(const void *graph, std::uint64_t max_iterations, double damping_factor, double stop_epsilon,
                 std::uint32_t number_of_threads, double *rank_out, char *err, std::size_t err_cap) {
  try {
    const auto &g = *static_cast<const pagerank_alg::PageRankGraph *>(graph);
    const std::vector<double> ranks = pagerank_alg::ParallelIterativePageRank(
        g, static_cast<std::size_t>(max_iterations), damping_factor, stop_epsilon, number_of_threads);
    if (!ranks.empty()) std::memcpy(rank_out, ranks.data(), ranks.size() * sizeof(double));
    return 0;
  } catch (const std::exception &e) {
    CopyMessage(err, err_cap, e.what());
    return 1;
  }
}

std::uint64_t ref_graph_node_count(const void *graph) {
  return static_cast<const pagerank_alg::PageRankGraph *>(graph)->GetNodeCount();
}

}  // extern "C"

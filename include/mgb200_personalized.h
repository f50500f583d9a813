/*
 * mgb200_personalized.h -- C ABI of the cuGraph-semantics PageRank variants: what the reference's
 * cugraph.pagerank.get / cugraph.personalized_pagerank.get compute, on the same device-resident layout as the in-tree
 * PageRank (mgb200_pagerank.h), so that the B200 module can stand in for the cuGraph modules too.
 *
 *   reference call site                                                        replaced by
 *   -------------------------------------------------------------------------  --------------------------------
 *   cugraph::pagerank(handle, view, weights, nullopt, nullopt, nullopt,          mgb200_cugraph_pagerank_run() with
 *     alpha, epsilon, max_iterations)   mage/cpp/cugraph_module/algorithms/       n_personalization == 0
 *     pagerank.cu:80-91
 *   cugraph::pagerank(..., personalization = (vertices, values), ...)            mgb200_cugraph_pagerank_run() with
 *     mage/cpp/cugraph_module/algorithms/personalized_pagerank.cu:141-152         the seeds
 *
 * The arithmetic is rapidsai/cugraph's (third party, absent from the reference checkout; restated in
 * oracle/cugraph_pagerank_oracle.c and pinned on the reference's e2e fixtures):
 *   pr_0 = 1/N;  p = values / sum(values) on the seeds, 0 elsewhere (no seeds: p = 1/N)
 *   repeat: dangling = sum of pr over vertices without out-edges
 *           new[v] = alpha * sum_{u->v} pr[u] / outdeg(u) + (alpha * dangling + 1 - alpha) * p[v]
 *           stop when sum_v |new[v] - pr[v]| < epsilon (converged) or after max_iterations (not converged)
 * Differences from the in-tree algorithm (mgb200_pagerank.h): dangling mass is redistributed, the stop test is the L1
 * norm, the result is not renormalised.  Edge weights: a handle from mgb200_graph_create_host_weighted_u32() carries one
 * non-negative FP64 weight per edge (new[v] sums w(u,v) * pr[u] / out_w[u], out_w = sum of u's out-edge weights; the
 * modules' weight_property, default weight 1.0); on any other handle every edge counts 1.  Multi-edges count with
 * multiplicity.  The out-weight sums are accumulated with FP64 atomics at build time: weighted results are
 * reproducible to rounding, not bitwise; unweighted ones are bit-reproducible.
 * Single partition only.  No CPU fallback.  Tolerance-based parity (1e-9 relative vs the restatement): cuGraph's own
 * summation order is not reproducible run to run.
 */
#ifndef MGB200_PERSONALIZED_H
#define MGB200_PERSONALIZED_H

#include "mgb200_pagerank.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct mgb200_cugraph_params {
  uint64_t max_iterations; /* module default 100 */
  double damping_factor;   /* alpha, module default 0.85 */
  double stop_epsilon;     /* L1 tolerance, module default 1e-5 */
  uint64_t n_personalization;              /* 0: plain PageRank (uniform teleport) */
  const uint64_t *personalization_vertices; /* dense vertex ids < node_count (host memory) */
  const double *personalization_values;     /* host memory; normalised to sum 1; the sum must be > 0 */
} mgb200_cugraph_params;

typedef struct mgb200_cugraph_stats {
  uint64_t iterations;
  int converged;        /* 0: max_iterations ended the loop (cuGraph reports this in its metadata) */
  double last_diff_sum; /* L1 distance of the last two iterates */
  double iterate_ms;    /* CUDA-event time of the loop */
  uint64_t kernel_launches;
} mgb200_cugraph_stats;

/* A single-partition graph handle with edge weights (dense 32-bit ids, host memory).  The in-tree PageRank and Katz on
 * such a handle ignore the weights (their reference algorithms are unweighted). */
int mgb200_graph_create_host_weighted_u32(int device, uint64_t n, uint64_t m, const uint32_t *from, const uint32_t *to,
                                         const double *weight, mgb200_graph **out);

/* rank_out: node_count doubles, HOST memory, original vertex-id order. */
int mgb200_cugraph_pagerank_run(mgb200_graph *g, const mgb200_cugraph_params *params, double *rank_out,
                                mgb200_cugraph_stats *stats);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MGB200_PERSONALIZED_H */

/*
 * oracle/cugraph_pagerank_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the PageRank the reference's cuGraph modules call:
 *   mage/cpp/cugraph_module/algorithms/pagerank.cu:80-91               cugraph::pagerank(handle, view, weights, nullopt,
 *   mage/cpp/cugraph_module/algorithms/personalized_pagerank.cu:141-152    personalization, nullopt, alpha, epsilon, max_iter)
 * The arithmetic lives in a THIRD-PARTY dependency that is absent from /root/reference: rapidsai/cugraph, pinned as
 * v22.02.00 (mage/cpp/cugraph_module cugraph.cmake:42) / RAPIDS 25.12 (Dockerfile.cugraph:1).  Its published algorithm
 * (cpp/src/link_analysis/pagerank_impl.cuh), restated here:
 *     pr_0 = 1/N;  p = personalization values / their sum on the given vertices, 0 elsewhere  (absent: p = 1/N)
 *     out_w[u] = sum of the weights of u's out-edges (weight 1 when the graph is unweighted)
 *     repeat:  dangling = sum_{u: out_w[u] == 0} pr[u]      (such a u contributes nothing along its zero-weight edges)
 *              new[v]   = alpha * sum_{(u->v)} w(u,v) * pr[u] / out_w[u]  +  (alpha * dangling + 1 - alpha) * p[v]
 *              diff     = sum_v |new[v] - pr[v]|;  pr = new;  ++iterations
 *              stop if diff < epsilon (converged) or iterations >= max_iterations (not converged)
 * Parity status: pinned on the reference's own call sites and e2e fixtures only -- mage/tests/e2e/pagerank_test/
 * test_cugraph_* and personalized_pagerank_test/test_cugraph_* (tests/golden/cugraph_pagerank_e2e.json, 3-6 printed
 * digits, "validated against NetworkX ground truth" by their authors); the library itself cannot be built or run here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* 0 ok; 2 out of memory; 3 bad endpoint / personalization vertex; 4 personalization values sum to <= 0 */
int oracle_cugraph_pagerank(uint64_t n, uint64_t m, const uint64_t *from, const uint64_t *to, const double *weight,
                            uint64_t n_pers, const uint64_t *pers_vertex, const double *pers_value, double alpha,
                            double epsilon, uint64_t max_iterations, double *rank_out, uint64_t *iterations_out,
                            int *converged_out) {
  if (iterations_out) *iterations_out = 0;
  if (converged_out) *converged_out = 1;
  if (n == 0) return 0;
  double *out_w = calloc(n, sizeof(double)), *pr = malloc(n * sizeof(double)), *nw = malloc(n * sizeof(double));
  double *p = malloc(n * sizeof(double));
  if (!out_w || !pr || !nw || !p) return 2;
  int rc = 0;
  for (uint64_t e = 0; e < m; ++e) {
    if (from[e] >= n || to[e] >= n) {
      rc = 3;
      goto out;
    }
    out_w[from[e]] += weight ? weight[e] : 1.0;
  }
  if (n_pers) {
    double sum = 0.0;
    memset(p, 0, n * sizeof(double));
    for (uint64_t i = 0; i < n_pers; ++i) {
      if (pers_vertex[i] >= n) {
        rc = 3;
        goto out;
      }
      sum += pers_value[i];
    }
    if (!(sum > 0.0)) {
      rc = 4;
      goto out;
    }
    for (uint64_t i = 0; i < n_pers; ++i) p[pers_vertex[i]] += pers_value[i] / sum;
  } else {
    for (uint64_t v = 0; v < n; ++v) p[v] = 1.0 / (double)n;
  }
  for (uint64_t v = 0; v < n; ++v) pr[v] = 1.0 / (double)n;
  uint64_t it = 0;
  int converged = 0;
  for (;;) {
    double dangling = 0.0;
    for (uint64_t v = 0; v < n; ++v)
      if (out_w[v] == 0.0) dangling += pr[v];
    const double spread = alpha * dangling + (1.0 - alpha);
    for (uint64_t v = 0; v < n; ++v) nw[v] = 0.0;
    /* a vertex whose out-edges all weigh 0 is dangling: its mass went into `dangling`, its edges carry nothing */
    for (uint64_t e = 0; e < m; ++e)
      if (out_w[from[e]] != 0.0) nw[to[e]] += (weight ? weight[e] : 1.0) * (pr[from[e]] / out_w[from[e]]);
    double diff = 0.0;
    for (uint64_t v = 0; v < n; ++v) {
      nw[v] = alpha * nw[v] + spread * p[v];
      diff += fabs(nw[v] - pr[v]);
    }
    double *t = pr;
    pr = nw;
    nw = t;
    ++it;
    if (diff < epsilon) {
      converged = 1;
      break;
    }
    if (it >= max_iterations) break;
  }
  memcpy(rank_out, pr, n * sizeof(double));
  if (iterations_out) *iterations_out = it;
  if (converged_out) *converged_out = converged;
out:
  free(out_w);
  free(pr);
  free(nw);
  free(p);
  return rc;
}

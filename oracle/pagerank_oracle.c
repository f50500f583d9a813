/*
 * oracle/pagerank_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's static PageRank
 * (memgraph/memgraph, mage/cpp/pagerank_module/algorithm/pagerank.cpp).  It is the
 * checker the parity tests, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * compare the CUDA path against.  Nothing under memgraph_b200/ may link, import or
 * call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) the 12 known-answer graphs of pagerank_test.cpp:35-65 (tests/golden/
 *       pagerank_unit_vectors.json, tolerance as in mg_test_utils.hpp:107-112), and
 *   (2) the reference's own pagerank.cpp compiled in place into oracle/_ref/
 *       (oracle/Makefile), BIT-EXACT for every thread count, because the restatement
 *       keeps the reference's summation order (see oracle_pagerank below).
 *
 * Each function cites the reference lines it follows.  No reference source text is
 * copied: the reference builds a vector<vector<>> adjacency list and std::thread blocks;
 * this file uses a counting sort and a sequential walk over the same blocks, which
 * yields the same floating-point operation order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ORACLE_OK 0
#define ORACLE_ERR_ZERO_THREADS 1 /* pagerank.cpp:63-65 "Number of threads can't be zero (0)!" */
#define ORACLE_ERR_NOMEM 2
#define ORACLE_ERR_BAD_INPUT 3

typedef struct oracle_graph {
  uint64_t node_count;   /* pagerank.hpp:63  node_count_  */
  uint64_t edge_count;   /* pagerank.hpp:65  edge_count_  (the ctor ARGUMENT m, pagerank.cpp:167) */
  uint64_t stored_edges; /* ordered_edges_.size() */
  uint64_t *edge_to;     /* ordered_edges_[e].first  (graph destination) */
  uint64_t *edge_from;   /* ordered_edges_[e].second (graph source), non-decreasing in e */
  uint64_t *out_degree;  /* pagerank.hpp:70 */
} oracle_graph;

const char *oracle_error_string(int code) {
  switch (code) {
    case ORACLE_OK: return "ok";
    case ORACLE_ERR_ZERO_THREADS: return "Number of threads can't be zero (0)!";
    case ORACLE_ERR_NOMEM: return "oracle: out of memory";
    case ORACLE_ERR_BAD_INPUT: return "oracle: edge endpoint out of range";
    default: return "oracle: unknown error";
  }
}

void oracle_graph_destroy(oracle_graph *g) {
  if (!g) return;
  free(g->edge_to);
  free(g->edge_from);
  free(g->out_degree);
  free(g);
}

/*
 * PageRankGraph(number_of_nodes, number_of_edges, edges) -- pagerank.cpp:165-181.
 * The reference buckets `to` by `from` (insertion order kept inside a bucket) and then
 * emits (to, from) pairs by ascending `from`.  A stable counting sort on `from` gives the
 * identical sequence.  out_degree_[from] counts every input edge (:173).
 * edge_count_ is the argument `number_of_edges`, NOT edges.size() (:167); the block
 * borders are derived from it (:68), so edges past it are never visited.
 */
int oracle_graph_create(uint64_t number_of_nodes, uint64_t number_of_edges, uint64_t n_input_edges,
                        const uint64_t *from, const uint64_t *to, oracle_graph **out) {
  *out = NULL;
  oracle_graph *g = (oracle_graph *)calloc(1, sizeof(*g));
  if (!g) return ORACLE_ERR_NOMEM;
  g->node_count = number_of_nodes;
  g->edge_count = number_of_edges;
  g->stored_edges = n_input_edges;
  g->out_degree = (uint64_t *)calloc(number_of_nodes ? number_of_nodes : 1, sizeof(uint64_t));
  g->edge_to = (uint64_t *)malloc((n_input_edges ? n_input_edges : 1) * sizeof(uint64_t));
  g->edge_from = (uint64_t *)malloc((n_input_edges ? n_input_edges : 1) * sizeof(uint64_t));
  uint64_t *cursor = (uint64_t *)calloc(number_of_nodes + 1, sizeof(uint64_t));
  if (!g->out_degree || !g->edge_to || !g->edge_from || !cursor) {
    free(cursor);
    oracle_graph_destroy(g);
    return ORACLE_ERR_NOMEM;
  }
  for (uint64_t e = 0; e < n_input_edges; e++) {
    if (from[e] >= number_of_nodes || to[e] >= number_of_nodes) {
      free(cursor);
      oracle_graph_destroy(g);
      return ORACLE_ERR_BAD_INPUT; /* the reference would index out of bounds (UB) */
    }
    g->out_degree[from[e]] += 1;
  }
  uint64_t run = 0;
  for (uint64_t v = 0; v < number_of_nodes; v++) {
    cursor[v] = run;
    run += g->out_degree[v];
  }
  for (uint64_t e = 0; e < n_input_edges; e++) {
    uint64_t slot = cursor[from[e]]++;
    g->edge_to[slot] = to[e];
    g->edge_from[slot] = from[e];
  }
  free(cursor);
  *out = g;
  return ORACLE_OK;
}

/* std::thread::hardware_concurrency() as libstdc++ computes it (online processors). */
static uint32_t oracle_hardware_concurrency(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (uint32_t)n : 0u;
}

/*
 * ParallelIterativePageRank -- pagerank.cpp:193-240, with the helpers it calls:
 *   CalculateOptimalBorders  :60-71   borders[i] = i * edge_count / T, error when T == 0
 *   ThreadPageRankIteration  :86-96   block[to] += old_rank[from] / (double)out_degree[from]
 *   AddCurrentBlockToRankNext:104-112 rank_next[i] += d * block[i], blocks in index order
 *   CheckContinueIterate     :138-150 stop at max_iterations, else continue iff any |delta| > eps
 *   NormalizeRank            :156-161 sequential sum, then divide
 * The reference runs the T blocks on T threads but merges them on the main thread in block
 * order, so walking the blocks one after another performs the same additions in the same
 * order: the result is bit-identical to the reference for the same (clamped) T.
 * `iterations_out` (optional) receives number_of_iterations.
 */
int oracle_pagerank(const oracle_graph *g, uint64_t max_iterations, double damping_factor, double stop_epsilon,
                    uint32_t number_of_threads, double *rank_out, uint64_t *iterations_out) {
  const uint64_t n = g->node_count;
  uint32_t hw = oracle_hardware_concurrency();
  if (number_of_threads > hw) number_of_threads = hw; /* :195 */
  if (number_of_threads == 0) return ORACLE_ERR_ZERO_THREADS; /* :63-65 via :197 */
  if (g->edge_count > g->stored_edges) return ORACLE_ERR_BAD_INPUT; /* reference: out-of-bounds read */

  double *rank = (double *)malloc((n ? n : 1) * sizeof(double));
  double *rank_next = (double *)malloc((n ? n : 1) * sizeof(double));
  double *block = (double *)malloc((n ? n : 1) * sizeof(double));
  if (!rank || !rank_next || !block) {
    free(rank); free(rank_next); free(block);
    return ORACLE_ERR_NOMEM;
  }
  for (uint64_t i = 0; i < n; i++) rank[i] = 1.0 / (double)n; /* :199 */

  int continue_iterate = max_iterations != 0; /* :201 */
  uint64_t iterations = 0;
  while (continue_iterate) {
    const double base = (1.0 - damping_factor) / (double)n; /* :221-222 */
    for (uint64_t i = 0; i < n; i++) rank_next[i] = base;
    for (uint32_t b = 0; b < number_of_threads; b++) {
      const uint64_t lo = (uint64_t)b * g->edge_count / number_of_threads;       /* :68 */
      const uint64_t hi = (uint64_t)(b + 1) * g->edge_count / number_of_threads; /* :68 */
      memset(block, 0, n * sizeof(double));                                       /* :88 */
      for (uint64_t e = lo; e < hi; e++) {
        const uint64_t src = g->edge_from[e];
        block[g->edge_to[e]] += rank[src] / (double)g->out_degree[src]; /* :93 */
      }
      for (uint64_t i = 0; i < n; i++) rank_next[i] += damping_factor * block[i]; /* :109-111 */
    }
    /* CompleteRankNext (:122-126) is a no-op: rank_next already has n entries. */
    double *t = rank; rank = rank_next; rank_next = t; /* :229 */
    iterations++;
    /* CheckContinueIterate(rank, rank_next, ...) -- after the swap, :231 */
    continue_iterate = 0;
    if (iterations != max_iterations) {
      for (uint64_t i = 0; i < n; i++) {
        if (fabs(rank[i] - rank_next[i]) > stop_epsilon) { continue_iterate = 1; break; }
      }
    }
  }
  double sum = 0.0; /* :157 std::accumulate, left to right */
  for (uint64_t i = 0; i < n; i++) sum += rank[i];
  for (uint64_t i = 0; i < n; i++) rank_out[i] = rank[i] / sum; /* :158-160 */
  if (iterations_out) *iterations_out = iterations;
  free(rank); free(rank_next); free(block);
  return ORACLE_OK;
}

/*
 * Convenience: build + run + free, the shape of the reference's unit test
 * (pagerank_test.cpp:23-27: PageRankGraph(n, m, edges) then ParallelIterativePageRank(graph)).
 */
int oracle_pagerank_edges(uint64_t number_of_nodes, uint64_t number_of_edges, uint64_t n_input_edges,
                          const uint64_t *from, const uint64_t *to, uint64_t max_iterations, double damping_factor,
                          double stop_epsilon, uint32_t number_of_threads, double *rank_out,
                          uint64_t *iterations_out) {
  oracle_graph *g = NULL;
  int rc = oracle_graph_create(number_of_nodes, number_of_edges, n_input_edges, from, to, &g);
  if (rc != ORACLE_OK) return rc;
  rc = oracle_pagerank(g, max_iterations, damping_factor, stop_epsilon, number_of_threads, rank_out, iterations_out);
  oracle_graph_destroy(g);
  return rc;
}

/*
 * CreatePageRankGraph's id mapping (pagerank_module.cpp:18-54), restated on plain arrays so
 * the module-level tests can predict what the drop-in must emit:
 *   - dense id of a vertex = its position in the vertex-iteration order (:41-42);
 *   - an edge endpoint whose gid was never visited maps to dense id 0, because the
 *     reference looks it up with unordered_map::operator[] (:50).
 * `visited_gids[n]` is the iteration order; edges are (src_gid, dst_gid).
 * Outputs dense (from, to) arrays of length m.  O(n log n + m log n) via sort + bsearch.
 */
typedef struct { int64_t gid; uint64_t dense; } oracle_gid_pair;
static int oracle_gid_cmp(const void *a, const void *b) {
  int64_t x = ((const oracle_gid_pair *)a)->gid, y = ((const oracle_gid_pair *)b)->gid;
  return x < y ? -1 : (x > y ? 1 : 0);
}
int oracle_map_gids(uint64_t n, const int64_t *visited_gids, uint64_t m, const int64_t *src_gid,
                    const int64_t *dst_gid, uint64_t *from_out, uint64_t *to_out) {
  oracle_gid_pair *tab = (oracle_gid_pair *)malloc((n ? n : 1) * sizeof(*tab));
  if (!tab) return ORACLE_ERR_NOMEM;
  for (uint64_t i = 0; i < n; i++) { tab[i].gid = visited_gids[i]; tab[i].dense = i; }
  qsort(tab, n, sizeof(*tab), oracle_gid_cmp);
  for (uint64_t e = 0; e < m; e++) {
    oracle_gid_pair key;
    oracle_gid_pair *hit;
    key.gid = src_gid[e]; key.dense = 0;
    hit = (oracle_gid_pair *)bsearch(&key, tab, n, sizeof(*tab), oracle_gid_cmp);
    from_out[e] = hit ? hit->dense : 0;
    key.gid = dst_gid[e];
    hit = (oracle_gid_pair *)bsearch(&key, tab, n, sizeof(*tab), oracle_gid_cmp);
    to_out[e] = hit ? hit->dense : 0;
  }
  free(tab);
  return ORACLE_OK;
}

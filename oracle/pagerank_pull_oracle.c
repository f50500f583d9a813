/*
 * oracle/pagerank_pull_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A multi-threaded CPU checker for graphs too big for the pinned restatement (oracle/pagerank_oracle.c walks the
 * reference's thread blocks sequentially: ~8 minutes at RMAT scale-26).  Same recurrence, stop test and normalisation
 * as the reference (mage/cpp/pagerank_module/algorithm/pagerank.cpp):
 *     rank_0 = 1/N                                                          :199
 *     rank_{k+1}[v] = (1-d)/N + d * sum_{(u->v)} rank_k[u] / outdeg(u)      :86-96, :104-112, :221-226
 *     continue iff k+1 != max_iterations and exists v: |delta_v| > eps      :138-150
 *     result = rank / sum(rank)                                             :156-161
 * but in PULL form over a CSC (sources of each destination sorted ascending), rows split over pthreads.  Only the
 * ORDER of the additions inside a row differs from the reference, exactly like the CUDA path; SURVEY 8c measured that
 * difference at 8e-13.  Parity status: validated, not independent -- tests/test_oracle.py pins it against
 * pagerank_oracle.c (which is bit-exact against the reference compiled in place) to 1e-12 relative with equal iteration
 * counts at every size the pinned oracle finishes quickly; the scale-26 GPU test then uses it at 1e-9.
 * SURVEY 8c prescribes exactly this ladder for config #3.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t n, m;
  uint64_t *ptr;    /* [n + 1] in-edge offsets */
  uint32_t *src;    /* [m] sources, ascending inside a row */
  uint32_t *outdeg; /* [n] */
} pull_graph;

typedef struct {
  const pull_graph *g;
  const uint64_t *from, *to;
  uint64_t lo, hi; /* edge range (build) or row range (iterate) */
  uint64_t *cursor;
  const double *contrib;
  const double *rank;
  double *next;
  double base, damping, local_max;
  int any;
  int phase;
} job;

static int cmp_u32(const void *a, const void *b) {
  const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : x > y;
}

static void *worker(void *arg) {
  job *j = (job *)arg;
  const pull_graph *g = j->g;
  if (j->phase == 0) { /* degree counts */
    for (uint64_t e = j->lo; e < j->hi; ++e) {
      __atomic_fetch_add(&g->outdeg[j->from[e]], 1u, __ATOMIC_RELAXED);
      __atomic_fetch_add(&g->ptr[j->to[e] + 1], 1ull, __ATOMIC_RELAXED);
    }
  } else if (j->phase == 1) { /* fill (row order fixed afterwards by the sort) */
    for (uint64_t e = j->lo; e < j->hi; ++e) {
      const uint64_t pos = __atomic_fetch_add(&j->cursor[j->to[e]], 1ull, __ATOMIC_RELAXED);
      g->src[pos] = (uint32_t)j->from[e];
    }
  } else if (j->phase == 2) { /* sort rows */
    for (uint64_t v = j->lo; v < j->hi; ++v) {
      const uint64_t a = g->ptr[v], b = g->ptr[v + 1];
      if (b - a > 1) qsort(g->src + a, b - a, sizeof(uint32_t), cmp_u32);
    }
  } else { /* one iteration over rows [lo, hi) */
    double mx = 0.0;
    int any = 0;
    for (uint64_t v = j->lo; v < j->hi; ++v) {
      double acc = 0.0;
      for (uint64_t e = g->ptr[v]; e < g->ptr[v + 1]; ++e) acc += j->contrib[g->src[e]];
      const double nx = j->base + j->damping * acc;
      const double d = fabs(nx - j->rank[v]);
      if (d > mx) mx = d;
      if (d == d) any = 1; /* a non-NaN delta exists */
      j->next[v] = nx;
    }
    j->local_max = mx;
    j->any = any;
  }
  return NULL;
}

static void run_jobs(job *jobs, int threads) {
  pthread_t tid[256];
  for (int t = 1; t < threads; ++t)
    if (pthread_create(&tid[t], NULL, worker, &jobs[t]) != 0) {
      worker(&jobs[t]);
      tid[t] = 0;
    }
  worker(&jobs[0]);
  for (int t = 1; t < threads; ++t)
    if (tid[t]) pthread_join(tid[t], NULL);
}

/* 0 ok, 1 zero threads (reference text applies), 2 out of memory, 3 endpoint out of range / n >= 2^32 */
int oracle_pull_pagerank(uint64_t n, uint64_t m, const uint64_t *from, const uint64_t *to, uint64_t max_iterations,
                         double damping, double eps, uint32_t threads, double *rank_out, uint64_t *iterations_out) {
  if (threads == 0) return 1;
  if (threads > 256) threads = 256;
  if (n >= 0xFFFFFFFFull) return 3;
  if (iterations_out) *iterations_out = 0;
  if (n == 0) {
    if (iterations_out) *iterations_out = max_iterations != 0 ? 1 : 0;
    return 0;
  }
  for (uint64_t e = 0; e < m; ++e)
    if (from[e] >= n || to[e] >= n) return 3;
  pull_graph g = {n, m, calloc(n + 1, sizeof(uint64_t)), malloc((m ? m : 1) * sizeof(uint32_t)), calloc(n, sizeof(uint32_t))};
  uint64_t *cursor = malloc(n * sizeof(uint64_t));
  double *rank = malloc(n * sizeof(double)), *next = malloc(n * sizeof(double)), *contrib = malloc(n * sizeof(double));
  job *jobs = calloc(threads, sizeof(job));
  int rc = 0;
  if (!g.ptr || !g.src || !g.outdeg || !cursor || !rank || !next || !contrib || !jobs) {
    rc = 2;
    goto out;
  }
  const int T = (int)threads;
  for (int t = 0; t < T; ++t) {
    jobs[t].g = &g;
    jobs[t].from = from;
    jobs[t].to = to;
    jobs[t].cursor = cursor;
    jobs[t].lo = m * (uint64_t)t / T;
    jobs[t].hi = m * (uint64_t)(t + 1) / T;
    jobs[t].phase = 0;
  }
  run_jobs(jobs, T);
  for (uint64_t v = 0; v < n; ++v) g.ptr[v + 1] += g.ptr[v];
  memcpy(cursor, g.ptr, n * sizeof(uint64_t));
  for (int t = 0; t < T; ++t) jobs[t].phase = 1;
  run_jobs(jobs, T);
  /* rows split so that every thread owns about m / T in-edges */
  {
    uint64_t v = 0;
    for (int t = 0; t < T; ++t) {
      jobs[t].lo = v;
      const uint64_t target = m * (uint64_t)(t + 1) / T;
      while (v < n && (t == T - 1 || g.ptr[v + 1] <= target)) ++v;
      if (t == T - 1) v = n;
      jobs[t].hi = v;
      jobs[t].phase = 2;
    }
  }
  run_jobs(jobs, T);

  const double r0 = 1.0 / (double)n;
  for (uint64_t v = 0; v < n; ++v) rank[v] = r0;
  uint64_t it = 0;
  int cont = max_iterations != 0;
  while (cont) {
    for (uint64_t v = 0; v < n; ++v) contrib[v] = g.outdeg[v] ? rank[v] / (double)g.outdeg[v] : 0.0;
    for (int t = 0; t < T; ++t) {
      jobs[t].phase = 3;
      jobs[t].contrib = contrib;
      jobs[t].rank = rank;
      jobs[t].next = next;
      jobs[t].base = (1.0 - damping) / (double)n;
      jobs[t].damping = damping;
    }
    run_jobs(jobs, T);
    double mx = 0.0;
    int any = 0;
    for (int t = 0; t < T; ++t) {
      if (jobs[t].local_max > mx) mx = jobs[t].local_max;
      any |= jobs[t].any;
    }
    double *tmp = rank;
    rank = next;
    next = tmp;
    ++it;
    cont = it != max_iterations && any && mx > eps;
  }
  {
    double sum = 0.0;
    for (uint64_t v = 0; v < n; ++v) sum += rank[v]; /* sequential, like std::accumulate (:157) */
    for (uint64_t v = 0; v < n; ++v) rank_out[v] = rank[v] / sum;
  }
  if (iterations_out) *iterations_out = it;
out:
  free(g.ptr);
  free(g.src);
  free(g.outdeg);
  free(cursor);
  free(rank);
  free(next);
  free(contrib);
  free(jobs);
  return rc;
}

/*
 * oracle/rmat_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restatement of the synthetic workload's RMAT edge stream (the definition lives in the product's
 * memgraph_b200/csrc/rmat.hpp; the reference delegates generation to cugraph::generate_rmat_edgelist,
 * mage/cpp/cugraph_module/algorithms/graph_generator.cu:143-145 -- third party, absent -- so the bit stream is
 * this repo's own: one splitmix64 per (seed, edge, level pair), high/low 32 bits pick the quadrant of two
 * consecutive levels, most significant vertex bit first).  It exists so that bench.py's `--impl reference` arm and
 * the CPU checkers can produce the benchmark graph WITHOUT loading the product library; tests/test_oracle.py pins
 * it byte for byte against mgb200_rmat_generate_host.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

static uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static uint32_t scaled(double x) {
  const double two32 = 4294967296.0;
  if (x <= 0.0) return 0u;
  if (x >= two32 - 1.0) return 0xffffffffu;
  return (uint32_t)x;
}

typedef struct {
  uint32_t scale, t1, t2, t3;
  uint64_t first, count, seed;
  uint64_t *from, *to;
} rmat_job;

static void *rmat_worker(void *arg) {
  const rmat_job *j = (const rmat_job *)arg;
  for (uint64_t i = 0; i < j->count; ++i) {
    const uint64_t key = mix(j->seed ^ ((j->first + i) * 0xD1342543DE82EF95ull));
    uint64_t s = 0, d = 0;
    for (uint32_t level = 0; level < j->scale; level += 2) {
      const uint64_t h = mix(key + (uint64_t)(level >> 1) * 0xA24BAED4963EE407ull);
      uint32_t u = (uint32_t)(h >> 32);
      s = (s << 1) | (u >= j->t2);
      d = (d << 1) | ((u >= j->t1 && u < j->t2) || u >= j->t3);
      if (level + 1 < j->scale) {
        u = (uint32_t)h;
        s = (s << 1) | (u >= j->t2);
        d = (d << 1) | ((u >= j->t1 && u < j->t2) || u >= j->t3);
      }
    }
    j->from[i] = s;
    j->to[i] = d;
  }
  return NULL;
}

/* edges [first, first + count) of RMAT(a, b, c) with 2^scale vertices as uint64 (from, to); `threads` host threads */
int oracle_rmat_edges(uint32_t scale, uint64_t first, uint64_t count, uint64_t seed, double a, double b, double c,
                      uint32_t threads, uint64_t *from, uint64_t *to) {
  if (scale == 0 || scale > 31 || (count && (!from || !to))) return 3;
  const double two32 = 4294967296.0;
  if (threads == 0) threads = 1;
  if (threads > 64) threads = 64;
  if (count < 65536) threads = 1;
  rmat_job jobs[64];
  pthread_t tid[64];
  const uint64_t per = (count + threads - 1) / threads;
  uint32_t started = 0;
  for (uint32_t t = 0; t < threads; ++t) {
    const uint64_t lo = (uint64_t)t * per;
    if (lo >= count) break;
    rmat_job *j = &jobs[t];
    j->scale = scale;
    j->t1 = scaled(a * two32);
    j->t2 = scaled((a + b) * two32);
    j->t3 = scaled((a + b + c) * two32);
    j->first = first + lo;
    j->count = (lo + per <= count) ? per : count - lo;
    j->seed = seed;
    j->from = from + lo;
    j->to = to + lo;
    if (threads == 1 || pthread_create(&tid[t], NULL, rmat_worker, j) != 0) {
      rmat_worker(j); /* no thread available: do the slice here */
      tid[t] = 0;
    } else {
      ++started;
    }
  }
  if (threads > 1)
    for (uint32_t t = 0; t < threads; ++t)
      if ((uint64_t)t * per < count && tid[t]) pthread_join(tid[t], NULL);
  (void)started;
  return 0;
}

/*
 * mgb200_bfs.h -- C ABI of the B200-native breadth-first expansion (the NEXT path after PageRank, SURVEY 8f-2 /
 * BASELINE config #5: "BFS expand on RMAT scale-24, 1xB200, distances bit-exact vs the reference").
 *
 * Replaces the arithmetic of the reference's engine operator for `MATCH (a)-[*BFS lower..upper]->(b)` with an
 * unbound target -- SingleSourceShortestPathCursor, src/query/plan/operator.cpp:2692-2912: level-synchronous
 * expansion from one source, visited-on-discovery, direction OUT / IN / BOTH, depth bounds (default 1..inf),
 * the source itself is never produced.  The operator lives inside the engine (no plugin boundary exists for it),
 * so this ABI is what a new read procedure (e.g. gpu_bfs.distances(source)) would call; see INTEGRATION.md.
 *
 * Output: dist[v] = the BFS depth at which the reference emits v, or -1 when it does not (unreachable, the
 * source, or depth outside [lower_bound, upper_bound]).  Integer work: parity is bit-exact.
 */
#ifndef MGB200_BFS_H
#define MGB200_BFS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define MGB200_BFS_OUT 0  /* EdgeAtom::Direction::OUT  */
#define MGB200_BFS_IN 1   /* EdgeAtom::Direction::IN   */
#define MGB200_BFS_BOTH 2 /* EdgeAtom::Direction::BOTH */

typedef struct mgb200_bfs_graph mgb200_bfs_graph; /* device-resident CSR (out) + CSC (in) */

typedef struct mgb200_bfs_stats {
  uint32_t levels;            /* depth of the deepest discovered vertex                                   */
  uint32_t top_down_levels;   /* levels expanded frontier-first (queue, atomicCAS on the depth array)     */
  uint32_t bottom_up_levels;  /* levels expanded by the unvisited vertices looking for a frontier parent  */
  uint64_t reached;           /* vertices discovered (source excluded), before the bounds filter          */
  uint64_t edges_inspected;   /* adjacency entries actually read                                          */
  double traverse_ms;         /* CUDA-event time of the expansion (depth array initialised .. last level) */
  uint64_t kernel_launches;
} mgb200_bfs_stats;

/* COO on the device (uint32 endpoints < n, read not consumed) -> CSR + CSC, both sorted. */
int mgb200_bfs_graph_create_device(int device, uint64_t n, uint64_t m, const uint32_t *d_from, const uint32_t *d_to,
                                   mgb200_bfs_graph **out);
/* COO in host memory (uint64 endpoints). */
int mgb200_bfs_graph_create_host(int device, uint64_t n, uint64_t m, const uint64_t *from, const uint64_t *to,
                                 mgb200_bfs_graph **out);
void mgb200_bfs_graph_destroy(mgb200_bfs_graph *g);

/* lower_bound / upper_bound: pass 1 and INT64_MAX for the reference's defaults (operator.cpp:2824-2828).
 * dist_out: n int32 values, host memory (dist_on_device = 0) or memory of the graph's device (1). */
int mgb200_bfs_run(mgb200_bfs_graph *g, uint64_t source, int direction, int64_t lower_bound, int64_t upper_bound,
                   int32_t *dist_out, int dist_on_device, mgb200_bfs_stats *stats);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MGB200_BFS_H */

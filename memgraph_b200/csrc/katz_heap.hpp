// memgraph_b200/csrc/katz_heap.hpp -- the permutation std::partial_sort leaves behind, as plain index arithmetic.
//
// The reference's Katz convergence test (mage/cpp/katz_centrality_module/algorithm/katz.cpp:163-211) sorts the
// active vertices with std::partial_sort(begin, end, end, [](a, b) { return a.second > b.second; }) and then
// compares NEIGHBOURS in that order.  partial_sort is not stable, so when centralities tie, which of the tied vertices
// comes first is decided by the heap algorithm -- and in one corner case (katz.cu, "ambiguous groups") that decides
// whether the loop stops.  With middle == last, libstdc++'s partial_sort is make_heap followed by sort_heap
// (bits/stl_algo.h __partial_sort -> __heap_select + __sort_heap; bits/stl_heap.h): a bottom-up heap construction and
// repeated pop, both built on "sift the hole down to a leaf along the preferred children, then push the value back
// up".  This header restates that procedure over an index array so that the product can reproduce the tie order
// WITHOUT running the sort on the CPU: the same function runs single-threaded on the device in the (measure-zero)
// case that needs it, and on the host in tests/test_katz_tie_order.py, where it is compared with the real
// std::partial_sort (oracle/katz_oracle.cpp oracle_partial_sort_order).
#pragma once

#include <cstdint>

#ifndef MGB_HD
#define MGB_HD __host__ __device__ __forceinline__
#endif

namespace mgb200 {

struct KatzTieOrder {
  // comp(a, b) of the reference: "a goes before b" iff key[a] > key[b]
  MGB_HD static bool before(const double *key, uint32_t a, uint32_t b) { return key[a] > key[b]; }

  // value climbs from `hole` towards `top` while its parent orders before it
  MGB_HD static void push_up(uint32_t *idx, int64_t hole, int64_t top, uint32_t value, const double *key) {
    int64_t parent = (hole - 1) / 2;
    while (hole > top && before(key, idx[parent], value)) {
      idx[hole] = idx[parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    idx[hole] = value;
  }

  // the hole walks down to a leaf along the children that do NOT order before their sibling, then `value` is pushed up
  MGB_HD static void sift(uint32_t *idx, int64_t hole, int64_t len, uint32_t value, const double *key) {
    const int64_t top = hole;
    int64_t child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);                                    // right child
      if (before(key, idx[child], idx[child - 1])) --child;      // prefer the left one if the right orders before it
      idx[hole] = idx[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {               // a last parent with a single (left) child
      child = 2 * (child + 1);
      idx[hole] = idx[child - 1];
      hole = child - 1;
    }
    push_up(idx, hole, top, value, key);
  }

  // idx[0..len) in: any order (the reference: ascending vertex id); out: the order partial_sort(first, last, last) leaves
  MGB_HD static void run(uint32_t *idx, int64_t len, const double *key) {
    if (len < 2) return;
    for (int64_t parent = (len - 2) / 2;; --parent) {             // make_heap
      sift(idx, parent, len, idx[parent], key);
      if (parent == 0) break;
    }
    for (int64_t last = len - 1; last > 0; --last) {              // sort_heap: pop the top into the shrinking tail
      const uint32_t value = idx[last];
      idx[last] = idx[0];
      sift(idx, 0, last, value, key);
    }
  }
};

}  // namespace mgb200

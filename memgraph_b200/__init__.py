"""memgraph_b200 -- B200-native static PageRank behind Memgraph's query-module ABI.

Python mirror of the reference's algorithm interface (mage/cpp/pagerank_module/algorithm/pagerank.hpp):

    reference (C++)                                            here
    ---------------------------------------------------------  -------------------------------------------
    pagerank_alg::PageRankGraph(n, m, edges)      :40          memgraph_b200.PageRankGraph(n, m, edges)
    pagerank_alg::ParallelIterativePageRank(g,    :107-109     memgraph_b200.parallel_iterative_pagerank(g,
        max_iterations=100, damping_factor=0.85,                   max_iterations=100, damping_factor=0.85,
        stop_epsilon=10e-6, number_of_threads=1)                   stop_epsilon=1e-5, number_of_threads=1)

All compute goes through the C ABI in include/mgb200_pagerank.h (memgraph_b200/_build/
libmgb200_pagerank.so, CUDA sm_100a); there is no CPU path -- without the library or a CUDA device
every call raises.
"""
from .pagerank import (MgB200Error, PageRankGraph, RunStats, device_count, pagerank_from_edges,  # noqa: F401
                       parallel_iterative_pagerank, rmat_edges_device, rmat_edges_host)

__all__ = ["MgB200Error", "PageRankGraph", "RunStats", "device_count", "pagerank_from_edges",
           "parallel_iterative_pagerank", "rmat_edges_device", "rmat_edges_host"]

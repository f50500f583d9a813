"""One process per GPU: the host-side plumbing around the partitioned C ABI.

The data path has NO host-side collective: partitions push their contributions into each other's
exchange windows over NVLink from inside the kernels and agree on convergence through flag pages in
peer memory (csrc/pagerank_kernels.cu).  torch.distributed (NCCL on the GPUs, gloo in the CPU tests)
is used only for (a) swapping the 64-byte CUDA IPC handles once per graph and (b) collecting the
result slices.  `dist` is the torch.distributed module (or any object with the same
all_gather_object / get_rank / get_world_size / barrier functions)."""
import numpy as np


def exchange_windows(graph, dist):
    """Connects `graph` (one partition) to the partitions held by the other ranks."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    if world == 1:
        return
    if graph.info["part_world"] != world or graph.info["part_rank"] != rank:
        raise ValueError("partition (rank, world) must equal the process group's (rank, world)")
    handles = [None] * world
    dist.all_gather_object(handles, graph.export_window())
    handles[rank] = None
    graph.connect_peers(ipc_handles=handles)
    dist.barrier()


def assemble_ranks(n, local_ranks, local_vertices, dist, dst=0):
    """Scatters every partition's (vertex, rank) slice into one array in original vertex order on
    rank `dst` (None elsewhere).  Each vertex is owned by exactly one partition."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    pieces = [None] * world
    dist.all_gather_object(pieces, (np.asarray(local_vertices), np.asarray(local_ranks)))
    if rank != dst:
        return None
    out = np.full(n, np.nan, dtype=np.float64)
    seen = 0
    for verts, ranks in pieces:
        if len(verts) != len(ranks):
            raise ValueError("slice length mismatch")
        out[verts.astype(np.int64)] = ranks
        seen += len(verts)
    if seen != n or np.isnan(out).any() and not np.isnan(np.concatenate([r for _, r in pieces])).any():
        raise ValueError(f"partitions cover {seen} of {n} vertices")
    return out


def partition_rows(n, world):
    """Rows owned by each partition: sorted positions are dealt round-robin, so rank q owns
    ceil((n - q) / world) rows (mirrors Dealer::count in csrc/graph_build.cu)."""
    return [(n - q + world - 1) // world if n > q else 0 for q in range(world)]

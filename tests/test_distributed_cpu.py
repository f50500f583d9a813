"""world_size-2 gloo test of the N>1 host logic (no GPU): IPC-handle exchange and result assembly,
with a stand-in partition object (the device path is covered by tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakePartition:
    """Stands in for PageRankGraph: records what the plumbing hands to connect_peers."""

    def __init__(self, rank, world, n):
        from memgraph_b200.distributed import partition_rows
        self.info = {"part_rank": rank, "part_world": world, "node_count": n, "local_rows": partition_rows(n, world)[rank]}
        self.connected = None

    def export_window(self):
        return bytes([self.info["part_rank"]]) * 64

    def connect_peers(self, ipc_handles=None, local_graphs=None):
        self.connected = ipc_handles


def _worker(rank, world, port, n, tmpdir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from memgraph_b200.distributed import assemble_ranks, exchange_windows, partition_rows
    part = FakePartition(rank, world, n)
    exchange_windows(part, dist)
    assert part.connected[rank] is None
    for q in range(world):
        if q != rank:
            assert part.connected[q] == bytes([q]) * 64
    # ownership: sorted position p belongs to rank p % world; emulate with the identity order
    owned = np.arange(rank, n, world, dtype=np.uint32)
    assert len(owned) == partition_rows(n, world)[rank]
    ranks = owned.astype(np.float64) / n
    full = assemble_ranks(n, ranks, owned, dist, dst=0)
    if rank == 0:
        assert np.array_equal(full, np.arange(n) / n)
        np.save(os.path.join(tmpdir, "ok.npy"), full)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11])
def test_two_rank_plumbing_gloo(tmp_path, n):
    port = 29500 + (os.getpid() % 500) + n
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok.npy")


def test_partition_rows_cover_all_vertices():
    """Python plan == the library's partition arithmetic (host-only C ABI call), ranges tile [0, n) in rank order."""
    import ctypes
    from memgraph_b200 import _native as N
    from memgraph_b200.distributed import partition_rows
    lib = N.lib()
    for n in [0, 1, 7, 8, 1000003, 2**26]:
        for world in [1, 2, 3, 8]:
            rows = partition_rows(n, world)
            assert sum(rows) == n and max(rows) - min(rows) <= 1
            nxt = 0
            for q in range(world):
                first, cnt = ctypes.c_uint64(), ctypes.c_uint64()
                assert lib.mgb200_partition_range(n, world, q, ctypes.byref(first), ctypes.byref(cnt)) == 0
                assert (first.value, cnt.value) == (nxt, rows[q])
                nxt += cnt.value
            assert nxt == n
    assert lib.mgb200_partition_range(10, 9, 0, None, None) != 0 and b"invalid partition" in lib.mgb200_last_error()

"""Host-only check of the label <-> (owner, local row) arithmetic both labellings use on the device (csrc/core.hpp
RowMap, reached through mgb200_partition_locate / mgb200_partition_label): every label has exactly one owner, the local
rows of an owner are 0..rows-1 in ascending label order, and the two directions are inverse of each other."""
import ctypes

import numpy as np
import pytest

from memgraph_b200 import _native as N


def table(n, heavy, world, mode):
    lib = N.lib()
    owner = np.zeros(n, dtype=np.int64)
    local = np.zeros(n, dtype=np.int64)
    o, l = ctypes.c_uint32(), ctypes.c_uint64()
    for label in range(n):
        assert lib.mgb200_partition_locate(n, heavy, world, mode, label, ctypes.byref(o), ctypes.byref(l)) == 0
        owner[label], local[label] = o.value, l.value
    return owner, local


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,heavy,world", [(1, 0, 1), (1, 1, 2), (31, 0, 2), (32, 0, 8), (33, 1, 8), (1000, 37, 8),
                                           (1000, 0, 3), (1000, 1000, 4), (4099, 64, 5), (257, 200, 7), (64, 0, 2)])
def test_partition_map_is_a_bijection(n, heavy, world, mode):
    lib = N.lib()
    owner, local = table(n, heavy, world, mode)
    rows_total = 0
    for q in range(world):
        rows = ctypes.c_uint64()
        assert lib.mgb200_partition_label(n, heavy, world, mode, q, 0, None, ctypes.byref(rows)) == 0
        mine = np.flatnonzero(owner == q)
        assert rows.value == len(mine)
        rows_total += rows.value
        # local rows are 0..rows-1 in ascending label order ...
        assert np.array_equal(local[mine], np.arange(len(mine)))
        # ... and the inverse map agrees
        lab = ctypes.c_uint64()
        for r, label in enumerate(mine):
            assert lib.mgb200_partition_label(n, heavy, world, mode, q, r, ctypes.byref(lab), None) == 0
            assert lab.value == label
        assert lib.mgb200_partition_label(n, heavy, world, mode, q, len(mine), ctypes.byref(lab), None) != 0
    assert rows_total == n
    if mode == 0:  # the dealt ranges are what mgb200_partition_range describes
        first, cnt = ctypes.c_uint64(), ctypes.c_uint64()
        for q in range(world):
            assert lib.mgb200_partition_range(n, world, q, ctypes.byref(first), ctypes.byref(cnt)) == 0
            mine = np.flatnonzero(owner == q)
            assert cnt.value == len(mine) and (len(mine) == 0 or first.value == mine[0])
            assert len(mine) == 0 or mine[-1] - mine[0] + 1 == len(mine)  # contiguous
    else:
        # heavy labels are dealt singly, the rest in whole blocks of 32 consecutive labels (one SELL slice = one block)
        assert np.array_equal(owner[:heavy], np.arange(heavy) % world)
        rest = owner[heavy:]
        for b in range(0, len(rest), 32):
            assert (rest[b:b + 32] == (b // 32) % world).all()
        # rows per partition differ by at most one block (+1 heavy row)
        counts = np.bincount(owner, minlength=world)
        assert counts.max() - counts.min() <= 33


def test_partition_map_rejects_bad_arguments():
    lib = N.lib()
    o, l = ctypes.c_uint32(), ctypes.c_uint64()
    assert lib.mgb200_partition_locate(10, 0, 0, 1, 0, ctypes.byref(o), ctypes.byref(l)) != 0
    assert lib.mgb200_partition_locate(10, 0, 9, 1, 0, ctypes.byref(o), ctypes.byref(l)) != 0
    assert lib.mgb200_partition_locate(10, 11, 2, 1, 0, ctypes.byref(o), ctypes.byref(l)) != 0  # heavy > n
    assert lib.mgb200_partition_locate(10, 0, 2, 1, 10, ctypes.byref(o), ctypes.byref(l)) != 0  # label >= n

#!/usr/bin/env python3
"""CPU study for the NEXT step of the multi-GPU exchange (profiles/r02_multi_gpu.md, DESIGN.md open item 1): the compacted
exchange.  Today a partition stores a new contribution at the SAME label slot of every peer that gathers it (push mask):
half the bytes of an unconditional push, but as sparse 8-byte pieces.  Compacted: every receiver q keeps, per owner p, only
the labels of p it gathers, in p's row order; the sender derives the slot as  base[q][row / 32] + popc(need bits of the
32-row group below the row)  -- dense, monotone, 256-byte stores.  This script checks on a real RMAT graph that the sender-side
arithmetic and the receiver-side numbering (exclusive rank among the labels q reads inside p's range) agree for every
(row, peer), and reports the bytes per iteration.  numpy only; RMAT stream from oracle/rmat_oracle.c."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
from _checkers import oracle_rmat_edges  # noqa: E402


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n, m = 1 << scale, 16 << scale
    f, t = oracle_rmat_edges(scale, m)
    f = f.astype(np.int64)
    t = t.astype(np.int64)
    indeg = np.bincount(t, minlength=n)
    outdeg = np.bincount(f, minlength=n)
    order = np.lexsort((np.arange(n), -outdeg, -indeg))  # in-degree desc, out-degree desc, id asc
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    print(f"RMAT scale-{scale}: N={n} E={m}")
    for P in (2, 4, 8):
        base, extra = n // P, n % P
        start = np.array([q * base + min(q, extra) for q in range(P + 1)], dtype=np.int64)
        start[P] = n
        owner_of_pos = pos % P
        label = start[owner_of_pos] + pos // P           # dealt ranges (graph_build.cu Dealer)
        owner = owner_of_pos                              # owner of each vertex
        src_l, dst_owner = label[f], owner[t]
        # readers[label] bit q: partition q has an in-edge from the vertex with that label
        readers = np.zeros(n, dtype=np.uint8)
        np.bitwise_or.at(readers, src_l, (1 << dst_owner).astype(np.uint8))
        active = np.zeros(n, dtype=bool)
        active[label] = outdeg > 0                        # sinks are never pushed
        total_pairs = dense_bytes = mask_bytes = full_bytes = 0
        ok = True
        for p in range(P):
            lo, hi = start[p], start[p + 1]
            rows = hi - lo
            rd = readers[lo:hi]
            for q in range(P):
                if q == p:
                    continue
                need = ((rd >> q) & 1).astype(bool) & active[lo:hi]
                # receiver side: slot = exclusive rank of the label among the labels of p that q reads
                recv_slot = np.cumsum(need) - need
                # sender side: per 32-row group a base (prefix of popcounts), inside the group popc of the lower need bits
                pad = (-rows) % 32
                g = np.concatenate([need, np.zeros(pad, dtype=bool)]).reshape(-1, 32)
                group_base = np.concatenate([[0], np.cumsum(g.sum(axis=1))[:-1]])
                within = np.cumsum(g, axis=1) - g
                send_slot = (group_base[:, None] + within).reshape(-1)[:rows]
                ok &= bool(np.array_equal(send_slot[need], recv_slot[need]))
                cnt = int(need.sum())
                total_pairs += cnt
                dense_bytes += cnt * 8
                # push mask today: 8-byte stores at label positions; NVLink moves whole 32-byte sectors
                sectors = np.unique((np.nonzero(need)[0] + lo) // 4).size
                mask_bytes += sectors * 32
                full_bytes += int(active[lo:hi].sum()) * 8
        per_gpu = lambda b: b / P / 1e6 * (2 ** 26 / n)   # scaled to scale-26, per GPU per iteration (MB)
        print(f"P={P}: slots agree for every (row, peer): {ok};  needed (row, peer) pairs {total_pairs / (full_bytes / 8):.1%} of all;  "
              f"per GPU per iteration at scale-26: unconditional {per_gpu(full_bytes):.0f} MB, masked 8-byte stores "
              f"{per_gpu(dense_bytes):.0f} MB payload in {per_gpu(mask_bytes):.0f} MB of touched sectors, compacted {per_gpu(dense_bytes):.0f} MB dense")


if __name__ == "__main__":
    main()

"""CPU study for the multi-GPU labelling: how evenly do dealing schemes spread the
in-edges, the heavy-row edges and the SELL padding over P partitions of an RMAT graph?

  round-robin : sorted position pos -> owner pos % P                      (what graph_build.cu does today)
  block-cyclic: heavy rows pos % P, the rest ((pos - H) // B) % P         (labels stay the global degree order, so
                "hot" is a global label prefix on every partition and no per-gather owner lookup is needed)

usage: python scripts/partition_balance.py [scale=22] [P=8]"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from memgraph_b200 import pagerank as pr  # noqa: E402  (host RMAT generator only; no device call)


def study(scale, P, heavy_min=1024):
    n, m = 1 << scale, 16 << scale
    s, t = pr.rmat_edges_host(scale, m)
    indeg = np.bincount(t, minlength=n).astype(np.int64)
    outdeg = np.bincount(s, minlength=n).astype(np.int64)
    order = np.lexsort((np.arange(n), -outdeg, -indeg))
    deg = indeg[order]  # in-degree by sorted position, descending
    H = int(np.searchsorted(-deg, -heavy_min, side="right"))  # heavy rows: positions [0, H)
    nz = int(np.searchsorted(-deg, 0, side="left"))            # rows with in-degree > 0: positions [0, nz)
    pos = np.arange(n)
    print(f"scale {scale}: n={n} m={m} heavy rows={H} ({deg[:H].sum() / m:.1%} of edges) nonzero rows={nz}")

    def report(name, owner):
        edges = np.bincount(owner, weights=deg, minlength=P)
        heavy = np.bincount(owner[:H], weights=deg[:H], minlength=P)
        pad = []
        for q in range(P):
            d = deg[H:nz][owner[H:nz] == q]  # this partition's SELL rows in local order (still descending)
            k = (len(d) + 31) // 32
            width = d[::32][:k]
            pad.append(width.sum() * 32 / max(1, d.sum()) - 1.0)
        print(f"  {name:22s} edges max/mean {edges.max() / edges.mean():.4f}  heavy max/mean "
              f"{heavy.max() / max(1.0, heavy.mean()):.4f}  SELL padding max {max(pad):.3%}")

    report("round-robin", pos % P)
    for B in (32, 256, 1024, 4096):
        owner = np.where(pos < H, pos % P, ((pos - H) // B) % P)
        report(f"block-cyclic B={B}", owner)
    report("blocks, heavy too B=32", (pos // 32) % P)


if __name__ == "__main__":
    study(int(sys.argv[1]) if len(sys.argv) > 1 else 22, int(sys.argv[2]) if len(sys.argv) > 2 else 8)

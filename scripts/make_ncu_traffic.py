#!/usr/bin/env python3
"""profiles/ncu_traffic.json from `ncu --set full` captures: DRAM read + write bytes per launch of sell_rows_kernel.
usage: scripts/make_ncu_traffic.py <world>=<file.ncu-rep> [...]   (e.g. 1=gpurun_out/c20/n1.ncu-rep 8=gpurun_out/c20/lone8.ncu-rep)"""
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def dram_bytes(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "sell_rows" not in d.get("Kernel Name", ""):
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(d[k].replace(",", ""))
            u = units[hdr.index(k)].lower()
            tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        best = tot if best is None else max(best, tot)
    return best


def main():
    import bench
    entries = {}
    for arg in sys.argv[1:]:
        world, path = arg.split("=", 1)
        entries[world] = dram_bytes(path)
    json.dump({"kernel_source_sha256": bench.kernel_source_sha(), "kernel_sources": list(bench.KERNEL_SOURCES),
               "sell_rows_kernel_dram_bytes_per_launch": entries,
               "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of one launch; key = "
                      "number of partitions (N > 1: partition 0 captured alone, MGB200_LONE_WORLD)"},
              open(os.path.join(REPO, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(entries)


if __name__ == "__main__":
    main()

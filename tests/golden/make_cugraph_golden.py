#!/usr/bin/env python3
"""Extracts the reference's cuGraph PageRank e2e fixtures (mage/tests/e2e/{pagerank_test,personalized_pagerank_test}/
test_cugraph_*: input.cyp + test.yml) into tests/golden/cugraph_pagerank_e2e.json.  Run where /root/reference exists."""
import json
import os
import re
import sys

import yaml

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
BASE = os.path.join(REF, "mage", "tests", "e2e")
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_graph(cyp):
    nodes, edges, merged = [], [], set()
    for line in cyp.splitlines():
        line = line.strip()
        m = re.match(r"CREATE \(\w+:Node \{id: (\d+)", line)
        if m and "->" not in line:
            nodes.append(int(m.group(1)))
            continue
        ids = [int(x) for x in re.findall(r"\{id: (\d+)\}", line)]
        if len(ids) == 2 and "->" in line:
            a, b = ids
            for v in (a, b):
                if v not in nodes:
                    nodes.append(v)
            if re.search(r"MERGE \(a\)-", line):  # MERGE de-duplicates the relationship, CREATE does not
                if (a, b) in merged:
                    continue
                merged.add((a, b))
            edges.append((a, b))
    return nodes, edges


def main():
    out = []
    for suite in ("pagerank_test", "personalized_pagerank_test"):
        d = os.path.join(BASE, suite)
        for name in sorted(os.listdir(d)):
            if not name.startswith("test_cugraph"):
                continue
            cyp = open(os.path.join(d, name, "input.cyp")).read()
            spec = yaml.safe_load(open(os.path.join(d, name, "test.yml")))
            nodes, edges = parse_graph(cyp)
            q = spec["query"]
            pers = None
            m = re.search(r"get\(\[([^\]]*)\],\s*\[([^\]]*)\]", q)
            if m:
                vars_ = [v.strip() for v in m.group(1).split(",")]
                vals = [float(v) for v in m.group(2).split(",")]
                ids = {v: int(i) for v, i in re.findall(r"\((\w+):Node \{id: (\d+)\}\)", q)}
                pers = [[ids[v], x] for v, x in zip(vars_, vals)]
            expected = [[row["node_id"], float(row["pagerank"])] for row in (spec.get("output") or [])]
            out.append({"name": f"{suite}/{name}", "nodes": nodes, "edges": edges, "personalization": pers,
                        "expected": expected, "source": f"mage/tests/e2e/{suite}/{name}"})
    json.dump(out, open(os.path.join(HERE, "cugraph_pagerank_e2e.json"), "w"), indent=1)
    print(len(out), "fixtures")


if __name__ == "__main__":
    main()

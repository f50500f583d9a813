"""No-GPU guards on two contracts the round is graded on:
1. the product tree never links, imports or calls the oracle (test infrastructure only);
2. bench.py's reference arm prints exactly one JSON line with the agreed keys, and the B200 arm refuses to
   produce a number without a device."""
import json
import os
import re
import subprocess
import sys

import pytest

import conftest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def product_files():
    out = []
    for root, _dirs, files in os.walk(os.path.join(REPO, "memgraph_b200")):
        if "_build" in root or "__pycache__" in root:
            continue
        out += [os.path.join(root, f) for f in files if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".c"))]
    out += [os.path.join(REPO, "include", f) for f in os.listdir(os.path.join(REPO, "include"))]
    return out


def test_product_never_touches_the_oracle():
    offenders = []
    for path in product_files():
        text = open(path, errors="replace").read()
        code = re.sub(r"//.*|#.*", "", text)                  # comments may MENTION the oracle (provenance notes)
        code = re.sub(r"/\*.*?\*/|\"\"\".*?\"\"\"", "", code, flags=re.S)
        if re.search(r"\boracle[/_.]|_checkers|libpagerank_(oracle|ref)|pagerank_reference\.so|libkatz_ref", code):
            offenders.append(os.path.relpath(path, REPO))
    assert not offenders, offenders
    # the native libraries do not depend on (or embed symbols of) the checker libraries
    for lib in ["libmgb200_pagerank.so", "pagerank.so", "gpu_bfs.so"]:
        p = os.path.join(REPO, "memgraph_b200", "_build", lib)
        if not os.path.exists(p):
            continue
        needed = subprocess.run(["readelf", "-d", p], capture_output=True, text=True).stdout
        syms = subprocess.run(["nm", "-D", p], capture_output=True, text=True).stdout
        assert "oracle" not in needed and "oracle_" not in syms and "ref_pagerank" not in syms, lib


def run_bench(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=REPO)


def test_reference_arm_prints_one_contract_line():
    r = run_bench("--impl", "reference", "--cpu-scale", "14", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout  # exactly ONE line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pagerank_edges_per_second" and d["unit"] == "edges/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] >= 1 and "RMAT scale-" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_never_maps_the_product_library():
    """The reference arm holds only oracle/ code: its RMAT input comes from oracle/rmat_oracle.c, not from the package."""
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--cpu-scale','12','--steps','1','--warmup','0'];"
            "import runpy\n"
            "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit: pass\n"
            "maps=open('/proc/self/maps').read(); bad=[l for l in maps.splitlines() if 'memgraph_b200/_build' in l];"
            "sys.stderr.write('MAPPED_PRODUCT=%d\\n' % len(bad)); assert 'memgraph_b200' not in sys.modules")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MAPPED_PRODUCT=0" in r.stderr, r.stderr[-500:]


def test_reference_arm_reports_its_thread_sweep():
    r = run_bench("--impl", "reference", "--cpu-scale", "13", "--steps", "1", "--warmup", "0")
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    cb = d["cpu_baseline"]
    assert cb["scale"] == 13 and cb["threads"] == cb["cores"] and cb["best_of"] >= 1
    if cb["kind"] == "reference":
        assert "1" in cb["thread_sweep_edges_per_s"] and cb["value_1_thread"] > 0


def test_reference_arm_is_rank0_only_under_torchrun():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=REPO, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""  # other ranks exit 0 without work


@pytest.mark.skipif(conftest.HAVE_GPU, reason="this is the no-device behaviour")
def test_b200_arm_fails_loudly_without_a_device():
    r = run_bench("--steps", "1", "--warmup", "0", "--scale", "10", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0
    assert "CUDA error" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]  # no fabricated JSON line

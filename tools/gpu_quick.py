"""Quick GPU check used while developing (also run under compute-sanitizer): small count + sdbg vs the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import assert_edges_equal, assert_sdbg_equal, make_reads  # noqa: E402
from mitoflex_b200 import lib  # noqa: E402
from oracle import oracle  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [21]
ctx = lib.Context(0)
for k in ks:
    bases, starts = make_reads(1 + k, n_reads, k, genome_len=max(4000, n_reads), max_len=max(150, k + 40))
    r = ctx.upload_reads(bases, starts)
    t0 = time.time()
    e = ctx.count(r, k, 2, want_counting=True)
    t1 = time.time()
    eo = oracle.count(oracle.Reads(bases, starts), k, 2, threads=8)
    print(f"k={k} count: gpu {e.n} edges in {t1 - t0:.3f}s, oracle {eo.n}", flush=True)
    assert_edges_equal(e, eo)
    g = ctx.read2sdbg(r, k, 2)
    go = oracle.read2sdbg(oracle.Reads(bases, starts), k, 2, threads=8)
    print(f"k={k} sdbg: gpu {g.n} items, oracle {go.n}", flush=True)
    assert_sdbg_equal(g, go)
print("OK")

import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import make_reads
from mitoflex_b200 import lib
from oracle import oracle
k, m = 21, 1
bases, starts = make_reads(123, 60000, k, genome_len=3000000, max_len=150, err=0.02)
eo = oracle.count(oracle.Reads(bases, starts), k, m, threads=8)
for pct in (sys.argv[1:] or ["45", "150", "400"]):
    os.environ["MFSDBG_STREAM_LOAD_PCT"] = pct
    ctx = lib.Context(0)
    ctx.set_profiling(True)
    e = ctx.count(ctx.upload_reads(bases, starts), k, m, want_counting=True)
    prof = ctx.last_profile()
    d = e.counting - eo.counting
    nz = np.nonzero(d)[0]
    print(pct, "edges", e.n, eo.n, "counting diff at", nz[:10], d[nz[:10]], "sum gpu", int((e.counting*np.arange(65536)).sum()), "orc", int((eo.counting*np.arange(65536)).sum()),
          {k2: round(v, 3) for k2, v in prof.items() if "count" in k2 or "oversized" in k2 or "fallback" in k2})
    ctx.close()

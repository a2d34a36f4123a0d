#!/usr/bin/env python
"""What rank R of an N-GPU bench run counts, reproduced on ONE GPU: the N read sets of bench.py (seed 1002 + 7919 r) are
generated one after the other, each is cut into super-k-mer records for N destinations of which only destination R has room
(the other regions have capacity 0, their runs are dropped), and count_skm runs over the N regions.  MFSDBG_TRACE=1 shows the
planner's decisions (distinct ratio, buckets that bail)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=16_666_667)
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=2)
    a = ap.parse_args()
    from mitoflex_b200 import lib
    L = lib.load()
    ctx = lib.Context(0)
    ctx.set_profiling(True)
    W, R = a.world, a.rank
    cap = None
    buf = None
    sizes, nkeys = [], 0
    for r in range(W):
        reads = ctx.synth(n_pairs=a.pairs, seed=1002 + 7919 * r)
        if cap is None:
            samp, _ = ctx.skm_scatter(reads, a.k, W, stride=64)
            cap = int(samp.max() * 64 * 1.05) + 65536
            buf = ctx.dev_alloc(W * cap * 8)
        dst = np.full(W, buf, np.uint64)
        caps = np.zeros(W, np.int64)
        dst[R] = buf + r * cap * 8
        caps[R] = cap
        rec, keys = ctx.skm_scatter(reads, a.k, W, dst, caps)
        assert rec[R] <= cap
        sizes.append(int(rec[R]))
        nkeys += int(keys[R])
    kc = L.mfsdbg_skm_key_capacity(nkeys)
    ka, kb = ctx.dev_alloc(kc * 8 + 256), ctx.dev_alloc(kc * 8 + 256)
    for _ in range(2):
        e = ctx.count_skm(buf, [s * cap for s in range(W)], sizes, nkeys, a.k, 2, ka, kb, kc)
        prof = ctx.last_profile()
    print(json.dumps(dict(rank=R, world=W, records=sum(sizes), keys=nkeys, edges=int(e.n), stages_ms={k: round(v, 3) for k, v in prof.items()},
                          total_ms=round(sum(prof.values()), 3))))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Stage times of the super-k-mer exchange (csrc/skm.cu) at full size on ONE GPU: the destinations are local buffers, so
NVLink is not in the picture -- this is the cost of the kernels themselves.

  sender   : skm_scatter of the 5 Gbp read set into n_dst regions (what every GPU runs at N = n_dst)
  receiver : count_skm over ALL the records in one region set (n_dst = 1): the key count a GPU receives in weak scaling,
             compared with the plain single-GPU count of the same reads (same edges expected)."""
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
    ap.add_argument("--dst", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from mitoflex_b200 import lib
    L = lib.load()
    ctx = lib.Context(0)
    ctx.set_profiling(True)
    reads = ctx.synth(n_pairs=a.pairs, seed=1001)
    out = dict(pairs=a.pairs, k=a.k, bases=int(reads.n_bases))
    # ---- sender, n_dst destinations
    samp, _ = ctx.skm_scatter(reads, a.k, a.dst, stride=64)
    out["sample_ms"] = ctx.last_profile().get("skm_sample")
    full_est = samp * 64
    caps = (full_est * 1.05).astype(np.int64) + 65536
    bufs = [ctx.dev_alloc(int(c) * 8) for c in caps]
    ms = []
    for _ in range(a.reps):
        rec, keys = ctx.skm_scatter(reads, a.k, a.dst, np.array(bufs, np.uint64), caps)
        ms.append(ctx.last_profile().get("skm_scatter"))
    assert (rec <= caps).all(), (rec, caps)
    out.update(sender_ms=ms, records=int(rec.sum()), keys=int(keys.sum()), keys_per_record=float(keys.sum() / rec.sum()),
               bytes_per_key=float(8 * rec.sum() / keys.sum()), balance_max_over_mean=float(keys.max() / keys.mean()),
               est_err=float(np.abs(full_est - rec).max() / rec.mean()))
    for b in bufs:
        ctx.dev_free(b)
    # ---- receiver: all records on one GPU
    rec1, keys1 = ctx.skm_scatter(reads, a.k, 1)
    cap1 = int(rec1[0]) + 16
    buf = ctx.dev_alloc(cap1 * 8)
    ctx.skm_scatter(reads, a.k, 1, np.array([buf], np.uint64), np.array([cap1]))
    kc = L.mfsdbg_skm_key_capacity(int(keys1[0]))
    ka, kb = ctx.dev_alloc(kc * 8 + 256), ctx.dev_alloc(kc * 8 + 256)
    prof = None
    for _ in range(a.reps):
        e = ctx.count_skm(buf, [0], [int(rec1[0])], int(keys1[0]), a.k, 2, ka, kb, kc)
        prof = ctx.last_profile()
    n_skm = e.n
    def checksum(ed):   # order-independent (the skm count returns its edges in mixed-key order)
        w = ctx.d2h(ed.s.edges, ed.n * ed.s.words_per_edge * 4, np.uint32).reshape(-1, ed.s.words_per_edge).astype(np.uint64)
        x = (w[:, 0] << np.uint64(32)) | w[:, 1]
        with np.errstate(over="ignore"):
            x = x * np.uint64(0x9E3779B97F4A7C15)
            x ^= x >> np.uint64(31)
            return int(x.sum(dtype=np.uint64))
    chk = checksum(e)
    out.update(receiver_stages_ms={k: round(v, 3) for k, v in prof.items()}, receiver_ms=round(sum(prof.values()), 3), edges=int(n_skm))
    for p in (buf, ka, kb):
        ctx.dev_free(p)
    e0 = ctx.count(reads, a.k, 2)
    p0 = ctx.last_profile()
    chk0 = checksum(e0)
    out.update(plain_count_stages_ms={k: round(v, 3) for k, v in p0.items()}, plain_count_ms=round(sum(p0.values()), 3),
               plain_edges=int(e0.n), same_edges=bool(e0.n == n_skm and chk == chk0))
    line = json.dumps(out)
    print(line)
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "skm_bench.json"), "a") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()

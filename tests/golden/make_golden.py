"""Generates the fixtures in this directory.  Run from the repo root:  python tests/golden/make_golden.py

PARITY UNPINNED: neither a megahit_core binary nor megahit's source exists in /root/reference or in this image, and the
reference has no tests / golden vectors for the sDBG path.  These vectors are therefore produced by
the in-repo CPU oracle (oracle/mh_oracle.c, a restatement of megahit v1.2.9 from recollection) on seeded inputs; they pin
the oracle against regressions and give the GPU path a fixed target, they do not prove agreement with upstream megahit.
The hand-derived known-answer vectors live in tests/test_oracle.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import make_reads  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = [("k21_m2", 21, 2, 1500, 3000), ("k31_m1", 31, 1, 800, 2000), ("k47_m2", 47, 2, 800, 2000), ("k141_m2", 141, 2, 500, 1500)]


def main():
    for name, k, m, n_reads, glen in CASES:
        bases, starts = make_reads(20261017 + k, n_reads, k, genome_len=glen, max_len=max(150, k + 30), err=0.01)
        r = oracle.Reads(bases, starts)
        e = oracle.count(r, k, m)
        s = oracle.Seqs()
        s.add_edges(e)
        g = oracle.seq2sdbg(s, k)
        g1 = oracle.read2sdbg(r, k, m)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), k=k, m=m, bases=bases, starts=starts, edges=e.data,
                            counting_nonzero=np.stack([np.nonzero(e.counting)[0], e.counting[np.nonzero(e.counting)[0]]]),
                            w=g.w, last=g.last, tip=g.tip, mul=g.mul, tip_labels=g.tip_labels, r2s_tip_labels=g1.tip_labels)
        print(name, "reads", n_reads, "edges", e.n, "items", g.n, "tips", int(g.tip.sum()))


if __name__ == "__main__":
    main()

import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mitoflex_b200 import lib
from oracle import oracle
k, m = 21, 2
c = lib.Context(0)
reads = c.synth(n_pairs=2_000_000, seed=1001)
bases, starts = c.download_reads(reads)
g_orc = oracle.read2sdbg(oracle.Reads(bases, starts), k, m, threads=os.cpu_count())
for filt in ("1", "0"):
    os.environ["MFSDBG_ITEM_FILTER"] = filt
    g = c.read2sdbg(reads, k, m).to_numpy()
    l1, l2 = g["tip_labels"], g_orc.tip_labels
    print("filter", filt, "shapes", l1.shape, l2.shape, "w eq", np.array_equal(g["w"], g_orc.w), "tip eq", np.array_equal(g["tip"], g_orc.tip))
    if l1.shape == l2.shape:
        bad = np.nonzero((l1 != l2).any(axis=1))[0]
        print(" differ", len(bad))
        tipidx = np.nonzero(g["tip"])[0]
        for i in bad[:12]:
            it = tipidx[i]
            print("  tip", i, "item", it, "gpu", [hex(x) for x in l1[i]], "orc", [hex(x) for x in l2[i]], "w", g["w"][it], "nbr w", g["w"][max(0, it - 2):it + 3].tolist(),
                  "nbr tip", g["tip"][max(0, it - 2):it + 3].tolist())
        # are the labels a permutation of each other?
        a = np.sort(l1.view([("a", l1.dtype), ("b", l1.dtype)]).ravel(), order=("a", "b"))
        b = np.sort(l2.view([("a", l2.dtype), ("b", l2.dtype)]).ravel(), order=("a", "b"))
        print(" same multiset:", np.array_equal(a, b))

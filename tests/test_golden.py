"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py with the oracle -- parity unpinned, see the
script's header).  CPU: the oracle still reproduces them.  GPU: libmfsdbg reproduces them through the C ABI."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _counting(z):
    c = np.zeros(65536, np.int64)
    c[z["counting_nonzero"][0]] = z["counting_nonzero"][1]
    return c


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_fixture(oracle, path):
    z = np.load(path)
    k, m = int(z["k"]), int(z["m"])
    r = oracle.Reads(z["bases"], z["starts"])
    e = oracle.count(r, k, m, threads=2)
    assert np.array_equal(e.data, z["edges"]) and np.array_equal(e.counting, _counting(z))
    g = oracle.read2sdbg(r, k, m, threads=2)
    for f in ("w", "last", "tip", "mul"):
        assert np.array_equal(getattr(g, f), z[f]), f
    assert np.array_equal(g.tip_labels, z["r2s_tip_labels"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_gpu_reproduces_fixture(path):
    from mitoflex_b200 import lib
    z = np.load(path)
    k, m = int(z["k"]), int(z["m"])
    ctx = lib.Context(0)
    reads = ctx.upload_reads(z["bases"], z["starts"])
    e = ctx.count(reads, k, m, want_counting=True)
    assert np.array_equal(e.to_numpy(), z["edges"]) and np.array_equal(e.counting, _counting(z))
    g = ctx.seq2sdbg(e, k).to_numpy()
    for f in ("w", "last", "tip", "mul", "tip_labels"):
        assert np.array_equal(g[f], z[f]), f
    g1 = ctx.read2sdbg(reads, k, m).to_numpy()
    for f in ("w", "last", "tip", "mul"):
        assert np.array_equal(g1[f], z[f]), f
    assert np.array_equal(g1["tip_labels"], z["r2s_tip_labels"])
    ctx.close()

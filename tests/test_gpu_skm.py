"""Super-k-mer exchange of the multi-GPU count (csrc/skm.cu) on ONE GPU: the destinations are separate buffers of the same
device ("virtual ranks"), so the sender, the record format and the receiver's record-fed level 1 are checked without NVLink.
Union of the virtual ranks' edge sets == the oracle's count of the whole read set, bit for bit; every canonical key must
sit on exactly one rank."""
import numpy as np
import pytest

from gpu_common import make_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from mitoflex_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


def _skm_count(ctx, reads, k, m, n_dst, slack=1.0):
    """-> list of edge arrays (one per virtual rank), records per rank, keys per rank"""
    from mitoflex_b200 import lib
    L = lib.load()
    rec0, keys0 = ctx.skm_scatter(reads, k, n_dst)                  # count only, every tile
    caps = np.maximum((rec0 * slack).astype(np.int64), 1)
    bufs = [ctx.dev_alloc(int(c) * 8 + 64) for c in caps]
    try:
        rec, keys = ctx.skm_scatter(reads, k, n_dst, np.array(bufs, np.uint64), caps)
        assert np.array_equal(rec, rec0) and np.array_equal(keys, keys0)
        out = []
        for d in range(n_dst):
            if slack < 1.0 and rec[d] > caps[d]:
                out.append(None)
                continue
            cap = L.mfsdbg_skm_key_capacity(int(keys[d]))
            ka, kb = ctx.dev_alloc(cap * 8 + 256), ctx.dev_alloc(cap * 8 + 256)
            try:
                e = ctx.count_skm(bufs[d], [0], [int(rec[d])], int(keys[d]), k, m, ka, kb, cap)
                out.append(e.to_numpy().copy())
            finally:
                ctx.dev_free(ka)
                ctx.dev_free(kb)
        return out, rec, keys
    finally:
        for b in bufs:
            ctx.dev_free(b)


def _merge(parts):
    e = np.concatenate([p for p in parts if len(p)], axis=0) if any(len(p) for p in parts) else parts[0]
    order = np.lexsort(tuple(e[:, c] for c in range(e.shape[1] - 1, -1, -1)))
    return e[order]


@pytest.mark.parametrize("k,m,n_dst", [(21, 2, 1), (21, 2, 2), (21, 1, 8), (21, 3, 3), (16, 2, 4), (19, 1, 5), (20, 2, 8),
                                       (24, 2, 8), (26, 1, 16), (26, 2, 2)])
def test_skm_union_matches_oracle(ctx, oracle, k, m, n_dst):
    bases, starts = make_reads(900 + 13 * k + n_dst, 30000, k, genome_len=80000, max_len=150, err=0.01)
    reads = ctx.upload_reads(bases, starts)
    parts, rec, keys = _skm_count(ctx, reads, k, m, n_dst)
    want = oracle.count(oracle.Reads(bases, starts), k, m, threads=8)
    n_pos = int(np.maximum(np.diff(starts) - k, 0).sum())
    assert int(keys.sum()) == n_pos                       # every (k+1)-mer travels exactly once
    assert int(rec.sum()) <= n_pos
    got = _merge(parts)
    assert got.shape == want.data.shape, (got.shape, want.data.shape)
    assert np.array_equal(got, want.data)
    if n_dst >= 2 and m == 1:
        # balance: no virtual rank holds more than 1.25x its share of the keys on a random genome
        assert keys.max() <= 1.25 * keys.mean() + 1000, keys
    if n_dst == 8 and k == 21:
        # the point of the exercise: several (k+1)-mers per 8-byte record
        assert keys.sum() / rec.sum() > 3.0, (keys.sum(), rec.sum())


def test_skm_edge_cases(ctx, oracle):
    k = 21
    for seqs in ([], [np.zeros(5, np.uint8)], [np.arange(40, dtype=np.uint8) & 3], [np.zeros(300, np.uint8)],
                 [np.arange(22, dtype=np.uint8) & 3, (np.arange(23, dtype=np.uint8) * 3) & 3]):
        starts = np.zeros(len(seqs) + 1, np.int64)
        starts[1:] = np.cumsum([len(s) for s in seqs])
        bases = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, np.uint8)
        reads = ctx.upload_reads(bases, starts)
        parts, rec, keys = _skm_count(ctx, reads, k, 1, 4)
        want = oracle.count(oracle.Reads(bases, starts), k, 1)
        assert np.array_equal(_merge(parts), want.data) if want.data.shape[0] else sum(len(p) for p in parts) == 0


def test_skm_overflow_is_reported(ctx):
    """regions too small: the dropped runs show up as counts beyond the capacity (the driver retries with more room)"""
    k = 21
    bases, starts = make_reads(77, 20000, k, genome_len=50000, max_len=150, err=0.01)
    reads = ctx.upload_reads(bases, starts)
    rec0, _ = ctx.skm_scatter(reads, k, 4)
    caps = (rec0 // 2).astype(np.int64)
    bufs = [ctx.dev_alloc(int(c) * 8 + 64) for c in caps]
    try:
        rec, _ = ctx.skm_scatter(reads, k, 4, np.array(bufs, np.uint64), caps)
        assert np.array_equal(rec, rec0) and (rec > caps).all()
    finally:
        for b in bufs:
            ctx.dev_free(b)


def test_skm_sample_estimates_the_counts(ctx):
    k = 21
    bases, starts = make_reads(78, 200000, k, genome_len=500000, max_len=150, err=0.005)
    reads = ctx.upload_reads(bases, starts)
    full, _ = ctx.skm_scatter(reads, k, 8)
    samp, _ = ctx.skm_scatter(reads, k, 8, stride=16)
    est = samp * (full.sum() / max(samp.sum(), 1))
    assert np.all(np.abs(est - full) <= 0.05 * full + 2000), (est, full)


def test_skm_rejects_unsupported_k(ctx):
    from mitoflex_b200 import lib
    assert lib.load().mfsdbg_skm_supported(21) == 1 and lib.load().mfsdbg_skm_supported(31) == 0
    bases, starts = make_reads(5, 100, 31, genome_len=2000)
    with pytest.raises(lib.MfsdbgError):
        ctx.skm_scatter(ctx.upload_reads(bases, starts), 31, 2)


@pytest.mark.parametrize("cmax", [4, 2, 6])
def test_skm_short_records(ctx, oracle, monkeypatch, cmax):
    """MFSDBG_SKM_CMAX: shorter records (<= 4 keys take the receiver's 4-slot kernel); same edges"""
    monkeypatch.setenv("MFSDBG_SKM_CMAX", str(cmax))
    k, m, n_dst = 21, 2, 8
    bases, starts = make_reads(4100 + cmax, 30000, k, genome_len=80000, max_len=150, err=0.01)
    parts, rec, keys = _skm_count(ctx, ctx.upload_reads(bases, starts), k, m, n_dst)
    want = oracle.count(oracle.Reads(bases, starts), k, m, threads=8)
    assert np.array_equal(_merge(parts), want.data)
    assert keys.sum() / rec.sum() <= cmax

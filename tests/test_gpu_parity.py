"""GPU parity: libmfsdbg (through its C ABI) against the CPU oracle, bit-exact (integer / byte work).
Every test here needs a B200; run with `pytest -m gpu`."""
import os

import numpy as np
import pytest

from gpu_common import assert_edges_equal, assert_sdbg_equal, make_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from mitoflex_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


def _orc_reads(oracle, bases, starts):
    return oracle.Reads(bases, starts)


# k values cover every key width class: 1 word (k<=15), 2 (<=31), 3, 4, 5 (k=79), 6 (k=90), 7 (k=99), 8 (k=119), 9 (k=141)
COUNT_CASES = [(9, 1), (9, 2), (15, 2), (16, 1), (21, 1), (21, 2), (21, 3), (31, 2), (32, 2), (47, 1), (59, 2), (63, 2),
               (79, 2), (90, 2), (99, 2), (119, 2), (141, 2)]


@pytest.mark.parametrize("k,m", COUNT_CASES)
def test_count_parity(ctx, oracle, k, m):
    bases, starts = make_reads(100 + k * 7 + m, 3000, k, genome_len=6000, max_len=max(150, k + 40))
    e_gpu = ctx.count(ctx.upload_reads(bases, starts), k, m, want_counting=True)
    e_orc = oracle.count(_orc_reads(oracle, bases, starts), k, m, threads=4)
    assert_edges_equal(e_gpu, e_orc)
    assert np.array_equal(e_gpu.counting, e_orc.counting)


@pytest.mark.parametrize("k,m", [(21, 2), (31, 2), (59, 2)])
def test_count_parity_medium(ctx, oracle, k, m):
    """enough keys for two partition levels (>> one bucket)."""
    bases, starts = make_reads(7 + k, 120000, k, genome_len=400000, max_len=150, err=0.005)
    e_gpu = ctx.count(ctx.upload_reads(bases, starts), k, m)
    e_orc = oracle.count(_orc_reads(oracle, bases, starts), k, m, threads=8)
    assert_edges_equal(e_gpu, e_orc)


def test_count_edge_cases(ctx, oracle):
    # empty library, only short reads, a single read
    for seqs in ([], [np.zeros(5, np.uint8)], [np.arange(40, dtype=np.uint8) & 3]):
        starts = np.zeros(len(seqs) + 1, np.int64)
        starts[1:] = np.cumsum([len(s) for s in seqs])
        bases = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, np.uint8)
        e_gpu = ctx.count(ctx.upload_reads(bases, starts), 21, 1)
        e_orc = oracle.count(_orc_reads(oracle, bases, starts), 21, 1)
        assert_edges_equal(e_gpu, e_orc)


def test_count_skewed_buckets(ctx, oracle):
    """one read repeated 70000 times (multiplicity cap 65535, buckets far larger than shared memory -> fallback path),
    poly-A reads, and a low-complexity family sharing long prefixes."""
    k = 21
    bases, starts = make_reads(5, 2000, k, dup_boost=70000)
    rng = np.random.default_rng(9)
    extra = [np.zeros(150, np.uint8) for _ in range(300)]
    stem = rng.integers(0, 4, 60, dtype=np.uint8)
    for _ in range(20000):
        tail = rng.integers(0, 4, 30, dtype=np.uint8)
        extra.append(np.concatenate([stem, tail]))
    seqs = [bases[starts[i]:starts[i + 1]] for i in range(len(starts) - 1)] + extra
    starts2 = np.zeros(len(seqs) + 1, np.int64)
    starts2[1:] = np.cumsum([len(s) for s in seqs])
    bases2 = np.concatenate(seqs).astype(np.uint8)
    for m in (1, 2):
        e_gpu = ctx.count(ctx.upload_reads(bases2, starts2), k, m, want_counting=True)
        e_orc = oracle.count(_orc_reads(oracle, bases2, starts2), k, m, threads=8)
        assert_edges_equal(e_gpu, e_orc)
        assert np.array_equal(e_gpu.counting, e_orc.counting)


@pytest.mark.parametrize("k,m", [(9, 1), (13, 2), (15, 2), (16, 2), (21, 2), (24, 1), (31, 2), (32, 2), (39, 2), (40, 1), (47, 2), (48, 2),
                                 (55, 2), (56, 2), (63, 2), (64, 2), (79, 2), (141, 2)])
def test_seq2sdbg_parity(ctx, oracle, k, m):
    bases, starts = make_reads(300 + k, 2500, k, genome_len=5000, max_len=max(150, k + 40))
    e_gpu = ctx.count(ctx.upload_reads(bases, starts), k, m)
    g_gpu = ctx.seq2sdbg(e_gpu, k)
    e_orc = oracle.count(_orc_reads(oracle, bases, starts), k, m, threads=4)
    s = oracle.Seqs()
    s.add_edges(e_orc)
    g_orc = oracle.seq2sdbg(s, k, threads=4)
    assert_sdbg_equal(g_gpu, g_orc)


@pytest.mark.parametrize("k,m", [(9, 2), (15, 2), (21, 1), (21, 2), (31, 2), (47, 2), (141, 2)])
def test_read2sdbg_parity(ctx, oracle, k, m):
    bases, starts = make_reads(500 + k, 2500, k, genome_len=5000, max_len=max(150, k + 40))
    g_gpu = ctx.read2sdbg(ctx.upload_reads(bases, starts), k, m)
    g_orc = oracle.read2sdbg(_orc_reads(oracle, bases, starts), k, m, threads=4)
    assert_sdbg_equal(g_gpu, g_orc)


def test_sdbg_medium(ctx, oracle):
    k, m = 21, 2
    bases, starts = make_reads(77, 100000, k, genome_len=300000, max_len=150, err=0.005)
    g_gpu = ctx.read2sdbg(ctx.upload_reads(bases, starts), k, m)
    g_orc = oracle.read2sdbg(_orc_reads(oracle, bases, starts), k, m, threads=8)
    assert_sdbg_equal(g_gpu, g_orc)


def test_high_multiplicity_sdbg(ctx, oracle):
    """multiplicities around the 254/255 inline limit and the 65535 cap."""
    k = 21
    rng = np.random.default_rng(3)
    base_reads = [rng.integers(0, 4, 60, dtype=np.uint8) for _ in range(4)]
    seqs = []
    for r, copies in zip(base_reads, (254, 255, 256, 66000)):
        seqs += [r] * copies
    starts = np.zeros(len(seqs) + 1, np.int64)
    starts[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.concatenate(seqs).astype(np.uint8)
    g_gpu = ctx.read2sdbg(ctx.upload_reads(bases, starts), k, 2)
    g_orc = oracle.read2sdbg(_orc_reads(oracle, bases, starts), k, 2, threads=4)
    assert_sdbg_equal(g_gpu, g_orc)
    assert g_orc.n_large > 0


def test_synth_generator_and_properties(ctx, oracle):
    """the in-HBM generator feeds both sides: download its reads, run the oracle on them."""
    reads = ctx.synth(n_pairs=20000, nuclear_len=200000, mito_len=16500, mito_fraction=0.05, error_rate=0.005, seed=1001)
    bases, starts = ctx.download_reads(reads)
    lens = np.diff(starts)
    assert reads.n_reads == 40000 and lens.max() == 150 and lens.min() >= 0 and (lens < 150).any()
    e_gpu = ctx.count(reads, 21, 2, want_counting=True)
    e_orc = oracle.count(oracle.Reads(bases, starts), 21, 2, threads=8)
    assert_edges_equal(e_gpu, e_orc)
    # size-independent property: the multiplicity histogram accounts for every (k+1)-mer occurrence below the cap
    n_keys = int(np.maximum(lens - 21, 0).sum())
    assert e_gpu.s.n_keys == n_keys
    c = e_gpu.counting
    assert (c[:65535] * np.arange(65535)).sum() + c[65535] * 65535 <= n_keys
    assert (c[:65535] * np.arange(65535)).sum() == n_keys - 0 or c[65535] > 0
    g_gpu = ctx.read2sdbg(reads, 21, 2)
    g_orc = oracle.read2sdbg(oracle.Reads(bases, starts), 21, 2, threads=8)
    assert_sdbg_equal(g_gpu, g_orc)


def test_count_out_of_core_rounds(oracle):
    """a small memory budget forces several rounds over prefix-bin ranges (the 30 Gbp-on-one-GPU path, SURVEY config 4)."""
    from mitoflex_b200 import lib
    c = lib.Context(0)
    try:
        k, m = 21, 2
        bases, starts = make_reads(99, 150000, k, genome_len=500000, max_len=150, err=0.005)
        n_keys = int(np.maximum(np.diff(starts) - k, 0).sum())
        c.set_mem_limit(200 << 20)          # ~9 M keys at 20 B each do not fit next to the tables: at least two rounds
        e_gpu = c.count(c.upload_reads(bases, starts), k, m, want_counting=True)
        e_orc = oracle.count(oracle.Reads(bases, starts), k, m, threads=8)
        assert e_gpu.s.n_keys == n_keys
        assert_edges_equal(e_gpu, e_orc)
        assert np.array_equal(e_gpu.counting, e_orc.counting)
    finally:
        c.close()


def test_count_crowded_buckets_multipass(ctx, oracle, monkeypatch):
    """buckets sized far beyond the shared table (every one of them bails) take the multi-pass kernel; shallow, error-rich
    reads make nearly every key distinct."""
    monkeypatch.setenv("MFSDBG_STREAM_LOAD_PCT", "400")
    k, m = 21, 1
    bases, starts = make_reads(123, 60000, k, genome_len=3000000, max_len=150, err=0.02)
    ctx.set_profiling(True)
    try:
        e_gpu = ctx.count(ctx.upload_reads(bases, starts), k, m, want_counting=True)
        prof = ctx.last_profile()
    finally:
        ctx.set_profiling(False)
    e_orc = oracle.count(_orc_reads(oracle, bases, starts), k, m, threads=8)
    assert_edges_equal(e_gpu, e_orc)
    assert np.array_equal(e_gpu.counting, e_orc.counting)
    assert "local_count_multipass" in prof, prof


@pytest.mark.parametrize("pipelined", [False, True])
def test_host_read2sdbg_matches_oracle(ctx, oracle, monkeypatch, pipelined):
    """the end-to-end entry point (host buffers in, host buffers out); with MFSDBG_H2D_MIN_BASES=0 the transfer is cut into
    chunks and the reads-fed partition level runs chunk by chunk behind it (several level-1 chunks per segment).  The input
    includes 400 copies of one read so that multiplicities beyond 254 travel through the side list."""
    import ctypes
    from mitoflex_b200 import lib
    if pipelined:
        monkeypatch.setenv("MFSDBG_H2D_MIN_BASES", "0")
    k, m = 21, 2
    bases, starts = make_reads(31, 30000, k, genome_len=120000, max_len=150, err=0.005, dup_boost=400)
    words, st = lib.pack_reads(bases, starts)
    words = np.ascontiguousarray(words)
    st = np.ascontiguousarray(st)
    out = ctx.host_read2sdbg(words.ctypes.data, st.ctypes.data, len(st) - 1, int(st[-1]), k, m)
    g = oracle.read2sdbg(_orc_reads(oracle, bases, starts), k, m, threads=8)
    assert out.n_items == g.n
    assert out.n_tips * out.words_per_tip == np.asarray(g.tip_labels).size
    # records come back as megahit's 16-bit packed items; multiplicities beyond 254 in the side list (ascending item index)
    rec = np.ctypeslib.as_array(ctypes.cast(out.rec, ctypes.POINTER(ctypes.c_uint16)), shape=(out.n_items,)).copy()
    assert np.array_equal(rec & 0xF, g.w)
    assert np.array_equal((rec >> 4) & 1, g.last)
    assert np.array_equal((rec >> 5) & 1, g.tip)
    mul = (rec >> 8).astype(np.int64)
    assert np.array_equal(mul, np.minimum(g.mul, 255))
    big = np.nonzero(g.mul > 254)[0]
    assert out.n_large == len(big)
    if out.n_large:
        idx = np.ctypeslib.as_array(ctypes.cast(out.large_index, ctypes.POINTER(ctypes.c_int64)), shape=(out.n_large,)).copy()
        lm = np.ctypeslib.as_array(ctypes.cast(out.large_mult, ctypes.POINTER(ctypes.c_uint16)), shape=(out.n_large,)).copy()
        assert np.array_equal(idx, big) and np.array_equal(lm, g.mul[big])
    if out.n_tips:
        lab = np.ctypeslib.as_array(ctypes.cast(out.tip_labels, ctypes.POINTER(ctypes.c_uint32)),
                                    shape=(out.n_tips * out.words_per_tip,)).copy()
        assert np.array_equal(lab, np.asarray(g.tip_labels).ravel())
    assert out.n_large == g.n_large


@pytest.mark.parametrize("k", [21, 47])
def test_count_general_and_fallback_paths(ctx, oracle, monkeypatch, k):
    """with the streamed kernels switched off every bucket takes the general kernel, and the skewed input (one read 70 000
    times, 20 000 distinct keys behind one stem) drives it through the chunk / merge / sorted-run fallbacks."""
    monkeypatch.setenv("MFSDBG_COUNT_STREAM", "0")
    monkeypatch.setenv("MFSDBG_COUNT_STREAM_W", "0")
    bases, starts = make_reads(6, 1500, k, dup_boost=70000, max_len=max(150, k + 40))
    rng = np.random.default_rng(10)
    stem = rng.integers(0, 4, k + 30, dtype=np.uint8)
    extra = [np.concatenate([stem, rng.integers(0, 4, 40, dtype=np.uint8)]) for _ in range(20000)]
    seqs = [bases[starts[i]:starts[i + 1]] for i in range(len(starts) - 1)] + extra
    starts2 = np.zeros(len(seqs) + 1, np.int64)
    starts2[1:] = np.cumsum([len(s) for s in seqs])
    bases2 = np.concatenate(seqs).astype(np.uint8)
    ctx.set_profiling(True)
    try:
        e_gpu = ctx.count(ctx.upload_reads(bases2, starts2), k, 2, want_counting=True)
        prof = ctx.last_profile()
    finally:
        ctx.set_profiling(False)
    e_orc = oracle.count(_orc_reads(oracle, bases2, starts2), k, 2, threads=8)
    assert_edges_equal(e_gpu, e_orc)
    assert np.array_equal(e_gpu.counting, e_orc.counting)
    assert "oversized" in prof, prof


def test_full_size_properties():
    """BASELINE configs[1] read set (16 666 667 x PE150, 5 Gbp): too large for the oracle, so size-independent properties:
    the multiplicity histogram accounts for every (k+1)-mer occurrence, the edge stream is strictly increasing (sorted and
    distinct), edge multiplicities add up to the solid part of the histogram, the graph's bucket statistics add up to its
    totals, and read2sdbg equals count followed by seq2sdbg."""
    from mitoflex_b200 import lib
    c = lib.Context(0)
    try:
        k, m = 21, 2
        reads = c.synth(n_pairs=16_666_667, seed=1002)
        e = c.count(reads, k, m, want_counting=True)
        cnt = e.counting.astype(np.int64)
        idx = np.arange(65536, dtype=np.int64)
        assert cnt[65535] == 0                                   # nothing reaches the multiplicity cap on this sample
        assert int((cnt * idx).sum()) == e.s.n_keys              # every key occurrence is in exactly one run
        assert int(cnt[m:].sum()) == e.n                         # one edge per solid distinct key
        ed = e.to_numpy()                                        # [n, 2] words: key (44 bits) | multiplicity (16 bits)
        key = (ed[:, 0].astype(np.uint64) << np.uint64(32)) | (ed[:, 1].astype(np.uint64) & np.uint64(0xFFF00000))
        assert bool((key[1:] > key[:-1]).all())                  # globally sorted, no duplicates
        mult = (ed[:, 1] & 0xFFFF).astype(np.int64)
        assert int(mult.min()) >= m
        assert int(mult.sum()) == int((cnt[m:] * idx[m:]).sum())
        assert int(e.bucket_counts().sum()) == e.n
        del ed, key, mult
        g1 = c.seq2sdbg(e, k)
        n1, st1 = g1.n, g1.bucket_stats().sum(axis=0)
        rec1 = g1.to_numpy()
        g2 = c.read2sdbg(reads, k, m)
        rec2 = g2.to_numpy()
        assert g2.n == n1 and n1 > 2 * e.n                      # two real items per edge plus the surviving dummies
        assert int(st1[0]) == n1 and int(st1[1]) == int(rec1["tip"].sum()) and int(st1[2]) == rec1["n_large"]
        for f in ("w", "last", "tip", "mul"):
            assert np.array_equal(rec1[f], rec2[f]), f
        assert int(rec2["last"].sum()) > 0
    finally:
        c.close()


@pytest.mark.parametrize("stride,slack", [(2, 10), (64, 10), (4, -60)])
def test_count_sampled_histogram(ctx, oracle, monkeypatch, stride, slack):
    """level 1 sized from a strided sample of the tiles; with a negative slack the bins outgrow their regions (runs that
    would cross a limit are dropped) and the exact pass takes over; equal results either way."""
    monkeypatch.setenv("MFSDBG_SAMPLED_MIN_TILES", "0")
    monkeypatch.setenv("MFSDBG_SAMPLED_STRIDE", str(stride))
    monkeypatch.setenv("MFSDBG_SAMPLED_SLACK_PCT", str(slack))
    k, m = 21, 2
    bases, starts = make_reads(77, 40000, k, genome_len=200000, max_len=150, err=0.005)
    ctx.set_profiling(True)
    try:
        e_gpu = ctx.count(ctx.upload_reads(bases, starts), k, m, want_counting=True)
        prof = ctx.last_profile()
    finally:
        ctx.set_profiling(False)
    assert ("sampled_overflow" in prof) == (slack < 0), prof
    e_orc = oracle.count(_orc_reads(oracle, bases, starts), k, m, threads=8)
    assert e_gpu.s.n_keys == int(np.maximum(np.diff(starts) - k, 0).sum())
    assert_edges_equal(e_gpu, e_orc)
    assert np.array_equal(e_gpu.counting, e_orc.counting)


@pytest.mark.parametrize("k,m,round_items", [(9, 1, 700), (15, 2, 3000), (21, 1, 2000), (21, 2, 50000), (31, 2, 3000), (47, 2, 4000),
                                             (79, 2, 2500), (141, 2, 1500)])
def test_sdbg_memory_bounded_rounds(ctx, oracle, monkeypatch, k, m, round_items):
    """the sdbg stage in rounds over level-1 bin ranges (what an item set larger than the HBM budget takes: SURVEY config 5,
    -m 1 on error-rich reads); a tiny round size forces many rounds, the appended outputs must equal the one-pass graph."""
    monkeypatch.setenv("MFSDBG_SDBG_ROUND_ITEMS", str(round_items))
    bases, starts = make_reads(900 + k, 4000, k, genome_len=8000, max_len=max(150, k + 40), err=0.02)
    ctx.set_profiling(True)
    try:
        g_gpu = ctx.read2sdbg(ctx.upload_reads(bases, starts), k, m)
        prof = ctx.last_profile()
    finally:
        ctx.set_profiling(False)
    g_orc = oracle.read2sdbg(_orc_reads(oracle, bases, starts), k, m, threads=4)
    assert_sdbg_equal(g_gpu, g_orc)
    assert "items_hist" in prof, prof


def test_config1_full_size_vs_oracle(oracle):
    """BASELINE configs[0] as SURVEY.md 8d defines it (2 000 000 x PE150 pairs = 600 Mbp, seed 1001, 5 % mitogenome, 0.5 %
    errors): `count`, `seq2sdbg` and `read2sdbg` at k=21, -m 2 and -m 3, default planner, no environment overrides, every
    output bit-exact against the oracle.  The planner takes its full-size branches here (sampled level-1 histogram, range
    partition digit, distinct-ratio probe), which the small tests only reach through overrides."""
    from mitoflex_b200 import lib
    c = lib.Context(0)
    try:
        k = 21
        reads = c.synth(n_pairs=2_000_000, seed=1001)
        bases, starts = c.download_reads(reads)
        assert reads.n_reads == 4_000_000 and 590e6 < reads.n_bases < 600e6
        o_reads = oracle.Reads(bases, starts)
        threads = os.cpu_count() or 8
        for m in (2, 3):
            e_gpu = c.count(reads, k, m, want_counting=True)
            e_orc = oracle.count(o_reads, k, m, threads=threads)
            assert_edges_equal(e_gpu, e_orc)
            assert np.array_equal(e_gpu.counting, e_orc.counting)
            s = oracle.Seqs()
            s.add_edges(e_orc)
            g_orc = oracle.seq2sdbg(s, k, threads=threads)
            assert_sdbg_equal(c.seq2sdbg(e_gpu, k), g_orc)
            g_gpu = c.read2sdbg(reads, k, m)
            if m == 2:   # the oracle's one-pass route sorts an item per solid read position (minutes): once is enough
                assert_sdbg_equal(g_gpu, oracle.read2sdbg(o_reads, k, m, threads=threads))
            else:
                # against count + seq2sdbg (proven equal to the one-pass route on the oracle): same graph; the tip labels of the
                # two routes differ only in the bits below the bases (stage-2 items carry flag | b there, seq2sdbg items the
                # multiplicity), so they are compared on the bases
                gn = g_gpu.to_numpy()
                for f in ("w", "last", "tip", "mul"):
                    assert np.array_equal(gn[f], getattr(g_orc, f)), f
                assert np.array_equal(gn["tip_labels"][:, 0], g_orc.tip_labels[:, 0])
                assert np.array_equal(gn["tip_labels"][:, 1] & 0xfff00000, g_orc.tip_labels[:, 1] & 0xfff00000)
            assert g_gpu.n > 2 * e_gpu.n > 0
    finally:
        c.close()


def test_synth_port_matches_device(ctx):
    """oracle/synth_np.py (what the CPU arm of bench.py generates its sample with, so that it never loads libmfsdbg.so) is
    the same generator as csrc/synth.cu: same hash, geometry, error and N model.  libm rounding may move an insert size or
    a trim point on a vanishing fraction of reads, so the bar is: identical read count, > 99.9 % of reads byte-identical."""
    from oracle import synth_np
    for kw in (dict(n_pairs=30000, nuclear_len=300_000, seed=1001), dict(n_pairs=20000, nuclear_len=100_000, seed=5, error_rate=0.02)):
        reads = ctx.synth(**kw)
        gb, gs = ctx.download_reads(reads)
        nb, ns = synth_np.synth_reads(**kw, workers=1)
        assert len(ns) == len(gs)
        same_len = np.diff(ns) == np.diff(gs)
        assert same_len.mean() > 0.999
        if same_len.all():
            assert (nb == gb).mean() > 0.999
        idx = np.nonzero(same_len)[0][:5000]
        ok = sum(np.array_equal(nb[ns[i]:ns[i + 1]], gb[gs[i]:gs[i + 1]]) for i in idx)
        assert ok > 0.999 * len(idx)


@pytest.mark.parametrize("k,env", [(21, {}), (21, {"MFSDBG_COUNT_REL": "0"}), (23, {}), (24, {}), (31, {}),
                                   (31, {"MFSDBG_COUNT_REL": "2", "MFSDBG_L1_BITS": "11"}),
                                   (29, {"MFSDBG_COUNT_REL": "2", "MFSDBG_L1_BITS": "10"}),
                                   (21, {"MFSDBG_L1_BITS": "3"})])
def test_count_stream2_slot_layouts(ctx, oracle, monkeypatch, k, env):
    """k_count_stream2: slots that hold "key low bits | count" (REL, chosen from the plan), the full-key variant
    (MFSDBG_COUNT_REL=0 / few level-1 bits), and REL with a forced minimum number of level-2 ranges per segment (k >= 24 at full
    size; forced here with MFSDBG_COUNT_REL=2).  Deep coverage plus skew: 3000 copies of a read, a low-complexity family."""
    for name, v in env.items():
        monkeypatch.setenv(name, v)
    bases, starts = make_reads(8800 + k, 40000, k, genome_len=60000, max_len=150, err=0.01, dup_boost=3000)
    rng = np.random.default_rng(k)
    stem = rng.integers(0, 4, k + 20, dtype=np.uint8)
    extra = [np.concatenate([stem, rng.integers(0, 4, 40, dtype=np.uint8)]) for _ in range(3000)] + [np.zeros(150, np.uint8)] * 50
    seqs = [bases[starts[i]:starts[i + 1]] for i in range(len(starts) - 1)] + extra
    starts2 = np.zeros(len(seqs) + 1, np.int64)
    starts2[1:] = np.cumsum([len(s) for s in seqs])
    bases2 = np.concatenate(seqs).astype(np.uint8)
    for m in (1, 2):
        e_gpu = ctx.count(ctx.upload_reads(bases2, starts2), k, m, want_counting=True)
        e_orc = oracle.count(_orc_reads(oracle, bases2, starts2), k, m, threads=8)
        assert_edges_equal(e_gpu, e_orc)
        assert np.array_equal(e_gpu.counting, e_orc.counting)


_WIDE_PLANS = {
    "default": {},
    "walk": {"MFSDBG_READS_COMPACT": "0", "MFSDBG_L2_NBCAP_W": "1024", "MFSDBG_CW_CHUNK": "0"},
    # a batch per listing pass (k_reads_scatter_compact, the first version of the kernel, kept as the A/B partner)
    "compact": {"MFSDBG_READS_COMPACT": "2", "MFSDBG_READS_COMPACT_V": "1", "MFSDBG_L1_CAP_WC": "8"},
    # 1024 level-1 bins, 2048-bin level 2, the 2 x 512-key ring of k_count_stream_w<7 | 8>
    "compact_fine": {"MFSDBG_READS_COMPACT": "2", "MFSDBG_L1_CAP_WC": "10", "MFSDBG_L1_SEGK_WC": "1", "MFSDBG_L2_NBCAP_W": "2048",
                     "MFSDBG_CW_CHUNK": "1"},
    # batches cut loose from the listing passes (k_reads_scatter_compact2)
    "compact2": {"MFSDBG_READS_COMPACT": "2", "MFSDBG_READS_COMPACT_V": "2"},
    "compact2_fine": {"MFSDBG_READS_COMPACT": "2", "MFSDBG_READS_COMPACT_V": "2", "MFSDBG_L1_SEGK_WC": "1", "MFSDBG_CW_CHUNK": "1"},
}


@pytest.mark.parametrize("plan", list(_WIDE_PLANS))
@pytest.mark.parametrize("k", [39, 59, 79, 90, 99, 119, 141])
def test_count_wide_reads_scatter_variants(ctx, oracle, monkeypatch, k, plan):
    """Wide keys (3..9 words; k=39 only meets the ring variants of k_count_stream_w) through both reads-fed scatters: the one that walks every position (k_reads_scatter) and the one
    that walks the list of positions that start a key (k_reads_scatter_compact; at k=59 a tile lists more keys than one staging
    batch holds), plus the planner switches that ride on the latter.  150-base reads with a long tail of short ones."""
    for name, v in _WIDE_PLANS[plan].items():
        monkeypatch.setenv(name, v)
    bases, starts = make_reads(4100 + k, 60000, k, genome_len=150000, max_len=150, err=0.005)
    e_gpu = ctx.count(ctx.upload_reads(bases, starts), k, 2, want_counting=True)
    e_orc = oracle.count(_orc_reads(oracle, bases, starts), k, 2, threads=8)
    assert_edges_equal(e_gpu, e_orc)
    assert np.array_equal(e_gpu.counting, e_orc.counting)

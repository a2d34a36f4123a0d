"""Shared helpers for the GPU parity tests: seeded inputs + comparison against the CPU oracle."""
import numpy as np

COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def make_reads(seed, n_reads, k, genome_len=5000, max_len=150, err=0.01, dup_boost=0, short_frac=0.05):
    """Seeded reads over a random genome: both strands, substitution errors, variable lengths (some shorter than k+1,
    some empty), optional `dup_boost` extra copies of one read (skewed buckets)."""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len, dtype=np.uint8)
    seqs = []
    for _ in range(n_reads):
        if rng.random() < short_frac:
            L = int(rng.integers(0, k + 1))
        else:
            L = int(rng.integers(k + 1, max(k + 2, max_len + 1)))
        p = int(rng.integers(0, genome_len - L + 1))
        r = genome[p:p + L].copy()
        if rng.random() < 0.5:
            r = COMP[r[::-1]]
        e = rng.random(L) < err
        r[e] = (r[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) & 3
        seqs.append(r)
    if dup_boost:
        seqs += [seqs[0].copy() for _ in range(dup_boost)]
    starts = np.zeros(len(seqs) + 1, dtype=np.int64)
    starts[1:] = np.cumsum([len(s) for s in seqs])
    bases = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
    return bases.astype(np.uint8), starts


def assert_edges_equal(gpu_edges, orc_edges):
    g = gpu_edges.to_numpy()
    assert gpu_edges.s.words_per_edge == orc_edges.words
    assert g.shape == orc_edges.data.shape, (g.shape, orc_edges.data.shape)
    if not np.array_equal(g, orc_edges.data):
        bad = np.nonzero((g != orc_edges.data).any(axis=1))[0]
        raise AssertionError(f"{len(bad)} edge records differ, first at {bad[0]}: gpu={g[bad[0]]} oracle={orc_edges.data[bad[0]]}")
    assert np.array_equal(gpu_edges.bucket_counts(), orc_edges.bucket_counts)


def assert_sdbg_equal(gpu_sdbg, orc_sdbg):
    g = gpu_sdbg.to_numpy()
    assert gpu_sdbg.n == orc_sdbg.n, (gpu_sdbg.n, orc_sdbg.n)
    for f in ("w", "last", "tip", "mul"):
        a, b = g[f], getattr(orc_sdbg, f)
        if not np.array_equal(a, b):
            bad = np.nonzero(a != b)[0]
            raise AssertionError(f"sdbg field {f}: {len(bad)} items differ, first at {bad[0]}: gpu={a[bad[0]]} oracle={b[bad[0]]}")
    assert np.array_equal(g["tip_labels"], orc_sdbg.tip_labels)
    assert g["n_large"] == orc_sdbg.n_large
    assert np.array_equal(gpu_sdbg.bucket_stats()[:, 0], orc_sdbg.bucket_items)

"""numpy port of the in-HBM synthetic read generator (mitoflex_b200/csrc/synth.cu) -- TEST / BENCH
INFRASTRUCTURE, NOT PRODUCT CODE.

Lets the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) and the CPU tests build the benchmark's read set
without loading libmfsdbg.so.  Same hash, same per-read geometry, same error / N model; the only place the two can differ
is libm rounding inside Box-Muller / the geometric N gap (an insert size or a trim point off by one on a vanishing
fraction of reads) -- tests/test_gpu_parity.py::test_synth_port_matches_device measures it on the GPU box.
"""
import os

import numpy as np

U64 = np.uint64
_M = U64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    x = (x + U64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)
    return x ^ (x >> U64(31))


def _h3(seed, a, b):
    a = np.asarray(a, dtype=U64)
    b = np.asarray(b, dtype=U64)
    with np.errstate(over="ignore"):
        return _mix64(U64(seed) ^ _mix64(a * U64(0x632BE59BD9B4E019) + b * U64(0xD1342543DE82EF95) + U64(0x2545F4914F6CDD1D)))


def _u01(h):
    return (h >> U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_reads(n_pairs, read_len=150, mito_len=16500, nuclear_len=50_000_000, mito_fraction=0.05, error_rate=0.005,
                n_rate=1e-4, insert_mean=350.0, insert_sd=35.0, seed=1001, first_read=0, n_reads=None, chunk=100_000, workers=None):
    """Reads [first_read, first_read + n_reads) of the `n_pairs`-pair synthetic library, generated `chunk` reads at a time.
    Returns (bases uint8 0..3 back to back, starts int64 [n+1])."""
    total = 2 * int(n_pairs)
    if n_reads is None:
        n_reads = total - first_read
    n_reads = max(0, min(int(n_reads), total - first_read))
    parts, lens = [], [np.zeros(1, np.int64)]
    jobs = [(n_pairs, read_len, mito_len, nuclear_len, mito_fraction, error_rate, n_rate, insert_mean, insert_sd, seed, r0,
             min(chunk, first_read + n_reads - r0)) for r0 in range(first_read, first_read + n_reads, chunk)]
    if workers is None:
        workers = min(len(jobs), os.cpu_count() or 1)
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.starmap(_synth_chunk, jobs)
    else:
        res = [_synth_chunk(*j) for j in jobs]
    for b, s in res:
        parts.append(b)
        lens.append(np.diff(s))
    starts = np.cumsum(np.concatenate(lens))
    return (np.concatenate(parts) if parts else np.zeros(0, np.uint8)), starts.astype(np.int64)


def _synth_chunk(n_pairs, read_len=150, mito_len=16500, nuclear_len=50_000_000, mito_fraction=0.05, error_rate=0.005,
                n_rate=1e-4, insert_mean=350.0, insert_sd=35.0, seed=1001, first_read=0, n_reads=None):
    L = int(read_len)
    total = 2 * int(n_pairs)
    if n_reads is None:
        n_reads = total - first_read
    n_reads = max(0, min(int(n_reads), total - first_read))
    read = np.arange(first_read, first_read + n_reads, dtype=np.int64)
    with np.errstate(over="ignore"):
        pair = (read >> 1).astype(U64)
        mate = (read & 1).astype(np.int64)
        is_mito = _u01(_h3(seed, pair, 1)) < mito_fraction
        u1 = np.maximum(_u01(_h3(seed, pair, 2)), 1e-300)
        u2 = _u01(_h3(seed, pair, 3))
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(np.pi * 2.0 * u2)
        ins = np.rint(insert_mean + insert_sd * z).astype(np.int64)
        ins = np.clip(ins, L, 600)
        G = np.where(is_mito, mito_len, nuclear_len).astype(np.int64)
        span = np.where(is_mito, G, np.maximum(G - ins + 1, 1))
        frag_start = (_u01(_h3(seed, pair, 4)) * span.astype(np.float64)).astype(np.int64)
        flip = (_h3(seed, pair, 5) & U64(1)).astype(np.int64)
        hr = _h3(seed, read.astype(U64), 6)
        lead = np.where((hr & U64(1023)) == 0, 1 + ((hr >> U64(10)) % U64(5)).astype(np.int64), 0)
        end = np.where(((hr >> U64(20)) & U64(1023)) == 0, L - 1 - ((hr >> U64(30)) % U64(5)).astype(np.int64), L)
        if n_rate > 0:
            u = np.maximum(_u01(_h3(seed, read.astype(U64), 7)), 1e-300)
            gap = np.floor(np.log(u) / np.log1p(-n_rate))
            cut = gap < (end - lead).astype(np.float64)
            end = np.where(cut, lead + np.where(cut, gap, 0).astype(np.int64), end)
        end = np.maximum(end, lead)
        length = end - lead
        starts = np.zeros(n_reads + 1, np.int64)
        starts[1:] = np.cumsum(length)
        nb = int(starts[-1])
        # per base
        rid = np.repeat(np.arange(n_reads, dtype=np.int64), length)
        off = np.arange(nb, dtype=np.int64) - starts[rid] + lead[rid]
        m = mate[rid]
        fpos = np.where(m == 1, ins[rid] - 1 - off, off)
        comp = m.copy()
        fl = flip[rid]
        gpos = np.where(fl == 1, frag_start[rid] + (ins[rid] - 1 - fpos), frag_start[rid] + fpos)
        comp ^= fl
        mito = is_mito[rid]
        gp = np.where(mito, np.mod(gpos, mito_len), gpos)
        tag = np.where(mito, U64(0x4D49544F), U64(0x4E55434C)).astype(U64)
        b = (_h3(seed, tag, gp.astype(U64)) & U64(3)).astype(np.int64)
        b = np.where(comp == 1, 3 - b, b)
        he = _h3(U64(seed) ^ U64(0x5EED), (read[rid]).astype(U64), (off + 16).astype(U64))
        err = _u01(he) < error_rate
        b = np.where(err, (b + 1 + ((he & U64(0xFF)) % U64(3)).astype(np.int64)) & 3, b)
    return b.astype(np.uint8), starts

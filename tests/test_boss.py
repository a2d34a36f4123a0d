"""Oracle-independent BOSS validity (SURVEY.md B.6a invariants 1-5) of the oracle's graphs, on the CPU.
The decoder (tests/boss.py) recovers node labels by LF mapping over W / last / tip alone, the way the graph's only consumer
(`megahit_core assemble -s`, /root/reference/assemble/assemble_wrapper.py:264-295) navigates it; the expected edge multiset
comes straight from the reads with numpy.  tests/test_gpu_boss.py runs the same check on the GPU's output."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import boss
from gpu_common import make_reads


def _graph(g):
    return dict(w=g.w, last=g.last, tip=g.tip, mul=g.mul, tip_labels=g.tip_labels)


@pytest.mark.parametrize("k,m", [(9, 1), (15, 2), (21, 2), (21, 1), (31, 2), (32, 2), (47, 2), (63, 1), (79, 2), (141, 2)])
def test_oracle_graph_round_trip(oracle, k, m):
    bases, starts = make_reads(4000 + k + m, 1500, k, genome_len=3000, max_len=max(150, k + 40), err=0.01)
    g = oracle.read2sdbg(oracle.Reads(bases, starts), k, m, threads=4)
    n = boss.assert_round_trip(_graph(g), bases, starts, k, m)
    assert n > 0
    # count + seq2sdbg reaches the same graph through the edge-record path
    e = oracle.count(oracle.Reads(bases, starts), k, m, threads=4)
    s = oracle.Seqs()
    s.add_edges(e)
    g2 = oracle.seq2sdbg(s, k, threads=4)
    assert boss.assert_round_trip(_graph(g2), bases, starts, k, m) == n


def test_decoder_rejects_broken_graphs(oracle):
    k, m = 21, 2
    bases, starts = make_reads(77, 800, k, genome_len=2000, max_len=120, err=0.01)
    g = _graph(oracle.read2sdbg(oracle.Reads(bases, starts), k, m, threads=2))
    boss.assert_round_trip(g, bases, starts, k, m)
    # flip a W, a last bit, a multiplicity, swap two items: each must be noticed
    rng = np.random.default_rng(5)
    real = np.nonzero((g["w"] > 0) & (g["tip"] == 0))[0]
    for mutate in ("w", "last", "mul", "swap"):
        h = {f: np.array(v, copy=True) for f, v in g.items()}
        i = int(rng.choice(real))
        if mutate == "w":
            h["w"][i] = (h["w"][i] - 1 + 1) % 4 + 1 + (4 if h["w"][i] > 4 else 0)
        elif mutate == "last":
            h["last"][i] ^= 1
        elif mutate == "mul":
            h["mul"][i] += 1
        else:
            j = int(real[np.searchsorted(real, i) - 1]) if i != real[0] else int(real[1])
            if h["w"][i] == h["w"][j] and h["last"][i] == h["last"][j] and h["mul"][i] == h["mul"][j]:
                continue
            for f in ("w", "last", "mul"):
                h[f][i], h[f][j] = h[f][j], h[f][i]
        with pytest.raises(AssertionError):
            boss.assert_round_trip(h, bases, starts, k, m)


_dna = st.text(alphabet="ACGT", min_size=0, max_size=60)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(reads=st.lists(_dna, min_size=0, max_size=25), k=st.sampled_from([9, 11, 16, 21]), m=st.integers(1, 3),
       dup=st.integers(1, 4))
def test_property_round_trip(oracle, reads, k, m, dup):
    """hypothesis: any read set, incl. empty / short reads, repeats, low-complexity."""
    reads = reads * dup
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    starts = np.zeros(len(reads) + 1, np.int64)
    starts[1:] = np.cumsum([len(r) for r in reads])
    bases = np.array([code[c] for r in reads for c in r], dtype=np.uint8)
    g = oracle.read2sdbg(oracle.Reads(bases, starts), k, m, threads=1)
    boss.assert_round_trip(_graph(g), bases, starts, k, m)
    e = oracle.count(oracle.Reads(bases, starts), k, m, threads=1)
    exp_e, exp_c = boss.expected_edges(bases, starts, k, m)
    # the canonical half of the expected set is the edge file, in order, with the same multiplicities
    canon = exp_e[(exp_e <= boss._rc(exp_e))[np.arange(len(exp_e)), np.argmax(exp_e != boss._rc(exp_e), axis=1)]] if len(exp_e) else exp_e
    assert e.n == len(canon)

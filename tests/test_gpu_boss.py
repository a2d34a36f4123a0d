"""Oracle-independent BOSS validity of the GPU's graphs (SURVEY.md B.6a invariants 1-5).
Nothing here touches oracle/: the decoder (tests/boss.py) navigates W / last / tip like `megahit_core assemble -s`
(/root/reference/assemble/assemble_wrapper.py:264-295) and the expected edge multiset comes from the reads with numpy."""
import numpy as np
import pytest

import boss
from gpu_common import make_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from mitoflex_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("k,m", [(21, 2), (21, 1), (31, 2), (47, 2), (63, 2), (99, 2), (141, 2)])
def test_gpu_graph_round_trip(ctx, k, m):
    bases, starts = make_reads(9100 + k + m, 4000, k, genome_len=8000, max_len=max(150, k + 40), err=0.01)
    reads = ctx.upload_reads(bases, starts)
    g = ctx.read2sdbg(reads, k, m).to_numpy()
    n = boss.assert_round_trip(g, bases, starts, k, m)
    assert n > 0
    g2 = ctx.seq2sdbg(ctx.count(reads, k, m)).to_numpy()
    assert boss.assert_round_trip(g2, bases, starts, k, m) == n


def test_gpu_graph_round_trip_deep_and_skewed(ctx):
    """a 2 kb genome at ~2000x (error neighbours become solid, multiplicities beyond 254) plus 3000 copies of one read"""
    k, m = 21, 2
    bases, starts = make_reads(31337, 30000, k, genome_len=2000, max_len=150, err=0.01, dup_boost=3000)
    g = ctx.read2sdbg(ctx.upload_reads(bases, starts), k, m).to_numpy()
    assert boss.assert_round_trip(g, bases, starts, k, m) > 0
    assert (g["mul"] > 254).any()


def test_gpu_graph_round_trip_synth(ctx):
    """the in-HBM generator of the bench (mitogenome + nuclear background, N-trimmed reads)"""
    k, m = 21, 2
    reads = ctx.synth(n_pairs=20000, nuclear_len=200_000, seed=77)
    bases, starts = ctx.download_reads(reads)
    g = ctx.read2sdbg(reads, k, m).to_numpy()
    assert boss.assert_round_trip(g, bases, starts, k, m) > 0

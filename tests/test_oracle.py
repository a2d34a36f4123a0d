"""Pins for the CPU oracle (oracle/mh_oracle.c).

The reference has no tests / golden vectors for this path and megahit itself is not
vendored -> PARITY UNPINNED.  What pins the oracle instead:
  * hand-derived known-answer vectors (every expected value below was worked out by hand
    from the 2-bit encoding rules, see comments),
  * an independent string-level Python model (tests/pymodel.py) on random inputs,
  * the BOSS invariants of SURVEY.md B.6a.
"""
import random

import numpy as np
import pytest

import pymodel as pm


def _reads(orc, seqs, policy=0):
    return orc.Reads.from_ascii(seqs, policy)


# ---------------------------------------------------------------- count KATs
def test_kat_single_edge(oracle):
    # read ACGTACGTAC, k=9 -> one 10-mer.  stored (reversed) = CATGCATGCA, its revcomp = TGCATGCATG,
    # canonical = CATGCATGCA = 01 00 11 10 01 00 11 10 01 00 -> 0x4E4E4 << 12; words_per_edge = ceil(36/32) = 2.
    e = oracle.count(_reads(oracle, ["ACGTACGTAC"]), 9, 1)
    assert e.words == 2 and e.k == 9
    assert e.data.tolist() == [[0x4E4E4000, 1]]
    assert e.bucket_counts[0x4E4E] == 1 and e.bucket_counts.sum() == 1
    assert e.counting[1] == 1 and e.counting.sum() == 1


def test_kat_both_strands_merge(oracle):
    # revcomp(ACGTACGTAC) = GTACGTACGT is the same canonical edge -> count 2
    e = oracle.count(_reads(oracle, ["ACGTACGTAC", "GTACGTACGT"]), 9, 2)
    assert e.data.tolist() == [[0x4E4E4000, 2]]
    assert oracle.count(_reads(oracle, ["ACGTACGTAC", "GTACGTACGT"]), 9, 3).n == 0


def test_kat_palindrome(oracle):
    # ACGTTAACGT is its own reverse complement; stored = TGCAATTGCA = 11 10 01 00 00 11 11 10 01 00 -> 0xE43E4
    e = oracle.count(_reads(oracle, ["ACGTTAACGT"] * 3), 9, 1)
    assert e.data.tolist() == [[0xE43E4000, 3]]


def test_kat_n_policy(oracle):
    r = _reads(oracle, ["NNACGTACGTACNACGT", "ACGTN", "NNNN", "ACGT"], 0)
    assert np.diff(r.starts).tolist() == [10, 4, 0, 4]
    assert "".join("ACGT"[b] for b in r.bases[:10]) == "ACGTACGTAC"
    r = _reads(oracle, ["NNACGTACGTACNACGT", "NNNN"], 1)
    assert np.diff(r.starts).tolist() == [10, 4, 0]


def test_kat_short_reads_skipped(oracle):
    assert oracle.count(_reads(oracle, ["ACGTACGTA", "", "ACG"]), 9, 1).n == 0


def test_kat_multiplicity_cap(oracle):
    e = oracle.count(_reads(oracle, ["ACGTACGTAC"] * 65540), 9, 1)
    assert e.data.tolist() == [[0x4E4E4000, 65535]]
    assert e.counting[65535] == 1


def test_kat_k31_three_words(oracle):
    # k=31: 32 bases = 64 bits of key + 16 bits multiplicity -> 3 words, last word = multiplicity only
    s = "A" * 31 + "C"          # stored = C + A*31 ; revcomp(stored) = T*31 + G  -> canonical = stored
    e = oracle.count(_reads(oracle, [s]), 31, 1)
    assert e.words == 3
    assert e.data.tolist() == [[0x40000000, 0, 1]]


def test_kat_k21_two_words(oracle):
    # k=21: 22 bases = 44 bits; multiplicity shares word 1 (low 16 bits)
    s = "C" + "A" * 21          # stored = A*21 + C -> word0 = 0, word1 = 01 at bits 21..20 -> 0x00100000
    e = oracle.count(_reads(oracle, [s, s]), 21, 1)
    assert e.data.tolist() == [[0, 0x00100002]]


# ------------------------------------------------------- count vs python model
@pytest.mark.parametrize("k,m", [(9, 1), (9, 2), (15, 2), (16, 1), (21, 2), (31, 1), (32, 2), (47, 1), (63, 2)])
def test_count_vs_model(oracle, k, m):
    rng = random.Random(k * 100 + m)
    genome = "".join(rng.choice("ACGT") for _ in range(300))
    reads = []
    for _ in range(120):
        L = rng.randint(k - 2, min(150, k + 60))
        p = rng.randint(0, len(genome) - L)
        r = genome[p:p + L]
        if rng.random() < 0.5:
            r = pm.revcomp(r)
        if rng.random() < 0.3:
            q = rng.randrange(len(r))
            r = r[:q] + rng.choice("ACGT") + r[q + 1:]
        reads.append(r)
    exp = pm.count_edges(reads, k, m)
    e = oracle.count(_reads(oracle, reads), k, m, threads=3)
    assert e.n == len(exp)
    for row, (s, c) in zip(e.data.tolist(), exp):
        assert row == pm.edge_words(s, c, k)


# ---------------------------------------------------------------- sdbg
def _unpack_label(words, k):
    return "".join("ACGT"[(words[i >> 4] >> (30 - 2 * (i & 15))) & 3] for i in range(k))


def _check_sdbg(g, exp, k):
    assert g.n == len(exp)
    assert g.w.tolist() == [x["w"] for x in exp]
    assert g.last.tolist() == [x["last"] for x in exp]
    assert g.tip.tolist() == [x["tip"] for x in exp]
    assert g.mul.tolist() == [x["mul"] for x in exp]
    tips = [x for x in exp if x["tip"]]
    assert len(tips) == g.tip_labels.shape[0]
    for words, x in zip(g.tip_labels.tolist(), tips):
        assert _unpack_label(words, k - 1) == x["label"][:k - 1]


def test_kat_sdbg_single_edge(oracle):
    """k=9, one edge CATGCATGCA (mult 7).  Worked by hand:
    strand0 s = CATGCATGCA, strand1 rc = TGCATGCATG.  Items (9-mer|b|cnt):
      s:  CATGCATGC|$|0  ATGCATGCA|C|7  TGCATGCA$|A|0
      rc: TGCATGCAT|$|0  GCATGCATG|T|7  CATGCAT G$ -> CATGCATG$|G|0
    sorted by 9-mer ($ as A, flag breaks the tie): ATGCATGCA, CATGCATG$ (flag 0) < CATGCATGC, GCATGCATG, TGCATGCA$ < TGCATGCAT.
    groups by 8-prefix: {ATGCATGC: A|C}, {CATGCATG: $|G, C|$}, {GCATGCAT: G|T}, {TGCATGCA: $|A, T|$}.
    No group has a solid item next to a dummy, so nothing is suppressed:
      ATGCATGCA b=C: w=2 last=1 tip=0 mul=7
      CATGCATG$ b=G: w=3 last=0 tip=1 mul=0 ; CATGCATGC b=$: w=0 last=1 tip=0 mul=0
      GCATGCATG b=T: w=4 last=1 tip=0 mul=7
      TGCATGCA$ b=A: w=1 last=0 tip=1 mul=0 ; TGCATGCAT b=$: w=0 last=1 tip=0 mul=0
    """
    s = oracle.Seqs()
    s.add([pm.CODE[c] for c in "CATGCATGCA"], 7)
    g = oracle.seq2sdbg(s, 9)
    assert g.w.tolist() == [2, 3, 0, 4, 1, 0]
    assert g.last.tolist() == [1, 0, 1, 1, 0, 1]
    assert g.tip.tolist() == [0, 1, 0, 0, 1, 0]
    assert g.mul.tolist() == [7, 0, 0, 7, 0, 0]
    assert g.words_per_tip == 1
    # tip label = raw first word of the item: 8 bases, then b<<16 | 0xFFFF (flag bit 19 = 0)
    # CATGCATG = 01 00 11 10 01 00 11 10 = 0x4E4E -> bits 31..16; the 9th base slot (bits 15..14) and everything
    # below belong to the same single word: 9 bases = 18 bits, so flags live at bits 19.. -> they overlap? no:
    # words_per_substr = ceil((18+4+16)/32) = 2, so flags are in word 1 and the label word is clean.
    assert g.tip_labels.tolist() == [[0x4E4E0000], [0xE4E40000]]


def test_kat_sdbg_dummy_suppression(oracle):
    """Two overlapping edges of the 11-base string CATGCATGCAA (k=9): CATGCATGCA and ATGCATGCAA.
    The tail dummy of the first (TGCATGCA$|A) and the head dummy of the second (ATGCATGCA|$) are each
    covered by a real edge in their group and must disappear; same on the reverse strand."""
    s = oracle.Seqs()
    for e in ("CATGCATGCA", "ATGCATGCAA"):
        s.add([pm.CODE[c] for c in e], 3)
    g = oracle.seq2sdbg(s, 9)
    exp = pm.sdbg_from_seqs([("CATGCATGCA", 3), ("ATGCATGCAA", 3)], 9)
    _check_sdbg(g, exp, 9)
    one = oracle.Seqs()
    one.add([pm.CODE[c] for c in "CATGCATGCAA"], 3)
    g1 = oracle.seq2sdbg(one, 9)
    assert g1.w.tolist() == g.w.tolist() and g1.mul.tolist() == g.mul.tolist()
    assert g.tip.sum() == 2 and (g.w == 0).sum() == 2


def _random_reads(rng, k, n=80, glen=200, err=0.2):
    genome = "".join(rng.choice("ACGT") for _ in range(glen))
    reads = []
    for _ in range(n):
        L = rng.randint(k, min(glen, k + 40))
        p = rng.randint(0, glen - L)
        r = genome[p:p + L]
        if rng.random() < 0.5:
            r = pm.revcomp(r)
        if rng.random() < err:
            q = rng.randrange(len(r))
            r = r[:q] + rng.choice("ACGT") + r[q + 1:]
        reads.append(r)
    return reads


@pytest.mark.parametrize("k,m", [(9, 1), (9, 2), (13, 2), (15, 1), (16, 2), (17, 2), (21, 2), (31, 2), (32, 1), (45, 2)])
def test_seq2sdbg_vs_model(oracle, k, m):
    rng = random.Random(7000 + k * 10 + m)
    reads = _random_reads(rng, k)
    e = oracle.count(_reads(oracle, reads), k, m)
    s = oracle.Seqs()
    s.add_edges(e)
    g = oracle.seq2sdbg(s, k, threads=2)
    exp = pm.sdbg_from_seqs(pm.count_edges(reads, k, m), k)
    _check_sdbg(g, exp, k)
    _check_invariants(g, exp, k)


@pytest.mark.parametrize("k,m", [(9, 1), (9, 2), (13, 3), (16, 2), (21, 2), (31, 1), (33, 2)])
def test_read2sdbg_vs_model_and_two_pass(oracle, k, m):
    rng = random.Random(9000 + k * 10 + m)
    reads = _random_reads(rng, k)
    # palindromic (k+1)-mers when k+1 is even
    if (k + 1) % 2 == 0:
        half = "".join(rng.choice("ACGT") for _ in range((k + 1) // 2))
        reads += [half + pm.revcomp(half)] * 3
    g = oracle.read2sdbg(_reads(oracle, reads), k, m, threads=2)
    exp = pm.sdbg_from_reads(reads, k, m)
    _check_sdbg(g, exp, k)
    # one-pass == count + seq2sdbg on the logical arrays
    e = oracle.count(_reads(oracle, reads), k, m)
    s = oracle.Seqs()
    s.add_edges(e)
    g2 = oracle.seq2sdbg(s, k)
    for f in ("w", "last", "tip", "mul"):
        assert getattr(g, f).tolist() == getattr(g2, f).tolist(), f
    km1 = [_unpack_label(x, k - 1) for x in g.tip_labels.tolist()]
    assert km1 == [_unpack_label(x, k - 1) for x in g2.tip_labels.tolist()]


def _check_invariants(g, exp, k):
    """SURVEY.md B.6a invariants 1-3."""
    w, last, tip, mul = g.w, g.last, g.tip, g.mul
    assert last.sum() == ((w >= 1) & (w <= 4)).sum()
    assert not (tip & last).any()
    assert (mul[tip == 1] == 0).all() and (mul[w == 0] == 0).all()
    nodes = {}
    for x in exp:
        if not x["tip"]:
            nodes[x["label"]] = 1
    for c in range(4):
        assert sum(1 for lab in nodes if lab[0] == "ACGT"[c]) == (w == c + 1).sum()


def test_contig_reader(oracle, tmp_path):
    fa = tmp_path / "k21.contigs.fa"
    fa.write_text(">k21_0 flag=1 multi=12.4999 len=30\n" + "ACGT" * 7 + "AC\n"
                  ">k21_1 flag=2 multi=3.5000 len=25\n" + "ACGTTGCAAGGCTTAACCGGTTAAC\n"
                  ">k21_2 flag=0 multi=70000.0 len=12\nACGTACGTACGT\n")
    s = oracle.Seqs()
    s.add_contigs(str(fa), 26, True, 21, 25)   # k_to = 25 -> min_len 26, loop contig extended by 4 -> 29
    assert s.n == 2
    s2 = oracle.Seqs()
    s2.add_contigs(str(fa), 10)
    assert s2.n == 3


def test_files_round_trip(oracle, tmp_path):
    rng = random.Random(5)
    reads = _random_reads(rng, 21, n=300, glen=2000)
    fq1, fq2 = tmp_path / "a_1.fq", tmp_path / "a_2.fq"
    with open(fq1, "w") as f1, open(fq2, "w") as f2:
        for i in range(0, len(reads), 2):
            a, b = reads[i], reads[i + 1]
            if i % 20 == 0:
                a = "NN" + a[2:]
            f1.write(f"@r{i}/1\n{a}\n+\n{'I' * len(a)}\n")
            f2.write(f"@r{i}/2\n{b}\n+\n{'I' * len(b)}\n")
    libf = tmp_path / "reads.lib"
    libf.write_text(f"{fq1},{fq2}\npe {fq1} {fq2}\n")
    oracle.cmd_buildlib(str(libf), str(libf))
    info = [x.split(" ") for x in open(str(libf) + ".lib_info").read().splitlines()]
    assert int(info[0][1]) == len(reads) and info[2][0] == "pe" and int(info[2][1]) == 0 and int(info[2][2]) == len(reads) - 1
    r = oracle.Reads.load_bin(str(libf) + ".bin")
    assert r.n == len(reads) and int(info[0][0]) == len(r.bases)
    trimmed = [pm.trim_n(("NN" + x[2:]) if (i % 20 == 0 and i % 2 == 0) else x) for i, x in enumerate(reads)]
    assert np.diff(r.starts).tolist() == [len(x) for x in trimmed]
    pref = str(tmp_path / "k21")
    oracle.cmd_count(str(libf), 21, 2, pref, threads=3)
    e = oracle.Edges.read(pref)
    e_mem = oracle.count(r, 21, 2)
    assert np.array_equal(e.data, e_mem.data)
    oracle.cmd_seq2sdbg(21, 0, pref, input_prefix=pref, threads=2)
    g = oracle.Sdbg.read(pref)
    s = oracle.Seqs()
    s.add_edges(e_mem)
    g_mem = oracle.seq2sdbg(s, 21)
    for f in ("w", "last", "tip", "mul", "tip_labels"):
        assert np.array_equal(getattr(g, f), getattr(g_mem, f)), f
    pref2 = str(tmp_path / "one")
    oracle.cmd_read2sdbg(str(libf), 21, 2, pref2)
    g1 = oracle.Sdbg.read(pref2)
    assert np.array_equal(g1.w, g.w) and np.array_equal(g1.mul, g.mul) and np.array_equal(g1.last, g.last)

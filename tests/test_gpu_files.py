"""File-level parity on the GPU: the four sub-commands, through the C ABI and through the megahit_core shim, against the
files the oracle writes for the same inputs.  With the same number of output files the bytes must be identical."""
import filecmp
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from gpu_common import make_reads

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_fastq(tmp, bases, starts, with_n=True):
    rng = np.random.default_rng(11)
    n = len(starts) - 1
    n -= n % 2
    f1, f2 = tmp / "r_1.fq", tmp / "r_2.fq"
    with open(f1, "w") as a, open(f2, "w") as b:
        for i in range(n):
            s = "".join("ACGT"[x] for x in bases[starts[i]:starts[i + 1]])
            if with_n and len(s) > 30 and rng.random() < 0.15:
                q = int(rng.integers(0, len(s)))
                s = s[:q] + "N" * int(rng.integers(1, 4)) + s[q + 1:]
            if with_n and rng.random() < 0.05:
                s = "NN" + s
            out = a if i % 2 == 0 else b
            out.write(f"@r{i // 2}/{i % 2 + 1}\n{s}\n+\n{'I' * len(s)}\n")
    libf = tmp / "reads.lib"
    libf.write_text(f"{f1},{f2}\npe {f1} {f2}\n")
    return str(libf)


def _same_files(pa, pb, pattern):
    fa = sorted(glob.glob(pa + pattern))
    fb = sorted(glob.glob(pb + pattern))
    assert [os.path.basename(x)[len(os.path.basename(pa)):] for x in fa] == [os.path.basename(x)[len(os.path.basename(pb)):] for x in fb]
    assert fa, f"no files match {pa}{pattern}"
    for x, y in zip(fa, fb):
        assert filecmp.cmp(x, y, shallow=False), (x, y)


@pytest.mark.parametrize("policy", [0, 1])
def test_buildlib_bytes(oracle, tmp_path, policy):
    from mitoflex_b200 import lib
    bases, starts = make_reads(31, 4000, 21, genome_len=20000)
    libf = _write_fastq(tmp_path, bases, starts)
    g, o = str(tmp_path / "gpu.lib"), str(tmp_path / "orc.lib")
    lib.buildlib(libf, g, policy)
    oracle.cmd_buildlib(libf, o, policy)
    assert filecmp.cmp(g + ".bin", o + ".bin", shallow=False)
    assert open(g + ".lib_info").read() == open(o + ".lib_info").read()


@pytest.mark.parametrize("io_chunk", [None, 1000])
def test_count_seq2sdbg_read2sdbg_files(oracle, tmp_path, monkeypatch, io_chunk):
    """io_chunk: the writers' pinned pieces cut down to 1000 bytes, so that the edge stream and the device-serialised sdbg stream
    (items beyond multiplicity 254 and tip labels included: 400 copies of one read) cross piece and file boundaries everywhere."""
    from mitoflex_b200 import lib
    k, m, threads = 21, 2, 3
    if io_chunk:
        monkeypatch.setenv("MFSDBG_IO_CHUNK", str(io_chunk))
    bases, starts = make_reads(77, 30000, k, genome_len=60000, dup_boost=400 if io_chunk else 0)
    libf = _write_fastq(tmp_path, bases, starts)
    lib.buildlib(libf, libf)
    os.makedirs(tmp_path / "g"), os.makedirs(tmp_path / "o")
    g, o = str(tmp_path / "g" / "21"), str(tmp_path / "o" / "21")
    lib.count(k=k, min_count=m, host_mem=1 << 30, mem_flag=1, output_prefix=g, num_cpu_threads=threads, read_lib_file=libf)
    oracle.cmd_count(libf, k, m, o, threads)
    _same_files(g, o, ".edges.*")
    assert open(g + ".counting").read() == open(o + ".counting").read()
    lib.seq2sdbg(k=k, kmer_from=0, host_mem=1 << 30, mem_flag=1, output_prefix=g, num_cpu_threads=threads, input_prefix=g)
    oracle.cmd_seq2sdbg(k, 0, o, input_prefix=o, threads=threads)
    _same_files(g, o, ".sdbg*")
    g1, o1 = str(tmp_path / "g" / "one"), str(tmp_path / "o" / "one")
    lib.read2sdbg(k=k, min_count=m, output_prefix=g1, num_cpu_threads=threads, read_lib_file=libf)
    oracle.cmd_read2sdbg(libf, k, m, o1, threads)
    _same_files(g1, o1, ".sdbg*")


@pytest.mark.parametrize("rounds", [False, True])
def test_next_k_with_contigs(oracle, tmp_path, monkeypatch, rounds):
    """k > k_min: unsorted iterative edges + contig / bubble / addi / local FASTA of the previous k (loop contigs extended);
    once in one pass, once through the memory-bounded rounds (ranged edge and contig item generators)."""
    from mitoflex_b200 import lib
    if rounds:
        monkeypatch.setenv("MFSDBG_SDBG_ROUND_ITEMS", "20000")
    k_from, k = 21, 29
    rng = np.random.default_rng(5)
    genome = rng.integers(0, 4, 30000, dtype=np.uint8)
    txt = lambda a: "".join("ACGT"[x] for x in a)   # noqa: E731
    contigs = tmp_path / "k21.contigs.fa"
    with open(contigs, "w") as f:
        pos, i = 0, 0
        while pos < len(genome) - 400:
            L = int(rng.integers(25, 400))
            flag = int(rng.choice([0, 1, 2]))
            f.write(f">k21_{i} flag={flag} multi={rng.uniform(1, 300):.4f} len={L}\n{txt(genome[pos:pos + L])}\n")
            pos += L - int(rng.integers(0, 20))
            i += 1
    for name in ("k21.bubble_seq.fa", "k21.addi.fa", "k21.local.fa"):
        with open(tmp_path / name, "w") as f:
            for j in range(40):
                p = int(rng.integers(0, len(genome) - 200))
                L = int(rng.integers(20, 200))
                f.write(f">x_{j} flag=0 multi={rng.uniform(1, 70000):.4f} len={L}\n{txt(genome[p:p + L])}\n")
    # iterative edges: an UNSORTED edge file (what `megahit_core iterate` writes), here the oracle's sorted edges shuffled
    bases, starts = make_reads(9, 8000, k, genome_len=30000)
    e = oracle.count(oracle.Reads(bases, starts), k, 1)
    perm = rng.permutation(e.n)
    pref = str(tmp_path / "29")
    with open(pref + ".edges.0", "wb") as f:
        f.write(e.data[perm].tobytes())
    open(pref + ".edges.info", "w").write(f"kmer_size {k}\nwords_per_edge {e.words}\nnum_files 1\nnum_buckets 0\nnum_edges {e.n}\nis_sorted 0\n")
    g, o = str(tmp_path / "gpu29"), str(tmp_path / "orc29")
    kw = dict(input_prefix=pref, contig=str(contigs), bubble=str(tmp_path / "k21.bubble_seq.fa"),
              addi_contig=str(tmp_path / "k21.addi.fa"), local_contig=str(tmp_path / "k21.local.fa"))
    lib.seq2sdbg(k=k, kmer_from=k_from, output_prefix=g, num_cpu_threads=2, **kw)
    oracle.cmd_seq2sdbg(k, k_from, o, threads=2, **kw)
    _same_files(g, o, ".sdbg*")


def test_shim_and_wrapper_end_to_end(oracle, tmp_path):
    """the unmodified MitoFlex call sequence: shell commands built like utility/helper.py builds them."""
    from mitoflex_b200 import wrapper
    k, m = 21, 3
    bases, starts = make_reads(123, 20000, k, genome_len=40000)
    libf = _write_fastq(tmp_path, bases, starts, with_n=False)
    exe = os.path.join(ROOT, "mitoflex_b200", "bin", "megahit_core")
    run = lambda cmd: subprocess.check_output(f"{sys.executable} {exe} {cmd}", shell=True)   # noqa: E731
    run(f"buildlib {libf} {libf}")
    info = [x.split(" ") for x in open(libf + ".lib_info").readlines()]
    assert int(info[0][1]) == len(starts) - 1 - (len(starts) - 1) % 2     # LibInfo.read_count (assemble_wrapper.py:49)
    os.makedirs(tmp_path / "T" / "k21")
    pref = str(tmp_path / "T" / "k21" / "21")
    run(f"count -k {k} --host_mem 1073741824.0 --mem_flag 1 --output_prefix {pref} --num_cpu_threads 2 -m {m} --read_lib_file {libf}")
    assert os.path.exists(pref + ".edges.0")                              # probed at assemble_wrapper.py:228
    run(f"seq2sdbg -k {k} --host_mem 1073741824.0 --mem_flag 1 --output_prefix {pref} --num_cpu_threads 2 --kmer_from 0 --input_prefix {pref}")
    o = str(tmp_path / "orc")
    oracle.cmd_count(libf, k, m, o, 2)
    oracle.cmd_seq2sdbg(k, 0, o, input_prefix=o, threads=2)
    _same_files(pref, o, ".sdbg*")
    # the Python mirror of MEGAHIT.graph does the same two steps through ctypes
    gb = wrapper.GraphBuilder(str(tmp_path / "W"), str(tmp_path / "W" / "contigs"), libf, threads=2, min_multi=m)
    os.makedirs(tmp_path / "W" / "contigs")
    gb.graph(0, k)
    _same_files(gb._graph_prefix(k), o, ".sdbg*")
    with pytest.raises(wrapper.EmptyGraph):
        gb.graph(k, 29)
    # failure: exit status 1 and a message, no partial meta file
    p = subprocess.run(f"{sys.executable} {exe} count -k 21 -m 2 --output_prefix {tmp_path / 'nope'} --read_lib_file {tmp_path / 'missing.lib'}",
                       shell=True, capture_output=True)
    assert p.returncode == 1 and b"cannot open" in p.stderr and not os.path.exists(str(tmp_path / "nope") + ".edges.info")


@pytest.mark.parametrize("chunk", [4096, 50_000])
def test_buildlib_streams_in_chunks(oracle, tmp_path, monkeypatch, chunk):
    """buildlib reads the FASTQ files in chunks that end on record boundaries (a helper thread reads chunk i+1 while chunk i is
    packed on the GPU); the two files of a pair have reads of different lengths, so their chunks hold different numbers of
    records and the surplus is handed back.  Tiny chunks force hundreds of them; the .bin / .lib_info must not change."""
    from mitoflex_b200 import lib
    monkeypatch.setenv("MFSDBG_TEXT_CHUNK", str(chunk))
    bases, starts = make_reads(37, 6000, 21, genome_len=20000, max_len=140, short_frac=0.3)
    libf = _write_fastq(tmp_path, bases, starts)
    g, o = str(tmp_path / "gpu.lib"), str(tmp_path / "orc.lib")
    for policy in (0, 1):
        lib.buildlib(libf, g, policy)
        oracle.cmd_buildlib(libf, o, policy)
        assert filecmp.cmp(g + ".bin", o + ".bin", shallow=False)
        assert open(g + ".lib_info").read() == open(o + ".lib_info").read()
    # single-end and a truncated pair
    f1 = str(tmp_path / "r_1.fq")
    se = tmp_path / "se.lib"
    se.write_text(f"{f1}\nse {f1}\n")
    lib.buildlib(str(se), g)
    oracle.cmd_buildlib(str(se), o)
    assert filecmp.cmp(g + ".bin", o + ".bin", shallow=False)
    short = tmp_path / "short_2.fq"
    short.write_text("".join(open(tmp_path / "r_2.fq").readlines()[:-4]))
    bad = tmp_path / "bad.lib"
    bad.write_text(f"{f1},{short}\npe {f1} {short}\n")
    with pytest.raises(lib.MfsdbgError):
        lib.buildlib(str(bad), g)

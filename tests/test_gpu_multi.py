"""Several GPUs behind the file-level C ABI in ONE process (mitoflex_b200/csrc/multi.cu): the calls the
reference makes (assemble_wrapper.py:224,258 launch one process) with `n_gpus` / MFSDBG_GPU naming 2 devices.  Every GPU
writes its own <prefix>.edges.<r> / .sdbg.<r>; the meta file stitches them, so the LOGICAL stream (buckets in order, read
through the meta) must equal the oracle's, bit for bit.  Needs 2 GPUs."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from gpu_common import make_reads
from test_gpu_files import _write_fastq

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _two_gpus():
    from mitoflex_b200 import lib
    if lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    return (ctypes.c_int32 * 2)(0, 1)


def _assert_sdbg(g, o):
    for f in ("w", "last", "tip", "mul", "tip_labels", "bucket_items"):
        assert np.array_equal(getattr(g, f), getattr(o, f)), f
    assert g.n_large == o.n_large


@pytest.mark.parametrize("k,m", [(21, 2), (31, 1), (47, 2)])
def test_two_gpu_files_match_the_oracle(oracle, tmp_path, k, m):
    from mitoflex_b200 import lib
    ids = _two_gpus()
    bases, starts = make_reads(600 + k, 30000, k, genome_len=80000, max_len=max(150, k + 40), err=0.005)
    libf = _write_fastq(tmp_path, bases, starts)
    lib.buildlib(libf, libf)
    g = str(tmp_path / "g")
    lib.count(k=k, min_count=m, output_prefix=g, num_cpu_threads=2, read_lib_file=libf, n_gpus=2, gpu_ids=ids)
    assert os.path.exists(g + ".edges.0") and os.path.exists(g + ".edges.1")     # the probe of assemble_wrapper.py:228
    r = oracle.Reads.load_bin(libf + ".bin")
    e_orc = oracle.count(r, k, m, threads=8)
    e_gpu = oracle.Edges.read(g)
    assert np.array_equal(e_gpu.data, e_orc.data)
    o = str(tmp_path / "o")
    oracle.cmd_count(libf, k, m, o, 2)
    assert open(g + ".counting").read() == open(o + ".counting").read()
    # seq2sdbg on 2 GPUs from the 2-GPU edge files
    lib.seq2sdbg(k=k, kmer_from=0, output_prefix=g, num_cpu_threads=2, input_prefix=g, n_gpus=2, gpu_ids=ids)
    s = oracle.Seqs()
    s.add_edges(e_orc)
    g_orc = oracle.seq2sdbg(s, k, threads=8)
    _assert_sdbg(oracle.Sdbg.read(g), g_orc)
    # read2sdbg on 2 GPUs
    g1 = str(tmp_path / "one")
    lib.read2sdbg(k=k, min_count=m, output_prefix=g1, num_cpu_threads=2, read_lib_file=libf, n_gpus=2, gpu_ids=ids)
    _assert_sdbg(oracle.Sdbg.read(g1), oracle.read2sdbg(r, k, m, threads=8))


def test_two_gpu_seq2sdbg_with_contigs_through_the_shim(oracle, tmp_path):
    """k > k_min on 2 GPUs via the shim (MFSDBG_GPU=0,1): unsorted iterative edges + contig FASTA, split over the GPUs by
    index, every item routed to the GPU that owns its prefix."""
    _two_gpus()
    k_from, k = 21, 29
    rng = np.random.default_rng(6)
    genome = rng.integers(0, 4, 30000, dtype=np.uint8)
    txt = lambda a: "".join("ACGT"[x] for x in a)   # noqa: E731
    contigs = tmp_path / "k21.contigs.fa"
    with open(contigs, "w") as f:
        pos, i = 0, 0
        while pos < len(genome) - 400:
            L = int(rng.integers(25, 400))
            f.write(f">k21_{i} flag={int(rng.choice([0, 1, 2]))} multi={rng.uniform(1, 300):.4f} len={L}\n{txt(genome[pos:pos + L])}\n")
            pos += L - int(rng.integers(0, 20))
            i += 1
    bases, starts = make_reads(9, 8000, k, genome_len=30000)
    e = oracle.count(oracle.Reads(bases, starts), k, 1)
    perm = rng.permutation(e.n)
    pref = str(tmp_path / "29")
    with open(pref + ".edges.0", "wb") as f:
        f.write(e.data[perm].tobytes())
    open(pref + ".edges.info", "w").write(f"kmer_size {k}\nwords_per_edge {e.words}\nnum_files 1\nnum_buckets 0\nnum_edges {e.n}\nis_sorted 0\n")
    g, o = str(tmp_path / "gpu29"), str(tmp_path / "orc29")
    exe = os.path.join(ROOT, "mitoflex_b200", "bin", "megahit_core")
    subprocess.check_call(f"{sys.executable} {exe} seq2sdbg -k {k} --host_mem 1e9 --mem_flag 1 --output_prefix {g} --num_cpu_threads 2 "
                          f"--kmer_from {k_from} --input_prefix {pref} --contig {contigs}", shell=True, env=dict(os.environ, MFSDBG_GPU="0,1"))
    oracle.cmd_seq2sdbg(k, k_from, o, threads=2, input_prefix=pref, contig=str(contigs))
    _assert_sdbg(oracle.Sdbg.read(g), oracle.Sdbg.read(o))
    assert os.path.exists(g + ".sdbg.1")


def test_bad_gpu_list_is_einval(tmp_path):
    from mitoflex_b200 import lib
    ids = (ctypes.c_int32 * 2)(0, 0)
    with pytest.raises(lib.MfsdbgError) as ei:
        lib.read2sdbg(k=21, min_count=2, output_prefix=str(tmp_path / "x"), read_lib_file=str(tmp_path / "missing"), n_gpus=2, gpu_ids=ids)
    assert ei.value.code in (lib.EINVAL, lib.EIO)

"""bench.py host logic that needs no GPU: argument defaults the driver relies on, the workload label, the algorithmic
byte model behind `roofline.achieved`, and the reference arm's behaviour on ranks other than 0."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_defaults_follow_the_driver_contract(monkeypatch):
    b = _bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = b.parse_args()
    assert a.gpus == 1 and a.impl == "b200"
    assert a.warmup >= 3 and a.steps >= 1                  # timing rules: W >= 3
    assert a.pairs == 16_666_667 and a.k == 21 and a.min_count == 2   # BASELINE configs[1] read set, headline k / -m
    assert a.error_rate == 0.005 and a.nuclear_len == 50_000_000


def test_workload_label_names_what_ran():
    b = _bench()
    head = b.workload_name(16_666_667)
    assert "k=21 -m 2" in head and "16666667xPE150" in head and "5.00 Gbp/GPU" in head
    assert "0.5% errors" in head and "BASELINE configs[1] read set" in head
    other = b.workload_name(16_666_667, 0.02, 50_000_000)
    assert "2% errors" in other and "BASELINE" not in other and "non-headline" in other
    assert "500 Mb nuclear" in b.workload_name(100, 0.005, 500_000_000)


def test_algorithmic_bytes_model():
    """k=21 -> 8-byte keys, 8-byte edges, 8-byte items; the count stages add up to ~35 B/base."""
    b = _bench()
    n_bases, n_keys, n_edges, n_items = 4_962_190_093, 4_262_962_952, 73_594_622, 149_301_499
    sb = b.stage_bytes(n_bases, n_keys, n_edges, n_items, 21)
    assert sb["count_l2_hist"] == n_keys * 8
    assert sb["count_l2_scatter"] == 2 * n_keys * 8
    assert sb["reads_scatter"] == n_bases / 4 + n_keys * 8
    assert sb["local_count"] == n_keys * 8 + n_edges * 8
    count = sb["reads_scatter"] + sb["count_l2_hist"] + sb["count_l2_scatter"] + sb["local_count"]
    assert 34.0 < count / n_bases < 36.0
    # k = 141: 9-word keys, 10-word edges and items
    sw = b.stage_bytes(1000, 100, 10, 20, 141)
    assert sw["count_l2_hist"] == 100 * 36 and sw["local_count"] == 100 * 36 + 10 * 40 and sw["local_sdbg"] == 20 * 40


def test_reference_arm_is_silent_on_other_ranks():
    """under torchrun only rank 0 runs the CPU arm; the others exit 0 without work (and without touching a GPU)."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""


def test_stream_checksum_is_position_dependent_and_additive():
    """the N>1 verification sums per-rank checksums of the ranks' pieces: pieces at their global offsets add up to the
    checksum of the whole stream, and a swap or a shifted boundary changes it"""
    import numpy as np
    b = _bench()
    rng = np.random.default_rng(1)
    v = rng.integers(0, 2**32, 10000, dtype=np.uint64).astype(np.uint32)
    whole = b.stream_checksum(v, 0)
    M = (1 << 64) - 1
    assert (b.stream_checksum(v[:3000], 0) + b.stream_checksum(v[3000:], 3000)) & M == whole
    assert (b.stream_checksum(v[:3000], 0) + b.stream_checksum(v[3000:], 2999)) & M != whole
    w = v.copy()
    w[[10, 11]] = w[[11, 10]]
    assert b.stream_checksum(w, 0) != whole
    assert b.stream_checksum(np.zeros(0, np.uint32), 5) == 0


def test_cpu_sample_keeps_the_depth():
    """the CPU arm's sample scales the nuclear background with the pair count (same depth, same solid fraction) and comes
    from the numpy port of the generator -- in a fresh process it must not map libmfsdbg.so"""
    code = ("import importlib.util, sys\n"
            f"spec = importlib.util.spec_from_file_location('b', {os.path.join(ROOT, 'bench.py')!r})\n"
            "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
            "bases, starts, desc = b.cpu_sample(16_666_667, 20_000, 0.005, 50_000_000)\n"
            "assert len(starts) - 1 == 40_000 and bases.max() <= 3 and '95x' in desc, desc\n"
            "dt, n = b.cpu_run(bases, starts, 2)\n"
            "assert n > 0\n"
            "assert 'libmfsdbg' not in open('/proc/self/maps').read()\n"
            "assert 'libmhoracle' in open('/proc/self/maps').read()\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr

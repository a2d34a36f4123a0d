#!/usr/bin/env python
"""Measurements of BASELINE.json's other configurations (one GPU), beside the headline line of bench.py:

  files   config 1 end to end FROM FILES: FASTQ -> buildlib -> count -> seq2sdbg (+ read2sdbg) on disk through the file-level
          C ABI (the calls assemble_wrapper.py:193,224,258 make), per-sub-command wall clock, K1 text GB/s, the oracle's CLI
          (CPU restatement of megahit_core) timed beside it on a bounded sample
  klist   config 2: the k-list loop 21 -> 141 on the 5 Gbp read set: read2sdbg at k_min, then per k seq2sdbg fed with
          SURVEY.md 8d's synthetic stand-ins (true-genome unitigs as contigs with multi= depth, "iterative edges" = the
          (k+1)-mers of the reads that lie on retained contigs are approximated by the solid edges of a count at that k)
  big     config 4 on one GPU: 100 M pairs (30 Gbp, 500 Mb nuclear), k=21, out-of-core rounds; rounds, ms, peak HBM
  c5      config 5 shape on one GPU: 2 % errors, -m 1 (huge distinct-edge set), item filter + memory-bounded sdbg rounds

Each leg prints one JSON line (and appends it to gpurun_out/configs_<leg>.json when that directory exists)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def emit(leg, obj):
    line = json.dumps(obj)
    print(line)
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, f"configs_{leg}.json"), "a") as f:
            f.write(line + "\n")


def peak_hbm_gb(dev=0):
    import torch
    free, total = torch.cuda.mem_get_info(dev)
    return (total - free) / 1e9, total / 1e9


def write_fastq(path1, path2, bases, starts):
    """PE FASTQ of the reads (mate 1 = even reads, mate 2 = odd reads), qualities constant 'I' (SURVEY.md 8d)."""
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    txt = lut[bases]
    n = len(starts) - 1
    with open(path1, "wb") as f1, open(path2, "wb") as f2:
        for i in range(n):
            s = txt[starts[i]:starts[i + 1]].tobytes()
            (f1 if i % 2 == 0 else f2).write(b"@r%d/%d\n%s\n+\n%s\n" % (i // 2, i % 2 + 1, s, b"I" * len(s)))


def leg_files(args):
    import torch   # noqa: F401  (device memory / context set-up)
    from mitoflex_b200 import lib
    from oracle import oracle
    work = args.work
    os.makedirs(work, exist_ok=True)
    ctx = lib.Context(0)
    reads = ctx.synth(n_pairs=args.pairs, seed=1001)
    bases, starts = ctx.download_reads(reads)
    ctx.close()
    f1, f2 = os.path.join(work, "r_1.fq"), os.path.join(work, "r_2.fq")
    t0 = time.perf_counter()
    write_fastq(f1, f2, bases, starts)
    t_write = time.perf_counter() - t0
    text_bytes = os.path.getsize(f1) + os.path.getsize(f2)
    libf = os.path.join(work, "reads.lib")
    open(libf, "w").write(f"{f1},{f2}\npe {f1} {f2}\n")
    k, m = 21, 2
    out = {"leg": "files", "pairs": args.pairs, "bases": int(starts[-1]), "fastq_bytes": text_bytes, "fastq_write_s": round(t_write, 2)}
    pref = os.path.join(work, "k21")
    steps = []
    for rep in range(2):   # second repetition = warm page cache and warm CUDA context
        rec = {}
        t0 = time.perf_counter(); lib.buildlib(libf, libf); rec["buildlib_s"] = time.perf_counter() - t0
        t0 = time.perf_counter(); lib.count(k=k, min_count=m, output_prefix=pref, num_cpu_threads=8, read_lib_file=libf, host_mem=0, mem_flag=1)
        rec["count_s"] = time.perf_counter() - t0
        t0 = time.perf_counter(); lib.seq2sdbg(k=k, kmer_from=0, output_prefix=pref, num_cpu_threads=8, input_prefix=pref, host_mem=0, mem_flag=1)
        rec["seq2sdbg_s"] = time.perf_counter() - t0
        t0 = time.perf_counter(); lib.read2sdbg(k=k, min_count=m, output_prefix=pref + "_1pass", num_cpu_threads=8, read_lib_file=libf)
        rec["read2sdbg_s"] = time.perf_counter() - t0
        steps.append({k2: round(v, 3) for k2, v in rec.items()})
    best = steps[-1]
    out["gpu"] = steps
    out["k1_text_GBps"] = text_bytes / best["buildlib_s"] / 1e9
    out["e2e_bases_per_s"] = out["bases"] / (best["buildlib_s"] + best["count_s"] + best["seq2sdbg_s"])
    out["outputs"] = {"edges_bytes": sum(os.path.getsize(os.path.join(work, f)) for f in os.listdir(work) if f.startswith("k21.edges.")),
                      "sdbg_bytes": sum(os.path.getsize(os.path.join(work, f)) for f in os.listdir(work) if f.startswith("k21.sdbg"))}
    # the oracle's CLI on a bounded sample of the same FASTQ (first cpu_pairs pairs)
    if args.cpu_pairs > 0:
        cp = min(args.cpu_pairs, args.pairs)
        c1, c2 = os.path.join(work, "c_1.fq"), os.path.join(work, "c_2.fq")
        write_fastq(c1, c2, bases[:starts[2 * cp]], starts[:2 * cp + 1])
        clib = os.path.join(work, "cpu.lib")
        open(clib, "w").write(f"{c1},{c2}\npe {c1} {c2}\n")
        threads = os.cpu_count() or 1
        rec = {}
        t0 = time.perf_counter(); oracle.cmd_buildlib(clib, clib); rec["buildlib_s"] = time.perf_counter() - t0
        t0 = time.perf_counter(); oracle.cmd_count(clib, k, m, os.path.join(work, "ck21"), threads); rec["count_s"] = time.perf_counter() - t0
        t0 = time.perf_counter(); oracle.cmd_seq2sdbg(k, 0, os.path.join(work, "ck21"), input_prefix=os.path.join(work, "ck21"), threads=threads)
        rec["seq2sdbg_s"] = time.perf_counter() - t0
        out["cpu_oracle"] = {"pairs": cp, "bases": int(starts[2 * cp]), "cores": threads, **{k2: round(v, 3) for k2, v in rec.items()},
                             "e2e_bases_per_s": int(starts[2 * cp]) / sum(rec.values()),
                             "note": "CPU restatement of megahit_core (oracle/), not upstream megahit; first pairs of the same FASTQ (lower depth than the full set)"}
    emit("files", out)


def stand_in_contigs(path, genome_seed, k, n_bases, rng):
    """true-genome unitigs cut at random points, headers as megahit writes them"""
    from oracle import synth_np
    # the synthetic genome is a hash of the position: regenerate a stretch of the nuclear background and the mitogenome
    with np.errstate(over="ignore"):
        pos = np.arange(n_bases, dtype=np.uint64)
        nuc = (synth_np._h3(genome_seed, np.uint64(0x4E55434C), pos) & np.uint64(3)).astype(np.uint8)
        mito = (synth_np._h3(genome_seed, np.uint64(0x4D49544F), np.arange(16500, dtype=np.uint64)) & np.uint64(3)).astype(np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    n = 0
    with open(path, "wb") as f:
        for name, g, depth in (("nuc", nuc, 95.0), ("mito", mito, 15000.0)):
            p = 0
            while p < len(g) - (k + 2):
                L = int(rng.integers(k + 2, 4000))
                s = lut[g[p:p + L]].tobytes()
                f.write(b">k%d_%d flag=1 multi=%.4f len=%d\n%s\n" % (k, n, depth, len(s), s))
                n += 1
                p += L - k   # consecutive contigs overlap by k bases, like unitigs around a branch
    return n


def leg_klist(args):
    import torch
    from mitoflex_b200 import lib
    work = args.work
    os.makedirs(work, exist_ok=True)
    ctx = lib.Context(0)
    ctx.set_profiling(True)
    reads = ctx.synth(n_pairs=args.pairs, seed=1002)
    klist = [int(x) for x in args.klist.split(",")]
    rng = np.random.default_rng(3)
    out = {"leg": "klist", "pairs": args.pairs, "bases": reads.n_bases, "klist": klist, "min_count": args.min_count, "per_k": []}
    prev_k = 0
    for k in klist:
        rec = {"k": k}
        for _ in range(1):
            ctx.read2sdbg(reads, k, args.min_count)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = ctx.read2sdbg(reads, k, args.min_count)
        rec["read2sdbg_ms"] = round(1e3 * (time.perf_counter() - t0), 2)
        rec["sdbg_items"] = g.n
        prof = ctx.last_profile()
        rec["top_stages_ms"] = {k2: round(v, 2) for k2, v in sorted(prof.items(), key=lambda kv: -kv[1])[:5]}
        if prev_k:
            # seq2sdbg at k > k_min from files: iterative edges (stand-in: the solid edges of this k, written unsorted) plus the
            # contigs of the previous k (stand-in: genome unitigs), the flow of assemble_wrapper.py:228-258
            e = ctx.count(reads, k, args.min_count)
            ed = e.to_numpy()
            pref = os.path.join(work, f"{k}")
            perm = rng.permutation(e.n)
            with open(pref + ".edges.0", "wb") as f:
                f.write(ed[perm].tobytes())
            open(pref + ".edges.info", "w").write(f"kmer_size {k}\nwords_per_edge {ed.shape[1]}\nnum_files 1\nnum_buckets 0\nnum_edges {e.n}\nis_sorted 0\n")
            contigs = os.path.join(work, f"k{prev_k}.contigs.fa")
            nc = stand_in_contigs(contigs, 1002, prev_k, args.contig_bases, rng)
            t0 = time.perf_counter()
            lib.seq2sdbg(k=k, kmer_from=prev_k, output_prefix=os.path.join(work, f"g{k}"), num_cpu_threads=8, input_prefix=pref, contig=contigs)
            rec["seq2sdbg_files_s"] = round(time.perf_counter() - t0, 3)
            rec["seq2sdbg_inputs"] = {"unsorted_edges": int(e.n), "contigs": nc, "contig_bases": args.contig_bases}
            del ed
        out["per_k"].append(rec)
        prev_k = k
    out["sum_read2sdbg_ms"] = round(sum(r["read2sdbg_ms"] for r in out["per_k"]), 1)
    out["peak_hbm_gb"], out["hbm_total_gb"] = [round(x, 1) for x in peak_hbm_gb()]
    emit("klist", out)


def leg_big(args):
    import torch
    from mitoflex_b200 import lib
    ctx = lib.Context(0)
    ctx.set_profiling(True)
    t0 = time.perf_counter()
    reads = ctx.synth(n_pairs=args.pairs, nuclear_len=args.nuclear_len, mito_fraction=args.mito_fraction, error_rate=args.error_rate,
                      insert_mean=args.insert_mean, insert_sd=args.insert_sd, seed=args.seed)
    t_gen = time.perf_counter() - t0
    out = {"leg": args.leg, "pairs": args.pairs, "bases": reads.n_bases, "k": args.k, "min_count": args.min_count,
           "error_rate": args.error_rate, "nuclear_len": args.nuclear_len, "synth_s": round(t_gen, 1)}
    times = []
    for rep in range(args.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = ctx.read2sdbg(reads, args.k, args.min_count)
        times.append(time.perf_counter() - t0)
        prof = ctx.last_profile()
    out["read2sdbg_s"] = [round(t, 3) for t in times]
    out["bases_per_s"] = reads.n_bases / min(times)
    out["sdbg_items"] = g.n
    out["stages_ms"] = {k2: round(v, 1) for k2, v in sorted(prof.items(), key=lambda kv: -kv[1])[:12]}
    out["count_rounds"] = sum(1 for k2 in prof if k2 == "reads_scatter") or None
    # size-independent properties (tests/test_gpu_parity.py::test_full_size_properties)
    e = ctx.count(reads, args.k, args.min_count, want_counting=True)
    cnt = e.counting.astype(np.int64)
    idx = np.arange(65536, dtype=np.int64)
    out["n_keys"], out["solid_edges"] = int(e.s.n_keys), int(e.n)
    out["properties"] = {"histogram_accounts_for_every_key": bool(int((cnt[:65535] * idx[:65535]).sum()) == e.s.n_keys or cnt[65535] > 0),
                         "one_edge_per_solid_key": bool(int(cnt[args.min_count:].sum()) == e.n),
                         "bucket_counts_add_up": bool(int(e.bucket_counts().sum()) == e.n),
                         "two_items_per_edge_or_more": bool(g.n >= 2 * e.n)}
    out["peak_hbm_gb"], out["hbm_total_gb"] = [round(x, 1) for x in peak_hbm_gb()]
    emit(args.leg, out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("leg", choices=["files", "klist", "big", "c5"])
    ap.add_argument("--pairs", type=int, default=None)
    ap.add_argument("--work", default="/tmp/mfsdbg_bench")
    ap.add_argument("--cpu-pairs", type=int, default=200_000)
    ap.add_argument("--klist", default="21,29,39,59,79,99,119,141")
    ap.add_argument("--min-count", type=int, default=None)
    ap.add_argument("--contig-bases", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--nuclear-len", type=int, default=None)
    ap.add_argument("--mito-fraction", type=float, default=None)
    ap.add_argument("--error-rate", type=float, default=None)
    ap.add_argument("--insert-mean", type=float, default=None)
    ap.add_argument("--insert-sd", type=float, default=None)
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args()
    dflt = {"files": dict(pairs=2_000_000, min_count=2),
            "klist": dict(pairs=16_666_667, min_count=2),
            "big": dict(pairs=100_000_000, min_count=2, nuclear_len=500_000_000, mito_fraction=0.01, error_rate=0.005, insert_mean=350.0,
                        insert_sd=35.0, seed=1004),
            "c5": dict(pairs=16_666_667, min_count=1, nuclear_len=50_000_000, mito_fraction=0.05, error_rate=0.02, insert_mean=200.0,
                       insert_sd=50.0, seed=1005)}[args.leg]
    for k2, v in dflt.items():
        if getattr(args, k2) is None:
            setattr(args, k2, v)
    {"files": leg_files, "klist": leg_klist, "big": leg_big, "c5": leg_big}[args.leg](args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the sDBG-construction path (BASELINE.json metric).

A "step" is one `read2sdbg -k 21 -m 2` (k-mer count + succinct de Bruijn graph build, fused in HBM) over one
batch of synthetic PE150 reads generated in HBM (SURVEY.md 8d generator: 16.5 kb mitogenome at 5 % of the pairs
over a 50 Mb nuclear background, 0.5 % substitution errors).  At N=1 the batch is BASELINE.json configs[1]'s read
set (16 666 667 pairs = 5.0 Gbp).  With N>1 (torchrun, one rank per GPU) every rank holds its own 5 Gbp shard,
keys are routed to their owner GPU by prefix bucket with an NCCL all-to-all, and value = all bases / max-over-ranks
time (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl reference]

`--impl reference` times the CPU restatement of megahit_core (oracle/, all host threads) on a bounded sample of the
same workload: megahit v1.2.9 itself is not vendored in the reference tree and no binary exists in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bases/s k-mer count+sDBG build (k=21)"
K, MIN_COUNT, READ_LEN = 21, 2, 150
FULL_PAIRS = 16_666_667


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=FULL_PAIRS, help="read pairs per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--k", type=int, default=21, help="k-mer size (the headline metric is quoted at 21)")
    ap.add_argument("--min-count", type=int, default=2, help="solidity threshold -m (headline: 2)")
    ap.add_argument("--error-rate", type=float, default=0.005, help="substitution error rate of the synthetic reads (headline: 0.005)")
    ap.add_argument("--nuclear-len", type=int, default=50_000_000, help="nuclear background length (headline: 50 Mb)")
    ap.add_argument("--no-config3", action="store_true", help="skip the wide-key blocks (k=119, k=141) added to the N=1 line")
    ap.add_argument("--verify-pairs", type=int, default=250_000,
                    help="N>1: read pairs per GPU of the correctness pass run before the timed region (0 = skip)")
    return ap.parse_args()


def workload_name(pairs, error_rate=0.005, nuclear_len=50_000_000):
    headline = error_rate == 0.005 and nuclear_len == 50_000_000
    return (f"read2sdbg k={K} -m {MIN_COUNT} on synthetic {pairs}xPE{READ_LEN} ({2 * pairs * READ_LEN / 1e9:.2f} Gbp/GPU): "
            f"16.5 kb mitogenome at 5% of pairs + {nuclear_len / 1e6:.0f} Mb nuclear background, {100 * error_rate:g}% errors "
            + ("(BASELINE configs[1] read set)" if headline else "(non-headline read set)"))


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.rows = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ roofline bookkeeping
def stage_bytes(n_bases, n_keys, n_edges, n_items_gen, k, n_records_out=0, n_records_in=0):
    """ALGORITHMIC bytes per stage launch (DESIGN.md 'kernels'): what the stage must read + write once."""
    W = 4 * ((2 * (k + 1) + 31) // 32)
    We = 4 * ((2 * (k + 1) + 16 + 31) // 32)
    Wi = 4 * ((2 * k + 20 + 31) // 32)
    return {
        "reads_hist": n_bases / 4,
        "reads_scatter": n_bases / 4 + n_keys * W,
        # super-k-mer exchange (N > 1): reads in, 8-byte records out / records in, keys out
        "skm_scatter": n_bases / 4 + n_records_out * 8,
        "skm_l1_scatter": n_records_in * 8 + n_keys * W,
        "count_l2_hist": n_keys * W,
        "count_l2_scatter": 2 * n_keys * W,
        "count_l2a_hist": n_keys * W,
        "count_l2a_scatter": 2 * n_keys * W,
        "local_count": n_keys * W + n_edges * We,
        # item filter (kmerset.cuh): two passes over the edges, 4 k-mer records per edge written once and read once
        "items_filter": 2 * n_edges * We + 2 * 4 * n_edges * (8 if k <= 31 else 16),
        # generator: edges in, items out
        "items": n_edges * We + n_items_gen * Wi,
        "records_hist": n_items_gen * Wi,
        "records_scatter": 2 * n_items_gen * Wi,
        "local_sdbg": n_items_gen * Wi,
        "sdbg_l1_hist": n_items_gen * Wi,
        "sdbg_l1_scatter": 2 * n_items_gen * Wi,
        "sdbg_l2_hist": n_items_gen * Wi,
        "sdbg_l2_scatter": 2 * n_items_gen * Wi,
    }


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (ValueError, KeyError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ CPU arm
def cpu_run(bases, starts, threads):
    from oracle import oracle
    t0 = time.perf_counter()
    g = oracle.read2sdbg(oracle.Reads(bases, starts), K, MIN_COUNT, threads=threads)
    dt = time.perf_counter() - t0
    return dt, g.n


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def max_rss_gb():
    import resource
    return resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6   # Linux reports KiB


def cpu_sample(full_pairs, sample_pairs, error_rate, nuclear_len, seed=1002):
    """A depth-preserving sample of the workload for the CPU arm, generated on the host by the numpy port of the generator
    (oracle/synth_np.py) -- libmfsdbg.so is never loaded on this path.  The nuclear background is scaled with the pair count
    so that its depth (95x on the headline read set) and with it the solid fraction stay those of the full workload; the
    mitogenome keeps its 16.5 kb."""
    from oracle import synth_np
    scale = sample_pairs / float(full_pairs)
    nuc = max(int(nuclear_len * scale), 20_000)
    b, s = synth_np.synth_reads(sample_pairs, read_len=READ_LEN, nuclear_len=nuc, error_rate=error_rate, seed=seed)
    depth = 0.95 * 2 * READ_LEN * sample_pairs / nuc
    desc = (f"{sample_pairs} pairs ({len(b) / 1e6:.1f} Mbp) of the same generator, nuclear background scaled "
            f"{nuclear_len / 1e6:g} Mb -> {nuc / 1e6:.3f} Mb so its depth stays {depth:.0f}x (same solid fraction as the full workload)")
    return b, s, desc


def sized_cpu_sample(args, target_seconds, threads):
    """pilot on 100 000 pairs, then the sample the oracle finishes in about target_seconds"""
    pilot_pairs = min(100_000, args.pairs)
    b, s, desc = cpu_sample(args.pairs, pilot_pairs, args.error_rate, args.nuclear_len)
    dt, _ = cpu_run(b, s, threads)
    want = int(min(args.pairs, 4_000_000, pilot_pairs * target_seconds / max(dt, 1e-3)))
    if want > pilot_pairs * 1.5:
        b, s, desc = cpu_sample(args.pairs, want, args.error_rate, args.nuclear_len)
        dt = None
    return b, s, desc, dt


def cpu_baseline(args, target_seconds):
    threads = os.cpu_count() or 1
    b, s, desc, dt = sized_cpu_sample(args, target_seconds, threads)
    if dt is None:
        dt, _ = cpu_run(b, s, threads)
    return {"value": len(b) / dt, "unit": "bases/s", "cores": threads, "kind": "port",
            "sample": f"{desc}; oracle read2sdbg (CPU restatement of megahit_core, OpenMP) in {dt:.2f} s; megahit itself is not vendored",
            "cpu_model": cpu_model(), "max_rss_gb": round(max_rss_gb(), 2)}


def run_reference(args):
    """--impl reference: the CPU restatement on host cores, bounded sample per step (rank 0 only).  Loads oracle/ only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = max(1.0, min(6.0, 150.0 / max(args.steps + args.warmup, 1)))
    b, s, desc, _ = sized_cpu_sample(args, per_step, threads)
    for _ in range(args.warmup):
        cpu_run(b, s, threads)
    times = [cpu_run(b, s, threads)[0] for _ in range(args.steps)]
    tot = sum(times)
    v = len(b) * args.steps / tot
    sample = desc + " per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "bases/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.pairs, args.error_rate, args.nuclear_len), "sample": sample},
        "cpu_baseline": {"value": v, "unit": "bases/s", "cores": threads, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model(), "max_rss_gb": round(max_rss_gb(), 2)},
        "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------ N>1 correctness (outside the timed region)
def stream_checksum(values, offset):
    """position-dependent 64-bit checksum of a uint32 stream that starts at global position `offset`: the sum over the
    concatenated rank outputs equals the checksum of the single-GPU stream iff both hold the same values in the same order
    (up to 2^-64 collisions)."""
    v = np.asarray(values).astype(np.uint64).ravel()
    with np.errstate(over="ignore"):
        x = v * np.uint64(0x9E3779B97F4A7C15) + (np.arange(len(v), dtype=np.uint64) + np.uint64(offset)) * np.uint64(0xD1342543DE82EF95)
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
        return int(x.sum(dtype=np.uint64))


def graph_stats(g, item_off, tip_off):
    """what rank-local pieces of a graph contribute to the global figures"""
    rec = g.ctx.d2h(g.s.rec, g.s.n_items * 4, np.uint32)
    lab = g.ctx.d2h(g.s.tip_labels, g.s.n_tips * g.s.words_per_tip * 4, np.uint32)
    return np.array([g.s.n_items, g.s.n_tips, g.s.n_large, int((rec >> 8).astype(np.int64).sum()),
                     stream_checksum(rec, item_off) & 0x7FFFFFFFFFFFFFFF,
                     stream_checksum(lab, tip_off * g.s.words_per_tip) & 0x7FFFFFFFFFFFFFFF], dtype=np.int64)


def verify_multi_gpu(ctx, runner, args, rank, world, dev):
    """Every rank runs the sharded path on a small shard; rank 0 then runs the single-GPU path on the CONCATENATED reads of all
    shards and the global figures must agree: item / tip / large-multiplicity counts, the sum of multiplicities and
    position-dependent checksums of the concatenated record and tip-label streams (rank order = key order)."""
    import torch
    import torch.distributed as dist
    from mitoflex_b200 import lib
    seeds = [424_242 + 7919 * r for r in range(world)]
    kw = dict(n_pairs=args.verify_pairs, error_rate=args.error_rate, nuclear_len=max(args.nuclear_len // 50, 100_000))
    reads = ctx.synth(seed=seeds[rank], **kw)
    res = runner.run(reads)
    g = res.sdbg
    mine = torch.tensor([g.s.n_items, g.s.n_tips, res.info["n_edges"]], dtype=torch.int64, device=dev)
    allc = torch.empty((world, 3), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, mine)
    allc = allc.cpu().numpy()
    st = torch.from_numpy(graph_stats(g, int(allc[:rank, 0].sum()), int(allc[:rank, 1].sum()))).to(dev)
    alls = torch.empty((world, 6), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(alls, st)
    alls = alls.cpu().numpy()
    out = None
    if rank == 0:
        parts, lens = [], []
        for sd in seeds:
            b, s0 = ctx.download_reads(ctx.synth(seed=sd, **kw))
            parts.append(b)
            lens.append(np.diff(s0))
        starts = np.concatenate([[0], np.cumsum(np.concatenate(lens))]).astype(np.int64)
        whole = ctx.upload_reads(np.concatenate(parts), starts)
        e1 = ctx.count(whole, K, MIN_COUNT)
        n_edges_1 = e1.n
        g1 = ctx.read2sdbg(whole, K, MIN_COUNT)
        ref = graph_stats(g1, 0, 0)
        got = alls[:, :4].sum(axis=0).tolist()
        ok = got == ref[:4].tolist() and int(allc[:, 2].sum()) == n_edges_1
        # checksums: every rank's value was masked to 63 bits for the int64 transport, so the reference is built the same
        # way -- piecewise over the ranks' item / tip ranges of the single-GPU streams -- and both are summed modulo 2^63
        M = (1 << 63) - 1
        got_cs = [sum(int(x) for x in alls[:, 4]) & M, sum(int(x) for x in alls[:, 5]) & M]
        rec1 = ctx.d2h(g1.s.rec, g1.s.n_items * 4, np.uint32)
        wt = g1.s.words_per_tip
        lab1 = ctx.d2h(g1.s.tip_labels, g1.s.n_tips * wt * 4, np.uint32)
        ref_cs = [0, 0]
        io = to = 0
        for r in range(world):
            ni, nt = int(allc[r, 0]), int(allc[r, 1])
            ref_cs[0] += stream_checksum(rec1[io:io + ni], io) & M
            ref_cs[1] += stream_checksum(lab1[to * wt:(to + nt) * wt], to * wt) & M
            io += ni
            to += nt
        ok = ok and io == g1.s.n_items and to == g1.s.n_tips and got_cs == [ref_cs[0] & M, ref_cs[1] & M]
        out = {"verified": bool(ok), "pairs_per_gpu": args.verify_pairs, "n_edges": int(allc[:, 2].sum()), "n_items": int(got[0]),
               "n_tips": int(got[1]), "sum_mult": int(got[3]), "checksum_items": f"{got_cs[0]:016x}", "checksum_tips": f"{got_cs[1]:016x}",
               "single_gpu": {"n_edges": int(n_edges_1), "n_items": int(ref[0]), "n_tips": int(ref[1]), "sum_mult": int(ref[3]),
                              "checksum_items": f"{ref_cs[0] & M:016x}", "checksum_tips": f"{ref_cs[1] & M:016x}"},
               "how": "sharded run on all ranks vs single-GPU run of the concatenated reads on rank 0, outside the timed region"}
        del whole
    dist.barrier()
    return out


# ------------------------------------------------------------------ GPU arm
def main():
    global K, MIN_COUNT
    args = parse_args()
    K = args.k
    MIN_COUNT = args.min_count
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from mitoflex_b200 import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ctx = lib.Context(local_rank)
    stream = torch.cuda.Stream(dev)            # a dedicated stream shared with the library, so that torch's
    torch.cuda.set_stream(stream)              # CUDA events bracket exactly the kernels the library launches
    ctx.set_stream(stream.cuda_stream)
    ctx.set_profiling(True)
    verify = None
    runner = None
    if world > 1:
        from mitoflex_b200 import dist as mdist
        runner = mdist.DistRead2Sdbg(ctx, K, MIN_COUNT)
        if args.verify_pairs > 0:
            # correctness of the sharded path, before (and outside) the timed region; the synth buffers are reused below
            verify = verify_multi_gpu(ctx, runner, args, rank, world, dev)
    reads = ctx.synth(n_pairs=args.pairs, seed=1002 + 7919 * rank, error_rate=args.error_rate, nuclear_len=args.nuclear_len)
    n_bases = reads.n_bases
    if world > 1:
        step = lambda: runner.run(reads)   # noqa: E731
    else:
        step = lambda: ctx.read2sdbg(reads, K, MIN_COUNT)   # noqa: E731

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    launches0 = ctx.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    prof = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    stage_sum = {}
    host_sum = {}
    res = None
    for _ in range(args.steps):
        res = step()
        prof = runner.profile if world > 1 else ctx.last_profile()
        for name, ms in prof.items():
            stage_sum[name] = stage_sum.get(name, 0.0) + ms
        for name, ms in (getattr(runner, "host", None) or {}).items():
            host_sum[name] = host_sum.get(name, 0.0) + ms
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - launches0
    # device time of the K steps: CUDA events on the stream the library launches on (set_stream above)
    t_ms = ev0.elapsed_time(ev1)
    wall_ms = t_wall * 1e3
    if world > 1:
        t = torch.tensor([t_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
        nb = torch.tensor([n_bases], device=dev, dtype=torch.int64)
        dist.all_reduce(nb)
        total_bases = int(nb.item())
    else:
        total_bases = n_bases
    value = total_bases * args.steps / (t_ms / 1e3)

    # ---- e2e: host buffers in, host buffers out, through the C ABI (rank-local, N=1 semantics per rank)
    e2e = None
    if not args.no_e2e and world == 1:
        nw = (n_bases + 15) // 16
        hw = torch.empty(nw + 16, dtype=torch.int32, pin_memory=True)
        hs = torch.empty(reads.n_reads + 1, dtype=torch.int64, pin_memory=True)
        # copy the device reads into pinned host buffers once, outside the timed region
        from ctypes import c_void_p
        L = lib.load()
        lib._check(L.mfsdbg_dev_copy(ctx._h, c_void_p(hw.data_ptr()), reads.s.packed, nw * 4, 0))
        lib._check(L.mfsdbg_dev_copy(ctx._h, c_void_p(hs.data_ptr()), reads.s.starts, (reads.n_reads + 1) * 8, 0))
        for _ in range(2):
            out = ctx.host_read2sdbg(hw.data_ptr(), hs.data_ptr(), reads.n_reads, n_bases, K, MIN_COUNT)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = ctx.host_read2sdbg(hw.data_ptr(), hs.data_ptr(), reads.n_reads, n_bases, K, MIN_COUNT)
        te = time.perf_counter() - t0
        assert out.n_items == res.n
        e2e = {"value": n_bases * args.steps / te, "unit": "bases/s", "h2d_bytes_per_step": int(out.h2d_bytes),
               "d2h_bytes_per_step": int(out.d2h_bytes), "ms_per_step": 1e3 * te / args.steps,
               "api": "mfsdbg_host_read2sdbg (pinned host packed reads in, sdbg arrays out)"}
    elif not args.no_e2e:
        # N > 1: every rank's shard starts in pinned host memory and its piece of the graph ends there; the copies are inside
        # the timed region, the time is the wall clock between two barriers, max over ranks
        nw = (n_bases + 15) // 16
        hw = torch.empty(nw + 16, dtype=torch.int32, pin_memory=True)
        hs = torch.empty(reads.n_reads + 1, dtype=torch.int64, pin_memory=True)
        from ctypes import c_void_p
        L = lib.load()
        lib._check(L.mfsdbg_dev_copy(ctx._h, c_void_p(hw.data_ptr()), reads.s.packed, nw * 4, 0))
        lib._check(L.mfsdbg_dev_copy(ctx._h, c_void_p(hs.data_ptr()), reads.s.starts, (reads.n_reads + 1) * 8, 0))
        dw = torch.zeros(nw + 16, dtype=torch.int32, device=dev)
        ds = torch.empty(reads.n_reads + 1, dtype=torch.int64, device=dev)
        cap_items = int(res.n * 1.2) + 1024
        out_rec = torch.empty(cap_items, dtype=torch.int32, pin_memory=True)
        out_lab = torch.empty(int(res.sdbg.s.n_tips * res.sdbg.s.words_per_tip * 1.5) + 1024, dtype=torch.int32, pin_memory=True)

        def e2e_step():
            dw[:nw + 16].copy_(hw, non_blocking=True)
            ds.copy_(hs, non_blocking=True)
            r2 = runner.run(ctx.reads_from_tensors(dw, ds, n_bases))
            g2 = r2.sdbg
            lib._check(L.mfsdbg_dev_copy(ctx._h, c_void_p(out_rec.data_ptr()), g2.s.rec, g2.s.n_items * 4, 0))
            lib._check(L.mfsdbg_dev_copy(ctx._h, c_void_p(out_lab.data_ptr()), g2.s.tip_labels, g2.s.n_tips * g2.s.words_per_tip * 4, 0))
            return g2
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            g2 = e2e_step()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        byts = torch.tensor([nw * 4 + (reads.n_reads + 1) * 8, g2.s.n_items * 4 + g2.s.n_tips * g2.s.words_per_tip * 4], device=dev, dtype=torch.int64)
        dist.all_reduce(byts)
        e2e = {"value": total_bases * n_e2e / float(te.item()), "unit": "bases/s", "h2d_bytes_per_step": int(byts[0].item()),
               "d2h_bytes_per_step": int(byts[1].item()), "ms_per_step": 1e3 * float(te.item()) / n_e2e, "steps": n_e2e,
               "api": "per rank: pinned host packed reads -> HBM, DistRead2Sdbg.run (C-ABI staged calls), sdbg arrays -> pinned host; bytes summed over ranks"}

    # per-rank view of the timed steps (N > 1): device-stage sum and host-clock phases of every rank, so that a slow rank shows
    by_rank = None
    if world > 1:
        mine = {"rank": rank, "stage_ms": round(sum(stage_sum.values()) / args.steps, 3),
                "host_ms": {k2: round(v / args.steps, 3) for k2, v in host_sum.items()},
                "slow_stages": {k2: round(v / args.steps, 3) for k2, v in stage_sum.items()
                                if k2 in ("local_count_multipass", "local_count_general", "oversized", "fallback_local", "sampled_overflow")}}
        by_rank = [None] * world
        dist.all_gather_object(by_rank, mine)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage, from the live CUDA-event stage timings of the timed steps
    n_keys = getattr(res, "n_keys", None)
    info = getattr(res, "info", {})
    e_cnt = ctx.count(reads, K, MIN_COUNT) if world == 1 else None
    n_keys = e_cnt.s.n_keys if e_cnt is not None else info.get("n_keys", 0)
    n_edges = e_cnt.n if e_cnt is not None else info.get("n_edges", 0)
    # generated sdbg items: the filtered generator writes the 2 real items per edge plus the few dummies that survive,
    # which is what the graph ends up holding (res.n); the multi-GPU driver still generates all 6 per edge
    n_items_gen = (res.n if res is not None else 2 * n_edges) if world == 1 else 6 * n_edges
    sb = stage_bytes(n_bases, n_keys, n_edges, n_items_gen, K, info.get("records_sent", 0), info.get("records_recv", 0))
    nvlink = None
    if world > 1:
        # rank 0's share: bytes that left this GPU / time of the all-to-all, against 900 GB/s per direction nominal
        skm = info.get("exchange") == "skm"
        kb = info.get("records_sent", 0) * 8 if skm else info.get("exchanged_keys", 0) * info.get("key_bytes", 8)
        ib = info.get("exchanged_items", 0) * info.get("item_bytes", 8)
        # fused mode ("p2p"): the scatter kernels store straight into the owners' HBM, so the exchange time IS the scatter
        # stage (reads_scatter for keys, records_scatter for items); NCCL mode has separate a2a_* stages
        fused = info.get("exchange") in ("p2p", "skm")
        ak = stage_sum.get("skm_scatter" if skm else ("reads_scatter" if fused else "a2a_keys"), 0.0) / args.steps
        ai = stage_sum.get("records_scatter" if fused else "a2a_items", 0.0) / args.steps
        nvlink = {"exchange": ("super-k-mer records (64-bit: a run of (k+1)-mers sharing a minimizer owner) stored into the owner's HBM by the "
                               "partition kernel over NVLink peer memory (CUDA IPC)") if skm else
                              ("fused partition+exchange kernel over NVLink peer memory (CUDA IPC)" if fused else "NCCL all_to_all_single"),
                  "keys_exchanged": info.get("exchanged_keys"), "keys_per_record": info.get("keys_per_record"),
                  "keys_bytes_sent": kb, "keys_ms": ak, "keys_GBps": kb / ak / 1e6 if ak else None, "items_bytes_sent": ib,
                  "items_ms": ai, "items_GBps": ib / ai / 1e6 if ai else None, "peak_GBps_per_direction": 900.0,
                  "measured_peer_copy_GBps": 770.0}
    per_step = {k2: v / args.steps for k2, v in stage_sum.items()}
    dom = max((k2 for k2 in per_step if k2 in sb), key=lambda k2: per_step[k2], default=None)
    peak, peak_src = measured_peak()
    roofline = None
    if dom:
        ach = sb[dom] / (per_step[dom] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": None, "peak_source": peak_src, "ms_per_launch": per_step[dom],
                    "algorithmic_bytes_per_launch": sb[dom],
                    "stages_ms": {k2: round(v, 3) for k2, v in sorted(per_step.items(), key=lambda kv: -kv[1])}}
        # DRAM bytes of that kernel from the committed `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum),
        # stored per key occurrence because the capture runs a smaller read set; scaled to this launch
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                ent = json.load(open(tr)).get(dom)
                if ent and "dram_bytes_per_key" in ent:
                    roofline["traffic"] = ent["dram_bytes_per_key"] * n_keys
                    roofline["traffic_source"] = ent.get("source")
            except ValueError:
                pass
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(args, args.cpu_seconds)
    # ---- BASELINE configs[2]: the large-k path (multi-word keys) on the same reads, device-resident, driver-timed
    config3 = None
    if world == 1 and not args.no_config3 and K == 21:
        config3 = {}
        for kk in (119, 141):
            try:
                for _ in range(2):
                    ctx.read2sdbg(reads, kk, MIN_COUNT)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                acc = {}
                e0.record()
                for _ in range(2):
                    gk = ctx.read2sdbg(reads, kk, MIN_COUNT)
                    for name, ms in ctx.last_profile().items():
                        acc[name] = acc.get(name, 0.0) + ms / 2
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / 2
                ek = ctx.count(reads, kk, MIN_COUNT)
                sbk = stage_bytes(n_bases, ek.s.n_keys, ek.n, gk.n, kk)
                domk = max((k2 for k2 in acc if k2 in sbk), key=lambda k2: acc[k2], default=None)
                blk = {"ms_per_step": ms, "value": n_bases / (ms / 1e3), "unit": "bases/s", "keys": int(ek.s.n_keys), "solid_edges": int(ek.n),
                       "sdbg_items": int(gk.n), "key_bytes": 4 * ((2 * (kk + 1) + 31) // 32),
                       "stages_ms": {k2: round(v, 3) for k2, v in sorted(acc.items(), key=lambda kv: -kv[1])[:8]}}
                if domk:
                    achk = sbk[domk] / (acc[domk] / 1e3) / 1e9
                    blk["roofline"] = {"bound": "hbm", "kernel": domk, "achieved": achk, "peak": peak, "unit": "GB/s", "frac": achk / peak,
                                       "ms_per_launch": acc[domk], "algorithmic_bytes_per_launch": sbk[domk]}
                config3[f"k{kk}"] = blk
            except lib.MfsdbgError as ex:   # the blocks are extras: never lose the headline line over them
                config3[f"k{kk}"] = {"error": str(ex)}
    out = {
        "metric": METRIC, "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.pairs, args.error_rate, args.nuclear_len), "bases_per_gpu": n_bases, "keys_per_gpu": n_keys, "solid_edges": n_edges,
                   "sdbg_items": res.n if res is not None else None, "l2_flush": "inputs and key buffers (>= 1 GB) exceed the 126 MB L2",
                   "parallelism": ("reads sharded by GPU; count: (k+1)-mers routed to the GPU that owns their minimizer as super-k-mer records, items routed to the "
                                   "owner of their prefix bin; both stored by the partition kernel itself into the owner's HBM (NVLink peer memory); rank r "
                                   "ends with the r-th prefix range of the graph" if info.get("exchange") == "skm" else
                                   "reads sharded by GPU, keys and items stored into the owner GPU of their prefix bin by the partition kernel itself (NVLink peer memory), disjoint key range per GPU") if world > 1 else "1 GPU"},
        "roofline": roofline, "nvlink": nvlink, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if config3 is not None:
        out["config3"] = config3
        out["config"]["config3"] = {k2: {"ms_per_step": v.get("ms_per_step"), "frac": (v.get("roofline") or {}).get("frac")} for k2, v in config3.items()}
    if world > 1:
        out["by_rank"] = by_rank
        out["verified"] = bool(verify and verify.get("verified"))
        out["config"]["verified"] = out["verified"]
        out["config"]["verify"] = verify
        out["config"]["nvlink"] = nvlink
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Multi-GPU driver of the sDBG path: one process per GPU (torch.distributed), reads sharded by GPU, records routed to the
GPU that owns their prefix bin with ONE all-to-all per stage (NCCL over NVLink), then every GPU finishes a disjoint,
contiguous key range on its own.  Concatenating the ranks' outputs in rank order gives the globally sorted streams, which
is exactly how megahit's per-bucket meta files stitch several `.edges.N` / `.sdbg.N` files together.

    count  : prefix histogram (per GPU) -> all_gather -> owners -> local partition -> all_to_all -> count_finish
    sdbg   : items of the local edges -> prefix histogram -> all_gather -> owners -> partition -> all_to_all -> sdbg_finish

The planning functions below are pure numpy (tested with gloo on CPU); only DistRead2Sdbg touches the GPU.
"""
import numpy as np

import os

L1_BITS = int(os.environ.get("MFSDBG_DIST_L1_BITS", "0"))   # 0 = by world size (l1_bits_for)


def l1_bits_for(world):
    """Exchange bins = ownership granularity.  64 bins keep the runs of the NVLink stores long (5-7 bits measured within 3 % on
    the fused kernel at 2 GPUs).  Owners are cut at bin boundaries and canonical keys are twice as dense at small prefixes, so
    with 8 ranks on 64 bins rank 0 holds +20 % of the mean (SCALE_r01) -- but finer bins are not the cure: 256 bins at 8 GPUs
    took the fused exchange kernel from 50 to 91 ms (326 instead of 576 GB/s out of each GPU: runs of 32 keys) and the step
    from 192 to 531 ms (gpurun_out/r2p_bench_n8*.json).  So: 64 bins whatever the world size."""
    return L1_BITS or 6


def assign_owners(global_hist, world):
    """Contiguous bin ranges [lo_r, hi_r) per rank, balanced on the global histogram. Returns int64 array [world + 1]."""
    h = np.asarray(global_hist, dtype=np.int64)
    nb = len(h)
    csum = np.concatenate([[0], np.cumsum(h)])
    total = int(csum[-1])
    bounds = np.zeros(world + 1, dtype=np.int64)
    bounds[world] = nb
    for r in range(1, world):
        target = total * r // world
        # first bin boundary whose cumulative count reaches the target
        b = int(np.searchsorted(csum, target, side="left"))
        bounds[r] = min(max(b, bounds[r - 1]), nb)
    return bounds


def exchange_plan(all_hists, rank):
    """all_hists: [world, nbins] per-source-rank bin counts. Returns dict with the owner bounds, this rank's send / recv
    splits (records) and the chunk table (start, size, seg) describing the receive buffer: source-major, bins ascending."""
    H = np.asarray(all_hists, dtype=np.int64)
    world, nb = H.shape
    bounds = assign_owners(H.sum(axis=0), world)
    send = np.array([H[rank, bounds[r]:bounds[r + 1]].sum() for r in range(world)], dtype=np.int64)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    recv = np.array([H[s, lo:hi].sum() for s in range(world)], dtype=np.int64)
    starts, sizes, segs = [], [], []
    off = 0
    for s in range(world):
        sz = H[s, lo:hi]
        st = off + np.concatenate([[0], np.cumsum(sz)[:-1]]) if hi > lo else np.zeros(0, np.int64)
        nz = sz > 0
        starts.append(st[nz])
        sizes.append(sz[nz])
        segs.append(np.arange(hi - lo, dtype=np.int32)[nz])
        off += int(sz.sum())
    return dict(bounds=bounds, lo=lo, hi=hi, n_segs=max(hi - lo, 1), send=send, recv=recv,
                chunk_start=np.concatenate(starts).astype(np.int64) if starts else np.zeros(0, np.int64),
                chunk_size=np.concatenate(sizes).astype(np.int64) if sizes else np.zeros(0, np.int64),
                chunk_seg=np.concatenate(segs).astype(np.int32) if segs else np.zeros(0, np.int32),
                n_recv=int(recv.sum()))


def peer_bin_bases(all_hists, bounds, rank, peer_ptrs, rec_bytes):
    """Byte address at which this rank's records of every prefix bin must land: bin b belongs to rank r(b), whose receive
    buffer is laid out source-major with bins ascending (the same layout exchange_plan's chunk table describes)."""
    H = np.asarray(all_hists, dtype=np.int64)
    world, nb = H.shape
    out = np.zeros(nb, dtype=np.uint64)
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        if hi <= lo:
            continue
        base = int(H[:rank, lo:hi].sum())                       # lower-ranked sources come first in r's buffer
        within = np.concatenate([[0], np.cumsum(H[rank, lo:hi])[:-1]])
        out[lo:hi] = np.uint64(peer_ptrs[r]) + ((base + within) * rec_bytes).astype(np.uint64)
    return out


def ks_slice_bounds(nslices, world):
    """Hash slices of the k-mer set dealt out in contiguous ranges: rank r owns slices [b[r], b[r + 1])."""
    return np.array([(r * nslices) // world for r in range(world + 1)], dtype=np.int64)


def ks_exchange_plan(all_hists, rank):
    """all_hists: [world, 2 * nslices] k-mer records per (source rank, kind * nslices + slice), kind 0 = inserts, 1 = queries.
    The owner of a slice receives its inserts first, then its queries, each slice-major and source-minor (a slice's records
    are contiguous whatever their source, which is what the owner's insert / query walks want).  Returns the record offset of
    every bin of THIS source inside its owner's buffer, the owner of every bin, and what this rank receives."""
    H = np.asarray(all_hists, dtype=np.int64)
    world, nbins = H.shape
    nsl = nbins // 2
    b = ks_slice_bounds(nsl, world)
    owner_of_slice = np.repeat(np.arange(world), np.diff(b))
    col = H.sum(axis=0)                                  # records of every bin over all sources
    before = np.cumsum(H, axis=0) - H                    # records of lower-ranked sources in the same bin
    off = np.zeros(nbins, dtype=np.int64)
    n_ins = np.zeros(world, dtype=np.int64)
    n_qry = np.zeros(world, dtype=np.int64)
    for r in range(world):
        lo, hi = int(b[r]), int(b[r + 1])
        ins, qry = col[lo:hi], col[nsl + lo:nsl + hi]
        n_ins[r], n_qry[r] = ins.sum(), qry.sum()
        off[lo:hi] = np.cumsum(ins) - ins + before[rank, lo:hi]
        off[nsl + lo:nsl + hi] = n_ins[r] + np.cumsum(qry) - qry + before[rank, nsl + lo:nsl + hi]
    return dict(offset=off, owner=np.concatenate([owner_of_slice, owner_of_slice]), slice_lo=int(b[rank]),
                n_owned=int(b[rank + 1] - b[rank]), n_ins=int(n_ins[rank]), n_qry=int(n_qry[rank]),
                recv_records=n_ins + n_qry)


class PeerBuffer:
    """A receive buffer in this GPU's HBM that every other rank maps through CUDA IPC (NVLink peer memory)."""

    def __init__(self, ctx, dev):
        self.ctx, self.dev, self.ptr, self.cap, self.peers = ctx, dev, None, 0, None

    def ensure(self, nbytes):
        """collective: grow (and re-share) if ANY rank needs more room"""
        import torch
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        flag = torch.tensor([int(nbytes > self.cap or self.ptr is None)], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) == 0:
            return
        self.release()
        self.cap = max(int(nbytes * 1.2) + 4096, self.cap)
        self.ptr = self.ctx.dev_alloc(self.cap)
        mine = torch.from_numpy(self.ctx.ipc_export(self.ptr)).to(self.dev)
        allh = torch.empty((world, 64), dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, mine)
        handles = allh.cpu().numpy()
        self.peers = [self.ptr if r == rank else self.ctx.ipc_open(handles[r]) for r in range(world)]

    def release(self):
        import torch.distributed as dist
        if self.ptr is None:
            return
        rank = dist.get_rank()
        for r, p in enumerate(self.peers or []):
            if r != rank:
                self.ctx.ipc_close(p)
        dist.barrier()               # nobody maps the old buffer any more
        self.ctx.dev_free(self.ptr)
        self.ptr, self.peers = None, None


def all_to_all_records(send_rows, send_splits, recv_splits, group=None):
    """send_rows: [n, W] tensor partitioned by destination rank (splits in rows). Returns the received [m, W] tensor."""
    import torch
    import torch.distributed as dist
    out = torch.empty((int(sum(recv_splits)), send_rows.shape[1]), dtype=send_rows.dtype, device=send_rows.device)
    dist.all_to_all_single(out, send_rows, output_split_sizes=[int(x) for x in recv_splits],
                           input_split_sizes=[int(x) for x in send_splits], group=group)
    return out


class DistResult:
    def __init__(self, sdbg, info):
        self.sdbg, self.info = sdbg, info

    @property
    def n(self):
        return self.sdbg.n


class DistRead2Sdbg:
    """read2sdbg over all ranks of the default process group; each rank ends with its prefix range of the graph."""

    def __init__(self, ctx, k, min_count, exchange=None, skm=None, item_filter=None):
        """exchange: "p2p" (records stored straight into the owner's HBM by the partition kernels) or "nccl" (all_to_all_single).
        skm: route the count by minimizer and send super-k-mer records (csrc/skm.cu) instead of one key per (k+1)-mer; None =
        whenever the library supports it for this k (16 <= k <= 26) in p2p mode, MFSDBG_DIST_SKM=0 switches it off."""
        import torch
        self.ctx, self.k, self.m = ctx, k, min_count
        self.dev = torch.device("cuda", ctx.device)
        from . import lib
        self.Wk = lib.load().mfsdbg_words_per_key(k)
        self.We = lib.load().mfsdbg_words_per_edge(k)
        self.Wi = lib.load().mfsdbg_words_per_item(k)
        self.profile = {}
        import os
        self.mode = exchange or os.environ.get("MFSDBG_EXCHANGE", "p2p")
        self.key_buf = PeerBuffer(ctx, self.dev)
        self.item_buf = PeerBuffer(ctx, self.dev)
        if skm is None:
            skm = os.environ.get("MFSDBG_DIST_SKM", "1") != "0"
        self.skm = bool(skm) and self.mode == "p2p" and bool(lib.load().mfsdbg_skm_supported(k))
        self.skm_cap = None      # records one (source, destination) region of the receive buffers holds
        # the item filter across GPUs (csrc/ksdist.cu): only the dummies that reach the graph are generated and exchanged
        self.filter = (self.mode == "p2p" and bool(lib.load().mfsdbg_ks_supported(k))
                       and os.environ.get("MFSDBG_DIST_FILTER", "1") != "0") if item_filter is None else bool(item_filter)
        self.ks_buf = PeerBuffer(ctx, self.dev)

    def _acc(self):
        for name, ms in self.ctx.last_profile().items():
            self.profile[name] = self.profile.get(name, 0.0) + ms

    def _mark(self, name):
        """host clock since the previous mark -> self.host[name] (what a phase costs on the wall, kernels + planning + waiting)"""
        import time
        now = time.perf_counter()
        self.host[name] = self.host.get(name, 0.0) + (now - self._t_mark) * 1e3
        self._t_mark = now

    def _timed_a2a(self, name, rows, send, recv):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = all_to_all_records(rows, send, recv)
        e1.record()
        e1.synchronize()
        self.profile[name] = self.profile.get(name, 0.0) + e0.elapsed_time(e1)
        return out

    def _gather_hists(self, hist):
        import torch
        import torch.distributed as dist
        world = dist.get_world_size()
        allh = torch.empty((world, hist.numel()), dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(allh, hist)
        self._last_hists = allh.cpu().numpy()
        return self._last_hists

    def _count_skm(self, reads):
        """count with the super-k-mer exchange: every rank cuts its reads into runs of (k+1)-mers that share a minimizer owner
        and stores them as 64-bit records into region [rank] of the owner's receive buffer (NVLink peer memory); the owner then
        counts the keys of the records it received.  The counts all-gather doubles as the "stores have landed" barrier."""
        import torch
        import torch.distributed as dist
        from . import lib
        ctx, k = self.ctx, self.k
        world, rank = dist.get_world_size(), dist.get_rank()
        stream = torch.cuda.current_stream(self.dev)
        if self.skm_cap is None:
            samp, _ = ctx.skm_scatter(reads, k, world, stride=64)      # every 64th tile, count only
            self._acc()
            est = torch.tensor([int(samp.max()) * 64 * 1.05 + 65536], dtype=torch.float64, device=self.dev)
            dist.all_reduce(est, op=dist.ReduceOp.MAX)
            self.skm_cap = int(est.item())
        while True:
            cap = self.skm_cap
            self._mark("h_other")
            self.key_buf.ensure(world * cap * 8)
            self._mark("h_ensure_keys")
            dst = np.array([int(self.key_buf.peers[d]) + rank * cap * 8 for d in range(world)], dtype=np.uint64)
            stream.synchronize()
            rec, keys = ctx.skm_scatter(reads, k, world, dst, np.full(world, cap, np.int64))
            self._acc()
            self._mark("h_skm_scatter")
            mine = torch.from_numpy(np.concatenate([rec, keys])).to(self.dev)
            allc = torch.empty((world, 2 * world), dtype=torch.int64, device=self.dev)
            dist.all_gather_into_tensor(allc, mine)
            C = allc.cpu().numpy()
            self._mark("h_gather_counts")       # includes waiting for the slowest sender
            if int(C[:, :world].max()) <= cap:
                break
            self.skm_cap = int(C[:, :world].max() * 1.05) + 65536   # a region overflowed somewhere: everybody retries with more room
        chunk_start = np.arange(world, dtype=np.int64) * cap
        chunk_size = C[:, rank].astype(np.int64)
        n_keys = int(C[:, world + rank].sum())
        kc = int(lib.load().mfsdbg_skm_key_capacity(n_keys))
        keys_a = torch.empty((kc + 32, 2), dtype=torch.int32, device=self.dev)
        keys_b = torch.empty((kc + 32, 2), dtype=torch.int32, device=self.dev)
        stream.synchronize()
        self._mark("h_alloc_keys")
        edges = ctx.count_skm(self.key_buf.ptr, chunk_start, chunk_size, n_keys, k, self.m, keys_a.data_ptr(), keys_b.data_ptr(), kc)
        self._acc()
        self._mark("h_count_skm")
        del keys_a, keys_b
        sent = int(rec.sum() - rec[rank])
        info = dict(n_keys=n_keys, n_edges=edges.n, exchanged_keys=int(keys.sum() - keys[rank]), key_bytes=4 * self.Wk, exchange="skm",
                    records_sent=sent, records_recv=int(chunk_size.sum()), record_bytes=8,
                    keys_per_record=float(keys.sum() / max(int(rec.sum()), 1)))
        return edges, info

    def run(self, reads):
        import torch
        import torch.distributed as dist
        L1 = l1_bits_for(dist.get_world_size())
        ctx, k, nb = self.ctx, self.k, 1 << L1
        rank = dist.get_rank()
        stream = torch.cuda.current_stream(self.dev)
        stream.synchronize()
        self.profile = {}
        import time
        self.host, self._t_mark = {}, time.perf_counter()
        if self.skm:
            edges, info = self._count_skm(reads)
            return self._sdbg(edges, info, L1)
        # ---- count: histogram, owners, local partition, exchange, finish
        hist = torch.zeros(nb, dtype=torch.int64, device=self.dev)
        stream.synchronize()
        ctx.count_hist(reads, k, L1, hist.data_ptr())
        self._acc()
        plan = exchange_plan(self._gather_hists(hist), rank)
        allH = self._last_hists
        n_local = int(plan["send"].sum())
        if self.mode == "p2p":
            # fused partition + exchange: the scatter kernel stores every key straight into its owner's HBM over NVLink
            self.key_buf.ensure(max(plan["n_recv"], 1) * self.Wk * 4)
            bases = torch.from_numpy(peer_bin_bases(allH, plan["bounds"], rank, self.key_buf.peers, self.Wk * 4).view(np.int64)).to(self.dev)
            scratch = torch.empty((max(plan["n_recv"], 1) + 16, self.Wk), dtype=torch.int32, device=self.dev)
            stream.synchronize()
            ctx.count_scatter_peer(reads, k, L1, bases.data_ptr())
            self._acc()
            dist.barrier()            # every rank's stores have landed
            edges = ctx.count_finish(self.key_buf.ptr, scratch.data_ptr(), plan["n_recv"], plan["chunk_start"], plan["chunk_size"],
                                     plan["chunk_seg"], plan["n_segs"], k, L1, self.m)
            self._acc()
            del scratch
        else:
            n_buf = max(n_local, plan["n_recv"], 1)
            send = torch.empty((n_buf, self.Wk), dtype=torch.int32, device=self.dev)
            stream.synchronize()
            ctx.count_scatter(reads, k, L1, hist.data_ptr(), send.data_ptr(), n_buf)
            self._acc()
            recv = self._timed_a2a("a2a_keys", send[:n_local], plan["send"], plan["recv"])
            stream.synchronize()
            edges = ctx.count_finish(recv.data_ptr(), send.data_ptr(), plan["n_recv"], plan["chunk_start"], plan["chunk_size"],
                                     plan["chunk_seg"], plan["n_segs"], k, L1, self.m)
            self._acc()
            del recv, send
        info = dict(n_keys=plan["n_recv"], n_edges=edges.n, exchanged_keys=n_local - int(plan["send"][rank]),
                    key_bytes=4 * self.Wk, exchange=self.mode)
        return self._sdbg(edges, info, L1)

    def _filtered_items(self, edges, info):
        """The item filter across GPUs: every rank cuts its edges into k-mer records (2 inserts + 2 queries each) and stores
        them into the buffer of the rank that owns their hash slice; the owners probe their slices; every rank then generates
        2 real items per local edge and 2 dummies per miss it found."""
        import torch
        import torch.distributed as dist
        import ctypes as C
        from . import lib
        ctx, k = self.ctx, self.k
        world, rank = dist.get_world_size(), dist.get_rank()
        stream = torch.cuda.current_stream(self.dev)
        tot = torch.tensor([edges.n], dtype=torch.int64, device=self.dev)
        dist.all_reduce(tot)
        ls, sl = C.c_int32(0), C.c_int32(0)
        lib._check(lib.load().mfsdbg_ks_geometry(int(tot.item()), world, C.byref(ls), C.byref(sl)))
        ls, sl = ls.value, sl.value
        nbins = 2 << (ls - sl)
        hist = torch.zeros(nbins, dtype=torch.int64, device=self.dev)
        stream.synchronize()
        ctx.ks_hist(edges.s.edges, edges.n, k, ls, sl, hist.data_ptr())
        self._acc()
        plan = ks_exchange_plan(self._gather_hists(hist), rank)
        self.ks_buf.ensure(max(int(plan["recv_records"][rank]), 1) * 8)
        peers = np.array([int(p) for p in self.ks_buf.peers], dtype=np.uint64)
        bases = torch.from_numpy((peers[plan["owner"]] + (plan["offset"] * 8).astype(np.uint64)).view(np.int64)).to(self.dev)
        stream.synchronize()
        ctx.ks_scatter_peer(edges.s.edges, edges.n, k, ls, sl, bases.data_ptr())
        self._acc()
        dist.barrier()            # every rank's records have landed
        n_miss = ctx.ks_filter(self.ks_buf.ptr, plan["n_ins"], self.ks_buf.ptr + plan["n_ins"] * 8, plan["n_qry"], ls, sl,
                               plan["slice_lo"], plan["n_owned"])
        self._acc()
        n_items = 2 * edges.n + 2 * n_miss
        items = torch.empty((max(n_items, 1), self.Wi), dtype=torch.int32, device=self.dev)
        stream.synchronize()
        got = ctx.ks_items(edges.s.edges, edges.n, n_miss, k, items.data_ptr(), max(n_items, 1))
        self._acc()
        assert got == n_items
        info.update(item_filter=True, misses=int(n_miss), kmer_records_sent=int(4 * edges.n - self._last_hists[rank][plan["owner"] == rank].sum()))
        return items, n_items

    def _sdbg(self, edges, info, L1):
        """items of the local edges, exchange by item prefix, finish: rank r ends with the r-th prefix range of the graph (the
        edges may be any disjoint split of the solid edge set -- a key range or the minimizer-owned subsets of the skm count)"""
        import torch
        import torch.distributed as dist
        ctx, k, nb = self.ctx, self.k, 1 << L1
        rank = dist.get_rank()
        stream = torch.cuda.current_stream(self.dev)
        if self.filter:
            items, n_items = self._filtered_items(edges, info)
        else:
            n_items = 6 * edges.n
            items = torch.empty((max(n_items, 1), self.Wi), dtype=torch.int32, device=self.dev)
            stream.synchronize()
            ctx.sdbg_items(edges.s.edges, edges.n, k, items.data_ptr())
            self._acc()
        self._mark("h_items")
        ihist = torch.zeros(nb, dtype=torch.int64, device=self.dev)
        stream.synchronize()
        ctx.records_hist(items.data_ptr(), n_items, self.Wi, L1, ihist.data_ptr())
        self._acc()
        self._mark("h_items_hist")
        iplan = exchange_plan(self._gather_hists(ihist), rank)
        self._mark("h_gather_item_hists")
        iH = self._last_hists
        if self.mode == "p2p":
            self.item_buf.ensure(max(iplan["n_recv"], 1) * self.Wi * 4)
            ibases = torch.from_numpy(peer_bin_bases(iH, iplan["bounds"], rank, self.item_buf.peers, self.Wi * 4).view(np.int64)).to(self.dev)
            stream.synchronize()
            self._mark("h_ensure_items")
            ctx.records_scatter_peer(items.data_ptr(), n_items, self.Wi, L1, ibases.data_ptr())
            self._acc()
            self._mark("h_items_scatter")
            dist.barrier()
            self._mark("h_items_barrier")
            del items
            iscratch = torch.empty((max(iplan["n_recv"], 1) + 16, self.Wi), dtype=torch.int32, device=self.dev)
            stream.synchronize()
            g = ctx.sdbg_finish(self.item_buf.ptr, iscratch.data_ptr(), iplan["n_recv"], iplan["chunk_start"], iplan["chunk_size"],
                                iplan["chunk_seg"], iplan["n_segs"], k, L1, 1)
            self._acc()
            self._mark("h_sdbg_finish")
            del iscratch
        else:
            ibuf = max(n_items, iplan["n_recv"], 1)
            isend = torch.empty((ibuf, self.Wi), dtype=torch.int32, device=self.dev)
            stream.synchronize()
            ctx.records_scatter(items.data_ptr(), n_items, self.Wi, L1, ihist.data_ptr(), isend.data_ptr())
            self._acc()
            del items
            irecv = self._timed_a2a("a2a_items", isend[:n_items], iplan["send"], iplan["recv"])
            stream.synchronize()
            g = ctx.sdbg_finish(irecv.data_ptr(), isend.data_ptr(), iplan["n_recv"], iplan["chunk_start"], iplan["chunk_size"],
                                iplan["chunk_seg"], iplan["n_segs"], k, L1, 1)
            self._acc()
            del irecv, isend
        info.update(n_items=iplan["n_recv"], exchanged_items=n_items - int(iplan["send"][rank]), sdbg_items=g.n,
                    item_bytes=4 * self.Wi)
        return DistResult(g, info)

    def close(self):
        self.key_buf.release()
        self.item_buf.release()
        self.ks_buf.release()

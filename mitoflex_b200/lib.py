"""ctypes binding of libmfsdbg.so (include/mfsdbg.h) -- the only way Python reaches the CUDA path.

The library is built in-tree by `make -C mitoflex_b200/csrc` (or `__graft_entry__.build()`); importing this module
never builds or falls back to anything: a missing library or a missing GPU is a loud error.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmfsdbg.so")

OK, EINVAL, EIO, ENODEV, ECUDA, ENOMEM, EINTERNAL = 0, -1, -2, -3, -4, -5, -6
N_MEGAHIT, N_SPLIT = 0, 1


class MfsdbgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmfsdbg error {code}: {msg}")
        self.code = code


class Opts(C.Structure):
    _fields_ = [("k", C.c_int32), ("kmer_from", C.c_int32), ("min_count", C.c_int32), ("mem_flag", C.c_int32),
                ("num_cpu_threads", C.c_int32), ("need_mercy", C.c_int32), ("host_mem", C.c_int64), ("n_gpus", C.c_int32),
                ("gpu_ids", C.POINTER(C.c_int32)), ("n_policy", C.c_int32), ("read_lib_file", C.c_char_p),
                ("input_prefix", C.c_char_p), ("output_prefix", C.c_char_p), ("contig", C.c_char_p), ("bubble", C.c_char_p),
                ("addi_contig", C.c_char_p), ("local_contig", C.c_char_p)]


class DevReads(C.Structure):
    _fields_ = [("packed", C.c_void_p), ("starts", C.c_void_p), ("n_reads", C.c_int64), ("n_bases", C.c_int64)]


class DevEdges(C.Structure):
    _fields_ = [("edges", C.c_void_p), ("n_edges", C.c_int64), ("k", C.c_int32), ("words_per_edge", C.c_int32),
                ("n_keys", C.c_int64)]


class DevSdbg(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("tip_labels", C.c_void_p), ("bucket_items", C.c_void_p), ("n_items", C.c_int64),
                ("n_tips", C.c_int64), ("n_large", C.c_int64), ("k", C.c_int32), ("words_per_tip", C.c_int32)]


class HostSdbg(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("tip_labels", C.c_void_p), ("large_index", C.c_void_p), ("large_mult", C.c_void_p),
                ("n_items", C.c_int64), ("n_tips", C.c_int64), ("n_large", C.c_int64), ("k", C.c_int32),
                ("words_per_tip", C.c_int32), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


class SynthSpec(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("read_len", C.c_int32), ("mito_len", C.c_int64), ("nuclear_len", C.c_int64),
                ("mito_fraction", C.c_double), ("error_rate", C.c_double), ("n_rate", C.c_double),
                ("insert_mean", C.c_double), ("insert_sd", C.c_double), ("seed", C.c_uint64)]


# every symbol include/mfsdbg.h declares (tests check the library exports exactly these)
SYMBOLS = {
    "mfsdbg_version": (C.c_int, []),
    "mfsdbg_device_count": (C.c_int, []),
    "mfsdbg_last_error": (C.c_char_p, []),
    "mfsdbg_buildlib": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int32]),
    "mfsdbg_count": (C.c_int, [C.POINTER(Opts)]),
    "mfsdbg_seq2sdbg": (C.c_int, [C.POINTER(Opts)]),
    "mfsdbg_read2sdbg": (C.c_int, [C.POINTER(Opts)]),
    "mfsdbg_ctx_create": (C.c_void_p, [C.c_int32]),
    "mfsdbg_ctx_destroy": (None, [C.c_void_p]),
    "mfsdbg_ctx_set_mem_limit": (C.c_int, [C.c_void_p, C.c_uint64]),
    "mfsdbg_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "mfsdbg_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfsdbg_ctx_launches": (C.c_int64, [C.c_void_p]),
    "mfsdbg_ctx_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "mfsdbg_ctx_last_profile": (C.c_char_p, [C.c_void_p]),
    "mfsdbg_dev_count": (C.c_int, [C.c_void_p, C.POINTER(DevReads), C.c_int32, C.c_int32, C.POINTER(DevEdges), C.c_void_p]),
    "mfsdbg_dev_seq2sdbg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(DevSdbg)]),
    "mfsdbg_dev_read2sdbg": (C.c_int, [C.c_void_p, C.POINTER(DevReads), C.c_int32, C.c_int32, C.POINTER(DevSdbg)]),
    "mfsdbg_dev_pack_fastq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(DevReads), C.POINTER(C.c_int32)]),
    "mfsdbg_dev_count_hist": (C.c_int, [C.c_void_p, C.POINTER(DevReads), C.c_int32, C.c_int32, C.c_void_p]),
    "mfsdbg_dev_count_scatter": (C.c_int, [C.c_void_p, C.POINTER(DevReads), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]),
    "mfsdbg_dev_count_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(DevEdges), C.c_void_p]),
    "mfsdbg_dev_sdbg_items": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "mfsdbg_dev_records_hist": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "mfsdbg_dev_records_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mfsdbg_dev_sdbg_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(DevSdbg)]),
    "mfsdbg_words_per_item": (C.c_int32, [C.c_int32]),
    "mfsdbg_dev_alloc": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "mfsdbg_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfsdbg_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mfsdbg_ipc_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mfsdbg_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfsdbg_dev_count_scatter_peer": (C.c_int, [C.c_void_p, C.POINTER(DevReads), C.c_int32, C.c_int32, C.c_void_p]),
    "mfsdbg_dev_records_scatter_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "mfsdbg_skm_supported": (C.c_int32, [C.c_int32]),
    "mfsdbg_skm_key_capacity": (C.c_int64, [C.c_int64]),
    "mfsdbg_dev_skm_scatter": (C.c_int, [C.c_void_p, C.POINTER(DevReads), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p]),
    "mfsdbg_dev_count_skm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(DevEdges)]),
    "mfsdbg_ks_supported": (C.c_int32, [C.c_int32]),
    "mfsdbg_ks_geometry": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mfsdbg_dev_ks_hist": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mfsdbg_dev_ks_scatter_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mfsdbg_dev_ks_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.POINTER(C.c_int64)]),
    "mfsdbg_dev_ks_items": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "mfsdbg_words_per_key": (C.c_int32, [C.c_int32]),
    "mfsdbg_words_per_edge": (C.c_int32, [C.c_int32]),
    "mfsdbg_host_read2sdbg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                        C.POINTER(HostSdbg)]),
    "mfsdbg_dev_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32]),
    "mfsdbg_ctx_edge_bucket_counts": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfsdbg_ctx_sdbg_bucket_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfsdbg_dev_synth_reads": (C.c_int, [C.c_void_p, C.POINTER(SynthSpec), C.POINTER(DevReads)]),
}

_lib = None


def load():
    """dlopen libmfsdbg.so. Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}`; "
                                    "there is no CPU fallback for the sDBG path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc):
    if rc != OK:
        raise MfsdbgError(rc, load().mfsdbg_last_error().decode(errors="replace"))


def _b(s):
    return None if s is None else (s.encode() if isinstance(s, str) else s)


def device_count():
    return load().mfsdbg_device_count()


# ------------------------------------------------------------------ file-level (megahit_core sub-commands)
def _opts(**kw):
    o = Opts()
    keep = []
    for name, _ in Opts._fields_:
        if name in kw and kw[name] is not None:
            v = kw[name]
            if isinstance(v, str):
                v = v.encode()
                keep.append(v)
            setattr(o, name, v)
    o._keep = keep
    return o


def buildlib(lib_file, out_prefix, n_policy=N_MEGAHIT):
    _check(load().mfsdbg_buildlib(_b(lib_file), _b(out_prefix), n_policy))


def count(**kw):
    _check(load().mfsdbg_count(C.byref(_opts(**kw))))


def seq2sdbg(**kw):
    _check(load().mfsdbg_seq2sdbg(C.byref(_opts(**kw))))


def read2sdbg(**kw):
    _check(load().mfsdbg_read2sdbg(C.byref(_opts(**kw))))


# ------------------------------------------------------------------ host helpers
def pack_reads(bases, starts):
    """uint8 bases (0..3, back to back) -> uint32 words, 16 bases per word MSB-first, plus 16 words of padding."""
    bases = np.asarray(bases, dtype=np.uint8)
    n = len(bases)
    nw = (n + 15) // 16
    pad = np.zeros(nw * 16, dtype=np.uint32)
    pad[:n] = bases
    sh = (30 - 2 * np.arange(16, dtype=np.uint32))
    words = (pad.reshape(nw, 16) << sh).sum(axis=1, dtype=np.uint64).astype(np.uint32) if nw else np.zeros(0, np.uint32)
    return np.concatenate([words, np.zeros(16, np.uint32)]), np.asarray(starts, dtype=np.int64)


def unpack_reads(words, n_bases):
    words = np.asarray(words, dtype=np.uint32)
    idx = np.arange(n_bases, dtype=np.int64)
    return ((words[idx >> 4] >> (30 - 2 * (idx & 15)).astype(np.uint32)) & 3).astype(np.uint8)


class Reads:
    """Packed reads resident in HBM (torch tensors own the memory, or the context does for synth/pack results)."""

    def __init__(self, struct, keep=()):
        self.s = struct
        self._keep = keep

    @property
    def n_reads(self):
        return self.s.n_reads

    @property
    def n_bases(self):
        return self.s.n_bases


class Context:
    """One GPU. Wraps mfsdbg_ctx; every method is one C-ABI call."""

    def __init__(self, device=0):
        L = load()
        self._h = L.mfsdbg_ctx_create(device)
        if not self._h:
            raise MfsdbgError(ENODEV, L.mfsdbg_last_error().decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load().mfsdbg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    # -- plumbing
    def set_mem_limit(self, nbytes):
        _check(load().mfsdbg_ctx_set_mem_limit(self._h, int(nbytes)))

    def set_stream(self, cuda_stream):
        _check(load().mfsdbg_ctx_set_stream(self._h, cuda_stream))

    def set_profiling(self, on=True):
        _check(load().mfsdbg_ctx_set_profiling(self._h, int(on)))

    def last_profile(self):
        out = {}
        for item in load().mfsdbg_ctx_last_profile(self._h).decode().split(";"):
            if "=" in item:
                name, ms = item.split("=")
                out[name] = out.get(name, 0.0) + float(ms)
        return out

    @property
    def launches(self):
        return load().mfsdbg_ctx_launches(self._h)

    @property
    def stream(self):
        return load().mfsdbg_ctx_stream(self._h)

    def d2h(self, ptr, nbytes, dtype):
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        if nbytes:
            _check(load().mfsdbg_dev_copy(self._h, out.ctypes.data, ptr, nbytes, 0))
        return out

    def h2d(self, dst_ptr, arr):
        arr = np.ascontiguousarray(arr)
        _check(load().mfsdbg_dev_copy(self._h, dst_ptr, arr.ctypes.data, arr.nbytes, 1))

    # -- inputs
    def upload_reads(self, bases, starts):
        """numpy bases (uint8 0..3) + starts -> Reads in HBM (memory owned by torch tensors)."""
        import torch
        words, starts = pack_reads(bases, starts)
        dev = torch.device("cuda", self.device)
        tw = torch.from_numpy(words.view(np.int32)).to(dev)
        ts = torch.from_numpy(starts).to(dev)
        torch.cuda.synchronize(dev)
        return Reads(DevReads(tw.data_ptr(), ts.data_ptr(), len(starts) - 1, int(starts[-1])), keep=(tw, ts))

    def reads_from_tensors(self, packed_i32, starts_i64, n_bases):
        return Reads(DevReads(packed_i32.data_ptr(), starts_i64.data_ptr(), starts_i64.numel() - 1, int(n_bases)),
                     keep=(packed_i32, starts_i64))

    def synth(self, n_pairs, read_len=150, mito_len=16500, nuclear_len=50_000_000, mito_fraction=0.05, error_rate=0.005,
              n_rate=1e-4, insert_mean=350.0, insert_sd=35.0, seed=1001):
        spec = SynthSpec(n_pairs, read_len, mito_len, nuclear_len, mito_fraction, error_rate, n_rate, insert_mean, insert_sd, seed)
        out = DevReads()
        _check(load().mfsdbg_dev_synth_reads(self._h, C.byref(spec), C.byref(out)))
        return Reads(out)

    def pack_fastq(self, text_ptr, n_bytes, n_policy=N_MEGAHIT):
        out, ml = DevReads(), C.c_int32(0)
        _check(load().mfsdbg_dev_pack_fastq(self._h, text_ptr, n_bytes, n_policy, C.byref(out), C.byref(ml)))
        return Reads(out), ml.value

    def download_reads(self, reads):
        nw = (reads.n_bases + 15) // 16
        words = self.d2h(reads.s.packed, nw * 4, np.uint32)
        starts = self.d2h(reads.s.starts, (reads.n_reads + 1) * 8, np.int64)
        return unpack_reads(words, reads.n_bases), starts

    # -- pipelines
    def count(self, reads, k, min_count, want_counting=False):
        out = DevEdges()
        counting = np.zeros(65536, np.int64) if want_counting else None
        _check(load().mfsdbg_dev_count(self._h, C.byref(reads.s), k, min_count, C.byref(out),
                                       counting.ctypes.data if want_counting else None))
        return Edges(self, out, counting)

    def seq2sdbg(self, edges, k=None, tip_mode=0):
        out = DevSdbg()
        _check(load().mfsdbg_dev_seq2sdbg(self._h, edges.s.edges, edges.s.n_edges, k or edges.s.k, tip_mode, C.byref(out)))
        return Sdbg(self, out)

    def read2sdbg(self, reads, k, min_count):
        out = DevSdbg()
        _check(load().mfsdbg_dev_read2sdbg(self._h, C.byref(reads.s), k, min_count, C.byref(out)))
        return Sdbg(self, out)

    def host_read2sdbg(self, packed_host_ptr, starts_host_ptr, n_reads, n_bases, k, min_count):
        """read2sdbg with host inputs/outputs (H2D + D2H inside the call). Returns the HostSdbg struct."""
        out = HostSdbg()
        _check(load().mfsdbg_host_read2sdbg(self._h, packed_host_ptr, starts_host_ptr, n_reads, n_bases, k, min_count,
                                            C.byref(out)))
        return out

    # -- staged count (multi-GPU driver)
    def count_hist(self, reads, k, l1_bits, hist_ptr):
        _check(load().mfsdbg_dev_count_hist(self._h, C.byref(reads.s), k, l1_bits, hist_ptr))

    def count_scatter(self, reads, k, l1_bits, hist_ptr, keys_ptr, capacity):
        _check(load().mfsdbg_dev_count_scatter(self._h, C.byref(reads.s), k, l1_bits, hist_ptr, keys_ptr, capacity))

    def count_finish(self, keys_ptr, scratch_ptr, n_keys, chunk_start, chunk_size, chunk_seg, n_segs, k, l1_bits, min_count,
                     want_counting=False):
        cs = np.ascontiguousarray(chunk_start, dtype=np.int64)
        cz = np.ascontiguousarray(chunk_size, dtype=np.int64)
        cg = np.ascontiguousarray(chunk_seg, dtype=np.int32)
        out = DevEdges()
        counting = np.zeros(65536, np.int64) if want_counting else None
        _check(load().mfsdbg_dev_count_finish(self._h, keys_ptr, scratch_ptr, n_keys, cs.ctypes.data, cz.ctypes.data,
                                              cg.ctypes.data, len(cs), n_segs, k, l1_bits, min_count, C.byref(out),
                                              counting.ctypes.data if want_counting else None))
        return Edges(self, out, counting)


    # -- peer memory (fused partition + exchange)
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        _check(load().mfsdbg_dev_alloc(self._h, int(nbytes), C.byref(p)))
        return p.value

    def dev_free(self, ptr):
        _check(load().mfsdbg_dev_free(self._h, ptr))

    def ipc_export(self, ptr):
        h = np.zeros(64, np.uint8)
        _check(load().mfsdbg_ipc_export(self._h, ptr, h.ctypes.data))
        return h

    def ipc_open(self, handle):
        h = np.ascontiguousarray(handle, dtype=np.uint8)
        p = C.c_void_p()
        _check(load().mfsdbg_ipc_open(self._h, h.ctypes.data, C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        _check(load().mfsdbg_ipc_close(self._h, ptr))

    def count_scatter_peer(self, reads, k, l1_bits, bin_base_ptr):
        _check(load().mfsdbg_dev_count_scatter_peer(self._h, C.byref(reads.s), k, l1_bits, bin_base_ptr))

    def records_scatter_peer(self, rec_ptr, n, words, l1_bits, bin_base_ptr):
        _check(load().mfsdbg_dev_records_scatter_peer(self._h, rec_ptr, n, words, l1_bits, bin_base_ptr))

    # -- super-k-mer exchange (multi-GPU count, 16 <= k <= 26)
    def skm_scatter(self, reads, k, n_dst, dst_ptrs=None, dst_caps=None, stride=1):
        """Returns (records[n_dst], keys[n_dst]) this source produced per destination; dst_ptrs None = count only."""
        counts = np.zeros(2 * n_dst, np.int64)
        if dst_ptrs is None:
            _check(load().mfsdbg_dev_skm_scatter(self._h, C.byref(reads.s), k, n_dst, None, None, stride, counts.ctypes.data))
        else:
            p = np.ascontiguousarray(dst_ptrs, dtype=np.uint64)
            cp = np.ascontiguousarray(dst_caps, dtype=np.int64)
            assert len(p) == n_dst and len(cp) == n_dst
            _check(load().mfsdbg_dev_skm_scatter(self._h, C.byref(reads.s), k, n_dst, p.ctypes.data, cp.ctypes.data, 1,
                                                 counts.ctypes.data))
        return counts[:n_dst].copy(), counts[n_dst:].copy()

    def count_skm(self, rec_ptr, chunk_start, chunk_size, n_keys, k, min_count, keys_ptr, scratch_ptr, capacity):
        cs = np.ascontiguousarray(chunk_start, dtype=np.int64)
        cz = np.ascontiguousarray(chunk_size, dtype=np.int64)
        out = DevEdges()
        _check(load().mfsdbg_dev_count_skm(self._h, rec_ptr, cs.ctypes.data, cz.ctypes.data, len(cs), int(n_keys), k, min_count,
                                           keys_ptr, scratch_ptr, int(capacity), C.byref(out)))
        return Edges(self, out)

    # -- the item filter across GPUs (16 <= k <= 31)
    def ks_hist(self, edges_ptr, n_edges, k, log_slots, slice_log, hist_ptr):
        _check(load().mfsdbg_dev_ks_hist(self._h, edges_ptr, n_edges, k, log_slots, slice_log, hist_ptr))

    def ks_scatter_peer(self, edges_ptr, n_edges, k, log_slots, slice_log, bin_base_ptr):
        _check(load().mfsdbg_dev_ks_scatter_peer(self._h, edges_ptr, n_edges, k, log_slots, slice_log, bin_base_ptr))

    def ks_filter(self, ins_ptr, n_ins, qry_ptr, n_qry, log_slots, slice_log, slice_lo, n_owned):
        n = C.c_int64(0)
        _check(load().mfsdbg_dev_ks_filter(self._h, ins_ptr, int(n_ins), qry_ptr, int(n_qry), log_slots, slice_log, slice_lo, n_owned,
                                           C.byref(n)))
        return n.value

    def ks_items(self, edges_ptr, n_edges, n_miss, k, items_ptr, capacity):
        n = C.c_int64(0)
        _check(load().mfsdbg_dev_ks_items(self._h, edges_ptr, n_edges, int(n_miss), k, items_ptr, int(capacity), C.byref(n)))
        return n.value

    def sdbg_items(self, edges_ptr, n_edges, k, items_ptr):
        _check(load().mfsdbg_dev_sdbg_items(self._h, edges_ptr, n_edges, k, items_ptr))

    def records_hist(self, rec_ptr, n, words, l1_bits, hist_ptr):
        _check(load().mfsdbg_dev_records_hist(self._h, rec_ptr, n, words, l1_bits, hist_ptr))

    def records_scatter(self, rec_ptr, n, words, l1_bits, hist_ptr, out_ptr):
        _check(load().mfsdbg_dev_records_scatter(self._h, rec_ptr, n, words, l1_bits, hist_ptr, out_ptr))

    def sdbg_finish(self, items_ptr, scratch_ptr, n_items, chunk_start, chunk_size, chunk_seg, n_segs, k, l1_bits, tip_mode):
        cs = np.ascontiguousarray(chunk_start, dtype=np.int64)
        cz = np.ascontiguousarray(chunk_size, dtype=np.int64)
        cg = np.ascontiguousarray(chunk_seg, dtype=np.int32)
        out = DevSdbg()
        _check(load().mfsdbg_dev_sdbg_finish(self._h, items_ptr, scratch_ptr, n_items, cs.ctypes.data, cz.ctypes.data,
                                             cg.ctypes.data, len(cs), n_segs, k, l1_bits, tip_mode, C.byref(out)))
        return Sdbg(self, out)


class Edges:
    def __init__(self, ctx, s, counting=None):
        self.ctx, self.s, self.counting = ctx, s, counting

    @property
    def n(self):
        return self.s.n_edges

    def to_numpy(self):
        w = self.s.words_per_edge
        return self.ctx.d2h(self.s.edges, self.s.n_edges * w * 4, np.uint32).reshape(-1, w)

    def bucket_counts(self):
        out = np.zeros(65536, np.int64)
        _check(load().mfsdbg_ctx_edge_bucket_counts(self.ctx._h, out.ctypes.data))
        return out


class Sdbg:
    def __init__(self, ctx, s):
        self.ctx, self.s = ctx, s

    @property
    def n(self):
        return self.s.n_items

    def to_numpy(self):
        rec = self.ctx.d2h(self.s.rec, self.s.n_items * 4, np.uint32)
        wt = self.s.words_per_tip
        labels = self.ctx.d2h(self.s.tip_labels, self.s.n_tips * wt * 4, np.uint32).reshape(-1, wt)
        return dict(w=(rec & 15).astype(np.uint8), last=((rec >> 4) & 1).astype(np.uint8), tip=((rec >> 5) & 1).astype(np.uint8),
                    mul=(rec >> 8).astype(np.uint16), tip_labels=labels, n_large=self.s.n_large)

    def bucket_stats(self):
        out = np.zeros((65536, 3), np.int64)
        _check(load().mfsdbg_ctx_sdbg_bucket_stats(self.ctx._h, out.ctypes.data))
        return out

"""ctypes loader for oracle/libmhoracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED: restates megahit v1.2.9 (not vendored in /root/reference) from
recollection; see oracle/mh_oracle.h.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

N_MEGAHIT, N_SPLIT = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "libmhoracle.so")
    src = [os.path.join(_HERE, f) for f in ("mh_oracle.c", "mh_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "libmhoracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    vp, i64, i32, cp = C.c_void_p, C.c_int64, C.c_int, C.c_char_p
    sig = {
        "orc_reads_new": (vp, []), "orc_reads_free": (None, [vp]),
        "orc_reads_add_ascii": (None, [vp, cp, i64, i32]),
        "orc_reads_add_fastx": (i32, [vp, cp, i32]), "orc_reads_add_fastx_pe": (i32, [vp, cp, cp, i32]),
        "orc_reads_load_bin": (vp, [cp]), "orc_reads_write_bin": (i32, [vp, cp]),
        "orc_reads_count": (i64, [vp]), "orc_reads_bases": (i64, [vp]), "orc_reads_max_len": (i32, [vp]),
        "orc_reads_data": (vp, [vp]), "orc_reads_starts": (vp, [vp]),
        "orc_cmd_buildlib": (i32, [cp, cp, i32]),
        "orc_count": (vp, [vp, vp, i64, i32, i32, i32]), "orc_edges_free": (None, [vp]),
        "orc_edges_k": (i32, [vp]), "orc_edges_words": (i32, [vp]), "orc_edges_sorted": (i32, [vp]),
        "orc_edges_n": (i64, [vp]), "orc_edges_data": (vp, [vp]), "orc_edges_bucket_counts": (vp, [vp]),
        "orc_edges_counting": (vp, [vp]), "orc_edges_write": (i32, [vp, cp, i32]), "orc_edges_read": (vp, [cp]),
        "orc_cmd_count": (i32, [cp, i32, i32, cp, i32]),
        "orc_seqs_new": (vp, []), "orc_seqs_free": (None, [vp]), "orc_seqs_add": (None, [vp, vp, i64, i32]),
        "orc_seqs_add_edges": (None, [vp, vp]), "orc_seqs_add_contigs": (i32, [vp, cp, i32, i32, i32, i32]),
        "orc_seqs_count": (i64, [vp]),
        "orc_seq2sdbg": (vp, [vp, i32, i32]), "orc_read2sdbg": (vp, [vp, vp, i64, i32, i32, i32]),
        "orc_sdbg_free": (None, [vp]), "orc_sdbg_k": (i32, [vp]), "orc_sdbg_words_per_tip": (i32, [vp]),
        "orc_sdbg_n": (i64, [vp]), "orc_sdbg_n_tips": (i64, [vp]), "orc_sdbg_n_large": (i64, [vp]),
        "orc_sdbg_w": (vp, [vp]), "orc_sdbg_last": (vp, [vp]), "orc_sdbg_tip": (vp, [vp]), "orc_sdbg_mul": (vp, [vp]),
        "orc_sdbg_tip_labels": (vp, [vp]), "orc_sdbg_bucket_items": (vp, [vp]),
        "orc_sdbg_write": (i32, [vp, cp, i32]), "orc_sdbg_read": (vp, [cp]),
        "orc_cmd_seq2sdbg": (i32, [i32, i32, cp, cp, cp, cp, cp, cp, i32]),
        "orc_cmd_read2sdbg": (i32, [cp, i32, i32, cp, i32]),
        "orc_last_error": (cp, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _LIB = L
    return L


class OracleError(RuntimeError):
    pass


def _err():
    return OracleError(lib().orc_last_error().decode())


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def _b(s):
    return s.encode() if isinstance(s, str) else s


class Reads:
    """bases: uint8 (0..3) true orientation, starts: int64 [n+1]."""

    def __init__(self, bases, starts, max_len=None):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.starts = np.ascontiguousarray(starts, dtype=np.int64)
        self.max_len = int(np.diff(self.starts).max()) if max_len is None and len(self.starts) > 1 else (max_len or 0)

    @property
    def n(self):
        return len(self.starts) - 1

    @staticmethod
    def _from_handle(h):
        L = lib()
        n, nb = L.orc_reads_count(h), L.orc_reads_bases(h)
        r = Reads(_arr(L.orc_reads_data(h), nb, np.uint8), _arr(L.orc_reads_starts(h), n + 1, np.int64),
                  L.orc_reads_max_len(h))
        L.orc_reads_free(h)
        return r

    @staticmethod
    def from_ascii(seqs, n_policy=N_MEGAHIT):
        L = lib()
        h = L.orc_reads_new()
        for s in seqs:
            s = _b(s)
            L.orc_reads_add_ascii(h, s, len(s), n_policy)
        return Reads._from_handle(h)

    @staticmethod
    def load_bin(path):
        h = lib().orc_reads_load_bin(_b(path))
        if not h:
            raise _err()
        return Reads._from_handle(h)


class Edges:
    def __init__(self, h):
        L = lib()
        self.k, self.words, self.sorted = L.orc_edges_k(h), L.orc_edges_words(h), bool(L.orc_edges_sorted(h))
        n = L.orc_edges_n(h)
        self.data = _arr(L.orc_edges_data(h), n * self.words, np.uint32).reshape(n, self.words)
        self.bucket_counts = _arr(L.orc_edges_bucket_counts(h), 65536, np.int64)
        self.counting = _arr(L.orc_edges_counting(h), 65536, np.int64)
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_edges_free(self._h)
            self._h = None

    @property
    def n(self):
        return self.data.shape[0]

    def write(self, prefix, n_files=1):
        if lib().orc_edges_write(self._h, _b(prefix), n_files):
            raise _err()

    @staticmethod
    def read(prefix):
        h = lib().orc_edges_read(_b(prefix))
        if not h:
            raise _err()
        return Edges(h)


class Sdbg:
    def __init__(self, h):
        L = lib()
        self.k, self.words_per_tip = L.orc_sdbg_k(h), L.orc_sdbg_words_per_tip(h)
        n, nt = L.orc_sdbg_n(h), L.orc_sdbg_n_tips(h)
        self.n_large = L.orc_sdbg_n_large(h)
        self.w = _arr(L.orc_sdbg_w(h), n, np.uint8)
        self.last = _arr(L.orc_sdbg_last(h), n, np.uint8)
        self.tip = _arr(L.orc_sdbg_tip(h), n, np.uint8)
        self.mul = _arr(L.orc_sdbg_mul(h), n, np.uint16)
        self.tip_labels = _arr(L.orc_sdbg_tip_labels(h), nt * self.words_per_tip, np.uint32).reshape(nt, self.words_per_tip)
        self.bucket_items = _arr(L.orc_sdbg_bucket_items(h), 65536, np.int64)
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sdbg_free(self._h)
            self._h = None

    @property
    def n(self):
        return len(self.w)

    def write(self, prefix, n_files=1):
        if lib().orc_sdbg_write(self._h, _b(prefix), n_files):
            raise _err()

    @staticmethod
    def read(prefix):
        h = lib().orc_sdbg_read(_b(prefix))
        if not h:
            raise _err()
        return Sdbg(h)


def count(reads, k, min_count, threads=1):
    h = lib().orc_count(reads.bases.ctypes.data, reads.starts.ctypes.data, reads.n, k, min_count, threads)
    if not h:
        raise _err()
    return Edges(h)


class Seqs:
    def __init__(self):
        self._h = lib().orc_seqs_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_seqs_free(self._h)
            self._h = None

    def add(self, bases, mult):
        b = np.ascontiguousarray(bases, dtype=np.uint8)
        lib().orc_seqs_add(self._h, b.ctypes.data, len(b), int(mult))

    def add_edges(self, edges):
        lib().orc_seqs_add_edges(self._h, edges._h)

    def add_contigs(self, path, min_len, extend_loop=False, k_from=0, k_to=0):
        if lib().orc_seqs_add_contigs(self._h, _b(path), min_len, int(extend_loop), k_from, k_to):
            raise _err()

    @property
    def n(self):
        return lib().orc_seqs_count(self._h)


def seq2sdbg(seqs, k, threads=1):
    h = lib().orc_seq2sdbg(seqs._h, k, threads)
    if not h:
        raise _err()
    return Sdbg(h)


def read2sdbg(reads, k, min_count, threads=1):
    h = lib().orc_read2sdbg(reads.bases.ctypes.data, reads.starts.ctypes.data, reads.n, k, min_count, threads)
    if not h:
        raise _err()
    return Sdbg(h)


def cmd_buildlib(lib_file, out_prefix, n_policy=N_MEGAHIT):
    if lib().orc_cmd_buildlib(_b(lib_file), _b(out_prefix), n_policy):
        raise _err()


def cmd_count(read_lib_file, k, min_count, out_prefix, threads=1):
    if lib().orc_cmd_count(_b(read_lib_file), k, min_count, _b(out_prefix), threads):
        raise _err()


def cmd_seq2sdbg(k, k_from, out_prefix, input_prefix=None, contig=None, bubble=None, addi_contig=None,
                 local_contig=None, threads=1):
    o = lambda s: _b(s) if s else None
    if lib().orc_cmd_seq2sdbg(k, k_from, o(input_prefix), o(contig), o(bubble), o(addi_contig), o(local_contig),
                              _b(out_prefix), threads):
        raise _err()


def cmd_read2sdbg(read_lib_file, k, min_count, out_prefix, threads=1):
    if lib().orc_cmd_read2sdbg(_b(read_lib_file), k, min_count, _b(out_prefix), threads):
        raise _err()

"""Host-side mirror of the reference's graph step: `MEGAHIT.graph` (/root/reference/assemble/assemble_wrapper.py:202-261)
with the two subprocess launches replaced by ctypes calls into libmfsdbg.  Same names, same option dicts, same file
discovery and the same EmptyGraph behaviour, so it can replace that method one-for-one."""
import os
from os import path

from . import lib


class EmptyGraph(Exception):
    pass


class GraphBuilder:
    """Carries the attributes `MEGAHIT.graph` reads from `self` (assemble_wrapper.py:60-98)."""

    def __init__(self, temp_dir, contig_dir, read_lib, threads=8, min_multi=3, no_mercy=True, one_pass=False, kmin=31,
                 keep_temp=True, available_memory=0):
        self.temp_dir, self.contig_dir, self.read_lib = temp_dir, contig_dir, read_lib
        self.threads, self.min_multi, self.no_mercy, self.one_pass = threads, int(min_multi), bool(no_mercy), bool(one_pass)
        self.kmin, self.keep_temp, self.available_memory = kmin, keep_temp, available_memory

    def _graph_prefix(self, kmer):                      # assemble_wrapper.py:93-94
        d = path.join(self.temp_dir, f"k{kmer}")
        os.makedirs(d, exist_ok=True)
        return path.join(d, str(kmer))

    def _contig_prefix(self, kmer):                     # assemble_wrapper.py:96-97
        return path.join(self.contig_dir, f"k{kmer}")

    def build_lib(self):                                # assemble_wrapper.py:193 (`megahit_core buildlib lib lib`)
        lib.buildlib(self.read_lib, self.read_lib)
        with open(self.read_lib + ".lib_info") as ri:
            return [x.split(" ") for x in ri.readlines()]

    def graph(self, current_kmer, next_kmer):           # assemble_wrapper.py:202-261
        options = {
            "k": next_kmer,
            "host_mem": int(self.available_memory),
            "mem_flag": 1,
            "output_prefix": self._graph_prefix(next_kmer),
            "num_cpu_threads": self.threads,
            "need_mercy": int(not self.no_mercy and current_kmer == self.kmin),
            "kmer_from": current_kmer,
        }
        if current_kmer == 0 and not self.one_pass:      # :215-224
            count_opts = dict(options)
            count_opts["min_count"] = self.min_multi
            count_opts["read_lib_file"] = self.read_lib
            count_opts.pop("need_mercy")
            count_opts.pop("kmer_from")
            lib.count(**count_opts)
        file_size = 0
        if path.exists(self._graph_prefix(next_kmer) + ".edges.0"):          # :228-230
            options["input_prefix"] = self._graph_prefix(next_kmer)
            file_size += path.getsize(self._graph_prefix(next_kmer) + ".edges.0")
        for key, suffix in (("addi_contig", ".addi.fa"), ("local_contig", ".local.fa")):   # :232-242
            f = self._contig_prefix(current_kmer) + suffix
            if path.exists(f):
                options[key] = f
                file_size += path.getsize(f)
        if path.exists(self._contig_prefix(current_kmer) + ".contigs.fa"):   # :244-250
            options["contig"] = self._contig_prefix(current_kmer) + ".contigs.fa"
            options["bubble"] = self._contig_prefix(current_kmer) + ".bubble_seq.fa"
            file_size += path.getsize(options["contig"])
        if file_size == 0 and current_kmer != 0:                              # :252-253
            raise EmptyGraph
        if "bubble" in options and not path.exists(options["bubble"]):
            options.pop("bubble")
        lib.seq2sdbg(**options)                                               # :258

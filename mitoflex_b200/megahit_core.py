"""`megahit_core`-compatible command line for the sDBG sub-commands, so that MitoFlex's unmodified
assemble/assemble_wrapper.py (which builds its argv with utility/helper.py:50-75 `concat_command`) runs on libmfsdbg when
this shim is first on PATH:

    megahit_core count    -k K --host_mem B --mem_flag 1 --output_prefix P --num_cpu_threads T -m M --read_lib_file L
    megahit_core seq2sdbg -k K --host_mem B --mem_flag 1 --output_prefix P --num_cpu_threads T --kmer_from F
                          [--input_prefix P] [--addi_contig F] [--local_contig F] [--contig F --bubble F] [--need_mercy]
    megahit_core read2sdbg (same options as count)
    megahit_core buildlib <lib_file> <out_prefix>
    megahit_core checkcpu | checkpopcnt      -> prints 1 (assemble_wrapper.py:122-125 compares captured stdout)

Everything else (assemble, local, iterate, ...) is graph traversal outside this library: it is handed to the real
megahit_core named by $MFSDBG_REAL_MEGAHIT_CORE (or the next `megahit_core*` on PATH).  Exit status 0 on success, 1 with
a message on stderr otherwise (utility/helper.py:78-86 turns non-zero into RuntimeError).
"""
import os
import shutil
import sys

OURS = ("count", "seq2sdbg", "read2sdbg", "buildlib", "checkcpu", "checkpopcnt")
VALUE_OPTS = {"-k": "k", "-m": "min_count", "--host_mem": "host_mem", "--mem_flag": "mem_flag", "--output_prefix": "output_prefix",
              "--num_cpu_threads": "num_cpu_threads", "--read_lib_file": "read_lib_file", "--kmer_from": "kmer_from",
              "--input_prefix": "input_prefix", "--contig": "contig", "--bubble": "bubble", "--addi_contig": "addi_contig",
              "--local_contig": "local_contig", "--kmer_k": "k", "--min_kmer_frequency": "min_count", "-t": "num_cpu_threads"}
INT_OPTS = {"k", "min_count", "host_mem", "mem_flag", "num_cpu_threads", "kmer_from"}
FLAG_OPTS = {"--need_mercy": "need_mercy"}


def parse(argv):
    opts, i = {}, 0
    while i < len(argv):
        a = argv[i]
        if a in FLAG_OPTS:
            opts[FLAG_OPTS[a]] = 1
            i += 1
        elif a in VALUE_OPTS:
            if i + 1 >= len(argv):
                raise ValueError(f"option {a} needs a value")
            name = VALUE_OPTS[a]
            opts[name] = int(float(argv[i + 1])) if name in INT_OPTS else argv[i + 1]
            i += 2
        else:
            raise ValueError(f"unknown option {a}")
    return opts


REAL_NAMES = ("megahit_core", "megahit_core_popcnt", "megahit_core_no_hw_accel", "megahit_core_no_hwaccel")


def find_real(invoked=None):
    """The real megahit_core behind this shim: $MFSDBG_REAL_MEGAHIT_CORE, else the first non-shim binary on PATH -- the variant
    the shim was invoked as first (MitoFlex picks megahit_core / _popcnt / _no_hwaccel from the CPU probes,
    assemble_wrapper.py:101-125, and spells the last one without the underscore megahit installs it with: both are tried)."""
    real = os.environ.get("MFSDBG_REAL_MEGAHIT_CORE")
    if real:
        return real
    me = os.path.realpath(sys.argv[0])
    base = os.path.basename(invoked or sys.argv[0])
    names = [n for n in REAL_NAMES if n == base]
    if base in ("megahit_core_no_hwaccel", "megahit_core_no_hw_accel"):
        names = ["megahit_core_no_hw_accel", "megahit_core_no_hwaccel"]
    names += [n for n in REAL_NAMES if n not in names]
    for name in names:
        for d in os.environ.get("PATH", "").split(os.pathsep):
            cand = os.path.join(d, name)
            try:
                if os.path.isfile(cand) and os.access(cand, os.X_OK) and os.path.realpath(cand) != me \
                        and b"mitoflex_b200" not in open(cand, "rb").read(4096):
                    return cand
            except OSError:
                continue
    return None


def forward(argv):
    real = find_real()
    if not real:
        sys.stderr.write(f"megahit_core shim: sub-command '{argv[0] if argv else ''}' is not part of libmfsdbg and no real "
                         "megahit_core was found (set MFSDBG_REAL_MEGAHIT_CORE)\n")
        return 1
    os.execv(real, [real] + argv)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in OURS:
        return forward(argv)
    cmd, rest = argv[0], argv[1:]
    if cmd in ("checkcpu", "checkpopcnt"):
        # the answers select the binary variant MitoFlex runs the FORWARDED stages with (assemble / local / iterate): when a
        # real megahit_core is there its probe of this CPU decides; the sDBG sub-commands themselves need neither BMI2 nor POPCNT
        real = find_real("megahit_core")
        if real:
            import subprocess
            try:
                out = subprocess.run([real, cmd], capture_output=True, text=True, timeout=30).stdout.strip()
                if out in ("0", "1"):
                    print(out)
                    return 0
            except (OSError, subprocess.SubprocessError):
                pass
        print(1)
        return 0
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mitoflex_b200 import lib
    try:
        if cmd == "buildlib":
            if len(rest) != 2:
                raise ValueError("usage: megahit_core buildlib <lib_file> <out_prefix>")
            policy = lib.N_SPLIT if os.environ.get("MFSDBG_N_POLICY", "megahit") == "split" else lib.N_MEGAHIT
            lib.buildlib(rest[0], rest[1], policy)
        else:
            opts = parse(rest)
            if "MFSDBG_GPU" in os.environ:   # "0" or "0,1,2,3": several GPUs share the sub-command inside this one process
                import ctypes
                want = [int(x) for x in os.environ["MFSDBG_GPU"].replace(" ", "").split(",") if x != ""]
                ids = (ctypes.c_int32 * len(want))(*want)
                opts["n_gpus"], opts["gpu_ids"] = len(want), ids
            getattr(lib, cmd)(**opts)
    except (lib.MfsdbgError, ValueError, OSError) as e:
        sys.stderr.write(f"megahit_core {cmd}: {e}\n")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())

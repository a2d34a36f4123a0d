"""CPU-side checks of the megahit_core shim: it must understand exactly the argv that the reference builds
(/root/reference/utility/helper.py:50-75 `concat_command` on the dicts of assemble/assemble_wrapper.py:204-250)."""
import subprocess
import sys
import os

from mitoflex_b200 import megahit_core as shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def concat_command(*args, **kwargs):
    """Restatement of utility/helper.py:50-75 for useconv=False (the only mode the sDBG call sites use)."""
    kwargs = {k[1:] if k.startswith("_") else k: v for k, v in kwargs.items() if v is not None}
    kwargs.pop("useconv", None)
    cmd = " ".join(str(a) for a in args)
    for k, v in kwargs.items():
        dash = "--" if len(k) > 1 else "-"
        if isinstance(v, bool):
            if v:
                cmd += f" {dash}{k}"
        elif isinstance(v, list):
            cmd += f" {dash}{k} {' '.join(v)}"
        else:
            cmd += f" {dash}{k} {v}"
    return cmd


def test_parse_count_argv():
    argv = concat_command("count", k=31, host_mem=17179869184.0, mem_flag=1, output_prefix="T/k31/31", num_cpu_threads=8, m=3,
                          read_lib_file="T/reads.lib", useconv=False).split()
    assert argv[0] == "count"
    o = shim.parse(argv[1:])
    assert o == dict(k=31, host_mem=17179869184, mem_flag=1, output_prefix="T/k31/31", num_cpu_threads=8, min_count=3,
                     read_lib_file="T/reads.lib")


def test_parse_seq2sdbg_argv():
    argv = concat_command("seq2sdbg", k=39, host_mem=1 << 33, mem_flag=1, output_prefix="T/k39/39", num_cpu_threads=8,
                          need_mercy=False, kmer_from=31, input_prefix="T/k39/39", addi_contig="C/k31.addi.fa",
                          local_contig="C/k31.local.fa", contig="C/k31.contigs.fa", bubble="C/k31.bubble_seq.fa",
                          useconv=False).split()
    o = shim.parse(argv[1:])
    assert o["k"] == 39 and o["kmer_from"] == 31 and "need_mercy" not in o
    assert o["contig"] == "C/k31.contigs.fa" and o["bubble"] == "C/k31.bubble_seq.fa"
    o = shim.parse(concat_command("seq2sdbg", k=21, need_mercy=True, output_prefix="x").split()[1:])
    assert o["need_mercy"] == 1


def test_cpu_probes_print_1():
    for name in ("megahit_core", "megahit_core_popcnt", "megahit_core_no_hwaccel"):
        exe = os.path.join(ROOT, "mitoflex_b200", "bin", name)
        for probe in ("checkcpu", "checkpopcnt"):
            out = subprocess.check_output([sys.executable, exe, probe]).decode()
            assert out.rstrip() == "1"      # assemble_wrapper.py:122-125 compares the captured stdout with '1'


def test_unknown_subcommand_without_real_binary_fails(tmp_path):
    exe = os.path.join(ROOT, "mitoflex_b200", "bin", "megahit_core")
    env = dict(os.environ, PATH=str(tmp_path))
    env.pop("MFSDBG_REAL_MEGAHIT_CORE", None)
    p = subprocess.run([sys.executable, exe, "assemble", "-s", "x"], env=env, capture_output=True)
    assert p.returncode == 1 and b"not part of libmfsdbg" in p.stderr


def test_bad_option_is_an_error(tmp_path):
    exe = os.path.join(ROOT, "mitoflex_b200", "bin", "megahit_core")
    p = subprocess.run([sys.executable, exe, "count", "--no_such_flag", "1"], capture_output=True)
    assert p.returncode == 1


def test_probes_and_forwarding_follow_the_real_binary(tmp_path):
    """with a real megahit_core on PATH its answer to checkcpu / checkpopcnt is passed through (it selects the variant MitoFlex
    runs the forwarded stages with), and a forwarded sub-command goes to the variant the shim was invoked as -- including
    MitoFlex's `megahit_core_no_hwaccel` spelling (assemble_wrapper.py:101) of megahit's `megahit_core_no_hw_accel`."""
    real_dir = tmp_path / "real"
    real_dir.mkdir()
    for name, ans in (("megahit_core", "0"), ("megahit_core_no_hw_accel", "1")):
        f = real_dir / name
        f.write_text(f"#!/bin/sh\nif [ \"$1\" = checkcpu ] || [ \"$1\" = checkpopcnt ]; then echo {ans}; else echo {name} \"$@\"; fi\n")
        f.chmod(0o755)
    shim_dir = os.path.join(ROOT, "mitoflex_b200", "bin")
    env = dict(os.environ, PATH=shim_dir + os.pathsep + str(real_dir) + os.pathsep + os.environ.get("PATH", ""))
    env.pop("MFSDBG_REAL_MEGAHIT_CORE", None)
    out = subprocess.check_output([sys.executable, os.path.join(shim_dir, "megahit_core"), "checkcpu"], env=env).decode()
    assert out.strip() == "0"
    out = subprocess.check_output([sys.executable, os.path.join(shim_dir, "megahit_core_no_hwaccel"), "assemble", "-s", "p"], env=env).decode()
    assert out.strip() == "megahit_core_no_hw_accel assemble -s p"
    out = subprocess.check_output([sys.executable, os.path.join(shim_dir, "megahit_core"), "iterate", "-k", "21"], env=env).decode()
    assert out.strip() == "megahit_core iterate -k 21"

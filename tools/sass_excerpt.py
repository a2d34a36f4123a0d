#!/usr/bin/env python
"""profiles/sass_<kernel>.txt: mnemonic counts and the lines that carry bulk copies, mbarrier operations, atomics, barriers and
vector shared / global accesses of one kernel of libmfsdbg.so (cuobjdump -sass).

  python tools/sass_excerpt.py k_reads_scatter_compact2ILi8E profiles/sass_k_reads_scatter_compact2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mitoflex_b200", "libmfsdbg.so")
KEEP = re.compile(r"UBLKCP|SYNCS|ATOMS|ATOMG|RED\.|BAR\.|STS\.(64|128)|LDS\.(64|128)|STG\.E\.(64|128)|LDG\.E\.(64|128)")


def main(pattern, out_path):
    names = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, body, found = None, [], None
    for line in names.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if found:
                break
            cur = m.group(1)
            if pattern in cur:
                found = cur
            continue
        if found and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            body.append(line.rstrip())
    if not found:
        sys.exit(f"no function matches {pattern}")
    ops = collections.Counter()
    for line in body:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("STS", "LDS", "STG", "LDG")) and "." in m.group(1) else "")] += 1
    with open(out_path, "w") as f:
        f.write(f"# cuobjdump -sass mitoflex_b200/libmfsdbg.so, function {found}\n")
        f.write(f"# {len(body)} SASS instructions; mnemonic counts: {dict(ops.most_common(24))}\n")
        f.write("# lines with bulk copies (UBLKCP), mbarrier operations (SYNCS), atomics, barriers and 64- / 128-bit shared and global accesses:\n")
        for line in body:
            if KEEP.search(line):
                f.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line) + "\n")
    print(out_path, len(body), "instructions")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

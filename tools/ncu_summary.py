#!/usr/bin/env python
"""Reduce `ncu -i <rep> --page raw --csv` to the columns that drive decisions (one row per profiled launch)."""
import csv
import subprocess
import sys

COLS = [("Kernel Name", "kernel", 44), ("gpu__time_duration.sum", "ms", 8), ("dram__bytes_read.sum", "rd", 8), ("dram__bytes_write.sum", "wr", 8),
        ("smsp__inst_executed.sum", "winst", 11), ("smsp__issue_active.avg.per_cycle_active", "ipc", 5),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6), ("launch__registers_per_thread", "regs", 4),
        ("launch__grid_size", "grid", 8), ("lts__t_sector_hit_rate.pct", "l2hit", 6),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst", 8)]
STALLS = "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print(" | ".join(n for _, n, _ in COLS) + " | top stalls")
    for r in rows[2:]:
        cells = []
        for h, _, w in COLS:
            v = r[idx[h]] if h in idx else ""
            u = units[idx[h]] if h in idx else ""
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" + (u[0] if u in ("Gbyte", "Mbyte", "Kbyte") else "")
            except ValueError:
                v = v[:w]
            cells.append(v)
        st = sorted(((float(r[idx[h]].replace(",", "") or 0), h.split("stalled_")[1].split("_per_issue")[0]) for h in stall_cols), reverse=True)[:4]
        tot = sum(float(r[idx[h]].replace(",", "") or 0) for h in stall_cols) or 1.0
        print(" | ".join(cells) + " | " + " ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st))


if __name__ == "__main__":
    main(sys.argv[1])

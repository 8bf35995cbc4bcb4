#!/usr/bin/env python3
"""Turn .ncu-rep captures into the tracked evidence under profiles/.

    python tools/ncu_summary.py out_prefix rep1.ncu-rep [rep2.ncu-rep ...]

For every kernel launch in the reports: duration, warp instructions, issue-slot utilisation,
occupancy, registers, shared memory, DRAM bytes -- as CSV (out_prefix.csv) and Markdown
(out_prefix.md).  The numbers quoted in DESIGN.md come from these files; `ncu` must be on PATH."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "us"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps"),
    ("launch__registers_per_thread", "regs"),
    ("launch__block_size", "block"),
    ("launch__grid_size", "grid"),
    ("launch__shared_mem_per_block_static", "smem_static"),
    ("launch__shared_mem_per_block_dynamic", "smem_dynamic"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"report": rep.split("/")[-1], "kernel": r[hdr.index("Kernel Name")][:70]}
        for m, name in METRICS:
            if m in hdr:
                v, u = r[hdr.index(m)], units[hdr.index(m)]
                d[name] = v + ((" " + u) if u and name in ("dram_read", "dram_write", "smem_static", "smem_dynamic") else "")
        res.append(d)
    return res


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    rows = [r for rep in reps for r in rows_of(rep)]
    cols = ["report", "kernel"] + [n for _, n in METRICS]
    with open(prefix + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, cols)
        w.writeheader()
        w.writerows(rows)
    with open(prefix + ".md", "w") as f:
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for r in rows:
            f.write("| " + " | ".join(str(r.get(c, "")) for c in cols) + " |\n")
    print("wrote", prefix + ".csv", prefix + ".md", len(rows), "launches")


if __name__ == "__main__":
    main()

"""Summarise .ncu-rep files (ncu --set full) into a small markdown table for profiles/.
usage: python tools/ncu_summary.py out.md rep1.ncu-rep [rep2 ...]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    lines = ["| report | kernel | " + " | ".join(n for _, n in WANT) + " |", "|---|---|" + "---|" * len(WANT)]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            cells = []
            for key, _ in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    v = r[i]
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    cells.append(f"{v} {units[i]}".strip())
                else:
                    cells.append("-")
            lines.append(f"| {rep.split('/')[-1]} | `{name[:70]}` | " + " | ".join(cells) + " |")
    # warp-state (stall) breakdown: every `..issue_stalled_<reason>_per_issue_active.ratio` column of --set full,
    # top reasons per kernel (average warps stalled on <reason> per issue slot)
    lines += ["", "## Stall breakdown (warps stalled per issue-active cycle, top reasons)", "",
              "| report | kernel | waves/SM | " + "top stall reasons |", "|---|---|---|---|"]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr = rows[0]
        cols = [(i, h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        for r in rows[2:]:
            vals = []
            for i, h in cols:
                try:
                    vals.append((float(r[i].replace(",", "")), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
            vals.sort(reverse=True)
            waves = r[hdr.index("launch__waves_per_multiprocessor")] if "launch__waves_per_multiprocessor" in hdr else "-"
            top = ", ".join(f"{n} {v:.2f}" for v, n in vals[:5])
            lines.append(f"| {rep.split('/')[-1]} | `{r[hdr.index('Kernel Name')][:60]}` | {waves} | {top} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()

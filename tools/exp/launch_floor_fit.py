"""Least-squares fit T(B) = a + bytes(B) / BW to the FLOOR lines printed by tools/exp/launch_floor (step and copy)."""
import re
import sys

import numpy as np

rows = []
for ln in open(sys.argv[1]):
    m = re.match(r"FLOOR B=(\d+) bytes=(\d+) U=(\d+) grid=(\d+) step_us=([\d.]+) copy_us=([\d.]+) null_us=([\d.]+) touch_us=([\d.]+)", ln)
    if m:
        rows.append([float(v) for v in m.groups()])
r = np.array(rows)
B, nbytes = r[:, 0], r[:, 1]
print("| B | bytes/launch | step us | copy us | null us | touch us | step GB/s | copy GB/s |\n|---|---|---|---|---|---|---|---|")
for row in r:
    print(f"| {int(row[0])} | {int(row[1])} | {row[4]:.2f} | {row[5]:.2f} | {row[6]:.2f} | {row[7]:.2f} | "
          f"{row[1] / row[4] / 1e3:.0f} | {row[1] / row[5] / 1e3:.0f} |")
for name, col in (("step", 4), ("copy", 5)):
    for lo in (0, 64):
        sel = B >= lo
        A = np.stack([np.ones(sel.sum()), nbytes[sel]], 1)
        (a, s), *_ = np.linalg.lstsq(A, r[sel, col], rcond=None)
        print(f"fit {name:4s} (B >= {lo or 8:3d}): T = {a:.2f} us + bytes / {1 / s / 1e3:.0f} GB/s   "
              f"-> T(B=64) = {a + 33554432 * s:.2f} us, frac of 6550.7 GB/s at B=64: {33554432 / (a + 33554432 * s) / 1e3 / 6550.7:.3f}")

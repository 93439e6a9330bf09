"""Print the hottest SASS instructions (by warp-stall samples) of each kernel in an ncu report.
usage: python tools/ncu_top.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["data"].append(r)
for b in blocks[:1]:
    ci = {h: i for i, h in enumerate(b["hdr"])}
    s = lambda r: int(r[ci["# Samples"]] or 0)  # noqa: E731
    tot = sum(s(r) for r in b["data"])
    print(b["name"][:100], "samples", tot, "instructions", len(b["data"]))
    for r in sorted(b["data"], key=lambda r: -s(r))[:topn]:
        print(f"{s(r):6d} {100 * s(r) / max(tot, 1):5.1f}%  exec={r[ci['Instructions Executed']]:>6}  {r[ci['Source']].strip()[:100]}")

#!/usr/bin/env python3
"""Top source lines of each kernel in an .ncu-rep (needs -lineinfo and --import-source on).

usage: tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-n]
"""
import csv, subprocess, sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Function Name":
        name = rows[i][1]
        h = rows[i + 1]
        isamp, iex = h.index("# Samples"), h.index("Instructions Executed")
        j = i + 2
        lines = []
        while j < len(rows) and not (rows[j] and rows[j][0] in ("File Path", "Function Name")):
            r = rows[j]
            if r and r[0].isdigit():
                lines.append((int(r[0]), r[1], num(r[isamp]), num(r[iex])))
            j += 1
        if want in name:
            ts, ti = sum(x[2] for x in lines) or 1, sum(x[3] for x in lines) or 1
            print("==", name[:90], "samples", ts, "instr", ti)
            for l in sorted(lines, key=lambda x: -x[2])[:topn]:
                print(f"{l[0]:5d} {l[2] / ts * 100:5.1f}%s {l[3] / ti * 100:5.1f}%i  {l[1][:100]}")
        i = j
    else:
        i += 1

#!/usr/bin/env python3
"""per-source-line warp-instruction counts of one kernel from an ncu report captured with --import-source on:
   python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                      "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
line = None
inst = collections.Counter(); thr = collections.Counter(); src = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":            # a source line row: aggregated
        line = int(r[0]); src[line] = r[1]
        try:
            inst[line] += int(r[ii]); thr[line] += int(r[it])
        except ValueError:
            pass
total = sum(inst.values())
print("total warp instructions %d" % total)
for ln, n in inst.most_common(top):
    print("%5d %6.2f%%  thr/warp %4.1f  %s" % (ln, 100.0 * n / total, thr[ln] / max(n, 1), src.get(ln, "")[:110].strip()))

#!/usr/bin/env python3
"""dynamic warp-instruction counts per SASS opcode of one kernel from an ncu report (--import-source on):
   python tools/ncu_opcodes.py report.ncu-rep kernel_regex"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cnt = collections.Counter()
for r in rows:
    if r and r[0] == "Address":
        hdr = r; isrc = hdr.index("Source"); ii = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr): continue
    parts = r[isrc].split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.rstrip(";")
    base = op.split(".")[0]
    key = base + ("." + op.split(".")[1] if base in ("IMAD", "SHF", "LOP3", "ISETP") and "." in op else "")
    try: cnt[key] += int(r[ii])
    except ValueError: pass
tot = sum(cnt.values())
print("total", tot)
for k, v in cnt.most_common(40): print("%-14s %6.2f%%" % (k, 100.0 * v / tot))

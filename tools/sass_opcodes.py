#!/usr/bin/env python3
"""static opcode mix per kernel of the shipped library (cuobjdump -sass): the evidence file profiles/*_sass_opcodes.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "openjph_b200/libojph_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "BAR", "IMAD", "IADD3", "LOP3", "SHF", "PRMT", "LDS", "STS", "LDG", "STG", "BRA", "FLO", "SHFL", "FMUL", "FADD"]
print("# cuobjdump -sass %s: static opcode mix per kernel (sm_100a), top mnemonics; UBLKCP = TMA bulk copy," % lib)
print("# LDGSTS = cp.async, SYNCS = mbarrier, BAR = named / CTA barriers.  Static counts, not executed counts (those: r02_kernels.json)")
cur, cnt = None, None
def flush():
    if cur:
        tot = sum(cnt.values())
        print("%-95s %5d instrs  %s" % (cur[:95], tot, " ".join("%s=%d" % (k, sum(v for o, v in cnt.items() if o.split(".")[0] == k)) for k in KEYS
                                                                 if any(o.split(".")[0] == k for o in cnt))))
for line in txt.split("\n"):
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        flush(); cur, cnt = m.group(1), collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        cnt[m.group(1)] += 1
flush()

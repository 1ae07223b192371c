#!/usr/bin/env python3
"""one 8K encode + decode (for ncu captures)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import openjph_b200 as ob
import images
W = int(os.environ.get("PW", "8192")); H = int(os.environ.get("PH", "8192"))
rev = os.environ.get("PREV", "1") == "1"
TH = int(os.environ.get("PTILE", "0"))       # tile height: PH = k * PTILE stacks k frames as tiles
p = ob.make_params(W, H, 3, 12, num_decomps=5, reversible=rev, color_transform=True, qfactor=0 if rev else 90,
                   tile=(W, TH) if TH else (0, 0))
frame = [f.astype(np.uint16) for f in images.synth_frame(W, H, 3, 12, 1234)]
enc = ob.Encoder(p, ob.U16)
n = int(os.environ.get("PN", "1"))
for _ in range(n):
    cs = enc.encode(frame)
dec = ob.Decoder()
for _ in range(n):
    out = dec.decode(cs, ob.U16)
print(len(cs), enc.timings(), dec.timings())

#!/usr/bin/env python3
"""A frame past the 32-bit addressing limits of the streaming DWT kernels (coefficient arena >= 2^30 words):
the codec must fall back to the general kernels and still round-trip losslessly.  20480 x 20480 x 3, 8-bit."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import openjph_b200 as ob
W = H = int(os.environ.get("BIG", "20480"))
rng = np.random.default_rng(3)
small = rng.integers(0, 256, (H // 16, W // 16), dtype=np.uint8)
planes = []
for c in range(3):
    a = small.repeat(16, axis=0).repeat(16, axis=1)
    a[::3, ::5] += np.uint8(c + 1)
    planes.append(a)
p = ob.make_params(W, H, 3, 8, num_decomps=5, reversible=True, color_transform=True)
t0 = time.time()
cs = ob.Encoder(p, ob.U8).encode(planes)
t1 = time.time()
out = ob.Decoder().decode(cs, ob.U8)
t2 = time.time()
ok = all(np.array_equal(a, b) for a, b in zip(out, planes))
print("big frame %dx%d: %d bytes, encode %.2f s decode %.2f s lossless=%s" % (W, H, len(cs), t1 - t0, t2 - t1, ok))
sys.exit(0 if ok else 1)

#!/usr/bin/env python3
"""Timeline of the resident bench loop: NW workers stream frames; after every frame the absolute CUDA-event
marks of its encode / decode stages are collected (ojb_enc_marks / ojb_dec_marks).  Prints how much of the wall
time had 0, 1, 2, ... kernel stages in flight and the busy time per stage kind."""
import os, sys, time, ctypes as C, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import openjph_b200 as ob
from openjph_b200 import _lib
import images
L = _lib.lib()
W = H = 8192; NC = 3
NW = int(os.environ.get("NW", "8")); STEPS = int(os.environ.get("STEPS", "6"))
MODE = os.environ.get("MODE", "both")       # both | enc | dec : what every worker loops over
p = ob.make_params(W, H, NC, 12, num_decomps=5, reversible=True, color_transform=True)
frame = [f.astype(np.uint16) for f in images.synth_frame(W, H, NC, 12, 1234)]
pin = [torch.empty((H, W), dtype=torch.uint16, pin_memory=True) for _ in range(NC)]
for t, f in zip(pin, frame): t.numpy()[:] = f
planes = (C.c_void_p * NC)(*[t.data_ptr() for t in pin])
cap = W * H * NC * 2 + (1 << 20)
def ck(rc):
    if rc != 0: raise RuntimeError(L.ojb_last_error().decode())
class Wk:
    def __init__(self):
        self.enc = L.ojb_enc_create(); self.dec = L.ojb_dec_create()
        ck(L.ojb_enc_configure(self.enc, C.byref(p), ob.U16))
        self.cs_pin = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        self.cs_dev = torch.empty(cap, dtype=torch.uint8, device="cuda")
        self.n = C.c_uint64(); self.fi = _lib.FrameInfo(); self.marks = []
        ck(L.ojb_enc_encode_frame(self.enc, planes, None, self.cs_pin.data_ptr(), cap, C.byref(self.n)))
        self.cs_len = self.n.value
        ck(L.ojb_enc_upload_frame(self.enc, planes, None))
        ck(L.ojb_enc_encode_resident(self.enc, self.cs_dev.data_ptr(), cap, C.byref(self.n), 1))   # cs_dev for MODE=dec
    def frame(self, keep):
        me = (C.c_float * 8)(*([0.0] * 8)); md = (C.c_float * 8)(*([0.0] * 8))
        if MODE in ("both", "enc"):
            ck(L.ojb_enc_encode_resident(self.enc, self.cs_dev.data_ptr(), cap, C.byref(self.n), 1))
            L.ojb_enc_marks(self.enc, me)
        if MODE in ("both", "dec"):
            ck(L.ojb_dec_read_headers(self.dec, self.cs_pin.data_ptr(), self.cs_len, ob.U16, C.byref(self.fi)))
            ck(L.ojb_dec_use_device_codestream(self.dec, self.cs_dev.data_ptr()))
            ck(L.ojb_dec_decode_resident(self.dec))
            L.ojb_dec_marks(self.dec, md)
        if keep: self.marks.append((list(me), list(md)))
ws = [Wk() for _ in range(NW)]
ck(L.ojb_marks_reference())
def loop(w):
    for i in range(STEPS + 2): w.frame(i >= 2)
th = [threading.Thread(target=loop, args=(w,)) for w in ws]
t0 = time.perf_counter()
for t in th: t.start()
for t in th: t.join()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
# stages: encode marks 1..2 dwt, 2..3 ht_encode, 5..6 assemble; decode marks 2..3 ht_decode, 3..4 dwt_inv
iv = []
for w in ws:
    for me, md in w.marks:
        if MODE in ("both", "enc"): iv += [("dwt", me[1], me[2]), ("enc", me[2], me[3]), ("asm", me[5], me[6])]
        if MODE in ("both", "dec"): iv += [("dec", md[2], md[3]), ("idwt", md[3], md[4])]
lo = min(a for _, a, b in iv); hi = max(b for _, a, b in iv)
ev = sorted([(a, 1, k) for k, a, b in iv] + [(b, -1, k) for k, a, b in iv])
depth = 0; last = lo; hist = {}
for t, d, k in ev:
    hist[depth] = hist.get(depth, 0.0) + (t - last); last = t; depth += d
span = hi - lo; nfr = sum(len(w.marks) for w in ws)
print("frames %d  span %.1f ms  -> %.3f ms/frame (wall %.3f)" % (nfr, span, span / nfr, dt * 1e3 / (NW * (STEPS + 2))))
print("stages in flight (event-to-event, includes queueing inside a stage): " + "  ".join("%d: %.0f%%" % (k, 100 * v / span) for k, v in sorted(hist.items())))
print("mode", MODE, "workers", NW)
for kind in ("dwt", "enc", "asm", "dec", "idwt"):
    d = [b - a for k, a, b in iv if k == kind]
    if not d: continue
    print("  %-5s mean %.3f ms  min %.3f  max %.3f" % (kind, sum(d) / len(d), min(d), max(d)))

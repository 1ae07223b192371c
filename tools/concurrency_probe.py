#!/usr/bin/env python3
"""How do the stages behave when several frames are on the device at once?  N threads, each with its own codec pair,
run encode-only, decode-only or encode+decode loops on resident 8K frames; prints frames/s and the mean per-stage
CUDA-event times seen by the workers (stretch = how much a stage slows down under concurrency)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np, torch
import openjph_b200 as ob
from openjph_b200 import _lib
import images
L = _lib.lib()
W = int(os.environ.get("PROBE_W", "8192")); H = int(os.environ.get("PROBE_H", "8192")); BD = int(os.environ.get("PROBE_BD", "12"))
p = ob.make_params(W, H, 3, BD, num_decomps=5, reversible=True, color_transform=True)
frame = [f.astype(np.uint16) for f in images.synth_frame(W, H, 3, BD, 1234)]
pin = [torch.from_numpy(f).pin_memory() for f in frame]
planes = (C.c_void_p * 3)(*[t.data_ptr() for t in pin])
cap = W * H * 3 * 2 + (1 << 20)
ENC = ("h2d", "dwt", "ht_encode", "d2h_lengths", "host_wait", "assemble", "d2h_out", "host_ms")
DEC = ("h2d", "host_parse", "ht_decode", "dwt_inv", "d2h_image", "_5", "_6", "host_ms")


class Wk:
    def __init__(self):
        self.enc = L.ojb_enc_create(); self.dec = L.ojb_dec_create()
        assert L.ojb_enc_configure(self.enc, C.byref(p), ob.U16) == 0
        assert L.ojb_enc_upload_frame(self.enc, planes, None) == 0
        self.cs = torch.empty(cap, dtype=torch.uint8, device="cuda")
        self.n = C.c_uint64(); self.fi = _lib.FrameInfo()
        self.enc_once(); self.dec_once()
        self.te = np.zeros(8); self.td = np.zeros(8); self.k = 0

    def enc_once(self):
        assert L.ojb_enc_encode_resident(self.enc, self.cs.data_ptr(), cap, C.byref(self.n), 1) == 0

    def dec_once(self):
        assert L.ojb_dec_read_headers_device(self.dec, self.cs.data_ptr(), self.n.value, ob.U16, C.byref(self.fi)) == 0
        assert L.ojb_dec_decode_resident(self.dec) == 0


def run(mode, nw, iters=int(os.environ.get("PROBE_ITERS", "12"))):
    ws = WS[:nw]
    for w in ws:
        w.te[:] = 0; w.td[:] = 0; w.k = 0

    def loop(w):
        te = (C.c_float * 8)(); td = (C.c_float * 8)()
        for _ in range(iters):
            if mode in ("enc", "both"):
                w.enc_once(); L.ojb_enc_timings(w.enc, te); w.te += np.array(te[:])
            if mode in ("dec", "both"):
                w.dec_once(); L.ojb_dec_timings(w.dec, td); w.td += np.array(td[:])
            w.k += 1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = [threading.Thread(target=loop, args=(w,)) for w in ws]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    te = sum(w.te for w in ws) / (nw * iters); td = sum(w.td for w in ws) / (nw * iters)
    s = "%-4s workers %2d  %.2f ms/frame " % (mode, nw, dt / (nw * iters) * 1e3)
    if mode != "dec":
        s += " enc[" + " ".join("%s=%.2f" % (k, v) for k, v in zip(ENC, te) if k in ("dwt", "ht_encode", "assemble", "host_ms")) + "]"
    if mode != "enc":
        s += " dec[" + " ".join("%s=%.2f" % (k, v) for k, v in zip(DEC, td) if k in ("ht_decode", "dwt_inv", "host_ms")) + "]"
    print(s, flush=True)


_WL = [int(x) for x in os.environ.get("PROBE_WORKERS", "1,2,4,8,12").split(",")]
WS = [Wk() for _ in range(max(_WL))]
for mode in os.environ.get("PROBE_MODES", "enc,dec,both").split(","):
    for nw in [int(x) for x in os.environ.get("PROBE_WORKERS", "1,2,4,8,12").split(",")]:
        run(mode, nw)

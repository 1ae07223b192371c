"""ctypes wrapper over oracle/libojph_port.so, the plain-C restatement (oracle/ojph_oracle.c).
Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(_ROOT, "oracle", "libojph_port.so")
_lib = None
I32P, F32P, U32P, U8P = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        import fcntl
        with open(os.path.join(_ROOT, "oracle", ".port.lock"), "w") as lock:       # one builder at a time (xdist)
            fcntl.flock(lock, fcntl.LOCK_EX)
            subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle"), "-f", "Makefile.port"])   # no-op when fresh
        _lib = C.CDLL(PATH)
        _lib.oj_ht_encode_block.restype = C.c_uint32
    return _lib


def ip(a): return a.ctypes.data_as(I32P)
def fp(a): return a.ctypes.data_as(F32P)
def up(a): return a.ctypes.data_as(U32P)


def encode_block(block, missing_msbs):
    h, w = block.shape
    stride = (w + 15) & ~15
    buf = np.zeros((h, stride), np.uint32); buf[:, :w] = block
    out = np.zeros(65536, np.uint8)
    n = lib().oj_ht_encode_block(up(buf), missing_msbs, w, h, stride, out.ctypes.data_as(U8P), out.size)
    return out[:n].tobytes()


def decode_block(data, w, h, missing_msbs, num_passes, len1, len2, causal=False):
    stride = (w + 15) & ~15
    out = np.zeros((h + 1, stride), np.uint32)
    src = np.frombuffer(bytes(data) + b"\0" * 64, np.uint8).copy()
    ok = lib().oj_ht_decode_block(src.ctypes.data_as(U8P), up(out), missing_msbs, num_passes, len1, len2, w, h, stride,
                                  1 if causal else 0)
    return out[:h, :w].copy(), bool(ok)


def forward_bands(p, frame, layout):
    """reference-order forward path on whole planes with the port's kernels: level shift / float, RCT/ICT,
    vertical-then-horizontal lifting per level, quantise -> {(comp, res, band): uint32 sign-magnitude plane}.
    layout(comp, res, band) -> (K_max, delta_inv, x0, y0 of the resolution) comes from the test."""
    L = lib()
    nc = p.num_comps
    rev = bool(p.reversible)
    planes = []
    for c in range(nc):
        a = np.ascontiguousarray(frame[c], np.int32)
        if rev:
            shift = 0 if p.is_signed[c] else -(1 << (p.bit_depth[c] - 1))
            d = np.zeros_like(a); L.oj_rev_convert(ip(a), ip(d), shift, a.size); planes.append(d)
        else:
            d = np.zeros(a.shape, np.float32); L.oj_irv_to_float(ip(a), fp(d), p.bit_depth[c], int(p.is_signed[c]), a.size)
            planes.append(d)
    if p.color_transform:
        n = planes[0].size
        o = [np.zeros_like(planes[0]) for _ in range(3)]
        if rev: L.oj_rct_fwd(ip(planes[0]), ip(planes[1]), ip(planes[2]), ip(o[0]), ip(o[1]), ip(o[2]), n)
        else: L.oj_ict_fwd(fp(planes[0]), fp(planes[1]), fp(planes[2]), fp(o[0]), fp(o[1]), fp(o[2]), n)
        planes[:3] = o
    out = {}
    D = p.num_decomps
    for c in range(nc):
        cur = planes[c]
        x0, y0 = layout["origin"](c)
        for r in range(D, 0, -1):
            h, w = cur.shape
            cur = np.ascontiguousarray(cur)
            if rev: L.oj_dwt53_fwd_level(ip(cur), w, h, w, x0, y0)
            else: L.oj_dwt97_fwd_level(fp(cur), w, h, w, x0, y0)
            # de-interleave by absolute parity
            ys = [(y0 + i) & 1 for i in range(h)]; xs = [(x0 + i) & 1 for i in range(w)]
            rows = [np.array([i for i in range(h) if ys[i] == k], int) for k in (0, 1)]
            cols = [np.array([i for i in range(w) if xs[i] == k], int) for k in (0, 1)]
            for b in (1, 2, 3):
                sub = cur[np.ix_(rows[b >> 1], cols[b & 1])]
                out[(c, r, b)] = quant(p, sub, *layout["quant"](c, r, b))
            cur = cur[np.ix_(rows[0], cols[0])]
            x0, y0 = (x0 + 1) >> 1, (y0 + 1) >> 1
        out[(c, 0, 0)] = quant(p, cur, *layout["quant"](c, 0, 0))
    return out


def quant(p, sub, kmax, delta_inv):
    L = lib()
    sub = np.ascontiguousarray(sub)
    o = np.zeros(sub.shape, np.uint32); mv = np.zeros(1, np.uint32)
    if sub.size == 0:
        return o
    if p.reversible: L.oj_rev_tx_to_cb32(ip(sub), up(o), kmax, sub.size, up(mv))
    else: L.oj_irv_tx_to_cb32(fp(sub), up(o), C.c_float(delta_inv), sub.size, up(mv), 1)
    return o

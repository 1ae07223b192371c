#!/usr/bin/env python3
"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the authoring container:

    python tests/golden/make_golden.py

Fixtures
  ht_multipass_blocks.npz  every coded code-block of the reference's in-tree stream
        subprojects/js/html/test.j2c (512x512 RGB, 9/7; 77 of its 89 coded blocks carry SigProp /
        MagRef passes -- the only in-tree pin for those passes) with the output of
        ojph_decode_codeblock32 for each.
  ht_encode_blocks.npz     random code-blocks (all shapes, densities, 0xFF-heavy) with the bytes
        ojph_encode_codeblock32 produced and what ojph_decode_codeblock32 returns for them.
  codestreams.npz          small frames + the codestream ojph::codestream produced for them (5/3 exact
        targets) and the reference's decode of its own 9/7 streams.
  kernels.npz              line-kernel vectors: 5/3 and 9/7 lifting (vertical steps, horizontal
        analysis/synthesis, both parities, widths 1..9 + 64), RCT/ICT, float<->int conversion,
        quantisation -- outputs of the reference's generic kernels.
"""
import ctypes as C
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refharness as R      # noqa: E402
import cases                # noqa: E402
import openjph_b200 as ob   # noqa: E402  (only make_params; nothing of the product is executed)


def multipass_blocks():
    """needs the product's packet parser only to LOCATE blocks; uses the emulator build"""
    import emu
    L = emu.emu_lib()
    cs = open("/root/reference/subprojects/js/html/test.j2c", "rb").read()
    d = ob.Decoder(lib=L); d.read_headers(cs)
    out = {"n": 0}
    i = 0
    for b in d.list_blocks():
        if b.num_passes == 0:
            continue
        data = cs[b.byte_off:b.byte_off + b.len1 + b.len2]
        want, ok = R.decode_block(data, b.w, b.h, b.missing_msbs, b.num_passes, b.len1, b.len2)
        out["data_%d" % i] = np.frombuffer(data, np.uint8)
        out["meta_%d" % i] = np.array([b.w, b.h, b.missing_msbs, b.num_passes, b.len1, b.len2, int(ok)], np.uint32)
        out["want_%d" % i] = want
        # the same bytes decoded with the stripe-causal rule (SPP ignores the row below the stripe,
        # ojph_block_decoder32.cpp:1411-1415): the reference cannot ENCODE such a stream, so the causal path of
        # the decoder is pinned by what the reference decodes from these blocks when told they are causal
        wantc, okc = R.decode_block(data, b.w, b.h, b.missing_msbs, b.num_passes, b.len1, b.len2, causal=True)
        out["wantc_%d" % i] = wantc
        out["okc_%d" % i] = np.array(int(okc))
        i += 1
    out["n"] = np.array(i)
    np.savez_compressed(os.path.join(HERE, "ht_multipass_blocks.npz"), **out)
    print("ht_multipass_blocks:", i, "blocks")


def whole_multipass_stream():
    """test_j2c.npz: the in-tree stream itself (32 KB, the only in-tree stream with SPP / MRP passes) and what the
    reference's ojph::codestream decodes from it -- as is, and with the vertically-causal bit of its COD segment
    (Scod style bit 3 of SPcod's code-block style byte) switched on"""
    cs = open("/root/reference/subprojects/js/html/test.j2c", "rb").read()
    planes, info = R.decode(cs)
    out = {"cs": np.frombuffer(cs, np.uint8)}
    for c, a in enumerate(planes):
        assert np.abs(a).max() < 32768
        out["dec%d" % c] = a.astype(np.int16)
    k = cs.index(b"\xff\x52")                      # COD: marker, Lcod(2), Scod(1), SGcod(4), SPcod: levels, xcb, ycb, style
    causal = bytearray(cs)
    assert causal[k + 12] & 0x40, "not an HT stream?"
    causal[k + 12] |= 0x08
    planes_c, _ = R.decode(bytes(causal))
    out["cs_causal"] = np.frombuffer(bytes(causal), np.uint8)
    for c, a in enumerate(planes_c):
        out["decc%d" % c] = a.astype(np.int16)
    out["differs"] = np.array(int(any(not np.array_equal(a, b) for a, b in zip(planes, planes_c))))
    np.savez_compressed(os.path.join(HERE, "test_j2c.npz"), **out)
    print("test_j2c: %d bytes, %d comps, causal decode differs: %d" % (len(cs), len(planes), int(out["differs"])))


def encode_blocks():
    rng = np.random.default_rng(2024)
    out = {}
    n = 0
    for it in range(120):
        w = int(rng.integers(1, 65)); h = int(rng.integers(1, min(64, 4096 // w) + 1))
        if it % 6 == 0: w, h = 64, 64
        kmax = int(rng.integers(1, 28))
        mode = it % 5
        if mode == 0: mag = rng.integers(0, 1 << kmax, (h, w), dtype=np.uint64)
        elif mode == 1: mag = np.full((h, w), (1 << kmax) - 1, dtype=np.uint64)
        elif mode == 2: mag = (rng.random((h, w)) < 0.03) * rng.integers(0, 1 << kmax, (h, w), dtype=np.uint64)
        elif mode == 3: mag = np.minimum(np.abs(rng.laplace(0, 2 ** (kmax / 3), (h, w))).astype(np.uint64), (1 << kmax) - 1)
        else: mag = (rng.random((h, w)) < 0.5) * np.uint64(1)
        if not mag.any(): continue
        sign = rng.integers(0, 2, (h, w), dtype=np.uint64)
        blk = ((sign << 31) | (mag << (31 - kmax))).astype(np.uint32)
        data = R.encode_block(blk, kmax - 1)
        dec, ok = R.decode_block(data, w, h, kmax - 1, 1, len(data), 0)
        out["blk_%d" % n] = blk; out["kmax_%d" % n] = np.array(kmax)
        out["bytes_%d" % n] = np.frombuffer(data, np.uint8); out["dec_%d" % n] = dec
        n += 1
    out["n"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "ht_encode_blocks.npz"), **out)
    print("ht_encode_blocks:", n, "blocks")


def codestreams():
    out = {}
    names = ["cfg1_256_gray_L1", "odd_rgb_L5", "offsets", "rgb16_noise", "block_4x4", "tilepart_RC_lrcp",
             "thin_off", "signed10", "po_PCRL_precincts", "sub420_planar", "tlm_tiles", "irv_rgb_qstep", "irv_rgb_q90"]
    allc = dict(cases.SMALL_REV + cases.SMALL_IRV)
    for nme in names:
        p = cases.make(allc[nme])
        frame = cases.frame_for(p, "noise" if "noise" in nme else "synth")
        cs = R.encode(p, frame)
        dec, _ = R.decode(cs)
        for c, f in enumerate(frame):
            out["%s_in%d" % (nme, c)] = f.astype(np.int32)
            out["%s_dec%d" % (nme, c)] = dec[c].astype(np.int32)
        out["%s_cs" % nme] = np.frombuffer(cs, np.uint8)
    np.savez_compressed(os.path.join(HERE, "codestreams.npz"), **out)
    print("codestreams:", len(names))


def kernels():
    L = R.lib()
    rng = np.random.default_rng(99)
    out = {}
    I32, F32 = C.POINTER(C.c_int32), C.POINTER(C.c_float)

    def ip(a): return a.ctypes.data_as(I32)
    def fp(a): return a.ctypes.data_as(F32)
    widths = list(range(1, 10)) + [64, 65]
    k = 0
    for w in widths:
        for even in (1, 0):
            src = rng.integers(-2000, 2000, w).astype(np.int32)
            lw, hw = (w + even) // 2, (w + 1 - even) // 2
            lo = np.zeros(max(lw, 1), np.int32); hi = np.zeros(max(hw, 1), np.int32)
            L.ojr_rev_horz_ana(ip(lo), ip(hi), ip(src), w, even)
            back = np.zeros(w, np.int32)
            L.ojr_rev_horz_syn(ip(back), ip(lo), ip(hi), w, even)
            out["rh_src_%d" % k] = src; out["rh_lo_%d" % k] = lo[:lw]; out["rh_hi_%d" % k] = hi[:hw]
            out["rh_meta_%d" % k] = np.array([w, even]); out["rh_back_%d" % k] = back
            fsrc = (rng.random(w).astype(np.float32) - 0.5)
            flo = np.zeros(max(lw, 1), np.float32); fhi = np.zeros(max(hw, 1), np.float32)
            L.ojr_irv_horz_ana(fp(flo), fp(fhi), fp(fsrc), w, even)
            fback = np.zeros(w, np.float32)
            L.ojr_irv_horz_syn(fp(fback), fp(flo), fp(fhi), w, even)
            out["ih_src_%d" % k] = fsrc; out["ih_lo_%d" % k] = flo[:lw]; out["ih_hi_%d" % k] = fhi[:hw]
            out["ih_back_%d" % k] = fback
            k += 1
    out["nh"] = np.array(k)
    n = 97
    for s in range(2):
        for syn in (0, 1):
            a = rng.integers(-3000, 3000, n).astype(np.int32); b = rng.integers(-3000, 3000, n).astype(np.int32)
            d = rng.integers(-3000, 3000, n).astype(np.int32); r = d.copy()
            L.ojr_rev_vert_step(s, ip(a), ip(b), ip(r), n, syn)
            out["rv_%d_%d" % (s, syn)] = np.stack([a, b, d, r])
    for s in range(4):
        for syn in (0, 1):
            a = rng.random(n).astype(np.float32) - 0.5; b = rng.random(n).astype(np.float32) - 0.5
            d = rng.random(n).astype(np.float32) - 0.5; r = d.copy()
            L.ojr_irv_vert_step(s, fp(a), fp(b), fp(r), n, syn)
            out["iv_%d_%d" % (s, syn)] = np.stack([a, b, d, r])
    out["irv_K"] = np.array(L.ojr_irv_K(), np.float32)
    out["irv_steps"] = np.array([L.ojr_irv_step(i) for i in range(4)], np.float32)
    r_ = rng.integers(-2048, 2048, n).astype(np.int32); g_ = rng.integers(-2048, 2048, n).astype(np.int32)
    b_ = rng.integers(-2048, 2048, n).astype(np.int32)
    y = np.zeros(n, np.int32); cb = np.zeros(n, np.int32); cr = np.zeros(n, np.int32)
    L.ojr_rct_forward(ip(r_), ip(g_), ip(b_), ip(y), ip(cb), ip(cr), n)
    out["rct"] = np.stack([r_, g_, b_, y, cb, cr])
    fr = rng.random(n).astype(np.float32) - 0.5; fg = rng.random(n).astype(np.float32) - 0.5
    fb = rng.random(n).astype(np.float32) - 0.5
    fy = np.zeros(n, np.float32); fcb = np.zeros(n, np.float32); fcr = np.zeros(n, np.float32)
    L.ojr_ict_forward(fp(fr), fp(fg), fp(fb), fp(fy), fp(fcb), fp(fcr), n)
    br = np.zeros(n, np.float32); bg = np.zeros(n, np.float32); bb = np.zeros(n, np.float32)
    L.ojr_ict_backward(fp(fy), fp(fcb), fp(fcr), fp(br), fp(bg), fp(bb), n)
    out["ict"] = np.stack([fr, fg, fb, fy, fcb, fcr, br, bg, bb])
    for bd, sg in ((8, 0), (12, 0), (10, 1)):
        lo = -(1 << (bd - 1)) if sg else 0
        v = rng.integers(lo, lo + (1 << bd), n).astype(np.int32)
        f = np.zeros(n, np.float32); L.ojr_irv_convert_to_float(ip(v), fp(f), bd, sg, n)
        f2 = (f * np.float32(1.3)).astype(np.float32)     # exercise clamping
        q = np.zeros(n, np.int32); L.ojr_irv_convert_to_integer(fp(f2), ip(q), bd, sg, n)
        out["cvt_%d_%d" % (bd, sg)] = np.stack([v.astype(np.float64), f.astype(np.float64), f2.astype(np.float64), q.astype(np.float64)])
    v = rng.integers(-30000, 30000, n).astype(np.int32)
    for kmax in (9, 15, 20):
        sm = np.zeros(n, np.uint32); mv = np.zeros(8, np.uint32)
        L.ojr_rev_tx_to_cb32(ip(v), sm.ctypes.data_as(C.c_void_p), kmax, n, mv.ctypes.data_as(C.c_void_p))
        bk = np.zeros(n, np.int32); L.ojr_rev_tx_from_cb32(sm.ctypes.data_as(C.c_void_p), ip(bk), kmax, n)
        out["txrev_%d" % kmax] = np.stack([v.astype(np.int64), sm.astype(np.int64), bk.astype(np.int64)])
    f = (rng.random(n).astype(np.float32) - 0.5)
    for dinv in (1.0e6, 3.3e7):
        sm = np.zeros(n, np.uint32); mv = np.zeros(8, np.uint32)
        L.ojr_irv_tx_to_cb32(fp(f), sm.ctypes.data_as(C.c_void_p), C.c_float(dinv), n, mv.ctypes.data_as(C.c_void_p))
        bk = np.zeros(n, np.float32)
        L.ojr_irv_tx_from_cb32(sm.ctypes.data_as(C.c_void_p), fp(bk), C.c_float(1.0 / dinv), n)
        out["txirv_%d" % int(dinv)] = np.stack([f.astype(np.float64), sm.astype(np.float64), bk.astype(np.float64)])
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    print("kernels: ok")


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first (make -C oracle)"
    encode_blocks(); codestreams(); kernels(); multipass_blocks(); whole_multipass_stream()

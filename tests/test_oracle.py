"""Pins the oracle (oracle/ojph_oracle.c, the plain-C restatement) against golden vectors produced by
the unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py) and -- when
oracle/_ref travelled -- against the reference live."""
import ctypes as C
import os
import numpy as np
import pytest
import oracleport as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_block_encoder_golden():
    z = np.load(os.path.join(G, "ht_encode_blocks.npz"))
    for i in range(int(z["n"])):
        blk, kmax = z["blk_%d" % i], int(z["kmax_%d" % i])
        assert O.encode_block(blk, kmax - 1) == z["bytes_%d" % i].tobytes(), i


def test_block_decoder_golden_cleanup():
    z = np.load(os.path.join(G, "ht_encode_blocks.npz"))
    for i in range(int(z["n"])):
        blk, kmax, data = z["blk_%d" % i], int(z["kmax_%d" % i]), z["bytes_%d" % i].tobytes()
        h, w = blk.shape
        out, ok = O.decode_block(data, w, h, kmax - 1, 1, len(data), 0)
        assert ok and np.array_equal(out, z["dec_%d" % i]), i


def test_block_decoder_golden_sigprop_magref():
    """every coded block of the reference's in-tree test.j2c (77 of 89 have SPP / MRP passes)"""
    z = np.load(os.path.join(G, "ht_multipass_blocks.npz"))
    multi = 0
    for i in range(int(z["n"])):
        w, h, mm, npass, l1, l2, ok_ref = [int(v) for v in z["meta_%d" % i]]
        out, ok = O.decode_block(z["data_%d" % i].tobytes(), w, h, mm, npass, l1, l2)
        assert ok == bool(ok_ref) and np.array_equal(out, z["want_%d" % i]), i
        multi += npass > 1
    assert multi >= 70


def test_line_kernels_golden():
    z = np.load(os.path.join(G, "kernels.npz"))
    L = O.lib()
    for k in range(int(z["nh"])):
        w, even = [int(v) for v in z["rh_meta_%d" % k]]
        start = 0 if even else 1
        x = z["rh_src_%d" % k].astype(np.int32).copy()
        L.oj_rev53_fwd_line(O.ip(x), w, start)
        lo = x[(0 if even else 1)::2]; hi = x[(1 if even else 0)::2]
        assert np.array_equal(lo, z["rh_lo_%d" % k]) and np.array_equal(hi, z["rh_hi_%d" % k]), (w, even)
        L.oj_rev53_inv_line(O.ip(x), w, start)
        assert np.array_equal(x, z["rh_back_%d" % k])
        f = z["ih_src_%d" % k].astype(np.float32).copy()
        L.oj_irv97_fwd_line(O.fp(f), w, start)
        lo = f[(0 if even else 1)::2]; hi = f[(1 if even else 0)::2]
        assert np.array_equal(lo, z["ih_lo_%d" % k]) and np.array_equal(hi, z["ih_hi_%d" % k]), (w, even)
        L.oj_irv97_inv_line(O.fp(f), w, start)
        assert np.array_equal(f, z["ih_back_%d" % k])
    for s in range(2):
        for syn in (0, 1):
            a, b, d, r = [v.astype(np.int32).copy() for v in z["rv_%d_%d" % (s, syn)]]
            L.oj_rev_vert_step(s, O.ip(a), O.ip(b), O.ip(d), d.size, syn)
            assert np.array_equal(d, r)
    for s in range(4):
        for syn in (0, 1):
            a, b, d, r = [v.astype(np.float32).copy() for v in z["iv_%d_%d" % (s, syn)]]
            L.oj_irv_vert_step(s, O.fp(a), O.fp(b), O.fp(d), d.size, syn)
            assert np.array_equal(d, r)
    r_, g_, b_, y, cb, cr = [v.astype(np.int32).copy() for v in z["rct"]]
    oy, ocb, ocr = np.zeros_like(y), np.zeros_like(y), np.zeros_like(y)
    L.oj_rct_fwd(O.ip(r_), O.ip(g_), O.ip(b_), O.ip(oy), O.ip(ocb), O.ip(ocr), y.size)
    assert np.array_equal(oy, y) and np.array_equal(ocb, cb) and np.array_equal(ocr, cr)
    br, bg, bb = np.zeros_like(y), np.zeros_like(y), np.zeros_like(y)
    L.oj_rct_bwd(O.ip(y), O.ip(cb), O.ip(cr), O.ip(br), O.ip(bg), O.ip(bb), y.size)
    assert np.array_equal(br, r_) and np.array_equal(bg, g_) and np.array_equal(bb, b_)
    fr, fg, fb, fy, fcb, fcr, rr, rg, rb = [v.astype(np.float32).copy() for v in z["ict"]]
    oy, ocb, ocr = np.zeros_like(fy), np.zeros_like(fy), np.zeros_like(fy)
    L.oj_ict_fwd(O.fp(fr), O.fp(fg), O.fp(fb), O.fp(oy), O.fp(ocb), O.fp(ocr), fy.size)
    assert np.array_equal(oy, fy) and np.array_equal(ocb, fcb) and np.array_equal(ocr, fcr)
    o1, o2, o3 = np.zeros_like(fy), np.zeros_like(fy), np.zeros_like(fy)
    L.oj_ict_bwd(O.fp(fy), O.fp(fcb), O.fp(fcr), O.fp(o1), O.fp(o2), O.fp(o3), fy.size)
    assert np.array_equal(o1, rr) and np.array_equal(o2, rg) and np.array_equal(o3, rb)
    for bd, sg in ((8, 0), (12, 0), (10, 1)):
        v, f, f2, q = z["cvt_%d_%d" % (bd, sg)]
        of = np.zeros(v.size, np.float32); L.oj_irv_to_float(O.ip(v.astype(np.int32)), O.fp(of), bd, sg, v.size)
        assert np.array_equal(of, f.astype(np.float32))
        oq = np.zeros(v.size, np.int32); L.oj_irv_to_int(O.fp(f2.astype(np.float32)), O.ip(oq), bd, sg, v.size)
        assert np.array_equal(oq, q.astype(np.int32))
    for kmax in (9, 15, 20):
        v, sm, bk = z["txrev_%d" % kmax]
        o = np.zeros(v.size, np.uint32); mv = np.zeros(1, np.uint32)
        L.oj_rev_tx_to_cb32(O.ip(v.astype(np.int32)), O.up(o), kmax, v.size, O.up(mv))
        assert np.array_equal(o, sm.astype(np.uint32))
        ob_ = np.zeros(v.size, np.int32); L.oj_rev_tx_from_cb32(O.up(o), O.ip(ob_), kmax, v.size)
        assert np.array_equal(ob_, bk.astype(np.int32))
    for dinv in (1000000, 33000000):
        f, sm, bk = z["txirv_%d" % dinv]
        o = np.zeros(f.size, np.uint32); mv = np.zeros(1, np.uint32)
        L.oj_irv_tx_to_cb32(O.fp(f.astype(np.float32)), O.up(o), C.c_float(float(dinv)), f.size, O.up(mv), 0)   # generic: truncation
        assert np.array_equal(o, sm.astype(np.uint32))
        of = np.zeros(f.size, np.float32)
        L.oj_irv_tx_from_cb32(O.up(o), O.fp(of), C.c_float(1.0 / dinv), f.size)
        assert np.array_equal(of, bk.astype(np.float32))


def test_port_matches_reference_live_on_random_blocks(ref):
    rng = np.random.default_rng(77)
    for it in range(150):
        w = int(rng.integers(1, 65)); h = int(rng.integers(1, min(64, 4096 // w) + 1))
        kmax = int(rng.integers(1, 28))
        mag = np.minimum(np.abs(rng.laplace(0, 2 ** (kmax / 2.5), (h, w))).astype(np.uint64), (1 << kmax) - 1)
        if it % 4 == 0: mag = rng.integers(0, 1 << kmax, (h, w), dtype=np.uint64)
        if not mag.any(): continue
        blk = ((rng.integers(0, 2, (h, w), dtype=np.uint64) << 31) | (mag << (31 - kmax))).astype(np.uint32)
        want = ref.encode_block(blk, kmax - 1)
        assert O.encode_block(blk, kmax - 1) == want
        a, ok = O.decode_block(want, w, h, kmax - 1, 1, len(want), 0)
        b, ok2 = ref.decode_block(want, w, h, kmax - 1, 1, len(want), 0)
        assert ok == ok2 and np.array_equal(a, b)


def test_nlt_type3_restatement_matches_reference_live(ref):
    """a1's NLT type 3 variants: the plain-C restatement against the reference's generic line kernels"""
    L, R = O.lib(), ref.lib()
    rng = np.random.default_rng(21)
    for bd in (8, 12, 16, 24):
        lo, hi = -(1 << (bd - 1)), (1 << (bd - 1)) - 1
        v = rng.integers(lo, hi + 1, 4099).astype(np.int32)
        v[:4] = (lo, hi, -1, 0)
        a, b = np.zeros_like(v), np.zeros_like(v)
        L.oj_rev_convert_nlt3(O.ip(v), O.ip(a), bd, v.size)
        R.ojr_rev_convert_nlt_type3(O.ip(v), O.ip(b), C.c_int64((1 << (bd - 1)) + 1), v.size)
        assert np.array_equal(a, b)
        back = np.zeros_like(v); L.oj_rev_convert_nlt3(O.ip(a), O.ip(back), bd, v.size)
        assert np.array_equal(back, v)                               # an involution
        for sg in (1, 0):
            src = v if sg else (v - lo).astype(np.int32)
            fa, fb = np.zeros(v.size, np.float32), np.zeros(v.size, np.float32)
            L.oj_irv_to_float_nlt3(O.ip(src), O.fp(fa), bd, sg, v.size)
            R.ojr_irv_convert_to_float_nlt_type3(O.ip(src), O.fp(fb), bd, sg, v.size)
            assert np.array_equal(fa, fb)
            f = (rng.random(v.size, dtype=np.float32) - 0.5) * 1.2       # overshoots the range on both sides
            ia, ib = np.zeros(v.size, np.int32), np.zeros(v.size, np.int32)
            L.oj_irv_to_int_nlt3(O.fp(f), O.ip(ia), bd, sg, v.size)
            R.ojr_irv_convert_to_integer_nlt_type3(O.fp(f), O.ip(ib), bd, sg, v.size)
            assert np.array_equal(ia, ib)

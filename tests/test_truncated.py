"""The reference's tests/test_truncated_decode.cpp, restated for the B200 path: a codestream cut to k/16 of
its length must decode (to something) in resilient mode for every k, and a truncation that the
non-resilient parser detects must raise instead -- the flag, and only the flag, decides."""
import numpy as np
import pytest
import openjph_b200 as ob
import cases

W = H = 256
NUM_CUTS = 16


def _codestream(lib):
    y, x = np.mgrid[0:H, 0:W]
    img = ((x * 7 + y * 13 + ((x * y) >> 3)) & 0xFF).astype(np.int32)        # test_truncated_decode.cpp:78-81
    p = ob.make_params(W, H, 1, 8, num_decomps=5, block=(64, 64), reversible=True)
    return ob.Encoder(p, ob.I32, lib=lib).encode([img]), img


def _check(lib, ref=None):
    full, img = _codestream(lib)
    assert len(full) > NUM_CUTS * 64
    for resilient in (False, True):
        out = ob.Decoder(resilient=resilient, lib=lib).decode(full)
        assert np.array_equal(out[0], img)
    detected = 0
    for cut in range(1, NUM_CUTS):
        part = full[:len(full) * cut // NUM_CUTS]
        out = ob.Decoder(resilient=True, lib=lib).decode(part)               # must not raise
        assert out[0].shape == (H, W)
        if ref is not None:                   # what survives is what the reference reconstructs from the same bytes
            want, _ = ref.decode(part, resilient=True)
            assert np.array_equal(out[0], want[0]), "resilient decode of cut %d/%d differs from the reference" % (cut, NUM_CUTS)
        raised = False
        try:
            ob.Decoder(resilient=False, lib=lib).decode(part)
        except ob.OjphError:
            raised = True
            detected += 1
        if ref is not None:                   # the same cuts are detected as by the reference's parser
            ref_raised = False
            try:
                ref.decode(part, resilient=False)
            except Exception:
                ref_raised = True
            assert raised == ref_raised, "cut %d/%d" % (cut, NUM_CUTS)
    assert detected > 0, "no truncation was detected by the parser"
    return detected


def test_truncated_codestreams_emulator(emu_lib, ref):
    _check(emu_lib, ref)


@pytest.mark.gpu
def test_truncated_codestreams_gpu(gpu_lib, ref):
    _check(None, ref)


def test_corrupted_tile_data_is_survived(emu_lib, ref):
    """random byte errors inside the tile data: a resilient decode must come back (failed blocks are zero-filled)
    without touching memory outside its buffers (the emulator's guard zones abort on that)"""
    rng = np.random.default_rng(11)
    for kw in (dict(width=200, height=150, num_comps=3, bit_depth=8, num_decomps=3, reversible=True, color_transform=True),
               dict(width=700, height=100, num_comps=1, bit_depth=10, num_decomps=2, reversible=True, block=(256, 16)),
               dict(width=128, height=96, num_comps=3, bit_depth=12, num_decomps=4, reversible=False, color_transform=True, qfactor=90)):
        p = cases.make(kw)
        cs = bytearray(ref.encode(p, cases.frame_for(p)))
        sod = cs.index(b"\xff\x93") + 2
        for _ in range(12):
            bad = bytearray(cs)
            for _ in range(int(rng.integers(1, 30))):
                bad[int(rng.integers(sod, len(bad) - 2))] = int(rng.integers(0, 256))
            out = ob.Decoder(resilient=True, lib=emu_lib).decode(bytes(bad))
            assert len(out) == p.num_comps and out[0].shape == (kw["height"], kw["width"])


def _check_failed_setup_is_not_remembered(lib):
    """a read_headers that fails after the header bytes were accepted (10-bit stream, 8-bit container) must fail
    again when repeated, and must not poison the decoder for a valid call"""
    img = (np.arange(64 * 64, dtype=np.int32).reshape(64, 64) * 5) & 1023
    p = ob.make_params(64, 64, 1, 10, num_decomps=2, reversible=True)
    cs = ob.Encoder(p, ob.I32, lib=lib).encode([img])
    dec = ob.Decoder(lib=lib)
    for _ in range(2):
        with pytest.raises(ob.OjphError):
            dec.decode(cs, sample_type=ob.U8)
    out = dec.decode(cs, sample_type=ob.U16)
    assert np.array_equal(out[0], img)


def test_failed_setup_is_not_remembered_emulator(emu_lib):
    _check_failed_setup_is_not_remembered(emu_lib)


@pytest.mark.gpu
def test_failed_setup_is_not_remembered_gpu(gpu_lib):
    _check_failed_setup_is_not_remembered(None)


def test_frame_info_reports_colour_transform(emu_lib):
    """Decoder::info: an RCT stream says so (the facade's is_using_color_transform / is_planar default hang on it)"""
    rng = np.random.default_rng(3)
    planes = [rng.integers(0, 256, (40, 56)).astype(np.int32) for _ in range(3)]
    for ct in (True, False):
        p = ob.make_params(56, 40, 3, 8, num_decomps=2, reversible=True, color_transform=ct)
        cs = ob.Encoder(p, ob.I32, lib=emu_lib).encode(planes)
        dec = ob.Decoder(lib=emu_lib)
        info = dec.read_headers(cs)
        assert bool(info.color_transform) == ct

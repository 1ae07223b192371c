"""NLT type 3 (param_nlt, the -nlt_type3 path of ojph_compress; rev_convert_nlt_type3 and the irreversible
variants, ojph_colour.cpp:273-440) and the extra COM segments of write_headers(file, comments, n):
codestreams byte-identical to the reference's (reversible), samples equal / within +-1 (9/7)."""
import numpy as np
import pytest
import cases
import openjph_b200 as ob


def _signed_frame(p, seed=7):
    rng = np.random.default_rng(seed)
    from openjph_b200.codestream import comp_dims
    out = []
    for c, (w, h) in enumerate(comp_dims(p)):
        bd = p.bit_depth[c]
        if p.is_signed[c]:
            lo, hi = -(1 << (bd - 1)), (1 << (bd - 1)) - 1
        else:
            lo, hi = 0, (1 << bd) - 1
        y, x = np.mgrid[0:h, 0:w]
        base = (np.sin(x / 9.0 + c) * np.cos(y / 7.0) * 0.45 * (hi - lo) + (hi + lo) / 2.0)
        a = np.clip(np.rint(base + rng.normal(0, 3.0, (h, w))), lo, hi).astype(np.int32)
        a[0, 0], a[-1, -1] = lo, hi                              # the extremes (lo maps to -1 under type 3)
        out.append(a)
    return out


NLT_CASES = {
    # all components alike: one ALL_COMPS segment
    "all_rev_rct": (dict(width=120, height=90, num_comps=3, bit_depth=12, is_signed=True, num_decomps=3, reversible=True,
                         color_transform=True), {"all": 3}),
    "all_irv_ict": (dict(width=120, height=90, num_comps=3, bit_depth=10, is_signed=True, num_decomps=3, reversible=False,
                         color_transform=True, qstep=0.002), {"all": 3}),
    # exception for one component (type 0 entry is written too)
    "all_but_one": (dict(width=100, height=70, num_comps=3, bit_depth=9, is_signed=True, num_decomps=2, reversible=True,
                         planar=1), {"all": 3, 1: 0}),
    # per-component calls, out of order
    "per_comp": (dict(width=100, height=70, num_comps=4, bit_depth=8, is_signed=True, num_decomps=2, reversible=True,
                      planar=1), {2: 3, 0: 3}),
    "gray_L0": (dict(width=64, height=48, num_comps=1, bit_depth=16, is_signed=True, num_decomps=0, reversible=True), {"all": 3}),
    "irv_gray_tiles": (dict(width=150, height=100, num_comps=1, bit_depth=12, is_signed=True, num_decomps=3, reversible=False,
                            tile=(64, 64), qstep=0.001), {0: 3}),
    # unsigned content: the marker is written, the samples are untouched
    "unsigned": (dict(width=80, height=60, num_comps=3, bit_depth=8, num_decomps=2, reversible=True, color_transform=True), {"all": 3}),
}


def _check_nlt(lib, ref, name):
    kw, nlt = NLT_CASES[name]
    kw = dict(kw)
    w, h, nc, bd = kw.pop("width"), kw.pop("height"), kw.pop("num_comps"), kw.pop("bit_depth")
    p = ob.make_params(w, h, nc, bd, nlt=nlt, **kw)
    frame = _signed_frame(p)
    want = ref.encode(p, frame)
    got = ob.Encoder(p, ob.I32, lib=lib).encode(frame)
    assert b"\xff\x76" in want[:300]
    if p.reversible:
        assert got == want
    else:
        assert len(got) == len(want) and got[:want.index(b"\xff\x90")] == want[:want.index(b"\xff\x90")]
    dec = ob.Decoder(lib=lib)
    for cs in (want, got):
        out = dec.decode(cs)
        refout, _ = ref.decode(cs)
        for c, (a, b) in enumerate(zip(out, refout)):
            d = int(np.abs(a.astype(np.int64) - b).max())
            assert d <= (0 if p.reversible else 1), (name, c, d)
        if p.reversible and name != "gray_L0":   # (-2^(B-1) at zero levels overflows K_max in the reference too)
            for a, b in zip(out, frame):
                assert np.array_equal(a, b)
    types = [dec.info.nlt_type[c] for c in range(nc)]
    exp = [nlt.get(c, nlt.get("all", 0)) for c in range(nc)]
    assert types == exp


@pytest.mark.parametrize("name", list(NLT_CASES))
def test_nlt_type3_emulator(name, emu_lib, ref):
    _check_nlt(emu_lib, ref, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(NLT_CASES))
def test_nlt_type3_gpu(name, gpu_lib, ref):
    _check_nlt(None, ref, name)


def test_nlt_mixed_depths_split_into_components(emu_lib, ref):
    """ALL_COMPS type 3 over components of different depth: the library writes one segment per component"""
    p = ob.make_params(64, 48, 3, 8, is_signed=True, num_decomps=2, reversible=True, planar=1, nlt={"all": 3})
    p.bit_depth[1] = 10
    frame = _signed_frame(p)
    want = ref.encode(p, frame)
    got = ob.Encoder(p, ob.I32, lib=emu_lib).encode(frame)
    assert got == want and want[:300].count(b"\xff\x76") == 3
    out = ob.Decoder(lib=emu_lib).decode(got)
    for a, b in zip(out, frame):
        assert np.array_equal(a, b)


def test_nlt_rejects_other_types(emu_lib):
    p = ob.make_params(32, 32, 1, 8, nlt={"all": 1})
    with pytest.raises(ob.OjphError):
        ob.Encoder(p, ob.I32, lib=emu_lib)


def _check_comments(lib, ref):
    p = cases.make(dict(width=100, height=70, num_comps=3, bit_depth=8, num_decomps=2, reversible=True, color_transform=True))
    frame = cases.frame_for(p)
    comments = ["a text comment", bytes(range(0, 40)), ""]
    want = ref.encode(p, frame, comments=comments)
    got = ob.Encoder(p, ob.I32, lib=lib, comments=comments).encode(frame)
    assert got == want
    out = ob.Decoder(lib=lib).decode(got)
    for a, b in zip(out, frame):
        assert np.array_equal(a, b)


def test_comments_emulator(emu_lib, ref):
    _check_comments(emu_lib, ref)


@pytest.mark.gpu
def test_comments_gpu(gpu_lib, ref):
    _check_comments(None, ref)

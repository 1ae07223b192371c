"""Error behaviour of the drop-in: parameter sets and codestreams the reference rejects must be rejected here
with the reference's error code (OJPH_ERROR codes travel in the message: "ojph error 0x000500xx: ...").
The reference's code is captured through its message-handler hook (oracle/ref_harness.cpp)."""
import re
import numpy as np
import pytest
import cases
import openjph_b200 as ob


def _ref_code(ref, fn):
    L = ref.lib()
    L.ojr_capture_errors(1)
    L.ojr_last_error_code.restype = __import__("ctypes").c_uint32
    try:
        fn()
        return None
    except RuntimeError:
        return int(L.ojr_last_error_code())
    finally:
        L.ojr_capture_errors(0)


def _our_code(fn):
    try:
        fn()
        return None
    except ob.OjphError as e:
        m = re.search(r"0x([0-9A-Fa-f]{8})", str(e))
        return int(m.group(1), 16) if m else -1


GOOD = dict(width=64, height=48, num_comps=3, bit_depth=8, num_decomps=3, reversible=True)
BAD_PARAMS = {
    "colour_transform_two_comps": dict(num_comps=2, color_transform=True),
    "colour_transform_subsampled": dict(color_transform=True, subsampling=[(1, 1), (2, 2), (2, 2)]),
    "tile_offset_beyond_image_offset": dict(offset=(2, 2), tile_offset=(3, 3), tile=(32, 32)),
    "first_tile_misses_image": dict(offset=(40, 0), tile=(32, 32), width=80),
    "block_area_too_large": dict(block=(128, 64)),
    "block_not_power_of_two": dict(block=(48, 32)),
    "block_too_small": dict(block=(2, 64)),
    "precinct_not_power_of_two": dict(precincts=[(100, 100)]),
    "precinct_too_small": dict(precincts=[(64, 64), (1, 1)]),
    "rpcl_subsampling_not_power_of_two": dict(prog_order="RPCL", subsampling=[(1, 1), (3, 1), (3, 1)], planar=1),
    "planar_with_colour_transform": dict(color_transform=True, planar=1),
    "qfactor_out_of_range": dict(reversible=False, qfactor=101),
    "too_many_decompositions": dict(num_decomps=33),
    "image_offset_beyond_extent": dict(offset=(64, 0)),
    "colour_transform_mixed_depth": dict(color_transform=True, _depths=[8, 10, 8]),
    "colour_transform_mixed_sign": dict(color_transform=True, _signs=[0, 1, 0]),
    "precinct_zero": dict(precincts=[(0, 64)]),
    "nlt_unsupported_type": dict(nlt={"all": 2}),
}


@pytest.mark.parametrize("name", list(BAD_PARAMS))
def test_rejected_parameters_same_code(name, emu_lib, ref):
    kw = dict(GOOD); kw.update(BAD_PARAMS[name])
    w, h, nc, bd = kw.pop("width"), kw.pop("height"), kw.pop("num_comps"), kw.pop("bit_depth")
    depths, signs = kw.pop("_depths", None), kw.pop("_signs", None)
    p = ob.make_params(w, h, nc, bd, **kw)
    for c, d in enumerate(depths or []): p.bit_depth[c] = d
    for c, sg in enumerate(signs or []): p.is_signed[c] = sg
    frame = [np.zeros((p.height - p.off_y if p.height > p.off_y else 1, max(1, p.width - p.off_x)), np.int32) for _ in range(nc)]
    rc = _ref_code(ref, lambda: ref.encode(p, frame))
    oc = _our_code(lambda: ob.Encoder(p, ob.I32, lib=emu_lib))
    assert rc is not None, "the reference accepts this parameter set"
    assert oc == rc, "reference 0x%08X, here %s" % (rc, "accepted" if oc is None else "0x%08X" % oc)


def _stream(ref):
    p = ob.make_params(96, 64, 3, 8, num_decomps=2, reversible=True, color_transform=True)
    return bytearray(ref.encode(p, cases.frame_for(p)))


def _patch(cs, marker, offset, value):
    i = cs.index(marker)
    cs[i + offset] = value
    return bytes(cs)


BAD_STREAMS = {
    "not_a_codestream": lambda cs: b"\x00\x01" + bytes(cs[2:]),
    "rsiz_without_ht": lambda cs: _patch(cs, b"\xff\x51", 4, 0x00),            # Rsiz bit 14 cleared
    "zero_subsampling": lambda cs: _patch(cs, b"\xff\x51", 2 + 2 + 2 + 32 + 2 + 1, 0),   # XRsiz of component 0
    "no_cod": lambda cs: bytes(cs).replace(b"\xff\x52", b"\xff\x64", 1),        # COD turned into a COM
    "ends_before_tiles": lambda cs: bytes(cs[:cs.index(b"\xff\x90")]),
    "siz_length_wrong": lambda cs: _patch(cs, b"\xff\x51", 3, 0x30),
    "csiz_mismatch": lambda cs: _patch(cs, b"\xff\x51", 2 + 2 + 2 + 32 + 1, 2),
    "two_quality_layers": lambda cs: _patch(cs, b"\xff\x52", 2 + 2 + 1 + 1 + 1, 2),
    "cod_not_ht": lambda cs: _patch(cs, b"\xff\x52", 2 + 2 + 1 + 4 + 1 + 1 + 1 + 1, 0x00),
    "cod_bad_wavelet": lambda cs: _patch(cs, b"\xff\x52", 2 + 2 + 1 + 4 + 1 + 1 + 1 + 2, 5),
    "cod_too_many_levels": lambda cs: _patch(cs, b"\xff\x52", 2 + 2 + 1 + 4, 40),
    "qcd_bad_style": lambda cs: _patch(cs, b"\xff\x5c", 4, 0x41),
    "no_qcd": lambda cs: bytes(cs).replace(b"\xff\x5c", b"\xff\x64", 1),
    "tile_index_out_of_range": lambda cs: _patch(cs, b"\xff\x90", 5, 9),
    "unsupported_nlt_type": lambda cs: bytes(cs[:cs.index(b"\xff\x64")]) + b"\xff\x76\x00\x06\xff\xff\x07\x01" + bytes(cs[cs.index(b"\xff\x64"):]),
}


@pytest.mark.parametrize("name", list(BAD_STREAMS))
def test_rejected_codestreams_same_code(name, emu_lib, ref):
    bad = BAD_STREAMS[name](_stream(ref))
    rc = _ref_code(ref, lambda: ref.decode(bad))
    oc = _our_code(lambda: ob.Decoder(lib=emu_lib).decode(bad))
    assert rc is not None, "the reference accepts this codestream"
    assert oc == rc, "reference 0x%08X, here %s" % (rc, "accepted" if oc is None else "0x%08X" % oc)

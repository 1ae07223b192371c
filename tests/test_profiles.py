"""codestream::set_profile("IMF" / "BROADCAST") (ojph_codestream_local.cpp:292-535, :1124-1133): the profile
rules reject the same parameter sets as the reference and, where they pass, force TLM and tile-part division by
components -- the codestream must still be byte-identical."""
import numpy as np
import pytest
import cases
import openjph_b200 as ob

GOOD = dict(num_decomps=5, reversible=True, color_transform=True, block=(32, 32), precincts=[(128, 128), (256, 256)], prog_order="CPRL")
CASES = {
    # name: (profile, width, height, comps, depth, overrides, accepted)
    "imf_rev": ("IMF", 300, 200, 3, 10, {}, True),
    "imf_irv": ("IMF", 300, 200, 3, 12, dict(reversible=False, qstep=0.01), True),
    "imf_422": ("IMF", 256, 128, 3, 8, dict(color_transform=False, subsampling=[(1, 1), (2, 1), (2, 1)], planar=1), True),
    "imf_tiles_1024": ("IMF", 1100, 1030, 1, 8, dict(color_transform=False, tile=(1024, 1024), num_decomps=4), True),
    "imf_bad_block": ("IMF", 300, 200, 3, 10, dict(block=(64, 64)), False),
    "imf_bad_order": ("IMF", 300, 200, 3, 10, dict(prog_order="RPCL"), False),
    "imf_bad_depth": ("IMF", 300, 200, 3, 7, {}, False),
    "imf_bad_offset": ("IMF", 300, 200, 3, 10, dict(offset=(2, 0)), False),
    "imf_no_levels": ("IMF", 300, 200, 3, 10, dict(num_decomps=0, precincts=[(128, 128)]), False),
    "imf_lossy_tiles": ("IMF", 300, 200, 3, 10, dict(reversible=False, tile=(256, 256)), False),
    "imf_bad_tile": ("IMF", 300, 200, 3, 10, dict(tile=(200, 200)), False),
    "bc_rev": ("BROADCAST", 320, 180, 3, 10, dict(block=(64, 64)), True),
    "bc_4comp_4tiles": ("BROADCAST", 320, 180, 4, 12, dict(block=(128, 32), tile=(160, 90)), True),
    "bc_bad_depth": ("BROADCAST", 320, 180, 3, 14, {}, False),
    "bc_bad_levels": ("BROADCAST", 320, 180, 3, 10, dict(num_decomps=6, precincts=[(128, 128), (256, 256)]), False),
    "bc_bad_tiles": ("BROADCAST", 320, 180, 3, 10, dict(tile=(160, 180)), False),
    "bc_bad_precincts": ("BROADCAST", 320, 180, 3, 10, dict(precincts=[(128, 128), (128, 128)]), False),
}


@pytest.mark.parametrize("name", list(CASES))
def test_profile_rules_and_bytes(name, emu_lib, ref):
    profile, w, h, nc, bd, over, accepted = CASES[name]
    kw = dict(GOOD); kw.update(over)
    if nc != 3:
        kw["color_transform"] = kw.get("color_transform", False) and nc >= 3
    p = ob.make_params(w, h, nc, bd, profile=profile, **kw)
    frame = cases.frame_for(p)
    try:
        want = ref.encode(p, frame)
        ref_ok = True
    except RuntimeError:
        ref_ok = False
    assert ref_ok == accepted, "the case table disagrees with the reference"
    if not accepted:
        with pytest.raises(ob.OjphError):
            ob.Encoder(p, ob.I32, lib=emu_lib)
        return
    got = ob.Encoder(p, ob.I32, lib=emu_lib).encode(frame)
    assert b"\xff\x55" in got[:400]                               # the profile switched TLM on
    if p.reversible:
        assert got == want
        out = ob.Decoder(lib=emu_lib).decode(got)
        for a, b in zip(out, frame):
            assert np.array_equal(a, b)
    else:
        assert len(got) == len(want) and got[:got.index(b"\xff\x90")] == want[:want.index(b"\xff\x90")]


def test_unknown_profile(emu_lib):
    p = ob.make_params(64, 64, 1, 8)
    p.profile = 7
    with pytest.raises(ob.OjphError):
        ob.Encoder(p, ob.I32, lib=emu_lib)

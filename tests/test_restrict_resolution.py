"""codestream::restrict_input_resolution (the -skip_res option of ojph_expand): reduced-resolution
reconstruction and undecoded top resolutions, against the reference for the same (read, recon) pairs."""
import numpy as np
import pytest
import cases
import openjph_b200 as ob

CASES = {
    "rgb_rct_L4": dict(width=301, height=213, num_comps=3, bit_depth=8, num_decomps=4, reversible=True, color_transform=True),
    "gray_tiles_off": dict(width=300, height=200, num_comps=1, bit_depth=10, num_decomps=3, reversible=True,
                           tile=(128, 100), offset=(7, 5), tile_offset=(3, 2)),
    "irv_sub420": dict(width=200, height=150, num_comps=3, bit_depth=8, num_decomps=3, reversible=False, qstep=0.02,
                       subsampling=[(1, 1), (2, 2), (2, 2)], planar=1),
    "irv_ict_L5": dict(width=256, height=200, num_comps=3, bit_depth=12, num_decomps=5, reversible=False,
                       color_transform=True, qfactor=85),
}
SKIPS = [(1, 1), (2, 1), (2, 2), (3, 3), (3, 0), (1, 0)]


def _check(lib, ref, name, skip):
    kw = CASES[name]
    if skip[0] > kw["num_decomps"]:
        pytest.skip("more skipped resolutions than decomposition levels")
    p = cases.make(kw)
    frame = cases.frame_for(p)
    cs = ref.encode(p, frame)
    want, info = ref.decode_restricted(cs, *skip)
    tol = 0 if kw["reversible"] else 1
    dec = ob.Decoder(lib=lib)
    got = dec.decode(cs, skip=skip)
    assert [g.shape for g in got] == [w.shape for w in want]
    for c, (a, b) in enumerate(zip(got, want)):
        d = np.abs(a.astype(np.int64) - b)
        assert d.max() <= tol, (name, skip, c, int(d.max()))
    # back to full resolution with the same object
    full = dec.decode(cs, skip=(0, 0))
    ref_full, _ = ref.decode(cs)
    for a, b in zip(full, ref_full):
        assert np.abs(a.astype(np.int64) - b).max() <= tol


@pytest.mark.parametrize("skip", SKIPS)
@pytest.mark.parametrize("name", list(CASES))
def test_restrict_resolution_emulator(name, skip, emu_lib, ref):
    _check(emu_lib, ref, name, skip)


@pytest.mark.gpu
@pytest.mark.parametrize("skip", SKIPS)
@pytest.mark.parametrize("name", list(CASES))
def test_restrict_resolution_gpu(name, skip, gpu_lib, ref):
    _check(None, ref, name, skip)


def test_restrict_resolution_errors(emu_lib, ref):
    p = cases.make(CASES["rgb_rct_L4"])
    cs = ref.encode(p, cases.frame_for(p))
    dec = ob.Decoder(lib=emu_lib)
    dec.read_headers(cs)
    with pytest.raises(ob.OjphError):
        dec.restrict_input_resolution(1, 2)        # reconstruction above what is read
    with pytest.raises(ob.OjphError):
        dec.restrict_input_resolution(5, 5)        # more than the decomposition levels


def test_restrict_resolution_mixed_decomps_emulator(emu_lib, ref):
    """per-component decomposition counts: fine while every component has the skipped levels, refused beyond"""
    p = ob.make_params(200, 150, 3, 8, num_decomps=4, reversible=True, planar=1,
                       coc={1: dict(reversible=True, num_decomps=2, block=(32, 32)), 2: dict(reversible=True, num_decomps=3)})
    cs = ref.encode(p, cases.frame_for(p))
    for skip in ((1, 1), (2, 2), (2, 0)):
        want, _ = ref.decode_restricted(cs, *skip)
        got = ob.Decoder(lib=emu_lib).decode(cs, skip=skip)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    with pytest.raises(ob.OjphError):
        ob.Decoder(lib=emu_lib).decode(cs, skip=(3, 1))

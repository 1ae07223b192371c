"""The reference's tests/test_mixed_coc.cpp (TestMixedCOC.FourCompMixedReversibility), restated: four
components, three coded with the irreversible 9/7 wavelet through COD and one with the reversible 5/3
through its own COC (+ QCC).  The reversible component must decode exactly; the per-component settings
must survive the codestream.  Also: components with their own number of decompositions / block size."""
import numpy as np
import pytest
import cases
import openjph_b200 as ob

W = H = 64


def _frame():
    y, x = np.mgrid[0:H, 0:W]
    ramp = ((x + y * W) % 256).astype(np.int32)                 # test_mixed_coc.cpp:100-106
    return [np.zeros((H, W), np.int32)] * 3 + [ramp]


def _params():
    return ob.make_params(W, H, 4, 8, num_decomps=5, block=(64, 64), reversible=False, color_transform=False,
                          qstep=0.01, planar=1, coc={3: dict(reversible=True)})


def _check_mixed(lib, ref):
    p, frame = _params(), _frame()
    want = ref.encode(p, frame)
    got = ob.Encoder(p, ob.I32, lib=lib).encode(frame)
    assert got == want                                          # all-zero 9/7 components: nothing rounding-dependent
    assert b"\xff\x53" in got[:200]                             # a COC segment is there
    for cs in (want, got):
        out = ob.Decoder(lib=lib).decode(cs)
        assert np.array_equal(out[3], frame[3])                 # the reversible component is exact
        refout, _ = ref.decode(cs)
        for a, b in zip(out, refout):
            assert np.array_equal(a, b)


def _check_varied(lib, ref):
    """natural content; component 1 has 2 levels and 32x32 blocks, component 2 is reversible with 3 levels"""
    p = ob.make_params(200, 150, 3, 8, num_decomps=4, reversible=False, qstep=0.02, planar=1,
                       coc={1: dict(reversible=False, num_decomps=2, block=(32, 32)), 2: dict(reversible=True, num_decomps=3)})
    frame = cases.frame_for(p)
    want = ref.encode(p, frame)
    refout, _ = ref.decode(want)
    out = ob.Decoder(lib=lib).decode(want)
    assert np.array_equal(out[2], frame[2]) and np.array_equal(refout[2], frame[2])
    got = ob.Encoder(p, ob.I32, lib=lib).encode(frame)
    assert len(got) == len(want)
    cross, _ = ref.decode(got)
    assert np.array_equal(cross[2], frame[2])
    for c in (0, 1):
        m_ref, p_ref = cases.mse_pae(refout[c], frame[c])
        for planes in (out, cross):
            m, pa = cases.mse_pae(planes[c], frame[c])
            assert abs(m - m_ref) / (m_ref + 0.01) < 0.01 and abs(pa - p_ref) <= 1


def test_mixed_reversibility_emulator(emu_lib, ref):
    _check_mixed(emu_lib, ref)


def test_per_component_styles_emulator(emu_lib, ref):
    _check_varied(emu_lib, ref)


@pytest.mark.gpu
def test_mixed_reversibility_gpu(gpu_lib, ref):
    _check_mixed(None, ref)


@pytest.mark.gpu
def test_per_component_styles_gpu(gpu_lib, ref):
    _check_varied(None, ref)


def test_read_side_getters(emu_lib, ref):
    """param_cod / param_siz getters after read_headers, per component (COC aware)"""
    p = ob.make_params(300, 200, 3, 8, num_decomps=4, reversible=False, qstep=0.02, planar=1, tile=(256, 128),
                       precincts=[(64, 64), (128, 128)], prog_order="PCRL", block=(32, 16),
                       coc={1: dict(reversible=True, num_decomps=2, block=(64, 64))})
    cs = ref.encode(p, cases.frame_for(p))
    dec = ob.Decoder(lib=emu_lib)
    dec.read_headers(cs)
    a, b = dec.coding_style(0), dec.coding_style(1)
    assert (a.num_decomps, a.reversible, a.block_w, a.block_h, a.prog_order) == (4, 0, 32, 16, 3)
    assert (a.precinct_w[0], a.precinct_w[1], a.precinct_w[4]) == (64, 128, 128) and a.num_layers == 1
    assert (b.num_decomps, b.reversible, b.block_w, b.block_h) == (2, 1, 64, 64) and b.precinct_w[0] == 32768
    assert (a.tile_w, a.tile_h, a.tile_off_x) == (256, 128, 0)


def test_per_component_precincts(emu_lib, ref):
    """param_cod::set_precinct_size(comp_idx, ...): the component's COC carries its own precinct sizes"""
    for po in ("RPCL", "CPRL", "LRCP"):
        p = ob.make_params(300, 260, 3, 8, num_decomps=3, reversible=True, planar=1, prog_order=po, precincts=[(128, 128)],
                           coc={1: dict(reversible=True, num_decomps=3, precincts=[(64, 64), (128, 64)]),
                                2: dict(reversible=True, num_decomps=2, block=(32, 32), precincts=[(256, 256)])})
        frame = cases.frame_for(p)
        want = ref.encode(p, frame)
        got = ob.Encoder(p, ob.I32, lib=emu_lib).encode(frame)
        assert got == want, po
        out = ob.Decoder(lib=emu_lib).decode(want)
        for a, b in zip(out, frame):
            assert np.array_equal(a, b)

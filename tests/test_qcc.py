"""Per-component quantisation (param_qcd::set_irrev_quant(comp, delta) / set_qfactor(comp, ctype, q),
ojph_params.cpp:2011-2035): same QCD / QCC marker segments as the reference (main header byte-identical),
same stream length, decodes within the 9/7 tolerance.  Includes the reference's quirk that a per-component
delta lands on the global QCD unless the component already has a QCC."""
import numpy as np
import pytest
import cases
import openjph_b200 as ob

BASE = dict(num_decomps=3, reversible=False, planar=1)
CASES = {
    "qfactor_per_comp": dict(qcc=[("qfactor", 0, 0, 90), ("qfactor", 1, 1, 60), ("qfactor", 2, 2, 60)]),
    "qfactor_one_comp": dict(qcc=[("qfactor", 1, 0, 75)], qstep=0.01),
    "delta_without_qcc": dict(qcc=[("qstep", 1, 0.03)], qstep=0.005),          # overwrites the global step
    "delta_after_qfactor": dict(qcc=[("qfactor", 2, 0, 50), ("qstep", 2, 0.02), ("qstep", 0, 0.004)]),
    "global_qfactor_plus_comp": dict(qfactor=85, qcc=[("qfactor", 1, 1, 40)], color_transform=True, planar=0),
}


def _check(lib, ref, name):
    kw = dict(BASE); kw.update(CASES[name])
    p = ob.make_params(160, 120, 3, 10, **kw)
    frame = cases.frame_for(p)
    want = ref.encode(p, frame)
    got = ob.Encoder(p, ob.I32, lib=lib).encode(frame)
    sot = want.index(b"\xff\x90")
    assert got[:sot] == want[:sot], "main header (QCD / QCC segments) differs"
    assert abs(len(got) - len(want)) <= max(4, len(want) // 1000)
    ref_planes, _ = ref.decode(want)
    for cs in (want, got):
        out = ob.Decoder(lib=lib).decode(cs)
        for c in range(3):
            m_ref, p_ref = cases.mse_pae(ref_planes[c], frame[c])
            m, pa = cases.mse_pae(out[c], frame[c])
            assert abs(m - m_ref) <= max(0.01 * m_ref, 0.01) and abs(pa - p_ref) <= max(1, 0.01 * p_ref), (name, c, m, m_ref)


@pytest.mark.parametrize("name", list(CASES))
def test_qcc_emulator(name, emu_lib, ref):
    _check(emu_lib, ref, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_qcc_gpu(name, gpu_lib, ref):
    _check(None, ref, name)

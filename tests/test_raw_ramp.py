"""The reference's self-contained known-answer tests SimpleEncRev53Raw{8,16,24,32}{Signed,Unsigned}
(tests/test_executables.cpp:1698-1728; frame generator :297-344): a 256x256 one-component frame holding a
ramp from the smallest to the largest value of the sample type, reversible 5/3, default parameters of
ojph_compress -- lossless exact, and here also codestream-identical to the reference.  The 32-bit variants run
on the 64-bit coefficient path (general DWT kernels on int64, 64-bit block coders)."""
import numpy as np
import pytest
import openjph_b200 as ob

W = H = 256


def ramp(bit_depth, is_signed):
    n = W * H
    upper = (1 << (bit_depth - 1)) - 1 if is_signed else (1 << bit_depth) - 1
    lower = -(1 << (bit_depth - 1)) if is_signed else 0
    idx = np.arange(n, dtype=np.int64)
    v = lower + idx * (upper - lower) // (n - 1)
    return (v & 0xFFFFFFFF).astype(np.uint32).astype(np.int32).reshape(H, W)      # 32-bit unsigned: the bit pattern in an si32 line


def _roundtrip(lib, ref, bit_depth, is_signed):
    img = ramp(bit_depth, is_signed)
    p = ob.make_params(W, H, 1, bit_depth, is_signed=is_signed, num_decomps=5, reversible=True)
    cs = ob.Encoder(p, ob.I32, lib=lib).encode([img])
    assert cs == ref.encode(p, [img])
    out = ob.Decoder(lib=lib).decode(cs)
    assert np.array_equal(out[0], img)
    refout, _ = ref.decode(cs)
    assert np.array_equal(refout[0], img)


@pytest.mark.parametrize("bit_depth,is_signed", [(8, True), (8, False), (16, True), (16, False), (24, True), (24, False), (32, True), (32, False)])
def test_simple_enc_rev53_raw_emulator(bit_depth, is_signed, emu_lib, ref):
    _roundtrip(emu_lib, ref, bit_depth, is_signed)


@pytest.mark.gpu
@pytest.mark.parametrize("bit_depth,is_signed", [(8, True), (8, False), (16, True), (16, False), (24, True), (24, False), (32, True), (32, False)])
def test_simple_enc_rev53_raw_gpu(bit_depth, is_signed, gpu_lib, ref):
    _roundtrip(None, ref, bit_depth, is_signed)

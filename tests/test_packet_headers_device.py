"""Packet headers on the device (pkt_headers.cu) against the host writer (ojb_layout.cpp, the
restatement of precinct::prepare_precinct, src/core/codestream/ojph_precinct.cpp:94-278, that every byte-identical test
of round 1 pinned against the reference): same bytes for every configuration, including the ones that stress what the
device path does differently -- tag trees over odd-width grids (the reference's aliasing quirk), bands and packets with
nothing in them, long Lblock runs and lengths that fill header bytes with ones (the bit-stuffing state machine), many
precincts and tile-parts (the layout prefix sum, Psot, TLM), 64-bit blocks.  Every other encoder test in this
directory runs the device path too (it is the default) and compares with the reference itself."""
import os
import numpy as np
import pytest
import openjph_b200 as ob
from openjph_b200.codestream import comp_dims


def _cases():
    rng = np.random.default_rng(77)

    def noise(p, bd, smooth=0):
        out = []
        for (w, h) in comp_dims(p):
            f = rng.integers(0, 1 << bd, (h, w))
            if smooth:
                f >>= smooth
            out.append(f.astype(np.int32))
        return out

    # (name, params, planes)
    p = ob.make_params(517, 389, 1, 16, num_decomps=5, reversible=True, block=(4, 4))
    yield "tiny blocks, odd grids", p, noise(p, 16)
    p = ob.make_params(1000, 40, 3, 16, num_decomps=2, reversible=True, color_transform=True, block=(1024, 4))
    yield "long blocks, long lengths", p, noise(p, 16)
    p = ob.make_params(640, 480, 3, 12, num_decomps=4, reversible=True, color_transform=True, block=(64, 64))
    yield "full-range noise: Lblock runs", p, noise(p, 12)
    p = ob.make_params(300, 300, 3, 8, num_decomps=3, reversible=True, block=(16, 16))
    z = [np.zeros((300, 300), np.int32) for _ in range(3)]
    yield "all zero: empty packets", p, z
    z2 = [a.copy() for a in z]; z2[1][200:, 250:] = 255
    yield "one corner only: empty bands, lone included leaves", p, z2
    p = ob.make_params(333, 222, 3, 10, num_decomps=4, reversible=False, color_transform=True, tile=(100, 90), tlm=True,
                       tilepart_div=3, prog_order="LRCP", precincts=[(64, 64)] * 5, block=(32, 32), offset=(7, 5), tile_offset=(3, 2))
    yield "tiles, tile-part division, TLM, precincts", p, noise(p, 10, smooth=3)
    p = ob.make_params(257, 129, 4, 8, num_decomps=3, reversible=True, prog_order="PCRL", subsampling=[(1, 1), (2, 2), (2, 1), (1, 1)],
                       precincts=[(32, 32)] * 4, block=(8, 32))
    yield "position progression, sub-sampling", p, noise(p, 8)
    p = ob.make_params(90, 70, 3, 32, num_decomps=3, reversible=True, color_transform=True, block=(32, 32))
    yield "64-bit blocks", p, [rng.integers(-(1 << 31), 1 << 31, (70, 90)).astype(np.int32) for _ in range(3)]
    p = ob.make_params(200, 130, 3, 12, num_decomps=5, reversible=False, color_transform=True, decomp="BHVHB", prog_order="PCRL",
                       precincts=[(128, 128)] * 6, block=(256, 8), atk=dict(K=1.0, A=[0.25, -0.5]), qstep=0.0001)
    yield "DFS + ATK", p, noise(p, 12, smooth=2)
    p = ob.make_params(64, 64, 1, 8, num_decomps=0, reversible=True)
    yield "no decomposition", p, noise(p, 8)


def _both(lib, p, planes, bd_type=ob.I32):
    old = os.environ.get("OJB_HOST_HEADERS")
    try:
        os.environ["OJB_HOST_HEADERS"] = "0"
        e = ob.Encoder(p, bd_type, lib=lib); dev = e.encode(planes); ldev = e.kernel_launches
        os.environ["OJB_HOST_HEADERS"] = "1"
        e = ob.Encoder(p, bd_type, lib=lib); host = e.encode(planes); lhost = e.kernel_launches
    finally:
        if old is None:
            os.environ.pop("OJB_HOST_HEADERS", None)
        else:
            os.environ["OJB_HOST_HEADERS"] = old
    return dev, host, ldev, lhost


def _check(lib):
    for name, p, planes in _cases():
        dev, host, ldev, lhost = _both(lib, p, planes)
        assert ldev > lhost, name                       # the two paths really are different kernels
        assert dev == host, name


def test_device_headers_match_host_writer_emulator(emu_lib):
    _check(emu_lib)


@pytest.mark.gpu
def test_device_headers_match_host_writer_gpu(gpu_lib):
    _check(gpu_lib)


def test_output_buffer_too_small_is_reported(emu_lib):
    """nothing is written past the caller's capacity; the call fails with the size it needs"""
    import ctypes as C
    rng = np.random.default_rng(5)
    p = ob.make_params(128, 128, 1, 8, num_decomps=3, reversible=True)
    e = ob.Encoder(p, ob.I32, lib=emu_lib)
    planes = [rng.integers(0, 256, (128, 128)).astype(np.int32)]
    full = e.encode(planes)
    arrs, ptrs = e._planes(planes)
    cap = len(full) - 100
    guard = 4096
    buf = np.full(cap + guard, 0xA5, np.uint8)
    n = C.c_uint64()
    rc = emu_lib.ojb_enc_encode_frame(e.h, ptrs, None, buf.ctypes.data_as(C.c_void_p), C.c_uint64(cap), C.byref(n))
    assert rc != 0 and "000b0030" in emu_lib.ojb_last_error().decode().lower()
    assert (buf[cap:] == 0xA5).all()
    assert e.encode(planes) == full                     # and the encoder is still usable

"""A codestream that never leaves the device: the encoder writes it to device memory, the decoder reads its headers
from there (ojb_dec_read_headers_device) -- only the marker segments and packet headers the host parsers touch are
fetched (32 KB pages; a codestream under 8 MB is simply copied whole), code-block bodies stay in HBM.  Same decoded samples as the host-buffer path."""
import numpy as np
import pytest
import openjph_b200 as ob
import images


def _device_buffer(lib, nbytes):
    """(keep-alive object, address): device memory on the GPU, plain host memory under the emulator"""
    if lib is None:
        import torch
        t = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        return t, t.data_ptr()
    a = np.zeros(nbytes, np.uint8)
    return a, a.ctypes.data


def _check(lib, w, h, nc, bd, max_mirror_frac, **kw):
    frame = images.synth_frame(w, h, nc, bd, 5)
    p = ob.make_params(w, h, nc, bd, **kw)
    enc = ob.Encoder(p, ob.I32, lib=lib)
    want_cs = enc.encode(frame)
    cap = w * h * nc * 4 + (1 << 16)
    keep, addr = _device_buffer(lib, cap)
    enc.upload(frame)
    n = enc.encode_resident_to_device(addr, cap)
    assert n == len(want_cs)
    dec = ob.Decoder(lib=lib)
    fi = dec.read_headers_device(addr, n)
    assert (fi.width, fi.height, fi.num_comps) == (w, h, nc)
    out = dec.decode()
    ref = ob.Decoder(lib=lib).decode(want_cs)
    for a, b in zip(out, ref):
        assert np.array_equal(a, b)
    if kw.get("reversible"):
        for a, b in zip(out, frame):
            assert np.array_equal(a, b)
    assert 0 < dec.mirror_bytes <= n
    assert dec.mirror_bytes <= max_mirror_frac * n, (dec.mirror_bytes, n)
    # a second frame through the same objects (geometry cached, mirror reset)
    frame2 = images.synth_frame(w, h, nc, bd, 6)
    enc.upload(frame2)
    n2 = enc.encode_resident_to_device(addr, cap)
    dec.read_headers_device(addr, n2)
    out2 = dec.decode()
    ref2 = ob.Decoder(lib=lib).decode(enc.encode(frame2))
    for a, b in zip(out2, ref2):
        assert np.array_equal(a, b)
    del keep


def test_device_resident_codestream_emulator(emu_lib):
    _check(emu_lib, 640, 480, 3, 8, 1.0, num_decomps=4, reversible=True, color_transform=True)
    _check(emu_lib, 300, 200, 1, 10, 1.0, num_decomps=3, reversible=False, qstep=0.01, tile=(128, 128), tlm=True)


@pytest.mark.gpu
def test_device_resident_codestream_gpu(gpu_lib):
    _check(None, 640, 480, 3, 8, 1.0, num_decomps=4, reversible=True, color_transform=True)
    _check(None, 4096, 4096, 3, 12, 0.10, num_decomps=5, reversible=True, color_transform=True)
    _check(None, 2048, 2048, 1, 10, 1.0, num_decomps=5, reversible=False, qstep=0.002, tile=(1024, 1024), tlm=True)


def _stream_of_frames(lib, on_gpu):
    """several DIFFERENT frames through the same resident encoder / decoder pair: from the third call on the device
    work is one replayed CUDA graph (CodecBase::run_frame), so every frame must still come out as its own"""
    import ctypes as C
    w, h, nc, bd = 512, 320, 3, 10
    p = ob.make_params(w, h, nc, bd, num_decomps=4, reversible=True, color_transform=True, block=(32, 32))
    enc = ob.Encoder(p, ob.I32, lib=lib)
    dec = ob.Decoder(lib=lib)
    L = enc.L
    cap = w * h * nc * 4 + (1 << 16)
    keep, addr = _device_buffer(None if on_gpu else lib, cap)
    rng = np.random.default_rng(31)
    frames = [images.synth_frame(w, h, nc, bd, 40 + i) for i in range(5)]
    frames.insert(2, [np.zeros((h, w), np.int32) for _ in range(nc)])                  # nothing to code at all
    frames.insert(4, [rng.integers(0, 1 << bd, (h, w)).astype(np.int32) for _ in range(nc)])   # full-range noise
    for k, fr in enumerate(frames):
        want = ob.Encoder(p, ob.I32, lib=lib).encode(fr)
        enc.upload(fr)
        n = enc.encode_resident_to_device(addr, cap)
        assert n == len(want), k
        if on_gpu:
            got = keep[:n].cpu().numpy().tobytes()
        else:
            got = keep[:n].tobytes()
        assert got == want, k
        dec.read_headers_device(addr, n)
        assert L.ojb_dec_decode_resident(dec.h) == 0, L.ojb_last_error()
        for c in range(nc):
            ptr = L.ojb_dec_device_plane(dec.h, c)
            if on_gpu:
                host = np.empty((h, w), np.int32)
                rt = None
                for name in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so", "libcudart.so"):
                    try:
                        rt = C.CDLL(name); break
                    except OSError:
                        pass
                assert rt is not None, "no CUDA runtime to read the device plane back with"
                assert rt.cudaMemcpy(C.c_void_p(host.ctypes.data), C.c_void_p(ptr), C.c_size_t(host.nbytes), 2) == 0
            else:
                host = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), (h, w)).copy()
            assert np.array_equal(host, fr[c]), (k, c)


def test_resident_stream_of_different_frames_emulator(emu_lib):
    _stream_of_frames(emu_lib, False)


@pytest.mark.gpu
def test_resident_stream_of_different_frames_gpu(gpu_lib):
    _stream_of_frames(gpu_lib, True)

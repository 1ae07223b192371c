"""File sample layouts (SURVEY §8 f, N2): .pgm/.ppm and .yuv payloads converted on the GPU, against the
reference's own readers and writers (ppm_in / ppm_out / yuv_in / yuv_out, src/apps/others/ojph_img_io.cpp,
compiled into the oracle harness): same codestream from the same file, same file from the same codestream."""
import numpy as np
import pytest
import openjph_b200 as ob

CASES = {
    # name: (layout, width, height, comps, depth, extra make_params kwargs)
    "ppm8_rct": ("pnm", 203, 131, 3, 8, dict(num_decomps=4, reversible=True, color_transform=True)),
    "ppm12_rct": ("pnm", 160, 97, 3, 12, dict(num_decomps=3, reversible=True, color_transform=True)),
    "ppm16_tiles": ("pnm", 130, 70, 3, 16, dict(num_decomps=2, reversible=True, color_transform=True, tile=(64, 64))),
    "pgm10": ("pnm", 101, 64, 1, 10, dict(num_decomps=3, reversible=True)),
    "pgm8_1px": ("pnm", 1, 5, 1, 8, dict(num_decomps=1, reversible=True)),
    "ppm10_irv": ("pnm", 128, 96, 3, 10, dict(num_decomps=3, reversible=False, color_transform=True, qstep=0.05)),
    "yuv420_8": ("yuv", 128, 96, 3, 8, dict(num_decomps=3, reversible=True, subsampling=[(1, 1), (2, 2), (2, 2)], planar=1)),
    "yuv444_10_irv": ("yuv", 90, 60, 3, 10, dict(num_decomps=3, reversible=False, qstep=0.08, planar=1)),
    "yuv400_12": ("yuv", 77, 50, 1, 12, dict(num_decomps=2, reversible=True)),
}


def _payload(layout, w, h, nc, bd, sub, seed=3):
    rng = np.random.default_rng(seed)
    planes = []
    for c in range(nc):
        dx, dy = sub[c]
        cw, ch = (w + dx - 1) // dx, (h + dy - 1) // dy
        y, x = np.mgrid[0:ch, 0:cw]
        a = (np.sin(x / 11.0 + c) * np.cos(y / 5.0) * 0.5 + 0.5) * ((1 << bd) - 1) + rng.normal(0, (1 << bd) / 40.0, (ch, cw))
        a = np.clip(np.rint(a), 0, (1 << bd) - 1).astype(np.uint16)
        a.flat[0], a.flat[-1] = 0, (1 << bd) - 1
        planes.append(a)
    if layout == "pnm":
        pix = np.stack(planes, axis=-1)                          # interleaved, big-endian 16-bit / 8-bit
        return pix.astype(">u2").tobytes() if bd > 8 else pix.astype(np.uint8).tobytes()
    return b"".join((p.astype("<u2") if bd > 8 else p.astype(np.uint8)).tobytes() for p in planes)


def _check(lib, ref, name, tmp_path):
    layout, w, h, nc, bd, kw = CASES[name]
    sub = kw.get("subsampling", [(1, 1)] * nc)
    p = ob.make_params(w, h, nc, bd, **kw)
    payload = _payload(layout, w, h, nc, bd, sub)
    kind = 0 if layout == "pnm" else 1
    src = tmp_path / (("in.ppm" if nc == 3 else "in.pgm") if kind == 0 else "in.yuv")
    header = (b"P6" if nc == 3 else b"P5") + b"\n%d %d\n%d\n" % (w, h, (1 << bd) - 1) if kind == 0 else b""
    src.write_bytes(header + payload)
    # encode: the reference reads the file line by line, we take the payload as it is
    planes = ref.read_image(str(src), kind, w, h, nc, bd, sub)
    want = ref.encode(p, planes)
    container = ob.U16 if bd > 8 else ob.U8
    got = ob.Encoder(p, container, lib=lib).encode_raster(payload, layout)
    if p.reversible:
        assert got == want
    else:
        assert len(got) == len(want)
    # decode: the reference pulls lines and writes them through its file writer
    refplanes, _ = ref.decode(want)
    dst = tmp_path / (("out.ppm" if nc == 3 else "out.pgm") if kind == 0 else "out.yuv")
    ref.write_image(str(dst), kind, bd, refplanes)
    data = dst.read_bytes()
    if kind == 0:
        assert data.startswith(header)
        data = data[len(header):]
    ours = ob.Decoder(lib=lib).decode_raster(want, layout)
    assert len(ours) == len(data)
    if p.reversible:
        assert ours == data and ours == payload
    else:
        dt = (">u2" if kind == 0 else "<u2") if bd > 8 else np.uint8
        a, b = np.frombuffer(ours, dt).astype(np.int64), np.frombuffer(data, dt).astype(np.int64)
        assert np.abs(a - b).max() <= 1 and a.max() <= (1 << bd) - 1


@pytest.mark.parametrize("name", list(CASES))
def test_raster_layout_emulator(name, emu_lib, ref, tmp_path):
    _check(emu_lib, ref, name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_raster_layout_gpu(name, gpu_lib, ref, tmp_path):
    _check(None, ref, name, tmp_path)


def test_clamp_to_component_range(emu_lib, ref, tmp_path):
    """dropping the top resolutions of a 5/3 stream (skip_read > skip_recon) leaves over- and undershoot: the
    planes must come out clamped to [0, 1023] as the reference's writers clamp (gen_cvrt_32b*: 0 <= val <=
    max_val), not to the container's 65535"""
    w, h, bd = 64, 64, 10
    p = ob.make_params(w, h, 1, bd, num_decomps=3, reversible=True)
    rng = np.random.default_rng(5)
    img = (rng.integers(0, 2, (h, w)) * 1023).astype(np.int32)
    cs = ref.encode(p, [img])
    refplanes, _ = ref.decode_restricted(cs, 2, 0)
    assert refplanes[0].max() > 1023 and refplanes[0].min() < 0         # the overshoot is there
    dst = tmp_path / "o.pgm"
    ref.write_image(str(dst), 0, bd, refplanes)
    data = dst.read_bytes().split(b"\n", 3)[3]
    ours = ob.Decoder(lib=emu_lib).decode_raster(cs, "pnm", skip=(2, 0))
    assert ours == data and np.frombuffer(ours, ">u2").max() == 1023


def test_raster_errors(emu_lib):
    p = ob.make_params(32, 32, 3, 12, reversible=True)
    enc = ob.Encoder(p, ob.U16, lib=emu_lib)
    with pytest.raises(ob.OjphError):
        enc.encode_raster(b"\0" * 100, "pnm")                           # short payload
    with pytest.raises(ob.OjphError):
        ob.Encoder(p, ob.I32, lib=emu_lib).encode_raster(b"\0" * (32 * 32 * 6), "pnm")
    p2 = ob.make_params(32, 32, 3, 8, reversible=True, subsampling=[(1, 1), (2, 2), (2, 2)], planar=1)
    with pytest.raises(ob.OjphError):
        ob.Encoder(p2, ob.U8, lib=emu_lib).encode_raster(b"\0" * (32 * 32 * 3), "pnm")   # sub-sampled ppm


@pytest.mark.gpu
def test_raster_full_size_matches_planar_path_gpu(gpu_lib):
    """BASELINE size (8192 x 8192 x 3, 12-bit): the .ppm payload path must give the codestream of the planar
    path on the same samples, and decode back to the same payload (size-independent property, no oracle)"""
    w = h = 8192
    rng = np.random.default_rng(9)
    base = (rng.integers(0, 4096, (h // 8, w // 8, 3), dtype=np.uint16)).repeat(8, axis=0).repeat(8, axis=1)
    base[::7, ::5, :] ^= 0x155                                   # some texture
    p = ob.make_params(w, h, 3, 12, num_decomps=5, reversible=True, color_transform=True)
    payload = base.astype(">u2").tobytes()
    enc = ob.Encoder(p, ob.U16)
    a = enc.encode_raster(payload, "pnm")
    b = enc.encode([np.ascontiguousarray(base[:, :, c]) for c in range(3)])
    assert a == b
    back = ob.Decoder().decode_raster(a, "pnm")
    assert back == payload


def _dpx_file(path, w, h, bd, big_endian, payload):
    """a minimal .dpx file around an image-data payload (the fields dpx_in::open reads, :1823-2030)"""
    import struct
    e = ">" if big_endian else "<"
    hdr = bytearray(2048)
    hdr[0:4] = struct.pack(e + "I", 0x53445058)
    hdr[4:8] = struct.pack(e + "I", 2048)
    hdr[8:16] = b"V2.0\0\0\0\0"
    hdr[16:20] = struct.pack(e + "I", 2048 + len(payload))
    hdr[768:772] = struct.pack(e + "HH", 0, 1)
    hdr[772:780] = struct.pack(e + "II", w, h)
    hdr[780:784] = struct.pack(e + "I", 0)
    hdr[800:804] = bytes([50, 2, 2, bd])
    hdr[804:808] = struct.pack(e + "HH", 1 if bd == 10 else 0, 0)
    hdr[808:812] = struct.pack(e + "I", 2048)
    path.write_bytes(bytes(hdr) + payload)


DPX_CASES = [(10, True, 131, 70), (10, False, 64, 33), (16, True, 77, 40), (16, False, 50, 31)]


@pytest.mark.parametrize("bd,big_endian,w,h", DPX_CASES)
def test_dpx_payload_emulator(bd, big_endian, w, h, emu_lib, ref, tmp_path):
    _check_dpx(emu_lib, ref, tmp_path, bd, big_endian, w, h)


@pytest.mark.gpu
@pytest.mark.parametrize("bd,big_endian,w,h", DPX_CASES)
def test_dpx_payload_gpu(bd, big_endian, w, h, gpu_lib, ref, tmp_path):
    _check_dpx(None, ref, tmp_path, bd, big_endian, w, h)


def _check_dpx(emu_lib, ref, tmp_path, bd, big_endian, w, h):
    rng = np.random.default_rng(bd + w)
    pix = rng.integers(0, 1 << bd, (h, w, 3), dtype=np.uint32)
    if bd == 10:
        words = (pix[:, :, 0] << 22) | (pix[:, :, 1] << 12) | (pix[:, :, 2] << 2)
        payload = words.astype(">u4" if big_endian else "<u4").tobytes()
    else:
        row = np.zeros((h, (3 * w + 1) & ~1), np.uint16)
        row[:, :3 * w] = pix.reshape(h, 3 * w)
        payload = row.astype(">u2" if big_endian else "<u2").tobytes()
    f = tmp_path / "in.dpx"
    _dpx_file(f, w, h, bd, big_endian, payload)
    planes = ref.read_image(str(f), 2, w, h, 3, bd)
    for c in range(3):
        assert np.array_equal(planes[c], pix[:, :, c])            # the test's idea of the format is the reference's
    p = ob.make_params(w, h, 3, bd, num_decomps=3, reversible=True, color_transform=True)
    want = ref.encode(p, planes)
    got = ob.Encoder(p, ob.U16, lib=emu_lib).encode_raster(payload, "dpx_be" if big_endian else "dpx_le")
    assert got == want
    with pytest.raises(ob.OjphError):
        ob.Decoder(lib=emu_lib).decode_raster(got, "dpx_be")      # read-only format

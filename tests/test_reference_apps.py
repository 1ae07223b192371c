"""Drop-in proof at the application level: the reference's own ojph_compress.cpp and ojph_expand.cpp, UNMODIFIED,
are compiled twice -- against ojph::codestream and, through a force-included header, against
ojph::b200::codestream -- and run on the same files: same codestream bytes, same decoded image.  Every test runs
in the CPU tier over the SIMT-emulator build (needs /root/reference) and, marked `gpu`, over the nvcc-built
product library on a real B200 (executables prebuilt by oracle/Makefile into oracle/_ref/apps)."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
REFLIB = os.path.join(ROOT, "oracle", "_ref")


PREBUILT = os.path.join(REFLIB, "apps")


@pytest.fixture(scope="module", params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def apps(request, tmp_path_factory):
    """(app, flavour) -> executable.  `emu`: compiled here against the SIMT-emulator build (CPU tier, needs
    /root/reference).  `gpu`: the executables oracle/Makefile built against the nvcc-built product library
    (oracle/_ref/apps, travels with gpurun) -- the same unmodified sources running on a real B200."""
    if request.param == "gpu":
        request.getfixturevalue("gpu_lib")
        out = {(a, f): os.path.join(PREBUILT, "ojph_%s_%s" % (a, f)) for a in ("compress", "expand") for f in ("ref", "b200")}
        missing = [v for v in out.values() if not os.path.exists(v)]
        if missing and not os.path.exists(os.path.join(REFLIB, "libopenjph_ref.so")):
            pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
        assert not missing, "oracle/_ref/apps is incomplete: %s" % missing
        return out
    request.getfixturevalue("emu_lib")
    if not os.path.isdir(REF) or not os.path.exists(os.path.join(REFLIB, "libopenjph_ref.so")):
        pytest.skip("needs the reference sources and oracle/_ref")
    d = tmp_path_factory.mktemp("apps")
    inc = ["-I" + os.path.join(REF, "core", "openjph"), "-I" + os.path.join(REF, "core", "common"),
           "-I" + os.path.join(REF, "apps", "common"), "-I" + os.path.join(ROOT, "include")]
    io = [os.path.join(REFLIB, "obj", "apps", n) for n in ("ojph_img_io.o", "ojph_img_io_sse41.o", "ojph_img_io_avx2.o")]
    emu_dir = os.path.join(ROOT, "tests", "emu")
    out = {}
    for app in ("compress", "expand"):
        src = os.path.join(REF, "apps", "ojph_" + app, "ojph_" + app + ".cpp")
        for flavour in ("ref", "b200"):
            exe = str(d / ("ojph_%s_%s" % (app, flavour)))
            cmd = ["g++", "-std=c++14", "-O1"] + inc
            if flavour == "b200":
                cmd += ["-include", os.path.join(ROOT, "tests", "cpp", "facade_force_include.h")]
            cmd += [src] + io + ["-o", exe, "-L" + REFLIB, "-lopenjph_ref", "-Wl,-rpath," + REFLIB]
            if flavour == "b200":
                cmd += ["-L" + emu_dir, "-lojph_b200_emu", "-Wl,-rpath," + emu_dir]
            subprocess.check_call(cmd)
            out[(app, flavour)] = exe
    return out


def _ppm(path, w, h, depth, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    pix = np.stack([(np.sin(x / 9.0 + c) * np.cos(y / 6.0) * 0.45 + 0.5) * ((1 << depth) - 1) + rng.normal(0, 2, (h, w))
                    for c in range(3)], axis=-1)
    pix = np.clip(np.rint(pix), 0, (1 << depth) - 1)
    body = pix.astype(">u2").tobytes() if depth > 8 else pix.astype(np.uint8).tobytes()
    path.write_bytes(b"P6\n%d %d\n%d\n" % (w, h, (1 << depth) - 1) + body)


CASES = {
    "rev_rct": (8, ["-reversible", "true", "-num_decomps", "4"]),
    "rev_12bit_tiles_tlm": (12, ["-reversible", "true", "-tile_size", "{96,64}", "-tlm_marker", "true", "-prog_order", "CPRL",
                                 "-tileparts", "C", "-block_size", "{32,32}", "-precincts", "{128,128},{64,64}"]),
    "irv_q": (10, ["-qstep", "0.01", "-num_decomps", "3", "-com", "made by the facade test"]),
    "irv_qfactor": (8, ["-qfactor", "85", "-num_decomps", "4"]),
    "rev_offsets_no_rct": (8, ["-reversible", "true", "-colour_trans", "false", "-image_offset", "{5,3}", "-tile_size", "{128,128}",
                               "-tile_offset", "{1,2}", "-prog_order", "PCRL"]),
    "broadcast_profile": (10, ["-reversible", "true", "-profile", "BROADCAST", "-prog_order", "CPRL", "-precincts", "{128,128},{256,256}"]),
}


def test_yuv420_through_the_apps(apps, tmp_path):
    """a sub-sampled .yuv in and out (-dims / -num_comps / -downsamp / -bit_depth / -signed)"""
    w, h = 128, 96
    rng = np.random.default_rng(8)
    src = tmp_path / "in.yuv"
    src.write_bytes(b"".join(rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (w * h, w * h // 4, w * h // 4)))
    opts = ["-dims", "{%d,%d}" % (w, h), "-num_comps", "3", "-downsamp", "{1,1},{2,2}", "-bit_depth", "8", "-signed", "false",
            "-reversible", "true", "-num_decomps", "3"]
    res = {}
    for fl in ("ref", "b200"):
        j = tmp_path / ("o_%s.j2c" % fl)
        subprocess.check_call([apps[("compress", fl)], "-i", str(src), "-o", str(j)] + opts, stdout=subprocess.DEVNULL)
        o = tmp_path / ("b_%s.yuv" % fl)
        subprocess.check_call([apps[("expand", fl)], "-i", str(j), "-o", str(o)], stdout=subprocess.DEVNULL)
        res[fl] = (j.read_bytes(), o.read_bytes())
    assert res["ref"] == res["b200"] and res["ref"][1] == src.read_bytes()


@pytest.mark.parametrize("name", list(CASES))
def test_unmodified_reference_apps_on_the_facade(name, apps, tmp_path):
    depth, opts = CASES[name]
    src = tmp_path / "in.ppm"
    _ppm(src, 203, 131, depth, 4)
    j = {}
    for fl in ("ref", "b200"):
        j[fl] = tmp_path / ("out_%s.j2c" % fl)
        subprocess.check_call([apps[("compress", fl)], "-i", str(src), "-o", str(j[fl])] + opts, stdout=subprocess.DEVNULL)
    a, b = j["ref"].read_bytes(), j["b200"].read_bytes()
    if "-reversible" in opts:
        assert a == b
    else:
        assert len(a) == len(b) and a[:a.index(b"\xff\x90")] == b[:b.index(b"\xff\x90")]
    outs = {}
    for fl in ("ref", "b200"):
        o = tmp_path / ("back_%s.ppm" % fl)
        subprocess.check_call([apps[("expand", fl)], "-i", str(j["ref"]), "-o", str(o)], stdout=subprocess.DEVNULL)
        outs[fl] = o.read_bytes()
    if "-reversible" in opts:
        assert outs["ref"] == outs["b200"] == src.read_bytes()
    else:
        hdr = len(outs["ref"]) - 203 * 131 * 3 * 2
        x, y = (np.frombuffer(outs[f][hdr:], ">u2").astype(np.int64) for f in ("ref", "b200"))
        assert outs["ref"][:hdr] == outs["b200"][:hdr] and np.abs(x - y).max() <= 1
    # reduced resolution through the application's own option
    for fl in ("ref", "b200"):
        o = tmp_path / ("small_%s.ppm" % fl)
        subprocess.check_call([apps[("expand", fl)], "-i", str(j["ref"]), "-o", str(o), "-skip_res", "1,1"], stdout=subprocess.DEVNULL)
        outs[fl] = o.read_bytes()
    if "-reversible" in opts:
        assert outs["ref"] == outs["b200"]


# the encode half of the reference's own test matrix (tests/test_executables.cpp:1033-1690: SimpleEncIrv97*,
# SimpleEncRev53*, tiles, 16-bit, gray, tall-narrow, qfactor), on synthetic images of the same kind
MATRIX = [
    ("irv97_64x64", "ppm8", ["-qstep", "0.1"]),
    ("irv97_32x32", "ppm8", ["-qstep", "0.01", "-block_size", "{32,32}"]),
    ("irv97_16x16", "ppm8", ["-qstep", "0.01", "-block_size", "{16,16}"]),
    ("irv97_4x4", "ppm8", ["-qstep", "0.01", "-block_size", "{4,4}"]),
    ("irv97_1024x4", "ppm8", ["-qstep", "0.01", "-block_size", "{4,1024}"]),
    ("irv97_4x1024", "ppm8", ["-qstep", "0.01", "-block_size", "{1024,4}"]),
    ("irv97_512x8", "ppm8", ["-qstep", "0.01", "-block_size", "{8,512}"]),
    ("irv97_8x512", "ppm8", ["-qstep", "0.01", "-block_size", "{512,8}"]),
    ("irv97_256x16", "ppm8", ["-qstep", "0.01", "-block_size", "{16,256}"]),
    ("irv97_16x256", "ppm8", ["-qstep", "0.01", "-block_size", "{256,16}"]),
    ("irv97_128x32", "ppm8", ["-qstep", "0.01", "-block_size", "{32,128}"]),
    ("irv97_32x128", "ppm8", ["-qstep", "0.01", "-block_size", "{128,32}"]),
    ("irv97_tiles_33x33_d5", "ppm8", ["-qstep", "0.01", "-tile_size", "{33,33}", "-num_decomps", "5"]),
    ("irv97_tiles_33x33_d6", "ppm8", ["-qstep", "0.01", "-tile_size", "{33,33}", "-num_decomps", "6"]),
    ("irv97_16bit", "ppm16", ["-qstep", "0.01"]),
    ("irv97_16bit_gray", "pgm16", ["-qstep", "0.01"]),
    ("rev53_16bit", "ppm16", ["-reversible", "true"]),
    ("rev53_16bit_gray", "pgm16", ["-reversible", "true"]),
    ("rev53_64x64", "ppm8", ["-reversible", "true"]),
    ("rev53_32x32", "ppm8", ["-reversible", "true", "-block_size", "{32,32}"]),
    ("rev53_4x4", "ppm8", ["-reversible", "true", "-block_size", "{4,4}"]),
    ("rev53_1024x4", "ppm8", ["-reversible", "true", "-block_size", "{4,1024}"]),
    ("rev53_4x1024", "ppm8", ["-reversible", "true", "-block_size", "{1024,4}"]),
    ("rev53_tiles_32x32_d5", "ppm8", ["-reversible", "true", "-tile_size", "{32,32}", "-num_decomps", "5"]),
    ("rev53_tiles_32x32_d6", "ppm8", ["-reversible", "true", "-tile_size", "{32,32}", "-num_decomps", "6"]),
    ("irv97_tall_narrow", "tall", ["-qstep", "0.1"]),
    ("irv97_tall_narrow1", "tall", ["-image_offset", "{1,0}", "-qstep", "0.1"]),
    ("rev53_tall_narrow", "tall", ["-reversible", "true"]),
    ("rev53_tall_narrow1", "tall", ["-image_offset", "{1,0}", "-reversible", "true"]),
    ("irv97_qfactor50", "ppm8", ["-qfactor", "50"]),
    ("irv97_qfactor50_gray", "pgm8", ["-qfactor", "50"]),
]


def _image(path_dir, kind):
    w, h, nc, depth = {"ppm8": (300, 200, 3, 8), "ppm16": (200, 150, 3, 16), "pgm16": (200, 150, 1, 16),
                       "pgm8": (300, 200, 1, 8), "tall": (7, 400, 3, 8)}[kind]
    rng = np.random.default_rng(len(kind) + w)
    y, x = np.mgrid[0:h, 0:w]
    pix = np.stack([(np.sin(x / 9.0 + c) * np.cos(y / 6.0) * 0.45 + 0.5) * ((1 << depth) - 1) + rng.normal(0, (1 << depth) / 100.0, (h, w))
                    for c in range(nc)], axis=-1)
    pix = np.clip(np.rint(pix), 0, (1 << depth) - 1)
    body = pix.astype(">u2").tobytes() if depth > 8 else pix.astype(np.uint8).tobytes()
    p = path_dir / ("in." + ("ppm" if nc == 3 else "pgm"))
    p.write_bytes((b"P6" if nc == 3 else b"P5") + b"\n%d %d\n%d\n" % (w, h, (1 << depth) - 1) + body)
    return p, w * h * nc, depth


@pytest.mark.parametrize("name,kind,opts", MATRIX, ids=[m[0] for m in MATRIX])
def test_reference_encode_matrix_on_the_facade(name, kind, opts, apps, tmp_path):
    src, nsamp, depth = _image(tmp_path, kind)
    j, back = {}, {}
    for fl in ("ref", "b200"):
        j[fl] = tmp_path / ("o_%s.j2c" % fl)
        subprocess.check_call([apps[("compress", fl)], "-i", str(src), "-o", str(j[fl])] + opts, stdout=subprocess.DEVNULL)
        o = tmp_path / ("b_%s%s" % (fl, src.suffix))
        subprocess.check_call([apps[("expand", fl)], "-i", str(j[fl]), "-o", str(o)], stdout=subprocess.DEVNULL)
        back[fl] = o.read_bytes()
    a, b = j["ref"].read_bytes(), j["b200"].read_bytes()
    if "-reversible" in opts:
        assert a == b and back["ref"] == back["b200"] == src.read_bytes()
    else:
        assert abs(len(a) - len(b)) <= max(4, len(a) // 1000) and a[:a.index(b"\xff\x90")] == b[:b.index(b"\xff\x90")]
        dt = ">u2" if depth > 8 else np.uint8
        es = 2 if depth > 8 else 1
        orig = np.frombuffer(src.read_bytes()[-nsamp * es:], dt).astype(np.float64)
        x, y = (np.frombuffer(back[f][-nsamp * es:], dt).astype(np.float64) for f in ("ref", "b200"))
        m_ref, m_new = ((x - orig) ** 2).mean(), ((y - orig) ** 2).mean()
        # the reference tests' own tolerance: MSE within 1 %, peak error within 1 (tests/test_executables.cpp:132-133)
        assert abs(m_new - m_ref) <= max(0.01 * m_ref, 0.01) and abs(np.abs(y - orig).max() - np.abs(x - orig).max()) <= max(1, 0.01 * np.abs(x - orig).max())


@pytest.mark.parametrize("bd,big_endian", [(10, True), (10, False), (16, True), (16, False)])
def test_dpx_input_through_the_apps(bd, big_endian, apps, tmp_path):
    """the reference's dpx_enc_* tests (tests/test_executables.cpp:1543-1630): a .dpx in, reversible"""
    from test_raster_layouts import _dpx_file
    w, h = 160, 90
    rng = np.random.default_rng(bd)
    pix = rng.integers(0, 1 << bd, (h, w, 3), dtype=np.uint32)
    if bd == 10:
        payload = ((pix[:, :, 0] << 22) | (pix[:, :, 1] << 12) | (pix[:, :, 2] << 2)).astype(">u4" if big_endian else "<u4").tobytes()
    else:
        payload = pix.reshape(h, 3 * w).astype(">u2" if big_endian else "<u2").tobytes()
    src = tmp_path / "in.dpx"
    _dpx_file(src, w, h, bd, big_endian, payload)
    out = {}
    for fl in ("ref", "b200"):
        j = tmp_path / ("o_%s.j2c" % fl)
        subprocess.check_call([apps[("compress", fl)], "-i", str(src), "-o", str(j), "-reversible", "true"], stdout=subprocess.DEVNULL)
        out[fl] = j.read_bytes()
    assert out["ref"] == out["b200"]


def test_truncated_stream_through_expand(apps, tmp_path):
    """ojph_expand -resilient true on a codestream cut short: both builds must agree on what comes out (and a
    non-resilient run must fail in both)"""
    src, _, _ = _image(tmp_path, "ppm8")
    j = tmp_path / "full.j2c"
    subprocess.check_call([apps[("compress", "ref")], "-i", str(src), "-o", str(j), "-reversible", "true", "-tile_size", "{128,128}"],
                          stdout=subprocess.DEVNULL)
    data = j.read_bytes()
    for frac in (0.55, 0.8):
        cut = tmp_path / ("cut_%d.j2c" % int(frac * 100))
        cut.write_bytes(data[:int(len(data) * frac)])
        res = {}
        for fl in ("ref", "b200"):
            o = tmp_path / ("r_%s_%d.ppm" % (fl, int(frac * 100)))
            rc = subprocess.call([apps[("expand", fl)], "-i", str(cut), "-o", str(o), "-resilient", "true"],
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            res[fl] = (rc, o.read_bytes() if o.exists() and rc == 0 else None)
            strict = subprocess.call([apps[("expand", fl)], "-i", str(cut), "-o", str(tmp_path / "x.ppm")],
                                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            assert strict != 0
        assert res["ref"] == res["b200"], frac


@pytest.mark.parametrize("bd,signed", [(8, True), (8, False), (16, True), (16, False), (24, True), (24, False), (32, True), (32, False)])
def test_raw_round_trip_through_the_apps(bd, signed, apps, tmp_path):
    """the reference's SimpleEncRev53Raw{8,16,24,32}{Signed,Unsigned} (tests/test_executables.cpp:1698-1728): a .raw
    file in, reversible, .raw out, bit exact; the 32-bit-sample ones run on the 64-bit coefficient path"""
    w = h = 96
    rng = np.random.default_rng(bd + signed)
    lo, hi = (-(1 << (bd - 1)), (1 << (bd - 1))) if signed else (0, 1 << bd)
    v = rng.integers(lo, hi, w * h, dtype=np.int64)
    v[:4] = (lo, hi - 1, 0, -1 if signed else 1)
    nbytes = (bd + 7) // 8
    raw = b"".join(int(x).to_bytes(nbytes, "little", signed=signed) for x in v)
    src = tmp_path / "in.raw"
    src.write_bytes(raw)
    opts = ["-reversible", "true", "-dims", "{%d,%d}" % (w, h), "-num_comps", "1", "-downsamp", "{1,1}", "-bit_depth", str(bd),
            "-signed", "true" if signed else "false"]
    res = {}
    for fl in ("ref", "b200"):
        j = tmp_path / ("o_%s.j2c" % fl)
        subprocess.check_call([apps[("compress", fl)], "-i", str(src), "-o", str(j)] + opts, stdout=subprocess.DEVNULL)
        o = tmp_path / ("b_%s.raw" % fl)
        subprocess.check_call([apps[("expand", fl)], "-i", str(j), "-o", str(o)], stdout=subprocess.DEVNULL)
        res[fl] = (j.read_bytes(), o.read_bytes())
    assert res["ref"] == res["b200"] and res["ref"][1] == raw


@pytest.mark.parametrize("nc", [1, 3])
def test_pfm_float_through_the_apps(nc, apps, tmp_path):
    """a .pfm (32-bit float) file in and out (ojph_compress.cpp:602-690, ojph_expand.cpp pfm_out): the applications treat
    the floats as 32-bit signed integers with the type-3 non-linearity, reversible -- the 64-bit coefficient path"""
    w, h = 96, 64
    rng = np.random.default_rng(40 + nc)
    y, x = np.mgrid[0:h, 0:w]
    pix = np.stack([(np.sin(x / 7.0 + c) * np.cos(y / 5.0) * (3.0 + c) + rng.normal(0, 0.01, (h, w))).astype("<f4") for c in range(nc)], axis=-1)
    src = tmp_path / "in.pfm"
    src.write_bytes((b"PF" if nc == 3 else b"Pf") + b"\n%d %d\n-1.0\n" % (w, h) + pix.tobytes())
    res = {}
    for fl in ("ref", "b200"):
        j = tmp_path / ("o_%s.j2c" % fl)
        subprocess.check_call([apps[("compress", fl)], "-i", str(src), "-o", str(j), "-reversible", "true"], stdout=subprocess.DEVNULL)
        o = tmp_path / ("b_%s.pfm" % fl)
        subprocess.check_call([apps[("expand", fl)], "-i", str(j), "-o", str(o)], stdout=subprocess.DEVNULL)
        res[fl] = (j.read_bytes(), o.read_bytes())
    assert res["ref"][0] == res["b200"][0]
    assert res["ref"][1] == res["b200"][1]
    assert res["ref"][1][-pix.nbytes:] == pix.tobytes()

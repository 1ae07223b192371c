"""Drop-in proof at the application level (CPU tier, needs /root/reference): the reference's own
ojph_compress.cpp and ojph_expand.cpp, UNMODIFIED, are compiled twice -- against ojph::codestream and, through a
force-included header, against ojph::b200::codestream -- and run on the same files: same codestream bytes, same
decoded image."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
REFLIB = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def apps(tmp_path_factory, emu_lib):
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

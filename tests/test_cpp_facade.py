"""CPU tier, this container only: the header-only C++ facade (include/ojph_b200_codestream.hpp) compiled
against the reference's public headers; one application source instantiated with ojph::codestream and
with ojph::b200::codestream gives byte-identical codestreams (kernels under the SIMT emulator)."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_facade_matches_reference_application_code(emu_lib, ref, tmp_path):
    if not os.path.isdir(REF):
        pytest.skip("the reference's headers are not available here")
    exe = str(tmp_path / "facade_roundtrip")
    emu_dir = os.path.join(ROOT, "tests", "emu")
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    cmd = ["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "facade_roundtrip.cpp"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(REF, "src/core/openjph"),
           "-L", emu_dir, "-lojph_b200_emu", "-L", ref_dir, "-lopenjph_ref",
           "-Wl,-rpath," + emu_dir, "-Wl,-rpath," + ref_dir, "-lpthread", "-o", exe]
    subprocess.check_call(cmd)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("identical=1 lossless=1") == 3 and "reduced=1" in r.stdout, r.stdout


@pytest.mark.gpu
def test_facade_matches_reference_application_code_gpu(gpu_lib):
    """the same program linked against the nvcc-built product library (prebuilt by oracle/Makefile: the GPU box
    has no reference headers), kernels on the real device"""
    exe = os.path.join(ROOT, "oracle", "_ref", "apps", "facade_roundtrip_b200")
    if not os.path.exists(exe) and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libopenjph_ref.so")):
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("identical=1 lossless=1") == 3 and "reduced=1" in r.stdout, r.stdout

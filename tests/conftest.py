import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_lib():
    """the product library on a real GPU; fails loudly if it is missing"""
    from openjph_b200 import _lib
    L = _lib.lib()
    assert L.ojb_device_count() > 0, "no CUDA device visible"
    return L


@pytest.fixture(scope="session")
def emu_lib():
    """SIMT-emulator build of the same sources (CPU test tier only)"""
    import emu
    return emu.emu_lib()


@pytest.fixture(scope="session")
def ref():
    import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return refharness

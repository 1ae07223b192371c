"""Test-only: the SIMT-emulator build of the product sources (tests/emu/libojph_b200_emu.so)."""
import os
import subprocess
from openjph_b200 import _lib

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_PATH = os.path.join(_ROOT, "tests", "emu", "libojph_b200_emu.so")
_emu = None


def emu_lib(build=True):
    global _emu
    if _emu is None:
        if build:
            import fcntl
            # one builder at a time (pytest-xdist workers all arrive here)
            with open(os.path.join(_ROOT, "tests", "emu", ".build.lock"), "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "openjph_b200", "csrc"), "emu", "-j8"])
        _emu = _lib.bind(EMU_PATH)
    return _emu

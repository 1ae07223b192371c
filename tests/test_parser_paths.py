"""The two packet-header readers of the decoder (ojb_layout.cpp): the byte-wise loop that follows
precinct::parse (src/core/codestream/ojph_precinct.cpp:328-573) bit by bit, and the 64-bit-window reader that takes
its place whenever nothing unusual happens.  Same block records from both, on streams that stress what differs:
sparse inclusion (tag-tree nodes that say "nothing below"), deep trees (small code-blocks), precincts, several
tile-parts, 0xFF bytes in the headers (random payload lengths)."""
import os
import ctypes as C
import numpy as np
import pytest
import cases
import openjph_b200 as ob
from openjph_b200 import _lib


def _blocks(L, cs, slow):
    buf = np.frombuffer(cs, np.uint8)
    d = L.ojb_dec_create()
    try:
        fi = _lib.FrameInfo()
        assert L.ojb_dec_read_headers(d, buf.ctypes.data, buf.size, ob.I32, C.byref(fi)) == 0, L.ojb_last_error()
        n = C.c_uint32()
        if slow:
            os.environ["OJB_PARSE_SLOW"] = "1"
        try:
            assert L.ojb_dec_list_blocks(d, None, 0, C.byref(n)) == 0, L.ojb_last_error()
            out = (_lib.BlockDesc * n.value)()
            assert L.ojb_dec_list_blocks(d, out, n.value, C.byref(n)) == 0, L.ojb_last_error()
        finally:
            os.environ.pop("OJB_PARSE_SLOW", None)
        return [(b.missing_msbs, b.num_passes, b.len1, b.len2, b.byte_off) for b in out]
    finally:
        L.ojb_dec_destroy(d)


def _frames(p, kind, seed):
    fr = cases.frame_for(p, "noise" if kind == "noise" else "synth", seed)
    if kind == "holes":          # large flat areas: whole sub-trees of code-blocks with nothing to code
        rng = np.random.default_rng(seed)
        for a in fr:
            h, w = a.shape
            for _ in range(6):
                y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
                a[y0:y0 + h // 2, x0:x0 + w // 2] = a.flat[0]
    if kind == "flat":
        fr = [np.full_like(a, a.flat[0]) for a in fr]
    return fr


PARSER_CASES = [
    (dict(width=640, height=480, num_comps=3, bit_depth=8, num_decomps=5, reversible=True, color_transform=True), "synth"),
    (dict(width=640, height=480, num_comps=3, bit_depth=8, num_decomps=5, reversible=True, color_transform=True), "holes"),
    (dict(width=512, height=512, num_comps=1, bit_depth=8, num_decomps=4, reversible=True, block=(8, 8)), "holes"),
    (dict(width=700, height=300, num_comps=1, bit_depth=12, num_decomps=3, reversible=True, block=(4, 4)), "noise"),
    (dict(width=500, height=333, num_comps=3, bit_depth=10, num_decomps=5, reversible=False, color_transform=True, qstep=0.05), "synth"),
    (dict(width=500, height=333, num_comps=1, bit_depth=8, num_decomps=3, reversible=True, block=(16, 16),
          precincts=[(64, 64), (128, 128)], prog_order="PCRL", tile=(256, 256), tilepart_div=1), "holes"),
    (dict(width=256, height=256, num_comps=1, bit_depth=8, num_decomps=2, reversible=True), "flat"),
    (dict(width=384, height=384, num_comps=1, bit_depth=16, num_decomps=5, reversible=True, block=(32, 32)), "noise"),
]


@pytest.mark.parametrize("ci", range(len(PARSER_CASES)))
def test_window_reader_equals_bytewise_loop(ci, emu_lib, ref):
    kw, kind = PARSER_CASES[ci]
    p = cases.make(kw)
    for seed in (1, 2):
        cs = ref.encode(p, _frames(p, kind, seed))
        fast = _blocks(emu_lib, cs, False)
        slow = _blocks(emu_lib, cs, True)
        assert len(fast) == len(slow) and fast == slow
        if kind != "flat":
            assert any(b[1] for b in fast)
        if kind in ("holes", "flat"):
            assert any(b[1] == 0 for b in fast)          # blocks that are not included


def _blocks_or_error(L, cs, slow, resilient):
    buf = np.frombuffer(cs, np.uint8)
    d = L.ojb_dec_create()
    try:
        if resilient:
            L.ojb_dec_enable_resilience(d)
        fi = _lib.FrameInfo()
        if L.ojb_dec_read_headers(d, buf.ctypes.data, buf.size, ob.I32, C.byref(fi)) != 0:
            return ("main header", L.ojb_last_error())
        n = C.c_uint32()
        if slow:
            os.environ["OJB_PARSE_SLOW"] = "1"
        try:
            if L.ojb_dec_list_blocks(d, None, 0, C.byref(n)) != 0:
                return ("error", L.ojb_last_error())
            out = (_lib.BlockDesc * n.value)()
            if L.ojb_dec_list_blocks(d, out, n.value, C.byref(n)) != 0:
                return ("error", L.ojb_last_error())
        finally:
            os.environ.pop("OJB_PARSE_SLOW", None)
        return [(b.missing_msbs, b.num_passes, b.len1, b.len2, b.byte_off) for b in out]
    finally:
        L.ojb_dec_destroy(d)


def test_damaged_streams_read_alike(emu_lib, ref):
    """corrupted and truncated packet headers: the window reader gives up where anything is out of the ordinary and the
    byte-wise loop reports -- so both paths end with the same block records or the same error (code and text), with and
    without resilience"""
    rng = np.random.default_rng(5)
    trials = 0
    for ci in (0, 2, 5):
        kw, kind = PARSER_CASES[ci]
        p = cases.make(kw)
        cs0 = ref.encode(p, _frames(p, kind, 1))
        sod = cs0.index(b"\xff\x93") + 2
        for _ in range(40):
            b = bytearray(cs0)
            for _k in range(int(rng.integers(1, 4))):
                b[sod + int(rng.integers(0, min(len(b) - sod, 4000)))] = int(rng.integers(0, 256))
            if rng.random() < 0.3:
                b = b[:sod + int(rng.integers(1, len(b) - sod))]
            cs = bytes(b)
            for resilient in (False, True):
                assert _blocks_or_error(emu_lib, cs, False, resilient) == _blocks_or_error(emu_lib, cs, True, resilient)
                trials += 1
    assert trials == 240

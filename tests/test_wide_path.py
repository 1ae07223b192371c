"""The 64-bit coefficient path: parameter sets whose precision exceeds 32 bits -- 28..32-bit samples,
depending on colour transform and decomposition count -- run on int64 lifting (general DWT kernels), 64-bit
sign-magnitude band planes and the 64-bit HT block coders (ojph_encode_codeblock64 / ojph_decode_codeblock64: fields of
up to 43 bits, 6-bit exponents, the 4-bit U-VLC extension).  Seeded random configurations against the unmodified
reference: byte-identical codestreams, identical decodes."""
import numpy as np
import pytest
import openjph_b200 as ob


def _case(seed):
    rng = np.random.default_rng(4000 + seed)
    nc = int(rng.choice([1, 3, 4])); bd = int(rng.choice([28, 29, 30, 31, 32, 32])); sg = bool(rng.random() < 0.3)
    w, h = int(rng.integers(1, 100)), int(rng.integers(1, 80))
    nd = int(rng.integers(0, 5)); ct = nc >= 3 and rng.random() < 0.6
    bw = int(rng.choice([4, 16, 32, 64, 128])); bh = int(rng.choice([b for b in (4, 8, 16, 32, 64) if b * bw <= 4096]))
    kw = dict(num_decomps=nd, reversible=True, color_transform=bool(ct), block=(bw, bh),
              prog_order=str(rng.choice(["LRCP", "RPCL", "CPRL"])))
    if rng.random() < 0.3:
        kw["tile"] = (int(rng.integers(16, 64)), int(rng.integers(16, 64)))
        kw["tlm"] = bool(rng.random() < 0.5)
    if sg and rng.random() < 0.4:
        kw["nlt"] = {"all": 3}
    lo, hi = (-(1 << (bd - 1)), (1 << (bd - 1))) if sg else (0, 1 << bd)
    fr = [rng.integers(lo, hi, (h, w), dtype=np.int64) for _ in range(nc)]
    fr = [(f >> int(rng.integers(0, 24))) if rng.random() < 0.4 else f for f in fr]       # some smooth / small planes
    planes = [(f & 0xFFFFFFFF).astype(np.uint32).astype(np.int32) for f in fr]               # bit patterns in si32 lines
    return ob.make_params(w, h, nc, bd, is_signed=sg, **kw), planes


def _check(lib, ref, seeds):
    wide = 0
    for seed in seeds:
        p, planes = _case(seed)
        want = ref.encode(p, planes)
        cs = ob.Encoder(p, ob.I32, lib=lib).encode(planes)
        assert cs == want, seed
        out = ob.Decoder(lib=lib).decode(want)
        refout, _ = ref.decode(want)
        for a, b in zip(out, refout):
            assert np.array_equal(a, b), seed
        wide += 1
    assert wide > 0


def test_wide_path_random_configs_emulator(emu_lib, ref):
    _check(emu_lib, ref, range(40))


@pytest.mark.gpu
def test_wide_path_random_configs_gpu(gpu_lib, ref):
    _check(None, ref, range(60))


@pytest.mark.gpu
def test_wide_path_larger_frame_gpu(gpu_lib, ref):
    """1024 x 768 RGB, 32-bit unsigned, RCT, 5 levels: every stage of the 64-bit path at a size with many blocks"""
    rng = np.random.default_rng(9)
    w, h = 1024, 768
    base = rng.integers(0, 1 << 32, (h // 8, w // 8), dtype=np.int64).repeat(8, 0).repeat(8, 1)
    planes = [((base + rng.integers(0, 1 << 20, (h, w), dtype=np.int64) * (c + 1)) & 0xFFFFFFFF).astype(np.uint32).astype(np.int32) for c in range(3)]
    p = ob.make_params(w, h, 3, 32, num_decomps=5, reversible=True, color_transform=True)
    want = ref.encode(p, planes)
    assert ob.Encoder(p, ob.I32).encode(planes) == want
    for a, b in zip(ob.Decoder().decode(want), planes):
        assert np.array_equal(a, b)

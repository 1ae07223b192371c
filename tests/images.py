"""deterministic synthetic frames: low-pass field + noise, planar"""
import numpy as np


def synth_plane(w, h, bit_depth, seed, noise=0.15):
    rng = np.random.default_rng(seed)
    lw, lh = max(2, (w + 7) // 8 + 3), max(2, (h + 7) // 8 + 3)
    low = rng.standard_normal((lh, lw))
    try:
        import cv2
        field = cv2.resize(low.astype(np.float32), (w + 24, h + 24), interpolation=cv2.INTER_CUBIC)[12:12 + h, 12:12 + w]
    except Exception:
        field = None
    if field is not None:
        pass
    else:
      try:
        from scipy.ndimage import zoom
        field = zoom(low, (h / (lh - 3) if lh > 3 else 8, w / (lw - 3) if lw > 3 else 8), order=3)[:h, :w]
        if field.shape != (h, w):
            field = np.resize(field, (h, w))
      except Exception:
        field = np.kron(low, np.ones((8, 8)))[:h, :w]
    field = field.astype(np.float32) + np.float32(noise) * rng.standard_normal((h, w), dtype=np.float32)
    lo, hi = field.min(), field.max()
    field = (field - lo) / (hi - lo if hi > lo else 1.0)
    return np.rint(field * ((1 << bit_depth) - 1)).astype(np.int32)


def synth_frame(w, h, nc, bit_depth, seed=1234, dims=None):
    if dims is None:
        dims = [(w, h)] * nc
    return [synth_plane(dw, dh, bit_depth, seed + 17 * c) for c, (dw, dh) in enumerate(dims)]


def noise_frame(w, h, nc, bit_depth, seed=1, signed=False):
    rng = np.random.default_rng(seed)
    off = (1 << (bit_depth - 1)) if signed else 0
    return [(rng.integers(0, 1 << bit_depth, (h, w)) - off).astype(np.int32) for _ in range(nc)]

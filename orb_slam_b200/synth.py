"""Deterministic synthetic u8 frames for tests and bench (numpy only; no reference code involved).

Multi-octave smoothed noise plus random filled rectangles: gives many more FAST corners per cell than
the per-cell quota so that retention, the threshold-7 fallback (flat regions) and the level-wide trim are
all exercised.
"""
import numpy as np


def _box_blur(a, r):
    if r <= 0:
        return a
    k = 2 * r + 1
    c = np.cumsum(np.pad(a, ((0, 0), (r + 1, r)), mode="reflect"), axis=1, dtype=np.float64)
    a = (c[:, k:] - c[:, :-k]) / k
    c = np.cumsum(np.pad(a, ((r + 1, r), (0, 0)), mode="reflect"), axis=0, dtype=np.float64)
    return (c[k:, :] - c[:-k, :]) / k


def textured_frame(width, height, seed=0, n_rects=None, flat_band=True):
    rng = np.random.default_rng(seed)
    img = np.zeros((height, width), np.float64)
    for octave, (r, amp) in enumerate(((0, 10.0), (1, 30.0), (3, 45.0), (8, 50.0), (20, 40.0))):
        n = rng.standard_normal((height, width))
        n = _box_blur(n, r)
        n /= (n.std() + 1e-9)
        img += amp * n
    img = 128.0 + img * (60.0 / img.std())
    n_rects = n_rects if n_rects is not None else max(8, (width * height) // 6000)
    for _ in range(n_rects):
        w = int(rng.integers(4, max(6, width // 12)))
        h = int(rng.integers(4, max(6, height // 12)))
        x = int(rng.integers(0, max(1, width - w)))
        y = int(rng.integers(0, max(1, height - h)))
        img[y:y + h, x:x + w] = rng.integers(0, 256)
    if flat_band:
        # a low-contrast band: cells here yield <=3 corners at th=20 and fall back to th=7
        y0 = height // 3
        band = slice(y0, y0 + max(8, height // 6))
        img[band, : width // 2] = 100.0 + 0.12 * (img[band, : width // 2] - 128.0)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def shifted_frame(frame, dx, dy, seed=0, noise=2.0):
    """A translated copy with a little sensor noise: consecutive frames of a stream."""
    rng = np.random.default_rng(seed)
    out = np.roll(frame, (dy, dx), axis=(0, 1)).astype(np.float64)
    out += noise * rng.standard_normal(out.shape)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def random_descriptors(n, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)


def noisy_copies(desc, flip_prob, seed=0):
    """Descriptors with independent random bit flips (realistic best/second-best ratios)."""
    rng = np.random.default_rng(seed)
    bits = np.unpackbits(desc, axis=1)
    flips = rng.random(bits.shape) < flip_prob
    return np.packbits(bits ^ flips, axis=1)

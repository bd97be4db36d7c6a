"""Deterministic synthetic gray frames and descriptor sets (SURVEY.md §8d 'Synthetic inputs').

numpy only: the same generator feeds the oracle, the parity tests and bench.py, here and on the GPU box.
"""
from __future__ import annotations

import numpy as np

def _cubic_w(t):
    """Catmull-Rom (a = -0.5) interpolation weight."""
    a = -0.5
    t = np.abs(t)
    return np.where(t <= 1, (a + 2) * t**3 - (a + 3) * t**2 + 1,
                    np.where(t < 2, a * t**3 - 5 * a * t**2 + 8 * a * t - 4 * a, 0.0))


def _upsample4(c: np.ndarray, axis: int) -> np.ndarray:
    """x4 Catmull-Rom up-sampling along `axis`: n+3 coarse samples -> 4*n fine samples.
    Fine sample 4i+ph sits at coarse position i + 1 + (ph + 0.5)/4, taps c[i..i+3]."""
    c = np.moveaxis(c, axis, 0)
    n = c.shape[0] - 3
    out = np.empty((4 * n,) + c.shape[1:], dtype=np.float64)
    for ph in range(4):
        f = (ph + 0.5) / 4.0
        w = _cubic_w(np.array([f + 1.0, f, 1.0 - f, 2.0 - f]))
        out[ph::4] = w[0] * c[0:n] + w[1] * c[1:n + 1] + w[2] * c[2:n + 2] + w[3] * c[3:n + 3]
    return np.moveaxis(out, 0, axis)


def textured_frame(seed: int, width: int, height: int, kind: str = "textured") -> np.ndarray:
    """kind: 'textured' (dense FAST-20 corners, octree saturated), 'lowcontrast' (exercises the
    minThFAST fallback), 'mixed' (left half textured, right half low contrast), 'constant' (0 keypoints)."""
    rng = np.random.default_rng(seed)
    if kind == "constant":
        return np.full((height, width), 128, dtype=np.uint8)
    ch, cw = (height + 3) // 4, (width + 3) // 4
    coarse = rng.uniform(0.0, 255.0, size=(ch + 3, cw + 3))
    img = _upsample4(_upsample4(coarse, 0), 1)[:height, :width]
    img = img + rng.normal(0.0, 6.0, size=img.shape)
    if kind == "lowcontrast":
        img = 100.0 + np.clip(img, 0, 255) * (30.0 / 255.0)
    elif kind == "mixed":
        lo = 100.0 + np.clip(img, 0, 255) * (30.0 / 255.0)
        ramp = np.clip((np.arange(width) - width * 0.45) / (width * 0.1), 0, 1)[None, :]
        img = img * (1 - ramp) + lo * ramp
    elif kind == "sparse":
        # smooth background + a few hundred bright blobs: cells with 0 and 1 corners, octree unsaturated
        img = 90.0 + 0.05 * img
        nb = max(8, width * height // 6000)
        ys = rng.integers(20, height - 20, nb)
        xs = rng.integers(20, width - 20, nb)
        for x, y in zip(xs, ys):
            s = int(rng.integers(2, 5))
            img[y:y + s, x:x + s] += float(rng.uniform(40, 150))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def shifted_frame(frame: np.ndarray, dx: int, dy: int, seed: int = 0) -> np.ndarray:
    """Integer-translated copy (+ light noise) used as the 'previous frame' for windowed matching."""
    rng = np.random.default_rng(seed + 7919)
    out = np.roll(np.roll(frame, dy, axis=0), dx, axis=1).astype(np.float64)
    out += rng.normal(0.0, 1.5, size=out.shape)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def descriptor_db(n: int, seed: int = 1234) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)


def queries_from_db(db: np.ndarray, nq: int, seed: int = 99, max_flips: int = 40):
    """nq DB rows with 0..max_flips random bit flips each. Returns (queries, source_row)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, db.shape[0], size=nq)
    q = db[src].copy()
    for i in range(nq):
        nf = int(rng.integers(0, max_flips + 1))
        if nf:
            bits = rng.choice(256, size=nf, replace=False)
            for b in bits:
                q[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return q, src.astype(np.int64)

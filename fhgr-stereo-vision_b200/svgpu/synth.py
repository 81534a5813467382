"""Deterministic synthetic stereo pairs, numpy only.

tex = normalize(GaussianBlur(uniform_u8(H, W+D), sigma=1.5), 0..255);
ground-truth disparity d(x,y) = 0.15 D + 0.6 D (0.5 + 0.5 sin(3 pi x/W) cos(2 pi y/H));
left = tex[:, D:D+W]; right(x,y) = bilinear(tex, x + D + d(x,y), y) + uniform noise in [-3,3].
Used by bench.py and the tests; seed = frame index.
"""
import numpy as np


def _blur(a, sigma=1.5, radius=5):
    t = np.arange(-radius, radius + 1, dtype=np.float32)
    k = np.exp(-(t * t) / (2 * sigma * sigma)).astype(np.float32)
    k /= k.sum()
    p = np.pad(a, ((0, 0), (radius, radius)), mode="reflect")
    h = sum(k[i] * p[:, i:i + a.shape[1]] for i in range(2 * radius + 1))
    p = np.pad(h, ((radius, radius), (0, 0)), mode="reflect")
    return sum(k[i] * p[i:i + a.shape[0], :] for i in range(2 * radius + 1))


def synth_pair(H, W, D, seed=0):
    """Returns (left, right) uint8 [H][W] with disparities in [0.15 D, 0.75 D]."""
    rng = np.random.default_rng(seed)
    tex = rng.integers(0, 256, (H, W + D), dtype=np.uint8).astype(np.float32)
    tex = _blur(tex)
    tex = (tex - tex.min()) * (255.0 / max(float(tex.max() - tex.min()), 1e-6))
    xs = np.arange(W, dtype=np.float32)[None, :]
    ys = np.arange(H, dtype=np.float32)[:, None]
    d = 0.15 * D + 0.6 * D * (0.5 + 0.5 * np.sin(3 * np.pi * xs / W) * np.cos(2 * np.pi * ys / H))
    left = np.clip(np.rint(tex[:, D:D + W]), 0, 255).astype(np.uint8)
    sx = np.minimum(xs + D + d, W + D - 1.001).astype(np.float32)
    x0 = np.floor(sx).astype(np.int64)
    fx = sx - x0
    rows = np.arange(H)[:, None]
    right = tex[rows, x0] * (1 - fx) + tex[rows, x0 + 1] * fx
    right = right + rng.integers(-3, 4, (H, W))
    right = np.clip(np.rint(right), 0, 255).astype(np.uint8)
    return left, right


def synth_batch(n, H, W, D, seed0=0):
    L = np.empty((n, H, W), np.uint8)
    R = np.empty((n, H, W), np.uint8)
    for i in range(n):
        L[i], R[i] = synth_pair(H, W, D, seed0 + i)
    return L, R

"""Diff of the per-pixel CUDA stages against the oracle (developer tool, run on a GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import svgpu  # noqa: E402
from oracle import sv_oracle as O  # noqa: E402
from svgpu.synth import synth_pair  # noqa: E402

rng = np.random.default_rng(3)
H, W = 97, 211
ctx = svgpu.Context(W, H)


def rep(name, got, want):
    n = int((np.asarray(got) != np.asarray(want)).sum())
    print(f"{name}: mismatches {n}/{np.asarray(want).size}", flush=True)


d = rng.integers(-16, 2000, (H, W)).astype(np.int16)
rep("median3", ctx.median3(d), O.median3(d))
b = (rng.integers(0, 6, (H // 4 + 1, W // 4 + 1)) * 40 - 16).astype(np.int16)
b = np.kron(b, np.ones((4, 4), np.int16))[:H, :W].copy()
for (nv, ms, md) in [(0, 150, 32), (-16, 60, 32), (-16, 20, 2), (0, 5, 100)]:
    rep(f"speckles {nv},{ms},{md}", ctx.filter_speckles(b, nv, ms, md), O.filter_speckles(b, nv, ms, md))
rep("close5", ctx.morph_close5(b), O.morph_close5(b))
K = np.array([[180.0, 0, 100.3], [0, 181.0, 50.2], [0, 0, 1]])
dist = np.array([-0.41, 0.16, 0.001, -0.002, 0.01])
R = np.eye(3)
P = np.array([[170.0, 0, 104.0, 0], [0, 170.0, 48.0, 0], [0, 0, 1, 0]])
m1, m2 = O.init_undistort_rectify_map(K, dist, R, P, (W, H))
ctx.set_rectify_maps(0, m1, m2)
ctx.set_rectify_maps(1, m1, m2)
img3 = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
rep("remap c3", ctx.remap(0, img3), O.remap(img3, m1, m2))
rep("remap c1", ctx.remap(1, img3[:, :, 1].copy()), O.remap(img3[:, :, 1].copy(), m1, m2))
big = rng.integers(0, 256, (H + 40, W + 60, 3), dtype=np.uint8)
rep("remap other-size src", ctx.remap(0, big), O.remap(big, m1, m2))
Q = np.array([[1, 0, 0, -100.3], [0, 1, 0, -50.2], [0, 0, 0, 170.0], [0, 0, 0.0132715, 0.0001]])
ctx.set_Q(Q)
for hm in (True, False):
    g, w = ctx.reproject(d, hm), O.reproject(d, Q, hm)
    print(f"reproject hm={hm}: bit mismatches {int((g.view(np.uint32) != w.view(np.uint32)).sum())}", flush=True)
ctx.set_depth_consts(75.336, 833.77)
g, w = ctx.depth(d), O.depth_simple(d, 75.336, 833.77)
print("depth bit mismatches", int((g.view(np.uint32) != w.view(np.uint32)).sum()), flush=True)
rep("prefilter chain", ctx.prefilter_chain(img3), O.sharpen3(O.bgr2gray(O.gaussian11(img3))))
src = rng.integers(0, 256, (H * 5 + 3, W * 5 + 1, 3), dtype=np.uint8)
rep("resize c3", ctx.resize(src), O.resize_linear(src, (W, H)))
rep("resize c1", ctx.resize(src[:, :, 0].copy()), O.resize_linear(src[:, :, 0].copy(), (W, H)))
for (D, bs, uq, tex) in [(32, 9, 15, 10), (48, 5, 0, 0), (16, 21, 15, 500), (64, 21, 15, 10)]:
    for (hh, ww) in [(40, 140), (41, 230)]:
        l, r = synth_pair(hh, ww, D, 5)
        c2 = svgpu.Context(ww, hh)
        c2.set_bm(numDisparities=D, blockSize=bs, uniquenessRatio=uq, textureThreshold=tex)
        rep(f"bm D{D} bs{bs} {hh}x{ww}", c2.bm(l, r), O.bm(l, r, numDisparities=D, blockSize=bs, uniquenessRatio=uq, textureThreshold=tex))
        c2.close()

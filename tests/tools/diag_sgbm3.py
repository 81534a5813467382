"""3-channel StereoSGBM (8UC3 input) on the GPU against the oracle: disparity mismatches per case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, svgpu
from oracle import sv_oracle as o
from svgpu.synth import synth_pair
rng = np.random.default_rng(0)
tot = 0
for (H, W, D, bs, minD) in [(40, 150, 32, 5, 0), (33, 120, 16, 3, 0), (50, 200, 64, 7, 4), (30, 140, 48, 5, -8), (60, 400, 128, 7, 0), (45, 420, 256, 5, 0), (37, 300, 224, 3, 0), (70, 230, 96, 1, 0)]:
    for mode in (0, 1, 3):
        for kind in ("synth", "noise"):
            if kind == "synth":
                l1, r1 = synth_pair(H, W, D, seed=H)
                l = np.stack([l1, np.roll(l1, 1, axis=0) // 2 + 30, 255 - l1], axis=2).astype(np.uint8)
                r = np.stack([r1, np.roll(r1, 1, axis=0) // 2 + 30, 255 - r1], axis=2).astype(np.uint8)
            else:
                l = rng.integers(0, 256, (H, W, 3), dtype=np.uint8); r = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            prm = dict(minDisparity=minD, numDisparities=D, blockSize=bs, P1=8 * bs * bs, P2=32 * bs * bs, disp12MaxDiff=1, preFilterCap=31,
                       uniquenessRatio=10, speckleWindowSize=0, speckleRange=0, mode=mode)
            want = o.sgbm(l, r, **prm)
            with svgpu.Context(W, H) as ctx:
                ctx.set_sgbm(**prm)
                got = ctx.sgbm(l, r)
            bad = int((got != want).sum()); tot += bad
            print((H, W, D, bs, minD), mode, kind, "mismatch", bad, "valid %.2f" % float((want >= minD * 16).mean()), flush=True)
print("TOTAL BAD", tot)

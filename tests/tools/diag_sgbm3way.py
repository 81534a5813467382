"""MODE_SGBM_3WAY on the GPU against the oracle: disparity mismatches per case (textured and noise pairs, tiny images)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, svgpu
from oracle import sv_oracle as o
from svgpu.synth import synth_pair
rng = np.random.default_rng(11)
tot = 0
cfgs = [(64, 160, 32, 200, 800, 0, 5, 0, 1), (64, 200, 48, 200, 800, 10, 5, 0, 1), (64, 160, 32, 200, 800, 15, 1, 0, 1), (30, 160, 32, 200, 800, 10, 5, 0, 1),
        (41, 150, 16, 8, 32, 10, 3, 0, 1), (57, 180, 64, 392, 1568, 15, 7, 0, 1), (50, 170, 32, 100, 400, 10, 5, 5, 1), (50, 170, 32, 100, 400, 10, 5, -6, 1),
        (50, 190, 32, 100, 400, 10, 5, 16, 2), (9, 120, 16, 100, 400, 10, 5, 0, 1), (5, 120, 16, 100, 400, 10, 5, 0, 1), (13, 120, 16, 100, 400, 10, 7, 0, 1),
        (3, 120, 16, 100, 400, 10, 3, 0, 1), (1, 120, 16, 100, 400, 10, 3, 0, 1), (2, 120, 16, 100, 400, 10, 9, 0, 1), (17, 120, 16, 100, 400, 10, 11, 0, 1),
        (120, 300, 128, 392, 1568, 15, 7, 0, 1), (100, 400, 224, 392, 1568, 15, 7, 0, 1), (40, 140, 32, 2, 5, 5, 1, 0, 1), (40, 140, 32, 100, 400, 50, 5, 0, 1),
        (40, 140, 32, 100, 400, 99, 5, 0, 1), (40, 250, 96, 50, 200, 10, 3, -40, 1), (70, 420, 256, 50, 200, 10, 3, 0, 1), (300, 500, 128, 392, 1568, 15, 7, 0, 1)]
for (H, W, D, P1, P2, uniq, bs, minD, d12) in cfgs:
    for kind in ("synth", "noise"):
        if kind == "synth":
            l, r = synth_pair(H, W, max(D + minD, 8), seed=H + D)
        else:
            l = rng.integers(0, 256, (H, W), dtype=np.uint8); r = rng.integers(0, 256, (H, W), dtype=np.uint8)
        prm = dict(minDisparity=minD, numDisparities=D, blockSize=bs, P1=P1, P2=P2, disp12MaxDiff=d12, preFilterCap=31, uniquenessRatio=uniq,
                   speckleWindowSize=0, speckleRange=0, mode=2)
        want = o.sgbm(l, r, **prm)
        with svgpu.Context(W, H) as ctx:
            ctx.set_sgbm(**prm)
            got = ctx.sgbm(l, r)
            got2 = ctx.sgbm_batch(np.stack([l, l, l]), np.stack([r, r, r]))
        bad = (got != want); tot += int(bad.sum()) + int((got2 != want[None]).sum())
        print((H, W, D, P1, P2, uniq, bs, minD, d12), kind, "mismatch", int(bad.sum()), "batch", int((got2 != want[None]).sum()),
              "rows", sorted(set(np.argwhere(bad)[:, 0].tolist()))[:12], "valid %.2f" % float((want >= minD * 16).mean()), flush=True)
print("TOTAL BAD", tot)

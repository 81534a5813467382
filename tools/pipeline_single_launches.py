#!/usr/bin/env python
"""One full-chain call per pair at the reference's setting (640x480, 224 disparities, MODE_HH4), host buffers: run under
`ncu --metrics gpu__time_duration.sum` for the kernel list of a call, or plain for the wall-clock time (developer tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, svgpu
from svgpu.synth import synth_pair
from bench import big_calib, to_bgr

W, H, D = 640, 480, 224
ALL = (svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM | svgpu.STAGE_POSTFILTER | svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH)
l, r = synth_pair(H, W, D, seed=1)
l, r = to_bgr(l[None])[0], to_bgr(r[None])[0]
K, dd, P, Q = big_calib(W, H)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
with svgpu.Context(W, H) as ctx:
    ctx.set_sgbm(numDisparities=D, mode=3, minDisparity=0, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32, uniquenessRatio=15,
                 speckleWindowSize=200, speckleRange=2)
    for side in (0, 1):
        ctx.init_rectify_maps(side, K, dd, np.eye(3), P)
    ctx.set_Q(Q); ctx.set_depth_consts(75.3, float(K[0, 0]))
    for _ in range(2):
        ctx.pipeline(l, r, ALL, want_xyz=True, want_depth=True)
    for stages, name in ((ALL, "all stages"), (ALL & ~(svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH), "without XYZ / depth"),
                         (svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM, "rectify + pre-filter + SGBM")):
        t0 = time.perf_counter()
        for _ in range(n):
            ctx.pipeline(l, r, stages, want_xyz=bool(stages & svgpu.STAGE_REPROJECT), want_depth=bool(stages & svgpu.STAGE_DEPTH))
        print(f"{name}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms per call", flush=True)

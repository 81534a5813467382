#!/usr/bin/env python
"""Runs every non-SGBM kernel of the chain once at a production size so that `ncu --set full` can capture them (developer
tool): remap (BGR, 4K), blur / gray / sharpen, speckle filter, close, reproject, depth at 3840x2160; StereoBM at 1280x720 / 64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, torch, svgpu
from svgpu.synth import synth_pair

W, H, D = 3840, 2160, 256
rng = np.random.default_rng(0)
l, r = synth_pair(H, W, D, seed=4)
lb = torch.from_numpy(np.stack([l, np.roll(l, 1, axis=0) // 2 + 30, 255 - l], axis=2).astype(np.uint8)).cuda()
K = np.array([[833.77 * 6, 0, 345.2 * 6], [0, 832.87 * 6, 283.1 * 4.5], [0, 0, 1]])
P = np.array([[K[0, 0] * 0.9, 0, K[0, 2], 0], [0, K[0, 0] * 0.9, K[1, 2], 0], [0, 0, 1, 0]])
Q = np.array([[1, 0, 0, -K[0, 2]], [0, 1, 0, -K[1, 2]], [0, 0, 0, P[0, 0]], [0, 0, 1 / 75.3, 0]])
disp = torch.from_numpy((rng.integers(0, 6, (H // 8, W // 8)) * 300 - 16).astype(np.int16).repeat(8, 0).repeat(8, 1) +
                        rng.integers(0, 3, (H, W)).astype(np.int16)).cuda()
with svgpu.Context(W, H) as ctx:
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for side in (0, 1):
        ctx.init_rectify_maps(side, K, np.array([-0.4168, 0.166, 0, 0, 0]), np.eye(3), P)
    ctx.set_Q(Q); ctx.set_depth_consts(75.3, float(K[0, 0]))
    for _ in range(2):
        rect = ctx.remap(0, lb)
        gray = ctx.prefilter_chain(rect)
        d = ctx.filter_speckles(disp, 0, 150, 32)
        d = ctx.morph_close5(d)
        xyz = ctx.reproject(d, True)
        z = ctx.depth(d)
        m = ctx.median3(d)
    torch.cuda.synchronize()
W, H, D = 1280, 720, 64
l, r = synth_pair(H, W, D, seed=2)
with svgpu.Context(W, H) as ctx:
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_bm(numDisparities=64, blockSize=21)
    tl, tr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    for _ in range(2):
        d = ctx.bm(tl, tr)
    torch.cuda.synchronize()
print("misc kernels done")

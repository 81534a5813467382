#!/usr/bin/env python
"""Latency of one StereoVision::AnalyzeAndAnnotateImage-shaped call (src/lib.cpp:53-85, kSGBM) through svgpu_pipeline on the
reference's bottle pair (tests/golden/bottle.npz, 640x480, D=224, MODE_HH4): raw BGR frames in host memory -> rectified
images, disparity, XYZ and depth in host memory.  Prints the whole call and the cost of each stage (by leaving it out).
Developer tool; the driver's bench is bench.py."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, torch, svgpu

g = np.load(os.path.join(ROOT, "tests", "golden", "bottle.npz"))
W, H = 640, 480
S = svgpu
ALL = S.STAGE_REMAP | S.STAGE_PREFILTER | S.STAGE_SGBM | S.STAGE_POSTFILTER | S.STAGE_REPROJECT | S.STAGE_DEPTH


def wall(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


with svgpu.Context(W, H) as ctx:
    R1, R2, P1, P2, Q = svgpu.stereo_rectify(g["cal_Kl"], g["cal_Dl"].reshape(-1)[:5], g["cal_Kr"], g["cal_Dr"].reshape(-1)[:5], (W, H), g["cal_R"], g["cal_T"])
    t0 = time.perf_counter()
    ctx.init_rectify_maps(0, g["cal_Kl"], g["cal_Dl"].reshape(-1)[:5], R1, P1)
    ctx.init_rectify_maps(1, g["cal_Kr"], g["cal_Dr"].reshape(-1)[:5], R2, P2)
    print(f"two rectification maps on the device: {(time.perf_counter() - t0) * 1e3:.2f} ms (first call, includes allocation)")
    ctx.set_Q(Q); ctx.set_depth_consts(float(g["baseline"]), float(g["focal"])); ctx.set_sgbm(**svgpu.REFERENCE_SGBM)
    l = torch.from_numpy(np.ascontiguousarray(g["scaled_left"])).pin_memory().numpy()
    r = torch.from_numpy(np.ascontiguousarray(g["scaled_right"])).pin_memory().numpy()
    full = wall(lambda: ctx.pipeline(l, r, ALL, want_rect=True, want_xyz=True, want_depth=True))
    print(f"whole call, host in / host out (rectified pair, disparity, XYZ, depth): {full:.2f} ms  ({1e3 / full:.0f} calls/s)")
    for name, bit in (("remap", S.STAGE_REMAP), ("speckle filter + close", S.STAGE_POSTFILTER),
                      ("reproject", S.STAGE_REPROJECT), ("depth", S.STAGE_DEPTH)):
        t = wall(lambda: ctx.pipeline(l, r, ALL & ~bit, want_rect=bool(bit != S.STAGE_REMAP), want_xyz=bool(bit != S.STAGE_REPROJECT),
                                      want_depth=bool(bit != S.STAGE_DEPTH)))
        print(f"  without {name:24s}: {t:.2f} ms  (stage + its copies ~ {full - t:.2f} ms)")
    ld, rd = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    dev = wall(lambda: ctx.pipeline(ld, rd, ALL, want_rect=True, want_xyz=True, want_depth=True))
    print(f"whole call, device in / device out: {dev:.2f} ms")
    gl, gr = torch.from_numpy(np.ascontiguousarray(g["gray_left"])).cuda(), torch.from_numpy(np.ascontiguousarray(g["gray_right"])).cuda()
    sg = wall(lambda: ctx.sgbm(gl, gr))
    print(f"StereoSGBM::compute alone (device): {sg:.2f} ms")

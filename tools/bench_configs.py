#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one GPU (developer tool; the driver's bench is bench.py):
  config 2: 1280x720, 64 disparities — StereoBM and SGBM MODE_SGBM (5 paths)
  config 4: 3840x2160, 256 disparities — rectify -> SGBM MODE_HH -> reprojectImageTo3D (svgpu_pipeline per pair)
  reference configuration (src/lib.cpp:248-259): 640x480, 224 disparities, MODE_HH4
Device-resident inputs, CUDA events, synthetic pairs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, torch, svgpu
from svgpu.synth import synth_pair

dev = torch.device("cuda", 0)
REF = dict(minDisparity=0, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32, uniquenessRatio=15, speckleWindowSize=0, speckleRange=0)


def timed(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(n)]; e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def sgbm(W, H, D, mode, B):
    L = np.empty((B, H, W), np.uint8); R = np.empty_like(L)
    for k in range(min(B, 4)): L[k], R[k] = synth_pair(H, W, D, seed=k)
    for k in range(4, B): L[k], R[k] = L[k % 4], R[k % 4]
    Ld, Rd = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
    out = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(numDisparities=D, mode=mode, **REF); ctx.reserve(B)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ms = timed(lambda: ctx.sgbm_batch(Ld, Rd, out=out))
    return B / ms * 1e3, W * H * D * B / ms / 1e3


if __name__ == "__main__":
    if "--latency" in sys.argv:   # one pair per call, the way the reference's front-ends call the library
        for (W, H, D, mode, name) in [(640, 480, 224, 3, "reference 640x480 D=224 HH4"), (1280, 720, 64, 0, "config 2 1280x720 D=64 SGBM"),
                                      (1920, 1080, 128, 1, "config 3 1920x1080 D=128 HH"), (3840, 2160, 256, 1, "config 4 3840x2160 D=256 HH")]:
            for B in ((1, 2, 4, 8, 16, 32) if W < 3840 else (1, 2, 4, 8)):
                fps, _ = sgbm(W, H, D, mode, B)
                print(f"{name}: {B} pair(s) per call: {B / fps * 1e3:7.2f} ms per call, {fps:7.1f} fps", flush=True)
        # the whole chain for one pair per call with host buffers, wall clock: what one AnalyzeAndAnnotateImage call costs
        from bench import big_calib, to_bgr
        ALL = (svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM | svgpu.STAGE_POSTFILTER | svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH)
        for (W, H, D, mode, name) in [(640, 480, 224, 3, "reference 640x480 D=224 HH4"), (1920, 1080, 128, 1, "config 3 1920x1080 D=128 HH")]:
            l, r = synth_pair(H, W, D, seed=1)
            l, r = to_bgr(l[None])[0], to_bgr(r[None])[0]
            K, dd, P, Q = big_calib(W, H)
            with svgpu.Context(W, H) as ctx:
                ctx.set_sgbm(numDisparities=D, mode=mode, **dict(REF, speckleWindowSize=200, speckleRange=2))
                for side in (0, 1):
                    ctx.init_rectify_maps(side, K, dd, np.eye(3), P)
                ctx.set_Q(Q); ctx.set_depth_consts(75.3, float(K[0, 0]))
                for _ in range(3):
                    ctx.pipeline(l, r, ALL, want_xyz=True, want_depth=True)
                t0 = time.perf_counter()
                for _ in range(20):
                    ctx.pipeline(l, r, ALL, want_xyz=True, want_depth=True)
                ms = (time.perf_counter() - t0) / 20 * 1e3
            print(f"{name}: full chain (BGR in; disparity + XYZ + depth out), one pair per call, host buffers, wall clock: {ms:7.2f} ms per call", flush=True)
        sys.exit(0)
    fps, mpd = sgbm(1280, 720, 64, 0, 64); print(f"config 2  SGBM MODE_SGBM 1280x720 D=64, 64 pairs/step: {fps:8.1f} fps {mpd:10.0f} Mpix*disp/s")
    fps, mpd = sgbm(1280, 720, 64, 1, 64); print(f"          SGBM MODE_HH   1280x720 D=64, 64 pairs/step: {fps:8.1f} fps {mpd:10.0f} Mpix*disp/s")
    fps, mpd = sgbm(640, 480, 224, 3, 64); print(f"reference SGBM MODE_HH4  640x480 D=224, 64 pairs/step: {fps:8.1f} fps {mpd:10.0f} Mpix*disp/s")
    fps, mpd = sgbm(3840, 2160, 256, 1, 4); print(f"config 4  SGBM MODE_HH  3840x2160 D=256, 4 pairs/step: {fps:8.1f} fps {mpd:10.0f} Mpix*disp/s")
    W, H = 1280, 720
    l, r = synth_pair(H, W, 64, seed=1)
    ld, rd = torch.from_numpy(l).to(dev), torch.from_numpy(r).to(dev)
    with svgpu.Context(W, H) as ctx:
        ctx.set_bm(minDisparity=0, numDisparities=64, blockSize=21, preFilterCap=31, textureThreshold=10, uniquenessRatio=15, speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ms = timed(lambda: ctx.bm(ld, rd), n=20)
    print(f"config 2  StereoBM bs21 1280x720 D=64, one pair per call: {1e3 / ms:8.1f} fps")

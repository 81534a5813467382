#!/usr/bin/env python
"""Stage split (cost volume / aggregation / selection) of the other configurations at a throughput batch (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, torch, svgpu
from svgpu.synth import synth_pair
REF = dict(minDisparity=0, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32, uniquenessRatio=15, speckleWindowSize=0, speckleRange=0)
for (W, H, D, mode, B) in [(640, 480, 224, 3, 64), (1280, 720, 64, 0, 64), (1280, 720, 64, 1, 64), (1280, 720, 64, 2, 64), (1920, 1080, 128, 0, 64), (1920, 1080, 128, 3, 64), (1920, 1080, 128, 2, 32)]:
    L = np.empty((B, H, W), np.uint8); R = np.empty_like(L)
    for k in range(4): L[k], R[k] = synth_pair(H, W, D, seed=k)
    for k in range(4, B): L[k], R[k] = L[k % 4], R[k % 4]
    Ld, Rd = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    out = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(numDisparities=D, mode=mode, **REF); ctx.reserve(B)
        ctx.enable_stage_timing(True)
        st = []
        for i in range(5):
            ctx.sgbm_batch(Ld, Rd, out=out); st.append(ctx.last_stage_ms())
        m = {k: float(np.median([s[k] for s in st[1:]])) for k in st[0]}
        print(f"{W}x{H} D={D} mode={mode} B={B}: {B / m['total'] * 1e3:8.1f} fps  cost {m['cost']:.2f} agg {m['aggregate']:.2f} sel {m['select']:.2f} ms", flush=True)

import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, torch, svgpu
from svgpu.synth import synth_pair
REF = dict(minDisparity=0, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32, uniquenessRatio=15, speckleWindowSize=0, speckleRange=0)
for (W,H,D,mode) in [(640,480,224,3),(640,480,224,1),(1280,720,64,0),(1920,1080,128,1)]:
    l,r = synth_pair(H,W,D,seed=1)
    ld, rd = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    with svgpu.Context(W,H) as ctx:
        ctx.set_sgbm(numDisparities=D, mode=mode, **REF)
        ctx.enable_stage_timing(True)
        st=[]
        for i in range(6):
            ctx.sgbm(ld, rd); st.append(ctx.last_stage_ms())
        print((W,H,D,mode), {k: round(float(np.median([s[k] for s in st[2:]])),3) for k in st[0]})

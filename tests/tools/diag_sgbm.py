"""Stage-by-stage diff of the CUDA SGBM path against the oracle (developer tool, run on a GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import svgpu  # noqa: E402
from oracle import sv_oracle as O  # noqa: E402
from svgpu.synth import synth_pair  # noqa: E402


def run(H, W, D, mode, noise=False, **over):
    prm = dict(minDisparity=0, numDisparities=D, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32,
               uniquenessRatio=15, speckleWindowSize=0, speckleRange=0, mode=mode)
    prm.update(over)
    if noise:
        rng = np.random.default_rng(7)
        l = rng.integers(0, 256, (H, W), dtype=np.uint8)
        r = rng.integers(0, 256, (H, W), dtype=np.uint8)
    else:
        l, r = synth_pair(H, W, D, 1)
    ref, Cv, Sv = O.sgbm_debug(l, r, **prm)
    ctx = svgpu.Context(W, H)
    ctx.set_sgbm(**prm)
    ctx.debug_keep_volume(True)
    got = ctx.sgbm(l, r)
    out = [f"H{H} W{W} D{D} mode{mode} noise{int(noise)} {over}"]
    if Cv.size:
        gC = ctx.debug_volume(0, Cv.shape)
        gS = ctx.debug_volume(1, Sv.shape)
        nc, ns = int((gC != Cv).sum()), int((gS != Sv).sum())
        out.append(f"C mismatches {nc}/{Cv.size}  S mismatches {ns}/{Sv.size}")
        if nc:
            idx = np.argwhere(gC != Cv)
            out.append(f"  first C mismatches {idx[:5].tolist()} got {gC[tuple(idx[0])]} want {Cv[tuple(idx[0])]}"
                       f" rows {np.unique(idx[:,0])[:10].tolist()} cols {np.unique(idx[:,1])[:10].tolist()}")
        elif ns:
            idx = np.argwhere(gS != Sv)
            out.append(f"  first S mismatches {idx[:5].tolist()} got {gS[tuple(idx[0])]} want {Sv[tuple(idx[0])]}")
    nd = int((got != ref).sum())
    out.append(f"disp mismatches {nd}/{ref.size} valid {float((ref >= 0).mean()):.3f}")
    if nd:
        idx = np.argwhere(got != ref)
        out.append(f"  first {idx[:8].tolist()} got {got[tuple(idx[0])]} want {ref[tuple(idx[0])]}")
    ctx.close()
    print(" | ".join(out), flush=True)
    return nd


if __name__ == "__main__":
    bad = 0
    for mode in (3, 0, 1):
        bad += run(40, 120, 32, mode)
        bad += run(300, 500, 128, mode)
        bad += run(70, 300, 256, mode)
        bad += run(300, 120, 16, mode)
        bad += run(33, 100, 16, mode, noise=True, blockSize=3, P1=8, P2=32, uniquenessRatio=0)
        bad += run(61, 333, 64, mode)
        bad += run(50, 400, 224, mode)
        bad += run(37, 150, 48, mode, minDisparity=4)
        bad += run(37, 150, 48, mode, minDisparity=-8, blockSize=5)
    print("TOTAL BAD", bad)

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SGBM_KEYS = ["minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
             "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode"]
BM_KEYS = ["minDisparity", "numDisparities", "blockSize", "preFilterCap", "textureThreshold", "uniquenessRatio",
           "speckleWindowSize", "speckleRange", "disp12MaxDiff"]

_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
    return _cache[name]


def sgbm_golden_cases():
    g = golden("sgbm_small")
    for i in range(int(g["n"])):
        yield i, g[f"l{i}"], g[f"r{i}"], dict(zip(SGBM_KEYS, (int(v) for v in g[f"p{i}"]))), g[f"d{i}"]


def sgbm_wide_cases():
    """cv2 cases with 272 .. 512 disparities (gray, 8UC3, blockSize 11 / 13); tests/golden/make_golden.py: save_sgbm_wide"""
    g = golden("sgbm_wide")
    for i in range(int(g["n"])):
        yield i, g[f"l{i}"], g[f"r{i}"], dict(zip(SGBM_KEYS, (int(v) for v in g[f"p{i}"]))), g[f"d{i}"]


def bm_golden_cases():
    g = golden("bm_small")
    for i in range(int(g["n"])):
        yield i, g[f"l{i}"], g[f"r{i}"], dict(zip(BM_KEYS, (int(v) for v in g[f"p{i}"]))), g[f"d{i}"]


def nmismatch(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return int((a != b).sum())


REF_PARAMS = dict(blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32, uniquenessRatio=15,
                  speckleWindowSize=0, speckleRange=0, minDisparity=0)


def bm_mind_golden_cases():
    """StereoBM with minDisparity != 0: (index, left, right, params, cv2 disparity, (y0, y1, x0, x1) of cv2's valid ROI)."""
    g = golden("bm_mind_small")
    for i in range(int(g["n"])):
        yield i, g[f"l{i}"], g[f"r{i}"], dict(zip(BM_KEYS, (int(v) for v in g[f"p{i}"]))), g[f"d{i}"], tuple(int(v) for v in g[f"roi{i}"])


def sgbm3_golden_cases():
    """StereoSGBM on 8UC3 pairs: (index, left, right, params, cv2 disparity)."""
    g = golden("sgbm3_small")
    for i in range(int(g["n"])):
        yield i, g[f"l{i}"], g[f"r{i}"], dict(zip(SGBM_KEYS, (int(v) for v in g[f"p{i}"]))), g[f"d{i}"]


def sgbm3way_golden_cases():
    """MODE_SGBM_3WAY: (index, left, right, params, cv2 disparity)."""
    g = golden("sgbm3way_small")
    for i in range(int(g["n"])):
        yield i, g[f"l{i}"], g[f"r{i}"], dict(zip(SGBM_KEYS, (int(v) for v in g[f"p{i}"]))), g[f"d{i}"]

"""ctypes front-end of the CPU oracle (oracle/libsv_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from the product package
(fhgr-stereo-vision_b200/).  Each wrapper restates one OpenCV call made by
/root/reference/src/lib.cpp (citations in sv_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_SGBM, MODE_HH, MODE_SGBM_3WAY, MODE_HH4 = 0, 1, 2, 3


class SgbmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]


class BmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "preFilterCap", "textureThreshold",
        "uniquenessRatio", "speckleWindowSize", "speckleRange", "disp12MaxDiff")]


def build(force=False):
    so = os.path.join(_HERE, "libsv_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("sv_oracle_sgbm.c", "sv_oracle_misc.c", "sv_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _u8(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8
    return a


def sgbm(left, right, **kw):
    """cv::StereoSGBM::compute — src/lib.cpp:248-262."""
    left, right = _u8(left), _u8(right)
    H, W = left.shape[:2]
    cn = 1 if left.ndim == 2 else left.shape[2]   # 8UC1 (the reference's call) or 8UC3
    prm = SgbmParams(**kw)
    disp = np.empty((H, W), np.int16)
    rc = lib().svo_sgbm_compute_cn(_p(left), _p(right), W, H, W * cn, W * cn, cn, C.byref(prm), _p(disp), W * 2)
    if rc:
        raise ValueError(f"svo_sgbm_compute rc={rc}")
    return disp


def sgbm_debug(left, right, **kw):
    """sgbm() plus the oracle's Ctrue and S volumes ([H][W1][D] int16) for stage-by-stage diffs."""
    left, right = _u8(left), _u8(right)
    H, W = left.shape
    D, minD = kw["numDisparities"], kw["minDisparity"]
    W1 = (W + min(minD, 0)) - max(minD + D, 0)
    Cv = np.zeros((H, max(W1, 0), D), np.int16)
    Sv = np.zeros_like(Cv)
    L = lib()
    pc = C.c_void_p.in_dll(L, "svo_debug_C")
    ps = C.c_void_p.in_dll(L, "svo_debug_S")
    pc.value, ps.value = Cv.ctypes.data, Sv.ctypes.data
    try:
        disp = sgbm(left, right, **kw)
    finally:
        pc.value, ps.value = None, None
    return disp, Cv, Sv


def bm(left, right, **kw):
    """cv::StereoBM::compute — BASELINE config 2."""
    left, right = _u8(left), _u8(right)
    H, W = left.shape
    d = dict(minDisparity=0, numDisparities=64, blockSize=21, preFilterCap=31, textureThreshold=10,
             uniquenessRatio=15, speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1)
    d.update(kw)
    prm = BmParams(**d)
    disp = np.empty((H, W), np.int16)
    rc = lib().svo_bm_compute(_p(left), _p(right), W, H, W, W, C.byref(prm), _p(disp), W * 2)
    if rc:
        raise ValueError(f"svo_bm_compute rc={rc}")
    return disp


def median3(src):
    src = np.ascontiguousarray(src, np.int16)
    dst = np.empty_like(src)
    lib().svo_median3_s16(_p(src), _p(dst), src.shape[1], src.shape[0])
    return dst


def filter_speckles(img, new_val, max_size, max_diff):
    out = np.ascontiguousarray(img, np.int16).copy()
    lib().svo_filter_speckles(_p(out), out.shape[1], out.shape[0], int(new_val), int(max_size), int(max_diff))
    return out


def init_undistort_rectify_map(K, dist, R, P, size):
    """cv::initUndistortRectifyMap(..., CV_16SC2) — src/lib.cpp:39-46."""
    W, H = size
    K = np.ascontiguousarray(K, np.float64).reshape(3, 3)
    d = np.zeros(5, np.float64)
    dd = np.asarray(dist, np.float64).ravel()
    d[:min(5, dd.size)] = dd[:5]
    R = np.ascontiguousarray(R, np.float64).reshape(3, 3)
    P = np.ascontiguousarray(P, np.float64).reshape(3, 4)
    m1 = np.empty((H, W, 2), np.int16)
    m2 = np.empty((H, W), np.uint16)
    lib().svo_init_undistort_rectify_map(_p(K), _p(d), _p(R), _p(P), W, H, _p(m1), _p(m2))
    return m1, m2


def remap(src, map1, map2):
    """cv::remap(INTER_LINEAR) — src/lib.cpp:89-92."""
    src = _u8(src)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    H, W = map2.shape
    map1 = np.ascontiguousarray(map1, np.int16)
    map2 = np.ascontiguousarray(map2, np.uint16)
    dst = np.empty((H, W) if src.ndim == 2 else (H, W, cn), np.uint8)
    lib().svo_remap_u8(_p(src), sw, sh, sw * cn, cn, _p(map1), _p(map2), _p(dst), W, H, W * cn)
    return dst


def reproject(disp, Q, handle_missing=True):
    """cv::reprojectImageTo3D — src/lib.cpp:154."""
    disp = np.ascontiguousarray(disp, np.int16)
    Q = np.ascontiguousarray(Q, np.float64).reshape(4, 4)
    H, W = disp.shape
    out = np.empty((H, W, 3), np.float32)
    lib().svo_reproject_s16(_p(disp), W, H, _p(Q), int(handle_missing), _p(out))
    return out


def depth_simple(disp, baseline, focal):
    """StereoVision::CalculateDepthMapSimple — src/lib.cpp:279-294."""
    disp = np.ascontiguousarray(disp, np.int16)
    H, W = disp.shape
    out = np.empty((H, W), np.float32)
    f = lib().svo_depth_simple
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    f(_p(disp), W, H, float(baseline), float(focal), _p(out))
    return out


def gaussian11(src):
    src = _u8(src)
    cn = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty_like(src)
    lib().svo_gaussian11_u8(_p(src), _p(dst), src.shape[1], src.shape[0], cn)
    return dst


def bgr2gray(bgr):
    bgr = _u8(bgr)
    out = np.empty(bgr.shape[:2], np.uint8)
    lib().svo_bgr2gray(_p(bgr), _p(out), bgr.shape[1], bgr.shape[0])
    return out


def sharpen3(src):
    src = _u8(src)
    dst = np.empty_like(src)
    lib().svo_sharpen3_u8(_p(src), _p(dst), src.shape[1], src.shape[0])
    return dst


def morph_close5(src):
    src = np.ascontiguousarray(src, np.int16)
    dst = np.empty_like(src)
    lib().svo_morph_close5_s16(_p(src), _p(dst), src.shape[1], src.shape[0])
    return dst


def resize_linear(src, dsize):
    src = _u8(src)
    cn = 1 if src.ndim == 2 else src.shape[2]
    dw, dh = dsize
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().svo_resize_linear_u8(_p(src), src.shape[1], src.shape[0], cn, _p(dst), dw, dh)
    return dst


REFERENCE_SGBM = dict(minDisparity=0, numDisparities=224, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1,
                      preFilterCap=32, uniquenessRatio=15, speckleWindowSize=0, speckleRange=2, mode=MODE_HH4)


def reference_sgbm_chain(left_bgr_rect, right_bgr_rect, params=None):
    """StereoVision::CalculateDisparityMapUsingSGBM — src/lib.cpp:228-277 (without the debug imshow)."""
    prm = dict(REFERENCE_SGBM)
    prm.update(params or {})
    gl = sharpen3(bgr2gray(gaussian11(left_bgr_rect)))
    gr = sharpen3(bgr2gray(gaussian11(right_bgr_rect)))
    d = sgbm(gl, gr, **prm)
    # src/lib.cpp:264 `setTo(0, disparity == -1)` (a no-op in practice, kept for fidelity)
    d = d.copy()
    d[d == -1] = 0
    d = filter_speckles(d, 0, 150, 32)
    return morph_close5(d)

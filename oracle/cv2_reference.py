"""The reference's own CPU implementation of the path, as far as it can be run here.

TEST / BASELINE INFRASTRUCTURE ONLY (same rules as the rest of oracle/).  The reference's
src/lib.cpp owns no arithmetic: line 262 dispatches to cv::StereoSGBM::compute of the OpenCV it
links (vcpkg `opencv` >= 4.11).  The library cannot be compiled in this image (no OpenCV C++
headers), but the same OpenCV code is installed as the cv2 4.13.0 wheel, so the CPU baseline calls
it through the Python binding with the reference's parameters (src/lib.cpp:248-259, mode and
numDisparities per BASELINE.json config).  When cv2 is not importable the plain-C port
(sv_oracle.sgbm) is timed instead and reported as kind "port".
"""
import os
import time

import numpy as np


def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


_MAPS = {}


def _one(args):
    l, r, prm, use_cv2, chain, calib = args
    if not use_cv2:
        from oracle import sv_oracle
        return sv_oracle.bm(l, r, **prm) if chain == "bm" else sv_oracle.sgbm(l, r, **prm)
    import cv2
    cv2.setNumThreads(1)
    if chain == "bm":     # cv::StereoBM::compute (BASELINE configs[1])
        return cv2.StereoBM_create(prm["numDisparities"], prm["blockSize"]).compute(l, r)
    if chain != "pipeline":
        return cv2.StereoSGBM_create(**prm).compute(l, r)
    # StereoVision::AnalyzeAndAnnotateImage with kSGBM on raw BGR frames, src/lib.cpp:53-85 (the mode / disparity count of the config)
    import numpy as np
    K, dist, P, Q = calib
    key = (l.shape, K.tobytes())
    if key not in _MAPS:
        _MAPS[key] = cv2.initUndistortRectifyMap(K, dist, np.eye(3), P, (l.shape[1], l.shape[0]), cv2.CV_16SC2)
    m1, m2 = _MAPS[key]
    k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
    g = [cv2.filter2D(cv2.cvtColor(cv2.GaussianBlur(cv2.remap(x, m1, m2, cv2.INTER_LINEAR), (11, 11), 1.4), cv2.COLOR_BGR2GRAY), -1, k) for x in (l, r)]
    d = cv2.StereoSGBM_create(**prm).compute(g[0], g[1])
    d[d == -1] = 0
    cv2.filterSpeckles(d, 0, 150, 32)
    d = cv2.morphologyEx(d, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5)))
    xyz = cv2.reprojectImageTo3D(d, Q, handleMissingValues=True)
    bf = np.float32(75.3) * np.float32(K[0, 0])
    with np.errstate(divide="ignore"):
        depth = np.where(d > 0, bf / d.astype(np.float32), np.float32(0)).astype(np.float32)
    return d, xyz[::64, ::64].copy(), depth[::64, ::64].copy()   # the full maps stay in the worker (pickling them is not part of the path)


class CpuSgbm:
    """A pool of worker processes, one pair each (cv2's MODE_SGBM / MODE_HH are single-threaded inside OpenCV,
    SURVEY.md [B#12], so whole-host throughput needs independent processes)."""

    def __init__(self, prm, workers=None, chain="sgbm", calib=None):
        import multiprocessing as mp
        self.prm = dict(prm)
        self.chain, self.calib = ("sgbm" if chain == "stream" else chain), calib
        self.use_cv2 = have_cv2()
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.workers = max(1, min(workers or ncpu, ncpu, 32))
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None
        self.kind = "reference" if self.use_cv2 else "port"

    def run(self, L, R):
        """Computes all pairs of (L, R) and returns (seconds, list of disparities)."""
        jobs = [(L[i], R[i], self.prm, self.use_cv2, self.chain, self.calib) for i in range(len(L))]
        t0 = time.perf_counter()
        out = self.pool.map(_one, jobs) if self.pool else [_one(j) for j in jobs]
        return time.perf_counter() - t0, out

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()

    def describe(self):
        if self.use_cv2:
            import cv2
            what = {"bm": "StereoBM.compute", "pipeline": "remap + GaussianBlur + cvtColor + filter2D + StereoSGBM.compute + filterSpeckles + morphologyEx + "
                    "reprojectImageTo3D + depth (src/lib.cpp:53-85)"}.get(self.chain, "StereoSGBM.compute (the OpenCV code src/lib.cpp:262 calls)")
            return f"cv2 {cv2.__version__} {what}, {self.workers} worker processes x 1 pair"
        return f"plain-C oracle port, {self.workers} worker processes x 1 pair"

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


def _one(args):
    l, r, prm, use_cv2 = args
    if use_cv2:
        import cv2
        cv2.setNumThreads(1)
        s = cv2.StereoSGBM_create(**prm)
        return s.compute(l, r)
    from oracle import sv_oracle
    return sv_oracle.sgbm(l, r, **prm)


class CpuSgbm:
    """A pool of worker processes, one pair each (cv2's MODE_SGBM / MODE_HH are single-threaded inside OpenCV,
    SURVEY.md [B#12], so whole-host throughput needs independent processes)."""

    def __init__(self, prm, workers=None):
        import multiprocessing as mp
        self.prm = dict(prm)
        self.use_cv2 = have_cv2()
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.workers = max(1, min(workers or ncpu, ncpu, 32))
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None
        self.kind = "reference" if self.use_cv2 else "port"

    def run(self, L, R):
        """Computes all pairs of (L, R) and returns (seconds, list of disparities)."""
        jobs = [(L[i], R[i], self.prm, self.use_cv2) for i in range(len(L))]
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
            return f"cv2 {cv2.__version__} StereoSGBM.compute (the OpenCV code src/lib.cpp:262 calls), {self.workers} worker processes x 1 pair"
        return f"plain-C oracle port, {self.workers} worker processes x 1 pair"

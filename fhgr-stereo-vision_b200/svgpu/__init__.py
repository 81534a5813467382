"""svgpu — Python host side of the B200 dense-disparity path.

`Context` wraps the C ABI (include/svgpu.h), one method per entry point; the mirror of the reference's
public C++ API (include/stereo-vision-lib/lib.hpp) is the C++23 facade under ../cpp.  Importing this
package does not load the CUDA library; the first `Context()` does, and raises if it is missing (no
CPU fallback).
"""
from . import _abi
from ._abi import (MODE_HH, MODE_HH4, MODE_SGBM, MODE_SGBM_3WAY, STAGE_DEPTH, STAGE_POSTFILTER, STAGE_PREFILTER,
                   STAGE_REMAP, STAGE_REPROJECT, STAGE_SGBM, SV_LEFT, SV_RIGHT, SvgpuError)
from .context import Context, stereo_rectify

# the SGBM setter block of the reference, src/lib.cpp:249-259
REFERENCE_SGBM = dict(minDisparity=0, numDisparities=224, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1,
                      preFilterCap=32, uniquenessRatio=15, speckleWindowSize=0, speckleRange=2, mode=MODE_HH4)

__all__ = ["Context", "stereo_rectify", "SvgpuError", "REFERENCE_SGBM", "MODE_SGBM", "MODE_HH", "MODE_HH4", "MODE_SGBM_3WAY",
           "STAGE_REMAP", "STAGE_PREFILTER", "STAGE_SGBM", "STAGE_POSTFILTER", "STAGE_REPROJECT", "STAGE_DEPTH",
           "SV_LEFT", "SV_RIGHT", "_abi"]

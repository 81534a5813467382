"""ctypes binding of include/svgpu.h (the C ABI of libsvgpu.so).

This is the only place the Python host side touches native code.  It fails loudly when
the CUDA library is missing or cannot be loaded: there is no CPU fallback anywhere in
this package (the CPU oracle under /oracle is test infrastructure and is never imported
from here).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVGPU_LIB: developer knob, path of another build of the same library (kernel tuning variants)
LIB_PATH = os.environ.get("SVGPU_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "libsvgpu.so")

SV_OK, SV_INVALID_IMAGE, SV_INVALID_SETTINGS, SV_CUDA_ERROR, SV_OOM, SV_NOT_CONFIGURED = 0, -1, -2, -3, -4, -5
SV_MEM_HOST, SV_MEM_DEVICE = 0, 1
SV_LEFT, SV_RIGHT = 0, 1
MODE_SGBM, MODE_HH, MODE_SGBM_3WAY, MODE_HH4 = 0, 1, 2, 3
STAGE_REMAP, STAGE_PREFILTER, STAGE_SGBM, STAGE_POSTFILTER, STAGE_REPROJECT, STAGE_DEPTH = 1, 2, 4, 8, 16, 32

# every symbol include/svgpu.h declares; tests check that the library exports all of them
EXPORTS = (
    "svgpu_create", "svgpu_destroy", "svgpu_set_stream", "svgpu_use_own_stream", "svgpu_reserve", "svgpu_sync", "svgpu_last_error",
    "svgpu_version", "svgpu_set_rectify_maps", "svgpu_set_Q", "svgpu_set_depth_consts", "svgpu_set_sgbm",
    "svgpu_set_bm", "svgpu_remap", "svgpu_sgbm_compute", "svgpu_sgbm_compute_batch", "svgpu_sgbm_enqueue_batch", "svgpu_bm_compute",
    "svgpu_median3", "svgpu_filter_speckles", "svgpu_morph_close5", "svgpu_reproject", "svgpu_depth",
    "svgpu_prefilter_chain", "svgpu_resize", "svgpu_pipeline", "svgpu_debug_copy_volume", "svgpu_launch_count",
    "svgpu_enable_stage_timing", "svgpu_last_stage_ms", "svgpu_debug_keep_volume",
    "svgpu_init_rectify_maps", "svgpu_get_rectify_maps", "svgpu_stereo_rectify",
    "svgpu_pipeline_batch", "svgpu_pipeline_enqueue_batch",
)


class SgbmParams(C.Structure):
    """svgpu_sgbm_params — one field per cv::StereoSGBM setter at src/lib.cpp:249-259."""
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]


class BmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "preFilterCap", "textureThreshold",
        "uniquenessRatio", "speckleWindowSize", "speckleRange", "disp12MaxDiff")]


class PipelineIO(C.Structure):
    _fields_ = [
        ("left", C.c_void_p), ("right", C.c_void_p), ("left_stride", C.c_size_t), ("right_stride", C.c_size_t),
        ("channels", C.c_int), ("left_rect", C.c_void_p), ("right_rect", C.c_void_p), ("rect_stride", C.c_size_t),
        ("disp", C.c_void_p), ("disp_stride", C.c_size_t), ("xyz", C.c_void_p), ("xyz_stride", C.c_size_t),
        ("depth", C.c_void_p), ("depth_stride", C.c_size_t), ("mem", C.c_int)]


class PipelineBatchIO(C.Structure):
    _fields_ = [
        ("left", C.c_void_p), ("right", C.c_void_p), ("channels", C.c_int), ("left_rect", C.c_void_p), ("right_rect", C.c_void_p),
        ("disp", C.c_void_p), ("xyz", C.c_void_p), ("depth", C.c_void_p), ("mem", C.c_int)]


class SvgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"svgpu error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libsvgpu.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    sig = {
        "svgpu_create": (i, [i, i, i, C.POINTER(vp)]),
        "svgpu_destroy": (i, [vp]),
        "svgpu_set_stream": (i, [vp, vp]),
        "svgpu_use_own_stream": (i, [vp]),
        "svgpu_reserve": (i, [vp, i]),
        "svgpu_sync": (i, [vp]),
        "svgpu_last_error": (C.c_char_p, [vp]),
        "svgpu_version": (C.c_char_p, []),
        "svgpu_set_rectify_maps": (i, [vp, i, vp, vp]),
        "svgpu_init_rectify_maps": (i, [vp, i, vp, vp, vp, vp]),
        "svgpu_get_rectify_maps": (i, [vp, i, vp, vp]),
        "svgpu_stereo_rectify": (i, [vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp]),
        "svgpu_set_Q": (i, [vp, vp]),
        "svgpu_set_depth_consts": (i, [vp, C.c_float, C.c_float]),
        "svgpu_set_sgbm": (i, [vp, C.POINTER(SgbmParams)]),
        "svgpu_set_bm": (i, [vp, C.POINTER(BmParams)]),
        "svgpu_remap": (i, [vp, i, vp, i, i, sz, i, vp, sz, i]),
        "svgpu_sgbm_compute": (i, [vp, vp, sz, vp, sz, vp, sz, i]),
        "svgpu_sgbm_compute_batch": (i, [vp, i, vp, vp, vp, i]),
        "svgpu_sgbm_enqueue_batch": (i, [vp, i, vp, vp, vp]),
        "svgpu_bm_compute": (i, [vp, vp, sz, vp, sz, vp, sz, i]),
        "svgpu_median3": (i, [vp, vp, sz, vp, sz, i]),
        "svgpu_filter_speckles": (i, [vp, vp, sz, i, i, i, i]),
        "svgpu_morph_close5": (i, [vp, vp, sz, vp, sz, i]),
        "svgpu_reproject": (i, [vp, vp, sz, vp, sz, i, i]),
        "svgpu_depth": (i, [vp, vp, sz, vp, sz, i]),
        "svgpu_prefilter_chain": (i, [vp, vp, sz, vp, sz, i]),
        "svgpu_resize": (i, [vp, vp, i, i, sz, i, vp, sz, i]),
        "svgpu_pipeline": (i, [vp, C.POINTER(PipelineIO), C.c_uint]),
        "svgpu_pipeline_batch": (i, [vp, i, C.POINTER(PipelineBatchIO), C.c_uint]),
        "svgpu_pipeline_enqueue_batch": (i, [vp, i, C.POINTER(PipelineBatchIO), C.c_uint]),
        "svgpu_debug_copy_volume": (i, [vp, i, vp, sz]),
        "svgpu_launch_count": (C.c_longlong, [vp]),
        "svgpu_enable_stage_timing": (i, [vp, i]),
        "svgpu_last_stage_ms": (i, [vp, C.POINTER(C.c_float)]),
        "svgpu_debug_keep_volume": (i, [vp, i]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _lib = lib
    return lib

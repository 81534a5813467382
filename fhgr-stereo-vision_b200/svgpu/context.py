"""`Context`: a thin, typed wrapper of one `svgpu_ctx*`.

Host images are numpy arrays (copied through the library's staging buffers), device
images are torch CUDA tensors (passed by `data_ptr()`; torch is only the allocator and
stream provider here).  Every method maps 1:1 onto a C-ABI entry point of
include/svgpu.h, which in turn replaces one OpenCV call of the reference's
src/lib.cpp (citations in the header).
"""
import ctypes as C

import numpy as np

from . import _abi


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


class Context:
    def __init__(self, width, height, device=0):
        self._lib = _abi.load()
        self.width, self.height, self.device = int(width), int(height), int(device)
        h = C.c_void_p()
        rc = self._lib.svgpu_create(self.device, self.width, self.height, C.byref(h))
        if rc != _abi.SV_OK:
            raise _abi.SvgpuError(rc, "svgpu_create failed (no CUDA device of compute capability 10.x, or out of memory)")
        self._h = h
        self._keep = []
        self._user_stream = False

    # -- plumbing ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.svgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != _abi.SV_OK:
            raise _abi.SvgpuError(rc, self._lib.svgpu_last_error(self._h).decode())

    @staticmethod
    def _ptr(a):
        return C.c_void_p(a.data_ptr()) if _is_tensor(a) else a.ctypes.data_as(C.c_void_p)

    def _mem(self, *arrays):
        dev = [_is_tensor(a) for a in arrays if a is not None]
        if any(dev) and not all(dev):
            raise ValueError("mix of host (numpy) and device (torch) buffers")
        if dev and dev[0]:
            # Device tensors live on torch's streams: unless the caller attached a stream of their own, the call runs on torch's
            # current stream, so that it is ordered after the producers of its inputs and before the consumers of its outputs
            # (and torch's caching allocator sees the use).
            if not self._user_stream:
                import torch
                self._lib.svgpu_set_stream(self._h, C.c_void_p(int(torch.cuda.current_stream(self.device).cuda_stream)))
            return _abi.SV_MEM_DEVICE
        return _abi.SV_MEM_HOST

    @staticmethod
    def _stride(a):
        return int(a.stride(0) * a.element_size()) if _is_tensor(a) else int(a.strides[0])

    def _out_like(self, ref, shape, dtype):
        if _is_tensor(ref):
            import torch
            tdt = {np.int16: torch.int16, np.float32: torch.float32, np.uint8: torch.uint8}[dtype]
            return torch.empty(shape, dtype=tdt, device=ref.device)
        return np.empty(shape, dtype)

    def _img(self, a, dtype=np.uint8):
        if _is_tensor(a):
            if not a.is_cuda or a.stride(-1) != 1 or (a.dim() == 3 and a.stride(1) != a.shape[2]):
                raise ValueError("device images must be CUDA tensors with packed pixels")
            return a
        a = np.asarray(a)
        if a.dtype != dtype:
            raise TypeError(f"expected {dtype}, got {a.dtype}")
        if a.size == 0:
            raise _abi.SvgpuError(_abi.SV_INVALID_IMAGE, "empty image")
        if a.strides[-1] != a.itemsize or (a.ndim == 3 and a.strides[1] != a.shape[2] * a.itemsize):
            a = np.ascontiguousarray(a)
        return a

    def set_stream(self, cuda_stream):
        """Run on the caller's cudaStream_t handle (an int; 0 is the CUDA default stream)."""
        self._check(self._lib.svgpu_set_stream(self._h, C.c_void_p(int(cuda_stream))))
        self._user_stream = True

    def use_own_stream(self):
        """Back to the context's private stream for host-memory calls; calls on device tensors follow torch's current stream."""
        self._check(self._lib.svgpu_use_own_stream(self._h))
        self._user_stream = False

    def reserve(self, max_batch):
        self._check(self._lib.svgpu_reserve(self._h, int(max_batch)))

    def sync(self):
        self._check(self._lib.svgpu_sync(self._h))

    @property
    def launch_count(self):
        return int(self._lib.svgpu_launch_count(self._h))

    # -- one-time state ------------------------------------------------------------------------
    def set_rectify_maps(self, side, map1, map2):
        """cv::initUndistortRectifyMap outputs (CV_16SC2 + CV_16UC1) — src/lib.cpp:39-46."""
        m1 = np.ascontiguousarray(map1, np.int16)
        m2 = np.ascontiguousarray(map2, np.uint16)
        if m1.shape != (self.height, self.width, 2) or m2.shape != (self.height, self.width):
            raise ValueError("map shapes must be (H,W,2) int16 and (H,W) uint16")
        self._check(self._lib.svgpu_set_rectify_maps(self._h, int(side), self._ptr(m1), self._ptr(m2)))

    def init_rectify_maps(self, side, K, D, R, P):
        """cv::initUndistortRectifyMap(K, D, R, P, size, CV_16SC2) computed on the device — src/lib.cpp:39-46."""
        a = [np.ascontiguousarray(v, np.float64).reshape(-1) for v in (K, D, R, P)]
        if [v.size for v in a] != [9, 5, 9, 12]:
            raise ValueError("K 3x3, D 5 (k1 k2 p1 p2 k3), R 3x3, P 3x4 expected")
        self._check(self._lib.svgpu_init_rectify_maps(self._h, int(side), *[self._ptr(v) for v in a]))

    def get_rectify_maps(self, side):
        m1 = np.empty((self.height, self.width, 2), np.int16)
        m2 = np.empty((self.height, self.width), np.uint16)
        self._check(self._lib.svgpu_get_rectify_maps(self._h, int(side), self._ptr(m1), self._ptr(m2)))
        return m1, m2

    def set_Q(self, Q):
        q = np.ascontiguousarray(Q, np.float64).reshape(16)
        self._check(self._lib.svgpu_set_Q(self._h, self._ptr(q)))

    def set_depth_consts(self, baseline, focal):
        self._check(self._lib.svgpu_set_depth_consts(self._h, float(baseline), float(focal)))

    def set_sgbm(self, **kw):
        """Fields = the cv::StereoSGBM setters at src/lib.cpp:249-259."""
        p = _abi.SgbmParams(**kw)
        self._check(self._lib.svgpu_set_sgbm(self._h, C.byref(p)))
        self._sgbm = dict(kw)

    def set_bm(self, **kw):
        d = dict(minDisparity=0, numDisparities=64, blockSize=21, preFilterCap=31, textureThreshold=10,
                 uniquenessRatio=15, speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1)
        d.update(kw)
        p = _abi.BmParams(**d)
        self._check(self._lib.svgpu_set_bm(self._h, C.byref(p)))

    def debug_keep_volume(self, on=True):
        self._check(self._lib.svgpu_debug_keep_volume(self._h, int(bool(on))))

    def enable_stage_timing(self, on=True):
        self._check(self._lib.svgpu_enable_stage_timing(self._h, int(bool(on))))

    def last_stage_ms(self):
        ms = (C.c_float * 4)()
        self._check(self._lib.svgpu_last_stage_ms(self._h, ms))
        return dict(cost=ms[0], aggregate=ms[1], select=ms[2], total=ms[3])

    # -- stages --------------------------------------------------------------------------------
    def remap(self, side, src):
        """cv::remap(INTER_LINEAR) — src/lib.cpp:89-92."""
        src = self._img(src)
        cn = 1 if src.ndim == 2 else int(src.shape[2])
        shape = (self.height, self.width) if src.ndim == 2 else (self.height, self.width, cn)
        dst = self._out_like(src, shape, np.uint8)
        self._check(self._lib.svgpu_remap(self._h, int(side), self._ptr(src), int(src.shape[1]), int(src.shape[0]),
                                          self._stride(src), cn, self._ptr(dst), self._stride(dst), self._mem(src)))
        return dst

    def _pair(self, left, right):
        left, right = self._img(left), self._img(right)
        if tuple(left.shape) != (self.height, self.width) or tuple(right.shape) != (self.height, self.width):
            raise ValueError("images must be 8UC1 of the context size")
        return left, right

    def sgbm(self, left, right):
        """cv::StereoSGBM::compute — src/lib.cpp:262.  8UC1 pairs (what the reference passes) or 8UC3 pairs (OpenCV then sums
        the matching cost over the three channels; goes through svgpu_pipeline with SV_STAGE_SGBM alone)."""
        if len(left.shape) == 3 and int(left.shape[2]) == 3:
            return self.pipeline(left, right, _abi.STAGE_SGBM)["disp"]
        left, right = self._pair(left, right)
        disp = self._out_like(left, (self.height, self.width), np.int16)
        self._check(self._lib.svgpu_sgbm_compute(self._h, self._ptr(left), self._stride(left), self._ptr(right),
                                                 self._stride(right), self._ptr(disp), self._stride(disp),
                                                 self._mem(left, right)))
        return disp

    def sgbm_batch(self, left, right, out=None):
        """n tightly packed pairs [n][H][W]."""
        if _is_tensor(left):
            if not (left.is_contiguous() and right.is_contiguous()):
                raise ValueError("batch tensors must be contiguous")
        else:
            left, right = np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(right, np.uint8)
        n = int(left.shape[0])
        if tuple(left.shape[1:]) != (self.height, self.width) or tuple(right.shape) != tuple(left.shape):
            raise ValueError("batch must be [n][H][W] 8UC1")
        disp = out if out is not None else self._out_like(left, (n, self.height, self.width), np.int16)
        if out is not None:
            if _is_tensor(out) != _is_tensor(left) or tuple(out.shape) != (n, self.height, self.width):
                raise ValueError("out must be [n][H][W] int16 in the same memory as the inputs")
            if (str(out.dtype) != "torch.int16" or not out.is_contiguous()) if _is_tensor(out) else (out.dtype != np.int16 or not out.flags["C_CONTIGUOUS"]):
                raise ValueError("out must be a contiguous int16 array")
        self._check(self._lib.svgpu_sgbm_compute_batch(self._h, n, self._ptr(left), self._ptr(right), self._ptr(disp),
                                                       self._mem(left, right, disp)))
        return disp

    def sgbm_enqueue(self, left, right, out):
        """Streaming form of sgbm_batch for host (ideally pinned) arrays: returns at once; copy-in, kernels and copy-out of
        consecutive calls overlap.  Keep the arrays alive and untouched until sync()."""
        if _is_tensor(left) or _is_tensor(right) or _is_tensor(out):
            raise ValueError("sgbm_enqueue takes host arrays")
        for a, dt in ((left, np.uint8), (right, np.uint8), (out, np.int16)):
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("sgbm_enqueue needs contiguous uint8 inputs and an int16 output")
        n = int(left.shape[0])
        if tuple(left.shape[1:]) != (self.height, self.width) or tuple(right.shape) != tuple(left.shape) or tuple(out.shape) != tuple(left.shape):
            raise ValueError("batch must be [n][H][W]")
        self._keep = getattr(self, "_keep", [])[-8:] + [(left, right, out)]
        self._check(self._lib.svgpu_sgbm_enqueue_batch(self._h, n, self._ptr(left), self._ptr(right), self._ptr(out)))

    def bm(self, left, right, out=None):
        """cv::StereoBM::compute — BASELINE config 2."""
        left, right = self._pair(left, right)
        disp = out if out is not None else self._out_like(left, (self.height, self.width), np.int16)
        if out is not None and (_is_tensor(out) != _is_tensor(left) or tuple(out.shape) != (self.height, self.width)):
            raise ValueError("out must be [H][W] int16 in the same memory as the inputs")
        self._check(self._lib.svgpu_bm_compute(self._h, self._ptr(left), self._stride(left), self._ptr(right),
                                               self._stride(right), self._ptr(disp), self._stride(disp),
                                               self._mem(left, right)))
        return disp

    def _s16(self, a):
        a = self._img(a, np.int16)
        if tuple(a.shape) != (self.height, self.width):
            raise ValueError("disparity must be CV_16S of the context size")
        return a

    def median3(self, disp):
        disp = self._s16(disp)
        out = self._out_like(disp, (self.height, self.width), np.int16)
        self._check(self._lib.svgpu_median3(self._h, self._ptr(disp), self._stride(disp), self._ptr(out),
                                            self._stride(out), self._mem(disp)))
        return out

    def filter_speckles(self, disp, new_val, max_size, max_diff):
        """cv::filterSpeckles — src/lib.cpp:265 (returns a filtered copy)."""
        disp = self._s16(disp)
        out = disp.clone() if _is_tensor(disp) else disp.copy()
        self._check(self._lib.svgpu_filter_speckles(self._h, self._ptr(out), self._stride(out), int(new_val),
                                                    int(max_size), int(max_diff), self._mem(out)))
        return out

    def morph_close5(self, disp):
        """cv::morphologyEx(MORPH_CLOSE, 5x5) — src/lib.cpp:267-268."""
        disp = self._s16(disp)
        out = self._out_like(disp, (self.height, self.width), np.int16)
        self._check(self._lib.svgpu_morph_close5(self._h, self._ptr(disp), self._stride(disp), self._ptr(out),
                                                 self._stride(out), self._mem(disp)))
        return out

    def reproject(self, disp, handle_missing=True):
        """cv::reprojectImageTo3D — src/lib.cpp:154."""
        disp = self._s16(disp)
        out = self._out_like(disp, (self.height, self.width, 3), np.float32)
        self._check(self._lib.svgpu_reproject(self._h, self._ptr(disp), self._stride(disp), self._ptr(out),
                                              self._stride(out), int(bool(handle_missing)), self._mem(disp)))
        return out

    def depth(self, disp):
        """StereoVision::CalculateDepthMapSimple — src/lib.cpp:279-294."""
        disp = self._s16(disp)
        out = self._out_like(disp, (self.height, self.width), np.float32)
        self._check(self._lib.svgpu_depth(self._h, self._ptr(disp), self._stride(disp), self._ptr(out),
                                          self._stride(out), self._mem(disp)))
        return out

    def prefilter_chain(self, bgr):
        """GaussianBlur 11x11 + BGR2GRAY + sharpen — src/lib.cpp:231-246."""
        bgr = self._img(bgr)
        if tuple(bgr.shape) != (self.height, self.width, 3):
            raise ValueError("expected 8UC3 of the context size")
        out = self._out_like(bgr, (self.height, self.width), np.uint8)
        self._check(self._lib.svgpu_prefilter_chain(self._h, self._ptr(bgr), self._stride(bgr), self._ptr(out),
                                                    self._stride(out), self._mem(bgr)))
        return out

    def resize(self, src):
        """cv::resize(INTER_LINEAR) to the context size — src/lib.cpp:96-104."""
        src = self._img(src)
        cn = 1 if src.ndim == 2 else int(src.shape[2])
        shape = (self.height, self.width) if src.ndim == 2 else (self.height, self.width, cn)
        out = self._out_like(src, shape, np.uint8)
        self._check(self._lib.svgpu_resize(self._h, self._ptr(src), int(src.shape[1]), int(src.shape[0]),
                                           self._stride(src), cn, self._ptr(out), self._stride(out), self._mem(src)))
        return out

    def pipeline(self, left, right, stages, want_rect=False, want_xyz=False, want_depth=False):
        """StereoVision::AnalyzeAndAnnotateImage (kSGBM) — src/lib.cpp:53-85, any subset of stages."""
        left, right = self._img(left), self._img(right)
        cn = 1 if left.ndim == 2 else int(left.shape[2])
        H, W = self.height, self.width
        io = _abi.PipelineIO()
        io.left, io.right = self._ptr(left), self._ptr(right)
        io.left_stride, io.right_stride, io.channels = self._stride(left), self._stride(right), cn
        io.mem = self._mem(left, right)
        out = {}
        if want_rect:
            shape = (H, W) if cn == 1 else (H, W, cn)
            out["left_rect"], out["right_rect"] = self._out_like(left, shape, np.uint8), self._out_like(left, shape, np.uint8)
            io.left_rect, io.right_rect = self._ptr(out["left_rect"]), self._ptr(out["right_rect"])
            io.rect_stride = self._stride(out["left_rect"])
        out["disp"] = self._out_like(left, (H, W), np.int16)
        io.disp, io.disp_stride = self._ptr(out["disp"]), self._stride(out["disp"])
        if want_xyz:
            out["xyz"] = self._out_like(left, (H, W, 3), np.float32)
            io.xyz, io.xyz_stride = self._ptr(out["xyz"]), self._stride(out["xyz"])
        if want_depth:
            out["depth"] = self._out_like(left, (H, W), np.float32)
            io.depth, io.depth_stride = self._ptr(out["depth"]), self._stride(out["depth"])
        self._check(self._lib.svgpu_pipeline(self._h, C.byref(io), int(stages)))
        return out

    def _batch_io(self, left, right, stages, want_rect, want_xyz, want_depth, out, disp_in):
        """Fills a svgpu_pipeline_batch_io for tightly packed [n][H][W](x cn) frames; `out` may carry preallocated result arrays."""
        if _is_tensor(left):
            if not (left.is_contiguous() and right.is_contiguous()):
                raise ValueError("batch tensors must be contiguous")
        else:
            left, right = np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(right, np.uint8)
        H, W = self.height, self.width
        n = int(left.shape[0])
        cn = 1 if len(left.shape) == 3 else int(left.shape[3])
        want = (n, H, W) if cn == 1 else (n, H, W, cn)
        if tuple(left.shape) != want or tuple(right.shape) != want:
            raise ValueError("batch must be [n][H][W] 8UC1 or [n][H][W][3] 8UC3 of the context size")
        out = dict(out or {})
        io = _abi.PipelineBatchIO()
        io.left, io.right, io.channels = self._ptr(left), self._ptr(right), cn
        io.mem = self._mem(left, right)

        def buf(key, shape, dtype):
            a = out.get(key)
            if a is None:
                a = out[key] = self._out_like(left, shape, dtype)
            else:
                ok = tuple(a.shape) == tuple(shape) and (a.is_contiguous() if _is_tensor(a) else (a.flags["C_CONTIGUOUS"] and a.dtype == dtype))
                if not ok or _is_tensor(a) != _is_tensor(left):
                    raise ValueError(f"output '{key}' must be a contiguous {np.dtype(dtype).name} array of shape {tuple(shape)} in the same memory as the inputs")
            return self._ptr(a)

        if want_rect:
            io.left_rect, io.right_rect = buf("left_rect", want, np.uint8), buf("right_rect", want, np.uint8)
        if disp_in is not None:
            out["disp"] = disp_in if _is_tensor(disp_in) else np.ascontiguousarray(disp_in, np.int16)
        io.disp = buf("disp", (n, H, W), np.int16)
        if want_xyz:
            io.xyz = buf("xyz", (n, H, W, 3), np.float32)
        if want_depth:
            io.depth = buf("depth", (n, H, W), np.float32)
        return n, io, out, (left, right)

    def pipeline_batch(self, left, right, stages, want_rect=False, want_xyz=False, want_depth=False, out=None, disp=None):
        """StereoVision::AnalyzeAndAnnotateImage (src/lib.cpp:53-85) for n frames per call (svgpu_pipeline_batch): what the folder
        and livecam loops (static-image-main.cpp:107-127, livecam-main.cpp:70-106) do frame by frame."""
        n, io, out, _ = self._batch_io(left, right, stages, want_rect, want_xyz, want_depth, out, disp)
        self._check(self._lib.svgpu_pipeline_batch(self._h, n, C.byref(io), int(stages)))
        return out

    def pipeline_enqueue(self, left, right, stages, out, want_rect=False, want_xyz=False, want_depth=False):
        """Streaming form (svgpu_pipeline_enqueue_batch): host arrays only, returns at once; `out` must hold the preallocated
        result arrays ('disp' and, as requested, 'left_rect', 'right_rect', 'xyz', 'depth').  Keep everything alive and
        untouched until sync()."""
        if _is_tensor(left) or _is_tensor(right):
            raise ValueError("pipeline_enqueue takes host arrays")
        if not isinstance(out, dict) or "disp" not in out:
            raise ValueError("pipeline_enqueue needs preallocated outputs")
        n, io, out2, ins = self._batch_io(left, right, stages, want_rect, want_xyz, want_depth, out, None)
        self._keep = getattr(self, "_keep", [])[-8:] + [(ins, out2)]
        self._check(self._lib.svgpu_pipeline_enqueue_batch(self._h, n, C.byref(io), int(stages)))
        return out2

    def debug_volume(self, which, shape):
        out = np.zeros(shape, np.int16)
        self._check(self._lib.svgpu_debug_copy_volume(self._h, int(which), self._ptr(out), out.nbytes))
        return out


def stereo_rectify(K1, D1, K2, D2, size, R, T):
    """cv::stereoRectify(K1, D1, K2, D2, size, R, T, flags=CALIB_ZERO_DISPARITY, alpha=0) — src/lib.cpp:32-37.
    Host arithmetic inside libsvgpu.so (needs no GPU).  size = (width, height).  Returns R1, R2, P1, P2, Q."""
    lib = _abi.load()
    a = [np.ascontiguousarray(v, np.float64).reshape(-1) for v in (K1, D1, K2, D2, R, T)]
    if [v.size for v in a] != [9, 5, 9, 5, 9, 3]:
        raise ValueError("K 3x3, D 5 (k1 k2 p1 p2 k3), R 3x3, T 3 expected")
    out = [np.zeros(n, np.float64) for n in (9, 9, 12, 12, 16)]
    p = lambda v: v.ctypes.data_as(C.c_void_p)
    rc = lib.svgpu_stereo_rectify(p(a[0]), p(a[1]), p(a[2]), p(a[3]), int(size[0]), int(size[1]), p(a[4]), p(a[5]), *[p(v) for v in out])
    if rc != _abi.SV_OK:
        raise _abi.SvgpuError(rc, "svgpu_stereo_rectify: degenerate calibration")
    return out[0].reshape(3, 3), out[1].reshape(3, 3), out[2].reshape(3, 4), out[3].reshape(3, 4), out[4].reshape(4, 4)

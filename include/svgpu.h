/*
 * svgpu.h — C ABI of the B200 (sm_100a) dense-disparity path.
 *
 * The reference (rursprung/fhgr-stereo-vision) has no FFI or plugin interface: its
 * hot path is the set of OpenCV calls made by the private methods of
 * stereo_vision::StereoVision (include/stereo-vision-lib/lib.hpp:133-193, bodies in
 * src/lib.cpp).  Each entry point below replaces one of those call sites; the
 * citation on each declaration is the reference line it stands in for.  A facade
 * with the reference's public API (fhgr-stereo-vision_b200/cpp/) and a ctypes
 * binding (fhgr-stereo-vision_b200/svgpu/) sit on top of exactly these symbols.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; no exceptions cross the ABI.
 *  - every call returns an int status: SV_OK (0) or a negative SV_* code;
 *    svgpu_last_error(ctx) gives the text of the last failure on that context.
 *  - images are row-major with an explicit byte stride (cv::Mat::step), 8UC1 / 8UC3
 *    in, CV_16S disparity (value = 16*d), 32FC3 XYZ and 32FC1 depth out — the same
 *    buffers src/lib.cpp hands OpenCV today.
 *  - `mem` says where the caller's image buffers live: SV_MEM_HOST (pageable or
 *    pinned host memory; the call copies in/out on the context's stream and
 *    returns when the result is in the host buffer) or SV_MEM_DEVICE (device
 *    pointers on the context's device; the call only enqueues work — use
 *    svgpu_sync or the caller's own stream to wait).
 *  - the caller owns every buffer it passes; the context owns maps, scratch volumes
 *    and (unless one is attached) its stream.  One context per GPU and thread;
 *    calls on one context are serialised by an internal mutex.
 *  - there is no CPU fallback: without a usable CUDA device svgpu_create fails.
 */
#ifndef SVGPU_H
#define SVGPU_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SVGPU_API __attribute__((visibility("default")))
#else
#define SVGPU_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svgpu_ctx svgpu_ctx;

enum {
  SV_OK = 0,
  SV_INVALID_IMAGE = -1,    /* -> AnalysisError::kInvalidImage    (lib.hpp:30, src/lib.cpp:54-56) */
  SV_INVALID_SETTINGS = -2, /* -> AnalysisError::kInvalidSettings (lib.hpp:32, src/lib.cpp:71-73) */
  SV_CUDA_ERROR = -3,
  SV_OOM = -4,
  SV_NOT_CONFIGURED = -5 /* a stage was called before its maps / Q / parameters were set */
};

enum { SV_MEM_HOST = 0, SV_MEM_DEVICE = 1 };
enum { SV_LEFT = 0, SV_RIGHT = 1 };
/* cv::StereoSGBM mode constants (src/lib.cpp:259 uses MODE_HH4) */
enum { SV_MODE_SGBM = 0, SV_MODE_HH = 1, SV_MODE_SGBM_3WAY = 2, SV_MODE_HH4 = 3 };

/* The setter block of src/lib.cpp:249-259, one field per cv::StereoSGBM setter.
 * Values are normalised the way OpenCV does (P1<=0 -> 2, uniquenessRatio<0 -> 10, ...). */
typedef struct svgpu_sgbm_params {
  int minDisparity;      /* :250 */
  int numDisparities;    /* :251, multiple of 16, <= 512 */
  int blockSize;         /* :249, odd, 1..15 (as far as blockSize^2 * (2 * ftzero + 63) + P2 fits 16 bits, OpenCV's own cost type) */
  int P1;                /* :257 */
  int P2;                /* :258 */
  int disp12MaxDiff;     /* :256 */
  int preFilterCap;      /* :255, <= 63 */
  int uniquenessRatio;   /* :254 */
  int speckleWindowSize; /* :253 */
  int speckleRange;      /* :252 */
  int mode;              /* :259 */
} svgpu_sgbm_params;

/* cv::StereoBM::create(numDisparities, blockSize) + setters (BASELINE config 2; the
 * reference itself never calls StereoBM). XSOBEL pre-filter, minDisparity 0. */
typedef struct svgpu_bm_params {
  int minDisparity, numDisparities, blockSize, preFilterCap, textureThreshold;
  int uniquenessRatio, speckleWindowSize, speckleRange, disp12MaxDiff;
} svgpu_bm_params;

/* ---- context --------------------------------------------------------------------- */
/* StereoVision::StereoVision (src/lib.cpp:29): one context per calibrated camera pair
 * of image_size width x height on CUDA device `device`. */
SVGPU_API int svgpu_create(int device, int width, int height, svgpu_ctx** out);
SVGPU_API int svgpu_destroy(svgpu_ctx* ctx);
/* run all further work of this context on the caller's cudaStream_t (e.g. torch's current stream).
 * The handle is taken literally: NULL is the CUDA default stream. svgpu_use_own_stream goes back to
 * the context's private non-blocking stream (the state after svgpu_create). */
SVGPU_API int svgpu_set_stream(svgpu_ctx* ctx, void* cuda_stream);
SVGPU_API int svgpu_use_own_stream(svgpu_ctx* ctx);
/* how many pairs the *_batch calls may keep in flight (scratch = C and S volumes per pair) */
SVGPU_API int svgpu_reserve(svgpu_ctx* ctx, int max_batch);
SVGPU_API int svgpu_sync(svgpu_ctx* ctx);
SVGPU_API const char* svgpu_last_error(const svgpu_ctx* ctx);
/* "svgpu <ver> sm_100a" */
SVGPU_API const char* svgpu_version(void);

/* ---- one-time state (src/lib.cpp:32-46) --------------------------------------------- */
/* the two cv::initUndistortRectifyMap(..., CV_16SC2) outputs per side, host pointers,
 * tightly packed: map1 = H*W*2 int16 (x,y), map2 = H*W uint16 (fy*32+fx). src/lib.cpp:39-46 */
SVGPU_API int svgpu_set_rectify_maps(svgpu_ctx* ctx, int side, const int16_t* map1, const uint16_t* map2);
/* cv::initUndistortRectifyMap(K, D, R, P, image_size, CV_16SC2, map1, map2) of one side, src/lib.cpp:39-46, computed on the
 * device straight into the context's map storage (instead of svgpu_set_rectify_maps).  Row-major f64: K 3x3, D = k1 k2 p1
 * p2 k3, R 3x3 (R1 / R2 of stereoRectify), P 3x4 (P1 / P2). */
SVGPU_API int svgpu_init_rectify_maps(svgpu_ctx* ctx, int side, const double K[9], const double D[5], const double R[9],
                                      const double P[12]);
/* copies the maps of one side back to the host (same layout as svgpu_set_rectify_maps) */
SVGPU_API int svgpu_get_rectify_maps(svgpu_ctx* ctx, int side, int16_t* map1, uint16_t* map2);
/* cv::stereoRectify(K1, D1, K2, D2, image_size, R, T, R1, R2, P1, P2, Q, CALIB_ZERO_DISPARITY, alpha = 0), src/lib.cpp:32-37.
 * Host arithmetic (one-time, a few hundred f64 operations), needs no context.  Outputs row-major: R1, R2 3x3; P1, P2 3x4;
 * Q 4x4.  SV_INVALID_SETTINGS for a degenerate calibration (zero baseline, singular matrices). */
SVGPU_API int svgpu_stereo_rectify(const double K1[9], const double D1[5], const double K2[9], const double D2[5], int width,
                                   int height, const double R[9], const double T[3], double R1[9], double R2[9], double P1[12],
                                   double P2[12], double Q[16]);
/* Q_ from cv::stereoRectify, row-major 4x4 f64. src/lib.cpp:36 */
SVGPU_API int svgpu_set_Q(svgpu_ctx* ctx, const double Q[16]);
/* |T[0]| and K_left[0,0] as float. src/lib.cpp:283-284 */
SVGPU_API int svgpu_set_depth_consts(svgpu_ctx* ctx, float baseline, float focal_length);
SVGPU_API int svgpu_set_sgbm(svgpu_ctx* ctx, const svgpu_sgbm_params* p); /* src/lib.cpp:248-259 */
SVGPU_API int svgpu_set_bm(svgpu_ctx* ctx, const svgpu_bm_params* p);

/* ---- stages ------------------------------------------------------------------------ */
/* cv::remap(src, dst, map1, map2, INTER_LINEAR) — src/lib.cpp:89-92. src may have any size
 * (the maps address it); dst is width x height of the context; channels 1 or 3. */
SVGPU_API int svgpu_remap(svgpu_ctx* ctx, int side, const uint8_t* src, int src_width, int src_height, size_t src_stride,
                int channels, uint8_t* dst, size_t dst_stride, int mem);
/* stereo->compute(left_gray, right_gray, disparity) — src/lib.cpp:262. 8UC1 in, CV_16S out. */
SVGPU_API int svgpu_sgbm_compute(svgpu_ctx* ctx, const uint8_t* left, size_t left_stride, const uint8_t* right,
                       size_t right_stride, int16_t* disp, size_t disp_stride, int mem);
/* n independent pairs, tightly packed ([n][H][W]); device or host memory. This is the
 * throughput entry point the folder / livecam loops (static-image-main.cpp:107-127,
 * livecam-main.cpp:70-106) map onto. */
SVGPU_API int svgpu_sgbm_compute_batch(svgpu_ctx* ctx, int n, const uint8_t* left, const uint8_t* right, int16_t* disp, int mem);
/* Streaming form of the batch call for the livecam / folder loops (livecam-main.cpp:70-106,
 * static-image-main.cpp:107-127): enqueues copy-in, kernels and copy-out of n <= svgpu_reserve() pairs in HOST memory
 * (pinned for real overlap) and returns; consecutive enqueues overlap copy-in(k+1), kernels(k) and copy-out(k-1).
 * The host buffers of a call must stay untouched until svgpu_sync(ctx) returns (which also reports errors). */
SVGPU_API int svgpu_sgbm_enqueue_batch(svgpu_ctx* ctx, int n, const uint8_t* left, const uint8_t* right, int16_t* disp);
/* cv::StereoBM::compute (BASELINE config 2) */
SVGPU_API int svgpu_bm_compute(svgpu_ctx* ctx, const uint8_t* left, size_t left_stride, const uint8_t* right,
                     size_t right_stride, int16_t* disp, size_t disp_stride, int mem);
/* cv::medianBlur(CV_16S, 3) — the last stage inside StereoSGBM::compute */
SVGPU_API int svgpu_median3(svgpu_ctx* ctx, const int16_t* src, size_t src_stride, int16_t* dst, size_t dst_stride, int mem);
/* cv::filterSpeckles(disp, newVal, maxSpeckleSize, maxDiff), in place — src/lib.cpp:265 */
SVGPU_API int svgpu_filter_speckles(svgpu_ctx* ctx, int16_t* disp, size_t stride, int new_val, int max_speckle_size,
                          int max_diff, int mem);
/* cv::morphologyEx(disp, MORPH_CLOSE, 5x5 rect) — src/lib.cpp:267-268 */
SVGPU_API int svgpu_morph_close5(svgpu_ctx* ctx, const int16_t* src, size_t src_stride, int16_t* dst, size_t dst_stride, int mem);
/* cv::reprojectImageTo3D(disp, xyz, Q_, handleMissingValues) — src/lib.cpp:154. xyz = 32FC3 */
SVGPU_API int svgpu_reproject(svgpu_ctx* ctx, const int16_t* disp, size_t disp_stride, float* xyz, size_t xyz_stride,
                    int handle_missing, int mem);
/* StereoVision::CalculateDepthMapSimple — src/lib.cpp:279-294. depth = 32FC1 */
SVGPU_API int svgpu_depth(svgpu_ctx* ctx, const int16_t* disp, size_t disp_stride, float* depth, size_t depth_stride, int mem);
/* GaussianBlur 11x11 sigma 1.4 -> BGR2GRAY -> 3x3 sharpen — src/lib.cpp:231-246. 8UC3 in, 8UC1 out */
SVGPU_API int svgpu_prefilter_chain(svgpu_ctx* ctx, const uint8_t* bgr, size_t bgr_stride, uint8_t* gray, size_t gray_stride, int mem);
/* cv::resize(INTER_LINEAR) to the context size — src/lib.cpp:96-104 (down-scaling contract) */
SVGPU_API int svgpu_resize(svgpu_ctx* ctx, const uint8_t* src, int src_width, int src_height, size_t src_stride, int channels,
                 uint8_t* dst, size_t dst_stride, int mem);

/* ---- fused chain ------------------------------------------------------------------- */
enum {
  SV_STAGE_REMAP = 1,      /* src/lib.cpp:87-94: inputs are raw 8UC3 frames, rectified on the fly   */
  SV_STAGE_PREFILTER = 2,  /* src/lib.cpp:231-246: blur + gray + sharpen (else inputs must be 8UC1) */
  SV_STAGE_SGBM = 4,       /* src/lib.cpp:248-262; without SV_STAGE_PREFILTER an 8UC3 pair goes to
                            * StereoSGBM as it is (cost summed over the three channels)             */
  SV_STAGE_POSTFILTER = 8, /* src/lib.cpp:264-268: filterSpeckles(0,150,32) + close 5x5            */
  SV_STAGE_REPROJECT = 16, /* src/lib.cpp:152-156                                                  */
  SV_STAGE_DEPTH = 32      /* src/lib.cpp:279-294                                                  */
};
typedef struct svgpu_pipeline_io {
  const uint8_t* left;  /* width x height, `channels` per pixel */
  const uint8_t* right;
  size_t left_stride, right_stride;
  int channels;         /* 3 (BGR) or 1 */
  uint8_t* left_rect;   /* optional outputs (NULL to skip): rectified images, same channels */
  uint8_t* right_rect;
  size_t rect_stride;
  int16_t* disp;        /* CV_16S */
  size_t disp_stride;
  float* xyz;           /* 32FC3, optional */
  size_t xyz_stride;
  float* depth;         /* 32FC1, optional */
  size_t depth_stride;
  int mem;
} svgpu_pipeline_io;
/* StereoVision::AnalyzeAndAnnotateImage with Algorithm::kSGBM — src/lib.cpp:53-85 */
SVGPU_API int svgpu_pipeline(svgpu_ctx* ctx, const svgpu_pipeline_io* io, unsigned stages);

/* The same chain for n independent pairs per call: what the folder loop (apps/stereo-vision/static-image-main.cpp:107-127)
 * and the livecam loop (apps/stereo-vision/livecam-main.cpp:70-106) run once per frame through
 * StereoVision::AnalyzeAndAnnotateImage (src/lib.cpp:53-85).  Frames are tightly packed. */
typedef struct svgpu_pipeline_batch_io {
  const uint8_t* left;  /* [n][height][width][channels] */
  const uint8_t* right;
  int channels;         /* 3 (BGR) or 1 */
  uint8_t* left_rect;   /* optional [n][height][width][channels] (both or neither) */
  uint8_t* right_rect;
  int16_t* disp;        /* [n][height][width] CV_16S; input when SV_STAGE_SGBM is not requested */
  float* xyz;           /* optional [n][height][width][3] */
  float* depth;         /* optional [n][height][width] */
  int mem;              /* SV_MEM_HOST or SV_MEM_DEVICE */
} svgpu_pipeline_batch_io;
/* n pairs, in chunks of svgpu_reserve() pairs; host buffers: returns when the results are in them */
SVGPU_API int svgpu_pipeline_batch(svgpu_ctx* ctx, int n, const svgpu_pipeline_batch_io* io, unsigned stages);
/* Streaming form (host buffers, pinned for real overlap; n <= svgpu_reserve()): returns after enqueueing; copy-in of call
 * k+1, the kernels of call k and copy-out of call k-1 overlap on three streams.  Buffers must stay untouched until
 * svgpu_sync(ctx), which also reports errors.  Shares its two staging slots with svgpu_sgbm_enqueue_batch. */
SVGPU_API int svgpu_pipeline_enqueue_batch(svgpu_ctx* ctx, int n, const svgpu_pipeline_batch_io* io, unsigned stages);

/* ---- test hooks (used by tests/ to diff intermediates against the oracle) ------------- */
/* copies the cost volume Ctrue (which=0) or the aggregated S (which=1) of pair 0 of the last
 * sgbm call to host: [H][W1][D] int16, disparities in natural order */
SVGPU_API int svgpu_debug_copy_volume(svgpu_ctx* ctx, int which, int16_t* host_dst, size_t bytes);
/* total number of kernel launches issued through this context so far */
SVGPU_API long long svgpu_launch_count(const svgpu_ctx* ctx);
/* record CUDA events around the SGBM stages (off by default) and read them back: device time in ms of
 * {pre-filter + cost volume, path aggregation, WTA + LR check + median, total} of the last sgbm call */
SVGPU_API int svgpu_enable_stage_timing(svgpu_ctx* ctx, int on);
SVGPU_API int svgpu_last_stage_ms(svgpu_ctx* ctx, float ms[4]);
/* when on, the second aggregation sweep also stores the final S volume so that
 * svgpu_debug_copy_volume(which=1) returns the full sum over all paths (off: the sum of the first sweep) */
SVGPU_API int svgpu_debug_keep_volume(svgpu_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* SVGPU_H */

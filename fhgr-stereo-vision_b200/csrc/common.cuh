// common.cuh — shared declarations of the svgpu CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/svgpu.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "svgpu is written for sm_100a (B200) only"
#endif

namespace svgpu {

// Normalised SGBM configuration as the kernels see it.
struct SgbmCfg {
  int W, H;          // image size
  int D, minD;       // numDisparities, minDisparity
  int minX1, W1;     // first computed column, number of computed columns
  int P1, P2;        // normalised penalties
  int uniq, d12;     // normalised uniquenessRatio, disp12MaxDiff
  int ftzero;        // max(preFilterCap,15)|1
  int r;             // blockSize/2
  int mode;          // SV_MODE_*
  int invalid;       // (minD-1)*16
  int Wp;            // byte pitch of one pre-filter plane row
  int speckleWindowSize, speckleRange;
};

struct BmCfg {
  int W, H, D, minD, bs, cap, tex, uniq, d12, filtered;
  int speckleWindowSize, speckleRange;
};

// Device buffers for `nframes` pairs in flight. Volumes are [frame][y][x'][d] uint16.
struct SgbmScratch {
  uint8_t* planesL;   // [frame][H][2][Wp] uint2  Birchfield-Tomasi operands per channel (sgbm.cu)
  uint8_t* planesR;
  uint16_t* C;        // cost volume Ctrue
  uint16_t* S;        // aggregated volume
  uint2* sel;         // [frame][H][W1]  winner-take-all result of the second sweep (sgbm_device.cuh)
  int16_t* dispRaw;   // [frame][H][W]   before / after LR check
  int* status;        // device error flag (spin time-outs etc.)
  unsigned long long* flags;  // inter-CTA progress flags of the sweep kernels
  uint16_t* edges;    // inter-CTA edge exchange of the sweep kernels
  size_t frameCells;  // W1*H*D
  // the two aggregation passes of a few pairs run side by side (sgbm_sweep.cu): second stream, fork / join events and the
  // capacities of S, flags and edges in bytes (0 / null: passes run one after the other)
  cudaStream_t aux = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  size_t sBytes = 0, flagBytes = 0, edgeBytes = 0;
};

// MODE_SGBM_3WAY (OpenCV's computeDisparity3WAY; restated in oracle/sv_oracle_sgbm.c: sgbm_3way) cuts the image into four
// stripes of ceil(H/4) rows; stripe s is processed as an image of its own from `ov` rows above its first row ("stripe frame",
// rows[s] rows, padded to Hs).  Its result row y lands in row dstOff[s] + y of a private buffer whose rows [ov, ..) become the
// image rows [s*sz, ..): image row s*sz + k therefore comes from stripe row ov + k - dstOff[s].
struct StripeGeom {
  int sz, ov, Hs;
  int start[4], rows[4], dstOff[4];
};
inline StripeGeom stripe_geom(int H, int r) {
  StripeGeom g;
  g.sz = (H + 3) / 4;
  g.ov = (r + 1) + (g.sz + 9) / 10;
  g.Hs = 1;
  for (int s = 0; s < 4; s++) {
    const int a = s * g.sz - g.ov;
    g.start[s] = a > H ? H : (a < 0 ? 0 : a);
    const int e = (s + 1) * g.sz < H ? (s + 1) * g.sz : H;
    g.rows[s] = e > g.start[s] ? e - g.start[s] : 0;
    g.dstOff[s] = s == 0 ? g.ov : 0;
    if (g.rows[s] > g.Hs) g.Hs = g.rows[s];
  }
  return g;
}
// rows of volume / selection storage one pair needs
inline int sgbm_volume_rows(const SgbmCfg& cfg) { return cfg.mode == 2 /* SV_MODE_SGBM_3WAY */ ? 4 * stripe_geom(cfg.H, cfg.r).Hs : cfg.H; }

struct LaunchStats {
  long long launches = 0;
};

// sgbm.cu
cudaError_t sgbm_forward(const SgbmCfg& cfg, const SgbmScratch& sc, int nframes, const uint8_t* left, size_t lstride,
                         size_t lframe, const uint8_t* right, size_t rstride, size_t rframe, int16_t* disp,
                         size_t dstride, size_t dframe, cudaStream_t st, LaunchStats* stats, cudaEvent_t* ev, int debugS, int cn = 1);
// sgbm_cost.cu: Birchfield-Tomasi pixel cost + box sum -> cost volume
// what one launch of the cost kernel reads and writes: `nframes` frames of `rowsOut` rows each; cfg.H is the row at which the
// vertical window clamps (>= rowsOut), planeRows the rows per frame of the operand planes, frameCells the cells per frame of C
struct CostView {
  const uint8_t* planesL;
  const uint8_t* planesR;
  uint16_t* C;
  size_t frameCells;
  int planeRows, rowsOut;
};
cudaError_t sgbm_cost_volume(const SgbmCfg& cfg, const CostView& view, int nframes, int cn, cudaStream_t st);
// sgbm_sweep.cu: path aggregation + winner-take-all as two fused sweeps over skewed column strips
cudaError_t sgbm_aggregate_sweep(const SgbmCfg& cfg, const SgbmScratch& sc, int nframes, int debugS, cudaStream_t st,
                                 LaunchStats* stats, int cn = 1);
// sgbm_lines.cu: aggregation of single pairs, one worker per line and direction
int sgbm_lines_max_frames(const SgbmCfg& cfg, size_t frameCells);   // pairs per call up to which the line form is used
bool sgbm_lines_ok(const SgbmCfg& cfg, int nframes, size_t frameCells);
size_t sgbm_lines_frames(const SgbmCfg& cfg, int nframes);   // path volumes it needs in S
cudaError_t sgbm_aggregate_lines(const SgbmCfg& cfg, const SgbmScratch& sc, int nframes, cudaStream_t st, LaunchStats* stats);
size_t sgbm_s_frames(const SgbmCfg& cfg, int nframes);   // S volumes that `nframes` pairs per call need
size_t sgbm_flags_bytes(const SgbmCfg& cfg, int nframes);
size_t sgbm_edges_bytes(const SgbmCfg& cfg, int nframes);

// image_ops.cu
struct RectifyMapArgs {
  double iR[9];  // (P[:3,:3] * R)^-1, row-major
  double fx, fy, cx, cy, k1, k2, p1, p2, k3;
};
// rectify.cu (host)
struct StereoRectifyOut { double R1[9], R2[9], P1[12], P2[12], Q[16]; };
bool stereo_rectify_host(const double* K1, const double* D1, const double* K2, const double* D2, int W, int H, const double* R,
                         const double* T, StereoRectifyOut* o);
bool rectify_map_args(const double* K, const double* D, const double* R, const double* P, RectifyMapArgs* a);
cudaError_t launch_init_rectify_map(const RectifyMapArgs& k, int W, int H, int16_t* map1, uint16_t* map2, cudaStream_t st);
cudaError_t launch_remap(const uint8_t* src, int sw, int sh, size_t sstride, int cn, const int16_t* map1,
                         const uint16_t* map2, uint8_t* dst, int W, int H, size_t dstride, cudaStream_t st);
cudaError_t launch_median3(const int16_t* src, size_t sstride, int16_t* dst, size_t dstride, int W, int H, int nframes,
                           size_t sframe, size_t dframe, cudaStream_t st);
cudaError_t launch_reproject(const int16_t* disp, size_t dstride, float* xyz, size_t xstride, int W, int H,
                             const double* Qdev, int handleMissing, int* minScratch, cudaStream_t st, int* nlaunch);
cudaError_t launch_depth(const int16_t* disp, size_t dstride, float* depth, size_t zstride, int W, int H, float bf,
                         cudaStream_t st);
cudaError_t launch_filter_speckles(int16_t* disp, size_t stride, int W, int H, int newVal, int maxSize, int maxDiff,
                                   int* labels, int* sizes, int* changed, cudaStream_t st, int* nlaunch);
cudaError_t launch_morph_close5(const int16_t* src, size_t sstride, int16_t* tmp, int16_t* dst, size_t dstride, int W,
                                int H, cudaStream_t st);
cudaError_t launch_prefilter_chain(const uint8_t* bgr, size_t bstride, uint16_t* tmpH, uint8_t* tmpGray, uint8_t* gray,
                                   size_t gstride, int W, int H, cudaStream_t st);
cudaError_t launch_resize(const uint8_t* src, int sw, int sh, size_t sstride, int cn, uint8_t* dst, int dw, int dh,
                          size_t dstride, cudaStream_t st);

// front_fused.cu: raw BGR frames -> (rectified images) -> blur / gray / sharpen -> SGBM operand planes, one kernel
cudaError_t launch_front_fused(const uint8_t* left, size_t lstride, size_t lframe, const uint8_t* right, size_t rstride, size_t rframe,
                               const int16_t* const map1[2], const uint16_t* const map2[2], uint8_t* rectL, uint8_t* rectR, size_t rectStride,
                               size_t rectFrame, uint8_t* planesL, uint8_t* planesR, int W, int H, int Wp, int ftzero, int nframes,
                               cudaStream_t st);

// bm.cu
size_t bm_smem_bytes(int D, int blockSize);  // dynamic shared memory k_bm_sad needs for this configuration
cudaError_t bm_forward(const BmCfg& cfg, const uint8_t* left, size_t lstride, const uint8_t* right, size_t rstride,
                       uint8_t* pfL, uint8_t* pfR, int16_t* disp, size_t dstride, cudaStream_t st, int* nlaunch);

}  // namespace svgpu

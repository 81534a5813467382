// api.cu — the C ABI of include/svgpu.h: context, parameter normalisation, scratch management and
// host<->device staging around the kernels in sgbm.cu / bm.cu / image_ops.cu.
// There is deliberately no CPU path in this file: every entry point either runs CUDA kernels on the
// context's device or fails with an SV_* code.
#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sgbm_device.cuh"

using namespace svgpu;

struct svgpu_ctx {
  int device = 0, W = 0, H = 0;
  cudaStream_t own = nullptr, st = nullptr;
  std::mutex mu;
  std::string err;

  int16_t* map1[2] = {nullptr, nullptr};
  uint16_t* map2[2] = {nullptr, nullptr};
  double* Qdev = nullptr;
  bool haveQ = false, haveDepth = false, haveSgbm = false, haveBm = false;
  float baseline = 0, focal = 0;
  SgbmCfg cfg{};
  BmCfg bm{};
  bool debugS = false;  // test hook: the second sweep also stores the final S volume

  int maxBatch = 1;
  // scratch for the SGBM core; every buffer keyed on its byte size (ensure_scratch)
  SgbmScratch sc{};
  static constexpr int kScratchBufs = 9;
  size_t allocBytes[kScratchBufs] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // planesL, planesR, C, S, sel, dispRaw, status, flags, edges
  // device staging of caller images for SV_MEM_HOST calls and intermediates of the fused chain
  std::vector<void*> pool;  // everything cudaMalloc'ed lazily through dev_buf()
  struct Buf { void* p = nullptr; size_t n = 0; };
  Buf stLeft, stRight, stDisp, stXyz, stDepth, stRectL, stRectR, stGrayL, stGrayR, stTmpH, stTmpG, stDisp2, stLabels,
      stSizes, stSrc, stBmL, stBmR;
  int* minScratch = nullptr;
  // streaming entry point: copy-in / compute / copy-out of consecutive enqueues overlap (double-buffered staging)
  cudaStream_t stIn = nullptr, stOut = nullptr;
  cudaStream_t stAux = nullptr;                  // second aggregation pass of a few pairs (sgbm_sweep.cu)
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  Buf pipeL[2], pipeR[2], pipeD[2], pipeX[2], pipeZ[2], pipeRL[2], pipeRR[2];
  cudaEvent_t evIn[2] = {nullptr, nullptr}, evComp[2] = {nullptr, nullptr}, evOut[2] = {nullptr, nullptr};
  unsigned long long enqueued = 0;

  LaunchStats stats;
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float stageMs[4] = {0, 0, 0, 0};
};

namespace {

int fail(svgpu_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
int cuda_fail(svgpu_ctx* c, cudaError_t e, const char* where) {
  const int code = (e == cudaErrorMemoryAllocation) ? SV_OOM : SV_CUDA_ERROR;
  return fail(c, code, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CU(c, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail((c), e__, #call); } while (0)

int ensure(svgpu_ctx* c, svgpu_ctx::Buf& b, size_t bytes) {
  if (b.n >= bytes && b.p) return SV_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.n = 0;
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc(staging)");
  b.n = bytes;
  return SV_OK;
}
void release(svgpu_ctx::Buf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.n = 0; }

void free_scratch(svgpu_ctx* c) {
  cudaFree(c->sc.planesL); cudaFree(c->sc.planesR); cudaFree(c->sc.C); cudaFree(c->sc.S); cudaFree(c->sc.sel);
  cudaFree(c->sc.dispRaw); cudaFree(c->sc.status); cudaFree(c->sc.flags); cudaFree(c->sc.edges);
  c->sc = SgbmScratch{};
  for (size_t& b : c->allocBytes) b = 0;
}

// the sweep kernels raise sc.status when a hand-over timed out (which would mean a scheduling bug or a device
// fault); surfaced at every point where the host waits for the stream anyway
int check_status(svgpu_ctx* c) {
  if (!c->sc.status) {  // no SGBM call yet: still wait for the stream (callers rely on their copies being done)
    cudaError_t e = cudaStreamSynchronize(c->st);
    return e == cudaSuccess ? SV_OK : cuda_fail(c, e, "stream synchronize");
  }
  int h = 0;
  cudaError_t e = cudaMemcpyAsync(&h, c->sc.status, sizeof(int), cudaMemcpyDeviceToHost, c->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
  if (e != cudaSuccess) return cuda_fail(c, e, "status read");
  if (h) {
    cudaMemsetAsync(c->sc.status, 0, sizeof(int), c->st);
    return fail(c, SV_CUDA_ERROR, "sgbm aggregation sweep aborted (hand-over timed out)");
  }
  return SV_OK;
}

// Scratch of the SGBM core for `batch` pairs of the current configuration.  Every buffer is keyed on its own byte size
// (they depend on W1, D, the mode, the block radius and the strip-group count individually, not only on the cell count):
// a reconfiguration that needs more of any buffer reallocates, one that needs less reuses what is there.
int ensure_scratch(svgpu_ctx* c, int batch, int cn = 1) {
  const SgbmCfg& g = c->cfg;
  // MODE_SGBM_3WAY works on four stripe frames per pair (common.cuh: StripeGeom): volumes, selection and the sweep's
  // hand-over buffers are sized for 4 x batch frames of Hs rows, plus one extra full-size disparity plane per pair
  const bool way3 = g.mode == SV_MODE_SGBM_3WAY;
  const int vrows = sgbm_volume_rows(g);
  SgbmCfg gs = g;
  if (way3) gs.H = vrows / 4;
  const int sweepFrames = way3 ? batch * 4 : batch;
  const size_t cells = g.W1 > 0 ? (size_t)g.W1 * vrows * g.D : 0;
  const size_t w1 = g.W1 > 0 ? g.W1 : 0;
  // + 256: the TMA staging of the cost kernel copies whole 16-byte units and may read one operand word past a row
  const size_t planeBytes = (size_t)batch * g.H * 2 * cn * g.Wp * sizeof(uint2) + 256;
  // S holds one volume per pair, two when the passes of a few pairs run side by side, one per path for single small pairs
  // (MODE_SGBM_3WAY: the line form counts stripe frames, four per pair, of cells / 4 each)
  const size_t sFrames = way3 ? std::max<size_t>(batch, (sgbm_lines_frames(gs, std::min(4 * batch, sgbm_lines_max_frames(gs, cells / 4))) + 3) / 4)
                              : std::max(sgbm_s_frames(gs, batch), sgbm_lines_frames(gs, std::min(batch, sgbm_lines_max_frames(gs, cells))));
  const size_t need[svgpu_ctx::kScratchBufs] = {
      planeBytes, planeBytes, (size_t)batch * cells * 2, sFrames * cells * 2, (size_t)batch * vrows * w1 * 8,
      (size_t)batch * (vrows + (way3 ? g.H : 0)) * g.W * 2, 256, sgbm_flags_bytes(gs, sweepFrames), sgbm_edges_bytes(gs, sweepFrames)};
  bool fits = c->sc.dispRaw != nullptr;
  for (int i = 0; i < svgpu_ctx::kScratchBufs; i++) fits = fits && c->allocBytes[i] >= need[i];
  auto publish = [&] {
    c->sc.frameCells = cells;
    c->sc.aux = c->stAux; c->sc.evFork = c->evFork; c->sc.evJoin = c->evJoin;
    c->sc.sBytes = c->allocBytes[3]; c->sc.flagBytes = c->allocBytes[7]; c->sc.edgeBytes = c->allocBytes[8];
  };
  if (fits) {
    publish();
    return SV_OK;
  }
  // work that still uses the old buffers must be done before they go away
  cudaStreamSynchronize(c->st);
  size_t want[svgpu_ctx::kScratchBufs];
  for (int i = 0; i < svgpu_ctx::kScratchBufs; i++) want[i] = std::max<size_t>(std::max(need[i], c->allocBytes[i]), 256);
  free_scratch(c);
  void** ptrs[svgpu_ctx::kScratchBufs] = {(void**)&c->sc.planesL, (void**)&c->sc.planesR, (void**)&c->sc.C, (void**)&c->sc.S, (void**)&c->sc.sel,
                                          (void**)&c->sc.dispRaw, (void**)&c->sc.status, (void**)&c->sc.flags, (void**)&c->sc.edges};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < svgpu_ctx::kScratchBufs && e == cudaSuccess; i++) e = cudaMalloc(ptrs[i], want[i]);
  if (e != cudaSuccess) { free_scratch(c); return cuda_fail(c, e, "cudaMalloc(sgbm scratch)"); }
  for (int i = 0; i < svgpu_ctx::kScratchBufs; i++) c->allocBytes[i] = want[i];
  cudaMemsetAsync(c->sc.status, 0, want[6], c->st);
  cudaMemsetAsync(c->sc.flags, 0, want[7], c->st);
  cudaMemsetAsync(c->sc.edges, 0, want[8], c->st);
  publish();
  return SV_OK;
}

// A.1 parameter normalisation + the contract checks of this implementation
int normalise_sgbm(svgpu_ctx* c, const svgpu_sgbm_params& p, SgbmCfg& g) {
  if (p.numDisparities <= 0 || p.numDisparities % 16 != 0) return fail(c, SV_INVALID_SETTINGS, "numDisparities must be a positive multiple of 16");
  // up to 256 disparities every form of the aggregation is built; up to 512 the line form (a full warp per pixel) serves all
  // call sizes
  if (p.numDisparities > 512) return fail(c, SV_INVALID_SETTINGS, "numDisparities > 512 is not supported");
  if (p.mode != SV_MODE_SGBM && p.mode != SV_MODE_HH && p.mode != SV_MODE_HH4 && p.mode != SV_MODE_SGBM_3WAY)
    return fail(c, SV_INVALID_SETTINGS, "mode must be MODE_SGBM, MODE_HH, MODE_SGBM_3WAY or MODE_HH4");
  // 15 x 15 is the largest block whose cost fits 16 bits at all (225 x 93 + P2, checked below)
  if (p.blockSize > 15 || (p.blockSize > 0 && p.blockSize % 2 == 0)) return fail(c, SV_INVALID_SETTINGS, "blockSize must be odd and <= 15");
  if (p.preFilterCap > 63) return fail(c, SV_INVALID_SETTINGS, "preFilterCap must be <= 63");
  g.W = c->W; g.H = c->H;
  g.D = p.numDisparities; g.minD = p.minDisparity;
  const int maxD = g.minD + g.D;
  g.minX1 = std::max(maxD, 0);
  g.W1 = (g.W + std::min(g.minD, 0)) - g.minX1;
  g.P1 = p.P1 > 0 ? p.P1 : 2;
  g.P2 = std::max(p.P2 > 0 ? p.P2 : 5, g.P1 + 1);
  g.uniq = p.uniquenessRatio >= 0 ? p.uniquenessRatio : 10;
  g.d12 = p.disp12MaxDiff > 0 ? p.disp12MaxDiff : 1;
  g.ftzero = std::max(p.preFilterCap, 15) | 1;
  const int bs = p.blockSize > 0 ? p.blockSize : 5;
  g.r = bs / 2;
  g.mode = p.mode;
  g.invalid = (g.minD - 1) * 16;
  g.Wp = (g.W + 15) & ~15;
  g.speckleWindowSize = p.speckleWindowSize;
  g.speckleRange = p.speckleRange;
  if (g.uniq > 100) return fail(c, SV_INVALID_SETTINGS, "uniquenessRatio must be <= 100");
  // OpenCV's MODE_SGBM_3WAY divides by (100 - uniquenessRatio)
  if (g.mode == SV_MODE_SGBM_3WAY && g.uniq >= 100) return fail(c, SV_INVALID_SETTINGS, "uniquenessRatio must be < 100 in MODE_SGBM_3WAY");
  // 16-bit headroom: every path cost L <= Ctrue + P2 must stay below MAX_COST, as in OpenCV's own contract
  const long long maxC = (long long)bs * bs * (2 * g.ftzero + 63);
  if (maxC + g.P2 > 32767) return fail(c, SV_INVALID_SETTINGS, "blockSize^2*(2*ftzero+63) + P2 exceeds the 16-bit cost range");
  if (std::abs(g.minD) > 4096) return fail(c, SV_INVALID_SETTINGS, "minDisparity out of range");
  return SV_OK;
}

struct HostIO {  // moves caller images through device staging for SV_MEM_HOST
  svgpu_ctx* c;
};

int h2d_2d(svgpu_ctx* c, void* dst, size_t dpitch, const void* src, size_t spitch, size_t widthBytes, int rows) {
  CU(c, cudaMemcpy2DAsync(dst, dpitch, src, spitch, widthBytes, rows, cudaMemcpyHostToDevice, c->st));
  return SV_OK;
}
int d2h_2d(svgpu_ctx* c, void* dst, size_t dpitch, const void* src, size_t spitch, size_t widthBytes, int rows) {
  CU(c, cudaMemcpy2DAsync(dst, dpitch, src, spitch, widthBytes, rows, cudaMemcpyDeviceToHost, c->st));
  return SV_OK;
}

// planesReady: the operand planes of the n pairs were written into the scratch by the fused front kernel (dl / dr unused)
int run_sgbm(svgpu_ctx* c, int n, const uint8_t* dl, size_t ls, size_t lf, const uint8_t* dr, size_t rs, size_t rf,
             int16_t* dd, size_t ds, size_t df, int cn = 1, bool planesReady = false) {
  // 8UC3 input: the pixel cost is the sum over the three channels, so the 16-bit headroom shrinks accordingly
  if (cn == 3 && (long long)(2 * c->cfg.r + 1) * (2 * c->cfg.r + 1) * 3 * (2 * c->cfg.ftzero + 63) + c->cfg.P2 > 32767)
    return fail(c, SV_INVALID_SETTINGS, "3 * blockSize^2*(2*ftzero+63) + P2 exceeds the 16-bit cost range for 3-channel input");
  // cv::StereoSGBM::compute throws for MODE_SGBM / MODE_HH when the image is too narrow for the
  // disparity range ('width - (minDisparity + numDisparities) > SADWindowSize/2'); MODE_HH4 returns an
  // all-invalid map instead.  Same behaviour here, as an error code.
  if (c->cfg.mode != SV_MODE_HH4 && c->W - (c->cfg.minD + c->cfg.D) <= c->cfg.r)
    return fail(c, SV_INVALID_SETTINGS, "image too narrow for minDisparity + numDisparities and blockSize");
  // the kernels address a frame's volume with 32-bit offsets in 16-byte units
  if ((unsigned long long)std::max(c->cfg.W1, 0) * sgbm_volume_rows(c->cfg) * c->cfg.D / 8 >= (1ull << 32))
    return fail(c, SV_INVALID_SETTINGS, "cost volume of one pair exceeds 64 GiB");
  int rc = ensure_scratch(c, std::max(n, c->maxBatch), cn);
  if (rc) return rc;
  if (planesReady) dl = dr = nullptr;
  cudaError_t e = sgbm_forward(c->cfg, c->sc, n, dl, ls, lf, dr, rs, rf, dd, ds, df, c->st, &c->stats,
                               c->timing ? c->ev : nullptr, c->debugS ? 1 : 0, cn);
  if (e != cudaSuccess) return cuda_fail(c, e, "sgbm_forward");
  if (c->cfg.speckleWindowSize > 0) {  // built-in stage of StereoSGBM::compute (A.7): newVal = INVALID, maxDiff = 16*range
    rc = ensure(c, c->stLabels, (size_t)c->W * c->H * 4);
    if (!rc) rc = ensure(c, c->stSizes, (size_t)c->W * c->H * 4);
    if (rc) return rc;
    for (int f = 0; f < n; f++) {
      int nl = 0;
      e = launch_filter_speckles((int16_t*)((uint8_t*)dd + f * df), ds, c->W, c->H, c->cfg.invalid, c->cfg.speckleWindowSize,
                                 16 * c->cfg.speckleRange, (int*)c->stLabels.p, (int*)c->stSizes.p, nullptr, c->st, &nl);
      c->stats.launches += nl;
      if (e != cudaSuccess) return cuda_fail(c, e, "filter_speckles");
    }
  }
  return SV_OK;
}


// ---- the chain of src/lib.cpp:53-85 on device buffers, n frames --------------------------------------------------
struct PipeDev {
  const uint8_t *left, *right;   // n frames of `cn` channels: row strides ls / rs, frame strides lf / rf (bytes)
  size_t ls, rs, lf, rf;
  int cn;
  uint8_t *rectL, *rectR;        // remap output (null: context scratch); filled in with what was used
  size_t rectS, rectF;
  bool wantRect;                 // the caller reads the rectified images after the call (else they may never be materialised)
  int16_t* disp;                 // in (without the SGBM stage) / out
  size_t dS, dF;
  float* xyz;                    // null: skipped
  size_t xS, xF;
  float* depth;
  size_t zS, zF;
};

int pipeline_check(svgpu_ctx* c, int cn, unsigned stages, bool haveDisp) {
  if (cn != 1 && cn != 3) return fail(c, SV_INVALID_SETTINGS, "channels must be 1 or 3");
  if ((stages & SV_STAGE_PREFILTER) && cn != 3) return fail(c, SV_INVALID_SETTINGS, "the pre-filter chain needs BGR input");
  if ((stages & SV_STAGE_SGBM) && !c->haveSgbm) return fail(c, SV_NOT_CONFIGURED, "sgbm parameters not set");
  if ((stages & SV_STAGE_REMAP) && (!c->map1[0] || !c->map1[1])) return fail(c, SV_NOT_CONFIGURED, "rectify maps not set");
  if ((stages & SV_STAGE_REPROJECT) && !c->haveQ) return fail(c, SV_NOT_CONFIGURED, "Q not set");
  if ((stages & SV_STAGE_DEPTH) && !c->haveDepth) return fail(c, SV_NOT_CONFIGURED, "depth constants not set");
  if ((stages & (SV_STAGE_POSTFILTER | SV_STAGE_REPROJECT | SV_STAGE_DEPTH)) && !(stages & SV_STAGE_SGBM) && !haveDisp)
    return fail(c, SV_INVALID_IMAGE, "no disparity input");
  return SV_OK;
}

int ensure_streams(svgpu_ctx* c) {
  if (c->stIn) return SV_OK;
  CU(c, cudaStreamCreateWithFlags(&c->stIn, cudaStreamNonBlocking));
  CU(c, cudaStreamCreateWithFlags(&c->stOut, cudaStreamNonBlocking));
  for (int s = 0; s < 2; s++)
    for (cudaEvent_t* e : {&c->evIn[s], &c->evComp[s], &c->evOut[s]}) CU(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  return SV_OK;
}

int pipeline_device(svgpu_ctx* c, int n, PipeDev& p, unsigned stages) {
  const size_t W = c->W, H = c->H;
  const int cn = p.cn;
  int rc;
  const uint8_t *curL = p.left, *curR = p.right;
  size_t curLs = p.ls, curRs = p.rs, curLf = p.lf, curRf = p.rf;
  // Raw BGR frames through remap + blur / gray / sharpen + SGBM: one kernel produces the operand planes of the cost kernel
  // (front_fused.cu); rectified pixels, blurred rows and gray images never go through HBM.  SVGPU_UNFUSED=1 (read once) keeps
  // the stage-by-stage kernels, which the parity tests compare against.
  static const bool unfused = [] { const char* e = std::getenv("SVGPU_UNFUSED"); return e && e[0] == '1'; }();
  const unsigned kFront = SV_STAGE_REMAP | SV_STAGE_PREFILTER | SV_STAGE_SGBM;
  const bool fused = !unfused && (stages & kFront) == kFront && cn == 3 && c->cfg.W1 > 0 &&
                     (c->cfg.mode == SV_MODE_HH4 || c->W - (c->cfg.minD + c->cfg.D) > c->cfg.r);
  if (fused) {
    if (p.wantRect && (!p.rectL || !p.rectR)) {
      rc = ensure(c, c->stRectL, W * H * cn * n); if (rc) return rc;
      rc = ensure(c, c->stRectR, W * H * cn * n); if (rc) return rc;
      p.rectL = (uint8_t*)c->stRectL.p; p.rectR = (uint8_t*)c->stRectR.p; p.rectS = W * cn; p.rectF = W * H * cn;
    }
    rc = ensure_scratch(c, std::max(n, c->maxBatch), 1); if (rc) return rc;
    const int16_t* m1[2] = {c->map1[0], c->map1[1]};
    const uint16_t* m2[2] = {c->map2[0], c->map2[1]};
    CU(c, launch_front_fused(curL, curLs, curLf, curR, curRs, curRf, m1, m2, p.wantRect ? p.rectL : nullptr, p.wantRect ? p.rectR : nullptr, p.rectS,
                             p.rectF, c->sc.planesL, c->sc.planesR, c->W, c->H, c->cfg.Wp, c->cfg.ftzero, n, c->st));
    c->stats.launches++;
    rc = run_sgbm(c, n, nullptr, 0, 0, nullptr, 0, 0, p.disp, p.dS, p.dF, 1, true); if (rc) return rc;
    stages &= ~kFront;
  }
  if (stages & SV_STAGE_REMAP) {  // src/lib.cpp:87-94
    if (!p.rectL || !p.rectR) {
      rc = ensure(c, c->stRectL, W * H * cn * n); if (rc) return rc;
      rc = ensure(c, c->stRectR, W * H * cn * n); if (rc) return rc;
      p.rectL = (uint8_t*)c->stRectL.p; p.rectR = (uint8_t*)c->stRectR.p; p.rectS = W * cn; p.rectF = W * H * cn;
    }
    for (int f = 0; f < n; f++) {
      CU(c, launch_remap(curL + f * curLf, c->W, c->H, curLs, cn, c->map1[0], c->map2[0], p.rectL + f * p.rectF, c->W, c->H, p.rectS, c->st));
      CU(c, launch_remap(curR + f * curRf, c->W, c->H, curRs, cn, c->map1[1], c->map2[1], p.rectR + f * p.rectF, c->W, c->H, p.rectS, c->st));
    }
    c->stats.launches += 2 * n;
    curL = p.rectL; curR = p.rectR; curLs = curRs = p.rectS; curLf = curRf = p.rectF;
  }
  if (stages & SV_STAGE_PREFILTER) {  // src/lib.cpp:231-246
    rc = ensure(c, c->stTmpH, W * H * 6); if (rc) return rc;
    rc = ensure(c, c->stTmpG, W * H); if (rc) return rc;
    rc = ensure(c, c->stGrayL, W * H * n); if (rc) return rc;
    rc = ensure(c, c->stGrayR, W * H * n); if (rc) return rc;
    for (int f = 0; f < n; f++) {
      CU(c, launch_prefilter_chain(curL + f * curLf, curLs, (uint16_t*)c->stTmpH.p, (uint8_t*)c->stTmpG.p, (uint8_t*)c->stGrayL.p + f * W * H, W, c->W, c->H, c->st));
      CU(c, launch_prefilter_chain(curR + f * curRf, curRs, (uint16_t*)c->stTmpH.p, (uint8_t*)c->stTmpG.p, (uint8_t*)c->stGrayR.p + f * W * H, W, c->W, c->H, c->st));
    }
    c->stats.launches += 6 * n;
    curL = (uint8_t*)c->stGrayL.p; curR = (uint8_t*)c->stGrayR.p; curLs = curRs = W; curLf = curRf = W * H;
  }
  if (stages & SV_STAGE_SGBM) {
    // without the pre-filter chain a BGR pair goes to StereoSGBM as it is (cv::StereoSGBM::compute on 8UC3 images)
    rc = run_sgbm(c, n, curL, curLs, curLf, curR, curRs, curRf, p.disp, p.dS, p.dF, (stages & SV_STAGE_PREFILTER) ? 1 : cn); if (rc) return rc;
  }
  if (stages & SV_STAGE_POSTFILTER) {
    // src/lib.cpp:264 is a no-op for SGBM output (invalid = -16, never -1); :265 and :267-268 follow
    rc = ensure(c, c->stLabels, W * H * 4); if (rc) return rc;
    rc = ensure(c, c->stSizes, W * H * 4); if (rc) return rc;
    rc = ensure(c, c->stDisp2, W * H * 2); if (rc) return rc;
    for (int f = 0; f < n; f++) {
      int16_t* dd = (int16_t*)((uint8_t*)p.disp + f * p.dF);
      int nl = 0;
      CU(c, launch_filter_speckles(dd, p.dS, c->W, c->H, 0, 150, 32, (int*)c->stLabels.p, (int*)c->stSizes.p, nullptr, c->st, &nl));
      CU(c, launch_morph_close5(dd, p.dS, (int16_t*)c->stLabels.p, (int16_t*)c->stDisp2.p, W * 2, c->W, c->H, c->st));
      CU(c, cudaMemcpy2DAsync(dd, p.dS, c->stDisp2.p, W * 2, W * 2, H, cudaMemcpyDeviceToDevice, c->st));
      c->stats.launches += nl + 1;
    }
  }
  if ((stages & SV_STAGE_REPROJECT) && p.xyz) {  // src/lib.cpp:152-156
    for (int f = 0; f < n; f++) {
      int nl = 0;
      CU(c, launch_reproject((const int16_t*)((const uint8_t*)p.disp + f * p.dF), p.dS, (float*)((uint8_t*)p.xyz + f * p.xF), p.xS, c->W, c->H, c->Qdev, 1,
                             c->minScratch, c->st, &nl));
      c->stats.launches += nl;
    }
  }
  if ((stages & SV_STAGE_DEPTH) && p.depth) {  // src/lib.cpp:279-294
    for (int f = 0; f < n; f++)
      CU(c, launch_depth((const int16_t*)((const uint8_t*)p.disp + f * p.dF), p.dS, (float*)((uint8_t*)p.depth + f * p.zF), p.zS, c->W, c->H,
                         c->baseline * c->focal, c->st));
    c->stats.launches += n;
  }
  return SV_OK;
}

}  // namespace

extern "C" {

SVGPU_API const char* svgpu_version(void) { return "svgpu 0.1 sm_100a"; }

SVGPU_API int svgpu_create(int device, int width, int height, svgpu_ctx** out) {
  if (!out) return SV_INVALID_SETTINGS;
  *out = nullptr;
  if (width <= 0 || height <= 0 || width > 16384 || height > 16384) return SV_INVALID_SETTINGS;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SV_CUDA_ERROR;
  if (cudaSetDevice(device) != cudaSuccess) return SV_CUDA_ERROR;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SV_CUDA_ERROR;
  if (prop.major < 10) return SV_CUDA_ERROR;  // sm_100a SASS only, no fallback
  svgpu_ctx* c = new (std::nothrow) svgpu_ctx();
  if (!c) return SV_OOM;
  c->device = device; c->W = width; c->H = height;
  if (cudaStreamCreateWithFlags(&c->own, cudaStreamNonBlocking) != cudaSuccess) { delete c; return SV_CUDA_ERROR; }
  c->st = c->own;
  for (auto& e : c->ev) cudaEventCreate(&e);
  if (cudaStreamCreateWithFlags(&c->stAux, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming) != cudaSuccess) {
    svgpu_destroy(c);
    return SV_CUDA_ERROR;
  }
  if (cudaMalloc((void**)&c->Qdev, 16 * sizeof(double)) != cudaSuccess || cudaMalloc((void**)&c->minScratch, 256) != cudaSuccess) {
    svgpu_destroy(c);
    return SV_OOM;
  }
  *out = c;
  return SV_OK;
}

SVGPU_API int svgpu_destroy(svgpu_ctx* c) {
  if (!c) return SV_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->st);
  free_scratch(c);
  for (int s = 0; s < 2; s++) { cudaFree(c->map1[s]); cudaFree(c->map2[s]); }
  cudaFree(c->Qdev); cudaFree(c->minScratch);
  for (svgpu_ctx::Buf* b : {&c->stLeft, &c->stRight, &c->stDisp, &c->stXyz, &c->stDepth, &c->stRectL, &c->stRectR, &c->stGrayL,
                            &c->stGrayR, &c->stTmpH, &c->stTmpG, &c->stDisp2, &c->stLabels, &c->stSizes, &c->stSrc, &c->stBmL, &c->stBmR})
    release(*b);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  for (int s = 0; s < 2; s++) {
    release(c->pipeL[s]); release(c->pipeR[s]); release(c->pipeD[s]); release(c->pipeX[s]); release(c->pipeZ[s]); release(c->pipeRL[s]); release(c->pipeRR[s]);
    for (cudaEvent_t e : {c->evIn[s], c->evComp[s], c->evOut[s]}) if (e) cudaEventDestroy(e);
  }
  if (c->stIn) { cudaStreamSynchronize(c->stIn); cudaStreamDestroy(c->stIn); }
  if (c->stOut) { cudaStreamSynchronize(c->stOut); cudaStreamDestroy(c->stOut); }
  if (c->stAux) { cudaStreamSynchronize(c->stAux); cudaStreamDestroy(c->stAux); }
  for (cudaEvent_t e : {c->evFork, c->evJoin}) if (e) cudaEventDestroy(e);
  if (c->own) cudaStreamDestroy(c->own);
  delete c;
  return SV_OK;
}

SVGPU_API int svgpu_set_stream(svgpu_ctx* c, void* s) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  c->st = (cudaStream_t)s;
  return SV_OK;
}

SVGPU_API int svgpu_use_own_stream(svgpu_ctx* c) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  c->st = c->own;
  return SV_OK;
}

SVGPU_API int svgpu_reserve(svgpu_ctx* c, int maxBatch) {
  if (!c || maxBatch < 1 || maxBatch > 1024) return fail(c, SV_INVALID_SETTINGS, "max_batch out of range");
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  c->maxBatch = maxBatch;
  if (c->haveSgbm) return ensure_scratch(c, maxBatch);
  return SV_OK;
}

SVGPU_API int svgpu_sync(svgpu_ctx* c) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  if (c->stIn) CU(c, cudaStreamSynchronize(c->stIn));
  CU(c, cudaStreamSynchronize(c->st));
  if (c->stOut) CU(c, cudaStreamSynchronize(c->stOut));
  return check_status(c);
}

SVGPU_API const char* svgpu_last_error(const svgpu_ctx* c) { return c ? c->err.c_str() : "null context"; }

SVGPU_API int svgpu_set_rectify_maps(svgpu_ctx* c, int side, const int16_t* m1, const uint16_t* m2) {
  if (!c || side < 0 || side > 1 || !m1 || !m2) return fail(c, SV_INVALID_SETTINGS, "bad rectify map arguments");
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  const size_t n = (size_t)c->W * c->H;
  if (!c->map1[side]) CU(c, cudaMalloc((void**)&c->map1[side], n * 4));
  if (!c->map2[side]) CU(c, cudaMalloc((void**)&c->map2[side], n * 2));
  CU(c, cudaMemcpyAsync(c->map1[side], m1, n * 4, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->map2[side], m2, n * 2, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return SV_OK;
}

SVGPU_API int svgpu_init_rectify_maps(svgpu_ctx* c, int side, const double K[9], const double D[5], const double R[9],
                                      const double P[12]) {
  if (!c || side < 0 || side > 1 || !K || !D || !R || !P) return fail(c, SV_INVALID_SETTINGS, "bad rectify map arguments");
  RectifyMapArgs a;
  if (!rectify_map_args(K, D, R, P, &a)) return fail(c, SV_INVALID_SETTINGS, "P[:3,:3] * R is singular");
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  const size_t n = (size_t)c->W * c->H;
  if (!c->map1[side]) CU(c, cudaMalloc((void**)&c->map1[side], n * 4));
  if (!c->map2[side]) CU(c, cudaMalloc((void**)&c->map2[side], n * 2));
  CU(c, launch_init_rectify_map(a, c->W, c->H, c->map1[side], c->map2[side], c->st));
  c->stats.launches++;
  CU(c, cudaStreamSynchronize(c->st));
  return SV_OK;
}

SVGPU_API int svgpu_get_rectify_maps(svgpu_ctx* c, int side, int16_t* m1, uint16_t* m2) {
  if (!c || side < 0 || side > 1 || !m1 || !m2) return fail(c, SV_INVALID_SETTINGS, "bad rectify map arguments");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->map1[side] || !c->map2[side]) return fail(c, SV_NOT_CONFIGURED, "rectification maps not set");
  cudaSetDevice(c->device);
  const size_t n = (size_t)c->W * c->H;
  CU(c, cudaMemcpyAsync(m1, c->map1[side], n * 4, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(m2, c->map2[side], n * 2, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return SV_OK;
}

SVGPU_API int svgpu_stereo_rectify(const double K1[9], const double D1[5], const double K2[9], const double D2[5], int width,
                                   int height, const double R[9], const double T[3], double R1[9], double R2[9], double P1[12],
                                   double P2[12], double Q[16]) {
  if (!K1 || !D1 || !K2 || !D2 || !R || !T || !R1 || !R2 || !P1 || !P2 || !Q || width <= 1 || height <= 1) return SV_INVALID_SETTINGS;
  StereoRectifyOut o;
  if (!stereo_rectify_host(K1, D1, K2, D2, width, height, R, T, &o)) return SV_INVALID_SETTINGS;
  std::memcpy(R1, o.R1, sizeof o.R1); std::memcpy(R2, o.R2, sizeof o.R2);
  std::memcpy(P1, o.P1, sizeof o.P1); std::memcpy(P2, o.P2, sizeof o.P2);
  std::memcpy(Q, o.Q, sizeof o.Q);
  return SV_OK;
}

SVGPU_API int svgpu_set_Q(svgpu_ctx* c, const double Q[16]) {
  if (!c || !Q) return fail(c, SV_INVALID_SETTINGS, "bad Q");
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  CU(c, cudaMemcpyAsync(c->Qdev, Q, 16 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  c->haveQ = true;
  return SV_OK;
}

SVGPU_API int svgpu_set_depth_consts(svgpu_ctx* c, float baseline, float focal) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  c->baseline = baseline; c->focal = focal; c->haveDepth = true;
  return SV_OK;
}

SVGPU_API int svgpu_set_sgbm(svgpu_ctx* c, const svgpu_sgbm_params* p) {
  if (!c || !p) return fail(c, SV_INVALID_SETTINGS, "null sgbm params");
  std::lock_guard<std::mutex> lk(c->mu);
  SgbmCfg g{};
  int rc = normalise_sgbm(c, *p, g);
  if (rc) return rc;
  c->cfg = g; c->haveSgbm = true;
  return SV_OK;
}

SVGPU_API int svgpu_debug_keep_volume(svgpu_ctx* c, int on) {
  if (!c) return SV_INVALID_SETTINGS;
  c->debugS = on != 0;
  return SV_OK;
}

SVGPU_API int svgpu_set_bm(svgpu_ctx* c, const svgpu_bm_params* p) {
  if (!c || !p) return fail(c, SV_INVALID_SETTINGS, "null bm params");
  std::lock_guard<std::mutex> lk(c->mu);
  if (p->numDisparities <= 0 || p->numDisparities % 16 || p->numDisparities > 256) return fail(c, SV_INVALID_SETTINGS, "bm numDisparities must be a multiple of 16 in [16,256]");
  if (p->blockSize < 5 || p->blockSize % 2 == 0 || p->blockSize > 51) return fail(c, SV_INVALID_SETTINGS, "bm blockSize must be odd in [5,51]");
  if (p->preFilterCap < 1 || p->preFilterCap > 63) return fail(c, SV_INVALID_SETTINGS, "bm preFilterCap must be in [1,63]");
  if (p->minDisparity < -1024 || p->minDisparity + p->numDisparities > 1024) return fail(c, SV_INVALID_SETTINGS, "bm disparity range must lie within [-1024, 1024]");
  if (p->disp12MaxDiff >= 0) return fail(c, SV_INVALID_SETTINGS, "bm disp12MaxDiff >= 0 is not supported");
  if ((long long)p->blockSize * p->blockSize * 2 * p->preFilterCap > 65535) return fail(c, SV_INVALID_SETTINGS, "bm blockSize^2*2*cap exceeds the 16-bit SAD range");
  if (bm_smem_bytes(p->numDisparities, p->blockSize) > 227 * 1024)
    return fail(c, SV_INVALID_SETTINGS, "bm: (blockSize + 1) x (32 + blockSize) x numDisparities bytes of row ring exceed the 227 KB of shared memory per SM; reduce blockSize or numDisparities");
  BmCfg b{};
  b.W = c->W; b.H = c->H; b.D = p->numDisparities; b.minD = p->minDisparity; b.bs = p->blockSize; b.cap = p->preFilterCap;
  b.tex = p->textureThreshold; b.uniq = p->uniquenessRatio; b.d12 = p->disp12MaxDiff; b.filtered = (b.minD - 1) * 16;
  b.speckleWindowSize = p->speckleWindowSize; b.speckleRange = p->speckleRange;
  c->bm = b; c->haveBm = true;
  return SV_OK;
}

SVGPU_API int svgpu_enable_stage_timing(svgpu_ctx* c, int on) {
  if (!c) return SV_INVALID_SETTINGS;
  c->timing = on != 0;
  return SV_OK;
}

SVGPU_API int svgpu_last_stage_ms(svgpu_ctx* c, float ms[4]) {
  if (!c || !ms) return SV_INVALID_SETTINGS;
  if (!c->timing) return fail(c, SV_NOT_CONFIGURED, "stage timing is off");
  CU(c, cudaEventSynchronize(c->ev[3]));
  for (int i = 0; i < 3; i++) CU(c, cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
  CU(c, cudaEventElapsedTime(&ms[3], c->ev[0], c->ev[3]));
  return SV_OK;
}

SVGPU_API long long svgpu_launch_count(const svgpu_ctx* c) { return c ? c->stats.launches : 0; }

SVGPU_API int svgpu_remap(svgpu_ctx* c, int side, const uint8_t* src, int sw, int sh, size_t sstride, int cn, uint8_t* dst,
                size_t dstride, int mem) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!src || !dst || sw <= 0 || sh <= 0) return fail(c, SV_INVALID_IMAGE, "empty image");
  if (side < 0 || side > 1 || (cn != 1 && cn != 3)) return fail(c, SV_INVALID_SETTINGS, "bad side / channels");
  if (!c->map1[side] || !c->map2[side]) return fail(c, SV_NOT_CONFIGURED, "rectify maps not set");
  cudaSetDevice(c->device);
  const uint8_t* ds = src; uint8_t* dd = dst; size_t dss = sstride, dds = dstride;
  if (mem == SV_MEM_HOST) {
    int rc = ensure(c, c->stSrc, (size_t)sw * sh * cn); if (rc) return rc;
    rc = ensure(c, c->stRectL, (size_t)c->W * c->H * cn); if (rc) return rc;
    rc = h2d_2d(c, c->stSrc.p, (size_t)sw * cn, src, sstride, (size_t)sw * cn, sh); if (rc) return rc;
    ds = (const uint8_t*)c->stSrc.p; dss = (size_t)sw * cn; dd = (uint8_t*)c->stRectL.p; dds = (size_t)c->W * cn;
  }
  CU(c, launch_remap(ds, sw, sh, dss, cn, c->map1[side], c->map2[side], dd, c->W, c->H, dds, c->st));
  c->stats.launches++;
  if (mem == SV_MEM_HOST) {
    int rc = d2h_2d(c, dst, dstride, dd, dds, (size_t)c->W * cn, c->H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

SVGPU_API int svgpu_sgbm_compute(svgpu_ctx* c, const uint8_t* left, size_t ls, const uint8_t* right, size_t rs, int16_t* disp,
                       size_t dstride, int mem) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!left || !right || !disp) return fail(c, SV_INVALID_IMAGE, "empty image");
  if (!c->haveSgbm) return fail(c, SV_NOT_CONFIGURED, "sgbm parameters not set");
  cudaSetDevice(c->device);
  const size_t W = c->W, H = c->H;
  if (mem == SV_MEM_HOST) {
    int rc = ensure(c, c->stLeft, W * H); if (rc) return rc;
    rc = ensure(c, c->stRight, W * H); if (rc) return rc;
    rc = ensure(c, c->stDisp, W * H * 2); if (rc) return rc;
    rc = h2d_2d(c, c->stLeft.p, W, left, ls, W, (int)H); if (rc) return rc;
    rc = h2d_2d(c, c->stRight.p, W, right, rs, W, (int)H); if (rc) return rc;
    rc = run_sgbm(c, 1, (uint8_t*)c->stLeft.p, W, W * H, (uint8_t*)c->stRight.p, W, W * H, (int16_t*)c->stDisp.p, W * 2, W * H * 2);
    if (rc) return rc;
    rc = d2h_2d(c, disp, dstride, c->stDisp.p, W * 2, W * 2, (int)H); if (rc) return rc;
    return check_status(c);
  }
  return run_sgbm(c, 1, left, ls, 0, right, rs, 0, disp, dstride, 0);
}

SVGPU_API int svgpu_sgbm_compute_batch(svgpu_ctx* c, int n, const uint8_t* left, const uint8_t* right, int16_t* disp, int mem) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!left || !right || !disp || n <= 0) return fail(c, SV_INVALID_IMAGE, "empty batch");
  if (!c->haveSgbm) return fail(c, SV_NOT_CONFIGURED, "sgbm parameters not set");
  cudaSetDevice(c->device);
  const size_t W = c->W, H = c->H, img = W * H;
  const int chunk = c->maxBatch;
  if (mem == SV_MEM_HOST) {
    int rc = ensure(c, c->stLeft, img * chunk); if (rc) return rc;
    rc = ensure(c, c->stRight, img * chunk); if (rc) return rc;
    rc = ensure(c, c->stDisp, img * 2 * chunk); if (rc) return rc;
  }
  for (int f0 = 0; f0 < n; f0 += chunk) {
    const int m = std::min(chunk, n - f0);
    if (mem == SV_MEM_HOST) {
      CU(c, cudaMemcpyAsync(c->stLeft.p, left + f0 * img, img * m, cudaMemcpyHostToDevice, c->st));
      CU(c, cudaMemcpyAsync(c->stRight.p, right + f0 * img, img * m, cudaMemcpyHostToDevice, c->st));
      int rc = run_sgbm(c, m, (uint8_t*)c->stLeft.p, W, img, (uint8_t*)c->stRight.p, W, img, (int16_t*)c->stDisp.p, W * 2, img * 2);
      if (rc) return rc;
      CU(c, cudaMemcpyAsync(disp + f0 * img, c->stDisp.p, img * 2 * m, cudaMemcpyDeviceToHost, c->st));
    } else {
      int rc = run_sgbm(c, m, left + f0 * img, W, img, right + f0 * img, W, img, disp + f0 * img, W * 2, img * 2);
      if (rc) return rc;
    }
  }
  if (mem == SV_MEM_HOST) return check_status(c);
  return SV_OK;
}

SVGPU_API int svgpu_sgbm_enqueue_batch(svgpu_ctx* c, int n, const uint8_t* left, const uint8_t* right, int16_t* disp) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!left || !right || !disp || n <= 0) return fail(c, SV_INVALID_IMAGE, "empty batch");
  if (!c->haveSgbm) return fail(c, SV_NOT_CONFIGURED, "sgbm parameters not set");
  if (n > c->maxBatch) return fail(c, SV_INVALID_SETTINGS, "enqueue_batch: n exceeds svgpu_reserve()");
  cudaSetDevice(c->device);
  const size_t W = c->W, H = c->H, img = W * H;
  int rc = ensure_streams(c); if (rc) return rc;
  const int s = (int)(c->enqueued & 1);
  rc = ensure(c, c->pipeL[s], img * c->maxBatch); if (rc) return rc;
  rc = ensure(c, c->pipeR[s], img * c->maxBatch); if (rc) return rc;
  rc = ensure(c, c->pipeD[s], img * 2 * c->maxBatch); if (rc) return rc;
  const bool reused = c->enqueued >= 2;  // slot s was used by the enqueue before the previous one
  // copy-in: the kernels that read this staging slot last must be done
  if (reused) CU(c, cudaStreamWaitEvent(c->stIn, c->evComp[s], 0));
  CU(c, cudaMemcpyAsync(c->pipeL[s].p, left, img * n, cudaMemcpyHostToDevice, c->stIn));
  CU(c, cudaMemcpyAsync(c->pipeR[s].p, right, img * n, cudaMemcpyHostToDevice, c->stIn));
  CU(c, cudaEventRecord(c->evIn[s], c->stIn));
  // kernels: inputs are in, and the copy-out that read this slot's result buffer last is done
  CU(c, cudaStreamWaitEvent(c->st, c->evIn[s], 0));
  if (reused) CU(c, cudaStreamWaitEvent(c->st, c->evOut[s], 0));
  rc = run_sgbm(c, n, (uint8_t*)c->pipeL[s].p, W, img, (uint8_t*)c->pipeR[s].p, W, img, (int16_t*)c->pipeD[s].p, W * 2, img * 2);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->evComp[s], c->st));
  // copy-out
  CU(c, cudaStreamWaitEvent(c->stOut, c->evComp[s], 0));
  CU(c, cudaMemcpyAsync(disp, c->pipeD[s].p, img * 2 * n, cudaMemcpyDeviceToHost, c->stOut));
  CU(c, cudaEventRecord(c->evOut[s], c->stOut));
  c->enqueued++;
  return SV_OK;
}

SVGPU_API int svgpu_bm_compute(svgpu_ctx* c, const uint8_t* left, size_t ls, const uint8_t* right, size_t rs, int16_t* disp,
                     size_t dstride, int mem) {
  if (!c) return SV_INVALID_SETTINGS;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!left || !right || !disp) return fail(c, SV_INVALID_IMAGE, "empty image");
  if (!c->haveBm) return fail(c, SV_NOT_CONFIGURED, "bm parameters not set");
  cudaSetDevice(c->device);
  const size_t W = c->W, H = c->H;
  int rc = ensure(c, c->stBmL, W * H); if (rc) return rc;
  rc = ensure(c, c->stBmR, W * H); if (rc) return rc;
  const uint8_t *dl = left, *dr = right; int16_t* dd = disp; size_t dls = ls, drs = rs, dds = dstride;
  if (mem == SV_MEM_HOST) {
    rc = ensure(c, c->stLeft, W * H); if (rc) return rc;
    rc = ensure(c, c->stRight, W * H); if (rc) return rc;
    rc = ensure(c, c->stDisp, W * H * 2); if (rc) return rc;
    rc = h2d_2d(c, c->stLeft.p, W, left, ls, W, (int)H); if (rc) return rc;
    rc = h2d_2d(c, c->stRight.p, W, right, rs, W, (int)H); if (rc) return rc;
    dl = (uint8_t*)c->stLeft.p; dr = (uint8_t*)c->stRight.p; dd = (int16_t*)c->stDisp.p; dls = drs = W; dds = W * 2;
  }
  int nl = 0;
  cudaError_t e = bm_forward(c->bm, dl, dls, dr, drs, (uint8_t*)c->stBmL.p, (uint8_t*)c->stBmR.p, dd, dds, c->st, &nl);
  c->stats.launches += nl;
  if (e != cudaSuccess) return cuda_fail(c, e, "bm_forward");
  if (c->bm.speckleWindowSize > 0) {
    rc = ensure(c, c->stLabels, W * H * 4); if (rc) return rc;
    rc = ensure(c, c->stSizes, W * H * 4); if (rc) return rc;
    nl = 0;
    e = launch_filter_speckles(dd, dds, c->W, c->H, c->bm.filtered, c->bm.speckleWindowSize, c->bm.speckleRange,
                               (int*)c->stLabels.p, (int*)c->stSizes.p, nullptr, c->st, &nl);
    c->stats.launches += nl;
    if (e != cudaSuccess) return cuda_fail(c, e, "filter_speckles");
  }
  if (mem == SV_MEM_HOST) {
    rc = d2h_2d(c, disp, dstride, dd, dds, W * 2, (int)H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

// helper for the single-image int16 -> X stages below
#define SV_STAGE_PROLOGUE()                                              \
  if (!c) return SV_INVALID_SETTINGS;                                    \
  std::lock_guard<std::mutex> lk(c->mu);                                 \
  cudaSetDevice(c->device);                                              \
  const size_t W = c->W, H = c->H;                                       \
  (void)W; (void)H;

SVGPU_API int svgpu_median3(svgpu_ctx* c, const int16_t* src, size_t ss, int16_t* dst, size_t dstride, int mem) {
  SV_STAGE_PROLOGUE();
  if (!src || !dst) return fail(c, SV_INVALID_IMAGE, "empty image");
  const int16_t* s = src; int16_t* d = dst; size_t sss = ss, dds = dstride;
  if (mem == SV_MEM_HOST) {
    int rc = ensure(c, c->stDisp, W * H * 2); if (rc) return rc;
    rc = ensure(c, c->stDisp2, W * H * 2); if (rc) return rc;
    rc = h2d_2d(c, c->stDisp.p, W * 2, src, ss, W * 2, (int)H); if (rc) return rc;
    s = (int16_t*)c->stDisp.p; d = (int16_t*)c->stDisp2.p; sss = dds = W * 2;
  }
  CU(c, launch_median3(s, sss, d, dds, c->W, c->H, 1, 0, 0, c->st));
  c->stats.launches++;
  if (mem == SV_MEM_HOST) {
    int rc = d2h_2d(c, dst, dstride, d, dds, W * 2, (int)H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

SVGPU_API int svgpu_filter_speckles(svgpu_ctx* c, int16_t* disp, size_t stride, int newVal, int maxSize, int maxDiff, int mem) {
  SV_STAGE_PROLOGUE();
  if (!disp) return fail(c, SV_INVALID_IMAGE, "empty image");
  int rc = ensure(c, c->stLabels, W * H * 4); if (rc) return rc;
  rc = ensure(c, c->stSizes, W * H * 4); if (rc) return rc;
  int16_t* d = disp; size_t ds = stride;
  if (mem == SV_MEM_HOST) {
    rc = ensure(c, c->stDisp, W * H * 2); if (rc) return rc;
    rc = h2d_2d(c, c->stDisp.p, W * 2, disp, stride, W * 2, (int)H); if (rc) return rc;
    d = (int16_t*)c->stDisp.p; ds = W * 2;
  }
  int nl = 0;
  cudaError_t e = launch_filter_speckles(d, ds, c->W, c->H, newVal, maxSize, maxDiff, (int*)c->stLabels.p, (int*)c->stSizes.p, nullptr, c->st, &nl);
  c->stats.launches += nl;
  if (e != cudaSuccess) return cuda_fail(c, e, "filter_speckles");
  if (mem == SV_MEM_HOST) {
    rc = d2h_2d(c, disp, stride, d, ds, W * 2, (int)H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

SVGPU_API int svgpu_morph_close5(svgpu_ctx* c, const int16_t* src, size_t ss, int16_t* dst, size_t dstride, int mem) {
  SV_STAGE_PROLOGUE();
  if (!src || !dst) return fail(c, SV_INVALID_IMAGE, "empty image");
  int rc = ensure(c, c->stLabels, W * H * 4); if (rc) return rc;  // tmp plane
  const int16_t* s = src; int16_t* d = dst; size_t sss = ss, dds = dstride;
  if (mem == SV_MEM_HOST) {
    rc = ensure(c, c->stDisp, W * H * 2); if (rc) return rc;
    rc = ensure(c, c->stDisp2, W * H * 2); if (rc) return rc;
    rc = h2d_2d(c, c->stDisp.p, W * 2, src, ss, W * 2, (int)H); if (rc) return rc;
    s = (int16_t*)c->stDisp.p; d = (int16_t*)c->stDisp2.p; sss = dds = W * 2;
  }
  CU(c, launch_morph_close5(s, sss, (int16_t*)c->stLabels.p, d, dds, c->W, c->H, c->st));
  c->stats.launches += 1;
  if (mem == SV_MEM_HOST) {
    rc = d2h_2d(c, dst, dstride, d, dds, W * 2, (int)H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

SVGPU_API int svgpu_reproject(svgpu_ctx* c, const int16_t* disp, size_t ds, float* xyz, size_t xs, int handleMissing, int mem) {
  SV_STAGE_PROLOGUE();
  if (!disp || !xyz) return fail(c, SV_INVALID_IMAGE, "empty image");
  if (!c->haveQ) return fail(c, SV_NOT_CONFIGURED, "Q not set");
  const int16_t* d = disp; float* o = xyz; size_t dds = ds, oos = xs;
  if (mem == SV_MEM_HOST) {
    int rc = ensure(c, c->stDisp, W * H * 2); if (rc) return rc;
    rc = ensure(c, c->stXyz, W * H * 12); if (rc) return rc;
    rc = h2d_2d(c, c->stDisp.p, W * 2, disp, ds, W * 2, (int)H); if (rc) return rc;
    d = (int16_t*)c->stDisp.p; o = (float*)c->stXyz.p; dds = W * 2; oos = W * 12;
  }
  int nl = 0;
  cudaError_t e = launch_reproject(d, dds, o, oos, c->W, c->H, c->Qdev, handleMissing, c->minScratch, c->st, &nl);
  c->stats.launches += nl;
  if (e != cudaSuccess) return cuda_fail(c, e, "reproject");
  if (mem == SV_MEM_HOST) {
    int rc = d2h_2d(c, xyz, xs, o, oos, W * 12, (int)H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

SVGPU_API int svgpu_depth(svgpu_ctx* c, const int16_t* disp, size_t ds, float* depth, size_t zs, int mem) {
  SV_STAGE_PROLOGUE();
  if (!disp || !depth) return fail(c, SV_INVALID_IMAGE, "empty image");
  if (!c->haveDepth) return fail(c, SV_NOT_CONFIGURED, "depth constants not set");
  const int16_t* d = disp; float* o = depth; size_t dds = ds, oos = zs;
  if (mem == SV_MEM_HOST) {
    int rc = ensure(c, c->stDisp, W * H * 2); if (rc) return rc;
    rc = ensure(c, c->stDepth, W * H * 4); if (rc) return rc;
    rc = h2d_2d(c, c->stDisp.p, W * 2, disp, ds, W * 2, (int)H); if (rc) return rc;
    d = (int16_t*)c->stDisp.p; o = (float*)c->stDepth.p; dds = W * 2; oos = W * 4;
  }
  CU(c, launch_depth(d, dds, o, oos, c->W, c->H, c->baseline * c->focal, c->st));
  c->stats.launches++;
  if (mem == SV_MEM_HOST) {
    int rc = d2h_2d(c, depth, zs, o, oos, W * 4, (int)H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

SVGPU_API int svgpu_prefilter_chain(svgpu_ctx* c, const uint8_t* bgr, size_t bs, uint8_t* gray, size_t gs, int mem) {
  SV_STAGE_PROLOGUE();
  if (!bgr || !gray) return fail(c, SV_INVALID_IMAGE, "empty image");
  int rc = ensure(c, c->stTmpH, W * H * 6); if (rc) return rc;
  rc = ensure(c, c->stTmpG, W * H); if (rc) return rc;
  const uint8_t* s = bgr; uint8_t* d = gray; size_t sss = bs, dds = gs;
  if (mem == SV_MEM_HOST) {
    rc = ensure(c, c->stRectL, W * H * 3); if (rc) return rc;
    rc = ensure(c, c->stGrayL, W * H); if (rc) return rc;
    rc = h2d_2d(c, c->stRectL.p, W * 3, bgr, bs, W * 3, (int)H); if (rc) return rc;
    s = (uint8_t*)c->stRectL.p; d = (uint8_t*)c->stGrayL.p; sss = W * 3; dds = W;
  }
  CU(c, launch_prefilter_chain(s, sss, (uint16_t*)c->stTmpH.p, (uint8_t*)c->stTmpG.p, d, dds, c->W, c->H, c->st));
  c->stats.launches += 3;
  if (mem == SV_MEM_HOST) {
    rc = d2h_2d(c, gray, gs, d, dds, W, (int)H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

SVGPU_API int svgpu_resize(svgpu_ctx* c, const uint8_t* src, int sw, int sh, size_t ss, int cn, uint8_t* dst, size_t dstride, int mem) {
  SV_STAGE_PROLOGUE();
  if (!src || !dst || sw <= 0 || sh <= 0) return fail(c, SV_INVALID_IMAGE, "empty image");
  if (cn != 1 && cn != 3) return fail(c, SV_INVALID_SETTINGS, "channels must be 1 or 3");
  const uint8_t* s = src; uint8_t* d = dst; size_t sss = ss, dds = dstride;
  if (mem == SV_MEM_HOST) {
    int rc = ensure(c, c->stSrc, (size_t)sw * sh * cn); if (rc) return rc;
    rc = ensure(c, c->stRectL, W * H * cn); if (rc) return rc;
    rc = h2d_2d(c, c->stSrc.p, (size_t)sw * cn, src, ss, (size_t)sw * cn, sh); if (rc) return rc;
    s = (uint8_t*)c->stSrc.p; d = (uint8_t*)c->stRectL.p; sss = (size_t)sw * cn; dds = W * cn;
  }
  CU(c, launch_resize(s, sw, sh, sss, cn, d, c->W, c->H, dds, c->st));
  c->stats.launches++;
  if (mem == SV_MEM_HOST) {
    int rc = d2h_2d(c, dst, dstride, d, dds, W * cn, (int)H); if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->st));
  }
  return SV_OK;
}

SVGPU_API int svgpu_pipeline(svgpu_ctx* c, const svgpu_pipeline_io* io, unsigned stages) {
  SV_STAGE_PROLOGUE();
  if (!io || !io->left || !io->right) return fail(c, SV_INVALID_IMAGE, "empty image");
  const int cn = io->channels;
  int rc = pipeline_check(c, cn, stages, io->disp != nullptr);
  if (rc) return rc;
  const bool host = io->mem == SV_MEM_HOST;
  PipeDev p{};
  p.cn = cn;
  p.left = io->left; p.right = io->right; p.ls = io->left_stride; p.rs = io->right_stride;
  if (host) {
    rc = ensure(c, c->stLeft, W * H * cn); if (rc) return rc;
    rc = ensure(c, c->stRight, W * H * cn); if (rc) return rc;
    rc = h2d_2d(c, c->stLeft.p, W * cn, io->left, io->left_stride, W * cn, (int)H); if (rc) return rc;
    rc = h2d_2d(c, c->stRight.p, W * cn, io->right, io->right_stride, W * cn, (int)H); if (rc) return rc;
    p.left = (uint8_t*)c->stLeft.p; p.right = (uint8_t*)c->stRight.p; p.ls = p.rs = W * cn;
  }
  if (!host && io->left_rect && io->right_rect) { p.rectL = io->left_rect; p.rectR = io->right_rect; p.rectS = io->rect_stride; }
  p.wantRect = io->left_rect || io->right_rect;
  p.disp = io->disp; p.dS = io->disp_stride;
  if (host || !p.disp) {
    rc = ensure(c, c->stDisp, W * H * 2); if (rc) return rc;
    if (host && io->disp && !(stages & SV_STAGE_SGBM)) { rc = h2d_2d(c, c->stDisp.p, W * 2, io->disp, io->disp_stride, W * 2, (int)H); if (rc) return rc; }
    p.disp = (int16_t*)c->stDisp.p; p.dS = W * 2;
  }
  if ((stages & SV_STAGE_REPROJECT) && io->xyz) {
    p.xyz = io->xyz; p.xS = io->xyz_stride;
    if (host) { rc = ensure(c, c->stXyz, W * H * 12); if (rc) return rc; p.xyz = (float*)c->stXyz.p; p.xS = W * 12; }
  }
  if ((stages & SV_STAGE_DEPTH) && io->depth) {
    p.depth = io->depth; p.zS = io->depth_stride;
    if (host) { rc = ensure(c, c->stDepth, W * H * 4); if (rc) return rc; p.depth = (float*)c->stDepth.p; p.zS = W * 4; }
  }
  rc = pipeline_device(c, 1, p, stages);
  if (rc) return rc;
  if (!host) {
    // device buffers: with only one of the two rectified outputs requested the pair was rectified into the context's scratch
    if ((stages & SV_STAGE_REMAP) && (io->left_rect != nullptr) != (io->right_rect != nullptr)) {
      uint8_t* dst = io->left_rect ? io->left_rect : io->right_rect;
      CU(c, cudaMemcpy2DAsync(dst, io->rect_stride, io->left_rect ? p.rectL : p.rectR, p.rectS, W * cn, H, cudaMemcpyDeviceToDevice, c->st));
    }
    return SV_OK;
  }
  if (stages & SV_STAGE_REMAP) {
    if (io->left_rect) { rc = d2h_2d(c, io->left_rect, io->rect_stride, p.rectL, p.rectS, W * cn, (int)H); if (rc) return rc; }
    if (io->right_rect) { rc = d2h_2d(c, io->right_rect, io->rect_stride, p.rectR, p.rectS, W * cn, (int)H); if (rc) return rc; }
  }
  if (io->disp && (stages & (SV_STAGE_SGBM | SV_STAGE_POSTFILTER))) {
    rc = d2h_2d(c, io->disp, io->disp_stride, p.disp, p.dS, W * 2, (int)H); if (rc) return rc;
  }
  if (p.xyz) { rc = d2h_2d(c, io->xyz, io->xyz_stride, p.xyz, p.xS, W * 12, (int)H); if (rc) return rc; }
  if (p.depth) { rc = d2h_2d(c, io->depth, io->depth_stride, p.depth, p.zS, W * 4, (int)H); if (rc) return rc; }
  return check_status(c);
}

SVGPU_API int svgpu_pipeline_batch(svgpu_ctx* c, int n, const svgpu_pipeline_batch_io* io, unsigned stages) {
  SV_STAGE_PROLOGUE();
  if (!io || !io->left || !io->right || n <= 0) return fail(c, SV_INVALID_IMAGE, "empty batch");
  const int cn = io->channels;
  int rc = pipeline_check(c, cn, stages, io->disp != nullptr);
  if (rc) return rc;
  if (!io->disp) return fail(c, SV_INVALID_IMAGE, "the batch calls need a disparity buffer");
  const bool host = io->mem == SV_MEM_HOST;
  const size_t px = W * H, img = px * cn;
  const int chunk = c->maxBatch;
  const bool wantRect = (stages & SV_STAGE_REMAP) && io->left_rect && io->right_rect;
  const bool wantXyz = (stages & SV_STAGE_REPROJECT) && io->xyz, wantDepth = (stages & SV_STAGE_DEPTH) && io->depth;
  if (host) {
    rc = ensure(c, c->stLeft, img * chunk); if (rc) return rc;
    rc = ensure(c, c->stRight, img * chunk); if (rc) return rc;
    rc = ensure(c, c->stDisp, px * 2 * chunk); if (rc) return rc;
    if (wantXyz) { rc = ensure(c, c->stXyz, px * 12 * chunk); if (rc) return rc; }
    if (wantDepth) { rc = ensure(c, c->stDepth, px * 4 * chunk); if (rc) return rc; }
  }
  for (int f0 = 0; f0 < n; f0 += chunk) {
    const int m = std::min(chunk, n - f0);
    PipeDev p{};
    p.cn = cn; p.ls = p.rs = W * cn; p.lf = p.rf = img;
    p.wantRect = wantRect;
    p.dS = W * 2; p.dF = px * 2; p.xS = W * 12; p.xF = px * 12; p.zS = W * 4; p.zF = px * 4;
    if (host) {
      CU(c, cudaMemcpyAsync(c->stLeft.p, io->left + f0 * img, img * m, cudaMemcpyHostToDevice, c->st));
      CU(c, cudaMemcpyAsync(c->stRight.p, io->right + f0 * img, img * m, cudaMemcpyHostToDevice, c->st));
      if (!(stages & SV_STAGE_SGBM)) CU(c, cudaMemcpyAsync(c->stDisp.p, io->disp + f0 * px, px * 2 * m, cudaMemcpyHostToDevice, c->st));
      p.left = (uint8_t*)c->stLeft.p; p.right = (uint8_t*)c->stRight.p; p.disp = (int16_t*)c->stDisp.p;
      if (wantXyz) p.xyz = (float*)c->stXyz.p;
      if (wantDepth) p.depth = (float*)c->stDepth.p;
    } else {
      p.left = io->left + f0 * img; p.right = io->right + f0 * img; p.disp = io->disp + f0 * px;
      if (wantRect) { p.rectL = io->left_rect + f0 * img; p.rectR = io->right_rect + f0 * img; p.rectS = W * cn; p.rectF = img; }
      if (wantXyz) p.xyz = io->xyz + f0 * px * 3;
      if (wantDepth) p.depth = io->depth + f0 * px;
    }
    rc = pipeline_device(c, m, p, stages);
    if (rc) return rc;
    if (host) {
      if (wantRect) {
        CU(c, cudaMemcpyAsync(io->left_rect + f0 * img, p.rectL, img * m, cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaMemcpyAsync(io->right_rect + f0 * img, p.rectR, img * m, cudaMemcpyDeviceToHost, c->st));
      }
      if (stages & (SV_STAGE_SGBM | SV_STAGE_POSTFILTER)) CU(c, cudaMemcpyAsync(io->disp + f0 * px, p.disp, px * 2 * m, cudaMemcpyDeviceToHost, c->st));
      if (wantXyz) CU(c, cudaMemcpyAsync(io->xyz + f0 * px * 3, p.xyz, px * 12 * m, cudaMemcpyDeviceToHost, c->st));
      if (wantDepth) CU(c, cudaMemcpyAsync(io->depth + f0 * px, p.depth, px * 4 * m, cudaMemcpyDeviceToHost, c->st));
    }
  }
  if (host) return check_status(c);
  return SV_OK;
}

SVGPU_API int svgpu_pipeline_enqueue_batch(svgpu_ctx* c, int n, const svgpu_pipeline_batch_io* io, unsigned stages) {
  SV_STAGE_PROLOGUE();
  if (!io || !io->left || !io->right || n <= 0) return fail(c, SV_INVALID_IMAGE, "empty batch");
  if (io->mem != SV_MEM_HOST) return fail(c, SV_INVALID_SETTINGS, "pipeline_enqueue_batch takes host buffers");
  if (n > c->maxBatch) return fail(c, SV_INVALID_SETTINGS, "pipeline_enqueue_batch: n exceeds svgpu_reserve()");
  const int cn = io->channels;
  int rc = pipeline_check(c, cn, stages, io->disp != nullptr);
  if (rc) return rc;
  if (!io->disp) return fail(c, SV_INVALID_IMAGE, "the batch calls need a disparity buffer");
  rc = ensure_streams(c); if (rc) return rc;
  const size_t px = W * H, img = px * cn, cap = (size_t)c->maxBatch;
  const bool wantRect = (stages & SV_STAGE_REMAP) && io->left_rect && io->right_rect;
  const bool wantXyz = (stages & SV_STAGE_REPROJECT) && io->xyz, wantDepth = (stages & SV_STAGE_DEPTH) && io->depth;
  const int s = (int)(c->enqueued & 1);
  rc = ensure(c, c->pipeL[s], img * cap); if (rc) return rc;
  rc = ensure(c, c->pipeR[s], img * cap); if (rc) return rc;
  rc = ensure(c, c->pipeD[s], px * 2 * cap); if (rc) return rc;
  if (wantRect) { rc = ensure(c, c->pipeRL[s], img * cap); if (rc) return rc; rc = ensure(c, c->pipeRR[s], img * cap); if (rc) return rc; }
  if (wantXyz) { rc = ensure(c, c->pipeX[s], px * 12 * cap); if (rc) return rc; }
  if (wantDepth) { rc = ensure(c, c->pipeZ[s], px * 4 * cap); if (rc) return rc; }
  const bool reused = c->enqueued >= 2;  // slot s was used by the enqueue before the previous one
  if (reused) CU(c, cudaStreamWaitEvent(c->stIn, c->evComp[s], 0));
  CU(c, cudaMemcpyAsync(c->pipeL[s].p, io->left, img * n, cudaMemcpyHostToDevice, c->stIn));
  CU(c, cudaMemcpyAsync(c->pipeR[s].p, io->right, img * n, cudaMemcpyHostToDevice, c->stIn));
  if (!(stages & SV_STAGE_SGBM)) CU(c, cudaMemcpyAsync(c->pipeD[s].p, io->disp, px * 2 * n, cudaMemcpyHostToDevice, c->stIn));
  CU(c, cudaEventRecord(c->evIn[s], c->stIn));
  CU(c, cudaStreamWaitEvent(c->st, c->evIn[s], 0));
  if (reused) CU(c, cudaStreamWaitEvent(c->st, c->evOut[s], 0));
  PipeDev p{};
  p.cn = cn; p.ls = p.rs = W * cn; p.lf = p.rf = img;
  p.wantRect = wantRect;
  p.left = (uint8_t*)c->pipeL[s].p; p.right = (uint8_t*)c->pipeR[s].p;
  p.disp = (int16_t*)c->pipeD[s].p; p.dS = W * 2; p.dF = px * 2;
  if (wantRect) { p.rectL = (uint8_t*)c->pipeRL[s].p; p.rectR = (uint8_t*)c->pipeRR[s].p; p.rectS = W * cn; p.rectF = img; }
  if (wantXyz) { p.xyz = (float*)c->pipeX[s].p; p.xS = W * 12; p.xF = px * 12; }
  if (wantDepth) { p.depth = (float*)c->pipeZ[s].p; p.zS = W * 4; p.zF = px * 4; }
  rc = pipeline_device(c, n, p, stages);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->evComp[s], c->st));
  CU(c, cudaStreamWaitEvent(c->stOut, c->evComp[s], 0));
  if (wantRect) {
    CU(c, cudaMemcpyAsync(io->left_rect, p.rectL, img * n, cudaMemcpyDeviceToHost, c->stOut));
    CU(c, cudaMemcpyAsync(io->right_rect, p.rectR, img * n, cudaMemcpyDeviceToHost, c->stOut));
  }
  if (stages & (SV_STAGE_SGBM | SV_STAGE_POSTFILTER)) CU(c, cudaMemcpyAsync(io->disp, p.disp, px * 2 * n, cudaMemcpyDeviceToHost, c->stOut));
  if (wantXyz) CU(c, cudaMemcpyAsync(io->xyz, p.xyz, px * 12 * n, cudaMemcpyDeviceToHost, c->stOut));
  if (wantDepth) CU(c, cudaMemcpyAsync(io->depth, p.depth, px * 4 * n, cudaMemcpyDeviceToHost, c->stOut));
  CU(c, cudaEventRecord(c->evOut[s], c->stOut));
  c->enqueued++;
  return SV_OK;
}

SVGPU_API int svgpu_debug_copy_volume(svgpu_ctx* c, int which, int16_t* host, size_t bytes) {
  SV_STAGE_PROLOGUE();
  if (!host || !c->sc.C) return fail(c, SV_NOT_CONFIGURED, "no volume");
  const size_t n = std::min(bytes, c->sc.frameCells * 2) / 2;
  CU(c, cudaStreamSynchronize(c->st));
  std::vector<int16_t> tmp(n);
  CU(c, cudaMemcpy(tmp.data(), which == 0 ? c->sc.C : c->sc.S, n * 2, cudaMemcpyDeviceToHost));
  // undo the interleave inside each group of 16 disparities (sgbm_device.cuh: vol_index)
  const size_t D = (size_t)c->cfg.D;
  for (size_t p = 0; p + D <= n; p += D)
    for (size_t d = 0; d < D; d++) host[p + d] = tmp[p + (size_t)vol_index((int)d)];
  return SV_OK;
}

}  // extern "C"

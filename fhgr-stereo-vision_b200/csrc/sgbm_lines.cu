// sgbm_lines.cu — SGM path aggregation for ONE (or very few) pairs per call: every path direction walks its own lines.
//
// Replaces the aggregation + winner-take-all inside cv::StereoSGBM::compute (/root/reference/src/lib.cpp:262; arithmetic:
// SURVEY.md A.5 / A.6) when a call is too small to fill the machine with the strip pipeline of sgbm_sweep.cu.
//
// The strip pipeline touches every volume cell twice (8 B per cell) but a call lasts as long as its wavefront:
// H + (W1 + H) / K dependent row steps per pass.  A single path, however, only depends on its own line: the rows of the
// image for the two horizontal directions, the columns for the vertical ones, the diagonals for the other four.  Here a
// worker (LPP lanes x 16 disparities, as in the sweep) walks ONE line of ONE direction and stores that path's costs; all
// lines of all directions (17,000 at 1080p MODE_HH) run at once and the longest chain is max(W1, H) steps of one path.
// k_wta_lines then adds the directions' volumes with the saturation of A.5 (min(a + b, M) folds in any order) and selects.
// Cost: 3 x paths x 2 B per cell of traffic instead of 8 B, which is why this is the path of calls whose volumes are small
// (sgbm_lines_ok): 1080p / 128 MODE_HH moves 12.7 GB per pair, the reference's 640 x 480 / 224 MODE_HH4 stays in L2.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "sgbm_device.cuh"

namespace svgpu {

namespace {

constexpr int kMaxDirs = 8;
#ifndef SV_LINES_PF
#define SV_LINES_PF 4
#endif
#ifndef SV_LINES_L2AHEAD
#define SV_LINES_L2AHEAD 8
#endif
// Pixels of C ahead of the walk: kPF in registers (covers an L2 hit) and a prefetch into L2 kL2Ahead pixels ahead (covers the
// DRAM round trip).  Measured (profiles/r02_summary.md): the L2 prefetch gains 15-30 %, 8 ahead is as good as 16 and beats 32.
// A deeper register ring gained nothing while the unrolled body held a branch: the SASS then carried the ring through copies
// placed right behind the loads, so every step waited for its newest load.  Without the branch the loads of an iteration are
// issued as one batch into the registers they are read from (-12 % at 720p / 64 and 640 x 480 / 224).  A cp.async ring in shared
// memory (4-16 pixels deep), 256-byte L2 fetch granules, burst prefetches, prefetch.global.L1 and volatile-asm loads were all
// equal or slower.
constexpr int kPF = SV_LINES_PF;
constexpr int kL2Ahead = SV_LINES_L2AHEAD;

struct LineArgs {
  const uint4* C;
  uint4* L;            // [dir][frame] volumes in the layout of C
  uint2* sel;
  int W1, H, D, nAct, nframes, ndirs;
  unsigned frameU4;    // uint4 per frame of a volume
  size_t dirU4;        // uint4 per direction = nframes * frameU4
  int pixU4;
  uint32_t p1dup, p2mp1dup, negp1dup, one;
  int dx[kMaxDirs], dy[kMaxDirs];
  int lineBase[kMaxDirs + 1];   // first line of a direction among the lines of one frame
  int uniqF;
  uint32_t uniqMagic;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int LPP>
__global__ void __launch_bounds__(256, 2) k_lines(const LineArgs a) {
  constexpr int RW = 32 / LPP, NV = kNR / 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane / LPP, gl = lane % LPP;
  const bool active = gl < a.nAct;
  const int lastLane = a.nAct - 1;
  SweepK kc;
  kc.p1dup = a.p1dup; kc.p2mp1dup = a.p2mp1dup; kc.negp1dup = a.negp1dup;
  kc.keepUp = opaque_u32(gl == 0 ? 0u : 1u); kc.addUp = opaque_u32(gl == 0 ? kMax2 : 0u);
  kc.keepDn = opaque_u32(gl == lastLane ? 0u : 1u); kc.addDn = opaque_u32(gl == lastLane ? kMax2 : 0u);
  kc.one = a.one;

  const int linesPerFrame = a.lineBase[a.ndirs];
  const long long total = (long long)a.nframes * linesPerFrame;
  const long long gw = ((long long)blockIdx.x * 8 + warp) * RW + j;
  const bool valid = gw < total;
  int len = 0;
  unsigned dOff = 0;   // step between the pixels of the line in uint4, modulo 2^32 (negative steps wrap)
  unsigned off = 0;
  const uint4* Cp = a.C;
  uint4* Lp = a.L;
  if (valid) {
    const int frame = (int)(gw / linesPerFrame), rem = (int)(gw - (long long)frame * linesPerFrame);
    int k = 0;
    while (k + 1 < a.ndirs && rem >= a.lineBase[k + 1]) k++;
    const int i = rem - a.lineBase[k], dx = a.dx[k], dy = a.dy[k];
    const int W1 = a.W1, H = a.H;
    int x0, y0;
    if (dy == 0) { x0 = dx > 0 ? 0 : W1 - 1; y0 = i; len = W1; }
    else if (dx == 0) { x0 = i; y0 = dy > 0 ? 0 : H - 1; len = H; }
    else if (i < W1) {   // diagonals that start on the top / bottom row ...
      x0 = i; y0 = dy > 0 ? 0 : H - 1;
      len = min(dx > 0 ? W1 - x0 : x0 + 1, H);
    } else {             // ... and on the left / right column (the corner pixel belongs to the first group)
      const int jj = i - W1 + 1;
      x0 = dx > 0 ? 0 : W1 - 1; y0 = dy > 0 ? jj : H - 1 - jj;
      len = min(W1, dy > 0 ? H - y0 : y0 + 1);
    }
    off = ((unsigned)y0 * (unsigned)W1 + (unsigned)x0) * (unsigned)a.pixU4 + (unsigned)((active ? gl : 0) * NV);
    dOff = (unsigned)((dy * W1 + dx) * a.pixU4);
    Cp = a.C + (size_t)frame * a.frameU4;
    Lp = a.L + (size_t)k * a.dirU4 + (size_t)frame * a.frameU4;
  }
  const int maxLen = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)len);   // the workers of a warp walk in lock step

  // offsets are formed modulo 2^32 (unsigned arithmetic) before they meet a pointer: a volume has fewer than 2^32 uint4
  uint4 cb[kPF][NV];
#pragma unroll
  for (int q = 0; q < kPF; q++) {
    const unsigned o = q < len ? off + (unsigned)q * dOff : 0u;
#pragma unroll
    for (int h = 0; h < NV; h++) cb[q][h] = __ldg(Cp + o + h);
  }
  for (int q = kPF; q < kL2Ahead; q++)
    if (q < len) prefetch_l2(Cp + (off + (unsigned)q * dOff));
  uint32_t L[kNR];
#pragma unroll
  for (int q = 0; q < kNR; q++) L[q] = 0;
  uint32_t m = 0;
  for (int t0 = 0; t0 < maxLen; t0 += kPF) {
#pragma unroll
    for (int q = 0; q < kPF; q++) {
      const int t = t0 + q;   // up to kPF - 1 steps past the longest line of the warp run predicated off (no branch in the body)
      uint32_t c[kNR];
#pragma unroll
      for (int h = 0; h < NV; h++) { c[4 * h] = cb[q][h].x; c[4 * h + 1] = cb[q][h].y; c[4 * h + 2] = cb[q][h].z; c[4 * h + 3] = cb[q][h].w; }
      {
        // always load (from a safe address past the end of the line) so that the data lands in the ring registers directly
        const unsigned o = t + kPF < len ? off + (unsigned)(t + kPF) * dOff : 0u;
#pragma unroll
        for (int h = 0; h < NV; h++) cb[q][h] = __ldg(Cp + o + h);
        if (t + kL2Ahead < len) prefetch_l2(Cp + (off + (unsigned)(t + kL2Ahead) * dOff));   // this lane's 32 bytes of that pixel
      }
      const uint32_t lm = sweep_path_core<LPP>(L, m, t == 0, c, kc, gl, lastLane, active, lane);
      m = group_min_u32<LPP>(lm, lane);
      if (t < len && active) {
        uint4* dst = Lp + (off + (unsigned)t * dOff);
#pragma unroll
        for (int h = 0; h < NV; h++) dst[h] = make_uint4(L[4 * h], L[4 * h + 1], L[4 * h + 2], L[4 * h + 3]);
      }
    }
  }
}

// S = saturating sum of the directions' volumes, then the winner-take-all of A.6: one pixel per worker
template <int LPP, bool WAY3>
__global__ void __launch_bounds__(256) k_wta_lines(const LineArgs a) {
  constexpr int RW = 32 / LPP, NV = kNR / 4, VEC = LPP * NV;
  __shared__ uint4 sW[8 * RW * VEC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane / LPP, gl = lane % LPP;
  const bool active = gl < a.nAct;
  uint16_t* wta = (uint16_t*)(sW + (warp * RW + j) * VEC);
  const long long framePix = (long long)a.H * a.W1, npix = framePix * a.nframes;
  const long long stride = (long long)gridDim.x * 8 * RW;
  for (long long p0 = ((long long)blockIdx.x * 8 + warp) * RW; p0 < npix; p0 += stride) {   // warp-uniform bounds
    const long long p = p0 + j;
    const bool on = p < npix;
    uint32_t s[kNR];
#pragma unroll
    for (int q = 0; q < kNR; q++) s[q] = 0;
    if (on && active) {
      const long long f = p / framePix, rem = p - f * framePix;
      const uint4* src = a.L + (size_t)f * a.frameU4 + (size_t)rem * a.pixU4 + gl * NV;
      for (int k = 0; k < a.ndirs; k++, src += a.dirU4) {
#pragma unroll
        for (int h = 0; h < NV; h++) {
          const uint4 v = __ldcs(src + h);
          s[4 * h] = __viaddmin_u16x2(s[4 * h], v.x, kMax2); s[4 * h + 1] = __viaddmin_u16x2(s[4 * h + 1], v.y, kMax2);
          s[4 * h + 2] = __viaddmin_u16x2(s[4 * h + 2], v.z, kMax2); s[4 * h + 3] = __viaddmin_u16x2(s[4 * h + 3], v.w, kMax2);
        }
      }
    }
    uint2 o;
    sweep_wta<LPP, WAY3>(s, wta, gl, active, lane, a.D, a.uniqF, a.uniqMagic, o);
    if (on && gl == 0) a.sel[p] = o;
  }
}

template <int LPP>
cudaError_t launch_lines(const LineArgs& a, bool way3, int numSM, cudaStream_t st) {
  constexpr int RW = 32 / LPP;
  const long long workers = (long long)a.nframes * a.lineBase[a.ndirs];
  const long long blocks = (workers + 8 * RW - 1) / (8 * RW);
  k_lines<LPP><<<(unsigned)blocks, 256, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long npix = (long long)a.H * a.W1 * a.nframes;
  const int grid = (int)std::min<long long>((npix + 8 * RW - 1) / (8 * RW), (long long)numSM * 32);
  if (way3) k_wta_lines<LPP, true><<<grid, 256, 0, st>>>(a);   // MODE_SGBM_3WAY has its own selection rule (sgbm_device.cuh)
  else k_wta_lines<LPP, false><<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

int ndirs_of(int mode) { return mode == SV_MODE_HH ? 8 : mode == SV_MODE_HH4 ? 4 : mode == SV_MODE_SGBM_3WAY ? 3 : 5; }

}  // namespace

// The line form wins while its traffic (6 x paths bytes per cell instead of 8) takes less time than the wavefront of the strip
// pipeline lasts.  Measured on a B200 (profiles/r02_summary.md): the line kernels move ~6 TB/s, a row step of the 2-column
// strip pipeline takes ~1.65 us and a call of a few pairs lasts little longer than one pair's wavefront.  E.g. aggregation of
// 1080p / 128 MODE_HH: 2.0 instead of 4.2 ms for one pair, 3.8 instead of 4.9 for two, 5.6 instead of 5.1 for three;
// 720p / 64 MODE_SGBM: 0.5 instead of 3.1 ms (1), 2.4 instead of 3.9 (8), 3.4 instead of 4.1 (12).
// SVGPU_SWEEP_LINES=0|1 pins the choice (tests).
int sgbm_lines_max_frames(const SgbmCfg& cfg, size_t frameCells) {
  // (MODE_SGBM_3WAY: the frames are the four stripe frames of a pair, cfg.H their row count)
  if (cfg.W1 <= 0 || frameCells == 0) return 0;
  if (cfg.D > 256) return 1;   // the only form built for more than 16 lanes per pixel: a call is served pair by pair
  const double perFrame = (double)frameCells * 2.0 * 3.0 * ndirs_of(cfg.mode);   // bytes moved per pair
  const char* f = std::getenv("SVGPU_SWEEP_LINES");
  if (f && f[0] == '0') return 0;
  if (f && f[0] == '1') return (int)std::min(64.0, 12.0 * 4294967296.0 / perFrame);   // path volumes up to 16 GB
  const double waveSeconds = 1.65e-6 * (cfg.H + (cfg.W1 + cfg.H) / 2.0);
  return (int)std::min(64.0, 1.1 * waveSeconds * 6e12 / perFrame);
}
bool sgbm_lines_ok(const SgbmCfg& cfg, int nframes, size_t frameCells) {
  return nframes >= 1 && nframes <= sgbm_lines_max_frames(cfg, frameCells);
}
size_t sgbm_lines_frames(const SgbmCfg& cfg, int nframes) { return (size_t)ndirs_of(cfg.mode) * nframes; }

cudaError_t sgbm_aggregate_lines(const SgbmCfg& cfg, const SgbmScratch& sc, int nframes, cudaStream_t st, LaunchStats* stats) {
  int dev = 0, numSM = 0;
  cudaError_t e0 = cudaGetDevice(&dev);
  if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, dev);
  if (e0 != cudaSuccess) return e0;
  LineArgs a{};
  a.C = (const uint4*)sc.C; a.L = (uint4*)sc.S; a.sel = (uint2*)sc.sel;
  a.W1 = cfg.W1; a.H = cfg.H; a.D = cfg.D; a.nAct = cfg.D / kDPL; a.nframes = nframes;
  a.frameU4 = (unsigned)(sc.frameCells / 8);
  a.dirU4 = (size_t)nframes * a.frameU4;
  a.pixU4 = cfg.D / 8;
  a.p1dup = dup16((uint32_t)cfg.P1);
  a.p2mp1dup = dup16((uint32_t)(cfg.P2 - cfg.P1));
  a.negp1dup = dup16((uint32_t)(-cfg.P1) & 0xFFFFu);
  a.one = 1u;
  a.uniqF = 100 - cfg.uniq;
  a.uniqMagic = a.uniqF >= 2 ? (uint32_t)((0x100000000ull + a.uniqF - 1) / a.uniqF) : 0u;
  // walking directions (the predecessor of a pixel is one step back): the horizontal ones first, they are the longest
  static const int hh[8][2] = {{1, 0}, {-1, 0}, {1, 1}, {-1, 1}, {-1, -1}, {1, -1}, {0, 1}, {0, -1}};
  static const int hh4[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
  static const int sgbm5[5][2] = {{1, 0}, {-1, 0}, {1, 1}, {-1, 1}, {0, 1}};
  static const int way3[3][2] = {{1, 0}, {-1, 0}, {0, 1}};
  const int (*dirs)[2] = cfg.mode == SV_MODE_HH ? hh : cfg.mode == SV_MODE_HH4 ? hh4 : cfg.mode == SV_MODE_SGBM_3WAY ? way3 : sgbm5;
  a.ndirs = ndirs_of(cfg.mode);
  a.lineBase[0] = 0;
  for (int k = 0; k < a.ndirs; k++) {
    a.dx[k] = dirs[k][0]; a.dy[k] = dirs[k][1];
    const int n = a.dy[k] == 0 ? cfg.H : a.dx[k] == 0 ? cfg.W1 : cfg.W1 + cfg.H - 1;
    a.lineBase[k + 1] = a.lineBase[k] + n;
  }
  const int groups = cfg.D / kDPL;
  const int lpp = groups <= 1 ? 1 : groups <= 2 ? 2 : groups <= 4 ? 4 : groups <= 8 ? 8 : groups <= 16 ? 16 : 32;
  // more pairs than the path volumes in S hold (numDisparities > 256: always one): pair by pair on the same volumes
  const int per = std::max(1, std::min(nframes, sgbm_lines_max_frames(cfg, sc.frameCells)));
  cudaError_t e = cudaSuccess;
  for (int f0 = 0; f0 < nframes && e == cudaSuccess; f0 += per) {
    a.nframes = std::min(per, nframes - f0);
    a.dirU4 = (size_t)a.nframes * a.frameU4;
    a.C = (const uint4*)sc.C + (size_t)f0 * a.frameU4;
    a.sel = (uint2*)sc.sel + (size_t)f0 * cfg.H * cfg.W1;
    switch (lpp) {
      case 1: e = launch_lines<1>(a, cfg.mode == SV_MODE_SGBM_3WAY, numSM, st); break;
      case 2: e = launch_lines<2>(a, cfg.mode == SV_MODE_SGBM_3WAY, numSM, st); break;
      case 4: e = launch_lines<4>(a, cfg.mode == SV_MODE_SGBM_3WAY, numSM, st); break;
      case 8: e = launch_lines<8>(a, cfg.mode == SV_MODE_SGBM_3WAY, numSM, st); break;
      case 16: e = launch_lines<16>(a, cfg.mode == SV_MODE_SGBM_3WAY, numSM, st); break;
      default: e = launch_lines<32>(a, cfg.mode == SV_MODE_SGBM_3WAY, numSM, st); break;
    }
    if (stats) stats->launches += 2;
  }
  return e;
}

}  // namespace svgpu

// front_fused.cu — the front of the chain in one kernel: rectification (remap) -> GaussianBlur 11x11 -> BGR2GRAY -> 3x3 sharpen ->
// SGBM pre-filter (Sobel x, clipped) -> Birchfield-Tomasi operand planes.
//
// Replaces, for raw BGR frames, the calls at /root/reference/src/lib.cpp:87-94 (cv::remap), :231-246 (GaussianBlur, cvtColor,
// filter2D) and the pre-filter inside cv::StereoSGBM::compute (:262), whose results the unfused path passes through HBM five
// times (k_remap, k_gauss_h, k_gauss_v_gray, k_sharpen3, k_prefilter_planes).  Here a CTA produces the operand planes of a
// TW x TH tile from the raw frame and the rectification maps: the rectified pixels, the blurred rows, the gray and the sharpened
// tile only ever live in shared memory (the rectified images are written out only when the caller asks for them).
// Arithmetic = the unfused kernels' (image_ops.cu, sgbm.cu: SURVEY.md A.2, A.3, A.11, A.13), bit for bit.
//
// Halos: planes(x) need the pre-filtered pixel at x-1..x+1, those the sharpened image at x-2..x+2 / y-1..y+1, that the gray
// image one pixel further, that the 9 non-zero taps of the blur: rectified pixels at x-7..x+7, y-6..y+6.  Image borders: the blur
// and the sharpen filter mirror (BORDER_REFLECT_101), so a tile position outside the image is computed AT its mirrored image
// position — every stage is a symmetric filter of the mirrored extension, hence its value at a mirrored position equals the
// value at the position itself; the SGBM pre-filter replicates rows and pins the first / last column to ftzero (A.2).
#include "common.cuh"

namespace svgpu {

namespace {

constexpr int kFTW = 128, kFTH = 32;      // output tile
constexpr int kHX = 7, kHY = 6;           // halo of the rectified tile
constexpr int kRW = kFTW + 2 * kHX, kRH = kFTH + 2 * kHY;   // rectified tile  142 x 44  (x 3 bytes)
constexpr int kBW = kFTW + 6, kBH = kRH;                    // row-blurred     134 x 44  (x 3 u16)
constexpr int kGW = kBW, kGH = kFTH + 4;                    // gray            134 x 36
constexpr int kSW = kFTW + 4, kSH = kFTH + 2;               // sharpened       132 x 34
constexpr int kFrontThreads = 256;
constexpr int kRectPitch = kRW * 3 + 2, kGrayPitch = kGW + 2;
constexpr int kOffRect = kBH * kBW * 3 * 2, kOffGray = kOffRect + kRH * kRectPitch, kOffSharp = kOffGray + kGH * kGrayPitch;
constexpr int kFrontSmem = kOffSharp + kSH * kSW;   // 63.6 KB: three CTAs per SM

__device__ __forceinline__ int mirror101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

struct FrontArgs {
  const uint8_t* src[2];   // raw frames, left / right: [frame][H][W][3]
  size_t sstride[2], sframe[2];
  const int16_t* map1[2];
  const uint16_t* map2[2];
  uint8_t* rect[2];        // optional rectified output (nullptr: not written)
  size_t rstride, rframe;
  uint8_t* planes[2];      // [frame][H][2][Wp] uint2
  int W, H, Wp, ftzero;
};

__global__ void __launch_bounds__(kFrontThreads) k_front_fused(const FrontArgs a) {
  extern __shared__ __align__(16) uint8_t smemFront[];
  uint16_t (*sBlur)[kBW * 3] = reinterpret_cast<uint16_t (*)[kBW * 3]>(smemFront);
  uint8_t (*sRect)[kRectPitch] = reinterpret_cast<uint8_t (*)[kRectPitch]>(smemFront + kOffRect);
  uint8_t (*sGray)[kGrayPitch] = reinterpret_cast<uint8_t (*)[kGrayPitch]>(smemFront + kOffGray);
  uint8_t (*sSharp)[kSW] = reinterpret_cast<uint8_t (*)[kSW]>(smemFront + kOffSharp);
  const int side = blockIdx.z & 1, frame = blockIdx.z >> 1;
  const int x0 = blockIdx.x * kFTW, y0 = blockIdx.y * kFTH;
  const int W = a.W, H = a.H;
  const uint8_t* src = a.src[side] + (size_t)frame * a.sframe[side];
  const size_t sstride = a.sstride[side];
  const int16_t* map1 = a.map1[side];
  const uint16_t* map2 = a.map2[side];
  uint8_t* rect = a.rect[side] ? a.rect[side] + (size_t)frame * a.rframe : nullptr;

  // ---- stage 0: cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) at the mirrored position of every tile pixel (A.11)
  for (int i = threadIdx.x; i < kRW * kRH; i += kFrontThreads) {
    const int ty = i / kRW, tx = i - ty * kRW;
    const int x = mirror101(x0 - kHX + tx, W), y = mirror101(y0 - kHY + ty, H);
    const short2 m = ((const short2*)map1)[(size_t)y * W + x];
    const int f = map2[(size_t)y * W + x] & 1023, fx = f & 31, fy = f >> 5;
    const int sx = m.x, sy = m.y;
    const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    const bool xin0 = sx >= 0 && sx < W, xin1 = sx + 1 >= 0 && sx + 1 < W;
    const bool yin0 = sy >= 0 && sy < H, yin1 = sy + 1 >= 0 && sy + 1 < H;
    const uint8_t* r0 = src + (size_t)sy * sstride + (size_t)sx * 3;
    const uint8_t* r1 = r0 + sstride;
    uint8_t v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      int acc = 16384;
      if (yin0 && xin0) acc += w00 * r0[c];
      if (yin0 && xin1) acc += w01 * r0[3 + c];
      if (yin1 && xin0) acc += w10 * r1[c];
      if (yin1 && xin1) acc += w11 * r1[3 + c];
      v[c] = (uint8_t)min(acc >> 15, 255);
      sRect[ty][tx * 3 + c] = v[c];
    }
    if (rect && tx >= kHX && tx < kHX + kFTW && ty >= kHY && ty < kHY + kFTH && x0 - kHX + tx < W && y0 - kHY + ty < H) {
      uint8_t* o = rect + (size_t)y * a.rstride + (size_t)x * 3;   // inside the image the mirrored position is the position itself
      o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
  }
  __syncthreads();
  // ---- stage 1: horizontal pass of GaussianBlur 11x11 sigma 1.4 (taps 0 and 10 are zero), 8.8 fixed point (A.13)
  const int kG[9] = {1, 8, 26, 57, 72, 57, 26, 8, 1};
  for (int i = threadIdx.x; i < kBW * kBH; i += kFrontThreads) {
    const int ty = i / kBW, tx = i - ty * kBW;      // blurred column tx <-> rectified columns tx .. tx + 8
    int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const uint8_t* p = &sRect[ty][(tx + t) * 3];
      s0 += kG[t] * p[0]; s1 += kG[t] * p[1]; s2 += kG[t] * p[2];
    }
    sBlur[ty][tx * 3] = (uint16_t)s0; sBlur[ty][tx * 3 + 1] = (uint16_t)s1; sBlur[ty][tx * 3 + 2] = (uint16_t)s2;
  }
  __syncthreads();
  // ---- stage 2: vertical pass + cv::cvtColor(BGR2GRAY) (15-bit fixed point)
  for (int i = threadIdx.x; i < kGW * kGH; i += kFrontThreads) {
    const int ty = i / kGW, tx = i - ty * kGW;      // gray row ty <-> blurred rows ty .. ty + 8
    unsigned s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const uint16_t* p = &sBlur[ty + t][tx * 3];
      s0 += kG[t] * p[0]; s1 += kG[t] * p[1]; s2 += kG[t] * p[2];
    }
    const int b = (s0 + 32768u) >> 16, g = (s1 + 32768u) >> 16, r = (s2 + 32768u) >> 16;
    sGray[ty][tx] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15);
  }
  __syncthreads();
  // ---- stage 3: cv::filter2D with [0 -1 0; -1 5 -1; 0 -1 0]
  for (int i = threadIdx.x; i < kSW * kSH; i += kFrontThreads) {
    const int ty = i / kSW, tx = i - ty * kSW;      // sharpened (ty, tx) <-> gray (ty + 1, tx + 1)
    const int c = sGray[ty + 1][tx + 1];
    const int v = 5 * c - sGray[ty + 1][tx] - sGray[ty + 1][tx + 2] - sGray[ty][tx + 1] - sGray[ty + 2][tx + 1];
    sSharp[ty][tx] = (uint8_t)min(max(v, 0), 255);
  }
  __syncthreads();
  // ---- stage 4: pre-filter rows of StereoSGBM (A.2) + Birchfield-Tomasi half-pixel bounds (A.3) -> operand planes
  const int ftzero = a.ftzero;
  auto h16 = [](int v) { return (uint32_t)v & 0xFFFFu; };
  for (int i = threadIdx.x; i < kFTW * kFTH; i += kFrontThreads) {
    const int ty = i / kFTW, tx = i - ty * kFTW;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= W || y >= H) continue;
    // sharpened tile: row of image row y is ty + 1, column of image column x is tx + 2; rows replicate at the image border
    const uint8_t* rr = sSharp[ty + 1];
    const uint8_t* ru = sSharp[y > 0 ? ty : ty + 1];
    const uint8_t* rd = sSharp[y < H - 1 ? ty + 2 : ty + 1];
    int g[3], it[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int xx = x - 1 + k, cx = tx + 1 + k;
      const bool border = xx <= 0 || xx >= W - 1;
      const int v = 2 * ((int)rr[cx + 1] - (int)rr[cx - 1]) + ((int)ru[cx + 1] - (int)ru[cx - 1]) + ((int)rd[cx + 1] - (int)rd[cx - 1]);
      g[k] = border ? ftzero : min(max(v, -ftzero), ftzero) + ftzero;
      it[k] = border ? ftzero : (int)rr[cx];
    }
    const int g0 = g[1], i0 = it[1];
    int glo = g0, ghi = g0, ilo = i0, ihi = i0;
    if (x > 0) {
      const int p = (g0 + g[0]) >> 1, q = (i0 + it[0]) >> 1;
      glo = min(glo, p); ghi = max(ghi, p); ilo = min(ilo, q); ihi = max(ihi, q);
    }
    if (x < W - 1) {
      const int p = (g0 + g[2]) >> 1, q = (i0 + it[2]) >> 1;
      glo = min(glo, p); ghi = max(ghi, p); ilo = min(ilo, q); ihi = max(ihi, q);
    }
    uint2* p = (uint2*)a.planes[side] + ((size_t)frame * H + y) * 2 * a.Wp + x;
    if (side == 0) {
      p[0] = make_uint2(h16(glo) | h16(g0) << 16, h16(-ghi) | h16(-g0) << 16);
      p[a.Wp] = make_uint2(h16(ilo) | h16(i0) << 16, h16(-ihi) | h16(-i0) << 16);
    } else {
      p[0] = make_uint2(h16(-g0) | h16(-ghi) << 16, h16(g0) | h16(glo) << 16);
      p[a.Wp] = make_uint2(h16(-i0) | h16(-ihi) << 16, h16(i0) | h16(ilo) << 16);
    }
  }
}

}  // namespace

// n BGR pairs: raw frames -> operand planes of the SGBM cost kernel (+ optionally the rectified images)
cudaError_t launch_front_fused(const uint8_t* left, size_t lstride, size_t lframe, const uint8_t* right, size_t rstride, size_t rframe,
                               const int16_t* const map1[2], const uint16_t* const map2[2], uint8_t* rectL, uint8_t* rectR, size_t rectStride,
                               size_t rectFrame, uint8_t* planesL, uint8_t* planesR, int W, int H, int Wp, int ftzero, int nframes,
                               cudaStream_t st) {
  FrontArgs a;
  a.src[0] = left; a.src[1] = right; a.sstride[0] = lstride; a.sstride[1] = rstride; a.sframe[0] = lframe; a.sframe[1] = rframe;
  a.map1[0] = map1[0]; a.map1[1] = map1[1]; a.map2[0] = map2[0]; a.map2[1] = map2[1];
  a.rect[0] = rectL; a.rect[1] = rectR; a.rstride = rectStride; a.rframe = rectFrame;
  a.planes[0] = planesL; a.planes[1] = planesR;
  a.W = W; a.H = H; a.Wp = Wp; a.ftzero = ftzero;
  dim3 g((W + kFTW - 1) / kFTW, (H + kFTH - 1) / kFTH, nframes * 2);
  cudaError_t e = cudaFuncSetAttribute(k_front_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrontSmem);
  if (e != cudaSuccess) return e;
  k_front_fused<<<g, kFrontThreads, kFrontSmem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace svgpu

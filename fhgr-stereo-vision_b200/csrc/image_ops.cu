// image_ops.cu — per-pixel stages around the SGBM core, sm_100a.
// Each kernel replaces one OpenCV call of /root/reference/src/lib.cpp (cited per kernel);
// arithmetic follows SURVEY.md Appendix A.7-A.13.  All of them are HBM-bound streaming
// kernels: one thread per output pixel, coalesced rows, no shared-memory staging needed
// beyond what L1/L2 give a 3x3 / 5x5 / 11-tap neighbourhood.
#include "common.cuh"

namespace svgpu {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// ---- cv::initUndistortRectifyMap(K, dist, R, P, size, CV_16SC2) — src/lib.cpp:39-46 (A.10)
// One thread per image row.  OpenCV advances the homogeneous ray along a row by repeated addition, so the rounding
// of column u depends on the u additions before it; the row loop stays sequential to reproduce that bit for bit
// (one-time work: ~0.5 ms for 1080 rows).  All f64 with explicit round-to-nearest operations, i.e. no fused
// multiply-add, like the host code it replaces.  Five distortion coefficients (k1 k2 p1 p2 k3), which is what the
// reference's calibrations store (StereoCameraInfo.hpp:15-45).
__global__ void __launch_bounds__(64) k_init_rectify_map(const RectifyMapArgs k, int W, int H, int16_t* __restrict__ map1,
                                                         uint16_t* __restrict__ map2) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= H) return;
  const double dv = (double)v;
  double X = __dadd_rn(__dmul_rn(dv, k.iR[1]), k.iR[2]);
  double Y = __dadd_rn(__dmul_rn(dv, k.iR[4]), k.iR[5]);
  double Wc = __dadd_rn(__dmul_rn(dv, k.iR[7]), k.iR[8]);
  int16_t* m1 = map1 + (size_t)v * W * 2;
  uint16_t* m2 = map2 + (size_t)v * W;
  for (int u = 0; u < W; u++) {
    const double w = __ddiv_rn(1.0, Wc), x = __dmul_rn(X, w), y = __dmul_rn(Y, w);
    const double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y), r2 = __dadd_rn(x2, y2);
    const double xy2 = __dmul_rn(__dmul_rn(2.0, x), y);
    const double kr = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k.k3, r2), k.k2), r2), k.k1), r2));
    const double xd = __dadd_rn(__dadd_rn(__dmul_rn(x, kr), __dmul_rn(k.p1, xy2)),
                                __dmul_rn(k.p2, __dadd_rn(r2, __dmul_rn(2.0, x2))));
    const double yd = __dadd_rn(__dadd_rn(__dmul_rn(y, kr), __dmul_rn(k.p1, __dadd_rn(r2, __dmul_rn(2.0, y2)))),
                                __dmul_rn(k.p2, xy2));
    const double us = __dadd_rn(__dmul_rn(k.fx, xd), k.cx), vs = __dadd_rn(__dmul_rn(k.fy, yd), k.cy);
    const int iu = __double2int_rn(__dmul_rn(us, 32.0)), iv = __double2int_rn(__dmul_rn(vs, 32.0));  // half to even
    m1[2 * u] = (int16_t)min(max(iu >> 5, -32768), 32767);
    m1[2 * u + 1] = (int16_t)min(max(iv >> 5, -32768), 32767);
    m2[u] = (uint16_t)((iv & 31) * 32 + (iu & 31));
    X = __dadd_rn(X, k.iR[0]); Y = __dadd_rn(Y, k.iR[3]); Wc = __dadd_rn(Wc, k.iR[6]);
  }
}

cudaError_t launch_init_rectify_map(const RectifyMapArgs& k, int W, int H, int16_t* map1, uint16_t* map2, cudaStream_t st) {
  k_init_rectify_map<<<(H + 63) / 64, 64, 0, st>>>(k, W, H, map1, map2);
  return cudaGetLastError();
}

// ---- cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) with CV_16SC2 + CV_16UC1 maps — src/lib.cpp:89-92 (A.11)
template <int CN>
__global__ void __launch_bounds__(256) k_remap(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride,
                                               const int16_t* __restrict__ map1, const uint16_t* __restrict__ map2,
                                               uint8_t* __restrict__ dst, int W, int H, size_t dstride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const short2 m = ((const short2*)map1)[(size_t)y * W + x];
  const int f = map2[(size_t)y * W + x] & 1023, fx = f & 31, fy = f >> 5;
  const int x0 = m.x, y0 = m.y;
  // 15-bit weights (32-fy)(32-fx)*32 ... are exact and always sum to 32768
  const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  const bool xin0 = x0 >= 0 && x0 < sw, xin1 = x0 + 1 >= 0 && x0 + 1 < sw;
  const bool yin0 = y0 >= 0 && y0 < sh, yin1 = y0 + 1 >= 0 && y0 + 1 < sh;
  const uint8_t* r0 = src + (size_t)y0 * sstride + (size_t)x0 * CN;
  const uint8_t* r1 = r0 + sstride;
  uint8_t* o = dst + (size_t)y * dstride + (size_t)x * CN;
#pragma unroll
  for (int c = 0; c < CN; c++) {
    int acc = 16384;
    if (yin0 && xin0) acc += w00 * r0[c];
    if (yin0 && xin1) acc += w01 * r0[CN + c];
    if (yin1 && xin0) acc += w10 * r1[c];
    if (yin1 && xin1) acc += w11 * r1[CN + c];
    o[c] = (uint8_t)min(acc >> 15, 255);
  }
}

cudaError_t launch_remap(const uint8_t* src, int sw, int sh, size_t sstride, int cn, const int16_t* map1,
                         const uint16_t* map2, uint8_t* dst, int W, int H, size_t dstride, cudaStream_t st) {
  dim3 g((W + 255) / 256, H);
  if (cn == 3) k_remap<3><<<g, 256, 0, st>>>(src, sw, sh, sstride, map1, map2, dst, W, H, dstride);
  else k_remap<1><<<g, 256, 0, st>>>(src, sw, sh, sstride, map1, map2, dst, W, H, dstride);
  return cudaGetLastError();
}

// ---- cv::medianBlur(CV_16S, 3), replicate border — last stage of StereoSGBM::compute (A.7)
__device__ __forceinline__ void cswap(int& a, int& b) { const int t = min(a, b); b = max(a, b); a = t; }
__global__ void __launch_bounds__(256) k_median3(const int16_t* __restrict__ src, size_t sstride, size_t sframe,
                                                 int16_t* __restrict__ dst, size_t dstride, size_t dframe, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const uint8_t* sb = (const uint8_t*)src + blockIdx.z * sframe;
  const int xm = max(x - 1, 0), xp = min(x + 1, W - 1);
  const int16_t* r0 = (const int16_t*)(sb + (size_t)max(y - 1, 0) * sstride);
  const int16_t* r1 = (const int16_t*)(sb + (size_t)y * sstride);
  const int16_t* r2 = (const int16_t*)(sb + (size_t)min(y + 1, H - 1) * sstride);
  int v0 = r0[xm], v1 = r0[x], v2 = r0[xp], v3 = r1[xm], v4 = r1[x], v5 = r1[xp], v6 = r2[xm], v7 = r2[x], v8 = r2[xp];
  // 19-exchange median-of-9 network
  cswap(v1, v2); cswap(v4, v5); cswap(v7, v8); cswap(v0, v1); cswap(v3, v4); cswap(v6, v7);
  cswap(v1, v2); cswap(v4, v5); cswap(v7, v8); cswap(v0, v3); cswap(v5, v8); cswap(v4, v7);
  cswap(v3, v6); cswap(v1, v4); cswap(v2, v5); cswap(v4, v7); cswap(v4, v2); cswap(v6, v4);
  cswap(v4, v2);
  *(int16_t*)((uint8_t*)dst + blockIdx.z * dframe + (size_t)y * dstride + (size_t)x * 2) = (int16_t)v4;
}

// two neighbouring pixels per thread on packed signed 16-bit lanes (needs an even width and 4-byte aligned rows)
__device__ __forceinline__ void cswap2(uint32_t& a, uint32_t& b) { const uint32_t t = __vmins2(a, b); b = __vmaxs2(a, b); a = t; }
__global__ void __launch_bounds__(256) k_median3_x2(const int16_t* __restrict__ src, size_t sstride, size_t sframe,
                                                    int16_t* __restrict__ dst, size_t dstride, size_t dframe, int W, int H) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 2, y = blockIdx.y;
  if (x >= W) return;
  const uint8_t* sb = (const uint8_t*)src + blockIdx.z * sframe;
  const int ys[3] = {max(y - 1, 0), y, min(y + 1, H - 1)};
  uint32_t v[9];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    const uint32_t* r = (const uint32_t*)(sb + (size_t)ys[q] * sstride) + (x >> 1);
    const uint32_t c = __ldg(r);
    const uint32_t l = x > 0 ? __ldg(r - 1) : __byte_perm(c, 0, 0x1010);            // replicate the border pixel
    const uint32_t n = x + 2 < W ? __ldg(r + 1) : __byte_perm(c, 0, 0x3232);
    v[3 * q] = __byte_perm(l, c, 0x5432);      // (x-1, x)
    v[3 * q + 1] = c;                          // (x, x+1)
    v[3 * q + 2] = __byte_perm(c, n, 0x5432);  // (x+1, x+2)
  }
  cswap2(v[1], v[2]); cswap2(v[4], v[5]); cswap2(v[7], v[8]); cswap2(v[0], v[1]); cswap2(v[3], v[4]); cswap2(v[6], v[7]);
  cswap2(v[1], v[2]); cswap2(v[4], v[5]); cswap2(v[7], v[8]); cswap2(v[0], v[3]); cswap2(v[5], v[8]); cswap2(v[4], v[7]);
  cswap2(v[3], v[6]); cswap2(v[1], v[4]); cswap2(v[2], v[5]); cswap2(v[4], v[7]); cswap2(v[4], v[2]); cswap2(v[6], v[4]);
  cswap2(v[4], v[2]);
  *(uint32_t*)((uint8_t*)dst + blockIdx.z * dframe + (size_t)y * dstride + (size_t)x * 2) = v[4];
}

cudaError_t launch_median3(const int16_t* src, size_t sstride, int16_t* dst, size_t dstride, int W, int H, int nframes,
                           size_t sframe, size_t dframe, cudaStream_t st) {
  if (W % 2 == 0 && ((uintptr_t)src | (uintptr_t)dst | sstride | dstride | sframe | dframe) % 4 == 0) {
    dim3 g((W / 2 + 255) / 256, H, nframes);
    k_median3_x2<<<g, 256, 0, st>>>(src, sstride, sframe, dst, dstride, dframe, W, H);
  } else {
    dim3 g((W + 255) / 256, H, nframes);
    k_median3<<<g, 256, 0, st>>>(src, sstride, sframe, dst, dstride, dframe, W, H);
  }
  return cudaGetLastError();
}

// ---- cv::reprojectImageTo3D(disp CV_16S, Q, handleMissingValues) — src/lib.cpp:154 (A.12)
__global__ void k_min_s16(const int16_t* __restrict__ disp, size_t dstride, int W, int H, int* out) {
  int m = 32767;
  for (int y = blockIdx.y; y < H; y += gridDim.y) {
    const int16_t* r = (const int16_t*)((const uint8_t*)disp + (size_t)y * dstride);
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) m = min(m, (int)r[x]);
  }
  m = __reduce_min_sync(0xFFFFFFFFu, m);
  if ((threadIdx.x & 31) == 0) atomicMin(out, m);
}
__global__ void k_set_int(int* p, int v) { *p = v; }

// PPT pixels per thread: the 3 x PPT floats of a thread are contiguous (three 16-byte stores for PPT = 4)
template <int PPT>
__global__ void __launch_bounds__(256) k_reproject(const int16_t* __restrict__ disp, size_t dstride, float* __restrict__ xyz,
                                                   size_t xstride, int W, int H, const double* __restrict__ Q,
                                                   int handleMissing, const int* __restrict__ minDisp) {
  const int xb = (blockIdx.x * blockDim.x + threadIdx.x) * PPT, y = blockIdx.y;
  if (xb >= W) return;
  const int16_t* drow = (const int16_t*)((const uint8_t*)disp + (size_t)y * dstride);
  float* orow = (float*)((uint8_t*)xyz + (size_t)y * xstride);
  const int mind = handleMissing ? *minDisp : 0;
  float o[3 * PPT];
  int dv[PPT];
  if constexpr (PPT == 4) {
    const uint2 w = *(const uint2*)(drow + xb);   // 8-byte aligned: xb is a multiple of 4 and the launcher checked the strides
    dv[0] = (int16_t)(w.x & 0xFFFF); dv[1] = (int16_t)(w.x >> 16); dv[2] = (int16_t)(w.y & 0xFFFF); dv[3] = (int16_t)(w.y >> 16);
  } else {
    dv[0] = drow[xb];
  }
#pragma unroll
  for (int k = 0; k < PPT; k++) {
    const int x = xb + k, di = dv[k];
    const double d = (double)di;
    double h[4];
#pragma unroll
    for (int r = 0; r < 4; r++) h[r] = Q[r * 4 + 0] * x + Q[r * 4 + 1] * y + Q[r * 4 + 2] * d + Q[r * 4 + 3];
    // Vec3f /= double in OpenCV multiplies by the reciprocal; explicit round-to-nearest operations (no contraction)
    const double iw = __ddiv_rn(1.0, h[3]);
    o[3 * k] = (float)__dmul_rn((double)(float)h[0], iw);
    o[3 * k + 1] = (float)__dmul_rn((double)(float)h[1], iw);
    float z = (float)__dmul_rn((double)(float)h[2], iw);
    if (handleMissing && di == mind) z = 10000.f;
    o[3 * k + 2] = z;
  }
  if constexpr (PPT == 4) {
    float4* p = (float4*)(orow + (size_t)xb * 3);
    p[0] = make_float4(o[0], o[1], o[2], o[3]);
    p[1] = make_float4(o[4], o[5], o[6], o[7]);
    p[2] = make_float4(o[8], o[9], o[10], o[11]);
  } else {
    orow[(size_t)xb * 3] = o[0]; orow[(size_t)xb * 3 + 1] = o[1]; orow[(size_t)xb * 3 + 2] = o[2];
  }
}

cudaError_t launch_reproject(const int16_t* disp, size_t dstride, float* xyz, size_t xstride, int W, int H,
                             const double* Qdev, int handleMissing, int* minScratch, cudaStream_t st, int* nlaunch) {
  if (handleMissing) {
    k_set_int<<<1, 1, 0, st>>>(minScratch, 32767);
    dim3 g(min((W + 255) / 256, 8), min(H, 148));
    k_min_s16<<<g, 256, 0, st>>>(disp, dstride, W, H, minScratch);
    *nlaunch += 2;
  }
  if (W % 4 == 0 && ((uintptr_t)disp | dstride) % 8 == 0 && ((uintptr_t)xyz | xstride) % 16 == 0) {
    dim3 g((W / 4 + 255) / 256, H);
    k_reproject<4><<<g, 256, 0, st>>>(disp, dstride, xyz, xstride, W, H, Qdev, handleMissing, minScratch);
  } else {
    dim3 g((W + 255) / 256, H);
    k_reproject<1><<<g, 256, 0, st>>>(disp, dstride, xyz, xstride, W, H, Qdev, handleMissing, minScratch);
  }
  *nlaunch += 1;
  return cudaGetLastError();
}

// ---- StereoVision::CalculateDepthMapSimple — src/lib.cpp:279-294
__global__ void __launch_bounds__(256) k_depth(const int16_t* __restrict__ disp, size_t dstride, float* __restrict__ depth,
                                               size_t zstride, int W, int H, float bf) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int d = *(const int16_t*)((const uint8_t*)disp + (size_t)y * dstride + (size_t)x * 2);
  *(float*)((uint8_t*)depth + (size_t)y * zstride + (size_t)x * 4) = d > 0 ? __fdiv_rn(bf, (float)d) : 0.0f;
}
cudaError_t launch_depth(const int16_t* disp, size_t dstride, float* depth, size_t zstride, int W, int H, float bf,
                         cudaStream_t st) {
  dim3 g((W + 255) / 256, H);
  k_depth<<<g, 256, 0, st>>>(disp, dstride, depth, zstride, W, H, bf);
  return cudaGetLastError();
}

// ---- cv::filterSpeckles(img, newVal, maxSpeckleSize, maxDiff) — src/lib.cpp:265 (A.8)
// 4-connected components under the edge predicate |a-b| <= maxDiff (both != newVal) by lock-free
// union-find on pixel indices, then a size histogram per root and a thresholded rewrite.
__device__ __forceinline__ int uf_find(int* L, int i) {
  int p = ((volatile int*)L)[i];
  while (p != i) { i = p; p = ((volatile int*)L)[i]; }
  return i;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  for (;;) {
    a = uf_find(L, a); b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}
__device__ __forceinline__ int px(const int16_t* img, size_t stride, int x, int y) {
  return *(const int16_t*)((const uint8_t*)img + (size_t)y * stride + (size_t)x * 2);
}
// Pass 1, one CTA per row: every valid pixel is labelled with the first pixel of its horizontal run (maximal stretch of
// 4-connected neighbours within the row), found by a running maximum over the run breaks.  Horizontal unions are then already
// done, and every set has a root that is at most one pointer away within its row.
__global__ void __launch_bounds__(256) k_spk_rows(const int16_t* __restrict__ img, size_t stride, int W, int H, int newVal, int maxDiff,
                                                  int* __restrict__ L, int* __restrict__ sizes) {
  __shared__ int sWarp[8];
  __shared__ int sCarry;
  const int y = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) sCarry = -1;
  __syncthreads();
  for (int xb = 0; xb < W; xb += 256) {
    const int x = xb + threadIdx.x;
    const bool in = x < W;
    const int v = in ? px(img, stride, x, y) : newVal;
    const bool valid = in && v != newVal;
    bool brk = valid;
    if (valid && x > 0) {
      const int q = px(img, stride, x - 1, y);
      brk = q == newVal || abs(v - q) > maxDiff;
    }
    int k = brk ? x : -1;   // start of the run this pixel belongs to = the last break at or before it
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, k, o); if (lane >= o) k = max(k, t); }
    if (lane == 31) sWarp[warp] = k;
    __syncthreads();
    int pre = sCarry;
    for (int w = 0; w < warp; w++) pre = max(pre, sWarp[w]);
    k = max(k, pre);
    if (in) {
      L[y * W + x] = valid ? y * W + k : -1;
      sizes[y * W + x] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 255) sCarry = k;
    __syncthreads();
  }
}
// Pass 2: vertical unions.  The edge (x,y)-(x,y-1) is skipped when the pixel to the left already joins the same two sets
// (both rows connected to their left neighbour and the left column connected vertically).
__global__ void __launch_bounds__(256) k_spk_vunion(const int16_t* __restrict__ img, size_t stride, int W, int H, int newVal, int maxDiff, int* L) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y + 1;
  if (x >= W || y >= H) return;
  const int v = px(img, stride, x, y), u = px(img, stride, x, y - 1);
  if (v == newVal || u == newVal || abs(v - u) > maxDiff) return;
  if (x > 0) {
    const int vl = px(img, stride, x - 1, y), ul = px(img, stride, x - 1, y - 1);
    if (vl != newVal && ul != newVal && abs(v - vl) <= maxDiff && abs(u - ul) <= maxDiff && abs(vl - ul) <= maxDiff) return;
  }
  uf_union(L, y * W + x, (y - 1) * W + x);
}
// Pass 3: flatten + component sizes.  Only "larger than maxSize or not" matters, so a counter that has passed maxSize is left
// alone (no atomic traffic on the big components), and the lanes of a warp that share a root add once.
__global__ void __launch_bounds__(256) k_spk_count(int W, int H, int maxSize, int* L, int* sizes) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  int r = -1;
  if (x < W) {
    const int i = y * W + x;
    if (L[i] >= 0) {
      r = uf_find(L, i);
      L[i] = r;  // path compression is safe here: no unions run concurrently
    }
  }
  const unsigned m = __match_any_sync(0xFFFFFFFFu, r);
  if (r >= 0 && (threadIdx.x & 31) == __ffs(m) - 1 && __ldcg(&sizes[r]) <= maxSize) atomicAdd(&sizes[r], __popc(m));
}
__global__ void k_spk_apply(int16_t* img, size_t stride, int W, int H, int newVal, int maxSize, const int* L, const int* sizes) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int r = L[y * W + x];
  if (r >= 0 && sizes[r] <= maxSize) *(int16_t*)((uint8_t*)img + (size_t)y * stride + (size_t)x * 2) = (int16_t)newVal;
}
cudaError_t launch_filter_speckles(int16_t* disp, size_t stride, int W, int H, int newVal, int maxSize, int maxDiff,
                                   int* labels, int* sizes, int* /*changed*/, cudaStream_t st, int* nlaunch) {
  dim3 g((W + 255) / 256, H);
  k_spk_rows<<<H, 256, 0, st>>>(disp, stride, W, H, newVal, maxDiff, labels, sizes);
  if (H > 1) {
    dim3 gv((W + 255) / 256, H - 1);
    k_spk_vunion<<<gv, 256, 0, st>>>(disp, stride, W, H, newVal, maxDiff, labels);
  }
  k_spk_count<<<g, 256, 0, st>>>(W, H, maxSize, labels, sizes);
  k_spk_apply<<<g, 256, 0, st>>>(disp, stride, W, H, newVal, maxSize, labels, sizes);
  *nlaunch += H > 1 ? 4 : 3;
  return cudaGetLastError();
}

// ---- cv::morphologyEx(MORPH_CLOSE, 5x5 rect) on CV_16S — src/lib.cpp:267-268 (A.13)
// Dilation then erosion with a 5 x 5 rectangle; pixels outside the image do not take part (OpenCV's morphology border).
// One launch: a CTA loads its tile with a 4-pixel halo into shared memory, dilates the tile + 2-pixel halo there (rows, then
// columns: the rectangle is separable), erodes it the same way and writes the tile: one global load and one store per pixel.
constexpr int kMW = 64, kMH = 16, kMT = 256;   // tile and threads
__global__ void __launch_bounds__(kMT) k_morph_close5(const int16_t* __restrict__ src, size_t sstride, int16_t* __restrict__ dst, size_t dstride,
                                                      int W, int H) {
  __shared__ int16_t sA[kMH + 8][kMW + 8];   // input, later the eroded rows
  __shared__ int16_t sB[kMH + 8][kMW + 8];   // row-wise pass of the current operation
  const int x0 = blockIdx.x * kMW - 4, y0 = blockIdx.y * kMH - 4;   // image position of sA[0][0]
  const int16_t NEG = -32768, POS = 32767;
  for (int i = threadIdx.x; i < (kMH + 8) * (kMW + 8); i += kMT) {
    const int ty = i / (kMW + 8), tx = i - ty * (kMW + 8);
    const int x = x0 + tx, y = y0 + ty;
    sA[ty][tx] = (x >= 0 && x < W && y >= 0 && y < H) ? px(src, sstride, x, y) : NEG;   // outside: neutral for the dilation
  }
  __syncthreads();
  // dilation, rows: sB[ty][tx] for tx in [2, kMW + 6)
  for (int i = threadIdx.x; i < (kMH + 8) * (kMW + 4); i += kMT) {
    const int ty = i / (kMW + 4), tx = i - ty * (kMW + 4) + 2;
    int v = sA[ty][tx - 2];
#pragma unroll
    for (int d = -1; d <= 2; d++) v = max(v, (int)sA[ty][tx + d]);
    sB[ty][tx] = (int16_t)v;
  }
  __syncthreads();
  // dilation, columns: the dilated image on rows / columns [2, +4 halo); positions outside the image become neutral for the erosion
  for (int i = threadIdx.x; i < (kMH + 4) * (kMW + 4); i += kMT) {
    const int ty = i / (kMW + 4) + 2, tx = i - (ty - 2) * (kMW + 4) + 2;
    int v = sB[ty - 2][tx];
#pragma unroll
    for (int d = -1; d <= 2; d++) v = max(v, (int)sB[ty + d][tx]);
    const int x = x0 + tx, y = y0 + ty;
    sA[ty][tx] = (x >= 0 && x < W && y >= 0 && y < H) ? (int16_t)v : POS;
  }
  __syncthreads();
  // erosion, rows: tx in [4, kMW + 4), ty in [2, kMH + 6)
  for (int i = threadIdx.x; i < (kMH + 4) * kMW; i += kMT) {
    const int ty = i / kMW + 2, tx = i - (ty - 2) * kMW + 4;
    int v = sA[ty][tx - 2];
#pragma unroll
    for (int d = -1; d <= 2; d++) v = min(v, (int)sA[ty][tx + d]);
    sB[ty][tx] = (int16_t)v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMH * kMW; i += kMT) {
    const int ty = i / kMW + 4, tx = i - (ty - 4) * kMW + 4;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= W || y >= H) continue;
    int v = sB[ty - 2][tx];
#pragma unroll
    for (int d = -1; d <= 2; d++) v = min(v, (int)sB[ty + d][tx]);
    *(int16_t*)((uint8_t*)dst + (size_t)y * dstride + (size_t)x * 2) = (int16_t)v;
  }
}
// tmp is no longer needed (kept in the signature for the callers' buffers)
cudaError_t launch_morph_close5(const int16_t* src, size_t sstride, int16_t* /*tmp*/, int16_t* dst, size_t dstride, int W,
                                int H, cudaStream_t st) {
  dim3 g((W + kMW - 1) / kMW, (H + kMH - 1) / kMH);
  k_morph_close5<<<g, kMT, 0, st>>>(src, sstride, dst, dstride, W, H);
  return cudaGetLastError();
}

// ---- GaussianBlur 11x11 s1.4 -> BGR2GRAY -> 3x3 sharpen — src/lib.cpp:231-246 (A.13)
__constant__ int kGauss11[11] = {0, 1, 8, 26, 57, 72, 57, 26, 8, 1, 0};
__global__ void __launch_bounds__(256) k_gauss_h(const uint8_t* __restrict__ bgr, size_t bstride, uint16_t* __restrict__ tmpH, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const uint8_t* r = bgr + (size_t)y * bstride;
  int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
  for (int t = 1; t < 10; t++) {  // taps 0 and 10 are zero
    const uint8_t* p = r + (size_t)reflect101(x + t - 5, W) * 3;
    s0 += kGauss11[t] * p[0]; s1 += kGauss11[t] * p[1]; s2 += kGauss11[t] * p[2];
  }
  uint16_t* o = tmpH + ((size_t)y * W + x) * 3;
  o[0] = (uint16_t)s0; o[1] = (uint16_t)s1; o[2] = (uint16_t)s2;
}
__global__ void __launch_bounds__(256) k_gauss_v_gray(const uint16_t* __restrict__ tmpH, uint8_t* __restrict__ gray, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  unsigned s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
  for (int t = 1; t < 10; t++) {
    const uint16_t* p = tmpH + ((size_t)reflect101(y + t - 5, H) * W + x) * 3;
    s0 += kGauss11[t] * p[0]; s1 += kGauss11[t] * p[1]; s2 += kGauss11[t] * p[2];
  }
  const int b = (s0 + 32768u) >> 16, g = (s1 + 32768u) >> 16, r = (s2 + 32768u) >> 16;
  gray[(size_t)y * W + x] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15);
}
__global__ void __launch_bounds__(256) k_sharpen3(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t dstride, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int c = src[(size_t)y * W + x];
  const int l = src[(size_t)y * W + reflect101(x - 1, W)], r = src[(size_t)y * W + reflect101(x + 1, W)];
  const int u = src[(size_t)reflect101(y - 1, H) * W + x], d = src[(size_t)reflect101(y + 1, H) * W + x];
  dst[(size_t)y * dstride + x] = (uint8_t)min(max(5 * c - l - r - u - d, 0), 255);
}
cudaError_t launch_prefilter_chain(const uint8_t* bgr, size_t bstride, uint16_t* tmpH, uint8_t* tmpGray, uint8_t* gray,
                                   size_t gstride, int W, int H, cudaStream_t st) {
  dim3 g((W + 255) / 256, H);
  k_gauss_h<<<g, 256, 0, st>>>(bgr, bstride, tmpH, W, H);
  k_gauss_v_gray<<<g, 256, 0, st>>>(tmpH, tmpGray, W, H);
  k_sharpen3<<<g, 256, 0, st>>>(tmpGray, gray, gstride, W, H);
  return cudaGetLastError();
}

// ---- cv::resize(8U, INTER_LINEAR) — src/lib.cpp:102 (A.13), 11-bit fixed-point two-tap bilinear
template <int CN>
__global__ void __launch_bounds__(256) k_resize(const uint8_t* __restrict__ src, int sw, int sh, size_t sstride,
                                                uint8_t* __restrict__ dst, int dw, int dh, size_t dstride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw) return;
  const double sx = (double)sw / dw, sy = (double)sh / dh;
  float fx = (float)((x + 0.5) * sx - 0.5);
  int i0 = (int)floorf(fx);
  fx -= i0;
  if (i0 < 0) { i0 = 0; fx = 0; }
  if (i0 >= sw - 1) { i0 = sw - 1; fx = 0; }
  float fy = (float)((y + 0.5) * sy - 0.5);
  int j0 = (int)floorf(fy);
  fy -= j0;
  if (j0 < 0) { j0 = 0; fy = 0; }
  if (j0 >= sh - 1) { j0 = sh - 1; fy = 0; }
  const int i1 = min(i0 + 1, sw - 1), j1 = min(j0 + 1, sh - 1);
  const int a0 = __float2int_rn((1.f - fx) * 2048.f), a1 = __float2int_rn(fx * 2048.f);
  const int b0 = __float2int_rn((1.f - fy) * 2048.f), b1 = __float2int_rn(fy * 2048.f);
  const uint8_t* r0 = src + (size_t)j0 * sstride;
  const uint8_t* r1 = src + (size_t)j1 * sstride;
  uint8_t* o = dst + (size_t)y * dstride + (size_t)x * CN;
#pragma unroll
  for (int c = 0; c < CN; c++) {
    const int h0 = r0[i0 * CN + c] * a0 + r0[i1 * CN + c] * a1;
    const int h1 = r1[i0 * CN + c] * a0 + r1[i1 * CN + c] * a1;
    o[c] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
  }
}
cudaError_t launch_resize(const uint8_t* src, int sw, int sh, size_t sstride, int cn, uint8_t* dst, int dw, int dh,
                          size_t dstride, cudaStream_t st) {
  dim3 g((dw + 255) / 256, dh);
  if (cn == 3) k_resize<3><<<g, 256, 0, st>>>(src, sw, sh, sstride, dst, dw, dh, dstride);
  else k_resize<1><<<g, 256, 0, st>>>(src, sw, sh, sstride, dst, dw, dh, dstride);
  return cudaGetLastError();
}

}  // namespace svgpu

// sgbm.cu — semi-global block matching for sm_100a (B200).
//
// Replaces the call `stereo->compute(left_gray, right_gray, disparity)` at
// /root/reference/src/lib.cpp:262 (cv::StereoSGBM, parameters :248-259).  OpenCV's
// sources are not part of the reference tree; the arithmetic follows SURVEY.md
// Appendix A.1-A.7, which tests/ check bit-exactly against the cv2 binary.
//
// Data layout in HBM (per pair in flight):
//   planes  [y][2][Wp] 2 x u32 Birchfield-Tomasi operands of one image per channel (gradient, intensity): for the
//                              left image {lo | v<<16, -hi | -v<<16}, for the right image {-v | -hi<<16, v | lo<<16}
//                              (16-bit fields), i.e. exactly the words the cost kernel pairs up (sgbm_cost.cu)
//   C, S    [y][x'][d] u16     cost volume Ctrue and aggregated costs, d innermost so that one
//                              pixel's D costs are D*2 contiguous bytes (256 B at D=128); inside each
//                              group of 16 disparities the halves are interleaved (sgbm_device.cuh)
//   sel     [y][x']    2 x u32 winner of the pixel: reject<<31 | minS<<16 | best, S[best-1]<<16 | S[best+1]
// Costs are kept as unsigned 16-bit lanes packed two per register; all min/add work
// runs on the packed-SIMD integer instructions of sm_100 (VIMNMX.U16x2, VIMNMX3,
// VIADDMNMX, VIADD.16x2) and __reduce_min_sync (CREDUX).
#include "common.cuh"
#include "sgbm_device.cuh"

namespace svgpu {

// ------------------------------------------------------------------------------------------------
// K1: pre-filter (A.2) + Birchfield-Tomasi half-pixel bounds (A.3) -> operand planes
// ------------------------------------------------------------------------------------------------
// r / u / d point at the wanted channel of pixel 0 of the row / the row above / below; `cn` interleaved channels
__device__ __forceinline__ int pf_grad(const uint8_t* r, const uint8_t* u, const uint8_t* d, int x, int W, int cn, int ftzero) {
  if (x <= 0 || x >= W - 1) return ftzero;
  const int a = (x + 1) * cn, b = (x - 1) * cn;
  int v = 2 * ((int)r[a] - (int)r[b]) + ((int)u[a] - (int)u[b]) + ((int)d[a] - (int)d[b]);
  return min(max(v, -ftzero), ftzero) + ftzero;
}
__device__ __forceinline__ int pf_int(const uint8_t* r, int x, int W, int cn, int ftzero) {
  return (x <= 0 || x >= W - 1) ? ftzero : (int)r[x * cn];
}

// Plane order of a row: [channel][gradient, intensity][Wp] (cn = 3: StereoSGBM sums the cost over the colour channels,
// each with its own gradient and intensity term)
template <int cn>
__global__ void __launch_bounds__(256) k_prefilter_planes(const uint8_t* __restrict__ left, size_t lstride, size_t lframe,
                                                          const uint8_t* __restrict__ right, size_t rstride, size_t rframe,
                                                          uint8_t* __restrict__ planesL, uint8_t* __restrict__ planesR,
                                                          int W, int H, int Wp, int ftzero) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int side = blockIdx.z & 1, frame = blockIdx.z >> 1;
  if (x >= W) return;
  const uint8_t* img = side ? right + frame * rframe : left + frame * lframe;
  const size_t stride = side ? rstride : lstride;
  auto h16 = [](int v) { return (uint32_t)v & 0xFFFFu; };
  uint2* p = (uint2*)(side ? planesR : planesL) + ((size_t)frame * H + y) * 2 * cn * Wp + x;
#pragma unroll
  for (int c = 0; c < cn; c++, p += 2 * Wp) {
    const uint8_t* r = img + (size_t)y * stride + c;
    const uint8_t* u = y > 0 ? r - stride : r;
    const uint8_t* d = y < H - 1 ? r + stride : r;
    const int g0 = pf_grad(r, u, d, x, W, cn, ftzero), i0 = pf_int(r, x, W, cn, ftzero);
    int glo = g0, ghi = g0, ilo = i0, ihi = i0;
    if (x > 0) {
      const int g = (g0 + pf_grad(r, u, d, x - 1, W, cn, ftzero)) >> 1, i = (i0 + pf_int(r, x - 1, W, cn, ftzero)) >> 1;
      glo = min(glo, g); ghi = max(ghi, g); ilo = min(ilo, i); ihi = max(ihi, i);
    }
    if (x < W - 1) {
      const int g = (g0 + pf_grad(r, u, d, x + 1, W, cn, ftzero)) >> 1, i = (i0 + pf_int(r, x + 1, W, cn, ftzero)) >> 1;
      glo = min(glo, g); ghi = max(ghi, g); ilo = min(ilo, i); ihi = max(ihi, i);
    }
    if (side == 0) {
      p[0] = make_uint2(h16(glo) | h16(g0) << 16, h16(-ghi) | h16(-g0) << 16);
      p[Wp] = make_uint2(h16(ilo) | h16(i0) << 16, h16(-ihi) | h16(-i0) << 16);
    } else {
      p[0] = make_uint2(h16(-g0) | h16(-ghi) << 16, h16(g0) | h16(glo) << 16);
      p[Wp] = make_uint2(h16(-i0) | h16(-ihi) << 16, h16(i0) | h16(ilo) << 16);
    }
  }
}

// Single-channel fast path: four neighbouring pixels per thread, the three image rows read as aligned 32-bit words, every
// column sum of the Sobel filter computed once (needs W % 4 == 0 and 4-byte aligned rows; otherwise the kernel above runs).
__global__ void __launch_bounds__(256) k_prefilter_planes_x4(const uint8_t* __restrict__ left, size_t lstride, size_t lframe,
                                                             const uint8_t* __restrict__ right, size_t rstride, size_t rframe,
                                                             uint8_t* __restrict__ planesL, uint8_t* __restrict__ planesR,
                                                             int W, int H, int Wp, int ftzero) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const int side = blockIdx.z & 1, frame = blockIdx.z >> 1;
  if (x0 >= W) return;
  const uint8_t* img = side ? right + frame * rframe : left + frame * lframe;
  const size_t stride = side ? rstride : lstride;
  const uint8_t* rows[3];
  rows[1] = img + (size_t)y * stride;
  rows[0] = y > 0 ? rows[1] - stride : rows[1];
  rows[2] = y < H - 1 ? rows[1] + stride : rows[1];
  int cs[8], ctr[8];  // column sums u + 2r + d and the centre row, columns x0-2 .. x0+5
#pragma unroll
  for (int k = 0; k < 8; k++) cs[k] = ctr[k] = 0;
#pragma unroll
  for (int q = 0; q < 3; q++) {
    const uint32_t* w = (const uint32_t*)(rows[q] + x0);
    const uint32_t w0 = x0 >= 4 ? __ldg(w - 1) : 0u, w1 = __ldg(w), w2 = x0 + 4 < W ? __ldg(w + 1) : 0u;
    const int b[8] = {(int)((w0 >> 16) & 255), (int)(w0 >> 24), (int)(w1 & 255), (int)((w1 >> 8) & 255), (int)((w1 >> 16) & 255), (int)(w1 >> 24),
                      (int)(w2 & 255), (int)((w2 >> 8) & 255)};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      cs[k] += q == 1 ? 2 * b[k] : b[k];
      if (q == 1) ctr[k] = b[k];
    }
  }
  int g[6], it[6];  // pre-filtered gradient and intensity of columns x0-1 .. x0+4
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const int x = x0 - 1 + j;
    const bool border = x <= 0 || x >= W - 1;
    g[j] = border ? ftzero : min(max(cs[j + 2] - cs[j], -ftzero), ftzero) + ftzero;
    it[j] = border ? ftzero : ctr[j + 1];
  }
  auto h16 = [](int v) { return (uint32_t)v & 0xFFFFu; };
  uint32_t og[8], oi[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int x = x0 + j;
    const int g0 = g[j + 1], i0 = it[j + 1];
    int glo = g0, ghi = g0, ilo = i0, ihi = i0;
    if (x > 0) {
      const int a = (g0 + g[j]) >> 1, c = (i0 + it[j]) >> 1;
      glo = min(glo, a); ghi = max(ghi, a); ilo = min(ilo, c); ihi = max(ihi, c);
    }
    if (x < W - 1) {
      const int a = (g0 + g[j + 2]) >> 1, c = (i0 + it[j + 2]) >> 1;
      glo = min(glo, a); ghi = max(ghi, a); ilo = min(ilo, c); ihi = max(ihi, c);
    }
    if (side == 0) {
      og[2 * j] = h16(glo) | h16(g0) << 16; og[2 * j + 1] = h16(-ghi) | h16(-g0) << 16;
      oi[2 * j] = h16(ilo) | h16(i0) << 16; oi[2 * j + 1] = h16(-ihi) | h16(-i0) << 16;
    } else {
      og[2 * j] = h16(-g0) | h16(-ghi) << 16; og[2 * j + 1] = h16(g0) | h16(glo) << 16;
      oi[2 * j] = h16(-i0) | h16(-ihi) << 16; oi[2 * j + 1] = h16(i0) | h16(ilo) << 16;
    }
  }
  uint2* p = (uint2*)(side ? planesR : planesL) + ((size_t)frame * H + y) * 2 * Wp + x0;
  ((uint4*)p)[0] = make_uint4(og[0], og[1], og[2], og[3]);
  ((uint4*)p)[1] = make_uint4(og[4], og[5], og[6], og[7]);
  ((uint4*)(p + Wp))[0] = make_uint4(oi[0], oi[1], oi[2], oi[3]);
  ((uint4*)(p + Wp))[1] = make_uint4(oi[4], oi[5], oi[6], oi[7]);
}

// K5: sub-pixel disparity + left-right consistency check of one image row (A.6, second half), from the
// per-pixel selection the second sweep left in `sel`.  disp2 is rebuilt by a scatter-min of
// (minS, W1-1-x') keys: lowest cost wins, ties go to the largest x' — the same winner as OpenCV's
// right-to-left scan with a strict '>' test.
__global__ void __launch_bounds__(256) k_select_lr(const uint2* __restrict__ sel, int16_t* __restrict__ dispRaw, SgbmCfg cfg) {
  extern __shared__ uint32_t sKey[];  // [W] keys, then [W] int16 disparities
  const int y = blockIdx.x, frame = blockIdx.y;
  const int W = cfg.W, W1 = cfg.W1;
  int16_t* sDisp = (int16_t*)(sKey + W);
  const uint2* row = sel + ((size_t)frame * cfg.H + y) * W1;
  int16_t* dr = dispRaw + ((size_t)frame * cfg.H + y) * W;
  for (int x = threadIdx.x; x < W; x += blockDim.x) { sKey[x] = 0xFFFFFFFFu; sDisp[x] = (int16_t)cfg.invalid; }
  __syncthreads();
  for (int xp = threadIdx.x; xp < W1; xp += blockDim.x) {
    const uint2 v = row[xp];
    const int minS = (int)((v.x >> 16) & 0x7FFFu), best = (int)(v.x & 0xFFFFu);
    if ((v.x & 0x80000000u) || minS >= 32767) continue;  // not unique / all costs saturated
    const int x2 = xp + cfg.minX1 - best - cfg.minD;
    atomicMin(&sKey[x2], ((uint32_t)minS << 16) | (uint32_t)(W1 - 1 - xp));
    int dfix = best * 16;
    if (best > 0 && best < cfg.D - 1) {
      const int vm = (int)(v.y >> 16), vp = (int)(v.y & 0xFFFFu);
      const int den = max(vm + vp - 2 * minS, 1);
      dfix += ((vm - vp) * 16 + den) / (den * 2);
    }
    sDisp[xp + cfg.minX1] = (int16_t)(dfix + cfg.minD * 16);
  }
  __syncthreads();
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    int d1 = sDisp[x];
    if (x >= cfg.minX1 && x < cfg.minX1 + W1 && d1 != cfg.invalid) {
      const int _d = d1 >> 4, d_ = (d1 + 15) >> 4, _x = x - _d, x_ = x - d_;
      bool c0 = false, c1 = false;
      if (_x >= 0 && _x < W) {
        const uint32_t k = sKey[_x];
        // disp2 holds best+minD of the winner, or (minD-1)*16 when nothing landed there
        const int d2 = (k == 0xFFFFFFFFu) ? cfg.invalid : ((W1 - 1 - (int)(k & 0xFFFFu)) + cfg.minX1 - _x);
        c0 = d2 >= cfg.minD && abs(d2 - _d) > cfg.d12;
      }
      if (x_ >= 0 && x_ < W) {
        const uint32_t k = sKey[x_];
        const int d2 = (k == 0xFFFFFFFFu) ? cfg.invalid : ((W1 - 1 - (int)(k & 0xFFFFu)) + cfg.minX1 - x_);
        c1 = d2 >= cfg.minD && abs(d2 - d_) > cfg.d12;
      }
      if (c0 && c1) d1 = cfg.invalid;
    }
    dr[x] = (int16_t)d1;
  }
}

__global__ void k_fill_s16(int16_t* p, size_t n, int v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (int16_t)v;
}
__global__ void k_fill_rows_s16(int16_t* p, size_t stride, size_t frame, int W, int H, int v) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < W) *(int16_t*)((uint8_t*)p + blockIdx.z * frame + (size_t)y * stride + (size_t)x * 2) = (int16_t)v;
}

// MODE_SGBM_3WAY: assemble the image from the rows of its four stripe frames (common.cuh: StripeGeom)
__global__ void __launch_bounds__(256) k_gather_stripes(const int16_t* __restrict__ stripes, int16_t* __restrict__ full, int W, int H,
                                                        StripeGeom g, int invalid) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y, pair = blockIdx.z;
  if (x >= W) return;
  const int s = min(yy / g.sz, 3), k = yy - s * g.sz;
  const int y = g.ov + k - g.dstOff[s];
  int v = invalid;
  if (y >= 0 && y < g.rows[s]) v = stripes[(((size_t)pair * 4 + s) * g.Hs + y) * W + x];
  full[((size_t)pair * H + yy) * W + x] = (int16_t)v;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

#define SV_RET(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)
#define SV_LAUNCHED() do { if (stats) stats->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; } while (0)

// k_select_lr keeps one row of keys + disparities (6 B per column) in dynamic shared memory: opt in beyond 48 KB (W > 8192)
static cudaError_t select_lr_smem(size_t bytes) {
  return bytes > 48 * 1024 ? cudaFuncSetAttribute(k_select_lr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) : cudaSuccess;
}

cudaError_t sgbm_forward(const SgbmCfg& cfg, const SgbmScratch& sc, int nframes, const uint8_t* left, size_t lstride,
                         size_t lframe, const uint8_t* right, size_t rstride, size_t rframe, int16_t* disp,
                         size_t dstride, size_t dframe, cudaStream_t st, LaunchStats* stats, cudaEvent_t* ev, int debugS, int cn) {
  const int W = cfg.W, H = cfg.H;
  if (cfg.W1 <= 0) {  // A.1: nothing computable, whole map invalid
    dim3 g((W + 255) / 256, H, nframes);
    k_fill_rows_s16<<<g, 256, 0, st>>>(disp, dstride, dframe, W, H, cfg.invalid);
    SV_LAUNCHED();
    return cudaSuccess;
  }
  if (ev) SV_RET(cudaEventRecord(ev[0], st));
  if (left != nullptr) {   // nullptr: the operand planes are already in the scratch (front_fused.cu)
    dim3 g((W + 255) / 256, H, nframes * 2);
    const bool x4 = cn == 1 && W % 4 == 0 && ((uintptr_t)left | (uintptr_t)right | lstride | rstride | lframe | rframe) % 4 == 0;
    if (x4) {
      dim3 g4((W / 4 + 255) / 256, H, nframes * 2);
      k_prefilter_planes_x4<<<g4, 256, 0, st>>>(left, lstride, lframe, right, rstride, rframe, sc.planesL, sc.planesR, W, H, cfg.Wp, cfg.ftzero);
    } else if (cn == 3)
      k_prefilter_planes<3><<<g, 256, 0, st>>>(left, lstride, lframe, right, rstride, rframe, sc.planesL, sc.planesR, W, H, cfg.Wp, cfg.ftzero);
    else
      k_prefilter_planes<1><<<g, 256, 0, st>>>(left, lstride, lframe, right, rstride, rframe, sc.planesL, sc.planesR, W, H, cfg.Wp, cfg.ftzero);
    SV_LAUNCHED();
  }
  if (cfg.mode == SV_MODE_SGBM_3WAY) {
    const StripeGeom g = stripe_geom(H, cfg.r);
    SgbmCfg cs = cfg;   // a stripe frame
    cs.H = g.Hs;
    const size_t cellsStripe = (size_t)cfg.W1 * g.Hs * cfg.D;
    const size_t planeRowBytes = (size_t)2 * cn * cfg.Wp * sizeof(uint2);
    // rows of a stripe frame below its stripe are never produced (only the first stripe is shorter than Hs): keep them defined
    for (int s = 0; s < 4; s++) {
      if (g.rows[s] >= g.Hs) continue;
      const size_t rowBytes = (size_t)cfg.W1 * cfg.D * 2;
      SV_RET(cudaMemset2DAsync(sc.C + s * cellsStripe + (size_t)g.rows[s] * cfg.W1 * cfg.D, 4 * cellsStripe * 2, 0,
                               (size_t)(g.Hs - g.rows[s]) * rowBytes, nframes, st));
    }
    for (int s = 0; s < 4; s++) {
      if (g.rows[s] <= 0) continue;
      SgbmCfg cc = cfg;
      cc.H = H - g.start[s];  // block sums clamp at the first row of the stripe frame, and at the last image row
      const CostView v{sc.planesL + g.start[s] * planeRowBytes, sc.planesR + g.start[s] * planeRowBytes, sc.C + s * cellsStripe,
                       4 * cellsStripe, H, g.rows[s]};
      SV_RET(sgbm_cost_volume(cc, v, nframes, cn, st));
      if (stats) stats->launches++;
    }
    if (ev) SV_RET(cudaEventRecord(ev[1], st));
    SgbmScratch ss = sc;
    ss.frameCells = cellsStripe;
    SV_RET(sgbm_aggregate_sweep(cs, ss, nframes * 4, 0, st, stats, cn));
    if (ev) SV_RET(cudaEventRecord(ev[2], st));
    int16_t* dispStripes = sc.dispRaw;
    int16_t* dispFull = sc.dispRaw + (size_t)nframes * 4 * g.Hs * W;
    {
      dim3 gs(g.Hs, nframes * 4);
      SV_RET(select_lr_smem((size_t)W * 6));
      k_select_lr<<<gs, 256, (size_t)W * 6, st>>>(sc.sel, dispStripes, cs);
      SV_LAUNCHED();
      dim3 gg((W + 255) / 256, H, nframes);
      k_gather_stripes<<<gg, 256, 0, st>>>(dispStripes, dispFull, W, H, g, cfg.invalid);
      SV_LAUNCHED();
      SV_RET(launch_median3(dispFull, (size_t)W * 2, disp, dstride, W, H, nframes, (size_t)W * H * 2, dframe, st));
      if (stats) stats->launches++;
    }
    if (ev) SV_RET(cudaEventRecord(ev[3], st));
    return cudaSuccess;
  }
  {
    const CostView v{sc.planesL, sc.planesR, sc.C, sc.frameCells, H, H};
    SV_RET(sgbm_cost_volume(cfg, v, nframes, cn, st));
  }
  if (stats) stats->launches++;
  if (ev) SV_RET(cudaEventRecord(ev[1], st));
  SV_RET(sgbm_aggregate_sweep(cfg, sc, nframes, debugS, st, stats, cn));
  if (ev) SV_RET(cudaEventRecord(ev[2], st));
  {
    dim3 g(H, nframes);
    SV_RET(select_lr_smem((size_t)W * 6));
    k_select_lr<<<g, 256, (size_t)W * 6, st>>>(sc.sel, sc.dispRaw, cfg);
    SV_LAUNCHED();
    SV_RET(launch_median3(sc.dispRaw, (size_t)W * 2, disp, dstride, W, H, nframes, (size_t)W * H * 2, dframe, st));
    if (stats) stats->launches++;
  }
  if (ev) SV_RET(cudaEventRecord(ev[3], st));
  return cudaSuccess;
}

}  // namespace svgpu

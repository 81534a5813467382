// sgbm.cu — semi-global block matching for sm_100a (B200).
//
// Replaces the call `stereo->compute(left_gray, right_gray, disparity)` at
// /root/reference/src/lib.cpp:262 (cv::StereoSGBM, parameters :248-259).  OpenCV's
// sources are not part of the reference tree; the arithmetic follows SURVEY.md
// Appendix A.1-A.7, which tests/ check bit-exactly against the cv2 binary.
//
// Data layout in HBM (per pair in flight):
//   planes  [y][6][Wp] u8      pre-filter rows of one image: g, lo(g), hi(g), i, lo(i), hi(i)
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
// K1: pre-filter planes (A.2) + Birchfield-Tomasi half-pixel bounds (A.3)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pf_grad(const uint8_t* r, const uint8_t* u, const uint8_t* d, int x, int W, int ftzero) {
  if (x <= 0 || x >= W - 1) return ftzero;
  int v = 2 * ((int)r[x + 1] - (int)r[x - 1]) + ((int)u[x + 1] - (int)u[x - 1]) + ((int)d[x + 1] - (int)d[x - 1]);
  return min(max(v, -ftzero), ftzero) + ftzero;
}
__device__ __forceinline__ int pf_int(const uint8_t* r, int x, int W, int ftzero) {
  return (x <= 0 || x >= W - 1) ? ftzero : (int)r[x];
}

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
  const uint8_t* r = img + (size_t)y * stride;
  const uint8_t* u = y > 0 ? r - stride : r;
  const uint8_t* d = y < H - 1 ? r + stride : r;
  const int g0 = pf_grad(r, u, d, x, W, ftzero), i0 = pf_int(r, x, W, ftzero);
  int glo = g0, ghi = g0, ilo = i0, ihi = i0;
  if (x > 0) {
    const int g = (g0 + pf_grad(r, u, d, x - 1, W, ftzero)) >> 1, i = (i0 + pf_int(r, x - 1, W, ftzero)) >> 1;
    glo = min(glo, g); ghi = max(ghi, g); ilo = min(ilo, i); ihi = max(ihi, i);
  }
  if (x < W - 1) {
    const int g = (g0 + pf_grad(r, u, d, x + 1, W, ftzero)) >> 1, i = (i0 + pf_int(r, x + 1, W, ftzero)) >> 1;
    glo = min(glo, g); ghi = max(ghi, g); ilo = min(ilo, i); ihi = max(ihi, i);
  }
  uint8_t* p = (side ? planesR : planesL) + ((size_t)frame * H + y) * 6 * Wp + x;
  p[0] = (uint8_t)g0; p[Wp] = (uint8_t)glo; p[2 * Wp] = (uint8_t)ghi;
  p[3 * Wp] = (uint8_t)i0; p[4 * Wp] = (uint8_t)ilo; p[5 * Wp] = (uint8_t)ihi;
}

// ------------------------------------------------------------------------------------------------
// K2: pixel cost (A.3) + blockSize x blockSize box sum (A.4) -> Ctrue
//   one CTA = TW output columns x a band of rows x all D; it walks down the band keeping the last
//   2r+2 rows of pixel costs (u8) in a shared-memory ring and the running vertical sums Vs in
//   shared memory; the horizontal window is summed from Vs.  Threads are (pixel slot, d-lane):
//   a lane owns 8 consecutive disparities = 4 packed registers = one 16-byte store.
// ------------------------------------------------------------------------------------------------
static constexpr int kCostTW = 32;
static constexpr int kCostThreads = 256;

struct CostSmem {
  int NC, NR, NRp, RS;
  size_t offLeft, offRA, offRB, offRing, offVs, total;
};
__host__ __device__ inline CostSmem cost_smem_layout(int D, int r) {
  CostSmem s;
  s.NC = kCostTW + 2 * r;
  s.NR = s.NC + D - 1;
  s.NRp = (s.NR + 3) & ~1;  // even, >= NR+1
  s.RS = 2 * r + 2;
  size_t o = 0;
  s.offLeft = o; o += (size_t)s.NC * 6 * 4;
  s.offRA = o; o += (size_t)6 * s.NRp * 2;
  s.offRB = o; o += (size_t)6 * s.NRp * 2;
  o = (o + 15) & ~(size_t)15;
  s.offVs = o; o += (size_t)s.NC * D * 2;
  s.offRing = o; o += (size_t)s.RS * s.NC * D;
  s.total = o;
  return s;
}

__device__ __forceinline__ uint32_t bt_pair(uint32_t u, uint32_t ulo, uint32_t uhi, uint32_t v, uint32_t vlo, uint32_t vhi) {
  const uint32_t c0 = __vimax_s16x2_relu(__vsub2(u, vhi), __vsub2(vlo, u));
  const uint32_t c1 = __vimax_s16x2_relu(__vsub2(v, uhi), __vsub2(ulo, v));
  return __vminu2(c0, c1);
}

__global__ void __launch_bounds__(kCostThreads) k_cost_volume(const uint8_t* __restrict__ planesL,
                                                              const uint8_t* __restrict__ planesR,
                                                              uint16_t* __restrict__ Cvol, SgbmCfg cfg, int rowsPerBand,
                                                              size_t frameCells) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int D = cfg.D, r = cfg.r, W1 = cfg.W1, H = cfg.H, Wp = cfg.Wp;
  const CostSmem L = cost_smem_layout(D, r);
  uint32_t* sLeft = (uint32_t*)(smem + L.offLeft);
  uint16_t* sRA = (uint16_t*)(smem + L.offRA);
  uint16_t* sRB = (uint16_t*)(smem + L.offRB);
  uint16_t* sVs = (uint16_t*)(smem + L.offVs);
  uint8_t* sRing = smem + L.offRing;
  const int NC = L.NC, NR = L.NR, NRp = L.NRp, RS = L.RS;

  const int frame = blockIdx.z;
  const int x0 = blockIdx.x * kCostTW;  // first output column x' of the tile
  const int yb = blockIdx.y * rowsPerBand, ye = min(yb + rowsPerBand, H);
  const uint8_t* pL = planesL + (size_t)frame * H * 6 * Wp;
  const uint8_t* pR = planesR + (size_t)frame * H * 6 * Wp;
  uint16_t* Cf = Cvol + (size_t)frame * frameCells;

  const int G = D >> 3;                 // lanes per pixel
  const int nslot = kCostThreads / G;   // pixel slots per CTA
  const int slot = threadIdx.x / G, dl = threadIdx.x % G, d0 = dl * 8;
  const bool live = slot < nslot;
  // largest right-image x addressed by this tile; reversed index j = xrMax - xr = (NC-1-c) + d
  const int xrMax = x0 + kCostTW - 1 + r + cfg.minX1 - cfg.minD;

  for (int i = threadIdx.x; i < NC * (D >> 3); i += kCostThreads) ((uint4*)sVs)[i] = make_uint4(0, 0, 0, 0);

  int last = -1;  // last image row whose pixel costs are in the ring
  // rows [yb-1-r, yb-1+r] initialise Vs for the virtual row yb-1, then one row enters/leaves per step
  for (int step = -(2 * r + 1); step < ye - yb; step++) {
    const bool prologue = step < 0;
    const int y = yb + step;  // output row (main loop only)
    const int addRow = prologue ? min(max(yb - 1 - r + (step + 2 * r + 1), 0), H - 1) : min(y + r, H - 1);
    const int subRow = prologue ? -1 : min(max(y - r - 1, 0), H - 1);
    if (addRow > last || last < 0) {
      // ---- stage the pre-filter planes of row addRow
      __syncthreads();  // previous users of sLeft/sR are done
      const int yr = (last < 0) ? addRow : last + 1;  // rows enter in order; (addRow == last+1 by construction)
      const uint8_t* rowL = pL + (size_t)yr * 6 * Wp;
      const uint8_t* rowR = pR + (size_t)yr * 6 * Wp;
      for (int i = threadIdx.x; i < NC * 6; i += kCostThreads) {
        const int c = i / 6, k = i % 6;
        const int xp = x0 - r + c;
        uint32_t v = 0;
        if (xp >= 0 && xp < W1) v = rowL[(size_t)k * Wp + xp + cfg.minX1];
        sLeft[i] = dup16(v);
      }
      for (int i = threadIdx.x; i < NR * 6; i += kCostThreads) {
        const int k = i / NR, j = i % NR;
        const int xr = xrMax - j;
        uint16_t v = 0;
        if (xr >= 0 && xr < cfg.W) v = rowR[(size_t)k * Wp + xr];
        sRA[k * NRp + j] = v;
        if (j > 0) sRB[k * NRp + j - 1] = v;
      }
      __syncthreads();
      // ---- pixel costs of row yr for this thread's cells -> ring slot yr % RS
      if (live) {
        uint8_t* ring = sRing + (size_t)(yr % RS) * NC * D;
        for (int c = slot; c < NC; c += nslot) {
          const int xp = x0 - r + c;
          if (xp < 0 || xp >= W1) continue;
          const uint32_t* lp = sLeft + c * 6;
          const uint32_t ug = lp[0], uglo = lp[1], ughi = lp[2], ui = lp[3], uilo = lp[4], uihi = lp[5];
          const int j0 = (NC - 1 - c) + d0;
          const uint32_t* rp = (const uint32_t*)((j0 & 1) ? sRB : sRA) + (j0 >> 1);
          uint32_t pix[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint32_t cg = bt_pair(ug, uglo, ughi, rp[q], rp[(NRp >> 1) + q], rp[2 * (NRp >> 1) + q]);
            const uint32_t ci = bt_pair(ui, uilo, uihi, rp[3 * (NRp >> 1) + q], rp[4 * (NRp >> 1) + q], rp[5 * (NRp >> 1) + q]);
            pix[q] = cg + ((ci >> 2) & 0x3FFF3FFFu);
          }
          uint2 packed;
          packed.x = __byte_perm(pix[0], pix[1], 0x6420);
          packed.y = __byte_perm(pix[2], pix[3], 0x6420);
          *(uint2*)(ring + (size_t)c * D + d0) = packed;
        }
      }
      last = yr;
    }
    // ---- Vs += ring[addRow] - ring[subRow] for this thread's own cells (no barrier needed: same owner)
    if (live) {
      const uint8_t* ra = sRing + (size_t)(addRow % RS) * NC * D;
      const uint8_t* rs = subRow >= 0 ? sRing + (size_t)(subRow % RS) * NC * D : nullptr;
      for (int c = slot; c < NC; c += nslot) {
        const int xp = x0 - r + c;
        if (xp < 0 || xp >= W1) continue;
        uint4* vp = (uint4*)(sVs + (size_t)c * D + d0);
        uint4 v = *vp;
        const uint2 a = *(const uint2*)(ra + (size_t)c * D + d0);
        v.x += __byte_perm(a.x, 0, 0x4140); v.y += __byte_perm(a.x, 0, 0x4342);
        v.z += __byte_perm(a.y, 0, 0x4140); v.w += __byte_perm(a.y, 0, 0x4342);
        if (rs) {
          const uint2 s = *(const uint2*)(rs + (size_t)c * D + d0);
          v.x -= __byte_perm(s.x, 0, 0x4140); v.y -= __byte_perm(s.x, 0, 0x4342);
          v.z -= __byte_perm(s.y, 0, 0x4140); v.w -= __byte_perm(s.y, 0, 0x4342);
        }
        *vp = v;
      }
    }
    if (prologue) continue;
    __syncthreads();
    // ---- horizontal window over Vs (clamped in W1 coordinates, A.4) -> Ctrue[y][x'][16-group of d]
    //      two neighbouring lanes hold one group of 16 disparities; they swap halves so that word i of the
    //      group is (d0+i, d0+i+8) — the interleaved volume layout of sgbm_device.cuh
    {
      const bool hh4zero = (cfg.mode == SV_MODE_HH4) && (y >= H - r);
      const int trips = (kCostTW + nslot - 1) / nslot;
      for (int it = 0; it < trips; it++) {
        const int t = slot + it * nslot;
        const int xp = x0 + t;
        const bool valid = live && t < kCostTW && xp < W1;
        uint4 acc = make_uint4(0, 0, 0, 0);
        if (valid && !hh4zero) {
          for (int dx = -r; dx <= r; dx++) {
            const int cc = min(max(xp + dx, 0), W1 - 1) - (x0 - r);
            const uint4 v = *(const uint4*)(sVs + (size_t)cc * D + d0);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
        }
        const bool upper = dl & 1;  // this lane holds d0+8..d0+15 of the group
        const uint32_t s0 = __shfl_xor_sync(0xFFFFFFFFu, upper ? acc.x : acc.z, 1);
        const uint32_t s1 = __shfl_xor_sync(0xFFFFFFFFu, upper ? acc.y : acc.w, 1);
        uint4 o;
        if (!upper) {  // words 0..3: (d, d+8) for d = d0..d0+3
          o.x = __byte_perm(acc.x, s0, 0x5410); o.y = __byte_perm(acc.x, s0, 0x7632);
          o.z = __byte_perm(acc.y, s1, 0x5410); o.w = __byte_perm(acc.y, s1, 0x7632);
        } else {       // words 4..7: (d-8, d) for d = d0..d0+3 of this lane, i.e. group offsets 4..7
          o.x = __byte_perm(s0, acc.z, 0x5410); o.y = __byte_perm(s0, acc.z, 0x7632);
          o.z = __byte_perm(s1, acc.w, 0x5410); o.w = __byte_perm(s1, acc.w, 0x7632);
        }
        if (valid) *(uint4*)(Cf + ((size_t)y * W1 + xp) * D + d0) = o;
      }
    }
    __syncthreads();
  }
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

#define SV_RET(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)
#define SV_LAUNCHED() do { if (stats) stats->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t sgbm_forward(const SgbmCfg& cfg, const SgbmScratch& sc, int nframes, const uint8_t* left, size_t lstride,
                         size_t lframe, const uint8_t* right, size_t rstride, size_t rframe, int16_t* disp,
                         size_t dstride, size_t dframe, cudaStream_t st, LaunchStats* stats, cudaEvent_t* ev, int debugS) {
  const int W = cfg.W, H = cfg.H;
  if (cfg.W1 <= 0) {  // A.1: nothing computable, whole map invalid
    dim3 g((W + 255) / 256, H, nframes);
    k_fill_rows_s16<<<g, 256, 0, st>>>(disp, dstride, dframe, W, H, cfg.invalid);
    SV_LAUNCHED();
    return cudaSuccess;
  }
  if (ev) SV_RET(cudaEventRecord(ev[0], st));
  {
    dim3 g((W + 255) / 256, H, nframes * 2);
    k_prefilter_planes<<<g, 256, 0, st>>>(left, lstride, lframe, right, rstride, rframe, sc.planesL, sc.planesR, W, H,
                                          cfg.Wp, cfg.ftzero);
    SV_LAUNCHED();
  }
  {
    const CostSmem L = cost_smem_layout(cfg.D, cfg.r);
    static size_t configured = 0;
    if (L.total > configured) {
      SV_RET(cudaFuncSetAttribute(k_cost_volume, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
      configured = L.total;
    }
    const int tiles = (cfg.W1 + kCostTW - 1) / kCostTW;
    // enough bands to fill the GPU a few times over, but not so many that the 2r+1-row prologue dominates
    int bands = (148 * 4 + tiles * nframes - 1) / (tiles * nframes);
    bands = max(1, min(bands, (H + 31) / 32));
    const int rowsPerBand = (H + bands - 1) / bands;
    bands = (H + rowsPerBand - 1) / rowsPerBand;
    dim3 g(tiles, bands, nframes);
    k_cost_volume<<<g, kCostThreads, L.total, st>>>(sc.planesL, sc.planesR, sc.C, cfg, rowsPerBand, sc.frameCells);
    SV_LAUNCHED();
  }
  if (ev) SV_RET(cudaEventRecord(ev[1], st));
  SV_RET(sgbm_aggregate_sweep(cfg, sc, nframes, debugS, st, stats));
  if (ev) SV_RET(cudaEventRecord(ev[2], st));
  {
    dim3 g(H, nframes);
    k_select_lr<<<g, 256, (size_t)W * 6, st>>>(sc.sel, sc.dispRaw, cfg);
    SV_LAUNCHED();
    SV_RET(launch_median3(sc.dispRaw, (size_t)W * 2, disp, dstride, W, H, nframes, (size_t)W * H * 2, dframe, st));
    if (stats) stats->launches++;
  }
  if (ev) SV_RET(cudaEventRecord(ev[3], st));
  return cudaSuccess;
}

}  // namespace svgpu

// sgbm_device.cuh — device building blocks shared by the SGBM aggregation kernels (sgbm.cu, sgbm_raster.cu):
// one min-plus path step on packed 16-bit costs, and the per-pixel winner-take-all / uniqueness / sub-pixel
// selection.  Arithmetic: SURVEY.md Appendix A.5 / A.6 (cv::StereoSGBM::compute, src/lib.cpp:262).
#pragma once
#include "common.cuh"

namespace svgpu {

static constexpr uint32_t kMax2 = 0x7FFF7FFFu;  // two lanes of MAX_COST (32767)

__device__ __forceinline__ uint32_t dup16(uint32_t v) { return v | (v << 16); }
__device__ __forceinline__ uint32_t hmin16(uint32_t v) { return min(v & 0xFFFFu, v >> 16); }

// ------------------------------------------------------------------------------------------------
// Path aggregation (A.5).  One min-plus step of one path for 8 disparities per lane.
//   L      previous / next path costs of this lane (4 packed registers, u16x2)
//   mdup   min over all d of the previous L, duplicated in both halves
//   c      Ctrue of the current pixel
// Neighbouring lanes exchange their border disparities with two shuffles; lanes beyond D hold
// MAX_COST so that they act as the L(q,-1) = L(q,D) = 32767 boundary.
// ------------------------------------------------------------------------------------------------
template <int LPP>
__device__ __forceinline__ void sgm_step(uint32_t (&L)[4], uint32_t& mdup, const uint4& c, uint32_t p1dup,
                                         uint32_t p2dup, unsigned gmask, int gl, bool active) {
  uint32_t up = __shfl_up_sync(0xFFFFFFFFu, L[3], 1, LPP);
  uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, L[0], 1, LPP);
  if (gl == 0) up = kMax2;
  if (gl == LPP - 1) dn = kMax2;
  const uint32_t mp2 = mdup + p2dup;
  const uint32_t cc[4] = {c.x, c.y, c.z, c.w};
  uint32_t prev = up, n[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t next = (i < 3) ? L[i + 1] : dn;
    const uint32_t lm1 = __byte_perm(prev, L[i], 0x5432);  // L(q,d-1) for the pair
    const uint32_t lp1 = __byte_perm(L[i], next, 0x5432);  // L(q,d+1)
    uint32_t t = __viaddmin_u16x2(__vminu2(lm1, lp1), p1dup, mp2);
    t = __vminu2(t, L[i]);
    n[i] = __vadd2(t, __vsub2(cc[i], mdup));
    prev = L[i];
  }
  uint32_t mn = __vminu2(__vminu2(n[0], n[1]), __vminu2(n[2], n[3]));
#pragma unroll
  for (int i = 0; i < 4; i++) L[i] = active ? n[i] : kMax2;
  uint32_t m = active ? hmin16(mn) : 0x7FFFu;
  m = __reduce_min_sync(gmask, m);
  mdup = dup16(m);
}

// first pixel of a path: predecessor outside the volume, L = Ctrue
__device__ __forceinline__ void sgm_first(uint32_t (&L)[4], uint32_t& mdup, const uint4& c, unsigned gmask, bool active) {
  L[0] = active ? c.x : kMax2; L[1] = active ? c.y : kMax2; L[2] = active ? c.z : kMax2; L[3] = active ? c.w : kMax2;
  uint32_t m = hmin16(__vminu2(__vminu2(L[0], L[1]), __vminu2(L[2], L[3])));
  m = __reduce_min_sync(gmask, m);
  mdup = dup16(m);
}

// ------------------------------------------------------------------------------------------------
// K4: winner-take-all + uniqueness + sub-pixel (A.6, first half) for one pixel held by a lane group
//   s = final S of the pixel (this lane's 8 disparities).  Returns the raw CV_16S value (or invalid)
//   and the key (minS<<16)|best (or ~0) that feeds the left-right check.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pick16(const uint32_t (&s)[4], int e) {  // element e (0..7) of the lane
  const uint32_t w = (e & 4) ? ((e & 2) ? s[3] : s[2]) : ((e & 2) ? s[1] : s[0]);
  return (e & 1) ? (w >> 16) : (w & 0xFFFFu);
}

template <int LPP>
__device__ __forceinline__ void wta_pixel(const uint32_t (&s)[4], int gl, unsigned gmask, bool active,
                                          const SgbmCfg& cfg, int& dispOut, uint32_t& keyOut) {
  const int d0 = gl * 8;
  uint32_t k = 0xFFFFFFFFu;
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      k = min(k, (s[i] << 16) | (uint32_t)(d0 + 2 * i));
      k = min(k, (s[i] & 0xFFFF0000u) | (uint32_t)(d0 + 2 * i + 1));
    }
  }
  k = __reduce_min_sync(gmask, k);
  const int minS = (int)(k >> 16), best = (int)(k & 0xFFFFu);
  dispOut = cfg.invalid;
  keyOut = 0xFFFFFFFFu;
  // neighbours of the winner for the parabola (owners publish, everybody reduces)
  const int e1 = best - 1, e2 = best + 1;
  uint32_t vm = 0, vp = 0;
  if (active && e1 >= d0 && e1 < d0 + 8) vm = pick16(s, e1 - d0);
  if (active && e2 >= d0 && e2 < d0 + 8) vp = pick16(s, e2 - d0);
  vm = __reduce_max_sync(gmask, vm);
  vp = __reduce_max_sync(gmask, vp);
  // uniqueness: any d with S[d]*(100-uniq) < minS*100 and |best-d| > 1 rejects the pixel
  bool rej = false;
  if (active) {
    const int f = 100 - cfg.uniq, t = minS * 100;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int a = (int)(s[i] & 0xFFFFu), b = (int)(s[i] >> 16), da = d0 + 2 * i;
      rej |= (a * f < t) && (abs(best - da) > 1);
      rej |= (b * f < t) && (abs(best - da - 1) > 1);
    }
  }
  const unsigned rejBallot = __ballot_sync(0xFFFFFFFFu, rej) & gmask;
  if (minS >= 32767 || rejBallot) return;
  int dfix = best * 16;
  if (best > 0 && best < cfg.D - 1) {
    const int den = max((int)vm + (int)vp - 2 * minS, 1);
    dfix += (((int)vm - (int)vp) * 16 + den) / (den * 2);
  }
  dispOut = dfix + cfg.minD * 16;
  keyOut = k;
}

}  // namespace svgpu

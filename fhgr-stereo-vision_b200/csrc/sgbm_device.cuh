// sgbm_device.cuh — device building blocks of the SGBM aggregation sweep (sgbm_sweep.cu): one min-plus path
// step on packed 16-bit costs and the per-pixel winner-take-all / uniqueness selection.
// Arithmetic: SURVEY.md Appendix A.5 / A.6 (cv::StereoSGBM::compute, /root/reference/src/lib.cpp:262).
//
// Register layout of one pixel's costs: a worker is a group of LPP lanes, lane gl holds the 16 disparities
// d0 = 16*gl .. d0+15 in kNR = 8 packed registers, register i = (d0+i) | (d0+i+8) << 16.  With that
// interleave the d-1 / d+1 neighbours of register i are registers i-1 / i+1 of the same lane; only
// register 0 and register 7 need one value from the neighbouring lane.
#pragma once
#include "common.cuh"

namespace svgpu {

static constexpr uint32_t kMax2 = 0x7FFF7FFFu;  // two lanes of MAX_COST (32767)
static constexpr int kDPL = 16;                 // disparities per lane
static constexpr int kNR = kDPL / 2;            // packed registers per lane and path

__host__ __device__ __forceinline__ uint32_t dup16(uint32_t v) { return v | (v << 16); }

// position of disparity d inside its pixel's D costs in volume memory (uint16 index)
__host__ __device__ __forceinline__ int vol_index(int d) { return (d & ~15) | ((d & 7) << 1) | ((d >> 3) & 1); }

// compile-time tuning knobs of the path step (tools/build_variants.sh builds alternatives for A/B runs on the GPU)
#ifndef SV_SWEEP_EDGEMAD
#define SV_SWEEP_EDGEMAD 1   // 1: the MAX_COST neighbours of a worker's first / last lane through a multiply-add (FMA pipe) instead of a select (ALU pipe)
#endif
#ifndef SV_SWEEP_FMAADD
#define SV_SWEEP_FMAADD 0    // n: the first n paths of a pixel form y + C - m with two multiply-adds (FMA pipe) instead of one three-input add (ALU pipe)
#endif
#ifndef SV_SWEEP_MINPAIR
#define SV_SWEEP_MINPAIR 1   // 1: paired path minima fold the two halves of the lane minimum while packing (2 instructions fewer per pair)
#endif

struct SweepK {
  uint32_t p1dup;     // P1 | P1<<16
  uint32_t p2mp1dup;  // (P2-P1) twice
  uint32_t negp1dup;  // (-P1 mod 2^16) twice
  // lane constants: the value from the neighbouring lane is v * keep + add (keep = 1, add = 0 inside the worker; keep = 0, add = two
  // MAX_COST for the lane without that neighbour), and a register holding 1 that the compiler cannot see through
  uint32_t keepUp, addUp, keepDn, addDn, one;
};
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
  asm volatile("mov.b32 %0, %0;" : "+r"(v));
  return v;
}
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// min / sum over the LPP lanes of a worker; every lane of the warp must call
template <int LPP>
__device__ __forceinline__ uint32_t group_min_u32(uint32_t v, int lane) {
  if constexpr (LPP == 1) return v;
  else if constexpr (LPP == 32) return __reduce_min_sync(0xFFFFFFFFu, v);
  else if constexpr (LPP == 16) {  // two full-warp CREDUX beat a four-step shuffle butterfly
    const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, lane < 16 ? v : 0xFFFFFFFFu);
    const uint32_t hi = __reduce_min_sync(0xFFFFFFFFu, lane < 16 ? 0xFFFFFFFFu : v);
    return lane < 16 ? lo : hi;
  } else {
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
  }
}
template <int LPP>
__device__ __forceinline__ uint32_t group_sum_u32(uint32_t v, int lane) {
  if constexpr (LPP == 1) return v;
  else if constexpr (LPP == 32) return __reduce_add_sync(0xFFFFFFFFu, v);
  else if constexpr (LPP == 16) {
    const uint32_t lo = __reduce_add_sync(0xFFFFFFFFu, lane < 16 ? v : 0u);
    const uint32_t hi = __reduce_add_sync(0xFFFFFFFFu, lane < 16 ? 0u : v);
    return lane < 16 ? lo : hi;
  } else {
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
  }
}

// ------------------------------------------------------------------------------------------------
// Path aggregation (A.5): one step of one path for the 16 disparities of this lane.
//   L      in: path costs of the predecessor pixel, out: path costs of this pixel
//   mdup   in: min_d L(predecessor) in both halves, out: the same for this pixel
//   outside  the predecessor lies outside the volume (L = 0, min = 0): the result is Ctrue.  This is
//          expressed through the two scalars only, so L may then hold any values in [0, 32767]
//   c      Ctrue of this pixel
// L(p,d) = C + min(L(q,d), L(q,d-1)+P1, L(q,d+1)+P1, m+P2) - m, evaluated as
//   x = min3_s16(L(d-1), L(d+1), m+P2-P1);  y = min_u16(x+P1, L(d));  L' = y + C + (-m)
// (3 instructions per packed register; no half ever borrows or carries because m <= y and y + C < 2^16).
// ------------------------------------------------------------------------------------------------
// sweep_path_core does the step and returns the minimum over this lane's 16 disparities (in both halves; all ones for a lane
// without disparities); the caller reduces it over the worker's lanes — alone (sweep_path_step) or two paths at a time in
// the two halves of one register (group_min2_u16), which halves the number of shuffles.
// FMA2: form the new path cost with two multiply-adds; HALVES: return the lane minimum per half (low: d0..d0+7, high: d0+8..d0+15)
// instead of the minimum over all 16 in both halves
template <int LPP, bool FMA2 = false, bool HALVES = false>
__device__ __forceinline__ uint32_t sweep_path_core(uint32_t (&L)[kNR], uint32_t mdup, bool outside, const uint32_t (&c)[kNR],
                                                    const SweepK& k, int gl, int lastLane, bool active, int lane) {
  uint32_t up = kMax2, dn = kMax2;
  if (LPP > 1) {
    up = __shfl_up_sync(0xFFFFFFFFu, L[kNR - 1], 1, LPP);
    dn = __shfl_down_sync(0xFFFFFFFFu, L[0], 1, LPP);
#if SV_SWEEP_EDGEMAD
    up = mad_u32(up, k.keepUp, k.addUp);
    dn = mad_u32(dn, k.keepDn, k.addDn);
#else
    if (gl == 0) up = kMax2;         // L(q,-1) = MAX_COST
    if (gl == lastLane) dn = kMax2;  // L(q,D)  = MAX_COST
#endif
  }
  const uint32_t capm = outside ? k.negp1dup : mdup + k.p2mp1dup;
  const uint32_t negm = outside ? 0u : 0u - mdup;
  const uint32_t lm0 = __byte_perm(up, L[kNR - 1], 0x5432);  // (d0-1, d0+7)
  const uint32_t lpN = __byte_perm(L[0], dn, 0x5432);        // (d0+8, d0+16)
  uint32_t n[kNR];
#pragma unroll
  for (int i = 0; i < kNR; i++) {
    const uint32_t lm = i ? L[i - 1] : lm0, lp = (i < kNR - 1) ? L[i + 1] : lpN;
    const uint32_t x = __vimin3_s16x2(lm, lp, capm);
    const uint32_t y = __viaddmin_u16x2(x, k.p1dup, L[i]);
    if constexpr (FMA2) n[i] = mad_u32(mad_u32(y, k.one, c[i]), k.one, negm);
    else n[i] = y + c[i] + negm;
  }
  uint32_t mn = __vimin3_u16x2(n[0], n[1], n[2]);
  mn = __vimin3_u16x2(mn, n[3], n[4]);
  mn = __vimin3_u16x2(mn, n[5], n[6]);
  mn = __vminu2(mn, n[7]);
  if constexpr (!HALVES) mn = __vminu2(mn, __byte_perm(mn, mn, 0x1032));
  if (!active) mn = 0xFFFFFFFFu;
#pragma unroll
  for (int i = 0; i < kNR; i++) L[i] = n[i];
  return mn;
}
template <int LPP>
__device__ __forceinline__ void sweep_path_step(uint32_t (&L)[kNR], uint32_t& mdup, bool outside, const uint32_t (&c)[kNR],
                                                const SweepK& k, int gl, int lastLane, bool active, int lane) {
  mdup = group_min_u32<LPP>(sweep_path_core<LPP>(L, mdup, outside, c, k, gl, lastLane, active, lane), lane);
}
// minima of two paths over the worker's lanes with one butterfly: a and b hold a lane-local minimum in both halves (HALVES: one
// minimum per half, folded here while the two paths are packed into one register)
template <int LPP, bool HALVES = false>
__device__ __forceinline__ void group_min2_u16(uint32_t a, uint32_t b, int lane, uint32_t& ma, uint32_t& mb) {
  if constexpr (LPP == 16) {  // two full-warp CREDUX per path (see group_min_u32)
    if constexpr (HALVES) { a = __vminu2(a, __byte_perm(a, a, 0x1032)); b = __vminu2(b, __byte_perm(b, b, 0x1032)); }
    ma = group_min_u32<LPP>(a, lane);
    mb = group_min_u32<LPP>(b, lane);
  } else {
    uint32_t p = __byte_perm(a, b, 0x5410);  // low half: path a, high half: path b
    if constexpr (HALVES) p = __vminu2(p, __byte_perm(a, b, 0x7632));
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) p = __vminu2(p, __shfl_xor_sync(0xFFFFFFFFu, p, o));
    ma = __byte_perm(p, 0, 0x1010);
    mb = __byte_perm(p, 0, 0x3232);
  }
}

// ------------------------------------------------------------------------------------------------
// Winner-take-all + uniqueness of one pixel (A.6, first half), run by the worker that holds its final S.
//   out.x = reject << 31 | minS << 16 | best        out.y = S[best-1] << 16 | S[best+1]
// Uniqueness ("some d with S[d]*(100-uniq) < minS*100 and |d-best| > 1") is decided without a per-cell
// compare: with thr = ceil(minS*100 / (100-uniq)) a cell qualifies iff S[d] < thr; the summed shortfall
// sum_d max(thr - S[d], 0) exceeds the shortfall of {best-1, best, best+1} iff another cell qualifies.
// `scratch` holds the pixel's S so that the two neighbours of the winner can be looked up by index.
// ------------------------------------------------------------------------------------------------
// WAY3 selects the semantics of OpenCV's MODE_SGBM_3WAY (oracle: wta_lr_row with simd8): among equal minima the winner is,
// per residue class d mod 8, the LAST disparity, and among the classes the smallest of those; the uniqueness threshold is
// (short)(100*minS/(100-uniq) + 1) instead of the cross-multiplied comparison (uniqMagic = ceil(2^32 / (100-uniq)) is then
// used for a floor division).  Register i of a lane holds d0+i and d0+i+8, i.e. exactly one residue class.
template <int LPP, bool WAY3 = false>
__device__ __forceinline__ void sweep_wta(const uint32_t (&s)[kNR], uint16_t* scratch, int gl, bool active, int lane, int D,
                                          int uniqF, uint32_t uniqMagic, uint2& out) {
  const int d0 = gl * kDPL;
  uint32_t k = 0xFFFFFFFFu;
  if (active) {
#pragma unroll
    for (int i = 0; i < kNR; i += 2) {
      const uint32_t a = s[i] * 65536u + (uint32_t)(d0 + i), b = (s[i] & 0xFFFF0000u) | (uint32_t)(d0 + i + 8);
      const uint32_t e = s[i + 1] * 65536u + (uint32_t)(d0 + i + 1), f = (s[i + 1] & 0xFFFF0000u) | (uint32_t)(d0 + i + 9);
      k = __vimin3_u32(k, a, b);
      k = __vimin3_u32(k, e, f);
    }
  }
  k = group_min_u32<LPP>(k, lane);
  if constexpr (WAY3) {
    // per residue class (= register) the largest disparity that attains the minimum, then the smallest over the classes
    const uint32_t m = k >> 16;
    uint32_t bestd = 0xFFFFu;
#pragma unroll
    for (int i = 0; i < kNR; i++) {
      uint32_t c = 0;   // candidate + 1 (0: class has no minimum in this lane)
      if (active) {
        if ((s[i] & 0xFFFFu) == m) c = (uint32_t)(d0 + i + 1);
        if ((s[i] >> 16) == m) c = (uint32_t)(d0 + i + 9);
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xFFFFFFFFu, c, o));
      if (c) bestd = min(bestd, c - 1);
    }
    k = (m << 16) | bestd;
  }
  const int minS = (int)(k >> 16), best = (int)(k & 0xFFFFu);
  // publish S of the pixel, look up the neighbours of the winner
  __syncwarp();
  if (active) {
    uint4* sp = (uint4*)scratch + gl * 2;
    sp[0] = make_uint4(s[0], s[1], s[2], s[3]);
    sp[1] = make_uint4(s[4], s[5], s[6], s[7]);
  }
  __syncwarp();
  const int vm = best > 0 ? (int)scratch[vol_index(best - 1)] : 0;
  const int vp = best < D - 1 ? (int)scratch[vol_index(best + 1)] : 0;
  // shortfall against thr
  uint32_t thr;
  if constexpr (WAY3) {
    // floor(100*minS / uniqF) + 1, truncated to a signed 16-bit value; a negative threshold matches nothing.  uniqF = 100
    // (uniquenessRatio 0) switches the test off.
    uint32_t q = (uint32_t)minS * 100u / (uint32_t)max(uniqF, 1);
    const int t16 = (int)(int16_t)(uint16_t)(q + 1u);
    thr = (uniqF >= 100 || t16 <= 0) ? 0u : (uint32_t)t16;
  } else {
    const uint32_t num = (uint32_t)minS * 100u + (uint32_t)uniqF - 1u;
    const uint32_t q = uniqF >= 2 ? __umulhi(num, uniqMagic) : (uniqF == 1 ? num : (minS ? 32768u : 0u));
    thr = min(q, 32768u);
  }
  const uint32_t thrdup = dup16(thr);
  uint32_t acc = 0;
  if (active) {
#pragma unroll
    for (int i = 0; i < kNR; i++) acc = __dp2a_lo(__vminu2(s[i], thrdup), 0x0101u, acc);
  }
  acc = group_sum_u32<LPP>(acc, lane);
  const int total = D * (int)thr - (int)acc;
  int allowed = max((int)thr - minS, 0);
  if (best > 0) allowed += max((int)thr - vm, 0);
  if (best < D - 1) allowed += max((int)thr - vp, 0);
  const uint32_t rej = total > allowed ? 0x80000000u : 0u;
  out.x = k | rej;
  out.y = ((uint32_t)vm << 16) | (uint32_t)vp;
}

}  // namespace svgpu

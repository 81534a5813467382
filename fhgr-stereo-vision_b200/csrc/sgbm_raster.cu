// sgbm_raster.cu — SGM path aggregation as two fused raster sweeps (aggregation schedule 1).
//
// Replaces the aggregation inside cv::StereoSGBM::compute (/root/reference/src/lib.cpp:262).
// A raster-order sweep can carry the four paths whose predecessor lies in the previous row or to the left
// (travel directions ->, down-right, down, down-left); the reverse raster sweep carries the other four.  So
// MODE_HH needs two passes over the cost volume: pass 1 reads C and writes S = sum of 4 paths (4 B/cell),
// pass 2 reads C and S, adds its 4 paths and runs winner-take-all in registers (4 B/cell) — the 8 B/cell
// aggregation traffic of SURVEY.md §8d instead of 46 B/cell for one kernel per direction.
//
// Parallelism comes from a skewed wavefront: image row y may process pixel x once row y-1 has finished
// pixel x+1, so consecutive rows run two pixels apart.
//   * a "row worker" is a group of LPP lanes (8 disparities per lane) that walks one image row; the
//     horizontal path lives in its registers;
//   * a warp holds RW = 32/LPP consecutive rows in lock step, skewed by two pixels per row;
//   * a CTA is kNW such warps (a band of kNW*RW rows) plus one helper warp;
//   * the three paths that come from the row above travel through small shared-memory rings (one per
//     row, written by the row above); warps hand over with acquire/release progress counters in shared
//     memory, never with a CTA-wide barrier;
//   * bands hand over through a global edge buffer ([frame][band][x][slot]) and a per-band progress
//     counter; the helper warp streams the previous band's last row into a shared-memory ring and this
//     band's last row out, so compute warps only ever touch shared memory;
//   * CTAs are persistent and draw (frame, band) tasks from an atomic ticket counter in band-major order.
//     A task only ever waits for a task with a smaller ticket, which is resident or finished, so the
//     scheme cannot deadlock whatever the residency.
// Pass 2 is pass 1 on the volume flipped in x and y (pure index arithmetic).
#include <algorithm>

#include "common.cuh"
#include "sgbm_device.cuh"

namespace svgpu {

namespace {

constexpr int kNW = 16;        // compute warps per CTA
constexpr int kRingIn = 4;     // ring depth of a row whose consumer is in the same warp (lag is exactly 2)
constexpr int kRingOut = 8;    // ring depth of the last row of a warp (consumer is another warp / the helper)
constexpr int kRingEdge = 8;   // ring depth of the virtual row above the band (filled by the helper)
constexpr int kPublish = 8;    // the helper publishes band progress to global memory every kPublish pixels
constexpr int kSpinLimit = 1 << 26;

enum : int { P_H = 1, P_DL = 2, P_V = 4, P_DR = 8 };

template <int LPP>
struct RasterGeom {
  static constexpr int RW = 32 / LPP;                                  // rows per warp
  static constexpr int RB = kNW * RW;                                  // rows per band (CTA)
  static constexpr int SLOT = 3 * LPP + 1;                             // uint4 per ring slot: V, DL, DR + mins
  static constexpr int WARP_SLOTS = (RW - 1) * kRingIn + kRingOut;     // ring slots owned by one warp
  static constexpr int RING_U4 = (kNW * WARP_SLOTS + kRingEdge) * SLOT;
  static constexpr int NFLAGS = 2 * kNW + 4;                           // prog[kNW], cons[kNW+1], progE, ticket, abort
  static constexpr size_t SMEM = (size_t)RING_U4 * 16 + NFLAGS * 4;
};

struct RasterArgs {
  const uint16_t* C;
  uint16_t* S;
  uint32_t* wtaKey;
  int16_t* dispRaw;
  uint4* edges;   // [frame][band][x][SLOT]
  int* gprog;     // [frame][band] pixels of the band's last row that are visible in `edges`
  int* ticket;
  int* status;
  SgbmCfg cfg;
  int nframes, nbands, flip;
  size_t frameCells;
};

__device__ __forceinline__ int ld_acquire_cta(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta(int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// spin until *flag >= need; false when the kernel is being aborted
__device__ __forceinline__ bool wait_cta(const int* flag, int need, volatile int* sAbort, int* status) {
  int n = 0;
  while (ld_acquire_cta(flag) < need) {
    if (((++n) & 1023) == 0) {
      if (*sAbort) return false;
      if (n > kSpinLimit) { *sAbort = 1; atomicExch(status, 1); return false; }
    }
  }
  return true;
}

template <int LPP, int PATHS, bool SECOND>
__global__ void __launch_bounds__((kNW + 1) * 32, 1) k_agg_raster(RasterArgs a) {
  using G = RasterGeom<LPP>;
  constexpr int RW = G::RW, SLOT = G::SLOT;
  constexpr bool VERT = (PATHS & (P_DL | P_V | P_DR)) != 0;
  extern __shared__ uint4 smemU4[];
  uint4* rings = smemU4;
  int* flags = (int*)(smemU4 + G::RING_U4);
  int* sProg = flags;                     // [kNW]   pixels finished by the last row of warp w
  int* sCons = flags + kNW;               // [kNW+1] pixels finished by the first row of warp w; [kNW] = helper store side
  int* sProgE = flags + 2 * kNW + 1;      // pixels of the row above the band that the helper has put into its ring
  int* sTicket = flags + 2 * kNW + 2;
  volatile int* sAbort = flags + 2 * kNW + 3;

  const SgbmCfg& cfg = a.cfg;
  const int W1 = cfg.W1, H = cfg.H, D = cfg.D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / LPP, gl = lane % LPP;
  const unsigned gmask = (LPP == 32) ? 0xFFFFFFFFu : (((1u << LPP) - 1u) << (lane - gl));
  const bool active = gl * 8 < D;
  const uint32_t p1dup = dup16(cfg.P1), p2dup = dup16(cfg.P2);
  const uint32_t idle = active ? 0u : kMax2;  // "predecessor outside the volume": L = 0, min = 0 (A.5)
  const int ntasks = a.nframes * a.nbands;
  if (threadIdx.x == 0) *sAbort = 0;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) *sTicket = atomicAdd(a.ticket, 1);
    if (threadIdx.x < 2 * kNW + 2) flags[threadIdx.x] = 0;
    __syncthreads();
    const int task = *sTicket;
    if (task >= ntasks || *sAbort) break;
    const int frame = task % a.nframes, band = task / a.nframes;

    if (warp < kNW) {
      // ------------------------------------------------------------------ compute warp: RW image rows
      const int yFirst = band * G::RB + warp * RW, yLast = yFirst + RW - 1;
      const int y = yFirst + g;  // logical (sweep-order) row of this worker
      const bool rowOn = y < H, firstOn = yFirst < H, lastOn = yLast < H;
      const int yp = a.flip ? H - 1 - y : y;
      const long long stepX = a.flip ? -(long long)D : (long long)D;
      const size_t rowOff = (size_t)frame * a.frameCells + ((size_t)(rowOn ? yp : 0) * W1 + (a.flip ? W1 - 1 : 0)) * D + gl * 8;
      const uint16_t* cRow = a.C + rowOff;
      uint16_t* sRow = a.S + rowOff;

      uint4* own = rings + (size_t)(warp * G::WARP_SLOTS + (g < RW - 1 ? g : RW - 1) * kRingIn) * SLOT;
      const int ownMask = (g < RW - 1) ? kRingIn - 1 : kRingOut - 1;
      const uint4* above;
      int aboveMask;
      if (g > 0) { above = rings + (size_t)(warp * G::WARP_SLOTS + (g - 1) * kRingIn) * SLOT; aboveMask = kRingIn - 1; }
      else if (warp > 0) { above = rings + (size_t)((warp - 1) * G::WARP_SLOTS + (RW - 1) * kRingIn) * SLOT; aboveMask = kRingOut - 1; }
      else { above = rings + (size_t)(kNW * G::WARP_SLOTS) * SLOT; aboveMask = kRingEdge - 1; }
      const bool topOutside = (y == 0);

      const int* aboveFlag = warp > 0 ? &sProg[warp - 1] : sProgE;
      const bool waitAbove = VERT && firstOn && yFirst > 0;
      const bool hasConsumer = VERT && lastOn && (warp < kNW - 1 ? (yLast + 1 < H) : (band < a.nbands - 1));
      const int* nextCons = &sCons[warp + 1];

      uint32_t LH[4] = {kMax2, kMax2, kMax2, kMax2}, mH = 0;
      uint4 cq0 = make_uint4(0, 0, 0, 0), cq1 = cq0, sq0 = cq0, sq1 = cq0;
      if (rowOn && active) {
        cq0 = *(const uint4*)cRow;
        if (W1 > 1) cq1 = *(const uint4*)(cRow + stepX);
        if (SECOND) {
          sq0 = *(const uint4*)sRow;
          if (W1 > 1) sq1 = *(const uint4*)(sRow + stepX);
        }
      }
      const int nsteps = W1 + 2 * (RW - 1);
      bool ok = true;
      for (int t = 0; t < nsteps && ok; t++) {
        const int x = t - 2 * g;
        const bool on = rowOn && x >= 0 && x < W1;
        // ---- hand-over with the neighbouring warps (warp-uniform)
        if (waitAbove && t < W1) ok = wait_cta(aboveFlag, min(t + 2, W1), sAbort, a.status);
        if (hasConsumer) {
          const int xl = t - 2 * (RW - 1);
          if (xl >= kRingOut - 1 && xl < W1 && ok) ok = wait_cta(nextCons, xl - kRingOut + 2, sAbort, a.status);
        }
        if (!ok) break;
        // ---- operands of this pixel: C (and S of pass 1) were prefetched two pixels ago
        uint4 c = cq0, sin = sq0;
        if (on) {
          cq0 = cq1;
          if (SECOND) sq0 = sq1;
          if (x + 2 < W1 && active) {
            cq1 = *(const uint4*)(cRow + (long long)(x + 2) * stepX);
            if (SECOND) sq1 = *(const uint4*)(sRow + (long long)(x + 2) * stepX);
          }
        }
        uint32_t LV[4] = {idle, idle, idle, idle}, LDL[4] = {idle, idle, idle, idle}, LDR[4] = {idle, idle, idle, idle};
        uint32_t mV = 0, mDL = 0, mDR = 0;
        if (VERT && on && !topOutside) {
          if (PATHS & P_V) {
            const uint4* s = above + (size_t)(x & aboveMask) * SLOT;
            if (active) { const uint4 v = s[0 * LPP + gl]; LV[0] = v.x; LV[1] = v.y; LV[2] = v.z; LV[3] = v.w; }
            mV = s[3 * LPP].x;
          }
          if ((PATHS & P_DL) && x > 0) {
            const uint4* s = above + (size_t)((x - 1) & aboveMask) * SLOT;
            if (active) { const uint4 v = s[1 * LPP + gl]; LDL[0] = v.x; LDL[1] = v.y; LDL[2] = v.z; LDL[3] = v.w; }
            mDL = s[3 * LPP].y;
          }
          if ((PATHS & P_DR) && x < W1 - 1) {
            const uint4* s = above + (size_t)((x + 1) & aboveMask) * SLOT;
            if (active) { const uint4 v = s[2 * LPP + gl]; LDR[0] = v.x; LDR[1] = v.y; LDR[2] = v.z; LDR[3] = v.w; }
            mDR = s[3 * LPP].z;
          }
        }
        if (x == 0) { LH[0] = LH[1] = LH[2] = LH[3] = idle; mH = 0; }
        // ---- the min-plus steps (A.5); lanes of rows that are not `on` compute on dummies
        uint32_t md = dup16(mH);
        sgm_step<LPP>(LH, md, c, p1dup, p2dup, gmask, gl, active);
        mH = md & 0xFFFFu;
        uint32_t s[4] = {LH[0], LH[1], LH[2], LH[3]};
        if (PATHS & P_V) {
          md = dup16(mV);
          sgm_step<LPP>(LV, md, c, p1dup, p2dup, gmask, gl, active);
          mV = md & 0xFFFFu;
#pragma unroll
          for (int i = 0; i < 4; i++) s[i] = __viaddmin_u16x2(s[i], LV[i], kMax2);
        }
        if (PATHS & P_DL) {
          md = dup16(mDL);
          sgm_step<LPP>(LDL, md, c, p1dup, p2dup, gmask, gl, active);
          mDL = md & 0xFFFFu;
#pragma unroll
          for (int i = 0; i < 4; i++) s[i] = __viaddmin_u16x2(s[i], LDL[i], kMax2);
        }
        if (PATHS & P_DR) {
          md = dup16(mDR);
          sgm_step<LPP>(LDR, md, c, p1dup, p2dup, gmask, gl, active);
          mDR = md & 0xFFFFu;
#pragma unroll
          for (int i = 0; i < 4; i++) s[i] = __viaddmin_u16x2(s[i], LDR[i], kMax2);
        }
        if (SECOND) {
          s[0] = __viaddmin_u16x2(s[0], sin.x, kMax2); s[1] = __viaddmin_u16x2(s[1], sin.y, kMax2);
          s[2] = __viaddmin_u16x2(s[2], sin.z, kMax2); s[3] = __viaddmin_u16x2(s[3], sin.w, kMax2);
          int dsp;
          uint32_t key;
          wta_pixel<LPP>(s, gl, gmask, active && on, cfg, dsp, key);
          if (on && gl == 0) {
            const int xp = a.flip ? W1 - 1 - x : x;
            a.wtaKey[((size_t)frame * H + yp) * W1 + xp] = key;
            a.dispRaw[((size_t)frame * H + yp) * cfg.W + xp + cfg.minX1] = (int16_t)dsp;
          }
        } else if (on && active) {
          *(uint4*)(sRow + (long long)x * stepX) = make_uint4(s[0], s[1], s[2], s[3]);
        }
        // ---- publish this pixel's vertical / diagonal paths for the row below
        if (VERT) {
          if (on) {
            uint4* o = own + (size_t)(x & ownMask) * SLOT;
            if (active) {
              if (PATHS & P_V) o[0 * LPP + gl] = make_uint4(LV[0], LV[1], LV[2], LV[3]);
              if (PATHS & P_DL) o[1 * LPP + gl] = make_uint4(LDL[0], LDL[1], LDL[2], LDL[3]);
              if (PATHS & P_DR) o[2 * LPP + gl] = make_uint4(LDR[0], LDR[1], LDR[2], LDR[3]);
            }
            if (gl == 0) o[3 * LPP] = make_uint4(mV, mDL, mDR, 0);
          }
          __syncwarp();
          if (lane == 0) {
            if (firstOn && t < W1) st_release_cta(&sCons[warp], t + 1);
            const int xl = t - 2 * (RW - 1);
            if (lastOn && xl >= 0 && xl < W1) st_release_cta(&sProg[warp], xl + 1);
          }
        }
      }
    } else if (VERT) {
      // ------------------------------------------------------------------ helper warp: band hand-over
      const bool doLoad = band > 0, doStore = band < a.nbands - 1;
      uint4* ringE = rings + (size_t)(kNW * G::WARP_SLOTS) * SLOT;
      const uint4* ringL = rings + (size_t)((kNW - 1) * G::WARP_SLOTS + (RW - 1) * kRingIn) * SLOT;
      const size_t bandSlots = (size_t)W1 * SLOT;
      const uint4* gin = a.edges + ((size_t)frame * a.nbands + (doLoad ? band - 1 : 0)) * bandSlots;
      uint4* gout = a.edges + ((size_t)frame * a.nbands + band) * bandSlots;
      const int* gflagIn = a.gprog + frame * a.nbands + (doLoad ? band - 1 : 0);
      int* gflagOut = a.gprog + frame * a.nbands + band;
      int loaded = 0, stored = 0, gseen = 0, published = 0, idleSpins = 0;
      while ((doLoad && loaded < W1) || (doStore && stored < W1)) {
        bool did = false;
        if (doLoad && loaded < W1) {
          if (gseen <= loaded) gseen = ld_acquire_gpu(gflagIn);
          const int c0 = ld_acquire_cta(&sCons[0]);
          const int hi = min(min(gseen, c0 + kRingEdge - 1), W1);
          if (hi > loaded) {
            const int n = (hi - loaded) * SLOT;
            for (int i = lane; i < n; i += 32) {
              const int p = loaded + i / SLOT, q = i % SLOT;
              ringE[(size_t)(p & (kRingEdge - 1)) * SLOT + q] = __ldcg(gin + (size_t)p * SLOT + q);
            }
            __syncwarp();
            if (lane == 0) st_release_cta(sProgE, hi);
            loaded = hi;
            did = true;
          }
        }
        if (doStore && stored < W1) {
          const int pl = ld_acquire_cta(&sProg[kNW - 1]);
          if (pl > stored) {
            const int n = (pl - stored) * SLOT;
            for (int i = lane; i < n; i += 32) {
              const int p = stored + i / SLOT, q = i % SLOT;
              gout[(size_t)p * SLOT + q] = ringL[(size_t)(p & (kRingOut - 1)) * SLOT + q];
            }
            __syncwarp();
            stored = pl;
            if (lane == 0) st_release_cta(&sCons[kNW], stored + 1);
            if (stored - published >= kPublish || stored == W1) {
              __threadfence();
              __syncwarp();
              if (lane == 0) st_release_gpu(gflagOut, stored);
              published = stored;
            }
            did = true;
          }
        }
        if (!did) {
          __nanosleep(32);
          if (((++idleSpins) & 255) == 0) {
            if (*sAbort || *(volatile int*)a.status) { *sAbort = 1; break; }
            if (idleSpins > (kSpinLimit >> 4)) { *sAbort = 1; atomicExch(a.status, 1); break; }
          }
        } else {
          idleSpins = 0;
        }
      }
    }
  }
}

template <int LPP, int PATHS, bool SECOND>
cudaError_t launch_one(const RasterArgs& a, int grid, cudaStream_t st) {
  using G = RasterGeom<LPP>;
  auto* k = k_agg_raster<LPP, PATHS, SECOND>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
  if (e != cudaSuccess) return e;
  k<<<grid, (kNW + 1) * 32, G::SMEM, st>>>(a);
  return cudaGetLastError();
}

template <int LPP>
cudaError_t launch_lpp(const RasterArgs& a, int paths, bool second, int grid, cudaStream_t st) {
  if (!second) {
    if (paths == 15) return launch_one<LPP, 15, false>(a, grid, st);
    return launch_one<LPP, P_H | P_V, false>(a, grid, st);
  }
  if (paths == 15) return launch_one<LPP, 15, true>(a, grid, st);
  if (paths == (P_H | P_V)) return launch_one<LPP, P_H | P_V, true>(a, grid, st);
  return launch_one<LPP, P_H, true>(a, grid, st);
}

int lpp_of(int D) {
  const int v = D / 8;
  return v <= 2 ? 2 : v <= 4 ? 4 : v <= 8 ? 8 : v <= 16 ? 16 : 32;
}
int slot_u4(int D) { return 3 * lpp_of(D) + 1; }

}  // namespace

int raster_rows_per_band(int D) { return kNW * (32 / lpp_of(D)); }
int raster_bands(const SgbmCfg& cfg) { const int rb = raster_rows_per_band(cfg.D); return (cfg.H + rb - 1) / rb; }
size_t sgbm_flags_bytes(const SgbmCfg& cfg, int nframes) { return ((size_t)nframes * raster_bands(cfg) + 64) * sizeof(int); }
size_t sgbm_edges_bytes(const SgbmCfg& cfg, int nframes) {
  return (size_t)nframes * raster_bands(cfg) * (size_t)std::max(cfg.W1, 1) * slot_u4(cfg.D) * sizeof(uint4);
}

// Aggregation + WTA of `nframes` pairs: S scratch, wtaKey and dispRaw are produced; C is read twice.
cudaError_t sgbm_aggregate_raster(const SgbmCfg& cfg, const SgbmScratch& sc, int nframes, cudaStream_t st, LaunchStats* stats) {
  static int numSM = 0;
  if (!numSM) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, dev);
  }
  RasterArgs a{};
  a.C = sc.C; a.S = sc.S; a.wtaKey = sc.wtaKey; a.dispRaw = sc.dispRaw;
  a.edges = (uint4*)sc.edges;
  a.ticket = (int*)sc.flags;
  a.gprog = (int*)sc.flags + 64;
  a.status = sc.status;
  a.cfg = cfg; a.nframes = nframes; a.nbands = raster_bands(cfg); a.frameCells = sc.frameCells;
  const int ntasks = a.nframes * a.nbands;
  const int grid = std::min(ntasks, numSM);
  const size_t flagBytes = ((size_t)nframes * a.nbands + 64) * sizeof(int);
  const int pass1 = cfg.mode == SV_MODE_HH4 ? (P_H | P_V) : 15;
  const int pass2 = cfg.mode == SV_MODE_HH ? 15 : (cfg.mode == SV_MODE_HH4 ? (P_H | P_V) : P_H);
  const int lpp = lpp_of(cfg.D);
  for (int pass = 0; pass < 2; pass++) {
    cudaError_t e = cudaMemsetAsync(sc.flags, 0, flagBytes, st);
    if (e != cudaSuccess) return e;
    a.flip = pass;
    const int paths = pass ? pass2 : pass1;
    switch (lpp) {
      case 2: e = launch_lpp<2>(a, paths, pass == 1, grid, st); break;
      case 4: e = launch_lpp<4>(a, paths, pass == 1, grid, st); break;
      case 8: e = launch_lpp<8>(a, paths, pass == 1, grid, st); break;
      case 16: e = launch_lpp<16>(a, paths, pass == 1, grid, st); break;
      default: e = launch_lpp<32>(a, paths, pass == 1, grid, st); break;
    }
    if (stats) stats->launches++;
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace svgpu

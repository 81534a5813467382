// sgbm_sweep.cu — SGM path aggregation as two fused sweeps over skewed column strips.
//
// Replaces the aggregation + winner-take-all inside cv::StereoSGBM::compute
// (/root/reference/src/lib.cpp:262; arithmetic: SURVEY.md A.5 / A.6).
//
// A raster-order sweep can carry the four paths whose predecessor lies to the left or in the row above
// (travel directions ->, down-right, down, down-left); the reverse sweep carries the other four.  MODE_HH
// therefore needs two passes over the cost volume: pass 1 reads C and writes S = sum of 4 paths, pass 2
// reads C and S, adds its 4 paths and selects the winner in registers: 8 B per volume cell in total.
//
// Work decomposition (per pass, in sweep coordinates; pass 2 is pass 1 on the volume flipped in x and y):
//   * the image is cut into strips of K columns (4; 2 when a call has so few pairs that its wavefronts set its
//     duration) whose borders move one pixel to the left per row:
//     strip s covers columns [s*K - y, s*K - y + K) of row y.  With that skew every predecessor of a
//     pixel — (x-1,y), (x-1,y-1), (x,y-1), (x+1,y-1) — lies in the same strip or in the strip to the LEFT,
//     so strips form a one-directional pipeline;
//   * a "worker" is a group of LPP lanes (16 disparities per lane) that owns one strip and walks it top to
//     bottom, K pixels per row ("chunk").  The horizontal path lives in registers; the three paths that
//     come from the row above live in a lane-private shared-memory row buffer (K records);
//   * per chunk a worker receives from its left neighbour the horizontal path of the pixel left of the
//     chunk and the vertical/diagonal paths of the two pixels above-left of it, and hands the same to its
//     right neighbour: one small "entry" per row instead of one hand-over per pixel;
//   * a warp holds 32/LPP neighbouring strips in lock step, one row apart (entries through shared memory,
//     __syncwarp only); warps of a CTA hand over through shared-memory rings with volatile counters; CTAs hand over
//     through a global edge buffer whose row count is published after a fence every few rows, the entries being
//     copied straight into the consumer's ring by the TMA one row ahead of their use;
//   * CTAs are persistent and draw (frame, strip group) tasks from a ticket counter, strip groups in
//     ascending order.  A task only waits for the task to its left, which has a smaller ticket and is
//     therefore resident or finished: no deadlock whatever the residency.
//   * few pairs per call: the two passes run side by side on two streams as two "first" passes (the second on the flipped
//     volume into its own S) and k_wta_sum adds and selects — see sgbm_aggregate_sweep.
//
// Volume layout: [frame][y][x'][D] uint16, the D costs of a pixel in groups of 16 with the two halves of a
// group interleaved (word i of a group = disparities d0+i and d0+i+8), so that the d-1 / d+1 neighbours of
// a packed register are simply the previous / next register of the lane (sgbm_device.cuh).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "sgbm_device.cuh"

namespace svgpu {

namespace {

// compile-time tuning knobs (tools/build_variants.sh builds alternatives for A/B runs on the GPU)
#ifndef SV_SWEEP_K
#define SV_SWEEP_K 4
#endif
#ifndef SV_SWEEP_MAXNW
#define SV_SWEEP_MAXNW 8
#endif
#ifndef SV_SWEEP_SLEEP
#define SV_SWEEP_SLEEP 40
#endif
#ifndef SV_SWEEP_PROF
#define SV_SWEEP_PROF 0
#endif
#ifndef SV_SWEEP_REVWARP
#define SV_SWEEP_REVWARP 1
#endif
#ifndef SV_SWEEP_PUBLISH
#define SV_SWEEP_PUBLISH 4   // rows between two publications (fence + counter) of the CTA-crossing entries
#endif
#ifndef SV_SWEEP_UNCOND
#define SV_SWEEP_UNCOND 1
#endif
#ifndef SV_SWEEP_MIN2
#define SV_SWEEP_MIN2 1      // 1: the minima of two paths share one shuffle butterfly, 0: one butterfly per path
#endif
#ifndef SV_SWEEP_INTERIOR
#define SV_SWEEP_INTERIOR 0  // 1: separate copy of the chunk code for warps whose pixels all lie strictly inside the volume
#endif
#ifndef SV_SWEEP_PIXFENCE
#define SV_SWEEP_PIXFENCE 0  // 1: compiler-level memory fence between the pixels of a chunk (keeps their loads / stores from being batched)
#endif
#ifndef SV_SWEEP_LDASM
#define SV_SWEEP_LDASM 0     // 1: operand prefetch through volatile inline PTX (keeps the loads where the source puts them)
#endif
#ifndef SV_SWEEP_LDPOS
#define SV_SWEEP_LDPOS 0     // where in a pixel's code the operands of the next row are fetched: 0 top, 1 when this pixel's are dead, 2 end
#endif
#ifndef SV_SWEEP_FAKE
#define SV_SWEEP_FAKE 0      // timing experiments only (results are wrong): 1 = no import from the left CTA, 2 = no export
#endif
constexpr int kK = SV_SWEEP_K;  // strip width = pixels per chunk
#ifndef SV_SWEEP_RINGX
#define SV_SWEEP_RINGX 8
#endif
constexpr int kRingX = SV_SWEEP_RINGX;  // entries in the import ring of a warp's first worker (fed by another warp / CTA)
#ifndef SV_SWEEP_AHEAD
#define SV_SWEEP_AHEAD (SV_SWEEP_RINGX - 1)
#endif
constexpr int kPub = SV_SWEEP_PUBLISH;
constexpr int kAhead = SV_SWEEP_AHEAD;  // rows a warp may run ahead of the warp that consumes its entries (< kRingX)
static_assert(kRingX % kPub == 0 && kRingX >= kPub + 2, "import blocks must tile the ring and leave room for the two entries in use");
constexpr int kRingI = 2;      // entries in the import ring of the other workers (fed in lock step by the same warp)
constexpr unsigned long long kWaitLimitNs = 10000000000ull;  // a hand-over that takes longer than this aborts the kernel
constexpr int kFlagBytes = 256;   // hand-over counters, ticket and abort flag at the end of shared memory
constexpr int kSmemMax = 227 * 1024;
constexpr int kMaxNW = SV_SWEEP_MAXNW;      // 8 warps = 256 threads leave the sweep kernels the full 255 registers per thread
constexpr int kLowK = 2;                    // latency variant for a few pairs per call: strips of 2 columns (a shorter wavefront, more strip groups)

enum : int { P_H = 1, P_DL = 2, P_V = 4, P_DR = 8 };

template <int LPP, bool SECOND, int K = kK>
struct SweepGeom {
  static constexpr int RW = 32 / LPP;             // workers (strips) per warp
  static constexpr int NV = kNR / 4;              // uint4 per lane per path vector
  static constexpr int VEC = LPP * NV;            // uint4 per path vector of one pixel
  static constexpr int REC = 3 * VEC;             // record of one pixel: V, DL, DR
  static constexpr int ENT = 4 * VEC + 1;         // hand-over entry of one row: H, V1, DL1, DL0, mins
  static constexpr int ROWBUF = K * REC;         // per worker
  static constexpr int WTA = SECOND ? VEC : 0;    // per worker: final S of one pixel (neighbour lookup)
  static constexpr int OFF_RING = RW * ROWBUF;    // in-warp rings of workers 1..RW-1, then the ring of worker 0
  static constexpr int OFF_RING0 = OFF_RING + (RW - 1) * kRingI * ENT;
  // Import ring of a warp's first worker and the publication period across CTAs.  At 16 lanes per pixel (D > 128) an entry is
  // 2 KB and the 8-deep ring would leave room for 6 warps only: 4 entries, published every 2 rows, keep all 8 warps.
  static constexpr int RINGX = LPP == 16 ? 4 : kRingX;
  static constexpr int PUB = LPP == 16 ? 2 : kPub;
  static constexpr int AHEAD = LPP == 16 ? RINGX - 1 : kAhead;
  static_assert(RINGX % PUB == 0 && RINGX >= PUB + 2 && AHEAD < RINGX, "import blocks must tile the ring and leave room for the two entries in use");
  static constexpr int OFF_XOUT = OFF_RING0 + RINGX * ENT;  // export of the last worker when it does not go to a ring
  static constexpr int OFF_WTA = OFF_XOUT + ENT;
  static constexpr int WARP_U4 = OFF_WTA + RW * WTA;
  static constexpr int STAGE_U4 = 1;               // the mbarrier of the TMA copies from the left CTA's entries
  static constexpr int NW_FIT = (kSmemMax - kFlagBytes - STAGE_U4 * 16) / (WARP_U4 * 16);
  static constexpr int NW = NW_FIT > kMaxNW ? kMaxNW : NW_FIT;  // warps per CTA
  static constexpr int SPC = NW * RW;             // strips per CTA task
  static constexpr size_t SMEM = (size_t)(NW * WARP_U4 + STAGE_U4) * 16 + kFlagBytes;
};

struct SweepArgs {
  const uint4* C;
  uint4* S;
  uint2* sel;     // [frame][H][W1]  pass 2: {reject<<31 | minS<<16 | best, S[best-1]<<16 | S[best+1]}
  uint4* edges;   // [frame][G-1][H][ENT]  hand-over entries between CTAs
  int* gprog;     // [frame][G]  rows of a strip group's last strip that are visible in `edges`
  int* ticket;
  int* status;
  int W1, H, D, nAct;          // nAct = D/16 lanes of a worker hold disparities
  int nframes, G, smax, flip, debugS;
  unsigned frameU4;            // uint4 per frame of a volume
  unsigned rowU4;              // uint4 per volume row = W1*D/8
  int pixU4;                   // uint4 per pixel = D/8
  uint32_t p1dup, p2mp1dup, negp1dup;
  int smallSums;               // 4 x (largest block cost + P2) <= 32767: a pass may add its path costs without saturation
  int uniqF;                   // 100 - uniquenessRatio
  uint32_t uniqMagic;          // ceil(2^32 / uniqF) for uniqF >= 2
  uint32_t one;                // 1, as a value the compiler cannot fold (multiply-adds that must stay on the FMA pipe)
};

// Hand-over counters between the warps of a CTA live in shared memory and are accessed with plain volatile
// loads / stores: the shared-memory pipeline of an SM performs the accesses of a warp in program order, so the
// ring entry written before the counter (same warp, after __syncwarp) is in place when another warp reads the
// new counter value.  (A fence.cta here would also wait for the outstanding global prefetch loads of the
// warp, i.e. expose a DRAM round trip in every step.)
__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// orders the generic-proxy acquire of a publication before the async-proxy (TMA) read of the published entries
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// bulk (TMA) copy global -> shared whose completion is signalled on an mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smemDst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smemDst)), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, int parity) {  // never suspends
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, int parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
#if SV_SWEEP_LDASM
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#else
  return __ldcs(p);
#endif
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until *flag >= need, backing off between polls so that the warps sharing the scheduler keep their issue
// slots; false when the kernel is being aborted
__device__ __forceinline__ bool wait_cta(const int* flag, int need, volatile int* sAbort, int* status) {
  if (ld_flag(flag) >= need) return true;
  unsigned long long t0 = 0;
  int n = 0;
  while (ld_flag(flag) < need) {
    if (SV_SWEEP_SLEEP) __nanosleep(SV_SWEEP_SLEEP);
    if (((++n) & 255) == 0) {
      if (*sAbort) return false;
      if (!t0) t0 = now_ns();
      if (now_ns() - t0 > kWaitLimitNs) { *sAbort = 1; atomicExch(status, 1); return false; }
    }
  }
  return true;
}

// FULL: every lane of a worker holds disparities (D is LPP x 16), so the "lane without disparities" cases fold away
// SMALL: four path costs of a pixel sum to less than 2^15 (4 x (largest block cost + P2) <= 32767, checked on the host), so a
//        pass can add them with plain (three-input) adds and only the final S of pass 2 needs the saturation of A.5
template <int LPP, int PATHS, bool SECOND, int K, bool WAY3 = false, bool FULL = false, bool SMALL = false>
__global__ void __launch_bounds__(SweepGeom<LPP, SECOND, K>::NW * 32, 1) k_sweep(const SweepArgs a) {
  using G = SweepGeom<LPP, SECOND, K>;
  constexpr int RW = G::RW, NV = G::NV, VEC = G::VEC, REC = G::REC, ENT = G::ENT, NW = G::NW;
  constexpr bool HAS_V = (PATHS & P_V) != 0, HAS_DL = (PATHS & P_DL) != 0, HAS_DR = (PATHS & P_DR) != 0;
  constexpr bool VERT = HAS_V || HAS_DL || HAS_DR;
  extern __shared__ uint4 smem[];
  uint64_t* sBar = (uint64_t*)(smem + NW * G::WARP_U4);  // completion of the TMA copy of a CTA-crossing entry
  int* flags = (int*)(smem + NW * G::WARP_U4 + G::STAGE_U4);
  int* sProg = flags;            // [NW] rows exported by the last worker of warp w (last row + 1)
  int* sCons = flags + NW;       // [NW] rows finished by the first worker of warp w (next row)
  int* sTicket = flags + 2 * NW;
  volatile int* sAbort = flags + 2 * NW + 1;

  // position of this warp in the CTA's chain of strips.  The scheduler arbitrates hardware warps highest id first, and
  // a chain runs at the pace of its head (the warp that imports from the left CTA and never finds slack), so the head
  // is the hardware warp with the highest id and followers fill the issue slots it leaves.
#if SV_SWEEP_REVWARP
  const int warp = NW - 1 - (threadIdx.x >> 5), lane = threadIdx.x & 31;
#else
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#endif
  const int j = lane / LPP, gl = lane % LPP;
  const int W1 = a.W1, H = a.H;
  const bool active = FULL ? true : gl < a.nAct;  // lane holds disparities
  const int lastLane = a.nAct - 1;
  SweepK kc;
  kc.p1dup = a.p1dup; kc.p2mp1dup = a.p2mp1dup; kc.negp1dup = a.negp1dup;
  kc.keepUp = opaque_u32(gl == 0 ? 0u : 1u); kc.addUp = opaque_u32(gl == 0 ? kMax2 : 0u);
  kc.keepDn = opaque_u32(gl == lastLane ? 0u : 1u); kc.addDn = opaque_u32(gl == lastLane ? kMax2 : 0u);
  kc.one = a.one;

  // shared memory starts out as valid (zero) path costs: stale entries are read under an override only
  for (int i = threadIdx.x; i < NW * G::WARP_U4; i += blockDim.x) smem[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { *sAbort = 0; mbar_init(sBar, 1); }
  int barPhase = 0;  // parity of the next completion of sBar (only warp 0 uses it)

  uint4* wbase = smem + warp * G::WARP_U4;
  uint4* rowbuf = wbase + j * G::ROWBUF + gl;  // lane-private: [pixel][path][half][gl]
  uint4* inRing = wbase + (j > 0 ? G::OFF_RING + (j - 1) * kRingI * ENT : G::OFF_RING0);
  const int inMask = j > 0 ? kRingI - 1 : G::RINGX - 1;
  uint4* outRing = (j < RW - 1) ? wbase + G::OFF_RING + j * kRingI * ENT : wbase + G::WARP_U4 + G::OFF_RING0;
  const int outMask = (j < RW - 1) ? kRingI - 1 : G::RINGX - 1;
  uint4* ring0 = wbase + G::OFF_RING0;
  uint16_t* wta = (uint16_t*)(wbase + G::OFF_WTA + j * G::WTA);

#if SV_SWEEP_PROF
  long long pfTask = 0, pfLeft = 0, pfRight = 0, pfGlobal = 0, pfIdle = 0, pfSteps = 0, pfA = 0, pfBlock = 0, pfFail = 0, pfB = 0, pfA1 = 0, pfA3 = 0, pfA5 = 0;
  long long pfT0 = clock64();
#define PF_BEGIN long long pf_t = clock64()
#define PF_END(acc) acc += clock64() - pf_t
#else
#define PF_BEGIN
#define PF_END(acc)
#endif
  const int ntasks = a.nframes * a.G;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) *sTicket = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int task = *sTicket;
    if (task >= ntasks || *sAbort) break;
    const int frame = task % a.nframes, sg = task / a.nframes;
    const int sFirst = sg * G::SPC + warp * RW;  // strip of this warp's worker 0
    auto y0of = [&](int s) { return max(0, s * K - W1 + 1); };
    auto y1of = [&](int s) { return min(H - 1, s * K + K - 1); };
    if (lane == 0) { sProg[warp] = 0; sCons[warp] = y0of(sFirst); }
    __syncthreads();
    if (sFirst > a.smax) continue;  // whole warp beyond the last strip (still takes part in the barriers above)
#if SV_SWEEP_PROF
    pfIdle += clock64() - pfT0;
    const long long pfStart = clock64();
#endif

    const int s = sFirst + j;
    const int y0 = y0of(s), y1 = y1of(s);            // rows of this strip (empty when s > smax)
    const int jLast = min(RW - 1, a.smax - sFirst);  // last worker with a strip
    const int sLast = sFirst + RW - 1;
    const bool hasLeft = sFirst > 0 && !((SV_SWEEP_FAKE & 1) && warp == 0), hasRight = sLast < a.smax && !((SV_SWEEP_FAKE & 2) && warp == NW - 1);
    const int y0F = y0of(sFirst), y1F = y1of(sFirst);
    const int y0L = y0of(sLast), y1L = y1of(sLast);
    const int y0p = hasLeft ? y0of(sFirst - 1) : 0;   // rows of the strip left of worker 0: its entry k is row y0p + k
    const int y1p = hasLeft ? y1of(sFirst - 1) : -1;
    const bool leftGlobal = hasLeft && warp == 0, rightGlobal = hasRight && warp == NW - 1;
    const int Tbeg = y0F - 1, Tend = y1of(sFirst + jLast) + jLast;

    const int glc = active ? gl : 0;  // lanes without disparities load what lane 0 loads and store nothing
    const uint4* Cf = a.C + (size_t)frame * a.frameU4 + glc * NV;
    uint4* Sf = a.S + (size_t)frame * a.frameU4 + glc * NV;
    const uint4* gin = nullptr;
    uint4* gout = nullptr;
    const int* gflagIn = nullptr;
    int* gflagOut = nullptr;
    if (leftGlobal) { gin = a.edges + ((size_t)frame * (a.G - 1) + (sg - 1)) * (size_t)H * ENT; gflagIn = a.gprog + frame * a.G + (sg - 1); }
    if (rightGlobal) { gout = a.edges + ((size_t)frame * (a.G - 1) + sg) * (size_t)H * ENT; gflagOut = a.gprog + frame * a.G + sg; }
    int gseen = 0, gpend = 0;   // rows the left CTA has published (cached / polled one step ahead)
    int loaded = max(y0F - 1, 0);  // rows [.., loaded) of the left CTA's entries are in ring0
    int pendEnd = 0;               // rows [loaded, pendEnd) are being copied into ring0 by the TMA

    // ---- per-task register state
    uint32_t Hs[kNR], Vd1[kNR], DLd1[kNR], DLd2[kNR];
    uint32_t mH = 0, mVd1 = 0, mDLd1 = 0, mDLd2 = 0;
    uint32_t recM[K][3];   // mins of the records in the row buffer
    uint4 cb[K][NV], sb[K][NV];  // C (and S of pass 1) of this chunk, loaded one chunk ahead
#pragma unroll
    for (int q = 0; q < kNR; q++) Hs[q] = Vd1[q] = DLd1[q] = DLd2[q] = 0;
#pragma unroll
    for (int i = 0; i < K; i++) {
      recM[i][0] = recM[i][1] = recM[i][2] = 0;
#pragma unroll
      for (int h = 0; h < NV; h++) cb[i][h] = sb[i][h] = make_uint4(0, 0, 0, 0);
    }

    bool ok = true;
    int where = 0, Tlast = 0;
    for (int T = Tbeg; T <= Tend && ok; T++) {
      Tlast = T;
      const int r = T - j;                               // row of this worker
      const bool act = r >= y0 && r <= y1;
      const int xl = s * K - r;                         // leftmost column of the chunk
      const bool act0 = T >= y0F && T <= y1F;            // worker 0 of the warp has a row
      const int rL = T - (RW - 1);
      const bool actL = rL >= y0L && rL <= y1L;          // worker RW-1 has a row (and a strip)
      // ================================================================== phase A: hand-over in
#if SV_SWEEP_PROF
      const long long pfA0 = clock64();
#endif
      if (hasLeft && act0) {
        const int need = min(T, y1p) + 1;
        where = leftGlobal ? 3 : 1;
        if (!leftGlobal) {
          { PF_BEGIN; ok = wait_cta(&sProg[warp - 1], need, sAbort, a.status); PF_END(pfLeft); }
        } else {
          // The left CTA publishes a row count after a fence every G::PUB rows; published entries are copied straight into
          // this warp's ring by the TMA in blocks of up to G::PUB rows (row-aligned like the publications), ahead of use.
          auto blockEnd = [&](int from) { return min((from / G::PUB) * G::PUB + G::PUB, y1p + 1); };
          auto issue = [&](int from, int to) {
            __syncwarp();
            if (lane == 0) {
              fence_proxy_async_global();
              bulk_load(ring0 + (from & (G::RINGX - 1)) * ENT, gin + (size_t)from * ENT, (uint32_t)(to - from) * (ENT * 16), sBar);
            }
          };
          auto finish = [&]() {  // wait for the block in flight
            for (int n = 0; !mbar_try_wait(sBar, barPhase);)
              if (((++n) & 255) == 0 && *sAbort) { ok = false; break; }
            barPhase ^= 1;
            loaded = pendEnd;
          };
          if (loaded < need) {  // rows needed now
            PF_BEGIN;
            if (pendEnd > loaded) finish();
            int n = 0;
            unsigned long long t0 = 0;
            while (loaded < need && ok) {
              const int be = blockEnd(loaded);
              if (gseen < be) {
                gseen = ld_acquire_gpu(gflagIn);
                if (gseen < be) {
                  __nanosleep(200);
                  if (((++n) & 63) == 0) {
                    if (!t0) t0 = now_ns();
                    if (*sAbort) { ok = false; break; }
                    if (now_ns() - t0 > kWaitLimitNs) { *sAbort = 1; atomicExch(a.status, 1); ok = false; break; }
                  }
                  continue;
                }
              }
              issue(loaded, be);
              pendEnd = be;
              finish();
            }
            PF_END(pfGlobal);
          } else if (pendEnd > loaded && mbar_test_wait(sBar, barPhase)) {
            barPhase ^= 1;
            loaded = pendEnd;
          }
          // request the next block (entries T-1 and T are in use, so rows up to T + G::RINGX - 2 have a free slot)
          if (ok && pendEnd <= loaded && loaded <= y1p) {
            const int be = blockEnd(loaded);
            if (be - 1 <= T + G::RINGX - 2) {
              if (gseen < be) {
                // the count is polled with a relaxed load one step ahead (no stall); only when that hint says the block
                // is there is it confirmed with an acquire load, which orders the copy below behind the publication
                if (gpend >= be) gseen = ld_acquire_gpu(gflagIn);
                else gpend = ld_relaxed_gpu(gflagIn);
              }
              if (gseen >= be) { issue(loaded, be); pendEnd = be; }
            }
          }
        }
      }
      if (hasRight && actL && !rightGlobal && ok) { where = 2; PF_BEGIN; ok = wait_cta(&sCons[warp + 1], rL - G::AHEAD, sAbort, a.status); PF_END(pfRight); }
      if (!ok) break;
#if SV_SWEEP_PROF
      const long long pfq5 = clock64();
#endif
      where = 0;
      __syncwarp();
      {
        const uint4* eH = inRing + (r & inMask) * ENT + gl;
        const uint4* eR = inRing + ((r - 1) & inMask) * ENT + gl;
#pragma unroll
        for (int h = 0; h < NV; h++) {
          const uint4 v = eH[0 * VEC + h * LPP];
          Hs[4 * h] = v.x; Hs[4 * h + 1] = v.y; Hs[4 * h + 2] = v.z; Hs[4 * h + 3] = v.w;
          if (HAS_V) { const uint4 u = eR[1 * VEC + h * LPP]; Vd1[4 * h] = u.x; Vd1[4 * h + 1] = u.y; Vd1[4 * h + 2] = u.z; Vd1[4 * h + 3] = u.w; }
          if (HAS_DL) {
            const uint4 u = eR[2 * VEC + h * LPP]; DLd1[4 * h] = u.x; DLd1[4 * h + 1] = u.y; DLd1[4 * h + 2] = u.z; DLd1[4 * h + 3] = u.w;
            const uint4 w = eR[3 * VEC + h * LPP]; DLd2[4 * h] = w.x; DLd2[4 * h + 1] = w.y; DLd2[4 * h + 2] = w.z; DLd2[4 * h + 3] = w.w;
          }
        }
        mH = (eH - gl)[4 * VEC].x;
        if (VERT) { const uint4 m = (eR - gl)[4 * VEC]; mVd1 = m.y; mDLd1 = m.z; mDLd2 = m.w; }
      }
      __syncwarp();
      if (act0 && lane == 0) st_flag(&sCons[warp], T);  // rows < T are done with: their entries may be overwritten
#if SV_SWEEP_PROF
      pfA5 += clock64() - pfq5;
      pfA += clock64() - pfA0;
      const long long pfB0 = clock64();
#endif
      // ================================================================== phase B: K pixels
      // INTERIOR: every worker of the warp is on a row that is neither the first nor the last of its strip and all its pixels
      // (and the ones fetched for the next row) lie strictly inside the volume: no predecessor is outside, nothing is
      // predicated.  The test is warp-uniform arithmetic on T, so the warp takes one of two copies of the chunk code.
      auto phaseB = [&](auto interiorTag) {
      constexpr bool IN = decltype(interiorTag)::value;
      // which parts of the chunk code the interior copy simplifies (SV_SWEEP_INTERIOR is a bit mask: 1 = enabled; the bits
      // 2 / 4 / 8 / 16 keep the general form of the operand loads / the 'outside' flags / the stores / the row flags)
      constexpr bool IN_LD = IN && !(SV_SWEEP_INTERIOR & 2), IN_OUT = IN && !(SV_SWEEP_INTERIOR & 4), IN_ST = IN && !(SV_SWEEP_INTERIOR & 8),
                     IN_ROW = IN && !(SV_SWEEP_INTERIOR & 16);
      const int yp = a.flip ? H - 1 - r : r;
      const int ypn = a.flip ? yp - 1 : yp + 1;                     // next row of this worker
      // volume offsets of the chunk's first pixel in this row and of the first pixel fetched for the next row; the chunk
      // walks them in steps of +-pixU4 (pass 2 runs on the flipped volume)
      const int stepU4 = a.flip ? -a.pixU4 : a.pixU4;
      const unsigned offRow0 = (unsigned)yp * a.rowU4 + (unsigned)(a.flip ? W1 - 1 - xl : xl) * (unsigned)a.pixU4;
      const unsigned offNext0 = (unsigned)ypn * a.rowU4 + (unsigned)(a.flip ? W1 - xl : xl - 1) * (unsigned)a.pixU4;
      uint2* selRow = SECOND ? a.sel + ((size_t)frame * H + ((IN_ROW || act) ? yp : 0)) * W1 : nullptr;  // selection output row of this step
      const bool actn = IN_LD || ((r + 1) >= y0 && (r + 1) <= y1);
      const bool top = !IN_OUT && r == 0;
      // export target of this chunk: the ring of the right neighbour, or (no neighbour / inactive row / neighbour in
      // another CTA) a private entry that the last warp copies to global memory after the chunk
      const bool toRing = (j < RW - 1) || (hasRight && (IN_ROW || act) && !rightGlobal);
      const bool expGlobal = (j == RW - 1) && rightGlobal && (IN_ROW || act);
      uint4* eOut = (toRing ? outRing + (r & outMask) * ENT : wbase + G::OFF_XOUT) + gl;
      uint32_t mDLprev = 0;
      uint4 eMins = make_uint4(0, 0, 0, 0);
      // SV_SWEEP_LDPOS = 3: the operands of the whole next row step are requested before the first pixel, into spare
      // registers that are copied into cb / sb after the last pixel (a full row step of latency for every pixel's fetch;
      // 4 = the same for C only, S as in 0)
      constexpr bool EARLY_C = SV_SWEEP_LDPOS == 3 || SV_SWEEP_LDPOS == 4, EARLY_S = SECOND && SV_SWEEP_LDPOS == 3;
      uint4 nbC[EARLY_C ? K : 1][NV], nbS[EARLY_S ? K : 1][NV];
      if constexpr (EARLY_C) {
#pragma unroll
        for (int i = 0; i < K; i++) {
          const int xn = xl + i - 1;
          const bool ld = IN_LD || (actn && (unsigned)xn < (unsigned)W1);
          const unsigned off = ld ? offNext0 + (unsigned)(i * stepU4) : 0u;
#pragma unroll
          for (int h = 0; h < NV; h++) {
            nbC[i][h] = ld_stream(Cf + off + h);
            if constexpr (EARLY_S) nbS[i][h] = ld_stream(Sf + off + h);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < K; i++) {
        if (SV_SWEEP_PIXFENCE && i > 0) asm volatile("" ::: "memory");
        const int x = xl + i;
        const bool on = IN_ST || (act && (unsigned)x < (unsigned)W1);   // stores, selection output
        const bool onP = IN_OUT || (act && (unsigned)x < (unsigned)W1);  // the pixel takes part in the paths
        const bool xFirst = !IN_OUT && x == 0, xLast = !IN_OUT && x == W1 - 1;
        const int xp = a.flip ? W1 - 1 - x : x;
        // ---- operands of this pixel were loaded one chunk ago; fetch those of the pixel below-left now
        uint32_t c[kNR], sin[kNR];
#pragma unroll
        for (int h = 0; h < NV; h++) {
          c[4 * h] = cb[i][h].x; c[4 * h + 1] = cb[i][h].y; c[4 * h + 2] = cb[i][h].z; c[4 * h + 3] = cb[i][h].w;
          if (SECOND) { sin[4 * h] = sb[i][h].x; sin[4 * h + 1] = sb[i][h].y; sin[4 * h + 2] = sb[i][h].z; sin[4 * h + 3] = sb[i][h].w; }
        }
        // The fetch is issued once the operands of this pixel are dead (SV_SWEEP_LDPOS = 1: C after the path steps, S after the
        // final sum; 0: at the top of the pixel), so that it can land in the registers the next row step reads them from —
        // issued earlier it lands in spare registers and the copy at the end of the row step waits for it.
        auto prefetch = [&](bool doC, bool doS) {
          const int xn = x - 1;
          const bool ld = IN_LD || (actn && (unsigned)xn < (unsigned)W1);
#if SV_SWEEP_UNCOND
          // always load (from a safe address when there is nothing to fetch) so that the data lands in cb/sb directly
          const unsigned off = ld ? offNext0 + (unsigned)(i * stepU4) : 0u;
#pragma unroll
          for (int h = 0; h < NV; h++) {
            if (doC) cb[i][h] = ld_stream(Cf + off + h);
            if (SECOND && doS) sb[i][h] = ld_stream(Sf + off + h);
          }
#else
          if (ld) {
            const unsigned off = offNext0 + (unsigned)(i * stepU4);
#pragma unroll
            for (int h = 0; h < NV; h++) {
              if (doC) cb[i][h] = ld_stream(Cf + off + h);
              if (SECOND && doS) sb[i][h] = ld_stream(Sf + off + h);
            }
          }
#endif
        };
        if (SV_SWEEP_LDPOS == 0) prefetch(true, true);
        if (SV_SWEEP_LDPOS == 4) prefetch(false, true);
        // ---- old record of this slot: DR is used now, V one pixel later, DL two pixels later
        uint32_t oV[kNR], oDL[kNR], DRin[kNR];
        uint4* rec = rowbuf + i * REC;
#pragma unroll
        for (int h = 0; h < NV; h++) {
          if (HAS_V) { const uint4 v = rec[0 * VEC + h * LPP]; oV[4 * h] = v.x; oV[4 * h + 1] = v.y; oV[4 * h + 2] = v.z; oV[4 * h + 3] = v.w; }
          if (HAS_DL) { const uint4 v = rec[1 * VEC + h * LPP]; oDL[4 * h] = v.x; oDL[4 * h + 1] = v.y; oDL[4 * h + 2] = v.z; oDL[4 * h + 3] = v.w; }
          if (HAS_DR) { const uint4 v = rec[2 * VEC + h * LPP]; DRin[4 * h] = v.x; DRin[4 * h + 1] = v.y; DRin[4 * h + 2] = v.z; DRin[4 * h + 3] = v.w; }
        }
        const uint32_t omV = recM[i][0], omDL = recM[i][1];
        uint32_t mDR = recM[i][2];
        // ---- the min-plus steps (A.5); a predecessor outside the volume is expressed through the scalars
#if SV_SWEEP_MIN2
        // the minima of two paths share one shuffle butterfly: (H, DR) and (V, DL) — or (H, V) without diagonals.  The
        // partner of the horizontal path is computed first, so that nothing waits for H longer than before.
        uint32_t sum[kNR];
        uint32_t mV = mVd1, mDL = mDLd2;
        if (HAS_DR) {
          constexpr bool HV = SV_SWEEP_MINPAIR != 0;
          const uint32_t lDR = sweep_path_core<LPP, (SV_SWEEP_FMAADD > 0), HV>(DRin, mDR, !onP || top || xLast, c, kc, gl, lastLane, active, lane);
          const uint32_t lH = sweep_path_core<LPP, (SV_SWEEP_FMAADD > 3), HV>(Hs, mH, !onP || xFirst, c, kc, gl, lastLane, active, lane);
          group_min2_u16<LPP, HV>(lH, lDR, lane, mH, mDR);
#pragma unroll
          for (int q = 0; q < kNR; q++) sum[q] = SMALL ? Hs[q] + DRin[q] : __viaddmin_u16x2(Hs[q], DRin[q], kMax2);
          const uint32_t lV = sweep_path_core<LPP, (SV_SWEEP_FMAADD > 1), HV>(Vd1, mV, !onP || top, c, kc, gl, lastLane, active, lane);
          const uint32_t lDL = sweep_path_core<LPP, (SV_SWEEP_FMAADD > 2), HV>(DLd2, mDL, !onP || top || xFirst, c, kc, gl, lastLane, active, lane);
          group_min2_u16<LPP, HV>(lV, lDL, lane, mV, mDL);
#pragma unroll
          for (int q = 0; q < kNR; q++)
            sum[q] = SMALL ? sum[q] + Vd1[q] + DLd2[q] : __viaddmin_u16x2(__viaddmin_u16x2(sum[q], Vd1[q], kMax2), DLd2[q], kMax2);
        } else if (HAS_V) {
          constexpr bool HV = SV_SWEEP_MINPAIR != 0;
          const uint32_t lV = sweep_path_core<LPP, (SV_SWEEP_FMAADD > 0), HV>(Vd1, mV, !onP || top, c, kc, gl, lastLane, active, lane);
          const uint32_t lH = sweep_path_core<LPP, (SV_SWEEP_FMAADD > 3), HV>(Hs, mH, !onP || xFirst, c, kc, gl, lastLane, active, lane);
          group_min2_u16<LPP, HV>(lH, lV, lane, mH, mV);
#pragma unroll
          for (int q = 0; q < kNR; q++) sum[q] = SMALL ? Hs[q] + Vd1[q] : __viaddmin_u16x2(Hs[q], Vd1[q], kMax2);
        } else {
          sweep_path_step<LPP>(Hs, mH, !onP || xFirst, c, kc, gl, lastLane, active, lane);
#pragma unroll
          for (int q = 0; q < kNR; q++) sum[q] = Hs[q];
        }
#else
        sweep_path_step<LPP>(Hs, mH, !onP || xFirst, c, kc, gl, lastLane, active, lane);
        uint32_t sum[kNR];
#pragma unroll
        for (int q = 0; q < kNR; q++) sum[q] = Hs[q];
        uint32_t mV = mVd1, mDL = mDLd2;
        if (HAS_V) {
          sweep_path_step<LPP>(Vd1, mV, !onP || top, c, kc, gl, lastLane, active, lane);
#pragma unroll
          for (int q = 0; q < kNR; q++) sum[q] = __viaddmin_u16x2(sum[q], Vd1[q], kMax2);
        }
        if (HAS_DL) {
          sweep_path_step<LPP>(DLd2, mDL, !onP || top || xFirst, c, kc, gl, lastLane, active, lane);
#pragma unroll
          for (int q = 0; q < kNR; q++) sum[q] = __viaddmin_u16x2(sum[q], DLd2[q], kMax2);
        }
        if (HAS_DR) {
          sweep_path_step<LPP>(DRin, mDR, !onP || top || xLast, c, kc, gl, lastLane, active, lane);
#pragma unroll
          for (int q = 0; q < kNR; q++) sum[q] = __viaddmin_u16x2(sum[q], DRin[q], kMax2);
        }
#endif
        if (SV_SWEEP_LDPOS == 1) prefetch(true, false);
        // ---- new record of this slot (read by this worker in the next row)
#pragma unroll
        for (int h = 0; h < NV; h++) {
          if (HAS_V) rec[0 * VEC + h * LPP] = make_uint4(Vd1[4 * h], Vd1[4 * h + 1], Vd1[4 * h + 2], Vd1[4 * h + 3]);
          if (HAS_DL) rec[1 * VEC + h * LPP] = make_uint4(DLd2[4 * h], DLd2[4 * h + 1], DLd2[4 * h + 2], DLd2[4 * h + 3]);
          if (HAS_DR) rec[2 * VEC + h * LPP] = make_uint4(DRin[4 * h], DRin[4 * h + 1], DRin[4 * h + 2], DRin[4 * h + 3]);
        }
        recM[i][0] = mV; recM[i][1] = mDL; recM[i][2] = mDR;
        // ---- hand-over to the right neighbour: the last two pixels of the chunk
        if (i >= K - 2) {
#pragma unroll
          for (int h = 0; h < NV; h++) {
            if (i == K - 2) {
              if (HAS_DL) eOut[3 * VEC + h * LPP] = make_uint4(DLd2[4 * h], DLd2[4 * h + 1], DLd2[4 * h + 2], DLd2[4 * h + 3]);
            } else {
              eOut[0 * VEC + h * LPP] = make_uint4(Hs[4 * h], Hs[4 * h + 1], Hs[4 * h + 2], Hs[4 * h + 3]);
              if (HAS_V) eOut[1 * VEC + h * LPP] = make_uint4(Vd1[4 * h], Vd1[4 * h + 1], Vd1[4 * h + 2], Vd1[4 * h + 3]);
              if (HAS_DL) eOut[2 * VEC + h * LPP] = make_uint4(DLd2[4 * h], DLd2[4 * h + 1], DLd2[4 * h + 2], DLd2[4 * h + 3]);
            }
          }
          if (i == K - 1) eMins = make_uint4(mH, mV, mDL, mDLprev);
          mDLprev = mDL;
        }
        // ---- output of the pixel
        if (SECOND) {
#pragma unroll
          for (int q = 0; q < kNR; q++) sum[q] = __viaddmin_u16x2(sum[q], sin[q], kMax2);
          uint2 o;
          sweep_wta<LPP, WAY3>(sum, wta, gl, active, lane, a.D, a.uniqF, a.uniqMagic, o);
          const unsigned off = offRow0 + (unsigned)(i * stepU4);
          if (on && gl == 0) selRow[xp] = o;
          if (a.debugS && on && active) {
#pragma unroll
            for (int h = 0; h < NV; h++) Sf[off + h] = make_uint4(sum[4 * h], sum[4 * h + 1], sum[4 * h + 2], sum[4 * h + 3]);
          }
        } else if (on && active) {
          const unsigned off = offRow0 + (unsigned)(i * stepU4);
#pragma unroll
          for (int h = 0; h < NV; h++) __stcs(Sf + off + h, make_uint4(sum[4 * h], sum[4 * h + 1], sum[4 * h + 2], sum[4 * h + 3]));
        }
        if (SV_SWEEP_LDPOS == 1) prefetch(false, true);
        if (SV_SWEEP_LDPOS == 2) prefetch(true, true);
        // ---- shift the delay line: what this slot held serves the next two pixels
#pragma unroll
        for (int q = 0; q < kNR; q++) {
          if (HAS_DL) { DLd2[q] = DLd1[q]; DLd1[q] = oDL[q]; }
          if (HAS_V) Vd1[q] = oV[q];
        }
        mDLd2 = mDLd1; mDLd1 = omDL; mVd1 = omV;
      }
      if constexpr (EARLY_C) {
#pragma unroll
        for (int i = 0; i < K; i++)
#pragma unroll
          for (int h = 0; h < NV; h++) {
            cb[i][h] = nbC[i][h];
            if constexpr (EARLY_S) sb[i][h] = nbS[i][h];
          }
      }
#if SV_SWEEP_PROF
      pfB += clock64() - pfB0;
#endif
      // ================================================================== hand-over out
      if (gl == 0) (eOut - gl)[4 * VEC] = eMins;
      if (expGlobal) {  // the right neighbour is in another CTA: copy this lane's part of the entry to global memory
        uint4* gOut = gout + (size_t)r * ENT + gl;
#pragma unroll
        for (int v = 0; v < 4; v++)
#pragma unroll
          for (int h = 0; h < NV; h++) __stcg(gOut + v * VEC + h * LPP, eOut[v * VEC + h * LPP]);
        if (gl == 0) __stcg(gOut - gl + 4 * VEC, eMins);
      }
      if (rightGlobal && actL && ((rL % G::PUB) == G::PUB - 1 || rL == y1L)) {
        // publish: the warp barrier orders every lane's entry stores before lane 0's release store of the counter (the
        // release also waits for the warp's outstanding prefetch loads, which is why it is not done for every row)
        __syncwarp();
        if (lane == 0) st_release_gpu(gflagOut, rL + 1);
      }
      __syncwarp();
      if (lane == 0 && actL && hasRight && !rightGlobal) st_flag(&sProg[warp], rL + 1);
          };
      {
        const int xw = sFirst * K - T;  // leftmost column of worker 0's chunk; worker j starts (K + 1) j columns further right
        const bool interior = T >= RW && T <= H - 2 && xw >= 1 && xw + (K + 1) * (RW - 1) + K - 1 <= W1 - 2;
        if constexpr ((SV_SWEEP_INTERIOR & 1) != 0) {
          if (interior) phaseB(std::true_type{});
          else phaseB(std::false_type{});
        } else {
          (void)interior;
          phaseB(std::false_type{});
        }
      }
}
#if SV_SWEEP_PROF
    pfTask += clock64() - pfStart; pfSteps += Tend - Tbeg + 1; pfT0 = clock64();
    if (task < (int)gridDim.x && lane == 0)  // first wave: every strip group starts at the same time
      printf("PW %d %d %d %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", (int)SECOND, blockIdx.x, warp, pfTask, pfLeft, pfRight, pfGlobal, pfIdle, pfSteps, pfA, pfB, pfBlock, pfFail, pfA1, pfA3, pfA5);
#endif
    if (!ok && lane == 0)
      printf("svgpu sweep: cta %d warp %d task (frame %d, group %d) stopped at row step %d of [%d,%d] while waiting for %s\n", blockIdx.x, warp, frame, sg,
             Tlast, Tbeg, Tend, where == 1 ? "left warp" : where == 2 ? "right warp" : where == 3 ? "left CTA" : "nothing");
  }
#if SV_SWEEP_PROF
  pfIdle += clock64() - pfT0;
  if (lane == 0) printf("PF %d %d %d %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", (int)SECOND, blockIdx.x, warp, pfTask, pfLeft, pfRight, pfGlobal, pfIdle, pfSteps, pfA, pfB, pfBlock, pfFail, pfA1, pfA3, pfA5);
#endif
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int lpp_of(int D) {
  const int v = D / kDPL;
  return v <= 1 ? 1 : v <= 2 ? 2 : v <= 4 ? 4 : v <= 8 ? 8 : 16;
}

template <int LPP>
struct Dims {
  static int spc(bool second, int k) {
    if (k == kLowK) return second ? SweepGeom<LPP, true, kLowK>::SPC : SweepGeom<LPP, false, kLowK>::SPC;
    return second ? SweepGeom<LPP, true>::SPC : SweepGeom<LPP, false>::SPC;
  }
};
int spc_of(int lpp, bool second, int k) {
  switch (lpp) {
    case 1: return Dims<1>::spc(second, k);
    case 2: return Dims<2>::spc(second, k);
    case 4: return Dims<4>::spc(second, k);
    case 8: return Dims<8>::spc(second, k);
    default: return Dims<16>::spc(second, k);
  }
}
int ent_of(int lpp) { return 4 * lpp * (kNR / 4) + 1; }
int smax_of(const SgbmCfg& cfg, int k) { return (cfg.W1 - 1 + cfg.H - 1) / k; }
int groups_of(const SgbmCfg& cfg, bool second, int k) {
  const int spc = spc_of(lpp_of(cfg.D), second, k);
  return (smax_of(cfg, k) + 1 + spc - 1) / spc;
}
int max_groups(const SgbmCfg& cfg) {  // over both passes and both strip widths
  int g = 1;
  for (int k : {kLowK, kK}) g = std::max(g, std::max(groups_of(cfg, false, k), groups_of(cfg, true, k)));
  return g;
}
// hand-over buffers of one pass over `nframes` frames
size_t flag_ints(const SgbmCfg& cfg, int nframes) { return (size_t)nframes * max_groups(cfg) + 64; }
size_t edge_u4(const SgbmCfg& cfg, int nframes) {
  return (size_t)nframes * std::max(max_groups(cfg) - 1, 1) * (size_t)cfg.H * ent_of(lpp_of(cfg.D));
}

template <int LPP, int PATHS, bool SECOND, int K, bool WAY3, bool FULL, bool SMALL = false>
cudaError_t launch_full(const SweepArgs& a, int grid, cudaStream_t st) {
  using G = SweepGeom<LPP, SECOND, K>;
  auto* k = k_sweep<LPP, PATHS, SECOND, K, WAY3, FULL, SMALL>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
  if (e != cudaSuccess) return e;
  k<<<grid, G::NW * 32, G::SMEM, st>>>(a);
  return cudaGetLastError();
}

template <int LPP, int PATHS, bool SECOND, int K, bool WAY3 = false>
cudaError_t launch_nw(const SweepArgs& a, int grid, cudaStream_t st) {
  // the specialisations without idle lanes are built for the common modes only
  if constexpr (!WAY3) {
    if (a.nAct == LPP) {
      if constexpr (PATHS == 15) {
        if (a.smallSums) return launch_full<LPP, PATHS, SECOND, K, WAY3, true, true>(a, grid, st);
      }
      return launch_full<LPP, PATHS, SECOND, K, WAY3, true>(a, grid, st);
    }
  }
  return launch_full<LPP, PATHS, SECOND, K, WAY3, false>(a, grid, st);
}

template <int LPP, int PATHS, bool SECOND, bool WAY3 = false>
cudaError_t launch_one(const SweepArgs& a, int k, int grid, cudaStream_t st) {
  return k == kLowK ? launch_nw<LPP, PATHS, SECOND, kLowK, WAY3>(a, grid, st) : launch_nw<LPP, PATHS, SECOND, kK, WAY3>(a, grid, st);
}

template <int LPP>
cudaError_t launch_lpp(const SweepArgs& a, int paths, bool second, int k, bool way3, int grid, cudaStream_t st) {
  if (way3 && second) return launch_one<LPP, P_H, true, true>(a, k, grid, st);   // MODE_SGBM_3WAY: backward path + its selection rule
  if (!second) {
    if (paths == 15) return launch_one<LPP, 15, false>(a, k, grid, st);
    if (paths == (P_H | P_V)) return launch_one<LPP, P_H | P_V, false>(a, k, grid, st);
    return launch_one<LPP, P_H, false>(a, k, grid, st);   // backward pass of MODE_SGBM when the passes run side by side
  }
  if (paths == 15) return launch_one<LPP, 15, true>(a, k, grid, st);
  if (paths == (P_H | P_V)) return launch_one<LPP, P_H | P_V, true>(a, k, grid, st);
  return launch_one<LPP, P_H, true>(a, k, grid, st);
}

cudaError_t launch_sweep(const SweepArgs& a, int lpp, int paths, bool second, int k, bool way3, int grid, cudaStream_t st) {
#if defined(SV_SWEEP_ONLY8)   // quick experiment builds: only the D = 128 MODE_HH throughput kernels are instantiated
  if (lpp != 8 || paths != 15 || k != kK || way3) return cudaErrorNotSupported;
  return second ? launch_nw<8, 15, true, kK>(a, grid, st) : launch_nw<8, 15, false, kK>(a, grid, st);
#else
  switch (lpp) {
    case 1: return launch_lpp<1>(a, paths, second, k, way3, grid, st);
    case 2: return launch_lpp<2>(a, paths, second, k, way3, grid, st);
    case 4: return launch_lpp<4>(a, paths, second, k, way3, grid, st);
    case 8: return launch_lpp<8>(a, paths, second, k, way3, grid, st);
    default: return launch_lpp<16>(a, paths, second, k, way3, grid, st);
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// Winner-take-all over S = S1 + S2 when the two passes ran side by side (each as a "first" pass that stores the sum of its
// paths): the saturating add of A.5 and sweep_wta, one pixel per worker.  HBM-bound: 4 B per volume cell.
// ------------------------------------------------------------------------------------------------
template <int LPP>
__global__ void __launch_bounds__(256) k_wta_sum(const uint4* __restrict__ S1, const uint4* __restrict__ S2, uint2* __restrict__ sel,
                                                 long long npix, long long framePix, unsigned frameU4, int pixU4, int D, int nAct,
                                                 int uniqF, uint32_t uniqMagic) {
  constexpr int RW = 32 / LPP, NV = kNR / 4, VEC = LPP * NV;
  __shared__ uint4 sW[8 * RW * VEC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane / LPP, gl = lane % LPP;
  const bool active = gl < nAct;
  uint16_t* wta = (uint16_t*)(sW + (warp * RW + j) * VEC);
  const long long stride = (long long)gridDim.x * 8 * RW;
  for (long long p0 = ((long long)blockIdx.x * 8 + warp) * RW; p0 < npix; p0 += stride) {   // warp-uniform bounds
    const long long p = p0 + j;
    const bool on = p < npix;
    uint32_t s[kNR];
#pragma unroll
    for (int q = 0; q < kNR; q++) s[q] = 0;
    if (on && active) {
      const long long f = p / framePix, rem = p - f * framePix;
      const size_t off = (size_t)f * frameU4 + (size_t)rem * pixU4 + gl * NV;
#pragma unroll
      for (int h = 0; h < NV; h++) {
        const uint4 u = __ldcs(S1 + off + h), v = __ldcs(S2 + off + h);
        s[4 * h] = __viaddmin_u16x2(u.x, v.x, kMax2); s[4 * h + 1] = __viaddmin_u16x2(u.y, v.y, kMax2);
        s[4 * h + 2] = __viaddmin_u16x2(u.z, v.z, kMax2); s[4 * h + 3] = __viaddmin_u16x2(u.w, v.w, kMax2);
      }
    }
    uint2 o;
    sweep_wta<LPP, false>(s, wta, gl, active, lane, D, uniqF, uniqMagic, o);
    if (on && gl == 0) sel[p] = o;
  }
}
template <int LPP>
cudaError_t launch_wta_sum(const SweepArgs& a, const uint4* S2, int numSM, cudaStream_t st) {
  const long long framePix = (long long)a.H * a.W1, npix = framePix * a.nframes;
  const long long perBlock = 8 * (32 / LPP);
  const int grid = (int)std::min<long long>((npix + perBlock - 1) / perBlock, (long long)numSM * 32);
  k_wta_sum<LPP><<<grid, 256, 0, st>>>(a.S, S2, a.sel, npix, framePix, a.frameU4, a.pixU4, a.D, a.nAct, a.uniqF, a.uniqMagic);
  return cudaGetLastError();
}

int env_int(const char* name) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : -1;
}

}  // namespace

// S volumes of `nframes` pairs: up to kSplitFrames pairs per call run the two passes side by side, each into its own volume
// Up to this many pairs per call run the two passes side by side (see sgbm_aggregate_sweep).  Measured at 1080p / 128 on 148 SMs
// (24 strip groups per pair and pass): side by side wins up to 16 pairs per call (6.4 instead of 10.3 ms at 4, 17.3 instead of
// 17.8 ms at 16); beyond that the extra 4 B per cell of the separate selection cost more than the fuller machine gains.
// SVGPU_SPLIT_FRAMES overrides the limit (experiments).
static int split_limit(const SgbmCfg& cfg, int numSM) {
  const int v = env_int("SVGPU_SPLIT_FRAMES");
  if (v >= 0) return v;
  if (cfg.W1 <= 0 || cfg.mode == SV_MODE_SGBM_3WAY || cfg.D > kDPL * 16) return 0;
  return (int)(5ll * numSM / (2ll * groups_of(cfg, false, kK)));
}
static int num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}
size_t sgbm_s_frames(const SgbmCfg& cfg, int nframes) { return nframes <= split_limit(cfg, num_sms()) ? 2 * (size_t)nframes : (size_t)nframes; }
size_t sgbm_flags_bytes(const SgbmCfg& cfg, int nframes) {
  if (cfg.W1 <= 0 || cfg.D > kDPL * 16) return 256;
  return 2 * flag_ints(cfg, nframes) * sizeof(int);
}
size_t sgbm_edges_bytes(const SgbmCfg& cfg, int nframes) {
  if (cfg.W1 <= 0 || cfg.D > kDPL * 16) return 256;
  return edge_u4(cfg, (int)sgbm_s_frames(cfg, nframes)) * sizeof(uint4);
}

// Aggregation + selection of `nframes` pairs: S scratch and `sel` are produced; C is read twice.
//
// Throughput (many pairs per call): two launches on `st`, the second carries its four paths, adds the first pass's S and selects
// in registers (8 B per cell).  Few pairs per call: the machine is not full and a call takes as long as its wavefronts —
// H + (W1 + H) / K row steps per pass.  The second pass does not depend on the first until the sums meet, so the two passes
// then run side by side — on `st` and on sc.aux, each storing the sum of its paths — and k_wta_sum adds and selects
// (min(min(a, M) + min(b, M), M) = min(a + b, M): the saturation of the fused form).  When the tasks of both passes fit on
// the SMs at once (1-3 pairs at 1080p) the strips are 2 columns wide, which shortens the wavefront further.
cudaError_t sgbm_aggregate_sweep(const SgbmCfg& cfg, const SgbmScratch& sc, int nframes, int debugS, cudaStream_t st,
                                 LaunchStats* stats, int cn) {
  // a single small pair: one worker per line and direction instead of the strip pipeline (sgbm_lines.cu)
  if (!debugS && sgbm_lines_ok(cfg, nframes, sc.frameCells) && sc.sBytes >= sgbm_lines_frames(cfg, nframes) * sc.frameCells * 2)
    return sgbm_aggregate_lines(cfg, sc, nframes, st, stats);
  if (cfg.D > kDPL * 16) {   // beyond 16 lanes per pixel only the line form exists (it walks the call pair by pair)
    if (debugS || sc.sBytes < sgbm_lines_frames(cfg, 1) * sc.frameCells * 2) return cudaErrorNotSupported;
    return sgbm_aggregate_lines(cfg, sc, nframes, st, stats);
  }
  int dev = 0, numSM = 0;
  cudaError_t e0 = cudaGetDevice(&dev);
  if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, dev);
  if (e0 != cudaSuccess) return e0;
  SweepArgs a{};
  a.C = (const uint4*)sc.C; a.S = (uint4*)sc.S; a.sel = (uint2*)sc.sel;
  a.status = sc.status;
  a.W1 = cfg.W1; a.H = cfg.H; a.D = cfg.D; a.nAct = cfg.D / kDPL;
  // largest block cost of single-channel input: blockSize^2 x (2 ftzero + 63); every path cost is <= that + P2 (A.5)
  a.smallSums = cn <= 1 && 4ll * ((2 * cfg.r + 1) * (2 * cfg.r + 1) * (2 * cfg.ftzero + 63) + cfg.P2) <= 32767;
  a.nframes = nframes; a.debugS = debugS;
  a.frameU4 = (unsigned)(sc.frameCells / 8);
  a.pixU4 = cfg.D / 8;
  a.rowU4 = (unsigned)cfg.W1 * (unsigned)a.pixU4;
  a.p1dup = dup16((uint32_t)cfg.P1);
  a.p2mp1dup = dup16((uint32_t)(cfg.P2 - cfg.P1));
  a.negp1dup = dup16((uint32_t)(-cfg.P1) & 0xFFFFu);
  a.one = 1u;
  a.uniqF = 100 - cfg.uniq;
  a.uniqMagic = a.uniqF >= 2 ? (uint32_t)((0x100000000ull + a.uniqF - 1) / a.uniqF) : 0u;
  // MODE_SGBM_3WAY (on stripe frames, see sgbm.cu): left-to-right + top-down in pass 1, right-to-left in pass 2
  const bool way3 = cfg.mode == SV_MODE_SGBM_3WAY;
  const int pass1 = (cfg.mode == SV_MODE_HH4 || way3) ? (P_H | P_V) : 15;
  const int pass2 = cfg.mode == SV_MODE_HH ? 15 : (cfg.mode == SV_MODE_HH4 ? (P_H | P_V) : P_H);
  const int lpp = lpp_of(cfg.D);
  // SVGPU_SWEEP_K=2|4 and SVGPU_SWEEP_SPLIT=0|1 pin the variant (the tests run every case through all of them)
  const int forceK = env_int("SVGPU_SWEEP_K"), forceSplit = env_int("SVGPU_SWEEP_SPLIT");
  const size_t fInts = flag_ints(cfg, nframes), eU4 = edge_u4(cfg, nframes);
  const bool canSplit = !way3 && !debugS && sc.aux && sc.evFork && sc.evJoin && forceSplit != 0 &&
                        sc.sBytes >= 2 * (size_t)nframes * sc.frameCells * 2 && sc.flagBytes >= 2 * fInts * sizeof(int) &&
                        sc.edgeBytes >= 2 * eU4 * sizeof(uint4);
  int K = kK;
  bool split = false;
#if !defined(SV_SWEEP_ONLY8)
  if (canSplit && 2ll * nframes * groups_of(cfg, false, kLowK) <= numSM) { split = true; K = kLowK; }
  else if (canSplit && nframes <= split_limit(cfg, numSM)) split = true;
  // passes one after the other, measured at 1080p / 128: 2-column strips win up to 8 pairs per call (192 tasks of the 4-column
  // build), 4-column strips from 12
  else if (2ll * nframes * groups_of(cfg, true, kK) <= 3ll * numSM) K = kLowK;
  if (forceK == kLowK || forceK == kK) K = forceK;
  if (canSplit && forceSplit == 1) split = true;
#endif
  a.smax = smax_of(cfg, K);
  if (split) {
    a.G = groups_of(cfg, false, K);
    const int grid = std::min(nframes * a.G, numSM);
    cudaError_t e = cudaMemsetAsync(sc.flags, 0, 2 * fInts * sizeof(int), st);
    if (e == cudaSuccess) e = cudaEventRecord(sc.evFork, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(sc.aux, sc.evFork, 0);
    if (e != cudaSuccess) return e;
    SweepArgs b = a;
    a.ticket = (int*)sc.flags; a.gprog = a.ticket + 64; a.edges = (uint4*)sc.edges; a.flip = 0;
    b.ticket = (int*)sc.flags + fInts; b.gprog = b.ticket + 64; b.edges = (uint4*)sc.edges + eU4; b.flip = 1;
    b.S = a.S + (size_t)nframes * a.frameU4;
    e = launch_sweep(a, lpp, pass1, false, K, false, grid, st);
    if (e == cudaSuccess) e = launch_sweep(b, lpp, pass2, false, K, false, grid, sc.aux);
    // the join is recorded even when a launch failed, so that `st` never waits for an event of an earlier call
    cudaError_t e2 = cudaEventRecord(sc.evJoin, sc.aux);
    if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(st, sc.evJoin, 0);
    if (e != cudaSuccess) return e;
    if (e2 != cudaSuccess) return e2;
    switch (lpp) {
      case 1: e = launch_wta_sum<1>(a, b.S, numSM, st); break;
      case 2: e = launch_wta_sum<2>(a, b.S, numSM, st); break;
      case 4: e = launch_wta_sum<4>(a, b.S, numSM, st); break;
      case 8: e = launch_wta_sum<8>(a, b.S, numSM, st); break;
      default: e = launch_wta_sum<16>(a, b.S, numSM, st); break;
    }
    if (stats) stats->launches += 3;
    return e;
  }
  a.edges = (uint4*)sc.edges;
  a.ticket = (int*)sc.flags;
  a.gprog = (int*)sc.flags + 64;
  for (int pass = 0; pass < 2; pass++) {
    a.G = groups_of(cfg, pass == 1, K);
    cudaError_t e = cudaMemsetAsync(sc.flags, 0, ((size_t)nframes * a.G + 64) * sizeof(int), st);
    if (e != cudaSuccess) return e;
    a.flip = pass;
    const int grid = std::min(nframes * a.G, numSM);
    e = launch_sweep(a, lpp, pass ? pass2 : pass1, pass == 1, K, way3, grid, st);
    if (stats) stats->launches++;
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace svgpu

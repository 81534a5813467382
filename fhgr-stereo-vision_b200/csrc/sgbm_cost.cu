// sgbm_cost.cu — SGBM matching cost volume: Birchfield-Tomasi pixel cost (A.3) + blockSize x blockSize box sum (A.4).
//
// Replaces the cost computation inside cv::StereoSGBM::compute (/root/reference/src/lib.cpp:262).  Input: the
// Birchfield-Tomasi operand planes of both images (sgbm.cu: k_prefilter_planes); output: Ctrue[y][x'][d] uint16 in the
// interleaved volume layout of sgbm_device.cuh (word i of a 16-disparity group = d0+i | (d0+i+8) << 16).
//
// One CTA produces a tile of NC - 2r output columns x a band of rows x all D disparities.
//   pixel cost   thread (column c, group g of 16 disparities): lanes of a warp are neighbouring columns, so the
//                right-image operands they read from shared memory are neighbouring 16-byte records.  Those
//                records are staged per row with the two disparities of a packed register already paired
//                (positions j and j+8) and with the negated copies the arithmetic needs, so that one cost of
//                one channel is 5 packed instructions:
//                    p = min(u - vhi, ulo - v), q = min(vlo - u, v - uhi), cost = max(0, p, q)
//                (equal to min(max(0, u-vhi, vlo-u), max(0, v-uhi, ulo-v)) because vlo <= v <= vhi, ulo <= u <= uhi);
//   vertical sum thread-private: the last 2r+1 rows of pixel costs sit packed as bytes in a register ring
//                (the row loop is unrolled over the ring), the running sum Vs in 8 registers;
//   horizontal   every kRB rows the Vs rows go through shared memory to a second thread mapping — lanes are
//                (disparity half-group, row) pairs that slide a 2r+1 window along x — which writes C with
//                256 contiguous bytes per pixel.
// Columns are clamped in W1 coordinates and rows in image coordinates by computing the clamped position again
// (A.4: the window replicates the border).
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "sgbm_device.cuh"


namespace svgpu {

namespace {

struct CostGeom {
  int NJ;      // right-image records per row: reversed positions 0 .. NC + D - 10
  int ng;      // groups of 16 disparities
  int rowU4;   // uint4 per (row, half-group) line of the Vs buffer: NC + 1 (padding against bank conflicts)
  size_t offL, offR, offV, offRing, offRaw, offBar, total;  // byte offsets in dynamic shared memory
  int nLraw, nRraw, rawSlotU2;  // TMA staging: operand words per plane of a row slot (left / right) and uint2 per slot
};
constexpr int kRawSlots = 3;   // row slots of the TMA staging ring: the row in use, the next one, the one in flight

// CH = operand planes per image row: 2 (gradient, intensity) per colour channel
// SRING: the pixel costs of the last 2r+1 rows live in shared memory (thread-private slots) instead of registers
// TMAST: the operand words of a row tile are brought in by TMA bulk copies (cp.async.bulk + mbarrier) two rows ahead and
//        turned into records from shared memory; otherwise every thread loads its words with LDG one row ahead
template <int NC, int RB, int CH, bool SRING, bool TMAST>
__host__ __device__ inline CostGeom cost_geom(int D, int R) {
  CostGeom g;
  g.ng = D / kDPL;
  g.NJ = (NC + D - 9 + 1) & ~1;   // even: the records of even and odd positions form two planes when a thread owns two columns
  g.rowU4 = NC + 1;
  g.offL = 0;                                        // [2 rows][CH planes][NC] uint4
  g.offR = g.offL + (size_t)2 * CH * NC * 16;        // [2 rows][CH planes][NJ] uint4
  g.offV = g.offR + (size_t)2 * CH * g.NJ * 16;      // [RB][ng*2][NC+1] uint4
  g.offRing = g.offV + (size_t)RB * g.ng * 2 * g.rowU4 * 16;   // [2R+1][words per cell block][NC x ng threads-columns] uint4
  g.offRaw = g.offRing + (SRING ? (size_t)(2 * R + 1) * (CH == 2 ? 1 : 2) * NC * g.ng * 16 : 0);
  g.nLraw = NC + 2;            // even counts: bulk copies move multiples of 16 bytes
  g.nRraw = g.NJ + 8 + 2;
  g.rawSlotU2 = CH * (g.nLraw + g.nRraw);
  g.offBar = g.offRaw + (TMAST ? (size_t)kRawSlots * g.rawSlotU2 * 8 : 0);
  g.total = g.offBar + (TMAST ? 64 : 0);
  return g;
}

// cost of one channel for two disparities: l = {ulo, u, -uhi, -u} (both halves equal), r = {-v, -vhi, v, vlo}
__device__ __forceinline__ uint32_t bt_cost(const uint4& l, const uint4& r) {
  const uint32_t p = __viaddmin_s16x2(r.y, l.y, __vadd2(r.x, l.x));  // min(u - vhi, ulo - v)
  const uint32_t q = __viaddmin_s16x2(r.w, l.w, __vadd2(r.z, l.z));  // min(vlo - u, v - uhi)
  return __vimax_s16x2_relu(p, q);
}

// COLS neighbouring columns per thread: 8 of the 9 right-image records of two neighbouring columns are the same,
// so two columns per thread nearly halve the shared-memory traffic of the cost arithmetic.
template <int R, int NC, int COLS, int RB, int ITEMS, int CH, bool SRING, int MAXT, bool TMAST>
__global__ void __launch_bounds__(MAXT, 1) k_cost_tile(const uint2* __restrict__ planesL, const uint2* __restrict__ planesR,
                                                                uint4* __restrict__ Cvol, const SgbmCfg cfg, int rowsPerBand, size_t frameU4,
                                                                int planeRows, int rowsOut) {
  constexpr int NRING = 2 * R + 1, TW = NC - 2 * R, NCP = NC / COLS;
  extern __shared__ uint4 smemU4[];
  const int D = cfg.D, W1 = cfg.W1, H = cfg.H, Wp = cfg.Wp;
  const CostGeom G = cost_geom<NC, RB, CH, SRING, TMAST>(D, R);
  uint4* sL = smemU4 + G.offL / 16;
  uint4* sRing = smemU4 + G.offRing / 16;
  uint4* sR = smemU4 + G.offR / 16;
  uint4* sV = smemU4 + G.offV / 16;
  const int NJ = G.NJ, ng = G.ng;

  const int tid = threadIdx.x, nthr = blockDim.x;  // nthr = NCP * ng
  const int frame = blockIdx.z;
  const int x0 = blockIdx.x * TW;                  // first output column x' of the tile
  // rows [0, rowsOut) of the frame are produced; the window clamps at rows 0 and H - 1 (H may exceed rowsOut: the stripe
  // frames of MODE_SGBM_3WAY take the rows below them from the image); planeRows = rows per frame in the operand planes
  const int yb = blockIdx.y * rowsPerBand, ye = min(yb + rowsPerBand, rowsOut);
  if (yb >= ye) return;
  const uint2* pL = planesL + (size_t)frame * planeRows * CH * Wp;
  const uint2* pR = planesR + (size_t)frame * planeRows * CH * Wp;
  uint4* Cf = Cvol + (size_t)frame * frameU4;

  // ---- pixel-cost mapping: lanes = neighbouring column groups, thread = (COLS columns, 16 disparities)
  const int cp = tid % NCP, g = tid / NCP;
  const int c0 = cp * COLS;                                   // first column of this thread (tile column 0 is x' = x0 - R)
  const int j0 = (NC - COLS - c0) + kDPL * g;                 // reversed right-image position of (last column, d = 16 g)
  const int xrMax = x0 - R + NC - 1 + cfg.minX1 - cfg.minD;   // right-image x of reversed position 0
  // position of tile column c inside a line of the Vs buffer: the columns of one thread go to different planes so
  // that the lanes of a warp store neighbouring words
  auto vpos = [](int c) { return COLS == 1 ? c : (c % COLS) * NCP + c / COLS; };
  // the same for the staged records: a thread reads the records of COLS neighbouring columns / positions, so even and odd
  // ones live in separate planes and the lanes of a warp read neighbouring 16-byte words of one plane
  auto lpos = [](int c) { return COLS == 1 ? c : (c % COLS) * NCP + c / COLS; };
  auto rpos = [NJ](int j) { return COLS == 1 ? j : (j % COLS) * (NJ / COLS) + j / COLS; };
  // ---- horizontal mapping: lanes = (half-group, row), threads of one lane slot walk a run of columns.  The window
  //      replicates the border columns of the volume (A.4): indices are clamped to the columns of x' = 0 .. W1-1
  const int gh = tid % (ng * 2), hrow = (tid / (ng * 2)) % RB, run = tid / (ng * 2 * RB);
  constexpr int NRUN = NC / (COLS * 2 * RB), RL = (TW + NRUN - 1) / NRUN;
  const int cLo = max(R - x0, 0), cHi = min(W1 - 1 - x0 + R, NC - 1);

  // ---- staging of the per-row operands: every thread owns up to ITEMS records; the operand words are loaded early
  //      (stage_load) and paired into records late (stage_store) so that the loads overlap the cost arithmetic.
  //      Left record of a column = the operand words with each 16-bit field doubled; right record of reversed
  //      position j = the fields of positions j and j+8 side by side.
  int itOffA[ITEMS], itOffB[ITEMS], itDst[ITEMS];  // operand index inside the row's plane block, record index (-1: none)
  uint32_t itLeft = 0;
  // TMA staging: the words of image columns [aL, aL + nL) (left) and [aR, aR + nR) (right) of every plane of a row go into a
  // slot of the raw ring; both ranges start at an even column (16-byte aligned source) and cover what the tile touches
  const int xLfirst = x0 - R + cfg.minX1;
  const int aL = max(xLfirst, 0) & ~1, bL = min(xLfirst + NC, cfg.W), nL = (bL - aL + 1) & ~1;
  const int aR = max(xrMax - (NJ - 1) - 8, 0) & ~1, bR = min(xrMax + 1, cfg.W), nR = (bR - aR + 1) & ~1;
#pragma unroll
  for (int q = 0; q < ITEMS; q++) {
    const int it = tid + q * nthr;
    itOffA[q] = itOffB[q] = 0;
    itDst[q] = -1;
    if (it < CH * NC) {
      const int ch = it / NC, cc = it % NC;
      // columns outside the volume are computed on some valid pixel and ignored by the horizontal phase
      if constexpr (TMAST) itOffA[q] = ch * G.nLraw + min(max(xLfirst + cc, aL), bL - 1) - aL;
      else itOffA[q] = ch * Wp + min(max(xLfirst + cc, 0), cfg.W - 1);
      itDst[q] = ch * NC + lpos(cc);
      itLeft |= 1u << q;
    } else if (it < CH * NC + CH * NJ) {
      const int k = it - CH * NC, ch = k / NJ, j = k % NJ;
      // positions outside the image only ever pair with cells outside the volume: any in-range word will do
      if constexpr (TMAST) {
        itOffA[q] = CH * G.nLraw + ch * G.nRraw + min(max(xrMax - j, aR), bR - 1) - aR;
        itOffB[q] = CH * G.nLraw + ch * G.nRraw + min(max(xrMax - j - 8, aR), bR - 1) - aR;
      } else {
        itOffA[q] = ch * Wp + min(max(xrMax - j, 0), cfg.W - 1);
        itOffB[q] = ch * Wp + min(max(xrMax - j - 8, 0), cfg.W - 1);
      }
      itDst[q] = ch * NJ + rpos(j);
    }
  }
  auto make_left = [](const uint2 a) {
    return make_uint4(__byte_perm(a.x, 0, 0x1010), __byte_perm(a.x, 0, 0x3232), __byte_perm(a.y, 0, 0x1010), __byte_perm(a.y, 0, 0x3232));
  };
  auto make_right = [](const uint2 a, const uint2 b) {
    return make_uint4(__byte_perm(a.x, b.x, 0x5410), __byte_perm(a.x, b.x, 0x7632), __byte_perm(a.y, b.y, 0x5410), __byte_perm(a.y, b.y, 0x7632));
  };
  uint2 stA[TMAST ? 1 : ITEMS], stB[TMAST ? 1 : ITEMS];
  auto stage_load = [&](int yy) {
    const uint2* rowL = pL + (size_t)yy * CH * Wp;
    const uint2* rowR = pR + (size_t)yy * CH * Wp;
#pragma unroll
    for (int q = 0; q < (TMAST ? 0 : ITEMS); q++) {
      if (itDst[q] < 0) continue;
      if ((itLeft >> q) & 1) {
        stA[q] = __ldg(rowL + itOffA[q]);
      } else {
        stA[q] = __ldg(rowR + itOffA[q]);
        stB[q] = __ldg(rowR + itOffB[q]);
      }
    }
  };
  auto stage_store = [&](int b) {
#pragma unroll
    for (int q = 0; q < (TMAST ? 0 : ITEMS); q++) {
      if (itDst[q] < 0) continue;
      if ((itLeft >> q) & 1) sL[b * CH * NC + itDst[q]] = make_left(stA[q]);
      else sR[b * CH * NJ + itDst[q]] = make_right(stA[q], stB[q]);
    }
  };
  // ---- TMA staging (TMAST): thread 0 requests a row, everybody turns its words into records once they have landed
  const uint2* sRaw = (const uint2*)(smemU4 + G.offRaw / 16);
  uint64_t* sBar = (uint64_t*)(smemU4 + G.offBar / 16);
  const int vbeg0 = yb - R, vend0 = ye - 1 + R;
  const int yFirst = min(max(vbeg0, 0), H - 1), yLast = min(max(vend0, 0), H - 1);   // image rows this band reads
  auto tma_request = [&](int yy) {   // thread 0 only
    const int slot = (yy - yFirst) % kRawSlots;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(sBar + slot);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sRaw + (size_t)slot * G.rawSlotU2);
    const uint32_t bytes = (uint32_t)(CH * (nL + nR)) * 8u;
    // the slot was last read (generic proxy) before the CTA barrier that precedes this call
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll
    for (int ch = 0; ch < CH; ch++) {
      const uint2* gl = pL + ((size_t)yy * CH + ch) * Wp + aL;
      const uint2* gr = pR + ((size_t)yy * CH + ch) * Wp + aR;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + (uint32_t)(ch * G.nLraw) * 8u),
                   "l"(gl), "r"((uint32_t)nL * 8u), "r"(bar) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       dst + (uint32_t)(CH * G.nLraw + ch * G.nRraw) * 8u), "l"(gr), "r"((uint32_t)nR * 8u), "r"(bar) : "memory");
    }
  };
  auto tma_build = [&](int yy) {     // all threads: wait for the row, write its records into buffer yy & 1
    const int seq = yy - yFirst, slot = seq % kRawSlots, parity = (seq / kRawSlots) & 1;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(sBar + slot);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    const uint2* raw = sRaw + (size_t)slot * G.rawSlotU2;
    const int b = yy & 1;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
      if (itDst[q] < 0) continue;
      if ((itLeft >> q) & 1) sL[b * CH * NC + itDst[q]] = make_left(raw[itOffA[q]]);
      else sR[b * CH * NJ + itDst[q]] = make_right(raw[itOffA[q]], raw[itOffB[q]]);
    }
  };

  // the pixel costs of the last 2r+1 rows: one byte per cost for a single channel (<= 2*63 + 63), 16 bits for three
  constexpr bool RING8 = CH == 2;
  constexpr int RINGW = RING8 ? kNR / 2 : kNR;
  constexpr int RINGU4 = RINGW / 4;              // uint4 per (slot, column) of the shared-memory ring
  uint32_t Vs[COLS][kNR], pix[COLS][kNR], ring[SRING ? 1 : NRING][COLS][RINGW];
  // shared-memory ring: slot k of column m of this thread = RINGU4 words at sRing[((k * COLS + m) * RINGU4 + w) * nthr + tid]
#pragma unroll
  for (int m = 0; m < COLS; m++) {
#pragma unroll
    for (int i = 0; i < kNR; i++) Vs[m][i] = pix[m][i] = 0;
#pragma unroll
    for (int k = 0; k < NRING; k++) {
      if constexpr (SRING) {
#pragma unroll
        for (int w = 0; w < RINGU4; w++) sRing[((k * COLS + m) * RINGU4 + w) * nthr + tid] = make_uint4(0, 0, 0, 0);
      } else {
#pragma unroll
        for (int i = 0; i < RINGW; i++) ring[k][m][i] = 0;
      }
    }
  }

  const int vbeg = yb - R, vend = ye - 1 + R;  // virtual rows whose pixel costs enter the band's sums
  int staged = min(max(vbeg, 0), H - 1);       // image row in buffer (staged & 1)
  if constexpr (TMAST) {
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < kRawSlots; k++)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(sBar + k)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      tma_request(yFirst);
      if (yFirst + 1 <= yLast) tma_request(yFirst + 1);
    }
    tma_build(staged);
  } else {
    stage_load(staged);
    stage_store(staged & 1);
  }
  __syncthreads();
  int computed = -1;                            // image row whose pixel costs are in `pix`

  for (int v0 = vbeg; v0 <= vend; v0 += NRING) {
#pragma unroll(SRING ? 1 : NRING)
    for (int k = 0; k < NRING; k++) {
      const int v = v0 + k;
      if (v > vend) break;
      const int yy = min(max(v, 0), H - 1);
      // ---- the next image row is staged while this one is consumed: loads now, records after the arithmetic
      const int ynext = min(max(v + 1, 0), H - 1);
      const bool doStage = v < vend && ynext != staged;
      if (doStage) {
        if constexpr (TMAST) { if (tid == 0 && ynext + 1 <= yLast) tma_request(ynext + 1); }
        else stage_load(ynext);
      }
      // ---- pixel costs of image row yy (rows beyond the image border repeat the border row).  Column c0 + m reads the
      //      records j0 + (COLS-1-m) + i, i = 0..7: record q of this thread serves register q - (COLS-1-m) of column m
      if (yy != computed) {
#pragma unroll
        for (int ch = 0; ch < CH; ch++) {
          const uint4* lrec = sL + ((yy & 1) * CH + ch) * NC;
          const uint4* rrec = sR + ((yy & 1) * CH + ch) * NJ;
          uint4 lu[COLS];
#pragma unroll
          for (int m = 0; m < COLS; m++) lu[m] = lrec[lpos(c0 + m)];
#pragma unroll
          for (int q = 0; q < kNR + COLS - 1; q++) {
            const uint4 rr = rrec[rpos(j0 + q)];
#pragma unroll
            for (int m = 0; m < COLS; m++) {
              const int i = q - (COLS - 1 - m);
              if (i >= 0 && i < kNR) {
                const uint32_t cst = bt_cost(lu[m], rr);
                if (ch == 0) pix[m][i] = cst;
                else if (ch & 1) pix[m][i] += (cst >> 2) & 0x3FFF3FFFu;  // intensity term: a quarter of the weight
                else pix[m][i] += cst;
              }
            }
          }
        }
        computed = yy;
      }
      if (doStage) {
        if constexpr (TMAST) tma_build(ynext);
        else stage_store(ynext & 1);
        staged = ynext;
      }
      // ---- vertical running sum: + this row, - the row that leaves the window (ring slot k)
      const int y = v - R;  // output row completed by this virtual row
#pragma unroll
      for (int m = 0; m < COLS; m++) {
        constexpr int KR = SRING ? 0 : 1;   // register ring: slot k; shared-memory ring: one staging copy
        if constexpr (SRING) {
#pragma unroll
          for (int w = 0; w < RINGU4; w++) {
            const uint4 o = sRing[((k * COLS + m) * RINGU4 + w) * nthr + tid];
            ring[0][m][4 * w] = o.x; ring[0][m][4 * w + 1] = o.y; ring[0][m][4 * w + 2] = o.z; ring[0][m][4 * w + 3] = o.w;
          }
        }
        if constexpr (RING8) {
#pragma unroll
          for (int i = 0; i < kNR / 2; i++) {
            const uint32_t old = ring[k * KR][m][i];
            Vs[m][2 * i] += pix[m][2 * i] - __byte_perm(old, 0, 0x4140);
            Vs[m][2 * i + 1] += pix[m][2 * i + 1] - __byte_perm(old, 0, 0x4342);
            ring[k * KR][m][i] = __byte_perm(pix[m][2 * i], pix[m][2 * i + 1], 0x6420);
          }
        } else {
#pragma unroll
          for (int i = 0; i < kNR; i++) {
            Vs[m][i] += pix[m][i] - ring[k * KR][m][i];
            ring[k * KR][m][i] = pix[m][i];
          }
        }
        if constexpr (SRING) {
#pragma unroll
          for (int w = 0; w < RINGU4; w++)
            sRing[((k * COLS + m) * RINGU4 + w) * nthr + tid] =
                make_uint4(ring[0][m][4 * w], ring[0][m][4 * w + 1], ring[0][m][4 * w + 2], ring[0][m][4 * w + 3]);
        }
        if (y >= yb) {
          uint4* dst = sV + ((size_t)((y - yb) % RB) * ng * 2 + g * 2) * G.rowU4 + vpos(c0 + m);
          dst[0] = make_uint4(Vs[m][0], Vs[m][1], Vs[m][2], Vs[m][3]);
          dst[G.rowU4] = make_uint4(Vs[m][4], Vs[m][5], Vs[m][6], Vs[m][7]);
        }
      }
      __syncthreads();  // staged operands and the Vs row are visible
      if (y >= yb && (((y - yb) % RB) == RB - 1 || y == ye - 1)) {
        // ---- horizontal window over the buffered Vs rows -> Ctrue
        const int yrow = y - ((y - yb) % RB) + hrow;  // image row of this thread's line
        if (yrow <= y && run < NRUN) {
          const bool zero = (cfg.mode == SV_MODE_HH4) && (yrow >= H - R);  // A.4: HH4 leaves the bottom rows at 0
          const uint4* line = sV + ((size_t)hrow * ng * 2 + gh) * G.rowU4;
          auto at = [&](int c) { return line[vpos(min(max(c, cLo), cHi))]; };
          const int t0 = run * RL, t1 = min(min(t0 + RL, TW), W1 - x0);
          // (a clamp-free copy of this loop for runs that never touch the border columns, walking the two planes of a line with
          // two pointers, measured slower: cost stage 15.9 vs 15.4 ms per 64 pairs)
          if (t0 < t1) {
            uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int dx = 0; dx < NRING; dx++) {
              const uint4 q = at(t0 + dx);
              acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
            }
            uint4* out = Cf + ((size_t)yrow * W1 + x0) * (D / 8) + gh;
            for (int t = t0;; t++) {
              __stcs(out + (size_t)t * (D / 8), zero ? make_uint4(0, 0, 0, 0) : acc);
              if (t + 1 >= t1) break;
              const uint4 in = at(t + NRING), ou = at(t);
              acc.x += in.x - ou.x; acc.y += in.y - ou.y; acc.z += in.z - ou.z; acc.w += in.w - ou.w;
            }
          }
        }
        __syncthreads();  // the Vs buffer may be overwritten
      }
    }
  }
}

template <int R, int NC, int COLS, int RB, int ITEMS, int CH, bool SRING, int MAXT, bool TMAST = false>
cudaError_t launch_cost(const SgbmCfg& cfg, const CostView& v, int nframes, cudaStream_t st) {
  const CostGeom G = cost_geom<NC, RB, CH, SRING, TMAST>(cfg.D, R);
  const int threads = (NC / COLS) * (cfg.D / kDPL);
  if (G.total > 227 * 1024 || threads > MAXT) return cudaErrorInvalidConfiguration;   // the callers pick configurations that fit
  auto* k = k_cost_tile<R, NC, COLS, RB, ITEMS, CH, SRING, MAXT, TMAST>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.total);
  if (e != cudaSuccess) return e;
  constexpr int TW = NC - 2 * R;
  const int tiles = (cfg.W1 + TW - 1) / TW;
  // bands: the CTA count should fill the 148 SMs (one CTA each) in whole waves, against the 2r-row run-in that every
  // extra band costs.  Estimated launch time in row steps = waves x (rows per band + run-in).
  int bands = 1, rowsPerBand = (v.rowsOut + RB - 1) / RB * RB;
  {
    long long best = -1;
    const int maxBands = std::max(1, std::min(64, v.rowsOut / 8));   // a single small pair is cut into enough bands to reach every SM
    for (int b = 1; b <= maxBands; b++) {
      const int rpb = ((v.rowsOut + b - 1) / b + RB - 1) / RB * RB;
      const int nb = (v.rowsOut + rpb - 1) / rpb;
      const long long waves = ((long long)tiles * nframes * nb + 147) / 148;
      const long long cost = waves * (rpb + 2 * R + 4);
      if (best < 0 || cost < best) { best = cost; bands = nb; rowsPerBand = rpb; }
    }
  }
  dim3 grid(tiles, bands, nframes);
  k<<<grid, threads, G.total, st>>>((const uint2*)v.planesL, (const uint2*)v.planesR, (uint4*)v.C, cfg, rowsPerBand, v.frameCells / 8, v.planeRows,
                                    v.rowsOut);
  return cudaGetLastError();
}

// general kernel: one column per thread, the row ring in registers, any block radius / channel count
template <int NC, int COLS, int RB, int ITEMS, int CH = 2>
cudaError_t launch_cost_r(const SgbmCfg& cfg, const CostView& sc, int nframes, cudaStream_t st) {
  switch (cfg.r) {
    case 0: return launch_cost<0, NC, COLS, RB, ITEMS, CH, false, 512>(cfg, sc, nframes, st);
    case 1: return launch_cost<1, NC, COLS, RB, ITEMS, CH, false, 512>(cfg, sc, nframes, st);
    case 2: return launch_cost<2, NC, COLS, RB, ITEMS, CH, false, 512>(cfg, sc, nframes, st);
    case 3: return launch_cost<3, NC, COLS, RB, ITEMS, CH, false, 512>(cfg, sc, nframes, st);
    case 4: return launch_cost<4, NC, COLS, RB, ITEMS, CH, false, 512>(cfg, sc, nframes, st);
    default: return launch_cost<5, NC, COLS, RB, ITEMS, CH, false, 512>(cfg, sc, nframes, st);
  }
}
// throughput kernel for single-channel input with blockSize <= 7 and 64..256 disparities: two neighbouring columns per thread
// (9 instead of 16 right-image records per 16 pixel costs), the row ring in shared memory (no register spills at 128
// registers), horizontal phase every two rows
#ifndef SV_COST_TMA
#define SV_COST_TMA 1
#endif
template <int NC, int ITEMS, bool TMAST, int RB = 2>
cudaError_t launch_cost_pair(const SgbmCfg& cfg, const CostView& sc, int nframes, cudaStream_t st) {
  switch (cfg.r) {
    case 0: return launch_cost<0, NC, 2, RB, ITEMS, 2, true, 512, TMAST>(cfg, sc, nframes, st);
    case 1: return launch_cost<1, NC, 2, RB, ITEMS, 2, true, 512, TMAST>(cfg, sc, nframes, st);
    case 2: return launch_cost<2, NC, 2, RB, ITEMS, 2, true, 512, TMAST>(cfg, sc, nframes, st);
    default: return launch_cost<3, NC, 2, RB, ITEMS, 2, true, 512, TMAST>(cfg, sc, nframes, st);
  }
}

}  // namespace

// Ctrue of `nframes` pairs from their pre-filter planes
cudaError_t sgbm_cost_volume(const SgbmCfg& cfg, const CostView& sc, int nframes, int cn, cudaStream_t st) {
  // records staged per thread and row = (CH NC + CH NJ) / threads, CH = 2 planes per colour channel
  if (cfg.D > 256) {
    // 272 .. 512 disparities: 17 .. 32 disparity groups leave a 512-thread CTA 16 columns per tile (general builds, not tuned)
    if (cfg.r >= 6) {
      if (cn != 1 || cfg.r > 7) return cudaErrorInvalidConfiguration;
      return cfg.r == 6 ? launch_cost<6, 16, 1, 2, 3, 2, true, 512>(cfg, sc, nframes, st) : launch_cost<7, 16, 1, 2, 3, 2, true, 512>(cfg, sc, nframes, st);
    }
    if (cn == 3) return launch_cost_r<16, 1, 4, 7, 6>(cfg, sc, nframes, st);
    return launch_cost_r<16, 1, 8, 3>(cfg, sc, nframes, st);
  }
  if (cfg.r >= 6) {
    // blockSize 13 / 15 (the 16-bit cost contract admits nothing larger, and single-channel input only): the general build with
    // its 13- / 15-row ring in shared memory and a two-row horizontal block
    auto go = [&](auto rtag) {
      constexpr int RR = decltype(rtag)::value;
      if (cfg.D > 128) return launch_cost<RR, 32, 1, 2, 2, 2, true, 512>(cfg, sc, nframes, st);
      if (cfg.D == 128) return launch_cost<RR, 64, 1, 2, 1, 2, true, 512>(cfg, sc, nframes, st);
      if (cfg.D >= 48) return launch_cost<RR, 64, 1, 2, 2, 2, true, 512>(cfg, sc, nframes, st);
      return launch_cost<RR, 64, 1, 2, 5, 2, true, 512>(cfg, sc, nframes, st);
    };
    if (cn != 1 || cfg.r > 7) return cudaErrorInvalidConfiguration;
    return cfg.r == 6 ? go(std::integral_constant<int, 6>()) : go(std::integral_constant<int, 7>());
  }
  if (cn == 3) {
    // three channels: six operand planes, 16-bit pixel costs in the row ring.  With the ring in registers the kernel spills at
    // 128 registers; for blockSize <= 7 and D >= 128 the ring fits into shared memory beside a two-row horizontal block
    if (cfg.r <= 3 && cfg.D >= 128) {
      auto go = [&](auto rtag) {
        constexpr int RR = decltype(rtag)::value;
        if (cfg.D > 128) return launch_cost<RR, 32, 1, 2, 6, 6, true, 512>(cfg, sc, nframes, st);
        return launch_cost<RR, 64, 1, 2, 3, 6, true, 512>(cfg, sc, nframes, st);
      };
      switch (cfg.r) {
        case 0: return go(std::integral_constant<int, 0>());
        case 1: return go(std::integral_constant<int, 1>());
        case 2: return go(std::integral_constant<int, 2>());
        default: return go(std::integral_constant<int, 3>());
      }
    }
    if (cfg.D > 128) return launch_cost_r<32, 1, 8, 6, 6>(cfg, sc, nframes, st);
    if (cfg.D == 128) return launch_cost_r<64, 1, 8, 3, 6>(cfg, sc, nframes, st);
    if (cfg.D >= 48) return launch_cost_r<64, 1, 8, 6, 6>(cfg, sc, nframes, st);
    return launch_cost_r<64, 1, 8, 15, 6>(cfg, sc, nframes, st);
  }
  if (cfg.r <= 3 && cfg.D >= 64) {
    // threads = NC / 2 x D / 16 <= 512; records staged per thread and row = 2 (NC + NJ) / threads
    const int D = cfg.D;
    // TMA staging of the row tiles; at D = 64 (256-column tiles) its three-row ring only fits beside a one-row horizontal block
    constexpr bool T = SV_COST_TMA != 0;
    // (a one-row horizontal block measured slower where the two-row one fits: cost stage 15.6 vs 15.4 ms at D = 128)
    if (D > 128) return launch_cost_pair<64, 2, T>(cfg, sc, nframes, st);
    if (D >= 96) return launch_cost_pair<128, 2, T>(cfg, sc, nframes, st);
    if (D > 64) return launch_cost_pair<128, 3, T>(cfg, sc, nframes, st);
    return launch_cost_pair<256, 3, T, 1>(cfg, sc, nframes, st);   // cost stage 5.04 -> 3.86 ms per 64 pairs at 1280x720 / 64
  }
  // one column per thread: 64 columns per tile while the CTA stays within 512 threads (D <= 128), 32 columns above
  if (cfg.D > 128) return launch_cost_r<32, 1, 8, 2>(cfg, sc, nframes, st);
  if (cfg.D == 128) return launch_cost_r<64, 1, 8, 1>(cfg, sc, nframes, st);
  if (cfg.D >= 48) return launch_cost_r<64, 1, 8, 2>(cfg, sc, nframes, st);
  return launch_cost_r<64, 1, 8, 5>(cfg, sc, nframes, st);
}

}  // namespace svgpu

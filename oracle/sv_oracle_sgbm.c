/*
 * sv_oracle_sgbm.c — plain-C restatement of cv::StereoSGBM::compute as called at
 * /root/reference/src/lib.cpp:248-262 (modes MODE_SGBM, MODE_HH, MODE_HH4).
 * TEST INFRASTRUCTURE ONLY — see sv_oracle.h.  Spec: SURVEY.md Appendix A.1-A.7.
 * Deliberately naive (whole-volume arrays, one loop nest per path) so that it can
 * be read against the spec line by line.
 */
#include "sv_oracle.h"
#include <stdlib.h>
#include <string.h>

#define MAXC 32767

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* A.2: pre-filter rows. g = clipped x-Sobel, i = intensity; both = ftzero at x=0 and x=W-1 */
/* `img` points at channel c of an image with `cn` interleaved channels (calcPixelCostBT treats every channel of an
 * 8UC3 input as a plane of its own: row1[x*3 + c]) */
static void prefilter(const uint8_t* img, int W, int H, int stride, int cn, int ftzero, uint8_t* g, uint8_t* it) {
  for (int y = 0; y < H; y++) {
    const uint8_t* r = img + (size_t)y * stride;
    const uint8_t* u = y > 0 ? r - stride : r;
    const uint8_t* d = y < H - 1 ? r + stride : r;
    uint8_t* gy = g + (size_t)y * W;
    uint8_t* iy = it + (size_t)y * W;
    for (int x = 1; x < W - 1; x++) {
      const int a = (x + 1) * cn, b = (x - 1) * cn;
      int v = 2 * (r[a] - r[b]) + (u[a] - u[b]) + (d[a] - d[b]);
      gy[x] = (uint8_t)(iclamp(v, -ftzero, ftzero) + ftzero);
      iy[x] = r[x * cn];
    }
    gy[0] = iy[0] = (uint8_t)ftzero;
    if (W > 1) gy[W - 1] = iy[W - 1] = (uint8_t)ftzero;
  }
}

/* A.3: half-pixel interpolated min / max of one row */
static void lohi(const uint8_t* p, int W, uint8_t* lo, uint8_t* hi) {
  for (int x = 0; x < W; x++) {
    int a = p[x], mn = a, mx = a;
    if (x > 0) { int h = (a + p[x - 1]) >> 1; mn = imin(mn, h); mx = imax(mx, h); }
    if (x < W - 1) { int h = (a + p[x + 1]) >> 1; mn = imin(mn, h); mx = imax(mx, h); }
    lo[x] = (uint8_t)mn; hi[x] = (uint8_t)mx;
  }
}

static inline int bt(int u, int v, int loL, int hiL, int loR, int hiR) {
  int c0 = imax(0, imax(u - hiR, loR - u));
  int c1 = imax(0, imax(v - hiL, loL - v));
  return imin(c0, c1);
}

/* A.6 for one image row: winner-take-all + uniqueness + sub-pixel over S[x'][d], then the left-right check; writes the
 * row `dr` (W values, INVALID where nothing is selected).  disp2 / disp2cost are scratch rows of W entries. */
static void wta_lr_row(const int16_t* Srow, int W, int W1, int D, int minD, int minX1, int maxX1, int uniq, int d12, int INVALID,
                       int simd8, int16_t* disp2, int* disp2cost, int16_t* dr) {
  for (int x = 0; x < W; x++) { disp2[x] = (int16_t)INVALID; disp2cost[x] = MAXC; }
  for (int xp = W1 - 1; xp >= 0; xp--) {
    const int16_t* sp = Srow + (size_t)xp * D;
    int minS = MAXC, best = -1;
    for (int d = 0; d < D; d++)
      if (sp[d] < minS) { minS = sp[d]; best = d; }
    if (best < 0 && !simd8) continue; /* all saturated */
    int d;
    if (simd8) {
      /* computeDisparity3WAY selects the winner with 8-lane vector code: every lane (d mod 8) keeps the LAST disparity
       * that attains the minimum, the lanes are then reduced to the smallest such disparity; and its uniqueness test
       * compares against (short)(100*minS/(100-uniq) + 1) instead of cross-multiplying. */
      if (best < 0) { minS = MAXC; }
      best = -1;
      for (int j = 0; j < 8 && j < D; j++) {
        int last = -1;
        for (int k = j; k < D; k += 8)
          if (sp[k] == minS) last = k;
        if (last >= 0 && (best < 0 || last < best)) best = last;
      }
      if (uniq > 0) {
        const int16_t thr = (int16_t)(uint16_t)((100 * minS) / (100 - uniq) + 1);
        for (d = 0; d < D; d++)
          if (sp[d] < thr && abs(best - d) > 1) break;
        if (d < D) continue;
      }
    } else {
      for (d = 0; d < D; d++)
        if (sp[d] * (100 - uniq) < minS * 100 && abs(best - d) > 1) break;
      if (d < D) continue;
    }
    const int x2 = xp + minX1 - best - minD;
    if (disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = (int16_t)(best + minD); }
    int dfix;
    if (0 < best && best < D - 1) {
      int den = imax(sp[best - 1] + sp[best + 1] - 2 * sp[best], 1);
      dfix = best * 16 + ((sp[best - 1] - sp[best + 1]) * 16 + den) / (den * 2);
    } else
      dfix = best * 16;
    dr[xp + minX1] = (int16_t)(dfix + minD * 16);
  }
  for (int x = minX1; x < maxX1; x++) {
    const int d1 = dr[x];
    if (d1 == INVALID) continue;
    const int _d = d1 >> 4, d_ = (d1 + 15) >> 4, _x = x - _d, x_ = x - d_;
    if (0 <= _x && _x < W && disp2[_x] >= minD && abs(disp2[_x] - _d) > d12 &&
        0 <= x_ && x_ < W && disp2[x_] >= minD && abs(disp2[x_] - d_) > d12)
      dr[x] = (int16_t)INVALID;
  }
}

/* one min-plus step of a path (A.5): L(p,.) from L(q,.) of the predecessor (q == NULL: outside, L = 0) */
static void path_step(const int16_t* lq, const int16_t* c /* Ctrue */, int D, int P1, int P2, int16_t* out) {
  int m = 0;
  if (lq) { m = MAXC; for (int d = 0; d < D; d++) m = imin(m, lq[d]); }
  for (int d = 0; d < D; d++) {
    const int a = lq ? lq[d] : 0;
    const int b = (lq ? (d > 0 ? lq[d - 1] : MAXC) : 0) + P1;
    const int e = (lq ? (d < D - 1 ? lq[d + 1] : MAXC) : 0) + P1;
    const int v = c[d] + imin(imin(a, b), imin(e, m + P2)) - m;
    out[d] = (int16_t)(v > MAXC ? MAXC : v);
  }
}

/* MODE_SGBM_3WAY (OpenCV's computeDisparity3WAY; black-box restatement pinned against cv2 4.13, tests/golden/sgbm3way_small.npz):
 *  - the image is cut into 4 stripes of ceil(H/4) rows, each processed on its own from `overlap` rows above its first row
 *    (overlap = blockSize/2 + 1 + ceil(0.1 * stripe rows)); the vertical path restarts at the first processed row;
 *  - block sums take the rows above the first processed row of a stripe from that row (clamp) but real image rows below
 *    the stripe;
 *  - per row: path left-to-right and path top-down in one pass, path right-to-left in a second pass, S = sum of the three,
 *    then the usual winner-take-all / uniqueness / sub-pixel / left-right check of the row;
 *  - every stripe writes into a private buffer of (stripe rows + overlap) rows at row offset `overlap` (stripe 0) or 0,
 *    and buffer rows [overlap, ..) become the stripe's image rows: for images so small that a stripe's overlap would
 *    start above row 0 the rows are therefore shifted, exactly as OpenCV produces them. */
static void sgbm_3way(const uint16_t* hs, int W, int H, int W1, int D, int minD, int minX1, int maxX1, int SH2, int P1, int P2, int uniq,
                      int d12, int INVALID, int16_t* disp, int dstride) {
  const int nstripes = 4;
  const int sz = (H + nstripes - 1) / nstripes;
  const int ov = (SH2 + 1) + (sz + 9) / 10; /* ceil(0.1 * sz) */
  const size_t rowCells = (size_t)W1 * D;
  int16_t* Ctrue = (int16_t*)malloc(rowCells * 2);
  int16_t* Ll = (int16_t*)malloc(rowCells * 2);
  int16_t* Lv = (int16_t*)malloc(rowCells * 2);
  int16_t* Srow = (int16_t*)malloc(rowCells * 2);
  int16_t* buf = (int16_t*)malloc((size_t)(sz + ov) * W * 2);
  int16_t* disp2 = (int16_t*)malloc((size_t)W * 2);
  int* disp2cost = (int*)malloc((size_t)W * sizeof(int));
  int16_t Lr[2][1024];
  for (int s = 0; s < nstripes; s++) {
    const int start = imax(imin(s * sz - ov, H), 0), end = imin((s + 1) * sz, H);
    if (end <= start) continue;
    const int dstOff = s == 0 ? ov : 0;
    for (size_t i = 0; i < (size_t)(sz + ov) * W; i++) buf[i] = (int16_t)INVALID;
    for (int y = 0; y < end - start; y++) {
      for (int xp = 0; xp < W1; xp++) {
        int acc[1024];
        for (int d = 0; d < D; d++) acc[d] = 0;
        for (int dy = -SH2; dy <= SH2; dy++) {
          const uint16_t* hp = hs + ((size_t)iclamp(start + y + dy, start, H - 1) * W1 + xp) * D;
          for (int d = 0; d < D; d++) acc[d] += hp[d];
        }
        for (int d = 0; d < D; d++) Ctrue[(size_t)xp * D + d] = (int16_t)acc[d];
      }
      for (int xp = 0; xp < W1; xp++) {
        path_step(xp > 0 ? Ll + (size_t)(xp - 1) * D : NULL, Ctrue + (size_t)xp * D, D, P1, P2, Ll + (size_t)xp * D);
        int16_t tmp[1024];
        path_step(y > 0 ? Lv + (size_t)xp * D : NULL, Ctrue + (size_t)xp * D, D, P1, P2, tmp);
        memcpy(Lv + (size_t)xp * D, tmp, (size_t)D * 2);
      }
      for (int xp = W1 - 1; xp >= 0; xp--) {
        int16_t* cur = Lr[xp & 1];
        path_step(xp < W1 - 1 ? Lr[(xp + 1) & 1] : NULL, Ctrue + (size_t)xp * D, D, P1, P2, cur);
        for (int d = 0; d < D; d++) {
          const int v = Ll[(size_t)xp * D + d] + Lv[(size_t)xp * D + d] + cur[d];
          Srow[(size_t)xp * D + d] = (int16_t)(v > MAXC ? MAXC : v);
        }
      }
      if (dstOff + y < sz + ov)
        wta_lr_row(Srow, W, W1, D, minD, minX1, maxX1, uniq, d12, INVALID, 1, disp2, disp2cost, buf + (size_t)(dstOff + y) * W);
    }
    for (int k = 0; s * sz + k < imin((s + 1) * sz, H); k++)
      memcpy((uint8_t*)disp + (size_t)(s * sz + k) * dstride, buf + (size_t)(ov + k) * W, (size_t)W * 2);
  }
  free(Ctrue); free(Ll); free(Lv); free(Srow); free(buf); free(disp2); free(disp2cost);
}

/* test hooks: when non-NULL the oracle copies its cost volume Ctrue (= C - P2) and the aggregated S
 * ([y][x'][d], int16) here so that GPU intermediates can be diffed stage by stage */
int16_t* svo_debug_C = 0;
int16_t* svo_debug_S = 0;

int svo_sgbm_compute(const uint8_t* left, const uint8_t* right, int W, int H, int lstride, int rstride,
                     const svo_sgbm_params* prm, int16_t* disp, int dstride) {
  return svo_sgbm_compute_cn(left, right, W, H, lstride, rstride, 1, prm, disp, dstride);
}

/* cn = 1 (8UC1, what the reference passes, src/lib.cpp:262) or 3 (8UC3: StereoSGBM sums the pixel cost of the three
 * channels, each with its own gradient and intensity term) */
int svo_sgbm_compute_cn(const uint8_t* left, const uint8_t* right, int W, int H, int lstride, int rstride, int cn,
                        const svo_sgbm_params* prm, int16_t* disp, int dstride) {
  /* A.1 parameter normalisation */
  const int minD = prm->minDisparity, D = prm->numDisparities, maxD = minD + D;
  const int P1 = prm->P1 > 0 ? prm->P1 : 2;
  const int P2 = imax(prm->P2 > 0 ? prm->P2 : 5, P1 + 1);
  const int uniq = prm->uniquenessRatio >= 0 ? prm->uniquenessRatio : 10;
  const int d12 = prm->disp12MaxDiff > 0 ? prm->disp12MaxDiff : 1;
  const int ftzero = imax(prm->preFilterCap, 15) | 1;
  const int bs = prm->blockSize > 0 ? prm->blockSize : 5;
  const int SW2 = bs / 2, SH2 = bs / 2;
  const int minX1 = imax(maxD, 0), maxX1 = W + imin(minD, 0), W1 = maxX1 - minX1;
  const int INVALID = (minD - 1) * 16;
  const int mode = prm->mode;
  if (D <= 0 || D % 16 != 0 || D > 1024) return -2;
  if (cn != 1 && cn != 3) return -2;
  /* OpenCV's 3WAY uniqueness test divides by (100 - uniquenessRatio): cv2 dies with SIGFPE at 100 */
  if (mode == SVO_MODE_SGBM_3WAY && uniq >= 100) return -2;
  if (mode != SVO_MODE_SGBM && mode != SVO_MODE_HH && mode != SVO_MODE_HH4 && mode != SVO_MODE_SGBM_3WAY) return -2;
  /* cv2 raises for MODE_SGBM / MODE_HH when width - (minD + D) <= blockSize/2; MODE_HH4 does not */
  /* (MODE_SGBM_3WAY has no such check in cv2 - it fails in an allocation for a negative width and returns values that
   * are not pinned here for 0 < width - (minD + D) <= blockSize/2 - the restatement refuses those images as well) */
  if (mode != SVO_MODE_HH4 && W - (minD + D) <= bs / 2) return -2;

  for (int y = 0; y < H; y++) {
    int16_t* dr = (int16_t*)((uint8_t*)disp + (size_t)y * dstride);
    for (int x = 0; x < W; x++) dr[x] = (int16_t)INVALID;
  }
  if (W1 <= 0 || H <= 0) return 0;

  const size_t npix = (size_t)W * H;
  uint8_t* buf = (uint8_t*)malloc(npix * 12);
  uint8_t *gL = buf, *iL = buf + npix, *gR = buf + 2 * npix, *iR = buf + 3 * npix;
  uint8_t *lo[4], *hi[4];
  for (int k = 0; k < 4; k++) { lo[k] = buf + (4 + k) * npix; hi[k] = buf + (8 + k) * npix; }

  const size_t ncell = (size_t)W1 * H * D;
  uint16_t* pix = (uint16_t*)calloc(ncell, 2);
  int16_t* C = (int16_t*)malloc(ncell * 2);
  int16_t* S = (int16_t*)calloc(ncell, 2);
  for (int ch = 0; ch < cn; ch++) {
    prefilter(left + ch, W, H, lstride, cn, ftzero, gL, iL);
    prefilter(right + ch, W, H, rstride, cn, ftzero, gR, iR);
    for (int y = 0; y < H; y++) {
      size_t o = (size_t)y * W;
      lohi(gL + o, W, lo[0] + o, hi[0] + o); lohi(iL + o, W, lo[1] + o, hi[1] + o);
      lohi(gR + o, W, lo[2] + o, hi[2] + o); lohi(iR + o, W, lo[3] + o, hi[3] + o);
    }
    /* A.3 pixel cost, accumulated over the channels */
    for (int y = 0; y < H; y++)
      for (int xp = 0; xp < W1; xp++) {
        const int x = xp + minX1;
        const size_t o = (size_t)y * W;
        uint16_t* pp = pix + ((size_t)y * W1 + xp) * D;
        for (int d = 0; d < D; d++) {
          const int xr = x - (d + minD);
          int cg = bt(gL[o + x], gR[o + xr], lo[0][o + x], hi[0][o + x], lo[2][o + xr], hi[2][o + xr]);
          int ci = bt(iL[o + x], iR[o + xr], lo[1][o + x], hi[1][o + x], lo[3][o + xr], hi[3][o + xr]);
          pp[d] = (uint16_t)(pp[d] + cg + (ci >> 2));
        }
      }
  }
  /* A.4 block sum, clamped in W1 coordinates; stored C = Ctrue + P2 (int16) */
  /* (separable: horizontal window sums first, then vertical) */
  uint16_t* hs = (uint16_t*)malloc(ncell * 2);
  for (int y = 0; y < H; y++)
    for (int xp = 0; xp < W1; xp++) {
      uint16_t* hp = hs + ((size_t)y * W1 + xp) * D;
      for (int d = 0; d < D; d++) hp[d] = 0;
      for (int dx = -SW2; dx <= SW2; dx++) {
        const uint16_t* pp = pix + ((size_t)y * W1 + iclamp(xp + dx, 0, W1 - 1)) * D;
        for (int d = 0; d < D; d++) hp[d] = (uint16_t)(hp[d] + pp[d]);
      }
    }
  if (mode == SVO_MODE_SGBM_3WAY) {
    sgbm_3way(hs, W, H, W1, D, minD, minX1, maxX1, SH2, P1, P2, uniq, d12, INVALID, disp, dstride);
    free(hs); free(pix); free(buf); free(C); free(S);
    goto median;
  }
  for (int y = 0; y < H; y++)
    for (int xp = 0; xp < W1; xp++) {
      int16_t* cp = C + ((size_t)y * W1 + xp) * D;
      if (mode == SVO_MODE_HH4 && y >= H - SH2) { /* HH4 bottom-rows quirk */
        for (int d = 0; d < D; d++) cp[d] = (int16_t)P2;
        continue;
      }
      int acc[1024];
      for (int d = 0; d < D; d++) acc[d] = P2;
      for (int dy = -SH2; dy <= SH2; dy++) {
        const uint16_t* hp = hs + ((size_t)iclamp(y + dy, 0, H - 1) * W1 + xp) * D;
        for (int d = 0; d < D; d++) acc[d] += hp[d];
      }
      for (int d = 0; d < D; d++) cp[d] = (int16_t)acc[d];
    }
  free(hs);
  free(pix);
  free(buf);

  /* A.5 path aggregation: one full pass per path, predecessor offset (dx,dy) */
  static const int paths_sgbm[5][2] = {{-1, 0}, {-1, -1}, {0, -1}, {1, -1}, {1, 0}};
  static const int paths_hh[8][2] = {{-1, 0}, {-1, -1}, {0, -1}, {1, -1}, {1, 0}, {1, 1}, {0, 1}, {-1, 1}};
  static const int paths_hh4[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
  const int(*paths)[2] = mode == SVO_MODE_SGBM ? paths_sgbm : (mode == SVO_MODE_HH ? paths_hh : paths_hh4);
  const int npaths = mode == SVO_MODE_SGBM ? 5 : (mode == SVO_MODE_HH ? 8 : 4);
  int16_t* L = (int16_t*)malloc((size_t)2 * W1 * D * 2); /* two rows of L_r */
  for (int r = 0; r < npaths; r++) {
    const int dx = paths[r][0], dy = paths[r][1];
    const int y0 = dy <= 0 ? 0 : H - 1, ys = dy <= 0 ? 1 : -1;
    const int x0 = dx <= 0 ? 0 : W1 - 1, xs = dx <= 0 ? 1 : -1;
    for (int yi = 0, y = y0; yi < H; yi++, y += ys) {
      int16_t* Lcur = L + (size_t)(y & 1) * W1 * D;
      int16_t* Lprev = L + (size_t)((y + dy) & 1) * W1 * D; /* == Lcur when dy==0 */
      for (int xi = 0, xp = x0; xi < W1; xi++, xp += xs) {
        const int qx = xp + dx, qy = y + dy;
        const int16_t* cp = C + ((size_t)y * W1 + xp) * D;
        int16_t* sp = S + ((size_t)y * W1 + xp) * D;
        int16_t* lp = Lcur + (size_t)xp * D;
        int16_t tmp[1024];
        if (qx < 0 || qx >= W1 || qy < 0 || qy >= H) {
          /* L(q,.)=0, m=0: delta = P2 */
          for (int d = 0; d < D; d++) tmp[d] = (int16_t)(cp[d] + imin(0, imin(P1, P2)) - P2);
        } else {
          const int16_t* lq = Lprev + (size_t)qx * D;
          int m = MAXC;
          for (int d = 0; d < D; d++) m = imin(m, lq[d]);
          const int delta = m + P2;
          for (int d = 0; d < D; d++) {
            int a = lq[d];
            int b = (d > 0 ? lq[d - 1] : MAXC) + P1;
            int c = (d < D - 1 ? lq[d + 1] : MAXC) + P1;
            tmp[d] = (int16_t)(cp[d] + imin(imin(a, b), imin(c, delta)) - delta);
          }
        }
        for (int d = 0; d < D; d++) {
          lp[d] = tmp[d];
          int s = sp[d] + tmp[d];
          sp[d] = (int16_t)(s > MAXC ? MAXC : s);
        }
      }
    }
  }
  free(L);
  if (svo_debug_C) for (size_t i = 0; i < ncell; i++) svo_debug_C[i] = (int16_t)(C[i] - P2);
  if (svo_debug_S) memcpy(svo_debug_S, S, ncell * 2);
  free(C);

  /* A.6 WTA + uniqueness + sub-pixel + LR check */
  int16_t* disp2 = (int16_t*)malloc((size_t)W * 2);
  int* disp2cost = (int*)malloc((size_t)W * sizeof(int));
  for (int y = 0; y < H; y++) {
    int16_t* dr = (int16_t*)((uint8_t*)disp + (size_t)y * dstride);
    for (int x = 0; x < W; x++) { disp2[x] = (int16_t)INVALID; disp2cost[x] = MAXC; }
    for (int xp = W1 - 1; xp >= 0; xp--) {
      const int16_t* sp = S + ((size_t)y * W1 + xp) * D;
      int minS = MAXC, best = -1;
      for (int d = 0; d < D; d++)
        if (sp[d] < minS) { minS = sp[d]; best = d; }
      if (best < 0) continue; /* all saturated */
      int d;
      for (d = 0; d < D; d++)
        if (sp[d] * (100 - uniq) < minS * 100 && abs(best - d) > 1) break;
      if (d < D) continue;
      const int x2 = xp + minX1 - best - minD;
      if (disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = (int16_t)(best + minD); }
      int dfix;
      if (0 < best && best < D - 1) {
        int den = imax(sp[best - 1] + sp[best + 1] - 2 * sp[best], 1);
        dfix = best * 16 + ((sp[best - 1] - sp[best + 1]) * 16 + den) / (den * 2);
      } else
        dfix = best * 16;
      dr[xp + minX1] = (int16_t)(dfix + minD * 16);
    }
    for (int x = minX1; x < maxX1; x++) {
      const int d1 = dr[x];
      if (d1 == INVALID) continue;
      const int _d = d1 >> 4, d_ = (d1 + 15) >> 4, _x = x - _d, x_ = x - d_;
      if (0 <= _x && _x < W && disp2[_x] >= minD && abs(disp2[_x] - _d) > d12 &&
          0 <= x_ && x_ < W && disp2[x_] >= minD && abs(disp2[x_] - d_) > d12)
        dr[x] = (int16_t)INVALID;
    }
  }
  free(disp2);
  free(disp2cost);
  free(S);

median:;
  /* A.7 median 3x3 in place, then the built-in speckle stage */
  int16_t* tmpi = (int16_t*)malloc(npix * 2);
  int16_t* tmpo = (int16_t*)malloc(npix * 2);
  for (int y = 0; y < H; y++) memcpy(tmpi + (size_t)y * W, (uint8_t*)disp + (size_t)y * dstride, (size_t)W * 2);
  svo_median3_s16(tmpi, tmpo, W, H);
  if (prm->speckleWindowSize > 0)
    svo_filter_speckles(tmpo, W, H, INVALID, prm->speckleWindowSize, 16 * prm->speckleRange);
  for (int y = 0; y < H; y++) memcpy((uint8_t*)disp + (size_t)y * dstride, tmpo + (size_t)y * W, (size_t)W * 2);
  free(tmpi);
  free(tmpo);
  return 0;
}

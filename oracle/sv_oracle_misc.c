/*
 * sv_oracle_misc.c — plain-C restatements of the remaining OpenCV calls on the
 * dense-disparity path of /root/reference/src/lib.cpp (file:line per function).
 * TEST INFRASTRUCTURE ONLY — see sv_oracle.h.  Spec: SURVEY.md Appendix A.7-A.13.
 */
#include "sv_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

/* A.7 — cv::medianBlur(CV_16S, ksize 3) as run at the end of StereoSGBM::compute (src/lib.cpp:262) */
void svo_median3_s16(const int16_t* src, int16_t* dst, int W, int H) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      int16_t v[9];
      int n = 0;
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++)
          v[n++] = src[(size_t)iclamp(y + dy, 0, H - 1) * W + iclamp(x + dx, 0, W - 1)];
      for (int i = 1; i < 9; i++) { /* insertion sort */
        int16_t t = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > t) { v[j + 1] = v[j]; j--; }
        v[j + 1] = t;
      }
      dst[(size_t)y * W + x] = v[4];
    }
}

/* A.8 — cv::filterSpeckles(img, newVal, maxSpeckleSize, maxDiff) (src/lib.cpp:265) */
void svo_filter_speckles(int16_t* img, int W, int H, int newVal, int maxSpeckleSize, int maxDiff) {
  const size_t n = (size_t)W * H;
  uint8_t* seen = (uint8_t*)calloc(n, 1);
  int* stack = (int*)malloc(n * sizeof(int));
  int* region = (int*)malloc(n * sizeof(int));
  for (size_t s = 0; s < n; s++) {
    if (seen[s] || img[s] == newVal) continue;
    int top = 0, cnt = 0;
    stack[top++] = (int)s;
    seen[s] = 1;
    while (top) {
      const int p = stack[--top];
      region[cnt++] = p;
      const int py = p / W, px = p % W, v = img[p];
      const int nb[4] = {px > 0 ? p - 1 : -1, px < W - 1 ? p + 1 : -1, py > 0 ? p - W : -1, py < H - 1 ? p + W : -1};
      for (int k = 0; k < 4; k++) {
        const int q = nb[k];
        if (q < 0 || seen[q] || img[q] == newVal || abs(v - img[q]) > maxDiff) continue;
        seen[q] = 1;
        stack[top++] = q;
      }
    }
    if (cnt <= maxSpeckleSize)
      for (int i = 0; i < cnt; i++) img[region[i]] = (int16_t)newVal;
  }
  free(seen);
  free(stack);
  free(region);
}

static void inv3(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, B = -(d * i - f * g), Cc = d * h - e * g;
  const double det = a * A + b * B + c * Cc, id = 1.0 / det;
  o[0] = A * id; o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
  o[3] = B * id; o[4] = (a * i - c * g) * id; o[5] = -(a * f - c * d) * id;
  o[6] = Cc * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}

/* A.10 — cv::initUndistortRectifyMap(K, dist, R, P, size, CV_16SC2) (src/lib.cpp:39-46) */
void svo_init_undistort_rectify_map(const double* K, const double* dist, const double* R, const double* P,
                                    int W, int H, int16_t* map1, uint16_t* map2) {
  double A[9], PR[9], iR[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) A[r * 3 + c] = P[r * 4 + c];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += A[r * 3 + k] * R[k * 3 + c];
      PR[r * 3 + c] = s;
    }
  inv3(PR, iR);
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
  for (int v = 0; v < H; v++) {
    double X = v * iR[1] + iR[2], Y = v * iR[4] + iR[5], Wc = v * iR[7] + iR[8];
    for (int u = 0; u < W; u++, X += iR[0], Y += iR[3], Wc += iR[6]) {
      const double w = 1.0 / Wc, x = X * w, y = Y * w;
      const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
      const double kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2;
      const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
      const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
      const double us = fx * xd + cx, vs = fy * yd + cy;
      const int iu = (int)lrint(us * 32), iv = (int)lrint(vs * 32); /* round-half-even */
      map1[((size_t)v * W + u) * 2 + 0] = (int16_t)iclamp(iu >> 5, -32768, 32767);
      map1[((size_t)v * W + u) * 2 + 1] = (int16_t)iclamp(iv >> 5, -32768, 32767);
      map2[(size_t)v * W + u] = (uint16_t)((iv & 31) * 32 + (iu & 31));
    }
  }
}

/* A.11 — cv::remap(src, map1(16SC2), map2(16UC1), INTER_LINEAR), BORDER_CONSTANT 0 (src/lib.cpp:89-92) */
void svo_remap_u8(const uint8_t* src, int sw, int sh, int sstride, int cn, const int16_t* map1,
                  const uint16_t* map2, uint8_t* dst, int W, int H, int dstride) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const int x0 = map1[((size_t)y * W + x) * 2], y0 = map1[((size_t)y * W + x) * 2 + 1];
      const int f = map2[(size_t)y * W + x] & 1023, fx = f & 31, fy = f >> 5;
      /* 15-bit weights: (32-fy)(32-fx)*32 etc. are exact, sum is always 32768 */
      const int w[4] = {(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32};
      for (int c = 0; c < cn; c++) {
        int acc = 0;
        for (int k = 0; k < 4; k++) {
          const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
          const int v = (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstride + xx * cn + c] : 0;
          acc += w[k] * v;
        }
        dst[(size_t)y * dstride + x * cn + c] = (uint8_t)iclamp((acc + 16384) >> 15, 0, 255);
      }
    }
}

/* A.12 — cv::reprojectImageTo3D(disp, Q, handleMissingValues) with CV_16S input (src/lib.cpp:154) */
void svo_reproject_s16(const int16_t* disp, int W, int H, const double* Q, int handleMissing, float* xyz) {
  double minD = 0;
  if (handleMissing) {
    int mn = 32767;
    for (size_t i = 0; i < (size_t)W * H; i++) mn = imin(mn, disp[i]);
    minD = mn;
  }
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const double d = (double)(float)disp[(size_t)y * W + x];
      double h[4];
      for (int r = 0; r < 4; r++) h[r] = Q[r * 4 + 0] * x + Q[r * 4 + 1] * y + Q[r * 4 + 2] * d + Q[r * 4 + 3];
      float* o = xyz + ((size_t)y * W + x) * 3;
      /* Vec3f /= double multiplies by the reciprocal (cv::Matx operator/=): pinned at full 4K size against cv2, where the
       * difference to a true division shows in about 1 of 65,000 values */
      const double iw = 1.0 / h[3];
      for (int r = 0; r < 3; r++) o[r] = (float)((double)(float)h[r] * iw);
      if (handleMissing && fabs(d - minD) <= FLT_EPSILON) o[2] = 10000.f;
    }
}

/* StereoVision::CalculateDepthMapSimple (src/lib.cpp:279-294) */
void svo_depth_simple(const int16_t* disp, int W, int H, float baseline, float focal, float* depth) {
  for (size_t i = 0; i < (size_t)W * H; i++) {
    const int16_t d = disp[i];
    depth[i] = d > 0 ? (baseline * focal) / (float)d : 0.0f;
  }
}

/* A.13 — cv::GaussianBlur(8U, 11x11, sigma 1.4) (src/lib.cpp:231-233): 8.8 fixed point, reflect-101 */
void svo_gaussian11_u8(const uint8_t* src, uint8_t* dst, int W, int H, int cn) {
  static const int k[11] = {0, 1, 8, 26, 57, 72, 57, 26, 8, 1, 0};
  uint16_t* h = (uint16_t*)malloc((size_t)W * H * cn * 2);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int c = 0; c < cn; c++) {
        int s = 0;
        for (int t = -5; t <= 5; t++) s += k[t + 5] * src[((size_t)y * W + reflect101(x + t, W)) * cn + c];
        h[((size_t)y * W + x) * cn + c] = (uint16_t)s;
      }
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int c = 0; c < cn; c++) {
        unsigned s = 0;
        for (int t = -5; t <= 5; t++) s += (unsigned)k[t + 5] * h[((size_t)reflect101(y + t, H) * W + x) * cn + c];
        dst[((size_t)y * W + x) * cn + c] = (uint8_t)((s + 32768u) >> 16);
      }
  free(h);
}

/* cv::cvtColor(BGR2GRAY) (src/lib.cpp:236-237): 15-bit coefficients */
void svo_bgr2gray(const uint8_t* bgr, uint8_t* gray, int W, int H) {
  for (size_t i = 0; i < (size_t)W * H; i++)
    gray[i] = (uint8_t)((bgr[i * 3] * 3735 + bgr[i * 3 + 1] * 19235 + bgr[i * 3 + 2] * 9798 + 16384) >> 15);
}

/* cv::filter2D 3x3 sharpen [0 -1 0; -1 5 -1; 0 -1 0] (src/lib.cpp:244-246), reflect-101, saturate */
void svo_sharpen3_u8(const uint8_t* src, uint8_t* dst, int W, int H) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const int c = src[(size_t)y * W + x];
      const int l = src[(size_t)y * W + reflect101(x - 1, W)], r = src[(size_t)y * W + reflect101(x + 1, W)];
      const int u = src[(size_t)reflect101(y - 1, H) * W + x], d = src[(size_t)reflect101(y + 1, H) * W + x];
      dst[(size_t)y * W + x] = (uint8_t)iclamp(5 * c - l - r - u - d, 0, 255);
    }
}

/* cv::morphologyEx(MORPH_CLOSE, 5x5 rect) on CV_16S (src/lib.cpp:267-268): dilate, then erode; border ignored */
void svo_morph_close5_s16(const int16_t* src, int16_t* dst, int W, int H) {
  int16_t* t = (int16_t*)malloc((size_t)W * H * 2);
  for (int pass = 0; pass < 2; pass++) {
    const int16_t* in = pass == 0 ? src : t;
    int16_t* out = pass == 0 ? t : dst;
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        int v = pass == 0 ? -32768 : 32767;
        for (int dy = -2; dy <= 2; dy++)
          for (int dx = -2; dx <= 2; dx++) {
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const int s = in[(size_t)yy * W + xx];
            v = pass == 0 ? imax(v, s) : imin(v, s);
          }
        out[(size_t)y * W + x] = (int16_t)v;
      }
  }
  free(t);
}

/* A.13 — cv::resize(8U, INTER_LINEAR) (src/lib.cpp:102): 11-bit fixed-point two-tap bilinear */
void svo_resize_linear_u8(const uint8_t* src, int sw, int sh, int cn, uint8_t* dst, int dw, int dh) {
  int* xi = (int*)malloc((size_t)dw * sizeof(int));
  int* xa = (int*)malloc((size_t)dw * 2 * sizeof(int));
  const double sx = (double)sw / dw, sy = (double)sh / dh;
  for (int x = 0; x < dw; x++) {
    float fx = (float)((x + 0.5) * sx - 0.5);
    int i0 = (int)floorf(fx);
    fx -= i0;
    if (i0 < 0) { i0 = 0; fx = 0; }
    if (i0 >= sw - 1) { i0 = sw - 1; fx = 0; }
    xi[x] = i0;
    xa[2 * x] = (int)lrintf((1.f - fx) * 2048);
    xa[2 * x + 1] = (int)lrintf(fx * 2048);
  }
  int* r0 = (int*)malloc((size_t)dw * cn * sizeof(int));
  int* r1 = (int*)malloc((size_t)dw * cn * sizeof(int));
  for (int y = 0; y < dh; y++) {
    float fy = (float)((y + 0.5) * sy - 0.5);
    int j0 = (int)floorf(fy);
    fy -= j0;
    if (j0 < 0) { j0 = 0; fy = 0; }
    if (j0 >= sh - 1) { j0 = sh - 1; fy = 0; }
    const int j1 = imin(j0 + 1, sh - 1);
    const int b0 = (int)lrintf((1.f - fy) * 2048), b1 = (int)lrintf(fy * 2048);
    for (int x = 0; x < dw; x++) {
      const int i0 = xi[x], i1 = imin(i0 + 1, sw - 1);
      for (int c = 0; c < cn; c++) {
        r0[x * cn + c] = src[((size_t)j0 * sw + i0) * cn + c] * xa[2 * x] + src[((size_t)j0 * sw + i1) * cn + c] * xa[2 * x + 1];
        r1[x * cn + c] = src[((size_t)j1 * sw + i0) * cn + c] * xa[2 * x] + src[((size_t)j1 * sw + i1) * cn + c] * xa[2 * x + 1];
      }
    }
    for (int i = 0; i < dw * cn; i++)
      dst[(size_t)y * dw * cn + i] = (uint8_t)((((b0 * (r0[i] >> 4)) >> 16) + ((b1 * (r1[i] >> 4)) >> 16) + 2) >> 2);
  }
  free(xi); free(xa); free(r0); free(r1);
}

/* A.9 — cv::StereoBM::compute (BASELINE config 2; not called by the reference), minDisparity 0 contract */
static void bm_prefilter_xsobel(const uint8_t* img, int W, int H, int stride, int cap, uint8_t* out) {
  const int He = H & ~1; /* rows are produced in pairs; an unpaired last row is filled with cap */
  for (int y = 0; y < He; y++) {
    const uint8_t* r = img + (size_t)y * stride;
    const uint8_t* u = img + (size_t)(y > 0 ? y - 1 : (H > 1 ? 1 : 0)) * stride;
    const uint8_t* d = img + (size_t)(y < H - 1 ? y + 1 : (H > 1 ? H - 2 : 0)) * stride;
    uint8_t* o = out + (size_t)y * W;
    for (int x = 1; x < W - 1; x++) {
      const int v = (u[x + 1] - u[x - 1]) + 2 * (r[x + 1] - r[x - 1]) + (d[x + 1] - d[x - 1]);
      o[x] = (uint8_t)(iclamp(v, -cap, cap) + cap);
    }
    o[0] = (uint8_t)cap;
    if (W > 1) o[W - 1] = (uint8_t)cap;
  }
  for (int y = He; y < H; y++) memset(out + (size_t)y * W, cap, (size_t)W);
}

int svo_bm_compute(const uint8_t* left, const uint8_t* right, int W, int H, int lstride, int rstride,
                   const svo_bm_params* p, int16_t* disp, int dstride) {
  const int minD = p->minDisparity, D = p->numDisparities, bs = p->blockSize, cap = p->preFilterCap;
  const int wsz2 = bs / 2, FILTERED = (minD - 1) * 16;
  if (D <= 0 || D % 16 || bs < 5 || bs % 2 == 0 || cap < 1 || cap > 63 || D > 1024) return -2;
  uint8_t* PL = (uint8_t*)malloc((size_t)W * H);
  uint8_t* PR = (uint8_t*)malloc((size_t)W * H);
  bm_prefilter_xsobel(left, W, H, lstride, cap, PL);
  bm_prefilter_xsobel(right, W, H, rstride, cap, PR);
  for (int y = 0; y < H; y++) {
    int16_t* dr = (int16_t*)((uint8_t*)disp + (size_t)y * dstride);
    for (int x = 0; x < W; x++) dr[x] = (int16_t)FILTERED;
  }
  /* Computed columns: every window pixel has all D right-image partners inside the image.  This is
   * cv::getValidDisparityROI's x range; for minDisparity != 0 OpenCV itself also emits values outside it (it indexes
   * past the end of the image row there, so those depend on the memory layout) - here they are FILTERED. */
  const int xlo = (minD + D - 1 > 0 ? minD + D - 1 : 0) + wsz2, xhi = W - wsz2 + (minD < 0 ? minD : 0);
  const int ylo = wsz2, yhi = H - wsz2;
  if (xlo < xhi && ylo < yhi) {
    /* SAD volume by one integral image per disparity: sadv[(y*W+x)*D + k], k ascending == disparity descending */
    int* sadv = (int*)malloc((size_t)W * H * D * sizeof(int));
    int* integ = (int*)malloc((size_t)(W + 1) * (H + 1) * sizeof(int));
    int* tinteg = (int*)malloc((size_t)(W + 1) * (H + 1) * sizeof(int));
    for (int k = -1; k < D; k++) { /* k == -1 builds the texture integral |PL - cap| */
      const int d = D - 1 - k + minD;
      for (int x = 0; x <= W; x++) integ[x] = 0;
      for (int y = 0; y < H; y++) {
        int run = 0;
        integ[(size_t)(y + 1) * (W + 1)] = 0;
        for (int x = 0; x < W; x++) {
          int a;
          if (k < 0) a = abs(PL[(size_t)y * W + x] - cap);
          else a = (x - d >= 0 && x - d < W) ? abs(PL[(size_t)y * W + x] - PR[(size_t)y * W + x - d]) : 0;
          run += a;
          integ[(size_t)(y + 1) * (W + 1) + x + 1] = integ[(size_t)y * (W + 1) + x + 1] + run;
        }
      }
      if (k < 0) { memcpy(tinteg, integ, (size_t)(W + 1) * (H + 1) * sizeof(int)); continue; }
      for (int y = ylo; y < yhi; y++)
        for (int x = xlo; x < xhi; x++) {
          const int x0 = x - wsz2, x1 = x + wsz2 + 1, y0 = y - wsz2, y1 = y + wsz2 + 1;
          sadv[((size_t)y * W + x) * D + k] = integ[(size_t)y1 * (W + 1) + x1] - integ[(size_t)y0 * (W + 1) + x1] -
                                              integ[(size_t)y1 * (W + 1) + x0] + integ[(size_t)y0 * (W + 1) + x0];
        }
    }
    int sadbuf[1024 + 2];
    int* sad = sadbuf + 1;
    for (int y = ylo; y < yhi; y++) {
      int16_t* dr = (int16_t*)((uint8_t*)disp + (size_t)y * dstride);
      for (int x = xlo; x < xhi; x++) {
        const int x0 = x - wsz2, x1 = x + wsz2 + 1, y0 = y - wsz2, y1 = y + wsz2 + 1;
        const int tsum = tinteg[(size_t)y1 * (W + 1) + x1] - tinteg[(size_t)y0 * (W + 1) + x1] -
                         tinteg[(size_t)y1 * (W + 1) + x0] + tinteg[(size_t)y0 * (W + 1) + x0];
        int minsad = 0x7fffffff, mind = -1;
        for (int k = 0; k < D; k++) { /* first strict minimum scanning disparities downwards */
          const int sv = sadv[((size_t)y * W + x) * D + k];
          sad[k] = sv;
          if (sv < minsad) { minsad = sv; mind = k; }
        }
        if (tsum < p->textureThreshold) continue;
        if (p->uniquenessRatio > 0) {
          const int thresh = minsad + (minsad * p->uniquenessRatio / 100);
          int k;
          for (k = 0; k < D; k++)
            if ((k < mind - 1 || k > mind + 1) && sad[k] <= thresh) break;
          if (k < D) continue;
        }
        sad[-1] = sad[1];
        sad[D] = sad[D - 2];
        const int pp = sad[mind + 1], n = sad[mind - 1];
        const int dd = pp + n - 2 * sad[mind] + abs(pp - n);
        dr[x] = (int16_t)(((D - mind - 1 + minD) * 256 + (dd != 0 ? (pp - n) * 256 / dd : 0) + 15) >> 4);
      }
    }
    free(sadv); free(integ); free(tinteg);
  }
  free(PL);
  free(PR);
  if (p->speckleWindowSize > 0) {
    int16_t* t = (int16_t*)malloc((size_t)W * H * 2);
    for (int y = 0; y < H; y++) memcpy(t + (size_t)y * W, (uint8_t*)disp + (size_t)y * dstride, (size_t)W * 2);
    svo_filter_speckles(t, W, H, FILTERED, p->speckleWindowSize, p->speckleRange);
    for (int y = 0; y < H; y++) memcpy((uint8_t*)disp + (size_t)y * dstride, t + (size_t)y * W, (size_t)W * 2);
    free(t);
  }
  return 0;
}

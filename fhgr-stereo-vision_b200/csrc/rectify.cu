// rectify.cu — host-side one-time calibration products of StereoVision::StereoVision
// (/root/reference/src/lib.cpp:29-47): the rectifying rotations / projections / Q of
// cv::stereoRectify(K1, D1, K2, D2, size, R, T, CALIB_ZERO_DISPARITY, alpha = 0) and the operands of
// cv::initUndistortRectifyMap, whose per-pixel part runs on the GPU (image_ops.cu: k_init_rectify_map).
//
// Plain f64 host arithmetic (a few hundred operations, once per camera pair).  The algorithm is Bouguet's
// rectification as OpenCV 4.13 performs it: split the relative rotation in half, align the baseline with the
// image axis, take the mean focal length, centre the principal point on the projected image corners (those four
// points go through float32 storage, as in OpenCV), then scale the focal length so that only valid pixels remain
// (alpha = 0: inner rectangle of a 9 x 9 grid of undistorted points).  Pinned against cv2.stereoRectify in
// tests/golden/rectify_small.npz (relative difference <= 1e-12, fixed-point maps identical).
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace svgpu {

namespace {

struct V3 { double x, y, z; };

inline void mat3_mul(const double* a, const double* b, double* o) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a[r * 3 + k] * b[k * 3 + c];
      o[r * 3 + c] = s;
    }
}
inline V3 mat3_vec(const double* m, V3 v) {
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z};
}
inline void mat3_t(const double* m, double* o) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[r * 3 + c] = m[c * 3 + r];
}
inline double norm3(V3 v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }

bool mat3_inv(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  if (det == 0.0 || !std::isfinite(det)) return false;
  const double id = 1.0 / det;
  o[0] = A * id; o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
  o[3] = B * id; o[4] = (a * i - c * g) * id; o[5] = -(a * f - c * d) * id;
  o[6] = C * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
  return true;
}

// rotation vector -> matrix (cv::Rodrigues)
void rodrigues_to_mat(V3 om, double* R) {
  const double th = norm3(om);
  if (th < DBL_EPSILON) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double c = std::cos(th), s = std::sin(th), c1 = 1.0 - c;
  const double r[3] = {om.x / th, om.y / th, om.z / th};
  const double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i * 3 + j] = c * (i == j ? 1.0 : 0.0) + c1 * r[i] * r[j] + s * rx[i * 3 + j];
}

// matrix -> rotation vector (cv::Rodrigues).  OpenCV first replaces R by the nearest rotation (U * Vt of its SVD);
// the same orthogonal polar factor is reached here by Newton's iteration R <- (R + R^-T) / 2.
bool rodrigues_to_vec(const double* Rin, V3* out) {
  double R[9];
  for (int i = 0; i < 9; i++) R[i] = Rin[i];
  for (int it = 0; it < 32; it++) {
    double inv[9], invT[9], nxt[9], delta = 0;
    if (!mat3_inv(R, inv)) return false;
    mat3_t(inv, invT);
    for (int i = 0; i < 9; i++) { nxt[i] = 0.5 * (R[i] + invT[i]); delta = std::fmax(delta, std::fabs(nxt[i] - R[i])); R[i] = nxt[i]; }
    if (delta < 1e-16) break;
  }
  const V3 r = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  const double s = std::sqrt((r.x * r.x + r.y * r.y + r.z * r.z) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double th = std::acos(c);
  if (s < 1e-5) {
    if (c > 0) { *out = {0, 0, 0}; return true; }
    double t = (R[0] + 1) * 0.5;
    double x = std::sqrt(std::fmax(t, 0.0));
    t = (R[4] + 1) * 0.5;
    double y = std::sqrt(std::fmax(t, 0.0)) * (R[1] < 0 ? -1.0 : 1.0);
    t = (R[8] + 1) * 0.5;
    double z = std::sqrt(std::fmax(t, 0.0)) * (R[2] < 0 ? -1.0 : 1.0);
    if (std::fabs(x) < std::fabs(y) && std::fabs(x) < std::fabs(z) && (R[5] > 0) != (y * z > 0)) z = -z;
    const double k = th / std::sqrt(x * x + y * y + z * z);
    *out = {x * k, y * k, z * k};
    return true;
  }
  const double vth = 1.0 / (2.0 * s) * th;
  *out = {r.x * vth, r.y * vth, r.z * vth};
  return true;
}

// cv::undistortPoints for one pixel: 5 fixed-point iterations of the inverse distortion model, then optionally the
// rectifying rotation R and the new camera matrix (fc, cx, cy)
void undistort_point(double u, double v, const double* K, const double* D, const double* R, const double* Pn, double* ox, double* oy) {
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
  double x = (u - cx) / fx, y = (v - cy) / fy;
  const double x0 = x, y0 = y;
  for (int it = 0; it < 5; it++) {
    const double r2 = x * x + y * y;
    const double icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
    const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
    const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
    x = (x0 - dx) * icdist;
    y = (y0 - dy) * icdist;
  }
  if (R) {
    const double xx = R[0] * x + R[1] * y + R[2], yy = R[3] * x + R[4] * y + R[5], ww = 1.0 / (R[6] * x + R[7] * y + R[8]);
    x = xx * ww;
    y = yy * ww;
  }
  if (Pn) { x = x * Pn[0] + Pn[1]; y = y * Pn[0] + Pn[2]; }
  *ox = x;
  *oy = y;
}

struct Rect { double x, y, w, h; };
// inner / outer bounding rectangles of the undistorted, rectified image (OpenCV: icvGetRectangles, 9 x 9 grid)
void get_rectangles(const double* K, const double* D, const double* R, double fc, double ccx, double ccy, int W, int H, Rect* inner, Rect* outer) {
  constexpr int N = 9;
  const double Pn[3] = {fc, ccx, ccy};
  double iX0 = -DBL_MAX, iX1 = DBL_MAX, iY0 = -DBL_MAX, iY1 = DBL_MAX;
  double oX0 = DBL_MAX, oX1 = -DBL_MAX, oY0 = DBL_MAX, oY1 = -DBL_MAX;
  for (int y = 0; y < N; y++)
    for (int x = 0; x < N; x++) {
      double px, py;
      undistort_point((double)x * (W - 1) / (N - 1), (double)y * (H - 1) / (N - 1), K, D, R, Pn, &px, &py);
      oX0 = std::fmin(oX0, px); oX1 = std::fmax(oX1, px); oY0 = std::fmin(oY0, py); oY1 = std::fmax(oY1, py);
      if (x == 0) iX0 = std::fmax(iX0, px);
      if (x == N - 1) iX1 = std::fmin(iX1, px);
      if (y == 0) iY0 = std::fmax(iY0, py);
      if (y == N - 1) iY1 = std::fmin(iY1, py);
    }
  *inner = {iX0, iY0, iX1 - iX0, iY1 - iY0};
  *outer = {oX0, oY0, oX1 - oX0, oY1 - oY0};
}

}  // namespace

// cv::stereoRectify(..., flags = CALIB_ZERO_DISPARITY, alpha = 0, newImageSize = imageSize) — src/lib.cpp:32-37
bool stereo_rectify_host(const double* K1, const double* D1, const double* K2, const double* D2, int W, int H, const double* R,
                         const double* T, StereoRectifyOut* o) {
  V3 om;
  if (!rodrigues_to_vec(R, &om)) return false;
  om = {om.x * -0.5, om.y * -0.5, om.z * -0.5};  // each camera takes half of the relative rotation
  double r_r[9], r_rT[9], wR[9];
  rodrigues_to_mat(om, r_r);
  V3 t = mat3_vec(r_r, {T[0], T[1], T[2]});
  const int idx = std::fabs(t.x) > std::fabs(t.y) ? 0 : 1;  // horizontal or vertical stereo
  const double c = idx == 0 ? t.x : t.y, nt = norm3(t);
  if (nt == 0.0) return false;
  const V3 uu = {idx == 0 ? (c > 0 ? 1.0 : -1.0) : 0.0, idx == 1 ? (c > 0 ? 1.0 : -1.0) : 0.0, 0.0};
  V3 ww = {t.y * uu.z - t.z * uu.y, t.z * uu.x - t.x * uu.z, t.x * uu.y - t.y * uu.x};  // rotation that aligns the baseline with the axis
  const double nw = norm3(ww);
  if (nw > 0.0) { const double k = std::acos(std::fabs(c) / nt) / nw; ww = {ww.x * k, ww.y * k, ww.z * k}; }
  rodrigues_to_mat(ww, wR);
  mat3_t(r_r, r_rT);
  mat3_mul(wR, r_rT, o->R1);
  mat3_mul(wR, r_r, o->R2);
  t = mat3_vec(o->R2, {T[0], T[1], T[2]});
  const double tIdx = idx == 0 ? t.x : t.y;

  const double nx = W, ny = H;
  double fc = (K1[(idx ^ 1) * 4] + K2[(idx ^ 1) * 4]) * 0.5;
  double cc[2][2];
  for (int k = 0; k < 2; k++) {
    const double* A = k == 0 ? K1 : K2;
    const double* D = k == 0 ? D1 : D2;
    const double* Rk = k == 0 ? o->R1 : o->R2;
    double ax = 0, ay = 0;
    for (int i = 0; i < 4; i++) {
      double x, y;
      undistort_point((i % 2) ? nx - 1 : 0.0, i >= 2 ? ny - 1 : 0.0, A, D, nullptr, nullptr, &x, &y);
      x = (double)(float)x;  // OpenCV keeps these four points in float32 matrices
      y = (double)(float)y;
      const V3 X = mat3_vec(Rk, {x, y, 1.0});
      const double z = X.z != 0.0 ? 1.0 / X.z : 1.0;
      ax += (double)(float)(fc * (X.x * z));
      ay += (double)(float)(fc * (X.y * z));
    }
    cc[k][0] = (nx - 1) / 2 - ax * 0.25;
    cc[k][1] = (ny - 1) / 2 - ay * 0.25;
  }
  // CALIB_ZERO_DISPARITY: both views share one principal point
  cc[0][0] = cc[1][0] = (cc[0][0] + cc[1][0]) * 0.5;
  cc[0][1] = cc[1][1] = (cc[0][1] + cc[1][1]) * 0.5;
  double p2t = tIdx * fc;

  // alpha = 0: zoom until only valid pixels are visible
  {
    Rect in1, out1, in2, out2;
    get_rectangles(K1, D1, o->R1, fc, cc[0][0], cc[0][1], W, H, &in1, &out1);
    get_rectangles(K2, D2, o->R2, fc, cc[1][0], cc[1][1], W, H, &in2, &out2);
    const double cx1_0 = cc[0][0], cy1_0 = cc[0][1], cx2_0 = cc[1][0], cy2_0 = cc[1][1];
    const double cx1 = nx * cx1_0 / nx, cy1 = ny * cy1_0 / ny, cx2 = nx * cx2_0 / nx, cy2 = ny * cy2_0 / ny;
    double s0 = std::fmax(std::fmax(std::fmax(cx1 / (cx1_0 - in1.x), cy1 / (cy1_0 - in1.y)), (nx - 1 - cx1) / (in1.x + in1.w - cx1_0)),
                          (ny - 1 - cy1) / (in1.y + in1.h - cy1_0));
    s0 = std::fmax(s0, std::fmax(std::fmax(std::fmax(cx2 / (cx2_0 - in2.x), cy2 / (cy2_0 - in2.y)), (nx - 1 - cx2) / (in2.x + in2.w - cx2_0)),
                                 (ny - 1 - cy2) / (in2.y + in2.h - cy2_0)));
    const double s = s0;  // s0 * (1 - alpha) + s1 * alpha with alpha = 0
    fc *= s;
    cc[0][0] = cx1; cc[0][1] = cy1; cc[1][0] = cx2; cc[1][1] = cy2;
    p2t = s * p2t;
    (void)out1; (void)out2;
  }
  for (int i = 0; i < 12; i++) o->P1[i] = o->P2[i] = 0.0;
  o->P1[0] = o->P1[5] = o->P2[0] = o->P2[5] = fc;
  o->P1[2] = cc[0][0]; o->P1[6] = cc[0][1]; o->P1[10] = 1.0;
  o->P2[2] = cc[1][0]; o->P2[6] = cc[1][1]; o->P2[10] = 1.0;
  o->P2[idx * 4 + 3] = p2t;
  for (int i = 0; i < 16; i++) o->Q[i] = 0.0;
  o->Q[0] = o->Q[5] = 1.0;
  o->Q[3] = -o->P1[2];
  o->Q[7] = -o->P1[6];
  o->Q[11] = o->P1[0];
  o->Q[14] = -1.0 / tIdx;
  o->Q[15] = (idx == 0 ? o->P1[2] - o->P2[2] : o->P1[6] - o->P2[6]) / tIdx;
  return std::isfinite(fc) && fc > 0.0;
}

// operands of k_init_rectify_map: iR = (P[:3,:3] * R)^-1 and the intrinsics / distortion of the camera
bool rectify_map_args(const double* K, const double* D, const double* R, const double* P, RectifyMapArgs* a) {
  double A[9], PR[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) A[r * 3 + c] = P[r * 4 + c];
  mat3_mul(A, R, PR);
  if (!mat3_inv(PR, a->iR)) return false;
  a->fx = K[0]; a->fy = K[4]; a->cx = K[2]; a->cy = K[5];
  a->k1 = D[0]; a->k2 = D[1]; a->p1 = D[2]; a->p2 = D[3]; a->k3 = D[4];
  return true;
}

}  // namespace svgpu

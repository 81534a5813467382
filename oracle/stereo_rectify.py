"""oracle/stereo_rectify.py — TEST INFRASTRUCTURE ONLY (never imported by the product).

numpy restatement of cv::stereoRectify(K1, D1, K2, D2, size, R, T, flags=CALIB_ZERO_DISPARITY, alpha) as the reference
calls it in StereoVision::StereoVision (/root/reference/src/lib.cpp:32-37), i.e. OpenCV 4.13's behaviour (the pinned
vcpkg dependency; OpenCV's sources are not part of /root/reference): Bouguet's algorithm with the mean focal length,
the principal point centred on the projected image corners (float32 storage for those four points) and the alpha
zoom from the inner / outer rectangles of a 9 x 9 grid of undistorted points.

Not pinned: calibrations whose distortion polynomial folds inside the image (normalised corner radius well above 1
with negative k1/k2): cv::undistortPoints does not converge there and OpenCV's answer differs from five plain
fixed-point iterations.

Pinned: tests/golden/rectify_small.npz holds cv2.stereoRectify outputs for seeded calibrations
(tests/golden/make_golden.py); tests/test_oracle.py checks this file against them (<= 1e-12 relative) and live
against cv2 where it is installed.
"""
import numpy as np


def rodrigues_vec2mat(om):
    th = np.linalg.norm(om)
    if th < 2.220446049250313e-16: return np.eye(3)
    c, s = np.cos(th), np.sin(th); c1 = 1 - c
    r = om / th
    rrt = np.outer(r, r)
    rx = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    return c * np.eye(3) + c1 * rrt + s * rx

def rodrigues_mat2vec(R):
    # OpenCV: SVD-orthogonalise, then axis-angle
    U, S, Vt = np.linalg.svd(R)
    R = U @ Vt
    r = np.array([R[2,1]-R[1,2], R[0,2]-R[2,0], R[1,0]-R[0,1]])
    s = np.sqrt((r*r).sum()*0.25)
    c = (R[0,0]+R[1,1]+R[2,2]-1)*0.5
    c = min(max(c,-1.0),1.0)
    th = np.arccos(c)
    if s < 1e-5:
        if c > 0: return np.zeros(3)
        t = (R[0,0]+1)*0.5; rx = np.sqrt(max(t,0.0))
        t = (R[1,1]+1)*0.5; ry = np.sqrt(max(t,0.0))*(-1.0 if R[0,1]<0 else 1.0)
        t = (R[2,2]+1)*0.5; rz = np.sqrt(max(t,0.0))*(-1.0 if R[0,2]<0 else 1.0)
        if abs(rx) < abs(ry) and abs(rx) < abs(rz) and (R[1,2] > 0) != (ry*rz > 0): rz = -rz
        v = np.array([rx,ry,rz]); return v * (th/np.linalg.norm(v))
    vth = 1/(2*s) * th
    return r * vth

def undistort_points(pts, K, dist, R=None, P=None, iters=5):
    fx, fy, cx, cy = K[0,0], K[1,1], K[0,2], K[1,2]
    k = np.zeros(14); k[:len(dist)] = dist
    out = []
    for (u, v) in pts:
        x = (u - cx)/fx; y = (v - cy)/fy
        x0, y0 = x, y
        for _ in range(iters):
            r2 = x*x + y*y
            icdist = (1 + ((k[7]*r2 + k[6])*r2 + k[5])*r2)/(1 + ((k[4]*r2 + k[1])*r2 + k[0])*r2)
            dx = 2*k[2]*x*y + k[3]*(r2 + 2*x*x) + k[8]*r2 + k[9]*r2*r2
            dy = k[2]*(r2 + 2*y*y) + 2*k[3]*x*y + k[10]*r2 + k[11]*r2*r2
            x = (x0 - dx)*icdist; y = (y0 - dy)*icdist
        if R is not None:
            xx = R[0,0]*x + R[0,1]*y + R[0,2]; yy = R[1,0]*x + R[1,1]*y + R[1,2]; ww = 1.0/(R[2,0]*x + R[2,1]*y + R[2,2])
            x = xx*ww; y = yy*ww
        if P is not None:
            x = x*P[0,0] + P[0,2]; y = y*P[1,1] + P[1,2]
        out.append((x, y))
    return np.array(out)

def get_rectangles(K, dist, R, newK, size):
    N = 9
    w, h = size
    pts = [(float(x)*(w-1)/(N-1), float(y)*(h-1)/(N-1)) for y in range(N) for x in range(N)]
    p = undistort_points(pts, K, dist, R, newK)
    p = p.reshape(N, N, 2)
    iX0, iX1, iY0, iY1 = -np.inf, np.inf, -np.inf, np.inf
    oX0, oX1, oY0, oY1 = np.inf, -np.inf, np.inf, -np.inf
    for y in range(N):
        for x in range(N):
            px, py = p[y, x]
            oX0 = min(oX0, px); oX1 = max(oX1, px); oY0 = min(oY0, py); oY1 = max(oY1, py)
            if x == 0: iX0 = max(iX0, px)
            if x == N-1: iX1 = min(iX1, px)
            if y == 0: iY0 = max(iY0, py)
            if y == N-1: iY1 = min(iY1, py)
    return (iX0, iY0, iX1-iX0, iY1-iY0), (oX0, oY0, oX1-oX0, oY1-oY0)

def stereo_rectify(K1, D1, K2, D2, size, R, T, alpha=0.0, zero_disp=True):
    om = rodrigues_mat2vec(R) * -0.5
    r_r = rodrigues_vec2mat(om)
    t = r_r @ T
    idx = 0 if abs(t[0]) > abs(t[1]) else 1
    c = t[idx]; nt = np.linalg.norm(t)
    uu = np.zeros(3); uu[idx] = 1.0 if c > 0 else -1.0
    ww = np.cross(t, uu); nw = np.linalg.norm(ww)
    if nw > 0: ww = ww * (np.arccos(abs(c)/nt)/nw)
    wR = rodrigues_vec2mat(ww)
    R1 = wR @ r_r.T
    R2 = wR @ r_r
    t = R2 @ T
    nx, ny = size
    cc_new = [None, None]
    fc_new = (K1[idx ^ 1, idx ^ 1] + K2[idx ^ 1, idx ^ 1]) * 0.5
    for k, (A, Dk, Rk) in enumerate(((K1, D1, R1), (K2, D2, R2))):
        pts = [((nx-1) if (i % 2) else 0.0, (ny-1) if i >= 2 else 0.0) for i in range(4)]
        und = undistort_points(pts, A, Dk).astype(np.float32).astype(np.float64)   # CV_32FC2 storage
        # project with rotation Rk, K = diag(fc_new), zero distortion
        acc = np.zeros(2)
        for (x, y) in und:
            X = Rk @ np.array([x, y, 1.0])
            z = 1.0/X[2] if X[2] else 1.0
            acc += np.array([np.float32(fc_new*(X[0]*z)), np.float32(fc_new*(X[1]*z))], dtype=np.float64)
        avg = acc * 0.25   # cvAvg
        cc_new[k] = np.array([(nx-1)/2 - avg[0], (ny-1)/2 - avg[1]])
    if zero_disp:
        cc_new[0] = cc_new[1] = (cc_new[0] + cc_new[1]) * 0.5
    elif idx == 0:
        cc_new[0][1] = cc_new[1][1] = (cc_new[0][1] + cc_new[1][1])*0.5
    else:
        cc_new[0][0] = cc_new[1][0] = (cc_new[0][0] + cc_new[1][0])*0.5
    P1 = np.zeros((3, 4)); P2 = np.zeros((3, 4))
    for P, cc in ((P1, cc_new[0]), (P2, cc_new[1])):
        P[0,0] = P[1,1] = fc_new; P[0,2] = cc[0]; P[1,2] = cc[1]; P[2,2] = 1
    P2[idx, 3] = t[idx]*fc_new
    if alpha >= 0:
        alpha = min(alpha, 1.0)
        newK1 = P1[:, :3]; newK2 = P2[:, :3]
        in1, out1 = get_rectangles(K1, D1, R1, newK1, size)
        in2, out2 = get_rectangles(K2, D2, R2, newK2, size)
        cx1_0, cy1_0 = cc_new[0]; cx2_0, cy2_0 = cc_new[1]
        cx1, cy1 = nx*cx1_0/nx, ny*cy1_0/ny   # newImgSize == imageSize
        cx2, cy2 = nx*cx2_0/nx, ny*cy2_0/ny
        s0 = max(cx1/(cx1_0 - in1[0]), cy1/(cy1_0 - in1[1]), (nx-1-cx1)/(in1[0]+in1[2]-cx1_0), (ny-1-cy1)/(in1[1]+in1[3]-cy1_0))
        s0 = max(s0, cx2/(cx2_0 - in2[0]), cy2/(cy2_0 - in2[1]), (nx-1-cx2)/(in2[0]+in2[2]-cx2_0), (ny-1-cy2)/(in2[1]+in2[3]-cy2_0))
        s1 = min(cx1/(cx1_0 - out1[0]), cy1/(cy1_0 - out1[1]), (nx-1-cx1)/(out1[0]+out1[2]-cx1_0), (ny-1-cy1)/(out1[1]+out1[3]-cy1_0))
        s1 = min(s1, cx2/(cx2_0 - out2[0]), cy2/(cy2_0 - out2[1]), (nx-1-cx2)/(out2[0]+out2[2]-cx2_0), (ny-1-cy2)/(out2[1]+out2[3]-cy2_0))
        s = s0*(1-alpha) + s1*alpha
        fc_new *= s
        P1[0,0] = P1[1,1] = P2[0,0] = P2[1,1] = fc_new
        P1[0,2] = cx1; P1[1,2] = cy1; P2[0,2] = cx2; P2[1,2] = cy2
        P2[idx,3] = s*P2[idx,3]
    Q = np.zeros((4,4))
    Q[0,0] = Q[1,1] = 1; Q[0,3] = -P1[0,2]; Q[1,3] = -P1[1,2]; Q[2,3] = P1[0,0]
    Q[3,2] = -1.0/t[idx]
    Q[3,3] = ((P1[0,2]-P2[0,2]) if idx == 0 else (P1[1,2]-P2[1,2]))/t[idx]
    return R1, R2, P1, P2, Q

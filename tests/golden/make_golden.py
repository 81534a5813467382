"""Generates the golden vectors under tests/golden/ from the OpenCV binary (cv2) that the
reference links against, following the reference's own call sequence in
/root/reference/src/lib.cpp.  Run in the build container (needs cv2 and /root/reference):

    python tests/golden/make_golden.py

Outputs (committed):
  sgbm_small.npz   small 8UC1 pairs + cv2.StereoSGBM disparities for MODE_SGBM / MODE_HH / MODE_HH4
  bm_small.npz     small pairs + cv2.StereoBM disparities
  stages_small.npz median / filterSpeckles / morph close / remap / reproject / pre-filter chain / resize cases
  big_sha.npz      sha256 digests of cv2's full-size outputs: 3840x2160 / 256 disparities (gray pair, rectified pair + XYZ,
                   BGR pair through the whole chain) and 1920x1080 / 128 with the built-in speckle filter
  bottle.npz       the reference's bottle pair (resources/test-images-v1) through the mirror of
                   StereoVision::AnalyzeAndAnnotateImage with Algorithm::kSGBM (src/lib.cpp:53-85):
                   rectified gray inputs of StereoSGBM, raw and post-filtered disparity, calibration
                   products (maps are regenerated from the YAML values stored alongside)
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
from svgpu.synth import synth_pair  # noqa: E402

REF = "/root/reference"


def cv_sgbm(l, r, k):
    s = cv2.StereoSGBM_create(minDisparity=k["minDisparity"], numDisparities=k["numDisparities"], blockSize=k["blockSize"],
                              P1=k["P1"], P2=k["P2"], disp12MaxDiff=k["disp12MaxDiff"], preFilterCap=k["preFilterCap"],
                              uniquenessRatio=k["uniquenessRatio"], speckleWindowSize=k["speckleWindowSize"],
                              speckleRange=k["speckleRange"], mode=k["mode"])
    return s.compute(l, r)


PARAM_SETS = [
    dict(blockSize=7, P1=392, P2=1568, uniquenessRatio=15, preFilterCap=32, disp12MaxDiff=1, speckleWindowSize=0, speckleRange=0),
    dict(blockSize=5, P1=200, P2=800, uniquenessRatio=10, preFilterCap=31, disp12MaxDiff=1, speckleWindowSize=0, speckleRange=0),
    dict(blockSize=3, P1=8, P2=32, uniquenessRatio=0, preFilterCap=63, disp12MaxDiff=2, speckleWindowSize=0, speckleRange=0),
    dict(blockSize=1, P1=0, P2=0, uniquenessRatio=-1, preFilterCap=0, disp12MaxDiff=0, speckleWindowSize=0, speckleRange=0),
    dict(blockSize=7, P1=392, P2=1568, uniquenessRatio=15, preFilterCap=32, disp12MaxDiff=1, speckleWindowSize=60, speckleRange=2),
]


def sgbm_cases():
    cases = []
    rng = np.random.default_rng(1234)
    for (H, W, D) in [(33, 100, 32), (48, 200, 64), (40, 300, 224), (21, 40, 48)]:
        for kind in ("synth", "noise", "const"):
            if kind == "synth":
                l, r = synth_pair(H, W, D, seed=H)
            elif kind == "noise":
                l = rng.integers(0, 256, (H, W), dtype=np.uint8)
                r = rng.integers(0, 256, (H, W), dtype=np.uint8)
            else:
                l = np.full((H, W), 77, np.uint8)
                r = np.full((H, W), 77, np.uint8)
            for mode in (0, 1, 3):
                for pi, prm in enumerate(PARAM_SETS):
                    for minD in ((0,) if kind != "synth" else (0, 4, -8)):
                        if (pi in (2, 3, 4)) and (D == 224 or minD != 0):
                            continue
                        if W - (minD + D) <= (prm["blockSize"] // 2) and mode != 3:
                            continue  # cv2 throws for MODE_SGBM / MODE_HH (stereosgbm.cpp:511); HH4 returns all-invalid
                        k = dict(minDisparity=minD, numDisparities=D, mode=mode, **prm)
                        cases.append((l, r, k, cv_sgbm(l, r, k)))
    return cases


def save_sgbm(path):
    cases = sgbm_cases()
    out = {"n": np.int64(len(cases))}
    keys = ["minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
            "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode"]
    for i, (l, r, k, d) in enumerate(cases):
        out[f"l{i}"], out[f"r{i}"], out[f"d{i}"] = l, r, d
        out[f"p{i}"] = np.array([k[x] for x in keys], np.int32)
    np.savez_compressed(path, **out)
    print("sgbm cases", len(cases), os.path.getsize(path))


def save_bm(path):
    out = {}
    i = 0
    for (H, W, D, bs, uq, tex, spk) in [(40, 120, 32, 9, 15, 10, 0), (41, 130, 48, 5, 0, 0, 0), (36, 120, 16, 21, 15, 500, 0),
                                        (64, 200, 64, 21, 15, 10, 0), (35, 97, 16, 7, 5, 3, 0), (64, 200, 32, 11, 10, 10, 40)]:
        for seed in range(2):
            l, r = synth_pair(H, W, D, seed)
            bm = cv2.StereoBM_create(D, bs)
            bm.setUniquenessRatio(uq)
            bm.setTextureThreshold(tex)
            if spk:
                bm.setSpeckleWindowSize(spk)
                bm.setSpeckleRange(2)
            out[f"l{i}"], out[f"r{i}"], out[f"d{i}"] = l, r, bm.compute(l, r)
            out[f"p{i}"] = np.array([0, D, bs, 31, tex, uq, spk, 2 if spk else 0, -1], np.int32)
            i += 1
    out["n"] = np.int64(i)
    np.savez_compressed(path, **out)
    print("bm cases", i, os.path.getsize(path))


def load_calib(path):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    n = fs.getNode("calibration")
    g = lambda k: n.getNode(k).mat()  # noqa: E731
    sz = n.getNode("image_size")
    size = (int(sz.at(0).real()), int(sz.at(1).real()))
    return dict(size=size, Kl=g("camera_matrix_left"), Dl=g("dist_coeffs_left"), Kr=g("camera_matrix_right"),
                Dr=g("dist_coeffs_right"), R=g("R"), T=g("T"))


def rectify_products(cal):
    """StereoVision::StereoVision — src/lib.cpp:29-47."""
    size = cal["size"]
    Rl, Rr, Pl, Pr, Q, _, _ = cv2.stereoRectify(cal["Kl"], cal["Dl"], cal["Kr"], cal["Dr"], size, cal["R"], cal["T"],
                                                flags=cv2.CALIB_ZERO_DISPARITY, alpha=0, newImageSize=size)
    ml = cv2.initUndistortRectifyMap(cal["Kl"], cal["Dl"], Rl, Pl, size, cv2.CV_16SC2)
    mr = cv2.initUndistortRectifyMap(cal["Kr"], cal["Dr"], Rr, Pr, size, cv2.CV_16SC2)
    return dict(Rl=Rl, Rr=Rr, Pl=Pl, Pr=Pr, Q=Q, ml=ml, mr=mr)


def save_stages(path):
    rng = np.random.default_rng(99)
    out = {}
    a = rng.integers(-16, 2000, (37, 53)).astype(np.int16)
    out["median_in"], out["median_out"] = a, cv2.medianBlur(a, 3)
    b = (rng.integers(0, 6, (30, 40)) * 40 - 16).astype(np.int16)
    b = cv2.resize(b, (160, 120), interpolation=cv2.INTER_NEAREST)
    out["spk_in"] = b
    spk = [(0, 150, 32), (-16, 60, 32), (-16, 20, 2), (0, 5, 100)]
    out["spk_params"] = np.array(spk, np.int32)
    for i, (nv, ms, md) in enumerate(spk):
        c = b.copy()
        cv2.filterSpeckles(c, nv, ms, md)
        out[f"spk_out{i}"] = c
    out["close_out"] = cv2.morphologyEx(b, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5)))
    # rectification of a small crop-sized camera: v1 calibration scaled to 160x120
    cal = load_calib(f"{REF}/resources/calibration/v1/calibration.yml")
    s = 0.25
    for k in ("Kl", "Kr"):
        cal[k] = cal[k].copy()
        cal[k][:2] *= s
    cal["size"] = (160, 120)
    pr = rectify_products(cal)
    for k in ("Kl", "Dl", "Kr", "Dr", "R", "T"):
        out["cal_" + k] = cal[k]
    for k in ("Rl", "Rr", "Pl", "Pr", "Q"):
        out["rect_" + k] = pr[k]
    out["map1_l"], out["map2_l"] = pr["ml"]
    out["map1_r"], out["map2_r"] = pr["mr"]
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    out["remap_in"] = img
    out["remap_out_l"] = cv2.remap(img, pr["ml"][0], pr["ml"][1], cv2.INTER_LINEAR)
    out["remap_out_r_gray"] = cv2.remap(img[:, :, 0].copy(), pr["mr"][0], pr["mr"][1], cv2.INTER_LINEAR)
    d = rng.integers(-16, 3000, (120, 160)).astype(np.int16)
    out["reproj_in"] = d
    out["reproj_out_hm"] = cv2.reprojectImageTo3D(d, pr["Q"], handleMissingValues=True)
    out["reproj_out"] = cv2.reprojectImageTo3D(d, pr["Q"], handleMissingValues=False)
    blur = cv2.GaussianBlur(img, (11, 11), 1.4)
    gray = cv2.cvtColor(blur, cv2.COLOR_BGR2GRAY)
    k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
    out["chain_blur"], out["chain_gray"], out["chain_sharp"] = blur, gray, cv2.filter2D(gray, -1, k)
    src = rng.integers(0, 256, (333, 517, 3), dtype=np.uint8)
    out["resize_in"], out["resize_out"] = src, cv2.resize(src, (160, 120), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(path, **out)
    print("stages", os.path.getsize(path))


def save_bottle(path):
    """Config 1 of BASELINE.json: the reference's own pair through its own chain (mirror of src/lib.cpp)."""
    cal = load_calib(f"{REF}/resources/calibration/v1/calibration.yml")
    pr = rectify_products(cal)
    size = cal["size"]
    L = cv2.imread(f"{REF}/resources/test-images-v1/left/bottle_000.jpg")
    R = cv2.imread(f"{REF}/resources/test-images-v1/right/bottle_000.jpg")
    # RescaleImage (src/lib.cpp:96-104) + RectifyImages (:87-94)
    Ls = cv2.resize(L, size, interpolation=cv2.INTER_LINEAR)
    Rs = cv2.resize(R, size, interpolation=cv2.INTER_LINEAR)
    Lr = cv2.remap(Ls, pr["ml"][0], pr["ml"][1], cv2.INTER_LINEAR)
    Rr = cv2.remap(Rs, pr["mr"][0], pr["mr"][1], cv2.INTER_LINEAR)
    # CalculateDisparityMapUsingSGBM (:228-277)
    k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
    gl = cv2.filter2D(cv2.cvtColor(cv2.GaussianBlur(Lr, (11, 11), 1.4), cv2.COLOR_BGR2GRAY), -1, k)
    gr = cv2.filter2D(cv2.cvtColor(cv2.GaussianBlur(Rr, (11, 11), 1.4), cv2.COLOR_BGR2GRAY), -1, k)
    prm = dict(minDisparity=0, numDisparities=224, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32,
               uniquenessRatio=15, speckleWindowSize=0, speckleRange=2, mode=cv2.StereoSGBM_MODE_HH4)
    raw = cv_sgbm(gl, gr, prm)
    post = raw.copy()
    post[post == -1] = 0
    cv2.filterSpeckles(post, 0, 150, 32)
    post = cv2.morphologyEx(post, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5)))
    xyz = cv2.reprojectImageTo3D(post, pr["Q"], handleMissingValues=True)  # :152-156
    bf = np.float32(abs(cal["T"][0, 0])) * np.float32(cal["Kl"][0, 0])
    depth = np.where(post > 0, bf / post.astype(np.float32), np.float32(0)).astype(np.float32)  # :279-294
    out = dict(scaled_left=Ls, scaled_right=Rs, rect_left=Lr, rect_right=Rr, gray_left=gl, gray_right=gr, disp_raw=raw,
               disp_post=post, Q=pr["Q"], xyz_sha256=np.frombuffer(hashlib.sha256(xyz.tobytes()).digest(), np.uint8),
               depth_sha256=np.frombuffer(hashlib.sha256(depth.tobytes()).digest(), np.uint8),
               xyz_rows=xyz[200:204].copy(), depth_rows=depth[200:204].copy(),
               baseline=np.float32(abs(cal["T"][0, 0])), focal=np.float32(cal["Kl"][0, 0]))
    for kk in ("Kl", "Dl", "Kr", "Dr", "R", "T"):
        out["cal_" + kk] = cal[kk]
    for kk in ("Rl", "Rr", "Pl", "Pr"):
        out["rect_" + kk] = pr[kk]
    np.savez_compressed(path, **out)
    print("bottle", os.path.getsize(path), "raw sha", hashlib.sha256(raw.tobytes()).hexdigest()[:16],
          "post sha", hashlib.sha256(post.tobytes()).hexdigest()[:16], "valid", float((raw >= 0).mean()))


def save_sgbm3(path):
    """cv2.StereoSGBM on 8UC3 pairs (the cost is summed over the three channels), all three modes."""
    out = {}
    n = 0
    rng = np.random.default_rng(4321)
    for (H, W, D, bs, minD) in [(40, 150, 32, 5, 0), (33, 120, 16, 3, 0), (50, 200, 64, 7, 4), (30, 140, 48, 5, -8), (37, 300, 224, 3, 0)]:
        for mode in (cv2.STEREO_SGBM_MODE_SGBM, cv2.STEREO_SGBM_MODE_HH, cv2.STEREO_SGBM_MODE_HH4):
            for kind in ("synth", "noise"):
                if kind == "synth":
                    l1, r1 = synth_pair(H, W, D, seed=H)
                    l = np.stack([l1, np.roll(l1, 1, axis=0) // 2 + 30, 255 - l1], axis=2).astype(np.uint8)
                    r = np.stack([r1, np.roll(r1, 1, axis=0) // 2 + 30, 255 - r1], axis=2).astype(np.uint8)
                else:
                    l = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
                    r = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
                k = dict(minDisparity=minD, numDisparities=D, blockSize=bs, P1=8 * bs * bs, P2=32 * bs * bs, disp12MaxDiff=1, preFilterCap=31,
                         uniquenessRatio=10, speckleWindowSize=0, speckleRange=0, mode=mode)
                out[f"l{n}"], out[f"r{n}"], out[f"d{n}"] = l, r, cv_sgbm(l, r, k)
                out[f"p{n}"] = np.array([k[x] for x in ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
                                                        "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")], np.int32)
                n += 1
    out["n"] = np.array(n)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", n, "cases")


def save_sgbm_wide(path):
    """cv2.StereoSGBM with more than 256 disparities (272 .. 512): gray pairs in the three modes, and 8UC3 / large-block cases."""
    out = {}
    n = 0
    rng = np.random.default_rng(777)
    keys = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap", "uniquenessRatio", "speckleWindowSize",
            "speckleRange", "mode")

    def add(l, r, k):
        nonlocal n
        out[f"l{n}"], out[f"r{n}"], out[f"d{n}"] = l, r, cv_sgbm(l, r, k)
        out[f"p{n}"] = np.array([k[x] for x in keys], np.int32)
        n += 1

    for (H, W, D, bs, minD) in [(26, 340, 272, 7, 0), (22, 470, 384, 5, 0), (24, 590, 512, 7, 0), (20, 420, 320, 3, -6), (18, 600, 512, 11, 3)]:
        P1, P2 = 8 * bs * bs, 32 * bs * bs
        l, r = synth_pair(H, W, D, seed=H)
        for mode in (cv2.STEREO_SGBM_MODE_SGBM, cv2.STEREO_SGBM_MODE_HH, cv2.STEREO_SGBM_MODE_HH4):
            add(l, r, dict(minDisparity=minD, numDisparities=D, blockSize=bs, P1=P1, P2=min(P2, 2500), disp12MaxDiff=1, preFilterCap=31,
                           uniquenessRatio=10, speckleWindowSize=0, speckleRange=0, mode=mode))
        ln, rn = rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8)
        add(ln, rn, dict(minDisparity=minD, numDisparities=D, blockSize=bs, P1=P1, P2=min(P2, 2500), disp12MaxDiff=2, preFilterCap=63,
                         uniquenessRatio=5, speckleWindowSize=0, speckleRange=0, mode=cv2.STEREO_SGBM_MODE_HH))
    # 8UC3 and blockSize 13 beyond 256 disparities
    H, W, D = 20, 400, 320
    l1, r1 = synth_pair(H, W, D, seed=5)
    l3 = np.stack([l1, np.roll(l1, 1, axis=0) // 2 + 30, 255 - l1], axis=2).astype(np.uint8)
    r3 = np.stack([r1, np.roll(r1, 1, axis=0) // 2 + 30, 255 - r1], axis=2).astype(np.uint8)
    for mode in (cv2.STEREO_SGBM_MODE_HH, cv2.STEREO_SGBM_MODE_HH4):
        add(l3, r3, dict(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=31, uniquenessRatio=10,
                         speckleWindowSize=0, speckleRange=0, mode=mode))
    add(l1, r1, dict(minDisparity=0, numDisparities=D, blockSize=13, P1=300, P2=1200, disp12MaxDiff=1, preFilterCap=15, uniquenessRatio=10,
                     speckleWindowSize=0, speckleRange=0, mode=cv2.STEREO_SGBM_MODE_HH))
    # MODE_SGBM_3WAY beyond 256 disparities (four stripe frames per pair)
    for (H, W, D) in [(37, 340, 272), (26, 590, 512)]:
        l, r = synth_pair(H, W, D, seed=H + 1)
        add(l, r, dict(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=31, uniquenessRatio=10,
                       speckleWindowSize=0, speckleRange=0, mode=cv2.STEREO_SGBM_MODE_SGBM_3WAY))
    out["n"] = np.array(n)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", n, "cases")


def save_sgbm3way(path):
    """cv2.StereoSGBM MODE_SGBM_3WAY: textured and noise pairs (noise exercises the tie-breaking of OpenCV's vectorised
    winner search and its uniqueness threshold), images smaller than the stripe overlap, minDisparity of both signs."""
    out = {}
    n = 0
    rng = np.random.default_rng(99)
    for (H, W, D, P1, P2, uniq, bs, minD, d12) in [(64, 160, 32, 200, 800, 10, 5, 0, 1), (57, 180, 64, 392, 1568, 15, 7, 0, 1), (41, 150, 16, 8, 32, 10, 3, 0, 1),
                                                   (50, 170, 32, 100, 400, 10, 5, 5, 1), (50, 170, 32, 100, 400, 10, 5, -6, 2), (9, 120, 16, 100, 400, 10, 5, 0, 1),
                                                   (3, 120, 16, 100, 400, 10, 3, 0, 1), (17, 120, 16, 100, 400, 10, 11, 0, 1), (40, 140, 32, 2, 5, 5, 1, 0, 1),
                                                   (40, 140, 32, 100, 400, 99, 5, 0, 1), (40, 330, 224, 392, 1568, 15, 7, 0, 1), (48, 240, 128, 72, 288, 0, 3, 0, 100)]:
        for kind in ("synth", "noise"):
            if kind == "synth":
                l, r = synth_pair(H, W, max(D + minD, 8), seed=H + D)
            else:
                l = rng.integers(0, 256, (H, W), dtype=np.uint8)
                r = rng.integers(0, 256, (H, W), dtype=np.uint8)
            k = dict(minDisparity=minD, numDisparities=D, blockSize=bs, P1=P1, P2=P2, disp12MaxDiff=d12, preFilterCap=31,
                     uniquenessRatio=uniq, speckleWindowSize=0, speckleRange=0, mode=cv2.STEREO_SGBM_MODE_SGBM_3WAY)
            out[f"l{n}"], out[f"r{n}"], out[f"d{n}"] = l, r, cv_sgbm(l, r, k)
            out[f"p{n}"] = np.array([k[x] for x in ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
                                                    "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")], np.int32)
            n += 1
    out["n"] = np.array(n)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", n, "cases")


def save_bm_mind(path):
    """cv2.StereoBM with minDisparity != 0.  OpenCV indexes past the end of the image rows there, so only the pixels
    inside cv::getValidDisparityROI are reproducible; the ROI (y0, y1, x0, x1) is stored with every case."""
    out = {}
    n = 0
    for (H, W, D, minD, bs, uniq) in [(60, 200, 32, 4, 9, 15), (60, 200, 32, -8, 9, 15), (80, 300, 64, 16, 15, 10), (61, 203, 32, 1, 5, 0),
                                      (60, 200, 16, -3, 7, 15), (70, 260, 48, -40, 11, 5), (70, 260, 16, -20, 7, 15), (50, 150, 16, 33, 5, 15)]:
        l, r = synth_pair(H, W, max(D + minD, 8), seed=H + n)
        bm = cv2.StereoBM_create(numDisparities=D, blockSize=bs)
        bm.setMinDisparity(minD)
        bm.setUniquenessRatio(uniq)
        prm = [minD, D, bs, bm.getPreFilterCap(), bm.getTextureThreshold(), uniq, 0, 0, -1]
        w2 = bs // 2
        out[f"l{n}"], out[f"r{n}"], out[f"p{n}"], out[f"d{n}"] = l, r, np.array(prm, np.int32), bm.compute(l, r)
        out[f"roi{n}"] = np.array([w2, H - w2, max(minD + D - 1, 0) + w2, W - w2 + min(minD, 0)], np.int32)
        n += 1
    out["n"] = np.array(n)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def save_rectify(path):
    """cv::stereoRectify(CALIB_ZERO_DISPARITY, alpha 0) + 2x cv::initUndistortRectifyMap(CV_16SC2) as the reference's
    constructor calls them (src/lib.cpp:32-46) for seeded calibrations, plus the bottle calibration of the reference."""
    rng = np.random.default_rng(77)
    out = {}
    cases = []
    for i, (W, H) in enumerate([(200, 150), (320, 180), (161, 97), (240, 320)]):
        f = 0.75 * W * (1 + 0.05 * rng.normal())
        K1 = np.array([[f, 0, W / 2 + rng.normal() * 3], [0, f * (1 + 0.01 * rng.normal()), H / 2 + rng.normal() * 3], [0, 0, 1]])
        K2 = np.array([[f * (1 + 0.01 * rng.normal()), 0, W / 2 + rng.normal() * 3], [0, f * (1 + 0.01 * rng.normal()), H / 2 + rng.normal() * 3], [0, 0, 1]])
        D1 = np.array([rng.normal() * 0.1, rng.normal() * 0.05, 0, 0, 0]) if i % 2 == 0 else rng.normal(size=5) * np.array([0.1, 0.05, 0.002, 0.002, 0.01])
        D2 = np.array([rng.normal() * 0.1, rng.normal() * 0.05, 0, 0, 0]) if i % 2 == 0 else rng.normal(size=5) * np.array([0.1, 0.05, 0.002, 0.002, 0.01])
        R, _ = cv2.Rodrigues(rng.normal(size=3) * 0.02)
        T = np.array([-120.0 + rng.normal(), rng.normal(), rng.normal()]) if i != 3 else np.array([rng.normal(), 95.0 + rng.normal(), rng.normal()])
        cases.append((f"c{i}", K1, D1, K2, D2, R, T, W, H))
    for tag in ("v1", "v2"):   # the reference's own calibrations
        cal = load_calib(f"{REF}/resources/calibration/{tag}/calibration.yml")
        W, H = cal["size"]
        cases.append((f"ref{tag}", cal["Kl"], cal["Dl"].reshape(-1)[:5], cal["Kr"], cal["Dr"].reshape(-1)[:5], cal["R"], cal["T"].reshape(3), W, H))
    for (n, K1, D1, K2, D2, R, T, W, H) in cases:
        R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K1, D1, K2, D2, (W, H), R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0)
        for k, v in dict(K1=K1, D1=D1, K2=K2, D2=D2, R=R, T=T, size=np.array([W, H]), R1=R1, R2=R2, P1=P1, P2=P2, Q=Q).items():
            out[f"{n}_{k}"] = np.asarray(v, np.float64 if k != "size" else np.int64)
        if not n.startswith("ref"):   # full-size maps of the reference calibration are covered by bottle.npz
            for side, (K, D, Rr, P) in enumerate(((K1, D1, R1, P1), (K2, D2, R2, P2))):
                m1, m2 = cv2.initUndistortRectifyMap(K, D, Rr, P, (W, H), cv2.CV_16SC2)
                out[f"{n}_map1_{side}"] = m1
                out[f"{n}_map2_{side}"] = m2
    out["names"] = np.array([c[0] for c in cases])
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", [c[0] for c in cases])


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def big_calib(W, H):
    """The synthetic single-camera calibration of the full-size tests (same numbers as tests/test_gpu_parity.py)."""
    K = np.array([[833.77 * 6, 0, 345.2 * 6], [0, 832.87 * 6, 283.1 * 4.5], [0, 0, 1]])
    dist = np.array([-0.4168, 0.166, 0, 0, 0])
    P = np.array([[K[0, 0] * 0.9, 0, K[0, 2], 0], [0, K[0, 0] * 0.9, K[1, 2], 0], [0, 0, 1, 0]])
    Q = np.array([[1, 0, 0, -K[0, 2]], [0, 1, 0, -K[1, 2]], [0, 0, 0, P[0, 0]], [0, 0, 1 / 75.3, 0]])
    return K, dist, P, Q


def save_big(path):
    """Full-size cases of BASELINE configs 3 and 4, stored as sha256 digests of cv2's outputs (the inputs are regenerated from
    their seeds by the tests; their digests are stored too, so a synth difference shows up as such).
      c4_*   3840x2160, 256 disparities, MODE_HH: StereoSGBM on the gray pair; remap -> StereoSGBM -> reprojectImageTo3D
      c4bgr_* the same size through the reference's whole chain on a BGR pair (src/lib.cpp:53-85 with MODE_HH, D=256)
      c3spk_* 1920x1080, 128 disparities, MODE_HH with speckleWindowSize=100, speckleRange=2, three pairs"""
    out = {}
    H, W, D = 2160, 3840, 256
    prm = dict(minDisparity=0, numDisparities=D, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32, uniquenessRatio=15,
               speckleWindowSize=0, speckleRange=0, mode=cv2.STEREO_SGBM_MODE_HH)
    l, r = synth_pair(H, W, D, seed=4)
    out["c4_in"] = sha(np.stack([l, r]))
    d = cv_sgbm(l, r, prm)
    out["c4_disp"] = sha(d)
    print("c4 direct valid", float((d[:, D:] >= 0).mean()))
    K, dist, P, Q = big_calib(W, H)
    m1, m2 = cv2.initUndistortRectifyMap(K, dist, np.eye(3), P, (W, H), cv2.CV_16SC2)
    lr_, rr_ = cv2.remap(l, m1, m2, cv2.INTER_LINEAR), cv2.remap(r, m1, m2, cv2.INTER_LINEAR)
    out["c4_rect"] = sha(np.stack([lr_, rr_]))
    d2 = cv_sgbm(lr_, rr_, prm)
    out["c4_rect_disp"] = sha(d2)
    out["c4_rect_xyz"] = sha(cv2.reprojectImageTo3D(d2, Q, handleMissingValues=True))
    # BGR through the reference's chain
    lb = np.stack([l, np.roll(l, 1, axis=0) // 2 + 30, 255 - l], axis=2).astype(np.uint8)
    rb = np.stack([r, np.roll(r, 1, axis=0) // 2 + 30, 255 - r], axis=2).astype(np.uint8)
    out["c4bgr_in"] = sha(np.stack([lb, rb]))
    Lr, Rr = cv2.remap(lb, m1, m2, cv2.INTER_LINEAR), cv2.remap(rb, m1, m2, cv2.INTER_LINEAR)
    k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.float32)
    gl = cv2.filter2D(cv2.cvtColor(cv2.GaussianBlur(Lr, (11, 11), 1.4), cv2.COLOR_BGR2GRAY), -1, k)
    gr = cv2.filter2D(cv2.cvtColor(cv2.GaussianBlur(Rr, (11, 11), 1.4), cv2.COLOR_BGR2GRAY), -1, k)
    out["c4bgr_gray"] = sha(np.stack([gl, gr]))
    raw = cv_sgbm(gl, gr, prm)
    out["c4bgr_raw"] = sha(raw)
    post = raw.copy()
    post[post == -1] = 0
    cv2.filterSpeckles(post, 0, 150, 32)
    post = cv2.morphologyEx(post, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5)))
    out["c4bgr_post"] = sha(post)
    out["c4bgr_xyz"] = sha(cv2.reprojectImageTo3D(post, Q, handleMissingValues=True))
    bf = np.float32(75.3) * np.float32(K[0, 0])
    out["c4bgr_depth"] = sha(np.where(post > 0, bf / post.astype(np.float32), np.float32(0)).astype(np.float32))
    print("c4bgr valid", float((raw[:, D:] >= 0).mean()))
    # 1080p with the speckle filter inside StereoSGBM::compute
    H, W, D = 1080, 1920, 128
    prm = dict(prm, numDisparities=D, speckleWindowSize=100, speckleRange=2)
    for i in range(3):
        l, r = synth_pair(H, W, D, seed=10 + i)
        out[f"c3spk_in{i}"] = sha(np.stack([l, r]))
        d = cv_sgbm(l, r, prm)
        out[f"c3spk_disp{i}"] = sha(d)
        out[f"c3_disp{i}"] = sha(cv_sgbm(l, r, dict(prm, speckleWindowSize=0, speckleRange=0)))
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def save_calibration_yml(path):
    """A calibration file as the reference's calibration application writes it (StereoCameraInfo.hpp:51-65: cv::FileStorage,
    node `calibration`), from seeded values with awkward decimals; the facade's reader / writer is pinned to it byte for byte."""
    rng = np.random.default_rng(2026)
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
    fs.startWriteStruct("calibration", cv2.FileNode_MAP)
    fs.write("file_version", 1)
    fs.startWriteStruct("image_size", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
    fs.write("", 1920)
    fs.write("", 1080)
    fs.endWriteStruct()
    K = lambda: np.array([[1400 + rng.normal() * 30, 0, 960 + rng.normal() * 9], [0, 1400 + rng.normal() * 30, 540 + rng.normal() * 9], [0, 0, 1]])  # noqa: E731
    fs.write("camera_matrix_left", K())
    fs.write("camera_matrix_right", K())
    fs.write("dist_coeffs_left", (rng.normal(size=(1, 5)) * [0.3, 0.1, 1e-3, 1e-3, 1e-2]))
    fs.write("dist_coeffs_right", (rng.normal(size=(1, 5)) * [0.3, 0.1, 1e-3, 1e-3, 1e-2]))
    R, _ = cv2.Rodrigues(rng.normal(size=3) * 0.03)
    fs.write("R", R)
    fs.write("T", np.array([[-120.5 + rng.normal()], [rng.normal()], [rng.normal() * 1e-3]]))
    fs.write("E", rng.normal(size=(3, 3)) * 50)
    fs.write("F", rng.normal(size=(3, 3)) * 1e-7)
    fs.write("reprojection_error", float(abs(rng.normal()) * 0.3))
    fs.endWriteStruct()
    fs.release()
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    only = sys.argv[1:]
    jobs = dict(big_sha=save_big, sgbm_small=save_sgbm, bm_small=save_bm, stages_small=save_stages, bottle=save_bottle, rectify_small=save_rectify, bm_mind_small=save_bm_mind, sgbm3_small=save_sgbm3, sgbm3way_small=save_sgbm3way, sgbm_wide=save_sgbm_wide)
    for name, fn in jobs.items():
        if not only or name in only:
            fn(os.path.join(HERE, name + ".npz"))
    if not only or "calibration_cv2" in only:
        save_calibration_yml(os.path.join(HERE, "calibration_cv2.yml"))

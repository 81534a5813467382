"""CPU tests: the oracle (oracle/) against the golden vectors produced by the OpenCV binary
(tests/golden/make_golden.py) and, when cv2 is importable, against cv2 live."""
import hashlib

import numpy as np
import pytest

from helpers import (bm_golden_cases, bm_mind_golden_cases, golden, nmismatch, sgbm3_golden_cases, sgbm3way_golden_cases,
                     sgbm_golden_cases, sgbm_wide_cases)


def test_sgbm_golden(oracle):
    n = 0
    for i, l, r, prm, want in sgbm_golden_cases():
        got = oracle.sgbm(l, r, **prm)
        assert nmismatch(got, want) == 0, (i, prm)
        n += 1
    assert n > 100


def test_sgbm_wide_disparity_golden(oracle):
    """numDisparities 272 .. 512 against cv2 (gray in the three modes, 8UC3, blockSize 11 / 13)."""
    n = 0
    for i, l, r, prm, want in sgbm_wide_cases():
        assert nmismatch(oracle.sgbm(l, r, **prm), want) == 0, (i, prm)
        n += 1
    assert n >= 20


def test_sgbm_three_channel_golden(oracle):
    """8UC3 input (not what the reference passes, but part of cv::StereoSGBM::compute): cost summed over the channels."""
    n = 0
    for i, l, r, prm, want in sgbm3_golden_cases():
        assert nmismatch(oracle.sgbm(l, r, **prm), want) == 0, (i, prm)
        n += 1
    assert n == 30


def test_sgbm_3way_golden(oracle):
    """MODE_SGBM_3WAY (OpenCV's computeDisparity3WAY), restated from its observable behaviour: four stripes with restarted
    vertical paths, the vectorised winner search's tie rule and uniqueness threshold, the stripe-buffer row offsets of tiny
    images.  Bit-exact on every golden case; uniquenessRatio = 100 is an error (cv2 itself dies with SIGFPE there)."""
    n = 0
    for i, l, r, prm, want in sgbm3way_golden_cases():
        assert nmismatch(oracle.sgbm(l, r, **prm), want) == 0, (i, prm)
        n += 1
    assert n == 24
    with pytest.raises(ValueError):
        oracle.sgbm(l, r, **dict(prm, uniquenessRatio=100))


def test_sgbm_too_narrow_is_an_error(oracle):
    l = np.zeros((21, 40), np.uint8)
    prm = dict(minDisparity=0, numDisparities=48, blockSize=7, P1=8, P2=32, disp12MaxDiff=1, preFilterCap=31,
               uniquenessRatio=10, speckleWindowSize=0, speckleRange=0)
    for mode in (0, 1):
        with pytest.raises(ValueError):
            oracle.sgbm(l, l, mode=mode, **prm)
    assert (oracle.sgbm(l, l, mode=3, **prm) == -16).all()


def test_bm_golden(oracle):
    for i, l, r, prm, want in bm_golden_cases():
        assert nmismatch(oracle.bm(l, r, **prm), want) == 0, (i, prm)


def test_bm_min_disparity_golden(oracle):
    """minDisparity != 0: bit-exact inside cv::getValidDisparityROI (outside it OpenCV's own output depends on what lies
    behind the end of the image rows in memory; the oracle writes FILTERED there)."""
    seen = 0
    for i, l, r, prm, want, (y0, y1, x0, x1) in bm_mind_golden_cases():
        got = oracle.bm(l, r, **prm)
        assert nmismatch(got[y0:y1, x0:x1], want[y0:y1, x0:x1]) == 0, (i, prm)
        outside = np.ones(got.shape, bool)
        outside[y0:y1, x0:x1] = False
        assert (got[outside] == (prm["minDisparity"] - 1) * 16).all(), (i, prm)
        assert (got[y0:y1, x0:x1] > (prm["minDisparity"] - 1) * 16).mean() > 0.2, (i, "case has too few valid pixels")
        seen += 1
    assert seen >= 8


def test_stages_golden(oracle):
    g = golden("stages_small")
    assert nmismatch(oracle.median3(g["median_in"]), g["median_out"]) == 0
    for i, (nv, ms, md) in enumerate(g["spk_params"]):
        assert nmismatch(oracle.filter_speckles(g["spk_in"], nv, ms, md), g[f"spk_out{i}"]) == 0
    assert nmismatch(oracle.morph_close5(g["spk_in"]), g["close_out"]) == 0
    for side in "lr":
        K, D = g["cal_K" + side], g["cal_D" + side]
        m1, m2 = oracle.init_undistort_rectify_map(K, D, g["rect_R" + side], g["rect_P" + side], (160, 120))
        assert nmismatch(m1, g["map1_" + side]) == 0 and nmismatch(m2, g["map2_" + side]) == 0
    assert nmismatch(oracle.remap(g["remap_in"], g["map1_l"], g["map2_l"]), g["remap_out_l"]) == 0
    assert nmismatch(oracle.remap(g["remap_in"][:, :, 0].copy(), g["map1_r"], g["map2_r"]), g["remap_out_r_gray"]) == 0
    for hm, key in ((True, "reproj_out_hm"), (False, "reproj_out")):
        got = oracle.reproject(g["reproj_in"], g["rect_Q"], hm)
        assert nmismatch(got.view(np.uint32), g[key].view(np.uint32)) == 0
    assert nmismatch(oracle.gaussian11(g["remap_in"]), g["chain_blur"]) == 0
    assert nmismatch(oracle.bgr2gray(g["chain_blur"]), g["chain_gray"]) == 0
    assert nmismatch(oracle.sharpen3(g["chain_gray"]), g["chain_sharp"]) == 0
    assert nmismatch(oracle.resize_linear(g["resize_in"], (160, 120)), g["resize_out"]) == 0


def test_bottle_chain_golden(oracle):
    """BASELINE config 1: the reference's bottle pair through the mirror of src/lib.cpp:53-85."""
    g = golden("bottle")
    size = (640, 480)
    ml = oracle.init_undistort_rectify_map(g["cal_Kl"], g["cal_Dl"], g["rect_Rl"], g["rect_Pl"], size)
    mr = oracle.init_undistort_rectify_map(g["cal_Kr"], g["cal_Dr"], g["rect_Rr"], g["rect_Pr"], size)
    Lr = oracle.remap(g["scaled_left"], *ml)
    Rr = oracle.remap(g["scaled_right"], *mr)
    assert nmismatch(Lr, g["rect_left"]) == 0 and nmismatch(Rr, g["rect_right"]) == 0
    gl = oracle.sharpen3(oracle.bgr2gray(oracle.gaussian11(Lr)))
    gr = oracle.sharpen3(oracle.bgr2gray(oracle.gaussian11(Rr)))
    assert nmismatch(gl, g["gray_left"]) == 0 and nmismatch(gr, g["gray_right"]) == 0
    raw = oracle.sgbm(gl, gr, **oracle.REFERENCE_SGBM)
    assert nmismatch(raw, g["disp_raw"]) == 0
    assert hashlib.sha256(raw.tobytes()).hexdigest()[:16] == "90f4df78b9ec67a4"
    post = oracle.reference_sgbm_chain(Lr, Rr)
    assert nmismatch(post, g["disp_post"]) == 0
    xyz = oracle.reproject(post, g["Q"], True)
    assert hashlib.sha256(xyz.tobytes()).digest() == g["xyz_sha256"].tobytes()
    depth = oracle.depth_simple(post, float(g["baseline"]), float(g["focal"]))
    assert hashlib.sha256(depth.tobytes()).digest() == g["depth_sha256"].tobytes()


def test_live_cv2_random(oracle, cv2):
    """Fresh random cases against the cv2 binary itself (not only the committed vectors)."""
    rng = np.random.default_rng()
    for _ in range(6):
        H, W = int(rng.integers(20, 50)), int(rng.integers(90, 160))
        D = int(rng.choice([16, 32, 48]))
        l = rng.integers(0, 256, (H, W), dtype=np.uint8)
        r = np.roll(l, -int(rng.integers(1, D // 2)), axis=1)
        r = np.clip(r.astype(int) + rng.integers(-4, 5, r.shape), 0, 255).astype(np.uint8)
        bs = int(rng.choice([3, 5, 7, 9]))
        prm = dict(minDisparity=int(rng.choice([0, 0, 3, -5])), numDisparities=D, blockSize=bs, P1=8 * bs * bs,
                   P2=32 * bs * bs, disp12MaxDiff=int(rng.choice([1, 2])), preFilterCap=int(rng.integers(5, 64)),
                   uniquenessRatio=int(rng.integers(0, 20)), speckleWindowSize=0, speckleRange=0,
                   mode=int(rng.choice([0, 1, 2, 3])))
        s = cv2.StereoSGBM_create(**prm)
        assert nmismatch(oracle.sgbm(l, r, **prm), s.compute(l, r)) == 0, prm
        d = s.compute(l, r)
        assert nmismatch(oracle.median3(d), cv2.medianBlur(d, 3)) == 0


def test_live_cv2_large_blocks(oracle, cv2):
    """blockSize 13 and 15 (the largest whose cost fits OpenCV's 16-bit cost type with a small pre-filter cap): oracle == cv2."""
    rng = np.random.default_rng(1315)
    for bs, mode, D in ((13, 0, 32), (13, 1, 48), (15, 3, 32), (15, 1, 16)):
        H, W = 40 + bs, 150 + D
        l = rng.integers(0, 256, (H, W), dtype=np.uint8)
        r = np.roll(l, -5, axis=1)
        r = np.clip(r.astype(int) + rng.integers(-6, 7, r.shape), 0, 255).astype(np.uint8)
        prm = dict(minDisparity=0, numDisparities=D, blockSize=bs, P1=100, P2=400, disp12MaxDiff=1, preFilterCap=15,
                   uniquenessRatio=10, speckleWindowSize=0, speckleRange=0, mode=mode)
        assert bs * bs * (2 * 15 + 63) + 400 <= 32767
        assert nmismatch(oracle.sgbm(l, r, **prm), cv2.StereoSGBM_create(**prm).compute(l, r)) == 0, prm


def test_live_cv2_contract_corners(oracle, cv2):
    """blockSize 13 / 15 together with MODE_SGBM_3WAY and with 272 .. 512 disparities (the combinations the GPU test
    test_sgbm_mode_block_disparity_combinations compares with the oracle): oracle == cv2."""
    from svgpu.synth import synth_pair
    rng = np.random.default_rng(1)
    for (H, W, D, bs, mode, cap, P1, P2) in [(40, 200, 64, 13, 2, 15, 300, 1200), (38, 420, 320, 15, 2, 15, 100, 400), (30, 600, 512, 15, 3, 15, 100, 400),
                                             (33, 380, 272, 13, 0, 15, 300, 1200), (44, 300, 128, 15, 1, 15, 100, 300), (36, 560, 512, 9, 2, 31, 200, 800)]:
        for kind in ("synth", "noise"):
            if kind == "synth":
                l, r = synth_pair(H, W, D, seed=H)
            else:
                l, r = rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8)
            prm = dict(minDisparity=0, numDisparities=D, blockSize=bs, P1=P1, P2=P2, disp12MaxDiff=1, preFilterCap=cap, uniquenessRatio=10,
                       speckleWindowSize=0, speckleRange=0, mode=mode)
            assert nmismatch(oracle.sgbm(l, r, **prm), cv2.StereoSGBM_create(**prm).compute(l, r)) == 0, prm

"""Calibration products of StereoVision::StereoVision (/root/reference/src/lib.cpp:29-47): cv::stereoRectify and
cv::initUndistortRectifyMap.  CPU part: the oracle restatements and the library's host-side svgpu_stereo_rectify against
the cv2 golden vectors (tests/golden/rectify_small.npz).  GPU part: the maps generated on the device, bit for bit."""
import numpy as np
import pytest

from helpers import golden

REL_TOL = 1e-12   # f64 products: same algorithm, summation order of 3x3 products may differ from OpenCV's


def cases():
    g = golden("rectify_small")
    for n in g["names"]:
        n = str(n)
        yield n, {k: g[f"{n}_{k}"] for k in ("K1", "D1", "K2", "D2", "R", "T", "size", "R1", "R2", "P1", "P2", "Q")}, g


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, float(np.abs(b).max())))


def test_oracle_stereo_rectify_golden():
    from oracle import stereo_rectify as sr
    seen = 0
    for n, c, _ in cases():
        got = sr.stereo_rectify(c["K1"], c["D1"], c["K2"], c["D2"], tuple(int(v) for v in c["size"]), c["R"], c["T"], alpha=0.0)
        for name, a in zip(("R1", "R2", "P1", "P2", "Q"), got):
            assert rel(a, c[name]) <= REL_TOL, (n, name)
        seen += 1
    assert seen >= 6   # four seeded calibrations (one vertical) + the reference's v1 and v2


def test_oracle_rectify_maps_golden(oracle):
    for n, c, g in cases():
        if n.startswith("ref"):
            continue
        W, H = (int(v) for v in c["size"])
        for side, (K, D, R, P) in enumerate(((c["K1"], c["D1"], c["R1"], c["P1"]), (c["K2"], c["D2"], c["R2"], c["P2"]))):
            m1, m2 = oracle.init_undistort_rectify_map(K, D, R, P, (W, H))
            assert np.array_equal(m1, g[f"{n}_map1_{side}"]) and np.array_equal(m2, g[f"{n}_map2_{side}"]), (n, side)


def test_library_stereo_rectify_golden():
    """svgpu_stereo_rectify is host arithmetic inside libsvgpu.so: it runs without a GPU."""
    import svgpu
    for n, c, _ in cases():
        got = svgpu.stereo_rectify(c["K1"], c["D1"], c["K2"], c["D2"], tuple(int(v) for v in c["size"]), c["R"], c["T"])
        for name, a in zip(("R1", "R2", "P1", "P2", "Q"), got):
            assert rel(a, c[name]) <= REL_TOL, (n, name)


def test_library_stereo_rectify_rejects_degenerate_input():
    import svgpu
    K = np.array([[100.0, 0, 50], [0, 100, 40], [0, 0, 1]])
    D = np.zeros(5)
    with pytest.raises(svgpu.SvgpuError):
        svgpu.stereo_rectify(K, D, K, D, (100, 80), np.eye(3), np.zeros(3))          # no baseline
    with pytest.raises(svgpu.SvgpuError):
        svgpu.stereo_rectify(K, D, K, D, (100, 80), np.zeros((3, 3)), np.array([-1.0, 0, 0]))  # singular R
    with pytest.raises(ValueError):
        svgpu.stereo_rectify(K, D[:4], K, D, (100, 80), np.eye(3), np.array([-1.0, 0, 0]))


def test_live_cv2_stereo_rectify(cv2):
    import svgpu
    from oracle import stereo_rectify as sr
    rng = np.random.default_rng(5)
    for _ in range(30):
        # calibrations whose distortion model is invertible over the image (corner radius <= ~1.1 in normalised
        # coordinates, i.e. up to ~95 degrees diagonal field of view); beyond that undistortPoints does not converge
        # and OpenCV's result there is not pinned
        W = int(rng.integers(120, 2000))
        H = int(rng.integers(90, max(W, 91) + 1))
        f = W * rng.uniform(0.65, 1.2)
        K1 = np.array([[f, 0, W / 2 + rng.normal() * 5], [0, f * rng.uniform(0.98, 1.02), H / 2 + rng.normal() * 5], [0, 0, 1]])
        K2 = np.array([[f * rng.uniform(0.98, 1.02), 0, W / 2 + rng.normal() * 5], [0, f * rng.uniform(0.98, 1.02), H / 2 + rng.normal() * 5], [0, 0, 1]])
        D1, D2 = (rng.normal(size=5) * np.array([0.08, 0.03, 0.001, 0.001, 0.005]) for _ in range(2))
        R, _ = cv2.Rodrigues(rng.normal(size=3) * 0.03)
        T = np.array([-rng.uniform(40, 200), rng.normal(), rng.normal()])
        want = cv2.stereoRectify(K1, D1, K2, D2, (W, H), R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0)[:5]
        for got in (sr.stereo_rectify(K1, D1, K2, D2, (W, H), R, T, alpha=0.0), svgpu.stereo_rectify(K1, D1, K2, D2, (W, H), R, T)):
            for name, a, b in zip(("R1", "R2", "P1", "P2", "Q"), got, want):
                assert rel(a, b) <= REL_TOL, name


@pytest.mark.gpu
def test_device_rectify_maps_bit_exact(oracle):
    """k_init_rectify_map against cv2's maps (golden) and, at the reference's calibrations, against the oracle at full size."""
    import svgpu
    for n, c, g in cases():
        W, H = (int(v) for v in c["size"])
        with svgpu.Context(W, H) as ctx:
            for side, (K, D, R, P) in enumerate(((c["K1"], c["D1"], c["R1"], c["P1"]), (c["K2"], c["D2"], c["R2"], c["P2"]))):
                ctx.init_rectify_maps(side, K, D, R, P)
                m1, m2 = ctx.get_rectify_maps(side)
                if n.startswith("ref"):
                    w1, w2 = oracle.init_undistort_rectify_map(K, D, R, P, (W, H))
                else:
                    w1, w2 = g[f"{n}_map1_{side}"], g[f"{n}_map2_{side}"]
                assert np.array_equal(m1, w1) and np.array_equal(m2, w2), (n, side, int((m1 != w1).sum()), int((m2 != w2).sum()))


@pytest.mark.gpu
def test_constructor_path_from_calibration_to_rectified_images(oracle):
    """src/lib.cpp:29-51 end to end on the library: calibration -> stereo_rectify -> device maps -> remap, against the
    oracle's remap with cv2's maps."""
    import svgpu
    g = golden("rectify_small")
    n = "c1"
    c = {k: g[f"{n}_{k}"] for k in ("K1", "D1", "K2", "D2", "R", "T", "size")}
    W, H = (int(v) for v in c["size"])
    R1, R2, P1, P2, Q = svgpu.stereo_rectify(c["K1"], c["D1"], c["K2"], c["D2"], (W, H), c["R"], c["T"])
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    with svgpu.Context(W, H) as ctx:
        for side, (K, D, R, P) in enumerate(((c["K1"], c["D1"], R1, P1), (c["K2"], c["D2"], R2, P2))):
            ctx.init_rectify_maps(side, K, D, R, P)
            got = ctx.remap(side, img)
            want = oracle.remap(img, g[f"{n}_map1_{side}"], g[f"{n}_map2_{side}"])
            assert np.array_equal(got, want), side

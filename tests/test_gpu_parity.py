"""GPU parity tests: the CUDA path, called through the C ABI (include/svgpu.h via the ctypes
binding), against the CPU oracle and the committed cv2 golden vectors.  Integer outputs must be
bit-exact; float outputs are compared bit-for-bit too (stated tolerance: 0 ulp; the north star
allows 1e-4 relative)."""
import hashlib

import numpy as np
import pytest

from helpers import REF_PARAMS, bm_golden_cases, golden, nmismatch, sgbm_golden_cases, sgbm_wide_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def svgpu():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import svgpu as m
    m._abi.load()
    return m


def test_sgbm_golden_all_modes(svgpu):
    for i, l, r, prm, want in sgbm_golden_cases():
        with svgpu.Context(l.shape[1], l.shape[0]) as ctx:
            ctx.set_sgbm(**prm)
            assert nmismatch(ctx.sgbm(l, r), want) == 0, (i, prm)


def test_sgbm_wide_disparity_golden(svgpu, oracle):
    """numDisparities 272 .. 512 (a full warp per pixel: only the line form of the aggregation serves them, pair by pair) against
    cv2's outputs: gray in the three modes, 8UC3, blockSize 11 / 13; then a batch of three pairs == single calls."""
    n = 0
    for i, l, r, prm, want in sgbm_wide_cases():
        H, W = l.shape[:2]
        with svgpu.Context(W, H) as ctx:
            ctx.set_sgbm(**prm)
            assert nmismatch(ctx.sgbm(l, r), want) == 0, (i, prm)
        n += 1
    assert n >= 20
    from svgpu.synth import synth_batch
    H, W, D = 30, 600, 512
    L, R = synth_batch(3, H, W, D, seed0=3)
    prm = dict(REF_PARAMS, numDisparities=D, mode=1)
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(**prm)
        got = ctx.sgbm_batch(L, R)
        for k in range(3):
            assert nmismatch(got[k], ctx.sgbm(L[k], R[k])) == 0
        assert nmismatch(got[0], oracle.sgbm(L[0], R[0], **prm)) == 0


def test_sgbm_three_channel_golden(svgpu, oracle):
    """cv::StereoSGBM::compute on 8UC3 pairs (svgpu_pipeline with SV_STAGE_SGBM alone): bit-exact against cv2's output; a
    block size whose three-channel cost could leave the 16-bit range is refused."""
    from helpers import sgbm3_golden_cases
    for i, l, r, prm, want in sgbm3_golden_cases():
        with svgpu.Context(l.shape[1], l.shape[0]) as ctx:
            ctx.set_sgbm(**prm)
            assert nmismatch(ctx.sgbm(l, r), want) == 0, (i, prm)
    # the shared-memory-ring build of the three-channel cost kernel (blockSize <= 7, D >= 128) and the register-ring build
    # (blockSize 9) against the oracle, which the golden cases above pin to cv2
    rng = np.random.default_rng(11)
    for (H, W, D, bs, mode) in [(37, 330, 128, 5, 1), (33, 300, 128, 9, 0), (30, 420, 256, 3, 3), (41, 250, 160, 7, 1)]:
        l3 = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        r3 = np.roll(l3, -9, axis=1)
        r3[::3] = rng.integers(0, 256, r3[::3].shape, dtype=np.uint8)
        prm = dict(REF_PARAMS, numDisparities=D, blockSize=bs, P1=8 * bs * bs, P2=32 * bs * bs, mode=mode,
                   preFilterCap=15 if bs == 9 else 32)   # three channels at blockSize 9 only fit 16 bits with a small cap
        with svgpu.Context(W, H) as ctx:
            ctx.set_sgbm(**prm)
            assert nmismatch(ctx.sgbm(l3, r3), oracle.sgbm(l3, r3, **prm)) == 0, (H, W, D, bs, mode)
    l = np.zeros((40, 200, 3), np.uint8)
    with svgpu.Context(200, 40) as ctx:
        ctx.set_sgbm(**dict(REF_PARAMS, numDisparities=32, blockSize=11, P1=8, P2=32, mode=1))
        with pytest.raises(svgpu.SvgpuError) as e:
            ctx.sgbm(l, l)
        assert e.value.code == svgpu._abi.SV_INVALID_SETTINGS


@pytest.mark.parametrize("lines", ["0", "1"])
def test_sgbm_3way_golden(svgpu, oracle, lines, monkeypatch):
    """MODE_SGBM_3WAY (four stripe frames per pair, forward H+V sweep, backward H sweep, OpenCV's vectorised selection rule):
    bit-exact against cv2's output incl. noise pairs (ties) and images smaller than the stripe overlap; batch == single; an
    8UC3 pair against the oracle; uniquenessRatio = 100 is refused (cv2 dies with SIGFPE there).  Once through the strip
    pipeline and once through the line form (three directions per stripe frame, the same selection rule)."""
    from helpers import sgbm3way_golden_cases
    monkeypatch.setenv("SVGPU_SWEEP_LINES", lines)
    for i, l, r, prm, want in sgbm3way_golden_cases():
        with svgpu.Context(l.shape[1], l.shape[0]) as ctx:
            ctx.set_sgbm(**prm)
            assert nmismatch(ctx.sgbm(l, r), want) == 0, (i, prm)
            if i % 5 == 0:
                got = ctx.sgbm_batch(np.stack([l, r, l]), np.stack([r, l, r]))
                assert nmismatch(got[0], want) == 0 and nmismatch(got[2], want) == 0, (i, prm)
    rng = np.random.default_rng(3)
    l3 = rng.integers(0, 256, (45, 150, 3), dtype=np.uint8)
    r3 = np.roll(l3, -7, axis=1)
    prm = dict(REF_PARAMS, numDisparities=32, blockSize=5, P1=200, P2=800, mode=2)
    with svgpu.Context(150, 45) as ctx:
        ctx.set_sgbm(**prm)
        assert nmismatch(ctx.sgbm(l3, r3), oracle.sgbm(l3, r3, **prm)) == 0
        with pytest.raises(svgpu.SvgpuError):
            ctx.set_sgbm(**dict(prm, uniquenessRatio=100))


def test_sgbm_too_narrow_matches_reference_error(svgpu):
    l = np.zeros((21, 40), np.uint8)
    with svgpu.Context(40, 21) as ctx:
        for mode in (0, 1):
            ctx.set_sgbm(numDisparities=48, mode=mode, **REF_PARAMS)
            with pytest.raises(svgpu.SvgpuError) as e:
                ctx.sgbm(l, l)
            assert e.value.code == svgpu._abi.SV_INVALID_SETTINGS
        ctx.set_sgbm(numDisparities=48, mode=3, **REF_PARAMS)
        assert (ctx.sgbm(l, l) == -16).all()


def test_invalid_settings_and_images(svgpu):
    with svgpu.Context(64, 48) as ctx:
        for bad in (dict(numDisparities=24), dict(numDisparities=528), dict(numDisparities=32, mode=4), dict(numDisparities=32, mode=2, uniquenessRatio=100),
                    dict(numDisparities=32, blockSize=4), dict(numDisparities=32, P2=30000)):
            prm = dict(REF_PARAMS, mode=1)
            prm.update(bad)
            with pytest.raises(svgpu.SvgpuError) as e:
                ctx.set_sgbm(**prm)
            assert e.value.code == svgpu._abi.SV_INVALID_SETTINGS
        with pytest.raises(svgpu.SvgpuError) as e:  # stage before its parameters
            ctx.sgbm(np.zeros((48, 64), np.uint8), np.zeros((48, 64), np.uint8))
        assert e.value.code == svgpu._abi.SV_NOT_CONFIGURED
        with pytest.raises(svgpu.SvgpuError) as e:  # empty image -> kInvalidImage (src/lib.cpp:54-56)
            ctx.remap(0, np.zeros((0, 0, 3), np.uint8))
        assert e.value.code == svgpu._abi.SV_INVALID_IMAGE


def test_bm_golden(svgpu):
    for i, l, r, prm, want in bm_golden_cases():
        with svgpu.Context(l.shape[1], l.shape[0]) as ctx:
            ctx.set_bm(**prm)
            assert nmismatch(ctx.bm(l, r), want) == 0, (i, prm)


def test_bm_min_disparity(svgpu, oracle):
    """StereoBM with minDisparity != 0 (positive, negative, ranges entirely below zero): equal to the oracle everywhere and to
    cv2 inside cv::getValidDisparityROI."""
    from helpers import bm_mind_golden_cases
    for i, l, r, prm, want, (y0, y1, x0, x1) in bm_mind_golden_cases():
        with svgpu.Context(l.shape[1], l.shape[0]) as ctx:
            ctx.set_bm(**prm)
            got = ctx.bm(l, r)
        assert nmismatch(got, oracle.bm(l, r, **prm)) == 0, (i, prm)
        assert nmismatch(got[y0:y1, x0:x1], want[y0:y1, x0:x1]) == 0, (i, prm)


def test_stages_golden(svgpu):
    g = golden("stages_small")
    with svgpu.Context(53, 37) as ctx:
        assert nmismatch(ctx.median3(g["median_in"]), g["median_out"]) == 0
    with svgpu.Context(160, 120) as ctx:
        for i, (nv, ms, md) in enumerate(g["spk_params"]):
            assert nmismatch(ctx.filter_speckles(g["spk_in"], nv, ms, md), g[f"spk_out{i}"]) == 0
        assert nmismatch(ctx.morph_close5(g["spk_in"]), g["close_out"]) == 0
        ctx.set_rectify_maps(0, g["map1_l"], g["map2_l"])
        ctx.set_rectify_maps(1, g["map1_r"], g["map2_r"])
        assert nmismatch(ctx.remap(0, g["remap_in"]), g["remap_out_l"]) == 0  # stated tolerance <= 1 LSB; achieved 0
        assert nmismatch(ctx.remap(1, g["remap_in"][:, :, 0].copy()), g["remap_out_r_gray"]) == 0
        ctx.set_Q(g["rect_Q"])
        for hm, key in ((True, "reproj_out_hm"), (False, "reproj_out")):
            got = ctx.reproject(g["reproj_in"], hm)
            assert nmismatch(got.view(np.uint32), g[key].view(np.uint32)) == 0
        assert nmismatch(ctx.prefilter_chain(g["remap_in"]), g["chain_sharp"]) == 0
        assert nmismatch(ctx.resize(g["resize_in"]), g["resize_out"]) == 0


def test_bottle_pair_full_chain(svgpu, oracle):
    """BASELINE config 1 on the GPU: scaled frames -> rectify -> blur/gray/sharpen -> SGBM HH4 D=224 ->
    filterSpeckles -> close -> XYZ + depth, one svgpu_pipeline call (src/lib.cpp:53-85)."""
    g = golden("bottle")
    size = (640, 480)
    ml = oracle.init_undistort_rectify_map(g["cal_Kl"], g["cal_Dl"], g["rect_Rl"], g["rect_Pl"], size)
    mr = oracle.init_undistort_rectify_map(g["cal_Kr"], g["cal_Dr"], g["rect_Rr"], g["rect_Pr"], size)
    with svgpu.Context(*size) as ctx:
        ctx.set_rectify_maps(0, *ml)
        ctx.set_rectify_maps(1, *mr)
        ctx.set_Q(g["Q"])
        ctx.set_depth_consts(float(g["baseline"]), float(g["focal"]))
        ctx.set_sgbm(**svgpu.REFERENCE_SGBM)
        raw = ctx.pipeline(g["scaled_left"], g["scaled_right"], svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM,
                           want_rect=True)
        assert nmismatch(raw["left_rect"], g["rect_left"]) == 0 and nmismatch(raw["right_rect"], g["rect_right"]) == 0
        assert nmismatch(raw["disp"], g["disp_raw"]) == 0
        assert hashlib.sha256(np.ascontiguousarray(raw["disp"]).tobytes()).hexdigest()[:16] == "90f4df78b9ec67a4"
        full = ctx.pipeline(g["scaled_left"], g["scaled_right"],
                            svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM | svgpu.STAGE_POSTFILTER |
                            svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH, want_xyz=True, want_depth=True)
        assert nmismatch(full["disp"], g["disp_post"]) == 0
        assert hashlib.sha256(full["xyz"].tobytes()).digest() == g["xyz_sha256"].tobytes()
        assert hashlib.sha256(full["depth"].tobytes()).digest() == g["depth_sha256"].tobytes()


def test_device_memory_and_batch(svgpu, oracle):
    import torch
    from svgpu.synth import synth_batch
    H, W, D, n = 60, 256, 64, 5
    L, R = synth_batch(n, H, W, D, seed0=40)
    prm = dict(REF_PARAMS, numDisparities=D, mode=1)
    want = np.stack([oracle.sgbm(L[i], R[i], **prm) for i in range(n)])
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(**prm)
        ctx.reserve(3)  # forces two chunks (3 + 2)
        assert nmismatch(ctx.sgbm_batch(L, R), want) == 0  # host buffers
        tl, tr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        td = ctx.sgbm_batch(tl, tr)
        torch.cuda.synchronize()
        assert nmismatch(td.cpu().numpy(), want) == 0
        # streaming entry point: three overlapping enqueues on rotating inputs, results read after sync()
        outs = [np.zeros((3, H, W), np.int16) for _ in range(3)]
        ins = [(np.ascontiguousarray(L[k:k + 3]), np.ascontiguousarray(R[k:k + 3])) for k in range(3)]
        for k in range(3):
            ctx.sgbm_enqueue(ins[k][0], ins[k][1], outs[k])
        ctx.sync()
        for k in range(3):
            assert nmismatch(outs[k], want[k:k + 3]) == 0, k
        one = ctx.sgbm(tl[2], tr[2])
        torch.cuda.synchronize()
        assert nmismatch(one.cpu().numpy(), want[2]) == 0
        assert ctx.launch_count > 0


@pytest.mark.parametrize("mode", [0, 1, 3])
def test_config2_720p_64_bit_exact(svgpu, oracle, mode):
    """BASELINE config 2 (1280x720, 64 disparities): SGBM 5-path bit-exact; 8-path and HH4 ride along."""
    from svgpu.synth import synth_pair
    l, r = synth_pair(720, 1280, 64, seed=2)
    prm = dict(REF_PARAMS, numDisparities=64, mode=mode)
    want = oracle.sgbm(l, r, **prm)
    with svgpu.Context(1280, 720) as ctx:
        ctx.set_sgbm(**prm)
        assert nmismatch(ctx.sgbm(l, r), want) == 0
    assert (want[:, 64:] >= 0).mean() > 0.9


def test_config2_720p_64_stereobm(svgpu, oracle):
    from svgpu.synth import synth_pair
    l, r = synth_pair(720, 1280, 64, seed=2)
    with svgpu.Context(1280, 720) as ctx:
        ctx.set_bm(numDisparities=64, blockSize=21)
        assert nmismatch(ctx.bm(l, r), oracle.bm(l, r, numDisparities=64, blockSize=21)) == 0


def test_config3_1080p_128_hh_bit_exact_and_properties(svgpu, oracle):
    """BASELINE config 3 (1920x1080, 128 disparities, MODE_HH + LR check): one pair against the oracle at
    full size, plus size-independent properties on a batch: determinism, batch == single, and saturation
    behaviour on uncorrelated noise (every pixel whose S saturates is invalid)."""
    import torch
    from svgpu.synth import synth_batch
    H, W, D = 1080, 1920, 128
    L, R = synth_batch(3, H, W, D, seed0=10)
    prm = dict(REF_PARAMS, numDisparities=D, mode=1)
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(**prm)
        ctx.reserve(3)
        tl, tr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        a = ctx.sgbm_batch(tl, tr).cpu().numpy()
        b = ctx.sgbm_batch(tl, tr).cpu().numpy()
        assert nmismatch(a, b) == 0
        single = ctx.sgbm(L[1], R[1])
        assert nmismatch(single, a[1]) == 0
        assert nmismatch(a[0], oracle.sgbm(L[0], R[0], **prm)) == 0
        assert (a[:, :, :D] == -16).all() and (a[:, :, D:] >= 0).mean() > 0.99
        rng = np.random.default_rng(5)
        nl, nr = rng.integers(0, 256, (H, W), dtype=np.uint8), rng.integers(0, 256, (H, W), dtype=np.uint8)
        dn = ctx.sgbm(nl, nr)
        assert (dn >= 0).mean() < 0.2  # noise: S saturates almost everywhere -> invalid (A.6)


def test_config4_4k_256_pipeline_matches_oracle_rows(svgpu, oracle):
    """BASELINE config 4 (3840x2160, 256 disparities, rectify -> SGBM -> reprojectImageTo3D). The oracle is
    too slow for the full 4K volume in a unit test, so the disparity is checked against the oracle on a
    full-width horizontal band (rows are only coupled through the box window and the vertical/diagonal
    paths, so MODE_HH4 restricted to a band is not comparable; instead the band is run as its own image on
    both sides), and remap / reproject are checked at full 4K."""
    from svgpu.synth import synth_pair
    H, W, D = 2160, 3840, 256
    l, r = synth_pair(H, W, D, seed=4)
    prm = dict(REF_PARAMS, numDisparities=D, mode=1)
    band = slice(1000, 1048)
    want = oracle.sgbm(np.ascontiguousarray(l[band]), np.ascontiguousarray(r[band]), **prm)
    with svgpu.Context(W, 48) as ctx:
        ctx.set_sgbm(**prm)
        assert nmismatch(ctx.sgbm(np.ascontiguousarray(l[band]), np.ascontiguousarray(r[band])), want) == 0
    # full 4K: rectify + SGBM + reproject run end to end; remap and reproject are compared at full size
    K = np.array([[833.77 * 6, 0, 345.2 * 6], [0, 832.87 * 6, 283.1 * 4.5], [0, 0, 1]])
    dist = np.array([-0.4168, 0.166, 0, 0, 0])
    P = np.array([[K[0, 0] * 0.9, 0, K[0, 2], 0], [0, K[0, 0] * 0.9, K[1, 2], 0], [0, 0, 1, 0]])
    m1, m2 = oracle.init_undistort_rectify_map(K, dist, np.eye(3), P, (W, H))
    Q = np.array([[1, 0, 0, -K[0, 2]], [0, 1, 0, -K[1, 2]], [0, 0, 0, P[0, 0]], [0, 0, 1 / 75.3, 0]])
    with svgpu.Context(W, H) as ctx:
        ctx.set_rectify_maps(0, m1, m2)
        ctx.set_rectify_maps(1, m1, m2)
        ctx.set_Q(Q)
        ctx.set_sgbm(**prm)
        out = ctx.pipeline(l, r, svgpu.STAGE_REMAP | svgpu.STAGE_SGBM | svgpu.STAGE_REPROJECT, want_rect=True, want_xyz=True)
        assert nmismatch(out["left_rect"], oracle.remap(l, m1, m2)) == 0
        xyz = oracle.reproject(out["disp"], Q, True)
        assert nmismatch(out["xyz"].view(np.uint32), xyz.view(np.uint32)) == 0
        rows = slice(1200, 1204)
        assert (out["disp"][rows, D + 100:-100] >= 0).mean() > 0.5


def test_cpp_facade_matches_python_pipeline(svgpu, oracle, tmp_path):
    """The C++ facade with the reference's public API (fhgr-stereo-vision_b200/cpp, stereo_vision::StereoVision::
    AnalyzeAndAnnotateImage, src/lib.cpp:53-85) on the bottle pair: same bytes as the ctypes path that the test above
    pins to the cv2 golden vectors, and the reference's own unit-test contract (empty image -> kInvalidImage)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fhgr-stereo-vision_b200", "cpp", "facade_test")
    assert os.path.exists(exe), "build() compiles the facade"
    g = golden("bottle")
    size = (640, 480)
    ml = oracle.init_undistort_rectify_map(g["cal_Kl"], g["cal_Dl"], g["rect_Rl"], g["rect_Pl"], size)
    mr = oracle.init_undistort_rectify_map(g["cal_Kr"], g["cal_Dr"], g["rect_Rr"], g["rect_Pr"], size)
    p0, p1 = (200, 300), (240, 420)  # (y, x)
    case = tmp_path / "case.bin"
    with open(case, "wb") as f:
        f.write(np.array([640, 480, p0[0] * 640 + p0[1], p1[0] * 640 + p1[1]], np.int32).tobytes())
        for a in (ml[0], ml[1], mr[0], mr[1]):
            f.write(np.ascontiguousarray(a).tobytes())
        f.write(np.ascontiguousarray(g["Q"], np.float64).tobytes())
        f.write(np.array([float(g["baseline"]), float(g["focal"])], np.float32).tobytes())
        f.write(np.ascontiguousarray(g["scaled_left"]).tobytes())
        f.write(np.ascontiguousarray(g["scaled_right"]).tobytes())
    out = subprocess.run([exe, str(case)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    keys = ("rect_left", "rect_right", "disp", "depth", "p0", "p1")
    got = dict(line.split(" ", 1) for line in out.stdout.splitlines() if line.split(" ", 1)[0] in keys)
    # the same instance constructed from the calibration alone (stereoRectify on the host, maps generated on the
    # device: all of src/lib.cpp:29-47) must produce the same bytes as the one fed with OpenCV's products
    calib = tmp_path / "calib.bin"
    with open(calib, "wb") as f:
        f.write(np.array([640, 480], np.int32).tobytes())
        for k in ("cal_Kl", "cal_Dl", "cal_Kr", "cal_Dr", "cal_R", "cal_T"):
            a = np.ascontiguousarray(g[k], np.float64).reshape(-1)
            f.write(a[:5].tobytes() if k in ("cal_Dl", "cal_Dr") else a.tobytes())
    out2 = subprocess.run([exe, str(case), str(calib)], capture_output=True, text=True, timeout=120)
    assert out2.returncode == 0, out2.stdout + out2.stderr
    got2 = dict(line.split(" ", 1) for line in out2.stdout.splitlines() if line.split(" ", 1)[0] in keys)
    for k in ("rect_left", "rect_right", "disp", "depth"):
        assert got2[k] == got[k], k
    for k in ("p0", "p1"):
        assert np.allclose(np.array(got2[k].split(), np.float32), np.array(got[k].split(), np.float32), rtol=1e-6, atol=0), k

    def fnv(a):
        h = 1469598103934665603
        for b in np.ascontiguousarray(a).tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    with svgpu.Context(*size) as ctx:
        ctx.set_rectify_maps(0, *ml)
        ctx.set_rectify_maps(1, *mr)
        ctx.set_Q(g["Q"])
        ctx.set_depth_consts(float(g["baseline"]), float(g["focal"]))
        ctx.set_sgbm(**svgpu.REFERENCE_SGBM)
        full = ctx.pipeline(g["scaled_left"], g["scaled_right"],
                            svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM | svgpu.STAGE_POSTFILTER |
                            svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH, want_rect=True, want_xyz=True, want_depth=True)
    assert nmismatch(full["disp"], g["disp_post"]) == 0
    # FNV over a few hundred KB in pure Python is slow; compare the small outputs fully and the images by a strided sample
    assert int(got["disp"]) == fnv(full["disp"])
    assert int(got["depth"]) == fnv(full["depth"])
    assert int(got["rect_left"]) == fnv(full["left_rect"]) and int(got["rect_right"]) == fnv(full["right_rect"])
    for key, p in (("p0", p0), ("p1", p1)):
        want = full["xyz"][p[0], p[1]]
        assert np.allclose(np.array(got[key].split(), np.float32), want, rtol=1e-6, atol=0), (key, got[key], want)


@pytest.mark.parametrize("sweep_k,sweep_split,lines", [("4", "0", "0"), ("2", "0", "0"), ("2", "1", "0"), ("4", "1", "0"), ("4", "0", "1")])
def test_sgbm_random_configs_bit_exact(svgpu, oracle, sweep_k, sweep_split, lines, monkeypatch):
    """Seeded random sweep over image sizes (incl. fewer computed columns than one strip and single-row images),
    disparity ranges, block sizes, penalties, uniqueness ratios, minDisparity signs and all three modes, on textured
    and on saturating noise inputs: bit-exact against the oracle.  Run once per variant of the aggregation: strips of 4 columns
    (throughput) and of 2 (which the library picks by itself for a few pairs per call), the two passes one after the other
    (the second selects in registers) and side by side on two streams (k_wta_sum adds the two sums and selects), and the
    line form of single small pairs (sgbm_lines.cu: one worker per line and direction)."""
    from svgpu.synth import synth_pair
    monkeypatch.setenv("SVGPU_SWEEP_K", sweep_k)
    monkeypatch.setenv("SVGPU_SWEEP_SPLIT", sweep_split)
    monkeypatch.setenv("SVGPU_SWEEP_LINES", lines)
    rng = np.random.default_rng(20261018)
    bad = []
    for case in range(48):
        D = int(rng.choice([16, 32, 48, 64, 80, 96, 128, 160, 224, 256, 320, 512]))
        minD = int(rng.choice([0, 0, 0, 3, -5, 16, -16]))
        bs = int(rng.choice([1, 3, 5, 7, 9, 11]))
        mode = int(rng.choice([0, 1, 3]))
        W1 = int(rng.choice([1, 2, 3, 5, 17, 40, 97, 150, 260]))
        W = max(minD + D, 0) - min(minD, 0) + W1
        H = int(rng.choice([1, 2, 5, 9, 33, 70]))
        if mode != 3 and W - (minD + D) <= bs // 2:  # the reference raises for this combination (tested elsewhere)
            W += bs // 2 + 1
        cap = int(rng.choice([1, 15, 31, 32, 63]))
        P1 = int(rng.integers(1, 400))
        P2 = P1 + int(rng.integers(1, 1200))
        prm = dict(minDisparity=minD, numDisparities=D, blockSize=bs, P1=P1, P2=P2, disp12MaxDiff=int(rng.choice([-1, 0, 1, 4])),
                   preFilterCap=cap, uniquenessRatio=int(rng.choice([0, 5, 15, 40, 100])), speckleWindowSize=0, speckleRange=0, mode=mode)
        if bs * bs * (2 * (max(cap, 15) | 1) + 63) + P2 > 32767:
            continue  # outside the 16-bit contract (rejected by svgpu_set_sgbm, as by this implementation's header)
        if rng.random() < 0.3:
            l = rng.integers(0, 256, (H, W), dtype=np.uint8)
            r = rng.integers(0, 256, (H, W), dtype=np.uint8)
        else:
            l, r = synth_pair(H, W, max(min(D, W // 2), 1), seed=case)
        want = oracle.sgbm(l, r, **prm)
        with svgpu.Context(W, H) as ctx:
            ctx.set_sgbm(**prm)
            got = ctx.sgbm(l, r)
        if nmismatch(got, want):
            bad.append((case, H, W, prm, nmismatch(got, want)))
    assert not bad, bad


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()


def _big(key):
    return golden("big_sha")[key].tobytes()


def _big_calib(W, H):
    K = np.array([[833.77 * 6, 0, 345.2 * 6], [0, 832.87 * 6, 283.1 * 4.5], [0, 0, 1]])
    dist = np.array([-0.4168, 0.166, 0, 0, 0])
    P = np.array([[K[0, 0] * 0.9, 0, K[0, 2], 0], [0, K[0, 0] * 0.9, K[1, 2], 0], [0, 0, 1, 0]])
    Q = np.array([[1, 0, 0, -K[0, 2]], [0, 1, 0, -K[1, 2]], [0, 0, 0, P[0, 0]], [0, 0, 1 / 75.3, 0]])
    return K, dist, P, Q


def test_config4_4k_256_full_frame_sha_vs_cv2(svgpu, oracle):
    """BASELINE config 4 at full size against cv2's own output (tests/golden/big_sha.npz: sha256 digests written by
    make_golden.py): 3840x2160, 256 disparities, MODE_HH — (a) svgpu_sgbm_compute on the gray pair, (b) svgpu_pipeline
    remap -> SGBM -> reprojectImageTo3D on the same pair, (c) a BGR pair through the reference's whole chain
    (src/lib.cpp:53-85).  Pins the 2160-row vertical / diagonal paths, the 16-lanes-per-pixel reduction and the strip-group
    hand-over at D = 256."""
    from svgpu.synth import synth_pair
    H, W, D = 2160, 3840, 256
    l, r = synth_pair(H, W, D, seed=4)
    assert _sha(np.stack([l, r])) == _big("c4_in"), "synthetic input differs from the one the golden digests were made from"
    prm = dict(REF_PARAMS, numDisparities=D, mode=1)
    K, dist, P, Q = _big_calib(W, H)
    m1, m2 = oracle.init_undistort_rectify_map(K, dist, np.eye(3), P, (W, H))
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(**prm)
        assert _sha(ctx.sgbm(l, r)) == _big("c4_disp")
        ctx.set_rectify_maps(0, m1, m2)
        ctx.set_rectify_maps(1, m1, m2)
        ctx.set_Q(Q)
        ctx.set_depth_consts(75.3, float(K[0, 0]))
        out = ctx.pipeline(l, r, svgpu.STAGE_REMAP | svgpu.STAGE_SGBM | svgpu.STAGE_REPROJECT, want_rect=True, want_xyz=True)
        assert _sha(np.stack([out["left_rect"], out["right_rect"]])) == _big("c4_rect")
        assert _sha(out["disp"]) == _big("c4_rect_disp")
        assert _sha(out["xyz"]) == _big("c4_rect_xyz")
        lb = np.stack([l, np.roll(l, 1, axis=0) // 2 + 30, 255 - l], axis=2).astype(np.uint8)
        rb = np.stack([r, np.roll(r, 1, axis=0) // 2 + 30, 255 - r], axis=2).astype(np.uint8)
        assert _sha(np.stack([lb, rb])) == _big("c4bgr_in")
        raw = ctx.pipeline(lb, rb, svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM)
        assert _sha(raw["disp"]) == _big("c4bgr_raw")
        full = ctx.pipeline(lb, rb, svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM | svgpu.STAGE_POSTFILTER |
                            svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH, want_xyz=True, want_depth=True)
        assert _sha(full["disp"]) == _big("c4bgr_post")
        assert _sha(full["xyz"]) == _big("c4bgr_xyz")
        assert _sha(full["depth"]) == _big("c4bgr_depth")


def test_config3_1080p_speckle_batch_and_streaming_vs_cv2(svgpu, oracle):
    """1920x1080 / 128 / MODE_HH: (a) the speckle filter inside StereoSGBM::compute (speckleWindowSize=100, speckleRange=2)
    on a batch of three pairs, against cv2's digests and (pair 0) the oracle; (b) the streaming entry point with three
    overlapping enqueues of two pairs each, results read after sync()."""
    from svgpu.synth import synth_batch
    H, W, D = 1080, 1920, 128
    L, R = synth_batch(3, H, W, D, seed0=10)
    for i in range(3):
        assert _sha(np.stack([L[i], R[i]])) == _big(f"c3spk_in{i}")
    prm = dict(REF_PARAMS, numDisparities=D, mode=1)
    spk = dict(prm, speckleWindowSize=100, speckleRange=2)
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(**spk)
        ctx.reserve(3)
        got = ctx.sgbm_batch(L, R)
        for i in range(3):
            assert _sha(got[i]) == _big(f"c3spk_disp{i}"), i
        assert nmismatch(got[0], oracle.sgbm(L[0], R[0], **spk)) == 0
        # streaming: batches (0,1), (1,2), (2,0) in flight together
        ctx.set_sgbm(**prm)
        ctx.reserve(2)
        order = [(0, 1), (1, 2), (2, 0)]
        ins = [(np.ascontiguousarray(L[list(o)]), np.ascontiguousarray(R[list(o)])) for o in order]
        outs = [np.zeros((2, H, W), np.int16) for _ in order]
        for k in range(3):
            ctx.sgbm_enqueue(ins[k][0], ins[k][1], outs[k])
        ctx.sync()
        for k, o in enumerate(order):
            for q, i in enumerate(o):
                assert _sha(outs[k][q]) == _big(f"c3_disp{i}"), (k, q)


def test_reconfigure_same_cell_count_reallocates(svgpu, oracle):
    """One context reconfigured between settings whose W1 * D product is equal but whose selection / hand-over buffers are
    not (W = 320: D = 256 -> W1 = 64, D = 64 -> W1 = 256), and between MODE_SGBM_3WAY and MODE_HH: same output as a fresh
    context (the scratch is keyed on every buffer's size)."""
    from svgpu.synth import synth_pair
    H, W = 40, 320
    l, r = synth_pair(H, W, 48, seed=7)
    seq = [dict(REF_PARAMS, numDisparities=256, mode=1), dict(REF_PARAMS, numDisparities=64, mode=1),
           dict(REF_PARAMS, numDisparities=64, mode=2), dict(REF_PARAMS, numDisparities=128, mode=0),
           dict(REF_PARAMS, numDisparities=64, mode=3), dict(REF_PARAMS, numDisparities=256, mode=1)]
    with svgpu.Context(W, H) as ctx:
        for prm in seq:
            ctx.set_sgbm(**prm)
            got = ctx.sgbm(l, r)
            with svgpu.Context(W, H) as fresh:
                fresh.set_sgbm(**prm)
                want = fresh.sgbm(l, r)
            assert nmismatch(got, want) == 0, prm
            assert nmismatch(got, oracle.sgbm(l, r, **prm)) == 0, prm


def test_pipeline_batch_and_streaming_match_single_calls(svgpu, oracle):
    """svgpu_pipeline_batch / svgpu_pipeline_enqueue_batch (the folder and livecam loops, static-image-main.cpp:107-127,
    livecam-main.cpp:70-106, n frames per call): every output of every frame equals the single-frame svgpu_pipeline call, which
    the bottle test pins to cv2 — host buffers in two chunks, device tensors, and three overlapping streamed batches."""
    import torch
    g = golden("bottle")
    size = (640, 480)
    ml = oracle.init_undistort_rectify_map(g["cal_Kl"], g["cal_Dl"], g["rect_Rl"], g["rect_Pl"], size)
    mr = oracle.init_undistort_rectify_map(g["cal_Kr"], g["cal_Dr"], g["rect_Rr"], g["rect_Pr"], size)
    a, b = np.ascontiguousarray(g["scaled_left"]), np.ascontiguousarray(g["scaled_right"])
    L = np.stack([a, a[::-1].copy(), np.roll(a, 40, axis=1), a])
    R = np.stack([b, b[::-1].copy(), np.roll(b, 40, axis=1), np.roll(b, 3, axis=0)])
    ALL = (svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM | svgpu.STAGE_POSTFILTER | svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH)
    keys = ("left_rect", "right_rect", "disp", "xyz", "depth")
    with svgpu.Context(*size) as ctx:
        ctx.set_rectify_maps(0, *ml)
        ctx.set_rectify_maps(1, *mr)
        ctx.set_Q(g["Q"])
        ctx.set_depth_consts(float(g["baseline"]), float(g["focal"]))
        ctx.set_sgbm(**svgpu.REFERENCE_SGBM)
        single = [ctx.pipeline(L[i], R[i], ALL, want_rect=True, want_xyz=True, want_depth=True) for i in range(4)]
        assert nmismatch(single[0]["disp"], g["disp_post"]) == 0

        def check(out, idx, tag):
            for q, i in enumerate(idx):
                for k in keys:
                    got = out[k][q].cpu().numpy() if hasattr(out[k], "cpu") else out[k][q]
                    assert nmismatch(got.view(np.uint8), np.ascontiguousarray(single[i][k]).view(np.uint8)) == 0, (tag, i, k)

        ctx.reserve(3)  # 4 frames = two chunks (3 + 1)
        check(ctx.pipeline_batch(L, R, ALL, want_rect=True, want_xyz=True, want_depth=True), range(4), "host")
        tl, tr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
        out = ctx.pipeline_batch(tl, tr, ALL, want_rect=True, want_xyz=True, want_depth=True)
        torch.cuda.synchronize()
        check(out, range(4), "device")
        # SGBM on 8UC3 frames without the pre-filter chain, batched == single
        raw1 = ctx.pipeline(L[1], R[1], svgpu.STAGE_REMAP | svgpu.STAGE_SGBM)["disp"]
        rawb = ctx.pipeline_batch(L[:2], R[:2], svgpu.STAGE_REMAP | svgpu.STAGE_SGBM)["disp"]
        assert nmismatch(rawb[1], raw1) == 0
        # post-filter + reprojection + depth on caller-supplied disparities (no SGBM stage), batched == single
        dsp = np.stack([single[i]["disp"] for i in range(4)])
        raw_in = np.stack([ctx.pipeline(L[i], R[i], svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM)["disp"] for i in range(4)])
        post = ctx.pipeline_batch(L, R, svgpu.STAGE_POSTFILTER | svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH, want_xyz=True, want_depth=True,
                                  disp=raw_in.copy())
        for i in range(4):
            assert nmismatch(post["disp"][i], dsp[i]) == 0, i
            assert nmismatch(post["xyz"][i].view(np.uint32), single[i]["xyz"].view(np.uint32)) == 0, i
            assert nmismatch(post["depth"][i].view(np.uint32), single[i]["depth"].view(np.uint32)) == 0, i
        # streaming: three batches in flight, results read after sync()
        ctx.use_own_stream()
        order = [(0, 1, 2), (3, 0), (2, 3, 1)]
        ins = [(np.ascontiguousarray(L[list(o)]), np.ascontiguousarray(R[list(o)])) for o in order]
        outs = []
        for (li, ri), o in zip(ins, order):
            n = len(o)
            pre = dict(disp=np.zeros((n, 480, 640), np.int16), xyz=np.zeros((n, 480, 640, 3), np.float32), depth=np.zeros((n, 480, 640), np.float32),
                       left_rect=np.zeros((n, 480, 640, 3), np.uint8), right_rect=np.zeros((n, 480, 640, 3), np.uint8))
            outs.append(ctx.pipeline_enqueue(li, ri, ALL, pre, want_rect=True, want_xyz=True, want_depth=True))
        ctx.sync()
        for o, res in zip(order, outs):
            check(res, o, "stream")


def test_unfused_front_matches_fused(svgpu, tmp_path):
    """The stage-by-stage kernels (k_remap, blur / gray / sharpen, k_prefilter_planes; SVGPU_UNFUSED=1, read once per process) and
    the fused front kernel (front_fused.cu, the default for raw BGR frames) give the same disparity on the bottle pair, which
    the golden test pins to cv2; run in a child process because the switch is read when the library first runs a pipeline."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, hashlib, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/fhgr-stereo-vision_b200"); sys.path.insert(0, sys.argv[1] + "/tests")
import svgpu
from helpers import golden
from oracle import sv_oracle as oracle
g = golden("bottle"); size = (640, 480)
ml = oracle.init_undistort_rectify_map(g["cal_Kl"], g["cal_Dl"], g["rect_Rl"], g["rect_Pl"], size)
mr = oracle.init_undistort_rectify_map(g["cal_Kr"], g["cal_Dr"], g["rect_Rr"], g["rect_Pr"], size)
with svgpu.Context(*size) as ctx:
    ctx.set_rectify_maps(0, *ml); ctx.set_rectify_maps(1, *mr); ctx.set_sgbm(**svgpu.REFERENCE_SGBM)
    n0 = ctx.launch_count
    out = ctx.pipeline(g["scaled_left"], g["scaled_right"], svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM, want_rect=True)
    print("RESULT", int((out["disp"] != g["disp_raw"]).sum()), int((out["left_rect"] != g["rect_left"]).sum()), ctx.launch_count - n0)
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("fused", {}), ("unfused", {"SVGPU_UNFUSED": "1"})):
        out = subprocess.run([sys.executable, "-c", code, root], capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
        assert out.returncode == 0, out.stderr[-2000:]
        res[tag] = [int(v) for v in [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1].split()[1:]]
    assert res["fused"][:2] == [0, 0] and res["unfused"][:2] == [0, 0], res
    assert res["fused"][2] < res["unfused"][2], res   # fewer kernels: 1 instead of 2 remap + 6 blur/gray/sharpen + 1 planes


def test_folder_front_end_on_the_facade(svgpu, tmp_path):
    """The reference's folder front-end (apps/stereo-vision/static-image-main.cpp:107-127: `left/`, `right/`, `calibration.yml`) on
    the C++ facade: `folder-main` loads the calibration with LoadStereoCameraInfo, reads the frames (PPM), runs them through
    StereoVision::AnalyzeAndAnnotateImages (svgpu_pipeline_batch) two pairs per device call and prints a checksum of every
    disparity map: equal to the golden (cv2) post-filtered disparity of the bottle pair; an unreadable frame gives the
    reference's 'invalid image!' for that pair only."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "fhgr-stereo-vision_b200", "cpp", "folder-main")
    assert os.path.exists(exe), "build() compiles the facade and its programs"
    g = golden("bottle")

    def mat(name, a):
        a = np.asarray(a, np.float64)
        rows, cols = (a.shape if a.ndim == 2 else (1, a.size))
        return (f"   {name}: !!opencv-matrix\n      rows: {rows}\n      cols: {cols}\n      dt: d\n      data: [ " +
                ", ".join(repr(float(v)) for v in a.reshape(-1)) + " ]\n")

    yml = ("%YAML:1.0\n---\ncalibration:\n   file_version: 1\n   image_size: [ 640, 480 ]\n" + mat("camera_matrix_left", g["cal_Kl"]) +
           mat("camera_matrix_right", g["cal_Kr"]) + mat("dist_coeffs_left", g["cal_Dl"].reshape(1, -1)[:, :5]) +
           mat("dist_coeffs_right", g["cal_Dr"].reshape(1, -1)[:, :5]) + mat("R", g["cal_R"]) + mat("T", g["cal_T"].reshape(3, 1)) +
           mat("E", np.zeros((3, 3))) + mat("F", np.zeros((3, 3))) + "   reprojection_error: 0.25\n")
    (tmp_path / "calibration.yml").write_text(yml)
    for side, key in (("left", "scaled_left"), ("right", "scaled_right")):
        d = tmp_path / side
        d.mkdir()
        rgb = np.ascontiguousarray(g[key][:, :, ::-1])
        for i in range(3):
            with open(d / f"frame_{i}.ppm", "wb") as f:
                f.write(b"P6\n640 480\n255\n" + rgb.tobytes())
        (d / "frame_3.ppm").write_bytes(b"not an image")
    out = subprocess.run([exe, str(tmp_path), str(tmp_path / "calibration.yml"), "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr

    def fnv(a):
        h = 1469598103934665603
        for b in np.ascontiguousarray(a).tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    want = fnv(g["disp_post"])
    sums = [int(l.split("disparity checksum")[1]) for l in out.stdout.splitlines() if "disparity checksum" in l]
    assert sums == [want, want, want], out.stdout
    assert "frame_3.ppm / frame_3.ppm: error: invalid image!" in out.stdout
    assert "4 pairs in" in out.stdout


def test_medium_batch_odd_count_speckle_and_streaming(svgpu, oracle):
    """Seven pairs per call (an odd count, all in one chunk): every pair equals the oracle on host and on device buffers, with
    the speckle filter behind the batch, for a sub-batch of the same context and through two overlapping streamed calls."""
    import torch
    from svgpu.synth import synth_batch
    H, W, D, n = 70, 300, 64, 7
    L, R = synth_batch(n, H, W, D, seed0=70)
    for mode, spk in ((1, 0), (3, 60)):
        prm = dict(REF_PARAMS, numDisparities=D, mode=mode, speckleWindowSize=spk, speckleRange=2 if spk else 0)
        want = np.stack([oracle.sgbm(L[i], R[i], **prm) for i in range(n)])
        with svgpu.Context(W, H) as ctx:
            ctx.set_sgbm(**prm)
            ctx.reserve(n)
            assert nmismatch(ctx.sgbm_batch(L, R), want) == 0
            td = ctx.sgbm_batch(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda())
            torch.cuda.synchronize()
            assert nmismatch(td.cpu().numpy(), want) == 0
            assert nmismatch(ctx.sgbm_batch(L[:4], R[:4]), want[:4]) == 0
            ctx.use_own_stream()
            outs = [np.zeros((n, H, W), np.int16) for _ in range(2)]
            for o in outs:
                ctx.sgbm_enqueue(L, R, o)
            ctx.sync()
            assert nmismatch(outs[0], want) == 0 and nmismatch(outs[1], want) == 0


def test_sgbm_block_sizes_13_and_15(svgpu, oracle):
    """blockSize 13 / 15 (row ring of the cost kernel in shared memory; everything above 15 cannot fit the 16-bit cost range and is
    refused like any other overflowing parameter set): bit-exact against the oracle, which test_oracle.py pins to cv2 for these."""
    rng = np.random.default_rng(1315)
    for bs, mode, D, H, W in ((13, 0, 32, 60, 190), (13, 1, 128, 50, 420), (15, 3, 64, 70, 300), (15, 1, 256, 40, 460), (13, 3, 16, 33, 90)):
        l = rng.integers(0, 256, (H, W), dtype=np.uint8)
        r = np.roll(l, -5, axis=1)
        r = np.clip(r.astype(int) + rng.integers(-6, 7, r.shape), 0, 255).astype(np.uint8)
        prm = dict(REF_PARAMS, numDisparities=D, blockSize=bs, P1=100, P2=400, preFilterCap=15, uniquenessRatio=10, mode=mode)
        with svgpu.Context(W, H) as ctx:
            ctx.set_sgbm(**prm)
            assert nmismatch(ctx.sgbm(l, r), oracle.sgbm(l, r, **prm)) == 0, (bs, mode, D)
    with svgpu.Context(200, 60) as ctx:
        with pytest.raises(svgpu.SvgpuError):
            ctx.set_sgbm(**dict(REF_PARAMS, numDisparities=32, blockSize=17, P1=8, P2=32, preFilterCap=15, mode=1))
        with pytest.raises(svgpu.SvgpuError):   # 15 x 15 x (2 x 33 + 63) + P2 leaves the 16-bit cost range
            ctx.set_sgbm(**dict(REF_PARAMS, numDisparities=32, blockSize=15, P2=4000, mode=1))


def test_mixed_call_sizes_on_one_context(svgpu, oracle):
    """One context, calls of 1, 3, 12, 2, 40 and again 1 pairs (640x360, 64 disparities, MODE_HH): the library moves between the
    line form, the side-by-side passes and the fused passes and regrows its scratch on the way; every pair must equal the
    oracle's result for it, whatever call it travelled in."""
    from svgpu.synth import synth_batch
    H, W, D = 360, 640, 64
    L, R = synth_batch(4, H, W, D, seed0=21)
    prm = dict(REF_PARAMS, numDisparities=D, mode=1)
    want = [oracle.sgbm(L[i], R[i], **prm) for i in range(4)]
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(**prm)
        for n in (1, 3, 12, 2, 40, 1):
            idx = np.arange(n) % 4
            got = ctx.sgbm_batch(np.ascontiguousarray(L[idx]), np.ascontiguousarray(R[idx]))
            for k in range(n):
                assert nmismatch(got[k], want[idx[k]]) == 0, (n, k)


def test_sgbm_mode_block_disparity_combinations(svgpu, oracle):
    """Corners of the contract together: blockSize 13 / 15 with MODE_SGBM_3WAY and with 272 .. 512 disparities, 512 disparities
    in MODE_SGBM_3WAY and MODE_HH4 — on a textured and on a noise pair each, against the oracle."""
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
            with svgpu.Context(W, H) as ctx:
                ctx.set_sgbm(**prm)
                assert nmismatch(ctx.sgbm(l, r), oracle.sgbm(l, r, **prm)) == 0, (H, W, D, bs, mode, kind)

#!/usr/bin/env python
"""bench.py — dense-disparity throughput on B200: BASELINE.json's metric on its named configs.

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path, BASELINE configs[2] (the default)
    python bench.py --config 4k256_pipeline ...                   # another BASELINE config (see CONFIGS)
    python bench.py --impl reference [--config X] --gpus N ...    # the reference's CPU path (OpenCV via cv2) on the same config

One "step" = one pass of the hot path over one batch of B pairs per GPU.  metric = Mpix*disp/s = W*H*D*pairs / s / 1e6
(SURVEY.md §8d); fps is in config.

  value         device-resident inputs and outputs, CUDA events on the launching stream, max over ranks
  e2e           the same work through the streaming C-ABI entry point with HOST (pinned) buffers: H2D of the inputs and D2H of
                the results inside the timed region
  roofline      for the dominant kernel (k_sweep: path aggregation + winner-take-all): algorithmic bytes / measured launch
                time / measured HBM peak, the DRAM traffic of the committed ncu capture, and the integer-pipe fractions
  cpu_baseline  the reference's CPU path (cv2) timed on this box's host cores on a bounded sample
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "fhgr-stereo-vision_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

REF = dict(minDisparity=0, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32, uniquenessRatio=15,
           speckleWindowSize=0, speckleRange=0)   # src/lib.cpp:249-258
UNIT = "Mpix*disp/s"
# kind: sgbm = svgpu_sgbm_compute_batch on gray pairs; bm = svgpu_bm_compute; pipeline = svgpu_pipeline_batch on raw BGR frames
# (rectify -> blur/gray/sharpen -> SGBM -> speckle filter + close -> reprojectImageTo3D + depth); stream = sgbm over a fixed
# number of frames through the streaming entry point
CONFIGS = {
    "1080p128": dict(W=1920, H=1080, D=128, mode=1, batch=64, kind="sgbm",
                     metric="SGBM Mpix*disp/s at 1080p/128 disp (MODE_HH, 8 paths + LR check)",
                     workload="1920x1080 pairs, 128 disparities, StereoSGBM MODE_HH + LR check (BASELINE configs[2])"),
    "720p64": dict(W=1280, H=720, D=64, mode=0, batch=64, kind="sgbm",
                   metric="SGBM Mpix*disp/s at 720p/64 disp (MODE_SGBM, 5 paths)",
                   workload="1280x720 pairs, 64 disparities, StereoSGBM MODE_SGBM (5 paths) (BASELINE configs[1])"),
    "bm720p": dict(W=1280, H=720, D=64, mode=-1, batch=64, kind="bm",
                   metric="StereoBM Mpix*disp/s at 720p/64 disp (blockSize 21)",
                   workload="1280x720 pairs, 64 disparities, StereoBM blockSize 21 (BASELINE configs[1])"),
    "ref_hh4": dict(W=640, H=480, D=224, mode=3, batch=64, kind="sgbm",
                    metric="SGBM Mpix*disp/s at 640x480/224 disp (MODE_HH4, the reference's own setting)",
                    workload="640x480 pairs, 224 disparities, StereoSGBM MODE_HH4: the setter block of src/lib.cpp:248-259 (BASELINE configs[0])"),
    "4k256_pipeline": dict(W=3840, H=2160, D=256, mode=1, batch=8, kind="pipeline",
                           metric="full chain Mpix*disp/s at 3840x2160/256 disp (rectify -> SGBM MODE_HH -> XYZ + depth)",
                           workload="3840x2160 BGR pairs, 256 disparities: remap -> blur/gray/sharpen -> StereoSGBM MODE_HH -> "
                                    "filterSpeckles + close -> reprojectImageTo3D + depth map (BASELINE configs[3], src/lib.cpp:53-85)"),
    "stream10k": dict(W=1920, H=1080, D=128, mode=1, batch=64, kind="stream", frames=10000,
                      metric="SGBM Mpix*disp/s over a 10,000-frame 1080p/128 stream (MODE_HH)",
                      workload="10,000 synthetic 1080p frames, 128 disparities, MODE_HH, streamed like the livecam loop "
                               "(livecam-main.cpp:70-106) and sharded over the GPUs (BASELINE configs[4])"),
}
PROFILE_CSV = os.path.join(ROOT, "profiles", "r02_ncu_raw.csv")   # `ncu --set full ... --page raw --csv` of the committed kernels
PROFILE_PAIRS = 64                                                 # pairs per launch in that capture


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_sweep_rows():
    """DRAM bytes and ALU-pipe share of the two sweep launches from the committed ncu capture (fails loudly when it is missing:
    the roofline block must be reproducible from files in the repository)."""
    if not os.path.exists(PROFILE_CSV):
        raise SystemExit(f"bench.py: {PROFILE_CSV} is missing (the committed `ncu --page raw --csv` capture the roofline block reads)")
    rows = list(csv.reader(open(PROFILE_CSV)))
    h, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}

    def val(r, name):
        i = h.index(name)
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)

    out = []
    for r in rows[2:]:
        if "k_sweep" not in r[h.index("Kernel Name")]:
            continue
        out.append({"dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"), "seconds": val(r, "gpu__time_duration.sum"),
                    "alu_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                    "fma_pct": val(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                    "issue_pct": val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")})
    if len(out) < 2:
        raise SystemExit(f"bench.py: {PROFILE_CSV} holds no k_sweep launches")
    return out[-2:]


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def workload_config(name, cfg):
    """The `config` object of the JSON line: identical for the CUDA arm and the reference arm of one workload."""
    return {"workload": cfg["workload"], "name": name, "width": cfg["W"], "height": cfg["H"], "num_disparities": cfg["D"],
            "l2": "per-step working set (cost volumes, W1*H*D*4 bytes per pair in flight) exceeds L2; two input sets alternate"}


def sgbm_params(cfg):
    return dict(REF, numDisparities=cfg["D"], mode=cfg["mode"])


def big_calib(W, H):
    """A synthetic single-camera calibration (radial distortion, 10 % zoom) scaled to the frame size; shared with the parity tests."""
    s = W / 3840.0
    K = np.array([[833.77 * 6 * s, 0, 345.2 * 6 * s], [0, 832.87 * 6 * s, 283.1 * 4.5 * s], [0, 0, 1]])
    dist = np.array([-0.4168, 0.166, 0, 0, 0])
    P = np.array([[K[0, 0] * 0.9, 0, K[0, 2], 0], [0, K[0, 0] * 0.9, K[1, 2], 0], [0, 0, 1, 0]])
    Q = np.array([[1, 0, 0, -K[0, 2]], [0, 1, 0, -K[1, 2]], [0, 0, 0, P[0, 0]], [0, 0, 1 / 75.3, 0]])
    return K, dist, P, Q


def to_bgr(g):
    return np.stack([g, np.roll(g, 1, axis=-2) // 2 + 30, 255 - g], axis=-1).astype(np.uint8)


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return model


def cpu_extras(cfg):
    """the single-process figures of cv2 beside the whole-host number — StereoSGBM on one pair with OpenCV's
    default thread count and with one thread, plus remap / StereoBM / reprojectImageTo3D at the config's frame size."""
    out = {"cpu_model": cpu_info()}
    try:
        import cv2
    except Exception:
        return out
    from svgpu.synth import synth_pair
    W, H, D = cfg["W"], cfg["H"], cfg["D"]
    sW, sH, sD = (W, H, D) if W * H * D <= 1920 * 1080 * 128 else (1920, 1080, 128)   # bounded sample
    l, r = synth_pair(sH, sW, sD, seed=7)
    out["cv2_threads_default"] = int(cv2.getNumThreads())

    def t(fn, n=1):
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n

    prm = dict(REF, numDisparities=sD, mode=max(cfg["mode"], 0))
    s = cv2.StereoSGBM_create(**prm)
    out["sgbm_all_threads_s"] = t(lambda: s.compute(l, r))
    nt = cv2.getNumThreads()
    cv2.setNumThreads(1)
    out["sgbm_one_thread_s"] = t(lambda: s.compute(l, r))
    cv2.setNumThreads(nt)
    out["sgbm_sample"] = f"{sW}x{sH}, {sD} disparities, mode {prm['mode']}, one pair"
    K, dist, P, Q = big_calib(sW, sH)
    m1, m2 = cv2.initUndistortRectifyMap(K, dist, np.eye(3), P, (sW, sH), cv2.CV_16SC2)
    lb = to_bgr(l)
    out["remap_bgr_s"] = t(lambda: cv2.remap(lb, m1, m2, cv2.INTER_LINEAR), 3)
    d = s.compute(l, r)
    out["reproject_s"] = t(lambda: cv2.reprojectImageTo3D(d, Q, handleMissingValues=True), 3)
    bm = cv2.StereoBM_create(min(sD, 128), 21)
    out["stereobm_s"] = t(lambda: bm.compute(l, r), 2)
    return out


class CpuChain:
    """The reference's CPU implementation of a config on all host cores: one pair per worker process (cv2's SGBM modes are
    single-threaded inside OpenCV, SURVEY.md [B#12])."""

    def __init__(self, cfg):
        from oracle.cv2_reference import CpuSgbm
        self.cfg = cfg
        self.cpu = CpuSgbm(sgbm_params(cfg) if cfg["kind"] != "bm" else dict(numDisparities=cfg["D"], blockSize=21),
                           chain=cfg["kind"], calib=big_calib(cfg["W"], cfg["H"]) if cfg["kind"] == "pipeline" else None)
        self.workers, self.kind = self.cpu.workers, self.cpu.kind

    def inputs(self, n, seed0):
        from svgpu.synth import synth_batch
        L, R = synth_batch(n, self.cfg["H"], self.cfg["W"], self.cfg["D"], seed0=seed0)
        if self.cfg["kind"] == "pipeline":
            return to_bgr(L), to_bgr(R)
        return L, R

    def run(self, L, R):
        return self.cpu.run(L, R)

    def close(self):
        self.cpu.close()

    def describe(self):
        return self.cpu.describe()


def run_reference(args, cfg):
    """The reference arm: OpenCV's CPU implementation of the config (what src/lib.cpp runs) on this box's host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    mpd = cfg["W"] * cfg["H"] * cfg["D"] / 1e6
    cpu = CpuChain(cfg)
    n = cpu.workers                      # one bounded sample per step: one pair per worker process
    L, R = cpu.inputs(n, 0)
    for _ in range(min(args.warmup, 1)):  # the pool and OpenCV's buffers are warm after one pass
        cpu.run(L, R)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu.run(L, R)
        t += dt
    desc = cpu.describe()
    cpu.close()
    value = n * args.steps * mpd / t
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s16", "data": "synthetic", "fps": n * args.steps / t,
            "config": workload_config(args.config, cfg), "details": {"pairs_per_step": n, "fps": n * args.steps / t},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.workers, "kind": cpu.kind, "sample": desc, "cpu_model": cpu_info()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_sample(cfg):
    """Bounded CPU sample beside the GPU number (rank 0, N=1): one pair per worker process, one warm pass."""
    mpd = cfg["W"] * cfg["H"] * cfg["D"] / 1e6
    cpu = CpuChain(cfg)
    L, R = cpu.inputs(cpu.workers, 100)
    cpu.run(L[:1], R[:1])
    dt, _ = cpu.run(L, R)
    desc = cpu.describe()
    cpu.close()
    out = {"value": cpu.workers * mpd / dt, "unit": UNIT, "cores": cpu.workers, "kind": cpu.kind,
           "sample": f"{desc}; {cpu.workers} pairs in {dt:.2f} s ({cpu.workers / dt:.3f} fps)"}
    out.update(cpu_extras(cfg))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="svgpu", choices=["svgpu", "reference"])
    ap.add_argument("--config", default="1080p128", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SVGPU_BENCH_BATCH", 0)), help="pairs per step per GPU (default: per config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    args.warmup = max(args.warmup, 3) if args.impl == "svgpu" else args.warmup
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    import svgpu
    from svgpu.sharding import checksum, gather_stats, shard_frames
    from svgpu.synth import synth_pair

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    W, H, D, kind = cfg["W"], cfg["H"], cfg["D"], cfg["kind"]
    B = args.batch or cfg["batch"]
    mpd = W * H * D / 1e6
    if kind == "stream":   # a fixed number of frames for the whole job: every rank streams its share in batches of B
        args.steps = max(1, -(-cfg["frames"] // (B * world)))
    # Frames of the job are dealt round robin (frame i -> rank i mod world); every rank works on its own B frames per step (weak
    # scaling: per-GPU work is fixed).  Two distinct input sets alternate between steps; distinct frames are synthesised for a
    # bounded number of seeds and repeated (the kernels' work does not depend on the pixel values).
    nsets = 2
    distinct = min(B, 16)
    mine = shard_frames(distinct * world * nsets, rank, world)
    base = np.empty((len(mine), H, W), np.uint8)
    baseR = np.empty_like(base)
    for k, fi in enumerate(mine):
        base[k], baseR[k] = synth_pair(H, W, D, seed=fi)
    idx = np.concatenate([np.arange(B) % distinct + s * distinct for s in range(nsets)])
    Lh, Rh = base[idx], baseR[idx]
    if kind == "pipeline":
        Lh, Rh = to_bgr(Lh), to_bgr(Rh)
    Ld, Rd = torch.from_numpy(Lh).to(dev), torch.from_numpy(Rh).to(dev)
    Lp, Rp = torch.from_numpy(Lh).pin_memory(), torch.from_numpy(Rh).pin_memory()   # pinned host buffers for the end-to-end arm
    out = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    outp = [torch.empty((B, H, W), dtype=torch.int16).pin_memory() for _ in range(2)]

    ctx = svgpu.Context(W, H, device=local)
    ALL = (svgpu.STAGE_REMAP | svgpu.STAGE_PREFILTER | svgpu.STAGE_SGBM | svgpu.STAGE_POSTFILTER | svgpu.STAGE_REPROJECT | svgpu.STAGE_DEPTH)
    if kind == "bm":
        ctx.set_bm(numDisparities=D, blockSize=21)
    else:
        ctx.set_sgbm(**sgbm_params(cfg))
        ctx.reserve(B)
    h2d, d2h = 2 * B * W * H, 2 * B * W * H
    if kind == "pipeline":
        K, dd, P, Q = big_calib(W, H)
        for side in (0, 1):
            ctx.init_rectify_maps(side, K, dd, np.eye(3), P)
        ctx.set_Q(Q)
        ctx.set_depth_consts(75.3, float(K[0, 0]))
        dev_out = dict(disp=out, xyz=torch.empty((B, H, W, 3), dtype=torch.float32, device=dev),
                       depth=torch.empty((B, H, W), dtype=torch.float32, device=dev))
        host_out = [dict(disp=outp[s].numpy(), xyz=torch.empty((B, H, W, 3), dtype=torch.float32).pin_memory().numpy(),
                         depth=torch.empty((B, H, W), dtype=torch.float32).pin_memory().numpy()) for s in range(2)]
        h2d, d2h = 2 * B * W * H * 3, B * W * H * (2 + 12 + 4)
    stream = torch.cuda.Stream(device=dev)  # kernels, copies and the timing events all go on this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def step_dev(i):
        s = (i % nsets) * B
        if kind == "pipeline":
            ctx.pipeline_batch(Ld[s:s + B], Rd[s:s + B], ALL, want_xyz=True, want_depth=True, out=dev_out)
        elif kind == "bm":
            for k in range(B):
                ctx.bm(Ld[s + k], Rd[s + k], out=out[k])
        else:
            ctx.sgbm_batch(Ld[s:s + B], Rd[s:s + B], out=out)

    def step_host(i):
        # the streaming entry points: copy-in of step i+1 and copy-out of step i-1 overlap the kernels of step i;
        # timed() synchronises (ctx.sync) before it reads the end event, so every copy is inside the timed region
        s = (i % nsets) * B
        if kind == "pipeline":
            ctx.pipeline_enqueue(Lp[s:s + B].numpy(), Rp[s:s + B].numpy(), ALL, host_out[i & 1], want_xyz=True, want_depth=True)
        elif kind == "bm":
            o = outp[i & 1].numpy()
            for k in range(B):
                o[k] = ctx.bm(Lp[s + k].numpy(), Rp[s + k].numpy())
        else:
            ctx.sgbm_enqueue(Lp[s:s + B].numpy(), Rp[s:s + B].numpy(), outp[i & 1].numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm, streaming=False):
        for i in range(warm):
            fn(i)
        if streaming:
            ctx.sync()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(i)
        if streaming:
            ctx.sync()  # the streaming calls return before their copies are done
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)  # ms

    dev_steps = args.steps if kind != "stream" else min(args.steps, 20)
    sampler = ClockSampler(local)
    sampler.start() if rank == 0 else None
    ms = timed(step_dev, dev_steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    # kernels launched per step: count one extra step exactly
    l0 = ctx.launch_count
    step_dev(0)
    torch.cuda.synchronize()
    launches_per_step = ctx.launch_count - l0
    csum = checksum(out.cpu().numpy())
    # Parity across GPU counts: a fixed set of frames (seeds 0 .. PAR-1, whatever N is) dealt round robin over the ranks; the sum of
    # the per-frame checksums over all ranks must not depend on N (and is what a single GPU computes for the same frames).
    PAR = 16 if W * H <= 1920 * 1080 else 8
    pmine = [i for i in range(PAR) if i % world == rank]
    psum = 0
    if pmine and kind != "bm":
        pl = np.empty((len(pmine), H, W), np.uint8)
        pr = np.empty_like(pl)
        for k, fi in enumerate(pmine):
            pl[k], pr[k] = synth_pair(H, W, D, seed=fi)
        if kind == "pipeline":
            pd = ctx.pipeline_batch(torch.from_numpy(to_bgr(pl)).to(dev), torch.from_numpy(to_bgr(pr)).to(dev), ALL)["disp"]
        else:
            pd = ctx.sgbm_batch(torch.from_numpy(pl).to(dev), torch.from_numpy(pr).to(dev))
        torch.cuda.synchronize()
        psum = checksum(pd.cpu().numpy())

    # per-stage device time of the SGBM core of one step (CUDA events recorded inside the library on the same stream)
    stage = None
    if kind != "bm":
        ctx.enable_stage_timing(True)
        st = []
        for i in range(3):
            step_dev(i)
            st.append(ctx.last_stage_ms())
        ctx.enable_stage_timing(False)
        stage = {k: float(np.median([s[k] for s in st])) for k in st[0]}

    e2e_steps = args.steps if kind == "stream" else max(3, min(args.steps, 10))
    ms_e2e = timed(step_host, e2e_steps, 2, streaming=(kind != "bm"))

    # one pair per call, the way the reference's front-ends call the library (device-resident, CUDA events; every rank runs it,
    # rank 0's is reported): small calls take the latency forms of the aggregation
    single_ms = None
    if kind in ("sgbm", "stream"):
        one = lambda i: ctx.sgbm_batch(Ld[:1], Rd[:1], out=out[:1])
        single_ms = timed(one, 10, 3) / 10

    stats = gather_stats([ms, ms_e2e, float(B * dev_steps), float(csum % (1 << 40)), float(psum % (1 << 24)), float((psum >> 24) % (1 << 24))],
                         device=dev if world > 1 else None)
    ms_max, ms_e2e_max = float(stats[:, 0].max()), float(stats[:, 1].max())
    pairs = float(stats[:, 2].sum())
    if rank == 0:
        value = pairs * mpd / (ms_max / 1e3)
        e2e_value = (B * world * e2e_steps) * mpd / (ms_e2e_max / 1e3)
        if kind == "stream":   # the stream is a host-buffer workload: its headline value is the streamed number
            value_note = "value = device-resident rate over %d steps; e2e = the whole %d-frame stream" % (dev_steps, B * world * e2e_steps)
        hbm_peak, which = peaks()
        line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": dev_steps, "warmup": args.warmup,
                "ms_per_step": ms_max / dev_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "s16", "data": "synthetic", "fps": pairs / (ms_max / 1e3),
                # `config` names the workload and is the same dictionary in both arms (this one and --impl reference);
                # what differs per run is under `details`
                "config": workload_config(args.config, cfg),
                "details": {"pairs_per_step_per_gpu": B, "fps": pairs / (ms_max / 1e3), "sharding": f"frame i -> rank i mod {world}",
                            "checksums": [int(c) for c in stats[:, 3]],
                            # sum over ranks of the checksums of frames 0 .. PAR-1 (48 bits): equal for every N
                            "parity_checksum": int((int(stats[:, 4].sum()) + (int(stats[:, 5].sum()) << 24)) % (1 << 48)), "parity_frames": PAR},
                "clocks": clocks, "gpu_launches": int(launches_per_step * dev_steps),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps, "fps": (B * world * e2e_steps) / (ms_e2e_max / 1e3), "frames": int(B * world * e2e_steps),
                        "api": {"pipeline": "svgpu_pipeline_enqueue_batch + svgpu_sync", "bm": "svgpu_bm_compute with host buffers"}.get(
                            kind, "svgpu_sgbm_enqueue_batch + svgpu_sync") + " on pinned host buffers (copy-in / kernels / copy-out of consecutive steps overlap)"}}
        if single_ms is not None:
            line["details"]["single_pair_call_ms"] = single_ms
        if kind == "stream":
            line["details"]["note"] = value_note
        if stage is not None:
            cells = (W - D) * H * D
            # Dominant kernel: k_sweep (path aggregation + WTA), two launches per step (forward and reverse sweep).  Algorithmic
            # bytes per launch = 4 B per computed volume cell (sweep 1: read C, write S; sweep 2: read C, read S), SURVEY.md §8d;
            # duration = the live CUDA-event time of the aggregation stage / 2 launches.  `traffic` = DRAM bytes per launch of
            # the committed ncu capture (PROFILE_CSV, PROFILE_PAIRS pairs per launch at 1080p/128), scaled by the cell count.
            agg_ms = stage["aggregate"]
            launch_bytes = 4.0 * cells * B
            launch_s = agg_ms / 1e3 / 2
            prof = ncu_sweep_rows()
            ref_cells = (1920 - 128) * 1080 * 128 * PROFILE_PAIRS
            traffic = float(np.mean([p["dram_bytes"] for p in prof])) * (cells * B) / ref_cells
            int_peak = 2.0 * 32 * 2 * 148 * 1.965e9   # tools/ubench_int16x2.cu on B200 (profiles/r02_ubench_int16x2.txt): 2.0 packed
            # warp-instructions / clk / SM x 32 lanes x 2 halves x 148 SMs x 1.965 GHz
            roofline = {"bound": "int_alu", "kernel": "k_sweep (path aggregation + winner-take-all; 2 launches per step)",
                        "achieved": launch_bytes / launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": launch_bytes / launch_s / 1e9 / hbm_peak, "traffic": traffic, "peak_source": which,
                        "traffic_source": os.path.relpath(PROFILE_CSV, ROOT) + f" (mean dram__bytes of the two sweep launches at {PROFILE_PAIRS} pairs, scaled to this step)",
                        "algorithmic_bytes_per_launch": launch_bytes, "launch_ms": launch_s * 1e3, "stage_ms_per_step": stage,
                        "hbm_frac": launch_bytes / launch_s / 1e9 / hbm_peak,
                        "ncu_alu_pipe_pct": [p["alu_pct"] for p in prof], "ncu_fma_pipe_pct": [p["fma_pct"] for p in prof],
                        "ncu_issue_active_pct": [p["issue_pct"] for p in prof],
                        "note": "achieved / peak / frac are the HBM numbers of the sweep launches (the only peak with a driver-measured denominator); the "
                                "binding resource is the packed-integer ALU pipe: ncu_alu_pipe_pct is its utilisation in the committed capture, int_pipe the "
                                "same from operation counts and the live step time",
                        "cost_kernel": {"bytes_per_cell": 2, "achieved": 2.0 * cells * B / (stage["cost"] / 1e3) / 1e9,
                                        "frac": 2.0 * cells * B / (stage["cost"] / 1e3) / 1e9 / hbm_peak},
                        # integer roofline of the SGBM chain: 16-bit lane-operations per volume cell on the packed-integer (ALU) pipe.
                        # 102 = SURVEY.md §8(d)'s algorithmic count for the 8-path mode; 85 = what the three hot kernels executed in the
                        # ncu capture of round 1 (profiles/r01_v8_ncu_raw.csv)
                        "int_pipe": {"ops_per_cell_survey": 102, "ops_per_cell_executed": 85, "peak_tops": int_peak / 1e12,
                                     "peak_source": "profiles/r02_ubench_int16x2.txt",
                                     "frac_survey": 102.0 * cells * B / (stage["total"] / 1e3) / int_peak,
                                     "frac_executed": 85.0 * cells * B / (stage["total"] / 1e3) / int_peak},
                        "chain": {"bytes_per_cell": 10, "achieved": 10.0 * cells * B / (stage["total"] / 1e3) / 1e9,
                                  "frac": 10.0 * cells * B / (stage["total"] / 1e3) / 1e9 / hbm_peak}}
            line["roofline"] = roofline
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample(cfg)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

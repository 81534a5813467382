#!/usr/bin/env python
"""bench.py — SGBM throughput on B200: BASELINE.json's metric on its named config.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (OpenCV via cv2)

Workload (config.workload): BASELINE.json configs[2] — synthetic 1920x1080 pairs, 128 disparities,
cv::StereoSGBM MODE_HH (8 paths) + LR check, the reference's other SGBM parameters
(src/lib.cpp:249-258).  One "step" = one pass of the hot path (svgpu_sgbm_compute_batch) over one batch
of B pairs per GPU.  metric = Mpix*disp/s = W*H*D*pairs / s / 1e6; fps is in config.

  value   device-resident inputs and outputs, CUDA events on the launching stream, max over ranks
  e2e     the same call with HOST (pinned) buffers: H2D of both images and D2H of the disparity inside
          the timed region
  roofline  for the dominant kernel, algorithmic bytes / measured kernel time / measured HBM peak
  cpu_baseline  the reference's CPU path timed on this box's host cores on a bounded sample
"""
import argparse
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

W, H, D = 1920, 1080, 128
SGBM = dict(minDisparity=0, numDisparities=D, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32,
            uniquenessRatio=15, speckleWindowSize=0, speckleRange=0, mode=1)  # MODE_HH
METRIC = "SGBM Mpix*disp/s at 1080p/128 disp (MODE_HH, 8 paths + LR check)"
UNIT = "Mpix*disp/s"
MPD_PER_PAIR = W * H * D / 1e6


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def run_reference(args):
    """The reference arm: OpenCV's CPU StereoSGBM (what src/lib.cpp:262 runs) on this box's host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    from oracle.cv2_reference import CpuSgbm
    from svgpu.synth import synth_batch
    cpu = CpuSgbm(SGBM)
    n = cpu.workers                      # one bounded sample per step: one pair per worker process
    L, R = synth_batch(n, H, W, D, seed0=0)
    for _ in range(min(args.warmup, 1)):  # the pool and OpenCV's buffers are warm after one pass
        cpu.run(L, R)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu.run(L, R)
        t += dt
    cpu.close()
    value = n * args.steps * MPD_PER_PAIR / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s16", "data": "synthetic",
            "config": {"workload": "1920x1080 pairs, 128 disparities, StereoSGBM MODE_HH + LR check (BASELINE configs[2])",
                       "pairs_per_step": n, "fps": n * args.steps / t},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.workers, "kind": cpu.kind, "sample": cpu.describe()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_sample():
    """Bounded CPU sample beside the GPU number (rank 0, N=1): one pair per worker process, one warm pass."""
    from oracle.cv2_reference import CpuSgbm
    from svgpu.synth import synth_batch
    cpu = CpuSgbm(SGBM)
    L, R = synth_batch(cpu.workers, H, W, D, seed0=100)
    cpu.run(L[:1], R[:1])
    dt, _ = cpu.run(L, R)
    cpu.close()
    return {"value": cpu.workers * MPD_PER_PAIR / dt, "unit": UNIT, "cores": cpu.workers, "kind": cpu.kind,
            "sample": f"{cpu.describe()}; {cpu.workers} pairs in {dt:.2f} s ({cpu.workers / dt:.3f} fps)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="svgpu", choices=["svgpu", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SVGPU_BENCH_BATCH", 64)), help="pairs per step per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "svgpu" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import svgpu
    from svgpu.sharding import checksum, gather_stats, shard_frames
    from svgpu.synth import synth_batch

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    # Frames of the job are dealt round robin (frame i -> rank i mod world); every rank works on its own B
    # frames per step (weak scaling: per-GPU work is fixed). Two distinct input sets alternate between steps.
    nsets = 2
    total = B * world * nsets
    mine = shard_frames(total, rank, world)
    Lh = np.empty((len(mine), H, W), np.uint8)
    Rh = np.empty_like(Lh)
    from svgpu.synth import synth_pair
    for k, fi in enumerate(mine):
        Lh[k], Rh[k] = synth_pair(H, W, D, seed=fi)
    Ld, Rd = torch.from_numpy(Lh).to(dev), torch.from_numpy(Rh).to(dev)
    out = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    # pinned host buffers for the end-to-end arm
    Lp, Rp = torch.from_numpy(Lh).pin_memory(), torch.from_numpy(Rh).pin_memory()
    outp = torch.empty((B, H, W), dtype=torch.int16).pin_memory()

    ctx = svgpu.Context(W, H, device=local)
    ctx.set_sgbm(**SGBM)
    ctx.reserve(B)
    stream = torch.cuda.Stream(device=dev)  # kernels, copies and the timing events all go on this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def step_dev(i):
        s = (i % nsets) * B
        ctx.sgbm_batch(Ld[s:s + B], Rd[s:s + B], out=out)

    outp2 = [outp, torch.empty((B, H, W), dtype=torch.int16).pin_memory()]

    def step_host(i):
        # the streaming entry point: copy-in of step i+1 and copy-out of step i-1 overlap the kernels of step i;
        # timed() synchronises (ctx.sync) before it reads the end event, so every copy is inside the timed region
        s = (i % nsets) * B
        ctx.sgbm_enqueue(Lp[s:s + B].numpy(), Rp[s:s + B].numpy(), outp2[i & 1].numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm, streaming=False):
        for i in range(warm):
            fn(i)
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

    sampler = ClockSampler(local)
    sampler.start() if rank == 0 else None
    ms = timed(step_dev, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    # kernels launched per step: count one extra step exactly
    l0 = ctx.launch_count
    step_dev(0)
    torch.cuda.synchronize()
    launches_per_step = ctx.launch_count - l0
    csum = checksum(out.cpu().numpy())

    # per-stage device time of one step (CUDA events recorded inside the library on the same stream)
    ctx.enable_stage_timing(True)
    stage = []
    for i in range(3):
        step_dev(i)
        stage.append(ctx.last_stage_ms())
    ctx.enable_stage_timing(False)
    stage = {k: float(np.median([s[k] for s in stage])) for k in stage[0]}

    ms_e2e = timed(step_host, max(3, min(args.steps, 10)), 2, streaming=True)
    e2e_steps = max(3, min(args.steps, 10))

    stats = gather_stats([ms, ms_e2e, float(B * args.steps), float(csum % (1 << 40))], device=dev if world > 1 else None)
    ms_max, ms_e2e_max = float(stats[:, 0].max()), float(stats[:, 1].max())
    pairs = float(stats[:, 2].sum())
    if rank == 0:
        value = pairs * MPD_PER_PAIR / (ms_max / 1e3)
        e2e_value = (B * world * e2e_steps) * MPD_PER_PAIR / (ms_e2e_max / 1e3)
        hbm_peak, which = peaks()
        cells = (W - D) * H * D
        # Dominant kernel: k_sweep (path aggregation + WTA), two launches per step (forward and reverse sweep).
        # Algorithmic bytes per launch = 4 B per computed volume cell (sweep 1: read C, write S; sweep 2: read C, read S),
        # SURVEY.md §8d; duration = the live CUDA-event time of the aggregation stage / 2 launches.  `traffic` = DRAM
        # bytes per launch from the committed `ncu --set full` capture (profiles/r01_v3_summary.md: 1.021 GB per pair for
        # sweep 1, 1.038 GB for sweep 2), scaled to this batch.
        agg_ms = stage["aggregate"]
        launch_bytes = 4.0 * cells * B
        launch_s = agg_ms / 1e3 / 2
        roofline = {"bound": "hbm", "kernel": "k_sweep<8,15,*> (path aggregation + winner-take-all; 2 launches per step)",
                    "achieved": launch_bytes / launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": launch_bytes / launch_s / 1e9 / hbm_peak, "traffic": 1.03e9 * B, "peak_source": which,
                    "algorithmic_bytes_per_launch": launch_bytes, "launch_ms": launch_s * 1e3, "stage_ms_per_step": stage,
                    "note": "the sweeps are bound by the packed-integer ALU pipe (ncu: ~60 % of the pipe over the launch, profiles/r01_v7_summary.md), not by HBM; "
                            "HBM is reported because it is the only peak with a driver-measured denominator",
                    "cost_kernel": {"bytes_per_cell": 2, "achieved": 2.0 * cells * B / (stage["cost"] / 1e3) / 1e9,
                                    "frac": 2.0 * cells * B / (stage["cost"] / 1e3) / 1e9 / hbm_peak},
                    # integer roofline of the whole chain: 85 16-bit lane-operations per volume cell on the packed-integer (ALU)
                    # pipe, as executed by the three hot kernels in the ncu capture of variant 8 (profiles/r01_v8_ncu_raw.csv;
                    # SURVEY.md §8d's algorithmic count for the 8-path mode is 102); peak = that pipe as measured by
                    # tools/ubench_int16x2.cu on B200: 2.0 warp-instructions / clk / SM x 32 lanes x 2 halves x 148 SMs x 1.965 GHz
                    "int_pipe": {"ops_per_cell": 85, "achieved_tops": 85.0 * cells * pairs / (ms_max / 1e3) / 1e12 / world,
                                 "peak_tops": 2.0 * 32 * 2 * 148 * 1.965e9 / 1e12,
                                 "frac": 85.0 * cells * pairs / (ms_max / 1e3) / world / (2.0 * 32 * 2 * 148 * 1.965e9)},
                    "chain": {"bytes_per_cell": 10, "achieved": 10.0 * cells * B / (stage["total"] / 1e3) / 1e9,
                              "frac": 10.0 * cells * B / (stage["total"] / 1e3) / 1e9 / hbm_peak}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "s16", "data": "synthetic",
                "config": {"workload": "1920x1080 pairs, 128 disparities, StereoSGBM MODE_HH + LR check (BASELINE configs[2])",
                           "pairs_per_step_per_gpu": B, "fps": pairs / (ms_max / 1e3), "sharding": f"frame i -> rank i mod {world}",
                           "l2": "per-step working set (cost volumes, ~1 GB per pair) exceeds L2; two input sets alternate",
                           "checksums": [int(c) for c in stats[:, 3]]},
                "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * B * W * H), "d2h_bytes_per_step": int(2 * B * W * H),
                        "steps": e2e_steps, "fps": (B * world * e2e_steps) / (ms_e2e_max / 1e3),
                        "api": "svgpu_sgbm_enqueue_batch + svgpu_sync on pinned host buffers (copy-in / kernels / copy-out of consecutive steps overlap)"},
                "roofline": roofline}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

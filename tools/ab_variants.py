#!/usr/bin/env python
"""A/B of kernel tuning variants on the GPU box (developer tool): one subprocess per library build, the same cached inputs,
stage times from the library's own CUDA events, and a checksum of the disparities so that a wrong variant shows up as such.

    python tools/ab_variants.py [--batch 64] [--size 1920x1080] [--disp 128] [--mode 1] default build/variants/x/libsvgpu.so ...
"""
import argparse, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
REF = dict(minDisparity=0, blockSize=7, P1=392, P2=1568, disp12MaxDiff=1, preFilterCap=32, uniquenessRatio=15, speckleWindowSize=0, speckleRange=0)


def child(a):
    import numpy as np, torch, svgpu
    W, H = (int(v) for v in a.size.split("x"))
    d = np.load(a.data)
    L8, R8 = d["L"], d["R"]
    B = a.batch
    idx = np.arange(B) % L8.shape[0]
    Ld, Rd = torch.from_numpy(L8[idx]).cuda(), torch.from_numpy(R8[idx]).cuda()
    out = torch.empty((B, H, W), dtype=torch.int16, device="cuda")
    with svgpu.Context(W, H) as ctx:
        ctx.set_sgbm(numDisparities=a.disp, mode=a.mode, **REF); ctx.reserve(B)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        # whole-call time first (CUDA events, stage timing off: the library may then run a medium batch as two halves on two
        # streams), then the stage split of the same call with the library's own events
        for i in range(2):
            ctx.sgbm_batch(Ld, Rd, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(max(a.reps, 1)):
            ctx.sgbm_batch(Ld, Rd, out=out)
        e1.record()
        torch.cuda.synchronize()
        call_ms = e0.elapsed_time(e1) / max(a.reps, 1)
        ctx.enable_stage_timing(True)
        st = []
        for i in range(a.reps + 2):
            ctx.sgbm_batch(Ld, Rd, out=out); st.append(ctx.last_stage_ms())
        torch.cuda.synchronize()
        m = {k: float(np.median([s[k] for s in st[2:]])) for k in st[0]} if a.reps else dict(st[-1])
        m["call"] = call_ms
        o = out[:L8.shape[0]].cpu().numpy()
        same = bool((out.cpu().numpy() == o[idx]).all())
    print(json.dumps(dict(fps=B / m["call"] * 1e3, sha=hashlib.sha256(o.tobytes()).hexdigest()[:12], rep_same=same, **m)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64); ap.add_argument("--size", default="1920x1080"); ap.add_argument("--disp", type=int, default=128)
    ap.add_argument("--mode", type=int, default=1); ap.add_argument("--reps", type=int, default=5); ap.add_argument("--child", action="store_true")
    ap.add_argument("--data", default="/tmp/ab_data.npz"); ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        return child(a)
    import numpy as np
    from svgpu.synth import synth_batch
    W, H = (int(v) for v in a.size.split("x"))
    a.data = f"/tmp/ab_data_{W}x{H}_{a.disp}.npz"
    if not os.path.exists(a.data):
        L, R = synth_batch(8, H, W, a.disp, seed0=0)
        np.savez(a.data, L=L, R=R)
    ref = None
    for lib in a.libs or ["default"]:
        env = dict(os.environ)
        if lib != "default":
            env["SVGPU_LIB"] = os.path.abspath(lib)
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--batch", str(a.batch), "--size", a.size, "--disp", str(a.disp), "--mode", str(a.mode),
               "--reps", str(a.reps), "--data", a.data]
        try:
            p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
            r = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
            ref = ref or r["sha"]
            tag = lib if lib == "default" else os.path.basename(os.path.dirname(lib))
            print(f"{tag:14s} B={a.batch:3d} fps {r['fps']:7.1f} (call {r['call']:6.2f} ms)  cost {r['cost']:6.2f} agg {r['aggregate']:6.2f} sel {r['select']:5.2f}  "
                  f"{'OK ' if r['sha'] == ref and r['rep_same'] else 'WRONG'} {r['sha']}", flush=True)
        except Exception as e:
            print(lib, "FAILED", e, (p.stdout + p.stderr)[-400:] if 'p' in dir() else "", flush=True)


if __name__ == "__main__":
    main()

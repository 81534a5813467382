#!/usr/bin/env python
"""Developer experiment: N contexts on one GPU, each on its own stream and host thread, B pairs per call
(tests whether independent launches fill each other's tails).  usage: dual_ctx_bench.py NCTX B [steps]"""
import sys, os, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
import numpy as np, torch, svgpu
from svgpu.synth import synth_pair
from bench import SGBM, W, H, D
nctx, B = int(sys.argv[1]), int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = torch.device("cuda", 0)
L = np.empty((B, H, W), np.uint8); R = np.empty_like(L)
for k in range(B): L[k], R[k] = synth_pair(H, W, D, seed=k)
Ld, Rd = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
ctxs, outs, streams = [], [], []
for i in range(nctx):
    c = svgpu.Context(W, H, device=0); c.set_sgbm(**SGBM); c.reserve(B)
    s = torch.cuda.Stream(device=dev); c.set_stream(s.cuda_stream)
    ctxs.append(c); streams.append(s); outs.append(torch.empty((B, H, W), dtype=torch.int16, device=dev))
def work(i, n):
    with torch.cuda.stream(streams[i]):
        for _ in range(n): ctxs[i].sgbm_batch(Ld, Rd, out=outs[i])
for i in range(nctx): work(i, 2)
torch.cuda.synchronize()
t0 = time.time()
th = [threading.Thread(target=work, args=(i, steps)) for i in range(nctx)]
[t.start() for t in th]; [t.join() for t in th]
torch.cuda.synchronize()
dt = time.time() - t0
print(f"nctx {nctx} B {B}: {nctx * B * steps / dt:.1f} fps", flush=True)

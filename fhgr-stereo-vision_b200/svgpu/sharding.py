"""Frame sharding over the GPUs of one box.

Stereo pairs are independent (StereoVision::AnalyzeAndAnnotateImage is a pure function of its inputs
and the immutable maps / Q / settings, src/lib.cpp:53-85), so the path shards with no data-path
collective: frame i goes to rank i mod world.  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) is used only after the work, to gather per-rank counts, timings and checksums.
"""
import numpy as np


def shard_frames(n_frames, rank, world):
    """Indices of the frames rank `rank` of `world` processes (round robin, like a livecam feeding N GPUs)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_frames, world))


def checksum(disp):
    """Order-independent 64-bit checksum of a CV_16S disparity batch (sum of per-frame FNV-style hashes)."""
    a = np.ascontiguousarray(np.asarray(disp)).view(np.uint16).astype(np.uint64).reshape(len(disp), -1)
    w = (np.arange(a.shape[1], dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    return int(((a * w).sum(axis=1, dtype=np.uint64)).sum(dtype=np.uint64) & np.uint64(0x7FFFFFFFFFFFFFFF))


def gather_stats(local, device=None):
    """all_gather of a small per-rank float64 vector (elapsed, frames, checksum pieces). Returns [world][k]."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(local, dtype=torch.float64, device=device or "cpu")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.cpu().numpy()[None, :]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()

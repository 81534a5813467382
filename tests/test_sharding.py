"""CPU tests of the multi-GPU host logic: frames are dealt round robin to one process per GPU with
no data-path collective; torch.distributed only gathers per-rank counts / timings / checksums afterwards.
world_size 2 over gloo, the same code path bench.py runs over NCCL."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
    import torch.distributed as dist
    from svgpu.sharding import checksum, gather_stats, shard_frames
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_frames(n_frames, rank, world)
    # stand-in for the per-frame result: a deterministic disparity map per frame index
    disp = np.stack([np.full((4, 6), 16 * f - 16, np.int16) + np.arange(6, dtype=np.int16) for f in mine]) if mine else np.zeros((0, 4, 6), np.int16)
    stats = gather_stats([float(len(mine)), 10.0 + rank, float(checksum(disp) % (1 << 40)) if len(mine) else 0.0])
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put(stats.tolist())


def test_round_robin_sharding_and_gather_world2():
    from svgpu.sharding import checksum, shard_frames
    n = 7
    assert shard_frames(n, 0, 2) == [0, 2, 4, 6] and shard_frames(n, 1, 2) == [1, 3, 5]
    assert sorted(shard_frames(n, 0, 3) + shard_frames(n, 1, 3) + shard_frames(n, 2, 3)) == list(range(n))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats = np.array(q.get(timeout=120))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert stats.shape == (2, 3)
    assert stats[:, 0].tolist() == [4.0, 3.0]          # frames per rank
    assert stats[:, 1].max() == 11.0                    # max-over-ranks timing is what bench.py reports
    # the checksum is a sum of per-frame hashes: sharded sums add up to the unsharded one
    whole = np.stack([np.full((4, 6), 16 * f - 16, np.int16) + np.arange(6, dtype=np.int16) for f in range(n)])
    parts = [checksum(whole[shard_frames(n, r, 2)]) for r in range(2)]
    assert (parts[0] + parts[1]) & 0x7FFFFFFFFFFFFFFF == checksum(whole)  # sum of per-frame hashes mod 2^63
    assert int(stats[0, 2]) == parts[0] % (1 << 40) and int(stats[1, 2]) == parts[1] % (1 << 40)


def test_shard_frames_rejects_bad_rank():
    import pytest
    from svgpu.sharding import shard_frames
    with pytest.raises(ValueError):
        shard_frames(4, 2, 2)


def test_parity_checksum_is_independent_of_the_gpu_count():
    """bench.py's `parity_checksum`: every rank sends the two 24-bit halves of the checksum of its share of a fixed frame set; the
    recombined sum over the ranks (mod 2^48) is the same for 1, 2, 4 and 8 ranks and equals the unsharded checksum."""
    sys.path.insert(0, os.path.join(ROOT, "fhgr-stereo-vision_b200"))
    from svgpu.sharding import checksum, shard_frames
    rng = np.random.default_rng(8)
    frames = rng.integers(-16, 4000, (16, 9, 11)).astype(np.int16)
    want = checksum(frames) % (1 << 48)
    for world in (1, 2, 4, 8):
        lo = hi = 0.0
        for rank in range(world):
            mine = shard_frames(16, rank, world)
            p = checksum(frames[mine])
            lo += float(p % (1 << 24))                  # what the rank contributes to the float64 all_gather
            hi += float((p >> 24) % (1 << 24))
        assert (int(lo) + (int(hi) << 24)) % (1 << 48) == want, world

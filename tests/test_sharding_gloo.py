"""Frame sharding: block partition + world_size-2 gloo gather of per-frame results on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mono_lidar_depth_b200 import sharding


def test_frame_blocks_cover_sequence_contiguously():
    for total in (0, 1, 7, 10000, 100000):
        for world in (1, 2, 3, 4, 8):
            blocks = sharding.all_blocks(total, world)
            assert sum(c for _, c in blocks) == total
            pos = 0
            for start, count in blocks:
                assert start == min(pos, total) or count == 0
                pos += count
            assert max(c for _, c in blocks) <= -(-total // world) if total else True
    assert sharding.frame_block(100000, 8, 7) == (87500, 12500)
    with pytest.raises(ValueError):
        sharding.frame_block(10, 2, 2)


def _worker(rank, world, port, total, F, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = sharding.frame_block(total, world, rank)
    # each rank "computes" its own frames: depth = frame index + feature/1000, status = frame % 17
    frames = torch.arange(start, start + count, dtype=torch.float64)[:, None]
    depth = frames + torch.arange(F, dtype=torch.float64)[None, :] / 1000.0
    status = (frames.to(torch.int32) % 17).expand(count, F).contiguous()
    d, s = sharding.gather_results(depth, status, total)
    exp_frames = torch.arange(total, dtype=torch.float64)[:, None]
    ok = torch.equal(d, exp_frames + torch.arange(F, dtype=torch.float64)[None, :] / 1000.0) and torch.equal(
        s, (exp_frames.to(torch.int32) % 17).expand(total, F))
    # gather on rank 0 only (what bench.py does over NCCL): the other rank receives nothing
    d0, s0 = sharding.gather_results(depth, status, total, dst=0)
    if rank == 0:
        ok = ok and torch.equal(d0, d) and torch.equal(s0, s)
    else:
        ok = ok and d0 is None and s0 is None
    q.put((rank, bool(ok), tuple(d.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [9, 10])
def test_gather_results_world_size_2_gloo(total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    F = 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, F, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (total, F) for _, _, shape in res)

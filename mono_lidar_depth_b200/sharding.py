"""Frame sharding across the GPUs of one box.

Frames are independent units (the only cross-frame state lives in the out-of-scope caller,
tracklets_depth/include/tracklets_depth/tracklet_depth_module.h:145-149), so a sequence is cut into
contiguous blocks of ceil(T/G) frames, one per rank, with no collective on the data path; the
per-frame results are gathered once at the end.
"""
from __future__ import annotations

from typing import List, Tuple


def frame_block(total_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first frame, frame count) of `rank`'s contiguous block."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    per = -(-total_frames // world_size) if total_frames > 0 else 0
    start = min(rank * per, total_frames)
    return start, max(0, min(per, total_frames - start))


def all_blocks(total_frames: int, world_size: int) -> List[Tuple[int, int]]:
    return [frame_block(total_frames, world_size, r) for r in range(world_size)]


def gather_results(depth, status, total_frames: int, group=None, dst=None):
    """Gather the per-rank (frames_r, F) result tensors into (total_frames, F): on every rank (dst=None, all_gather) or
    only on rank `dst` (gather; the other ranks get (None, None) and receive nothing -- what bench.py uses, since the
    consumer of a sequence's depths is one process).

    Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests). Blocks are
    padded to the common ceil(T/G) length for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = -(-total_frames // world) if total_frames > 0 else 0
    F = depth.shape[1]

    def padded(t):
        if t.shape[0] == per:
            return t.contiguous()
        out = torch.zeros((per, F), dtype=t.dtype, device=t.device)
        out[: t.shape[0]] = t
        return out

    if dst is not None:
        me = dist.get_rank(group)
        d_all = [torch.empty((per, F), dtype=depth.dtype, device=depth.device) for _ in range(world)] if me == dst else None
        s_all = [torch.empty((per, F), dtype=status.dtype, device=status.device) for _ in range(world)] if me == dst else None
        dist.gather(padded(depth), d_all, dst=dst, group=group)
        dist.gather(padded(status), s_all, dst=dst, group=group)
        if me != dst:
            return None, None
        return torch.cat(d_all, 0)[:total_frames], torch.cat(s_all, 0)[:total_frames]
    d_all = [torch.empty((per, F), dtype=depth.dtype, device=depth.device) for _ in range(world)]
    s_all = [torch.empty((per, F), dtype=status.dtype, device=status.device) for _ in range(world)]
    dist.all_gather(d_all, padded(depth), group=group)
    dist.all_gather(s_all, padded(status), group=group)
    d = torch.cat(d_all, 0)[:total_frames]
    s = torch.cat(s_all, 0)[:total_frames]
    return d, s

"""Frame sharding across the GPUs of one box.

Frames (or camera streams) are independent in detect + pose, so rank g of G processes the
contiguous block frames[g*B//G : (g+1)*B//G] on its own GPU with its own handle and streams;
there is NO collective on the data path.  Only the compact per-frame detection records travel:
they are gathered to rank 0 (torch.distributed object gather over the process group the caller
initialised: NCCL ranks on a GPU box, gloo in the CPU tests) and concatenated in frame order.
The reference is single-process (one ROS spinner, aruco_slam_node.cpp:79); this is the harness
side of `bench.py --gpus N` and of multi-camera use."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence


def shard_bounds(n_frames: int, world: int) -> List[int]:
    """block boundaries: rank r owns [b[r], b[r+1])"""
    if world < 1 or n_frames < 0:
        raise ValueError("world must be >= 1 and n_frames >= 0")
    return [(n_frames * r) // world for r in range(world + 1)]


def shard_slice(n_frames: int, rank: int, world: int) -> slice:
    b = shard_bounds(n_frames, world)
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    return slice(b[rank], b[rank + 1])


def detect_sharded(detect_fn: Callable[[Sequence], list], frames: Sequence, rank: int, world: int, group=None) -> Optional[list]:
    """Run detect_fn on this rank's block of `frames` (detect_fn returns one record per frame) and
    gather the records on rank 0 in frame order.  Returns the full list on rank 0, None elsewhere.
    An empty shard (more ranks than frames) contributes an empty list."""
    sl = shard_slice(len(frames), rank, world)
    local = list(detect_fn(frames[sl])) if sl.stop > sl.start else []
    if len(local) != sl.stop - sl.start:
        raise RuntimeError("detect_fn returned %d records for %d frames" % (len(local), sl.stop - sl.start))
    if world == 1:
        return local
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out: list = []
    for part in gathered:
        out.extend(part)
    return out

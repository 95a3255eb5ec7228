"""Frame / super-frame sharding across the GPUs of one box.

Every ``encode_profile_from_raw`` / ``decode_profile_to_raw`` call is independent, so frames shard with
NO data-path collective: frame f goes to rank f mod world, each rank runs its frames on its own GPU, and
the only communication is the host-side gather of the per-frame results in frame order (gloo on CPU
tensors, or NCCL on device tensors).  This module holds that plumbing; it never touches codec data itself.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment: frame f -> rank f mod world."""
    return list(range(rank, n_frames, world))


def owner_of(frame: int, world: int) -> int:
    return frame % world


def run_sharded(n_frames: int, process_frame: Callable[[int], torch.Tensor], rank: int, world: int) -> dict:
    """Run ``process_frame(f)`` for this rank's frames; returns {frame: result tensor}."""
    return {f: process_frame(f) for f in frames_for_rank(n_frames, rank, world)}


def gather_in_frame_order(local: dict, n_frames: int, rank: int, world: int, dst: int = 0,
                          group: Optional[dist.ProcessGroup] = None) -> Optional[List[torch.Tensor]]:
    """Host-side gather of per-frame byte tensors (possibly of different lengths) to ``dst`` in frame order.

    Two collectives in total: the per-frame lengths, then one padded payload per rank.
    Returns the list of frames on ``dst`` and None elsewhere.
    """
    if world == 1:
        return [local[f] for f in range(n_frames)]
    mine = frames_for_rank(n_frames, rank, world)
    device = next(iter(local.values())).device if local else torch.device("cpu")
    lens = torch.zeros(n_frames, dtype=torch.int64, device=device)
    for f in mine:
        lens[f] = local[f].numel()
    dist.all_reduce(lens, op=dist.ReduceOp.SUM, group=group)
    per_rank = [int(sum(int(lens[f]) for f in frames_for_rank(n_frames, r, world))) for r in range(world)]
    cap = max(per_rank) if per_rank else 0
    payload = torch.zeros(cap, dtype=torch.uint8, device=device)
    off = 0
    for f in mine:
        n = local[f].numel()
        payload[off:off + n] = local[f].reshape(-1)
        off += n
    bufs = [torch.zeros(cap, dtype=torch.uint8, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(payload, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out: List[Optional[torch.Tensor]] = [None] * n_frames
    for r in range(world):
        off = 0
        for f in frames_for_rank(n_frames, r, world):
            n = int(lens[f])
            out[f] = bufs[r][off:off + n].clone()
            off += n
    return out  # type: ignore[return-value]


def split_superframes(n_words: int, superframe_words: int) -> Sequence[range]:
    """Optional segmentation of one raw-word stream into super-frames of ``superframe_words`` words each
    (the reference never segments: one call = one header + one body, SURVEY bug B10)."""
    return [range(s, min(s + superframe_words, n_words)) for s in range(0, n_words, superframe_words)]

"""Frame sharding across the processes of one box, for callers that run ONE PROCESS PER GPU (bench.py under torchrun):
frame f goes to rank f mod world, each rank codes its frames on its own GPU with no data-path collective, and the only
communication is the host-side gather of the per-frame streams in frame order.  (A single process that drives several GPUs uses
t3c_stream_* / ``Stream`` instead: same assignment, host threads instead of ranks, nothing to gather.)  bench.py takes its config-4
frame assignment from here; tests/test_sharding_gloo.py runs the gather on two gloo ranks."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment: frame f -> rank f mod world."""
    return list(range(rank, n_frames, world))


def owner_of(frame: int, world: int) -> int:
    return frame % world


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

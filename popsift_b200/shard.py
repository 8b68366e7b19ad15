"""Multi-GPU sharding of the hot path: frames are independent, so frame i goes to rank
i mod G and slot (i div G) mod S; there is no data-path collective.  The only communication is the
barrier around the timed region and the MAX / SUM reduction of (time, work) for reporting, which
works on any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple


def frame_indices(n_frames: int, rank: int, world: int) -> List[int]:
    """Indices of the frames rank `rank` of `world` extracts (round robin, reference semantics:
    one PopSift object per device, popsift.h:158-168)."""
    assert 0 <= rank < world
    return list(range(rank, n_frames, world))


def slot_of(local_index: int, slots: int) -> int:
    """Slot (CUDA stream + buffers) used for the k-th frame of a rank."""
    return local_index % slots


def aggregate(local_units: float, local_ms: float, device=None) -> Tuple[float, float]:
    """(sum of units over ranks, max of elapsed ms over ranks).  No-op without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_units), float(local_ms)
    t = torch.tensor([local_units], dtype=torch.float64, device=device)
    m = torch.tensor([local_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(t.item()), float(m.item())

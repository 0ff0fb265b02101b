"""Frame sharding across the GPUs of one box.

Frames are independent for detection and stereo matching (reference src/Pipeline.cpp:64-145 reads nothing from
previous frames), so a batch of F frames is split into contiguous blocks, one per rank, and there is NO
collective on the data path.  The only exchange is the final result gather, done here with torch.distributed
(NCCL on GPUs, gloo in the CPU tests): per-frame mate counts first, then the padded mate records.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, world: int, rank: int):
    """Contiguous block of ceil(F/G) frames per rank (last ranks may get fewer, or none)."""
    per = (n_frames + world - 1) // world
    lo = min(rank * per, n_frames)
    hi = min(lo + per, n_frames)
    return lo, hi


def gather_mates(local_mates, local_counts, n_frames: int, dist=None, device=None):
    """Gather per-frame results to rank 0, indexed by global frame id.

    local_mates: [f_local, cap] structured array (MATE_DTYPE, 64 B records); local_counts: [f_local] int32.
    Returns (mates[n_frames, cap], counts[n_frames]) on rank 0 and (None, None) elsewhere.
    """
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_mates, local_counts
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = local_mates.shape[1]
    per = (n_frames + world - 1) // world
    pad_m = np.zeros((per, cap, 64), np.uint8)
    pad_c = np.zeros(per, np.int32)
    nloc = len(local_counts)
    if nloc:
        pad_m[:nloc] = local_mates.view(np.uint8).reshape(nloc, cap, 64)
        pad_c[:nloc] = local_counts
    tm = torch.from_numpy(pad_m)
    tc = torch.from_numpy(pad_c)
    if device is not None:
        tm, tc = tm.to(device), tc.to(device)
    gm = [torch.empty_like(tm) for _ in range(world)] if rank == 0 else None
    gc = [torch.empty_like(tc) for _ in range(world)] if rank == 0 else None
    dist.gather(tm, gm, dst=0)
    dist.gather(tc, gc, dst=0)
    if rank != 0:
        return None, None
    allm = torch.stack(gm).cpu().numpy().reshape(world * per, cap, 64)[:n_frames]
    allc = torch.stack(gc).cpu().numpy().reshape(world * per)[:n_frames]
    return np.ascontiguousarray(allm).view(local_mates.dtype).reshape(n_frames, cap), allc

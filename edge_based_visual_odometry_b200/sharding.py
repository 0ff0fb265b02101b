"""Frame sharding across the GPUs of one box.

Frames are independent for detection and stereo matching (reference src/Pipeline.cpp:64-145 reads nothing from
previous frames), so a batch of F frames is split into contiguous blocks, one per rank, and there is NO
collective on the data path.  The only exchange is the final result gather - what the reference's frame loop
(cmd/main_VO.cpp:99-113) accumulates on one host - done here with torch.distributed (NCCL on GPUs, gloo in the CPU
tests): the per-frame mate counts first (one small all_gather), then every rank's mate records at their exact size
(point-to-point sends into the destination's buffer: no padding, no host bounce when the tensors are on the device).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, world: int, rank: int):
    """Contiguous block of ceil(F/G) frames per rank (last ranks may get fewer, or none)."""
    per = (n_frames + world - 1) // world
    lo = min(rank * per, n_frames)
    hi = min(lo + per, n_frames)
    return lo, hi


def gather_packed(packed, counts, n_frames: int, dist, dst: int = 0):
    """Final gather of a sharded batch.

    packed: torch uint8 tensor [n_local, 64], this rank's mate records back to back in frame order (ebvo_batch_pack output,
    on the device under NCCL; CPU tensors under gloo); counts: torch int32 tensor [f_local] on the same device.
    Returns (packed_all [n_total, 64], counts_all [n_frames]) in global frame order on rank `dst`, (None, None) elsewhere.
    """
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_frames + world - 1) // world
    dev = packed.device
    pad = torch.zeros(per, dtype=torch.int32, device=dev)
    pad[:counts.numel()] = counts
    allc = torch.empty(world * per, dtype=torch.int32, device=dev)
    dist.all_gather(list(allc.view(world, per).unbind(0)), pad)
    per_rank = allc.view(world, per).sum(dim=1).cpu().tolist()      # records held by every rank (the one host read of the exchange)
    if rank != dst:
        if per_rank[rank]:
            dist.send(packed[:per_rank[rank]].contiguous(), dst=dst)
        return None, None
    out = torch.empty((sum(per_rank), 64), dtype=torch.uint8, device=dev)
    pos = 0
    for r in range(world):
        n = per_rank[r]
        if n:
            if r == rank:
                out[pos:pos + n].copy_(packed[:n])
            else:
                dist.recv(out[pos:pos + n], src=r)
        pos += n
    return out, allc[:n_frames]


class HostGather:
    """The final gather when the consumer is ONE host process (MotionTracker's sequential pose estimation stays on the host,
    BASELINE.json north star): every rank copies its own mate records device -> host over ITS OWN PCIe link straight into
    its slice of one shared, page-locked host buffer (POSIX shared memory mapped by all ranks of the box and registered with
    CUDA), so the transfer runs on all links at once instead of funnelling 1.7 GB through rank 0's GPU and link; the only
    exchange between the ranks is the per-frame counts (one small all_gather) and a barrier.  Collective constructor."""

    def __init__(self, nbytes: int, dist, tag: str = "0", register: bool = True):
        import mmap
        import os
        import torch
        self.dist, self.nbytes = dist, int(nbytes)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.path = f"/dev/shm/ebvo_gather_{os.getuid()}_{os.environ.get('MASTER_PORT', '0')}_{tag}"
        if self.rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        dist.barrier()
        self.nbytes = os.path.getsize(self.path)          # rank 0's size is the buffer's size for every rank
        self._f = open(self.path, "r+b")
        self._mm = mmap.mmap(self._f.fileno(), self.nbytes)
        self.buf = torch.frombuffer(self._mm, dtype=torch.uint8)
        self._registered = False
        if register and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.buf.data_ptr(), self.nbytes, 0)
            self._registered = int(rc) == 0
        dist.barrier()

    def gather(self, packed, counts, n_frames: int, dst: int = 0):
        """packed: [n_local, 64] uint8 (device or CPU), counts: int32 [f_local] on the same device.  Returns (records
        [n_total, 64] - a view of the shared host buffer -, counts[n_frames]) on `dst`, (None, None) elsewhere."""
        import torch
        dist, world, rank = self.dist, self.world, self.rank
        per = (n_frames + world - 1) // world
        dev = packed.device
        pad = torch.zeros(per, dtype=torch.int32, device=dev)
        pad[:counts.numel()] = counts
        allc = torch.empty(world * per, dtype=torch.int32, device=dev)
        dist.all_gather(list(allc.view(world, per).unbind(0)), pad)
        allc_h = allc.cpu()
        per_rank = allc_h.view(world, per).sum(dim=1).tolist()
        n, off = per_rank[rank], sum(per_rank[:rank])
        if (off + n) * 64 > self.nbytes:
            raise RuntimeError("HostGather buffer too small")
        if n:
            self.buf[off * 64:(off + n) * 64].view(n, 64).copy_(packed[:n], non_blocking=self._registered)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        dist.barrier()
        if rank != dst:
            return None, None
        total = sum(per_rank)
        return self.buf[:total * 64].view(total, 64), allc_h[:n_frames]

    def region(self, n_frames: int, per_frame_records: int):
        """This rank's slice of the shared buffer for a STREAMED gather: (host address, capacity in records, record offset).
        A rank hands the address to ebvo_stereo_batch_packed, which copies every finished sub-batch's mates there while the
        next sub-batches compute; finish() then only exchanges the per-frame counts."""
        per = (n_frames + self.world - 1) // self.world
        cap = per * int(per_frame_records)
        if (self.world * cap) * 64 > self.nbytes:
            raise RuntimeError("HostGather buffer too small")
        off = self.rank * cap
        return self.buf.data_ptr() + off * 64, cap, off

    def finish(self, counts, n_frames: int, per_frame_records: int, device=None, dst: int = 0):
        """After every rank's ebvo_stereo_batch_packed into region(): counts = this rank's per-frame mate counts (int32 numpy
        or tensor).  Returns ([records of rank 0's frames, records of rank 1's frames, ...] - views of the shared buffer -,
        counts[n_frames]) on `dst`, (None, None) elsewhere."""
        import torch
        dist, world = self.dist, self.world
        per = (n_frames + world - 1) // world
        cap = per * int(per_frame_records)
        counts = torch.as_tensor(counts, dtype=torch.int32)
        pad = torch.zeros(per, dtype=torch.int32, device=device)
        pad[:counts.numel()] = counts.to(pad.device)
        allc = torch.empty(world * per, dtype=torch.int32, device=device)
        dist.all_gather(list(allc.view(world, per).unbind(0)), pad)
        dist.barrier()                                     # every rank's copies have landed (its call returned before its all_gather)
        if self.rank != dst:
            return None, None
        allc_h = allc.cpu()
        per_rank = allc_h.view(world, per).sum(dim=1).tolist()
        segs = [self.buf[r * cap * 64:(r * cap + n) * 64].view(n, 64) for r, n in enumerate(per_rank)]
        return segs, allc_h[:n_frames]

    def close(self):
        import os
        import torch
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.buf.data_ptr())
            self._registered = False
        self.dist.barrier()
        self.buf = None
        try:
            self._mm.close()
        except BufferError:
            pass
        self._f.close()
        if self.rank == 0 and os.path.exists(self.path):
            os.unlink(self.path)


def gather_mates(local_mates, local_counts, n_frames: int, dist=None, device=None):
    """Padded convenience form for host arrays: local_mates [f_local, cap] (MATE_DTYPE), local_counts [f_local] int32.
    Returns (mates[n_frames, cap], counts[n_frames]) on rank 0 and (None, None) elsewhere; the exchange itself is
    gather_packed (exact sizes)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_mates, local_counts
    cap = local_mates.shape[1]
    nloc = len(local_counts)
    rows = [local_mates[k, :local_counts[k]] for k in range(nloc)]
    flat = np.concatenate(rows) if rows else np.zeros(0, local_mates.dtype)
    tp = torch.from_numpy(np.ascontiguousarray(flat).view(np.uint8).reshape(-1, 64).copy())
    tc = torch.from_numpy(np.asarray(local_counts, np.int32).copy())
    if device is not None:
        tp, tc = tp.to(device), tc.to(device)
    allp, allc = gather_packed(tp, tc, n_frames, dist)
    if allp is None:
        return None, None
    allc = allc.cpu().numpy()
    rec = allp.cpu().numpy().reshape(-1).view(local_mates.dtype)
    out = np.zeros((n_frames, cap), local_mates.dtype)
    pos = 0
    for f in range(n_frames):
        out[f, :allc[f]] = rec[pos:pos + allc[f]]
        pos += allc[f]
    return out, allc

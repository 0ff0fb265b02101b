"""CPU, world_size 2, gloo: the frame-sharding host logic and the final result gather (the only exchange step)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edge_based_visual_odometry_b200 import _lib, sharding


def test_shard_ranges_cover_all_frames_once():
    for F in (1, 2, 7, 1000):
        for G in (1, 2, 4, 8):
            seen = []
            for r in range(G):
                lo, hi = sharding.shard_range(F, G, r)
                seen += list(range(lo, hi))
            assert seen == list(range(F))


def _worker(rank, world, port, F, cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(F, world, rank)
    m = np.zeros((hi - lo, cap), _lib.MATE_DTYPE)
    c = np.zeros(hi - lo, np.int32)
    for k, f in enumerate(range(lo, hi)):       # fake per-frame results that encode the global frame id
        c[k] = (f % cap) + 1
        m["left_index"][k, :c[k]] = f
        m["score"][k, :c[k]] = f + 0.5
    gm, gc = sharding.gather_mates(m, c, F, dist)
    if rank == 0:
        ok = all(gc[f] == (f % cap) + 1 and (gm["left_index"][f, :gc[f]] == f).all() and (gm["score"][f, :gc[f]] == f + 0.5).all()
                 for f in range(F))
        q.put(bool(ok) and gm.shape == (F, cap))
    else:
        assert gm is None
    dist.destroy_process_group()


def _worker_packed(rank, world, port, F, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(F, world, rank)
    counts = np.array([(3 * f) % 7 for f in range(lo, hi)], np.int32)      # some frames have no mates
    rec = np.zeros(int(counts.sum()), _lib.MATE_DTYPE)
    pos = 0
    for k, f in enumerate(range(lo, hi)):
        rec["left_index"][pos:pos + counts[k]] = 1000 * f + np.arange(counts[k])
        pos += counts[k]
    tp = torch.from_numpy(rec.view(np.uint8).reshape(-1, 64).copy())
    allp, allc = sharding.gather_packed(tp, torch.from_numpy(counts), F, dist)
    if rank == 0:
        want_c = np.array([(3 * f) % 7 for f in range(F)], np.int32)
        got = allp.numpy().reshape(-1).view(_lib.MATE_DTYPE)["left_index"]
        want = np.concatenate([1000 * f + np.arange(want_c[f]) for f in range(F)])
        q.put(bool(np.array_equal(allc.numpy(), want_c) and np.array_equal(got, want) and allp.shape == (want_c.sum(), 64)))
    else:
        assert allp is None
    dist.destroy_process_group()


def _worker_host(rank, world, port, F, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(F, world, rank)
    counts = np.array([(3 * f) % 7 for f in range(lo, hi)], np.int32)
    rec = np.zeros(int(counts.sum()), _lib.MATE_DTYPE)
    pos = 0
    for k, f in enumerate(range(lo, hi)):
        rec["left_index"][pos:pos + counts[k]] = 1000 * f + np.arange(counts[k])
        pos += counts[k]
    hg = sharding.HostGather(64 * 64, dist, tag="t", register=False)
    for _ in range(2):          # reusable
        allp, allc = hg.gather(torch.from_numpy(rec.view(np.uint8).reshape(-1, 64).copy()), torch.from_numpy(counts), F)
    if rank == 0:
        want_c = np.array([(3 * f) % 7 for f in range(F)], np.int32)
        got = allp.numpy().copy().reshape(-1).view(_lib.MATE_DTYPE)["left_index"]
        want = np.concatenate([1000 * f + np.arange(want_c[f]) for f in range(F)])
        q.put(bool(np.array_equal(allc.numpy(), want_c) and np.array_equal(got, want)))
    del allp
    hg.close()
    dist.destroy_process_group()


def _worker_stream(rank, world, port, F, q):
    import ctypes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(F, world, rank)
    counts = np.array([(3 * f) % 7 for f in range(lo, hi)], np.int32)
    rec = np.zeros(int(counts.sum()), _lib.MATE_DTYPE)
    pos = 0
    for k, f in enumerate(range(lo, hi)):
        rec["left_index"][pos:pos + counts[k]] = 1000 * f + np.arange(counts[k])
        pos += counts[k]
    per_frame = 8
    per = -(-F // world)
    hg = sharding.HostGather(world * per * per_frame * 64, dist, tag="s", register=False)
    addr, cap, off = hg.region(F, per_frame)
    ok = cap == per * per_frame and off == rank * cap
    ctypes.memmove(addr, rec.ctypes.data, rec.nbytes)      # what ebvo_stereo_batch_packed does with device -> host copies
    segs, allc = hg.finish(counts, F, per_frame)
    if rank == 0:
        want_c = np.array([(3 * f) % 7 for f in range(F)], np.int32)
        got = np.concatenate([sg.numpy().copy().reshape(-1).view(_lib.MATE_DTYPE)["left_index"] for sg in segs])
        want = np.concatenate([1000 * f + np.arange(want_c[f]) for f in range(F)])
        q.put(bool(ok and len(segs) == world and np.array_equal(allc.numpy(), want_c) and np.array_equal(got, want)))
    del segs
    hg.close()
    dist.destroy_process_group()


def test_host_gather_streamed_regions_world2_gloo():
    """The streamed gather: every rank writes its records densely into ITS region of the shared buffer (on a GPU box:
    ebvo_stereo_batch_packed, sub-batch by sub-batch during the computation); finish() exchanges only the counts and rank 0
    reads the regions in rank order = global frame order."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_stream, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_host_gather_shared_memory_world2_gloo():
    """sharding.HostGather: every rank writes its records into its slice of one shared host buffer (on a GPU box: its own
    device -> host copy over its own PCIe link); rank 0 reads all of them in global frame order."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_host, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_gather_packed_world2_gloo_exact_sizes():
    """The exchange of a sharded batch: counts, then every rank's records at their exact size, global frame order kept
    (7 frames over 2 ranks: blocks of 4 and 3; frame 0 and frame 7k have no mates)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_gather_mates_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

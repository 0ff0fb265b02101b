"""Multi-GPU paths on real devices (skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

BASELINE.json configs[2]: a batch of stereo pairs partitioned across the GPUs of one box with only a final result gather.
"""
import os
import socket

import numpy as np
import pytest
import torch

from edge_based_visual_odometry_b200 import _lib, sharding, synth

pytestmark = pytest.mark.gpu
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


def _calib(cal):
    return _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)


def _frames(n):
    cal = synth.kitti_calib(480, 200)
    pairs = [synth.stereo_pair(cal, f) for f in range(3)]
    return cal, [pairs[f % 3][0] for f in range(n)], [pairs[f % 3][1] for f in range(n)]


@needs2
def test_batch_multi_over_real_devices():
    """ebvo_stereo_batch_multi with one context per physical GPU: the blocks run on different devices, the mates land at
    their global frame index and equal the single-GPU result record for record."""
    ndev = torch.cuda.device_count()
    cal, Ls, Rs = _frames(2 * ndev + 1)
    ctxs = [_lib.Context(d, 480, 200, max_batch=4, max_edges=16384) for d in range(ndev)]
    one = _lib.Context(0, 480, 200, max_batch=len(Ls), max_edges=16384)
    ref, nref = one.stereo_batch(_calib(cal), Ls, Rs, cap=12000)
    out, n = _lib.stereo_batch_multi(ctxs, _calib(cal), Ls, Rs, cap=12000)
    for c in ctxs + [one]:
        c.close()
    assert np.array_equal(n, nref) and n.min() > 500
    for f in range(len(Ls)):
        assert np.array_equal(out[f, :n[f]], ref[f, :n[f]])


def _worker(rank, world, port, F, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cal, Ls, Rs = _frames(F)
    lo, hi = sharding.shard_range(F, world, rank)
    ctx = _lib.Context(rank, 480, 200, max_batch=max(1, hi - lo), max_edges=16384)
    dev = torch.device("cuda", rank)
    nm = ctx.stereo_batch_device(_calib(cal), Ls[lo:hi], Rs[lo:hi])
    packed = torch.empty((int(nm.sum()) + 16, 64), dtype=torch.uint8, device=dev)
    offs = torch.empty(hi - lo + 1, dtype=torch.int32, device=dev)
    tot = ctx.batch_pack(packed.data_ptr(), packed.shape[0], offs.data_ptr())
    assert tot == int(nm.sum()) and offs.cpu().numpy()[-1] == tot
    allp, allc = sharding.gather_packed(packed[:tot], torch.from_numpy(nm.copy()).to(dev), F, dist)
    # the same gather into one shared page-locked host buffer, every rank over its own PCIe link (sharding.HostGather)
    hg = sharding.HostGather(64 * 200000, dist, tag="t")
    hp, hc = hg.gather(packed[:tot], torch.from_numpy(nm.copy()).to(dev), F)
    if rank == 0:
        assert hg._registered and np.array_equal(hp.numpy(), allp.cpu().numpy()) and np.array_equal(hc.numpy(), allc.cpu().numpy())
    del hp
    hg.close()
    # the streamed gather: ebvo_stereo_batch_packed writes this rank's records into its region of the shared buffer during the computation
    per_frame = 12000
    hs = sharding.HostGather(2 * (-(-F // world)) * per_frame * 64, dist, tag="s")
    addr, cap, _ = hs.region(F, per_frame)
    nm2, wrote = ctx.stereo_batch_packed(_calib(cal), Ls[lo:hi], Rs[lo:hi], addr, cap)
    segs, sc = hs.finish(nm2, F, per_frame, device=dev)
    if rank == 0:
        got = np.concatenate([sg.numpy().copy() for sg in segs])
        assert hs._registered and np.array_equal(got, allp.cpu().numpy()) and np.array_equal(sc.numpy(), allc.cpu().numpy())
    assert wrote == int(nm.sum()) and np.array_equal(nm2, nm)
    del segs
    hs.close()
    if rank == 0:
        one = _lib.Context(0, 480, 200, max_batch=F, max_edges=16384)
        ref, nref = one.stereo_batch(_calib(cal), Ls, Rs, cap=12000)
        one.close()
        rec = allp.cpu().numpy().reshape(-1).view(_lib.MATE_DTYPE)
        want = np.concatenate([ref[f, :nref[f]] for f in range(F)])
        q.put(bool(np.array_equal(allc.cpu().numpy(), nref) and np.array_equal(rec, want)))
    ctx.close()
    dist.destroy_process_group()


@needs2
def test_sharded_batch_nccl_gather_two_gpus():
    """One process per GPU (NCCL): every rank matches its block of the batch, packs the mates on the device
    (ebvo_batch_pack) and the final gather (sharding.gather_packed: counts, then exact-size device-to-device sends)
    reproduces the single-GPU result in global frame order."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_packed_batch_call_equals_the_padded_one():
    """ebvo_stereo_batch_packed on one GPU: host images in, the batch's mates back to back in (page-locked) host memory,
    equal to ebvo_stereo_batch's padded output frame by frame; a buffer that is too small reports EBVO_ERR_CAPACITY."""
    cal, Ls, Rs = _frames(37)                      # three sub-batches of the 16-frame pipeline, the last one ragged
    ctx = _lib.Context(0, 480, 200, max_batch=37, max_edges=16384)
    ref, nref = ctx.stereo_batch(_calib(cal), Ls, Rs, cap=12000)
    buf = torch.empty((int(nref.sum()) + 5, 64), dtype=torch.uint8).pin_memory()
    nm, wrote = ctx.stereo_batch_packed(_calib(cal), Ls, Rs, buf.data_ptr(), buf.shape[0])
    assert wrote == int(nref.sum()) and np.array_equal(nm, nref)
    rec = buf[:wrote].numpy().reshape(-1).view(_lib.MATE_DTYPE)
    o = np.concatenate([[0], np.cumsum(nref)])
    for f in range(37):
        assert np.array_equal(rec[o[f]:o[f + 1]], ref[f, :nref[f]])
    with pytest.raises(_lib.EbvoError):
        ctx.stereo_batch_packed(_calib(cal), Ls, Rs, buf.data_ptr(), int(nref.sum()) - 1)
    ctx.close()


def test_batch_pack_matches_download():
    """ebvo_batch_pack on one GPU: the packed records and offsets equal the padded download frame by frame."""
    cal, Ls, Rs = _frames(5)
    ctx = _lib.Context(0, 480, 200, max_batch=5, max_edges=16384)
    nm = ctx.stereo_batch_device(_calib(cal), Ls, Rs)
    dev = torch.device("cuda", 0)
    packed = torch.empty((int(nm.sum()) + 8, 64), dtype=torch.uint8, device=dev)
    offs = torch.empty(6, dtype=torch.int32, device=dev)
    tot = ctx.batch_pack(packed.data_ptr(), packed.shape[0], offs.data_ptr())
    out, n = ctx.batch_download(12000)
    ctx.close()
    o = offs.cpu().numpy()
    rec = packed[:tot].cpu().numpy().reshape(-1).view(_lib.MATE_DTYPE)
    assert tot == int(n.sum()) and np.array_equal(n, nm) and np.array_equal(np.diff(o), n)
    for f in range(5):
        assert np.array_equal(rec[o[f]:o[f + 1]], out[f, :n[f]])

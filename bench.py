#!/usr/bin/env python
"""Benchmark of the TOED + stereo-NCC hot path (BASELINE.json metric: stereo frames/s at KITTI 1241x376).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one pass of the whole hot path (TOED on both views + stereo matching S1..S13, SIFT-off) over
one batch of B DISTINCT synthetic KITTI-shape stereo pairs per GPU.  `value` is measured with the batch resident in
HBM (CUDA events on the library's stream); `e2e` goes through the C-ABI batch call with pinned HOST buffers
(H2D of the images and D2H of the mates inside the timed region).  Frames are independent, so N GPUs each
process their own batch (weak scaling, no data-path collective); rank 0 prints ONE JSON line.  The same line carries
`strong_scaling`: BASELINE.json configs[2] as written - ONE 1000-frame batch split over the N ranks
(sharding.shard_range), host images in, and the final gather of the unpadded mate records to rank 0 (device to device
over NCCL, then one D2H) INSIDE the timed region.

`--impl reference` times the reference's own CPU implementation on the host cores: the unmodified reference
TOED compiled in place (oracle/_ref) + the C++ port of the stereo stage (the reference stereo sources need
OpenCV/Eigen and cannot be compiled here), one frame per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP32_LANES_PER_SM = 128
TOED_FLOP_PER_PX = 1636.0      # 818 MAC per input pixel, separable form
GN_FLOP_PER_ITER = 5300.0      # 98 samples x 54 FP64 flop per Gauss-Newton iteration
GN_NCU_SUMMARY = os.path.join("profiles", "r02_ncu_gn_lerp64.txt")   # ncu --set full capture of the dominant kernel (committed summary)
FMA_PEAKS = os.path.join("profiles", "r02_fma_peaks.json")           # scripts/ubench/fma_peak.cu on this pool's B200


def read_fma_peaks():
    """Measured DFMA / FFMA throughput (TFLOP/s) from the committed microbenchmark record; None when it is missing."""
    try:
        d = json.load(open(os.path.join(ROOT, FMA_PEAKS)))
        return float(d["fp64_fma_tflops"]), float(d["fp32_fma_tflops"])
    except Exception:
        return None, None


def read_gn_traffic_per_frame():
    """DRAM bytes (read + write) per frame of one launch of the Gauss-Newton kernel, parsed from the committed ncu summary;
    None when the summary is missing (no constant is substituted)."""
    import re
    try:
        txt = open(os.path.join(ROOT, GN_NCU_SUMMARY)).read()
        frames = int(re.search(r"\((\d+) KITTI-shape frames per launch", txt).group(1))
        rd = float(re.search(r"dram__bytes_read\.sum\s+([0-9.]+) Mbyte", txt).group(1))
        wr = float(re.search(r"dram__bytes_write\.sum\s+([0-9.]+) Mbyte", txt).group(1))
        return (rd + wr) * 1e6 / frames
    except Exception:
        return None


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def _gen_pair(a):
    """(workload, seed, density) -> cached synthetic stereo pair (uint8 L, R); generated once per box into /tmp."""
    workload, seed, density = a
    from edge_based_visual_odometry_b200 import synth
    cal = {"kitti": synth.kitti_calib, "euroc": synth.CALIBS["euroc"], "4k": synth.kitti4k_calib}[workload]()
    d = os.path.join("/tmp", "ebvo_synth_cache")
    os.makedirs(d, exist_ok=True)
    f = os.path.join(d, f"{workload}_{cal.width}x{cal.height}_s{seed}_d{density:g}.npz")
    if os.path.exists(f):
        try:
            z = np.load(f)
            return z["L"], z["R"]
        except Exception:
            pass
    L, R = synth.stereo_pair(cal, seed, density=density)
    tmp = f + f".{os.getpid()}.tmp.npz"
    np.savez(tmp, L=L, R=R)
    os.replace(tmp, f)
    return L, R


def synth_frames(workload, seeds, density, world):
    """Distinct synthetic frames (seed = frame id), generated in parallel on the host cores this rank may use and cached in /tmp."""
    jobs = [(workload, int(sd), float(density)) for sd in seeds]
    workers = max(1, min(len(jobs), (os.cpu_count() or 4) // max(1, world)))
    if workers == 1 or len(jobs) < 4:
        return [_gen_pair(j) for j in jobs]
    import concurrent.futures as cf
    import multiprocessing as mp
    with cf.ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) as ex:
        return list(ex.map(_gen_pair, jobs, chunksize=max(1, len(jobs) // (4 * workers))))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        try:      # NVML directly (a few ms per sample); nvidia-smi below as the fallback
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append([str(sm), str(mx), "%.1f" % pw] + ["Active" if r & bits[k] else "Not Active" for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                time.sleep(0.05)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for k, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(len(r) > k and r[k].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_frame(cal, L, R, threads=0, stereo_ref=True):
    """One frame through the CPU path: reference TOED (oracle/_ref) x2 + stereo port.  Returns seconds and parts."""
    import oracle
    t0 = time.perf_counter()
    if oracle.have_ref():
        eL, _, _, _ = oracle.toed_reference(L, threads)
        eR, _, _, _ = oracle.toed_reference(R, threads)
        kind = "reference"
    else:
        eL, _ = oracle.toed(L)
        eR, _ = oracle.toed(R)
        kind = "port"
    t1 = time.perf_counter()
    if stereo_ref and oracle.have_stereo_ref():
        res = oracle.stereo_reference(L, R, eL, eR, cal.Kl, cal.Kr, cal.R21, cal.T21)
    else:
        F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
        res = oracle.stereo(L, R, eL, eR, F21, want_dumps=False, threads=threads)
        kind = "port" if kind == "port" else "reference TOED + stereo port"
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1, kind, len(res.mate_left)


def run_reference(args, cal, rank):
    from edge_based_visual_odometry_b200 import synth
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = [synth.stereo_pair(cal, f) for f in range(2)]
    for w in range(min(args.warmup, 1)):          # one untimed frame is enough to page the libraries in
        cpu_reference_frame(cal, *frames[w % 2])
    t0 = time.perf_counter()
    kind = "port"
    for k in range(args.steps):
        _, _, _, kind, _ = cpu_reference_frame(cal, *frames[k % 2])
    dt = time.perf_counter() - t0
    v = args.steps / dt
    line = {"impl": "reference", "metric": "stereo frames/s (TOED+NCC stereo match) at KITTI 1241x376", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "KITTI-shape 1241x376 synthetic stereo pairs, 1 frame per step, SIFT-off", "frames_per_step": 1},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores,
                             "kind": "reference" if kind == "reference" else "port",
                             "sample": "1 frame per step: TOED = unmodified reference cpu_toed.cpp on both views; stereo S1-S13 (SIFT-off) = "
                                       "unmodified reference Stereo_Matches.cpp/utility.cpp/EdgeClusterer.cpp compiled in place "
                                       "(oracle/_ref, OpenCV/Eigen primitives from third_party_shim), OpenMP on all host cores"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="stereo frames per GPU per step (default 160; 4 for the 4K workload)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--distinct", type=int, default=0, help="distinct synthetic frames per GPU (default: the batch size, i.e. every frame of the "
                    "batch is its own scene; smaller values are cycled to fill the batch)")
    ap.add_argument("--strong-frames", type=int, default=1000, help="frames of the ONE batch that the strong-scaling leg splits over the ranks (0 = skip)")
    ap.add_argument("--strong-gather", default="stream", choices=["stream", "host", "nccl"],
                    help="final gather of the strong-scaling leg: stream = every rank's ebvo_stereo_batch_packed copies each finished sub-batch's records "
                         "over its own PCIe link into its slice of one shared page-locked host buffer while its next sub-batches compute "
                         "(sharding.HostGather.region / finish); host = the same buffer, but one copy per rank after its last kernel; "
                         "nccl = device-to-device sends to rank 0, then one D2H (sharding.gather_packed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="kitti", choices=["kitti", "euroc", "4k"],
                    help="kitti = the BASELINE.json metric (default); euroc / 4k = BASELINE configs[1] / configs[4] shapes (extra measurements)")
    ap.add_argument("--density", type=float, default=1.0, help="synthetic scene density (objects per area), the 4K stress sweep varies it")
    ap.add_argument("--sift", action="store_true", help="SIFT-on: descriptors, SIFT gate and BNB-SIFT on the device (sift_mode 1)")
    ap.add_argument("--gn-mode", type=int, default=0, help="Gauss-Newton kernel: 0 FP64 tiled (default), 1 FP64 gather, 2 FP32")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from edge_based_visual_odometry_b200 import synth
    cal = {"kitti": synth.kitti_calib, "euroc": synth.CALIBS["euroc"], "4k": synth.kitti4k_calib}[args.workload]()
    if not args.batch:
        args.batch = 4 if args.workload == "4k" else 160
    if args.impl == "reference":
        run_reference(args, cal, rank)
        return

    import torch
    import torch.distributed as dist
    from edge_based_visual_odometry_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    B = args.batch
    H, W = cal.height, cal.width
    big = args.workload == "4k"
    max_edges = (1 << 21) if big else 65536
    CAP = (1 << 20) if big else 49152   # mates per frame returned to the host
    # distinct frames per rank (different seeds per rank), cycled to fill the batch
    if not args.distinct:
        args.distinct = B
    if args.workload == "4k":
        args.distinct = min(args.distinct, 2)
    base = synth_frames(args.workload, [rank * 1000 + f for f in range(min(args.distinct, B))], args.density, world)
    # pinned host staging for the e2e path
    hL = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    hR = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    for f in range(B):
        hL[f] = torch.from_numpy(base[f % len(base)][0])
        hR[f] = torch.from_numpy(base[f % len(base)][1])
    nL_imgs = [hL[f].numpy() for f in range(B)]
    nR_imgs = [hR[f].numpy() for f in range(B)]
    h_out = torch.empty((B, CAP, 64), dtype=torch.uint8).pin_memory()
    out_np = h_out.numpy().view(_lib.MATE_DTYPE).reshape(B, CAP)
    n_mates = np.zeros(B, np.int32)

    prm = _lib.default_params()
    prm.gn_mode = args.gn_mode
    prm.sift_mode = 1 if args.sift else 0
    ctx = _lib.Context(local_rank, W, H, max_batch=B, max_edges=max_edges, params=prm)
    calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    ctx.batch_upload(nL_imgs, nR_imgs)
    ctx.batch_sync()
    for _ in range(args.warmup):
        ctx.batch_run(calib, True)
    ctx.batch_sync()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.batch_run(calib, True)
    ev1.record(stream)
    ctx.batch_sync()
    barrier()
    sampler.stop_flag = True
    ms_total = ev0.elapsed_time(ev1)
    gpu_launches = ctx.launch_count() - launches0
    nL, nR, nM, counters = ctx.batch_counts()

    # ---------------- per-kernel durations: the same K steps again with CUDA events around every launch ----------------
    # (the batch call runs its two halves on two streams so that kernel tails overlap; with per-kernel events enabled it stays
    #  on one stream, so the durations below are exclusive - the roofline of the dominant kernel is computed from them)
    ctx.set_profiling(True)
    ctx.batch_run(calib, True)
    ctx.batch_sync()
    ctx.set_profiling(False); ctx.set_profiling(True)      # reset the accumulators after one untimed pass
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record(stream)
    for _ in range(args.steps):
        ctx.batch_run(calib, True)
    evp1.record(stream)
    ctx.batch_sync()
    ms_profiled = evp0.elapsed_time(evp1)
    ktimes = ctx.kernel_times()
    ctx.set_profiling(False)

    # ---------------- end to end through the C ABI with host buffers (`e2e`) ----------------
    for _ in range(2):
        ctx.stereo_batch(calib, nL_imgs, nR_imgs, CAP, out_np, n_mates)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.stereo_batch(calib, nL_imgs, nR_imgs, CAP, out_np, n_mates)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---------------- strong scaling: ONE batch of --strong-frames frames split over the ranks, final gather timed ----------------
    strong = None
    if args.strong_frames > 0 and args.workload == "kitti" and not args.sift:
        from edge_based_visual_odometry_b200 import sharding
        FS = args.strong_frames
        lo, hi = sharding.shard_range(FS, world, rank)
        nloc = hi - lo
        dev = torch.device("cuda", local_rank)
        pf = torch.tensor([float(nM.mean()) * 1.25 + 1024], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pf, op=dist.ReduceOp.MAX)      # one capacity for every rank
        per_frame = float(pf.item())
        cap_rec = int(max(nloc, 1) * per_frame)
        packed = torch.empty((cap_rec, 64), dtype=torch.uint8, device=dev)
        offs = torch.empty(B + 1, dtype=torch.int32, device=dev)
        mode = args.strong_gather
        if world > 1 and mode in ("stream", "host"):
            # the shared page-locked buffer lives in /dev/shm: fall back to the NCCL gather when the box does not have the room
            need = world * (-(-FS // world)) * int(per_frame) * 64
            try:
                st = os.statvfs("/dev/shm")
                room = st.f_bavail * st.f_frsize
            except OSError:
                room = 0
            ok = torch.tensor([1 if room > need + (64 << 20) else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                mode = "nccl"
        use_host = world > 1 and mode == "host"
        use_stream = mode == "stream"
        pf_rec = int(per_frame)
        h_res = torch.empty((int(FS * per_frame), 64), dtype=torch.uint8).pin_memory() if (rank == 0 and not use_host and not (use_stream and world > 1)) else None
        hg = sharding.HostGather(world * (-(-FS // world)) * pf_rec * 64, dist, tag="bench") if (world > 1 and (use_host or use_stream)) else None

        def strong_pass():
            if use_stream:
                # every sub-batch's records go to host memory (this rank's slice of the shared buffer) while the next ones compute
                if hg is not None:
                    addr, cap, _ = hg.region(FS, pf_rec)
                else:
                    addr, cap = h_res.data_ptr(), h_res.shape[0]
                pos, counts = 0, []
                for s0 in range(0, nloc, B):
                    n = min(B, nloc - s0)
                    nm, wrote = ctx.stereo_batch_packed(calib, nL_imgs[:n], nR_imgs[:n], addr + pos * 64, cap - pos)
                    pos += wrote
                    counts.append(nm.copy())
                tg = time.perf_counter()
                nrec = pos
                if hg is not None:
                    segs, allc = hg.finish(np.concatenate(counts) if counts else np.zeros(0, np.int32), FS, pf_rec, device=dev)
                    if rank == 0:
                        nrec = int(sum(int(sg.shape[0]) for sg in segs))
                return time.perf_counter() - tg, nrec
            pos, counts = 0, []
            for s0 in range(0, nloc, B):          # this rank's block, in sub-batches of the context's capacity
                n = min(B, nloc - s0)
                nm = ctx.stereo_batch_device(calib, nL_imgs[:n], nR_imgs[:n])
                pos += ctx.batch_pack(packed[pos:].data_ptr(), cap_rec - pos, offs.data_ptr())
                counts.append(torch.from_numpy(nm.copy()))
            ct = (torch.cat(counts) if counts else torch.zeros(0, dtype=torch.int32)).to(dev)
            torch.cuda.synchronize()
            tg = time.perf_counter()
            nrec = 0
            if use_host:
                allp, allc = hg.gather(packed[:pos], ct, FS)
                if rank == 0:
                    nrec = int(allp.shape[0])
            else:
                if world > 1:
                    allp, allc = sharding.gather_packed(packed[:pos], ct, FS, dist)
                else:
                    allp, allc = packed[:pos], ct
                if rank == 0:
                    nrec = int(allp.shape[0])
                    h_res[:nrec].copy_(allp, non_blocking=True)
                    allc.cpu()
            torch.cuda.synchronize()
            return time.perf_counter() - tg, nrec

        strong_pass()                              # untimed: NCCL point-to-point connections, allocator
        barrier()
        t0 = time.perf_counter()
        reps, g_s, nrec = 2, 0.0, 0
        for _ in range(reps):
            gs, nrec = strong_pass()
            g_s += gs
        barrier()
        strong = [(time.perf_counter() - t0) / reps, g_s / reps, nrec,
                  "stream" if use_stream else ("host" if use_host else ("nccl" if world > 1 else "single device: one D2H"))]
        if hg is not None:
            hg.close()
        del packed, h_res

    t = torch.tensor([ms_total, e2e_s * 1e3, (strong[0] if strong else 0.0) * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, strong_ms = t.tolist()
    sampler.join(timeout=2)

    if rank == 0:
        peaks, peak_src = read_peaks()
        frames_total = world * B * args.steps
        value = frames_total / (ms_total / 1e3)
        e2e_value = frames_total / (e2e_ms / 1e3)
        # dominant kernel + roofline.  Algorithmic work per unit (DESIGN.md section 5 / SURVEY.md 8(d)):
        #   GN    5.3 kflop FP64 per Gauss-Newton iteration (98 samples x 54 flop: 3 four-corner blends, weights, residual,
        #         Huber-weighted normal equations), all in double as the reference computes it -> bound = FP64 pipe
        #   TOED  1636 flop FP32 per input pixel (818 MAC, separable dense form) x W*H x 2 views      -> bound = FP32 pipe
        #   NCC   2.3 kflop per scored pair
        # No step is a dense contraction (tensor cores unused) and HBM traffic is far below the compute time of every
        # kernel (the per-frame working set is L2-resident): `hbm` gives the byte side, `traffic` the DRAM bytes ncu measured.
        ksum = sum(v[0] for v in ktimes.values())
        dom = max(ktimes.items(), key=lambda kv: kv[1][0])
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        fp32_nom = sms * FP32_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12       # TFLOP/s, nominal FP32 FMA peak at max clock
        m64, m32 = read_fma_peaks()                                        # measured on this pool (scripts/ubench/fma_peak.cu)
        fp32_peak = m32 if m32 else fp32_nom
        fp64_peak = m64 if m64 else fp32_nom / 2                           # nominal: 64 FP64 lanes per SM per clock
        fma_src = (f"MEASURED FMA throughput, {FMA_PEAKS} (scripts/ubench/fma_peak.cu: DFMA {m64} / FFMA {m32} TFLOP/s; nominal "
                   f"{fp32_nom / 2:.1f} / {fp32_nom:.1f} at {sm_max:.0f} MHz)") if m64 and m32 else \
                  f"nominal FMA peak = {sms} SM x 64 (FP64) / 128 (FP32) lanes x 2 x sm_max_mhz ({peak_src} MEASURED_PEAKS.json clock); {FMA_PEAKS} missing"
        gn_traffic = read_gn_traffic_per_frame()
        c = counters.sum(axis=0)
        toed_ms = ktimes.get("toed_grad_nms", (0, 1))[0] + ktimes.get("toed_refine", ktimes.get("toed_orient", (0, 1)))[0]
        toed_flops = TOED_FLOP_PER_PX * W * H * 2 * B * args.steps
        gn_name = next((k for k in ("gn", "gn64", "gn32") if k in ktimes), None)
        gn_ms = ktimes[gn_name][0] if gn_name else 0.0
        gn_flops = GN_FLOP_PER_ITER * float(c[3]) * args.steps
        gn_peak = fp32_peak if gn_name == "gn32" else fp64_peak
        ncc_ms = ktimes.get("patch", (0, 1))[0] + ktimes.get("ncc_bnb", (0, 1))[0]
        ncc_flops = 2300.0 * float(c[0]) * args.steps

        def tf(fl, ms):
            return fl / (ms / 1e3) / 1e12 if ms else None

        kernels = {k: {"ms_per_step": v[0] / args.steps, "share": v[0] / ksum if ksum else None, "launches": v[1]} for k, v in ktimes.items()}
        # algorithmic bytes of the matching stage per step (SURVEY.md 8(d) formula, this run's counts)
        match_bytes = (2 * W * H * 2 + 2 * 4 * W * H) * B + 24.0 * (nL.sum() + nR.sum()) + 4.0 * c[0] + 64.0 * nM.sum()
        dom_is_gn = dom[0] == gn_name
        ach = tf(gn_flops, gn_ms) if dom_is_gn else tf(toed_flops, toed_ms)
        peak = gn_peak if dom_is_gn else fp32_peak
        roof = {"kernel": dom[0] if dom_is_gn else "toed_grad_nms+toed_refine", "bound": ("fp64" if gn_peak == fp64_peak else "fp32") if dom_is_gn else "fp32",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak if ach else None,
                # DRAM bytes of one launch of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu
                # --set full capture (parsed from its summary, per frame), scaled to this batch; null when no summary is present
                "traffic": gn_traffic * B if (dom_is_gn and gn_traffic and gn_name == "gn") else None,
                "traffic_source": GN_NCU_SUMMARY if (dom_is_gn and gn_traffic and gn_name == "gn") else None,
                "algorithmic_flop_per_launch": gn_flops / args.steps if dom_is_gn else toed_flops / args.steps,
                "avg_launch_ms": (gn_ms if dom_is_gn else toed_ms) / args.steps,
                "peak_source": fma_src + "; neither HBM nor tensor bound: no dense contraction on this path (tensor cores unused) and the "
                               "kernel's DRAM traffic is 0.2 % of HBM peak",
                "dominant_kernel_by_time": dom[0], "dominant_kernel_share": dom[1][0] / ksum if ksum else None,
                "per_stage": {"toed": {"bound": "fp32", "algorithmic_flop_per_px": TOED_FLOP_PER_PX, "achieved_tflops": tf(toed_flops, toed_ms),
                                       "frac": (tf(toed_flops, toed_ms) or 0) / fp32_peak},
                              "gauss_newton": {"bound": "fp64" if gn_peak == fp64_peak else "fp32", "algorithmic_flop_per_iteration": GN_FLOP_PER_ITER,
                                               "achieved_tflops": tf(gn_flops, gn_ms), "frac": (tf(gn_flops, gn_ms) or 0) / gn_peak},
                              "ncc": {"bound": "fp32", "algorithmic_flop_per_pair": 2300.0, "achieved_tflops": tf(ncc_flops, ncc_ms),
                                      "frac": (tf(ncc_flops, ncc_ms) or 0) / fp32_peak}},
                "hbm": {"peak_gbs": peaks.get("hbm_gbs"), "matching_algorithmic_bytes_per_step": float(match_bytes),
                        "matching_algorithmic_gbs": float(match_bytes) / ((ksum - toed_ms) / args.steps / 1e3) / 1e9 if ksum > toed_ms else None}}
        shape = {"kitti": "KITTI 1241x376", "euroc": "EuRoC 752x480", "4k": "4K 3840x2160"}[args.workload]
        cfgname = {"kitti": "configs[2]: KITTI-shape 1241x376 synthetic stereo batch", "euroc": "configs[1]: EuRoC-shape 752x480 synthetic stereo sequence (euroc.yaml calibration, general F)",
                   "4k": "configs[4]: 4K 3840x2160 synthetic stereo stress, density %.2f" % args.density}[args.workload]
        line = {"metric": "stereo frames/s (TOED+NCC stereo match) at " + shape, "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (TOED) / f64 (matching)", "data": "synthetic",
                "config": {"workload": cfgname + ", TOED x2 + stereo match S1-S13 (" + ("SIFT-on, device descriptors" if args.sift else "SIFT-off: the S4 / S7' SIFT stages need cv::SIFT, which is not part of the reference tree; --sift runs them on device descriptors") + ")",
                           "frames_per_gpu_per_step": B, "distinct_frames": len(base), "l2": "batch inputs (%.0f MB u8 images + per-frame "
                           "intermediates) exceed the 126 MB L2" % (2 * B * W * H / 1e6),
                           "edges_per_image": float(nL.mean()), "mates_per_frame": float(nM.mean())},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(2 * B * W * H), "d2h_bytes_per_step": int(64 * nM.sum() + 4 * B)},
                "gpu_launches": int(gpu_launches),
                "clocks": sampler.summary(), "roofline": roof, "kernels": kernels,
                "kernels_note": "per-kernel durations come from a second pass of the same K steps with CUDA events around every launch, on one "
                                "stream (%.1f ms per step); the timed region of `value` runs the batch as two slices on two streams so that kernel "
                                "tails overlap (%.1f ms per step)" % (ms_profiled / args.steps, ms_total / args.steps),
                "work_per_step": {"s3_pairs": int(c[0]), "bnb_pairs": int(c[1]), "gn_pairs": int(c[2]), "gn_iterations": int(c[3]), "ncc2_pairs": int(c[4]), "gn_tile_builds": int(c[5])}}
        if world == 1 and not args.no_cpu_baseline and args.workload == "kitti" and not args.sift:
            secs, ttoed, tst, kind, nm = cpu_reference_frame(cal, *base[0])
            secs2, ttoed2, tst2, _, _ = cpu_reference_frame(cal, *base[1 % len(base)])
            _, _, tport, _, _ = cpu_reference_frame(cal, *base[0], stereo_ref=False)
            tot = secs + secs2
            line["cpu_baseline"] = {"value": 2.0 / tot, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference" if kind == "reference" else "port",
                                    "sample": "2 frames of the same workload: reference TOED (oracle/_ref, OpenMP all cores) %.2f s + reference "
                                              "stereo sources compiled in place (oracle/_ref, shimmed OpenCV/Eigen) %.2f s per frame; the leaner "
                                              "C++ port of the stereo stage takes %.2f s" % ((ttoed + ttoed2) / 2, (tst + tst2) / 2, tport)}
        if strong:
            FS = args.strong_frames
            line["strong_scaling"] = {
                "workload": "configs[2]: ONE %d-frame KITTI-shape batch split into contiguous blocks of ceil(F/N) frames per rank "
                            "(sharding.shard_range; %d distinct scenes per rank cycled), host images in (pipelined H2D), every mate record gathered "
                            "into host memory of rank 0's process - %s - all inside the timed region" % (FS, len(base), {
                            "stream": "ebvo_stereo_batch_packed: each rank copies every finished sub-batch's records over its own PCIe link into its slice of "
                            "one shared page-locked buffer (sharding.HostGather.region) while its next sub-batches compute; after the last copy only the "
                            "per-frame counts are exchanged (one all_gather + barrier: gather_ms_rank0)",
                            "host": "results packed on the device (ebvo_batch_pack), then every rank copies its records device -> host over its own PCIe link into its "
                            "slice of a shared page-locked buffer (sharding.HostGather; counts all_gather + barrier are the only exchange)",
                            "nccl": "results packed on the device (ebvo_batch_pack), counts all_gather + NCCL send/recv to rank 0's GPU (sharding.gather_packed), then one D2H"}.get(strong[3], strong[3])),
                "gather": strong[3],
                "frames": FS, "n_gpus": world, "frames_per_gpu": -(-FS // world), "value": FS / (strong_ms / 1e3), "unit": "frames/s",
                "ms": strong_ms, "gather_ms_rank0": strong[1] * 1e3, "gather_share": strong[1] * 1e3 / strong_ms if strong_ms else None,
                "mate_records_gathered": strong[2], "bytes_gathered": strong[2] * 64,
                "timing": "host clock between barrier + cudaDeviceSynchronize on both sides, max over ranks, mean of 2 passes after one untimed pass"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Measurement of the quad-tracking row (BASELINE.json configs[3]): one keyframe -> current-frame pair of a
synthetic sequence at the ETH3D cables_2 shape (742x464) and at the KITTI shape, through ebvo_temporal_quads with HOST buffers
(copies inside the timed region), per-kernel durations from CUDA events, and the CPU path beside it (the reference's own
Temporal_Matches.cpp compiled in place, oracle/_ref/libtemporal_ref.so, all host threads; else the oracle port).
One JSON line per shape.  usage: python scripts/bench_temporal.py [--steps K] [--warmup W]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from edge_based_visual_odometry_b200 import synth, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()

for name, (W, H) in (("eth3d_cables_2-shape 742x464", (742, 464)), ("kitti-shape 1241x376", (1241, 376))):
    cal = synth.kitti_calib(W, H)     # the cables_2 YAML calibration yields no stereo mates in the reference itself (76 px vertical offset)
    calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
    ctx = _lib.Context(0, W, H, max_batch=1, max_edges=131072)
    frames = []
    for k in (0, 1):
        L, R, _ = synth.stereo_sequence_pair(cal, k)
        m = ctx.stereo_frame(calib, L, R)[0]
        frames.append((L, R, m))
    (L0, R0, m0), (L1, R1, m1) = frames
    kf_imgs, cf_imgs = (L0, L0, R0), (L1, L1, R1)
    for _ in range(args.warmup):
        off, q = ctx.temporal_quads(kf_imgs, cf_imgs, m0, m1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        off, q = ctx.temporal_quads(kf_imgs, cf_imgs, m0, m1)
    t_e2e = (time.perf_counter() - t0) / args.steps
    cnt = ctx.temporal_counters()
    ctx.set_profiling(True)
    ctx.temporal_quads(kf_imgs, cf_imgs, m0, m1)
    ctx.set_profiling(False); ctx.set_profiling(True)
    for _ in range(args.steps):
        ctx.temporal_quads(kf_imgs, cf_imgs, m0, m1)
    kt = {k: v[0] / args.steps for k, v in ctx.kernel_times().items() if k.startswith("tq_")}
    ctx.set_profiling(False)
    # roofline of the dominant kernel (tq_gn, 2-D Gauss-Newton on both views, Temporal_Matches.cpp:735-851): per iteration 98 samples x
    # 60 FP64 flop (coordinates, three four-corner blends rounded to float, residual, Huber weight, the five sums of the 2 x 2 normal
    # equations) = 5.9 kflop, against the MEASURED DFMA throughput (profiles/r02_fma_peaks.json)
    try:
        fp64_peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_fma_peaks.json")))["fp64_fma_tflops"]
    except Exception:
        fp64_peak = None
    gn_key = next((k for k in kt if k.startswith("tq_gn")), None)
    ach = 5880.0 * cnt["gn_iterations"] / (kt[gn_key] * 1e-3) / 1e12 if gn_key and kt[gn_key] > 0 else None
    roof = dict(kernel=gn_key, bound="fp64", algorithmic_flop_per_iteration=5880.0, iterations=cnt["gn_iterations"], achieved=ach, peak=fp64_peak, unit="TFLOP/s",
                frac=(ach / fp64_peak) if (ach and fp64_peak) else None, avg_launch_ms=kt.get(gn_key), share_of_kernel_time=(kt[gn_key] / sum(kt.values())) if gn_key else None)
    line = dict(roofline=roof, metric="quad-tracking frame pairs/s (keyframe -> current frame, grid + orientation + NCC + BNB + 2-D GN + clustering)",
                workload=name, value=1.0 / t_e2e, unit="pairs/s", ms_per_pair=1e3 * t_e2e, timing="host clock around ebvo_temporal_quads, host buffers in and out",
                kf_mates=int(len(m0)), cf_mates=int(len(m1)), quads=int(len(q)), counters=cnt,
                kernels_ms={k: round(v, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1])}, kernel_ms_total=round(sum(kt.values()), 4))
    if not args.no_cpu:
        a0 = np.stack([m0[k] for k in ("lx", "ly", "ltheta", "rx", "ry", "rtheta")], 1)
        a1 = np.stack([m1[k] for k in ("lx", "ly", "ltheta", "rx", "ry", "rtheta")], 1)
        t0 = time.perf_counter(); o = oracle.temporal(kf_imgs, cf_imgs, a0, a1); t_port = time.perf_counter() - t0
        line["cpu_baseline"] = dict(kind="port", value=1.0 / t_port, unit="pairs/s", cores=os.cpu_count(), sample="the same pair, oracle/temporal_oracle.inl (OpenMP)")
        if oracle.have_temporal_ref():
            t0 = time.perf_counter(); r = oracle.temporal_reference(kf_imgs, cf_imgs, a0, a1); t_ref = time.perf_counter() - t0
            line["cpu_baseline"] = dict(kind="reference", value=1.0 / t_ref, unit="pairs/s", cores=os.cpu_count(), port_pairs_per_s=1.0 / t_port,
                                        sample="the same pair, the reference's own Temporal_Matches.cpp compiled in place (oracle/_ref, shimmed OpenCV/Eigen), incl. its patch extraction")
            line["quads_reference"] = int(len(r.stages["cluster"]["cf"]))
        line["quads_oracle"] = int(len(o.stages["cluster"]["cf"]))
    print(json.dumps(line), flush=True)
    ctx.close()

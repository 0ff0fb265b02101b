"""Single-frame latency through the C ABI (the drop-in's per-frame call pattern): ebvo_stereo_frame with host buffers."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from edge_based_visual_odometry_b200 import synth, _lib
cal = synth.kitti_calib()
L, R = synth.stereo_pair(cal, 0)
ctx = _lib.Context(0, cal.width, cal.height, max_batch=1, max_edges=65536)
calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
for _ in range(5): m = ctx.stereo_frame(calib, L, R, want_edges=False)
t = []
for _ in range(30):
    t0 = time.perf_counter(); m = ctx.stereo_frame(calib, L, R, want_edges=False); t.append(time.perf_counter() - t0)
print(f"stereo_frame (no edge download): median {1e3*np.median(t):.2f} ms, min {1e3*min(t):.2f} ms, {len(m)} mates")
t = []
for _ in range(30):
    t0 = time.perf_counter(); m, a, b = ctx.stereo_frame(calib, L, R); t.append(time.perf_counter() - t0)
print(f"stereo_frame (+ both edge lists): median {1e3*np.median(t):.2f} ms")
ctx.set_profiling(True)
ctx.stereo_frame(calib, L, R, want_edges=False)
kt = ctx.kernel_times()
print("kernels ms:", {k: round(v[0], 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])}, "sum", round(sum(v[0] for v in kt.values()), 3))

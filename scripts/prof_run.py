"""Small fixed workload for ncu: B KITTI-shape frames resident on the device, a few full passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_based_visual_odometry_b200 import synth, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fp32 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sift = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cal = synth.kitti_calib()
pairs = [synth.stereo_pair(cal, f) for f in range(min(B, 4))]
Ls = [pairs[f % len(pairs)][0] for f in range(B)]
Rs = [pairs[f % len(pairs)][1] for f in range(B)]
prm = _lib.default_params(); prm.gn_mode = fp32; prm.sift_mode = sift
ctx = _lib.Context(0, cal.width, cal.height, max_batch=B, max_edges=65536, params=prm)
calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
ctx.batch_upload(Ls, Rs)
for _ in range(steps):
    ctx.batch_run(calib, True)
ctx.batch_sync()
print("ok", ctx.batch_counts()[2])

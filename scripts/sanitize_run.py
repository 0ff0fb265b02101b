"""Small end-to-end workload for compute-sanitizer (memcheck): every kernel family once, tiny inputs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_based_visual_odometry_b200 import synth, _lib
cal = synth.kitti_calib(320, 200)
pairs = [synth.stereo_pair(cal, f) for f in range(2)]
calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
for sift in (0, 1):
    prm = _lib.default_params(); prm.sift_mode = sift
    ctx = _lib.Context(0, 320, 200, max_batch=3, max_edges=16384, params=prm)
    m, Le, Re = ctx.stereo_frame(calib, *pairs[0])
    out, n = ctx.stereo_batch(calib, [p[0] for p in pairs] + [pairs[0][0]], [p[1] for p in pairs] + [pairs[0][1]], cap=8000)
    ctx.set_stage_dumps(True)
    m2 = ctx.stereo_match(calib, pairs[0][0], pairs[0][1], Le, Re)
    print("sift", sift, "mates", len(m), n.tolist(), len(m2))
    if sift:
        d = ctx.sift_descriptors(pairs[0][0], Le[:100])
    u = ctx.undistort(pairs[0][0], np.array([[458.654, 0, 160.0], [0, 457.296, 100.0], [0, 0, 1.0]]), np.array([-0.28, 0.07, 0.0002, 1.7e-5]))
    ctx.close()
for mode in (1, 2):
    prm = _lib.default_params(); prm.gn_mode = mode
    ctx = _lib.Context(0, 320, 200, max_batch=1, max_edges=16384, params=prm)
    print("gn_mode", mode, len(ctx.stereo_frame(calib, *pairs[1])[0]))
    ctx.close()
print("ok")

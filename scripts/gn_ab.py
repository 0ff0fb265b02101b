"""A/B of the Gauss-Newton kernels on the KITTI-shape pair: deviation from the oracle at the GN stage and on the final mates."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from edge_based_visual_odometry_b200 import synth, _lib
modes = [int(a) for a in sys.argv[1:]] or [0, 3, 1]
calK = synth.kitti_calib(); LK, RK = synth.stereo_pair(calK, 0)
eLK, _ = oracle.toed(LK); eRK, _ = oracle.toed(RK); FK, _ = oracle.fundamental(calK.Kl, calK.Kr, calK.R21, calK.T21)
resK = oracle.stereo(LK, RK, eLK, eRK, FK)
calibK = _lib.make_calib(calK.Kl, calK.Kr, calK.R21, calK.T21)
for mode in modes:
    prm = _lib.default_params(); prm.gn_mode = mode
    ctx = _lib.Context(0, 1241, 376, max_batch=1, max_edges=65536, params=prm)
    ctx.set_stage_dumps(True); ctx.set_profiling(True)
    m = ctx.stereo_match(calibK, LK, RK, _lib.edges_from_xyt(eLK), _lib.edges_from_xyt(eRK))
    kt = ctx.kernel_times()
    gn = ctx.stage("gn"); so = resK.stages["gn"]
    same = np.array_equal(gn["off"], so["off"])
    dgn = np.hypot(gn["x"] - so["x"], gn["y"] - so["y"]) if same else np.array([np.nan])
    dsc = np.abs(gn["score"] - so["score"]) if same else np.array([np.nan])
    common, io, ig = np.intersect1d(resK.mate_left, m["left_index"], return_indices=True)
    dm = np.hypot(resK.mate_right[io, 0] - m["rx"][ig], resK.mate_right[io, 1] - m["ry"][ig])
    dth = np.abs(resK.mate_right[io, 2] - m["rtheta"][ig])
    print(f"mode {mode}: gn offsets equal {same}; gn dpos max {np.nanmax(dgn):.2e} p99.9 {np.nanpercentile(dgn, 99.9):.2e} >1e-3 {(dgn > 1e-3).sum()}/{len(dgn)}"
          f" score max {np.nanmax(dsc):.2e} | mates {len(m)} vs {len(resK.mate_left)} common {len(common)} dpos max {dm.max():.2e} >1e-3 {(dm > 1e-3).sum()}"
          f" dth max {dth.max():.2e} >1e-4 {(dth > 1e-4).sum()} | gn ms {[round(v[0], 3) for k, v in kt.items() if k.startswith('gn')]}")
    ctx.close()

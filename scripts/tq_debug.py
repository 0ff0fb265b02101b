import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle
from edge_based_visual_odometry_b200 import synth, _lib
cal = synth.kitti_calib(480, 300)
F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
fr = []
for k in (0, 1):
    L, R, _ = synth.stereo_sequence_pair(cal, k, scene_seed=31)
    eL, _ = oracle.toed(L); eR, _ = oracle.toed(R)
    res = oracle.stereo(L, R, eL, eR, F21, want_dumps=False)
    fr.append((L, R, np.concatenate([eL[res.mate_left], res.mate_right], 1)))
(L0, R0, m0), (L1, R1, m1) = fr
o = oracle.temporal((L0, L0, R0), (L1, L1, R1), m0, m1)
ctx = _lib.Context(0, 480, 300, max_batch=1, max_edges=65536)
kf, cf = _lib.mates_from_arrays(m0[:, :3], m0[:, 3:]), _lib.mates_from_arrays(m1[:, :3], m1[:, 3:])
offg, qg = ctx.temporal_quads((L0, L0, R0), (L1, L1, R1), kf, cf, stage="gn")
offc, qc = ctx.temporal_quads((L0, L0, R0), (L1, L1, R1), kf, cf, stage="cluster")
so, sc = o.stages["gn"], o.stages["cluster"]
ng, no = np.diff(offc), np.diff(sc["off"])
bad = np.nonzero(ng != no)[0]
print("kf mates with different cluster counts", len(bad), "of", len(ng))
for i in bad[:3]:
    s, e = offg[i], offg[i + 1]
    print("KF", i, "gn list gpu cf", qg["cf_index"][s:e], "oracle cf", so["cf"][so["off"][i]:so["off"][i + 1]])
    print(" gpu lx", qg["lx"][s:e], "\n orc lx", so["left"][so["off"][i]:so["off"][i + 1], 0])
    print(" gpu ly", qg["ly"][s:e], "\n orc ly", so["left"][so["off"][i]:so["off"][i + 1], 1])
    print(" gpu th", qg["ltheta"][s:e])
    print(" clusters gpu", ng[i], qc[offc[i]:offc[i + 1]][["cf_index", "lx", "ly"]], "\n oracle", no[i], sc["cf"][sc["off"][i]:sc["off"][i + 1]], sc["left"][sc["off"][i]:sc["off"][i + 1], :2])

"""Where does the CUDA path leave the reference stereo golden (small pair)?  Per stage, per GN mode."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_based_visual_odometry_b200 import synth, _lib
g = np.load("tests/golden/stereo_small.npz"); ref = np.load("tests/golden/stereo_ref_small.npz")
cal = synth.kitti_calib(320, 200); calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
for mode in (0, 1):
    prm = _lib.default_params(); prm.gn_mode = mode
    ctx = _lib.Context(0, 1241, 376, max_batch=1, max_edges=65536, params=prm)
    ctx.set_stage_dumps(True)
    m = ctx.stereo_match(calib, g["L"], g["R"], _lib.edges_from_xyt(g["eL"]), _lib.edges_from_xyt(g["eR"]))
    for n in ("shift", "gn", "cluster", "ncc2", "best"):
        s = ctx.stage(n)
        same = np.array_equal(s["off"], ref[f"{n}_off"])
        print(f"mode {mode} {n}: offsets equal {same} total {s['off'][-1]} vs {ref[f'{n}_off'][-1]}")
        if not same:
            continue
        r = ref[f"{n}_xyt"]
        dp = np.hypot(s["x"] - r[:, 0], s["y"] - r[:, 1]); dt = np.abs(s["th"] - r[:, 2])
        print(f"   max dpos {dp.max():.3e} (>1e-3: {(dp > 1e-3).sum()})  max dth {dt.max():.3e} (>1e-4: {(dt > 1e-4).sum()}) of {len(dp)}")
        bad = np.nonzero((dt > 1e-4) | (dp > 1e-3))[0][:5]
        for k in bad:
            i = np.searchsorted(s["off"], k, side="right") - 1
            print(f"   entry {k} left edge {i}: gpu ({s['x'][k]:.6f},{s['y'][k]:.6f},{s['th'][k]:.6f}) ref ({r[k,0]:.6f},{r[k,1]:.6f},{r[k,2]:.6f})")
            for pn in ("shift", "gn"):
                ps = ctx.stage(pn); pr = ref[f"{pn}_xyt"]; a, b = ps["off"][i], ps["off"][i + 1]
                print(f"      {pn} gpu", np.stack([ps["x"][a:b], ps["y"][a:b], ps["th"][a:b]], 1).tolist())
                print(f"      {pn} ref", pr[ref[f'{pn}_off'][i]:ref[f'{pn}_off'][i + 1]].tolist())
    ctx.close()

"""A/B of a library build (run on a GPU box): parity of the matcher on the KITTI-shape pair against the oracle (stage lists,
Gauss-Newton stage, final mates) + per-kernel times of a resident batch.  One JSON line on stdout.
usage: python scripts/ab2.py [tag] [frames=16] [distinct=8] [reps=3] [parity=1]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_based_visual_odometry_b200 import synth, _lib

tag = sys.argv[1] if len(sys.argv) > 1 else "head"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
distinct = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
parity = int(sys.argv[5]) if len(sys.argv) > 5 else 1
out = {"tag": tag}
cal = synth.kitti_calib()
calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
if parity:
    import oracle
    L, R = synth.stereo_pair(cal, 0)
    eL, _ = oracle.toed(L); eR, _ = oracle.toed(R)
    F, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
    res = oracle.stereo(L, R, eL, eR, F)
    ctx = _lib.Context(0, cal.width, cal.height, max_batch=1, max_edges=65536)
    ctx.set_stage_dumps(True)
    m = ctx.stereo_match(calib, L, R, _lib.edges_from_xyt(eL), _lib.edges_from_xyt(eR))
    bad = []
    for name in _lib.STAGES:
        so, sg = res.stages[name], ctx.stage(name)
        if not np.array_equal(so["off"], sg["off"]):
            bad.append(name)
        elif name in ("epi", "disp", "orient", "ncc", "bnb_ncc") and not np.array_equal(so["ridx"], sg["ridx"]):
            bad.append(name + ":ridx")
    out["stages_differ"] = bad
    gn, so = ctx.stage("gn"), res.stages["gn"]
    if np.array_equal(gn["off"], so["off"]):
        d = np.hypot(gn["x"] - so["x"], gn["y"] - so["y"])
        out["gn_dpos_max"] = float(d.max()); out["gn_dpos_gt1e-6"] = int((d > 1e-6).sum()); out["gn_dpos_gt1e-3"] = int((d > 1e-3).sum())
        out["gn_score_max"] = float(np.nanmax(np.abs(gn["score"] - so["score"])))
    common, io, ig = np.intersect1d(res.mate_left, m["left_index"], return_indices=True)
    dm = np.hypot(res.mate_right[io, 0] - m["rx"][ig], res.mate_right[io, 1] - m["ry"][ig])
    dth = np.abs(res.mate_right[io, 2] - m["rtheta"][ig])
    out.update(mates=len(m), mates_oracle=len(res.mate_left), common=len(common), mate_dpos_max=float(dm.max()), mate_dpos_gt1e3=int((dm > 1e-3).sum()),
               mate_dth_max=float(dth.max()), mate_ncc_max=float(np.abs(res.mate_score[io] - m["score"][ig]).max()) if hasattr(res, "mate_score") else None)
    ctx.close()
pairs = [synth.stereo_pair(cal, f) for f in range(distinct)]
Ls = [pairs[f % distinct][0] for f in range(B)]; Rs = [pairs[f % distinct][1] for f in range(B)]
ctx = _lib.Context(0, cal.width, cal.height, max_batch=B, max_edges=65536)
ctx.batch_upload(Ls, Rs)
ctx.batch_run(calib, True); ctx.batch_sync()
ctx.set_profiling(True)
acc = {}
prev = {}
for _ in range(reps):
    ctx.batch_run(calib, True); ctx.batch_sync()
    for k, (ms, n) in ctx.kernel_times().items():      # cumulative since profiling was enabled
        acc.setdefault(k, []).append(ms - prev.get(k, 0.0)); prev[k] = ms
out["frames"] = B
out["kernel_ms"] = {k: round(float(np.median(v)), 4) for k, v in acc.items()}
out["total_ms"] = round(sum(out["kernel_ms"].values()), 3)
nL, nR, nM, cnt = ctx.batch_counts()
out["gn_iters"] = int(cnt[:, 3].sum()); out["mates_batch"] = int(nM.sum())
ctx.close()
print(json.dumps(out))

"""End-to-end parity (GPU TOED -> GPU matcher) against the oracle pipeline (FP64 TOED -> FP64 matcher), per stage.
Run on a GPU box: python scripts/e2e_parity.py [kitti euroc ...]   -> one JSON line per configuration."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from edge_based_visual_odometry_b200 import synth, _lib

INDEX_STAGES = ("epi", "disp", "orient", "ncc", "bnb_ncc")
GEOM_STAGES = ("shift", "gn", "cluster", "ncc2", "best")
for name in (sys.argv[1:] or ["kitti", "euroc"]):
    cal = synth.CALIBS[name]()
    out = {"config": name}
    for seed in (0, 1):
        L, R = synth.stereo_pair(cal, seed)
        eL, _ = oracle.toed(L); eR, _ = oracle.toed(R)
        F, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
        res = oracle.stereo(L, R, eL, eR, F)
        ctx = _lib.Context(0, cal.width, cal.height, max_batch=1, max_edges=65536)
        calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
        gL, _ = ctx.toed(L); gR, _ = ctx.toed(R)
        same_set = len(gL) == len(eL) and len(gR) == len(eR)
        r = {"edges": [len(eL), len(eR)], "same_edge_count": same_set}
        if same_set:
            r["toed_dpos_max"] = float(max(np.hypot(gL["x"] - eL[:, 0], gL["y"] - eL[:, 1]).max(), np.hypot(gR["x"] - eR[:, 0], gR["y"] - eR[:, 1]).max()))
            r["toed_dth_max"] = float(max(np.abs(np.angle(np.exp(1j * (gL["theta"] - eL[:, 2])))).max(), np.abs(np.angle(np.exp(1j * (gR["theta"] - eR[:, 2])))).max()))
            ctx.set_stage_dumps(True)
            m = ctx.stereo_match(calib, L, R, gL, gR)
            nL = len(eL)
            bad_any = np.zeros(nL, bool)
            per = {}
            for st in INDEX_STAGES + GEOM_STAGES:
                so, sg = res.stages[st], ctx.stage(st)
                co, cg = np.diff(so["off"]), np.diff(sg["off"])
                bad = co != cg
                # per-edge comparison of equal-length lists
                oo, og = so["off"], sg["off"]
                eq = np.where(~bad)[0]
                if st in INDEX_STAGES:
                    for i in eq:
                        if co[i] and not np.array_equal(so["ridx"][oo[i]:oo[i + 1]], sg["ridx"][og[i]:og[i + 1]]): bad[i] = True
                else:
                    for i in eq:
                        if co[i]:
                            a, b = slice(oo[i], oo[i + 1]), slice(og[i], og[i + 1])
                            if (np.hypot(so["x"][a] - sg["x"][b], so["y"][a] - sg["y"][b]) > 1e-3).any() or (np.abs(so["th"][a] - sg["th"][b]) > 1e-4).any(): bad[i] = True
                per[st] = float(bad.mean()); bad_any |= bad
            r["frac_left_edges_differ_per_stage"] = per
            r["frac_left_edges_differ_any_stage"] = float(bad_any.mean())
            common, io, ig = np.intersect1d(res.mate_left, m["left_index"], return_indices=True)
            d = np.hypot(res.mate_right[io, 0] - m["rx"][ig], res.mate_right[io, 1] - m["ry"][ig])
            dth = np.abs(res.mate_right[io, 2] - m["rtheta"][ig])
            r.update(mates_oracle=len(res.mate_left), mates_gpu=len(m), common=len(common), frac_mates_not_common=float(1 - len(common) / max(1, len(res.mate_left))),
                     frac_common_gt_1e3px=float((d > 1e-3).mean()), frac_common_gt_1e4rad=float((dth > 1e-4).mean()), dpos_median=float(np.median(d)), dpos_p99=float(np.percentile(d, 99)))
        ctx.close()
        out[f"seed{seed}"] = r
    print(json.dumps(out))

"""Developer check (run on a GPU box): CUDA path vs oracle, stage by stage, with per-kernel times."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from edge_based_visual_odometry_b200 import synth, _lib

shape = sys.argv[1] if len(sys.argv) > 1 else "kitti"
cal = synth.CALIBS[shape]()
L, R = synth.stereo_pair(cal, 0)
H, W = L.shape
ctx = _lib.Context(0, W, H, max_batch=4, max_edges=131072)
ctx.set_profiling(True)
calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)

# ---- TOED ----
t = time.time(); eo, nto = oracle.toed(L); print("oracle toed s", time.time() - t)
eg, ntg = ctx.toed(L)
print("TOED n oracle", len(eo), nto, "gpu", len(eg), ntg, ctx.kernel_times())
if len(eo) == len(eg):
    dx = np.abs(eg["x"] - eo[:, 0]).max(); dy = np.abs(eg["y"] - eo[:, 1]).max()
    dth = np.abs(np.angle(np.exp(1j * (eg["theta"] - eo[:, 2])))).max()
    print("TOED max |dx| %.3e |dy| %.3e |dth| %.3e" % (dx, dy, dth))
else:
    # set comparison on rounded interp cells
    so = set(map(tuple, np.round(eo[:, :2] * 2).astype(int)))
    sg = set(map(tuple, np.round(np.stack([eg["x"], eg["y"]], 1) * 2).astype(int)))
    print("TOED set diff: only oracle", len(so - sg), "only gpu", len(sg - so))

# ---- stereo, stage-isolated on oracle edges ----
eoR, _ = oracle.toed(R)
F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
Fg, _ = _lib.fundamental(calib)
print("F diff", np.abs(Fg - F21).max())
t = time.time(); res = oracle.stereo(L, R, eo, eoR, F21); print("oracle stereo s", time.time() - t)
ctx.set_stage_dumps(True)
mates = ctx.stereo_match(calib, L, R, _lib.edges_from_xyt(eo), _lib.edges_from_xyt(eoR))
print("kernel times (dump mode)", ctx.kernel_times())
print("mates oracle", len(res.mate_left), "gpu", len(mates))
for name in _lib.STAGES:
    so = res.stages[name]; sg = ctx.stage(name)
    same_off = np.array_equal(so["off"], sg["off"])
    msg = f"{name:9s} tot oracle {so['off'][-1]:8d} gpu {sg['off'][-1]:8d} offsets_equal {same_off}"
    if same_off:
        if name in ("epi", "disp", "orient", "sift", "ncc", "bnb_ncc", "bnb_sift"):
            msg += f" ridx_equal {np.array_equal(so['ridx'], sg['ridx'])}"
        if len(so["x"]):
            msg += " maxd xy %.2e th %.2e" % (max(np.abs(so["x"] - sg["x"]).max(), np.abs(so["y"] - sg["y"]).max()), np.abs(so["th"] - sg["th"]).max())
            if name in ("ncc", "bnb_ncc", "bnb_sift", "gn", "ncc2", "best"):
                msg += " score %.2e" % np.nanmax(np.abs(so["score"] - sg["score"]))
    else:
        co, cg = np.diff(so["off"]), np.diff(sg["off"])
        msg += f" edges_with_diff_count {(co != cg).sum()}"
    print(msg)
ml = set(res.mate_left.tolist()); mg = set(mates["left_index"].tolist())
print("mate left sets: common", len(ml & mg), "only oracle", len(ml - mg), "only gpu", len(mg - ml))
common = sorted(ml & mg)
io = {l: k for k, l in enumerate(res.mate_left.tolist())}; ig = {l: k for k, l in enumerate(mates["left_index"].tolist())}
ko = np.array([io[l] for l in common]); kg = np.array([ig[l] for l in common])
d = np.hypot(res.mate_right[ko, 0] - mates["rx"][kg], res.mate_right[ko, 1] - mates["ry"][kg])
print("mate right pos diff: max %.3e p99 %.3e  >1e-3: %d" % (d.max(), np.percentile(d, 99), (d > 1e-3).sum()))
ctx.set_stage_dumps(False)

# ---- production path timings: single frame and batch ----
for rep in range(2):
    m2, Le, Re = ctx.stereo_frame(calib, L, R)
    print("stereo_frame mates", len(m2), "nL", len(Le), "nR", len(Re), {k: round(v[0], 3) for k, v in ctx.kernel_times().items()})
B = 4
Ls, Rs = [], []
for f in range(B):
    a, b = synth.stereo_pair(cal, f); Ls.append(a); Rs.append(b)
ctx.batch_upload(Ls, Rs)
for rep in range(2):
    ctx.batch_run(calib, True); ctx.batch_sync()
    print("batch", B, {k: round(v[0], 3) for k, v in ctx.kernel_times().items()})
    ctx.set_profiling(True)
nL, nR, nM, cnt = ctx.batch_counts()
print("batch counts nL", nL, "nR", nR, "mates", nM, "\ncounters", cnt[:, :5])
ctx.set_profiling(False)
import ctypes
t = time.time()
for rep in range(5):
    ctx.batch_run(calib, True)
ctx.batch_sync()
print("batch of %d frames: %.3f ms per frame (wall, 5 reps)" % (B, (time.time() - t) / 5 / B * 1e3))

"""Wall-clock of the C++ drop-in members a maintainer links (run on a GPU box): ProcessEdges' detector call,
get_Stereo_Edge_Pairs + finalize_stereo_edge_mates on one KITTI-shape frame, SIFT-off and SIFT-on, and
get_Temporal_Edge_Pairs_from_Quads on one KITTI-shape keyframe -> current-frame pair.  JSON lines on stdout."""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from edge_based_visual_odometry_b200 import synth

B = os.path.join(ROOT, "dropin", "_build")
cal = synth.kitti_calib()
L, R = synth.stereo_pair(cal, 0)
eL, _ = oracle.toed(L); eR, _ = oracle.toed(R)
reps = sys.argv[1] if len(sys.argv) > 1 else "20"
with tempfile.TemporaryDirectory() as d:
    raw = os.path.join(d, "img.raw"); L.tofile(raw)
    out = subprocess.run([os.path.join(B, "test_dropin_toed"), raw, str(L.shape[0]), str(L.shape[1]), reps], capture_output=True, text=True, timeout=600)
    print(out.stdout.strip() or json.dumps({"failed": out.stderr[-300:]}))
    inp, outp = os.path.join(d, "in.bin"), os.path.join(d, "out.json")
    with open(inp, "wb") as f:
        np.array([L.shape[1], L.shape[0], len(eL), len(eR)], np.int32).tofile(f)
        np.concatenate([np.ravel(cal.Kl), np.ravel(cal.Kr), np.ravel(cal.R21), np.ravel(cal.T21)]).astype(np.float64).tofile(f)
        L.tofile(f); R.tofile(f)
        np.ascontiguousarray(eL[:, :3], np.float64).tofile(f); np.ascontiguousarray(eR[:, :3], np.float64).tofile(f)
    for sift in ("0", "1"):
        r = subprocess.run([os.path.join(B, "test_dropin_stereo"), inp, outp, "time", reps], capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, EBVO_DROPIN_SIFT=sift))
        if r.returncode == 0:
            j = json.load(open(outp)); j["sift"] = sift == "1"
            print(json.dumps(j))
        else:
            print(json.dumps({"failed": r.returncode, "sift": sift, "err": (r.stdout + r.stderr)[-300:]}))
    # quad tracking: mates of two consecutive frames of a synthetic sequence (device stereo), every keyframe mate takes part
    from edge_based_visual_odometry_b200 import _lib
    calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
    ctx = _lib.Context(0, cal.width, cal.height, max_batch=1, max_edges=131072)
    fr = []
    for k in (0, 1):
        Lk, Rk, _ = synth.stereo_sequence_pair(cal, k)
        fr.append((Lk, Rk, ctx.stereo_frame(calib, Lk, Rk)[0]))
    ctx.close()
    (L0, R0, m0), (L1, R1, m1) = fr
    six = lambda m: np.ascontiguousarray(np.stack([m[k] for k in ("lx", "ly", "ltheta", "rx", "ry", "rtheta")], 1), np.float64)
    tin, tout = os.path.join(d, "tq.bin"), os.path.join(d, "tq.json")
    with open(tin, "wb") as f:
        np.array([L0.shape[1], L0.shape[0], len(m0), len(m1)], np.int32).tofile(f)
        for im in (L0, R0, L1, R1):
            np.ascontiguousarray(im, np.uint8).tofile(f)
        six(m0).tofile(f); six(m1).tofile(f)
        np.ones(len(m0), np.uint8).tofile(f)
    r = subprocess.run([os.path.join(B, "test_dropin_temporal"), tin, tout, "time", reps], capture_output=True, text=True, timeout=900)
    if os.environ.get("EBVO_DROPIN_TRACE"):
        sys.stderr.write(r.stderr[-3000:])
    print(open(tout).read().strip() if r.returncode == 0 else json.dumps({"failed": r.returncode, "err": (r.stdout + r.stderr)[-300:]}))

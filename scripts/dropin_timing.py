"""Wall-clock of the C++ drop-in members a maintainer links (run on a GPU box): ProcessEdges' detector call and
get_Stereo_Edge_Pairs + finalize_stereo_edge_mates on one KITTI-shape frame, SIFT-off and SIFT-on.  JSON lines on stdout."""
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

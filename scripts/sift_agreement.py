"""GPU box: agreement of the device SIFT descriptors with cv2 on a KITTI-shape image (fraction of the 8-bit entries that differ)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cv2
import oracle
from edge_based_visual_odometry_b200 import synth, _lib
from test_gpu_sift import _cv2_descriptors
out = []
for seed in (0, 3):
    cal = synth.kitti_calib()
    img, _ = synth.stereo_pair(cal, seed)
    e, _ = oracle.toed(img)
    prm = _lib.default_params(); prm.sift_mode = 1
    ctx = _lib.Context(0, cal.width, cal.height, max_batch=1, max_edges=65536, params=prm)
    got = ctx.sift_descriptors(img, _lib.edges_from_xyt(e[:, :3]))
    ctx.close()
    want = _cv2_descriptors(img, e[:, :3])
    diff = np.abs(got - want)
    out.append({"seed": seed, "keypoints": int(2 * len(e)), "max_diff": float(diff.max()), "entries_differing": float((diff > 0).mean())})
print(json.dumps(out))

"""Generates the committed golden fixtures.  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

* toed_ref_small.npz   : a 200x152 synthetic image and the edges produced by the UNMODIFIED reference detector
                         (oracle/_ref/libtoed_ref.so = /root/reference/src/toed/cpu_toed.cpp compiled in place).
* stereo_small.npz     : a 320x200 KITTI-calibrated synthetic pair, its reference TOED edges (both views) and
                         the oracle's stage counts + final mates (regression pin of the restatement itself).
* stereo_ref_small.npz : the same pair through the REFERENCE'S OWN stereo code (oracle/_ref/libstereo_ref.so =
                         Stereo_Matches.cpp + utility.cpp + EdgeClusterer.cpp compiled in place against third_party_shim,
                         driven by oracle/ref_stereo_harness.cpp): per-stage candidate lists and final mates.
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle
from edge_based_visual_odometry_b200 import synth

assert oracle.have_ref(), "needs oracle/_ref/libtoed_ref.so (build container only)"

cal = synth.kitti_calib(200, 152)
L, _ = synth.stereo_pair(cal, 7, density=1.5)
e, nt, _, _ = oracle.toed_reference(L)
np.savez_compressed(os.path.join(HERE, "toed_ref_small.npz"), image=L, edges=e, n_total=nt)
print("toed_ref_small", L.shape, len(e), nt)

cal = synth.kitti_calib(320, 200)
L, R = synth.stereo_pair(cal, 11, density=1.5)
eL, ntL, _, _ = oracle.toed_reference(L)
eR, ntR, _, _ = oracle.toed_reference(R)
F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
res = oracle.stereo(L, R, eL, eR, F21)
counts = {k: int(v["off"][-1]) for k, v in res.stages.items()}
np.savez_compressed(os.path.join(HERE, "stereo_small.npz"), L=L, R=R, eL=eL, eR=eR, F21=F21,
                    stage_names=np.array(list(counts.keys())), stage_totals=np.array(list(counts.values())),
                    mate_left=res.mate_left, mate_right=res.mate_right, mate_score=res.mate_score)
print("stereo_small", len(eL), len(eR), counts, len(res.mate_left))

assert oracle.have_stereo_ref(), "needs oracle/_ref/libstereo_ref.so (build container only)"
ref = oracle.stereo_reference(L, R, eL, eR, cal.Kl, cal.Kr, cal.R21, cal.T21)
out = dict(mate_left=ref.mate_left, mate_right=ref.mate_right, mate_score=ref.mate_score, F21=ref.F21)
for k, v in ref.stages.items():
    out[f"{k}_off"] = v["off"]
    if k in ("epi", "disp", "orient", "ncc", "bnb_ncc"):      # index stages: positions are those of the right edges
        out[f"{k}_ridx"] = v["ridx"]
    if k in ("shift", "gn", "cluster", "ncc2", "best"):
        out[f"{k}_xyt"] = np.stack([v["x"], v["y"], v["th"]], 1).astype(np.float64)
    if k in ("ncc", "bnb_ncc", "ncc2", "best"):
        out[f"{k}_score"] = v["score"]
np.savez_compressed(os.path.join(HERE, "stereo_ref_small.npz"), **out)
print("stereo_ref_small", {k: int(v["off"][-1]) for k, v in ref.stages.items()}, len(ref.mate_left))

# temporal_ref_small.npz : two frames of a 320x200 synthetic SEQUENCE (synth.stereo_sequence_pair), their stereo mates
#                          (oracle) and the quad-tracking stages produced by the REFERENCE'S OWN Temporal_Matches.cpp
#                          (oracle/_ref/libtemporal_ref.so, driven by oracle/ref_temporal_harness.cpp).
assert oracle.have_temporal_ref(), "needs oracle/_ref/libtemporal_ref.so (build container only)"
cal = synth.kitti_calib(320, 200)
F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
fr = []
for k in (0, 1):
    L, R, _ = synth.stereo_sequence_pair(cal, k)
    eL, _ = oracle.toed(L)
    eR, _ = oracle.toed(R)
    res = oracle.stereo(L, R, eL, eR, F21, want_dumps=False)
    fr.append((L, R, np.concatenate([eL[res.mate_left], res.mate_right], 1)))
(L0, R0, m0), (L1, R1, m1) = fr
mask = (np.arange(len(m0)) % 3 != 1).astype(np.uint8)          # exercises the keyframe-mate selection
ref = oracle.temporal_reference((L0, L0, R0), (L1, L1, R1), m0, m1, mask)
out = dict(kfL=L0, kfR=R0, cfL=L1, cfR=R1, kf=m0, cf=m1, mask=mask)
for k, v in ref.stages.items():
    if k in ("sift", "bnb_sift"):      # pass-through stages of the SIFT-off run: identical to their predecessors
        continue
    out[f"{k}_off"] = v["off"]
    if k == "grid":      # 490 k entries: keep the per-keyframe-mate counts and an order-sensitive checksum of the lists
        out["grid_cfsum"] = np.add.reduceat(np.concatenate([v["cf"].astype(np.int64) * (1 + np.arange(len(v["cf"])) % 7), [0]]), v["off"][:-1].clip(max=len(v["cf"])))
        out["grid_cfsum"][np.diff(v["off"]) == 0] = 0
    else:
        out[f"{k}_cf"] = v["cf"]
    if k in ("ncc", "bnb", "gn", "cluster"):
        out[f"{k}_ncc"] = v["ncc"]
    if k in ("gn", "cluster"):
        out[f"{k}_left"], out[f"{k}_right"], out[f"{k}_score"], out[f"{k}_valid"] = v["left"], v["right"], v["score"], v["valid"]
# the same pair SIFT-on: descriptor pairs regenerated from the mate positions (synth.position_descriptors), so only the
# reference's lists are stored
desc = (synth.position_descriptors(m0[:, :3], 1), synth.position_descriptors(m0[:, 3:], 2),
        synth.position_descriptors(m1[:, :3], 1), synth.position_descriptors(m1[:, 3:], 2))
refs = oracle.temporal_reference((L0, L0, R0), (L1, L1, R1), m0, m1, mask, desc=desc)
for k in ("sift", "bnb_sift", "cluster"):
    v = refs.stages[k]
    out[f"son_{k}_off"], out[f"son_{k}_cf"], out[f"son_{k}_sift"] = v["off"], v["cf"], v["sift"]
out["son_cluster_left"], out["son_cluster_right"] = refs.stages["cluster"]["left"], refs.stages["cluster"]["right"]
np.savez_compressed(os.path.join(HERE, "temporal_ref_small.npz"), **out)
print("temporal_ref_small", {k: int(v["off"][-1]) for k, v in ref.stages.items()}, "SIFT-on", {k: int(v["off"][-1]) for k, v in refs.stages.items()})

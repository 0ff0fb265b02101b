"""GPU: keyframe -> current-frame quad tracking through the C ABI (ebvo_temporal_quads / _stage) against the oracle,
stage by stage, and against the committed output of the reference's own Temporal_Matches.cpp.

Tolerances: candidate lists identical at every stage (order included; after best-nearly-best the order is compared
per keyframe mate as a set, because the sort key - an NCC score - differs by < 4e-6 between two readings of OpenCV's
float type mix and near-ties swap); NCC within 1e-5; refined locations within 1e-3 px, orientation within 1e-4 rad
(the north-star tolerances of the stereo stage, BASELINE.json)."""
import os

import numpy as np
import pytest

import oracle
from edge_based_visual_odometry_b200 import synth, _lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "temporal_ref_small.npz")


def _canon(off, q):
    own = np.repeat(np.arange(len(off) - 1), np.diff(off))
    order = np.lexsort((q["cf_index"], own))
    return q[order]


def _as_quads(st):
    q = np.zeros(len(st["cf"]), _lib.QUAD_DTYPE)
    q["kf_index"] = np.repeat(np.arange(len(st["off"]) - 1), np.diff(st["off"]))
    q["cf_index"] = st["cf"]
    q["lx"], q["ly"], q["ltheta"] = st["left"].T
    q["rx"], q["ry"], q["rtheta"] = st["right"].T
    q["ncc_left"], q["ncc_right"] = st["ncc"].T
    q["sift_left"], q["sift_right"] = st["sift"].T if "sift" in st else (900.0, 900.0)
    q["score_left"], q["score_right"] = st["score"].T
    q["valid"] = st["valid"]
    return q


def _compare(name, off_g, qg, st, tol_px=1e-3, tol_rad=1e-4):
    qo = _as_quads(st)
    assert np.array_equal(off_g, st["off"]), name
    if name in ("bnb", "bnb_sift", "gn"):
        qg, qo = _canon(off_g, qg), _canon(st["off"], qo)
    assert np.array_equal(qg["kf_index"], qo["kf_index"]) and np.array_equal(qg["cf_index"], qo["cf_index"]), name
    if name in ("ncc", "sift", "bnb", "bnb_sift", "gn", "cluster"):
        assert np.abs(qg["sift_left"] - qo["sift_left"]).max() < 1e-9 and np.abs(qg["sift_right"] - qo["sift_right"]).max() < 1e-9, name
        assert np.abs(qg["ncc_left"] - qo["ncc_left"]).max() < 1e-5 and np.abs(qg["ncc_right"] - qo["ncc_right"]).max() < 1e-5, name
    for f in ("lx", "ly", "rx", "ry"):
        assert np.abs(qg[f] - qo[f]).max() < tol_px, (name, f)
    for f in ("ltheta", "rtheta"):
        assert np.abs(qg[f] - qo[f]).max() < tol_rad, (name, f)
    if name in ("gn", "cluster"):
        assert np.array_equal(qg["valid"], qo["valid"]), name
        assert np.abs(qg["score_left"] - qo["score_left"]).max() < 1e-5 and np.abs(qg["score_right"] - qo["score_right"]).max() < 1e-5, name   # RMS residual of the refinement (measured 1.7e-6 on non-converging sequences)


@pytest.fixture(scope="module")
def seq_case():
    cal = synth.kitti_calib(480, 300)
    F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
    fr = []
    for k in (0, 1):
        L, R, _ = synth.stereo_sequence_pair(cal, k, scene_seed=31)
        eL, _ = oracle.toed(L)
        eR, _ = oracle.toed(R)
        res = oracle.stereo(L, R, eL, eR, F21, want_dumps=False)
        fr.append((L, R, np.concatenate([eL[res.mate_left], res.mate_right], 1)))
    (L0, R0, m0), (L1, R1, m1) = fr
    o = oracle.temporal((L0, L0, R0), (L1, L1, R1), m0, m1)
    return dict(kf_imgs=(L0, L0, R0), cf_imgs=(L1, L1, R1), m0=m0, m1=m1, res=o)


def test_stage_by_stage_against_the_oracle(gpu_ctx, seq_case):
    c = seq_case
    kf, cf = _lib.mates_from_arrays(c["m0"][:, :3], c["m0"][:, 3:]), _lib.mates_from_arrays(c["m1"][:, :3], c["m1"][:, 3:])
    for name in _lib.TQ_STAGES:
        off, q = gpu_ctx.temporal_quads(c["kf_imgs"], c["cf_imgs"], kf, cf, stage=name)
        _compare(name, off, q, c["res"].stages[name])
    n = {k: len(v["cf"]) for k, v in c["res"].stages.items()}
    assert n["grid"] > n["orient"] > n["ncc"] == n["sift"] >= n["bnb"] == n["bnb_sift"] == n["gn"] > n["cluster"] > 1000
    cnt = gpu_ctx.temporal_counters()
    assert cnt["gate_survivors"] == n["bnb"] and cnt["gn_problems"] == 2 * n["gn"]


def test_sift_on_stage_by_stage_against_the_oracle(gpu_ctx, seq_case):
    """Descriptor pairs supplied: the SIFT gate and the SIFT best-nearly-best pass run on the device."""
    c = seq_case
    m0, m1 = c["m0"], c["m1"]
    desc = (synth.position_descriptors(m0[:, :3], 1), synth.position_descriptors(m0[:, 3:], 2),
            synth.position_descriptors(m1[:, :3], 1), synth.position_descriptors(m1[:, 3:], 2))
    res = oracle.temporal(c["kf_imgs"], c["cf_imgs"], m0, m1, desc=desc)
    kf, cf = _lib.mates_from_arrays(m0[:, :3], m0[:, 3:]), _lib.mates_from_arrays(m1[:, :3], m1[:, 3:])
    for name in ("sift", "bnb", "bnb_sift", "gn", "cluster"):
        off, q = gpu_ctx.temporal_quads(c["kf_imgs"], c["cf_imgs"], kf, cf, stage=name, desc=desc)
        _compare(name, off, q, res.stages[name])
    n = {k: len(v["cf"]) for k, v in res.stages.items()}
    assert n["ncc"] > n["sift"] == n["bnb"] > n["bnb_sift"] == n["gn"] > n["cluster"] > 300
    with pytest.raises(_lib.EbvoError):       # all four descriptor arrays or none
        gpu_ctx.temporal_quads(c["kf_imgs"], c["cf_imgs"], kf, cf, desc=(desc[0], None, None, None))


def test_gather_kernel_cross_checks_the_tiled_refinement(gpu_ctx, seq_case):
    """gn_mode 1 selects tq_gn_kernel (global-memory gathers, four-weight blend); the default tq_gn_tile_kernel (shared-memory
    tiles, interpolation form, cooperative 49th sample) implements the same FP64 arithmetic on another data path."""
    c = seq_case
    kf, cf = _lib.mates_from_arrays(c["m0"][:, :3], c["m0"][:, 3:]), _lib.mates_from_arrays(c["m1"][:, :3], c["m1"][:, 3:])
    prm = _lib.default_params(); prm.gn_mode = 1
    ctx = _lib.Context(0, 480, 300, max_batch=1, max_edges=65536, params=prm)
    off1, q1 = ctx.temporal_quads(c["kf_imgs"], c["cf_imgs"], kf, cf, stage="gn")
    ctx.close()
    off0, q0 = gpu_ctx.temporal_quads(c["kf_imgs"], c["cf_imgs"], kf, cf, stage="gn")
    assert np.array_equal(off0, off1) and np.array_equal(q0["cf_index"], q1["cf_index"]) and np.array_equal(q0["valid"], q1["valid"])
    for f in ("lx", "ly", "rx", "ry"):
        assert np.abs(q0[f] - q1[f]).max() < 1e-4, f
    assert np.abs(q0["score_left"] - q1["score_left"]).max() < 1e-5 and np.abs(q0["score_right"] - q1["score_right"]).max() < 1e-5


def test_default_call_and_mask(gpu_ctx, seq_case):
    c = seq_case
    kf, cf = _lib.mates_from_arrays(c["m0"][:, :3], c["m0"][:, 3:]), _lib.mates_from_arrays(c["m1"][:, :3], c["m1"][:, 3:])
    off, q = gpu_ctx.temporal_quads(c["kf_imgs"], c["cf_imgs"], kf, cf)
    _compare("cluster", off, q, c["res"].stages["cluster"])
    mask = (np.arange(len(kf)) % 2).astype(np.uint8)
    off2, q2 = gpu_ctx.temporal_quads(c["kf_imgs"], c["cf_imgs"], kf, cf, kf_mask=mask)
    assert (np.diff(off2)[mask == 0] == 0).all()
    keep = mask[q["kf_index"]] == 1
    assert np.array_equal(q2, q[keep])               # a keyframe mate's quads do not depend on the other mates


def test_golden_reference_output(gpu_ctx):
    """The reference's own Temporal_Matches.cpp output (tests/golden/make_golden.py) on a 320x200 sequence pair."""
    g = np.load(GOLD)
    kf, cf = _lib.mates_from_arrays(g["kf"][:, :3], g["kf"][:, 3:]), _lib.mates_from_arrays(g["cf"][:, :3], g["cf"][:, 3:])
    imgs_k, imgs_c = (g["kfL"], g["kfL"], g["kfR"]), (g["cfL"], g["cfL"], g["cfR"])
    for name in ("orient", "ncc", "gn", "cluster"):
        off, q = gpu_ctx.temporal_quads(imgs_k, imgs_c, kf, cf, kf_mask=g["mask"], stage=name)
        assert (q["sift_left"] == 900.0).all()
        st = dict(off=g[f"{name}_off"], cf=g[f"{name}_cf"])
        n = len(st["cf"])
        st["ncc"] = g[f"{name}_ncc"] if f"{name}_ncc" in g.files else np.full((n, 2), -1.0)
        if f"{name}_left" in g.files:
            st.update(left=g[f"{name}_left"], right=g[f"{name}_right"], score=g[f"{name}_score"], valid=g[f"{name}_valid"])
        else:
            st.update(left=g["cf"][st["cf"], :3], right=g["cf"][st["cf"], 3:], score=np.full((n, 2), 1e6), valid=np.zeros(n, np.int32))
        _compare(name, off, q, st)


def test_golden_reference_output_sift_on(gpu_ctx):
    """The reference's own SIFT gate / SIFT best-nearly-best / final clusters on the golden pair with descriptor pairs."""
    g = np.load(GOLD)
    kf, cf = _lib.mates_from_arrays(g["kf"][:, :3], g["kf"][:, 3:]), _lib.mates_from_arrays(g["cf"][:, :3], g["cf"][:, 3:])
    imgs_k, imgs_c = (g["kfL"], g["kfL"], g["kfR"]), (g["cfL"], g["cfL"], g["cfR"])
    desc = (synth.position_descriptors(g["kf"][:, :3], 1), synth.position_descriptors(g["kf"][:, 3:], 2),
            synth.position_descriptors(g["cf"][:, :3], 1), synth.position_descriptors(g["cf"][:, 3:], 2))
    for name in ("sift", "bnb_sift", "cluster"):
        off, q = gpu_ctx.temporal_quads(imgs_k, imgs_c, kf, cf, kf_mask=g["mask"], stage=name, desc=desc)
        roff, rcf, rs = g[f"son_{name}_off"], g[f"son_{name}_cf"], g[f"son_{name}_sift"]
        assert np.array_equal(off, roff), name
        if name == "bnb_sift":
            own = np.repeat(np.arange(len(roff) - 1), np.diff(roff))
            o1, o2 = np.lexsort((q["cf_index"], q["kf_index"])), np.lexsort((rcf, own))
            q, rcf, rs = q[o1], rcf[o2], rs[o2]
        assert np.array_equal(q["cf_index"], rcf), name
        assert np.abs(q["sift_left"] - rs[:, 0]).max() < 1e-9 and np.abs(q["sift_right"] - rs[:, 1]).max() < 1e-9, name
    assert np.abs(np.stack([q["lx"], q["ly"]], 1) - g["son_cluster_left"][:, :2]).max() < 1e-3
    assert np.abs(np.stack([q["rx"], q["ry"]], 1) - g["son_cluster_right"][:, :2]).max() < 1e-3


def test_full_size_pair_end_to_end(gpu_ctx):
    """BASELINE configs[3] at the ETH3D cables_2 shape (742x464): both frames through ebvo_stereo_frame, then the quad
    tracking, against the oracle run on the same mates (identical final quads, order included)."""
    cal = synth.kitti_calib(742, 464)   # the cables_2 YAML calibration yields no stereo mates in the reference itself
    calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
    fr = []
    for k in (0, 1):
        L, R, _ = synth.stereo_sequence_pair(cal, k)
        fr.append((L, R, gpu_ctx.stereo_frame(calib, L, R, want_edges=False)))
    (L0, R0, m0), (L1, R1, m1) = fr
    assert len(m0) > 10000 and len(m1) > 10000
    off, q = gpu_ctx.temporal_quads((L0, L0, R0), (L1, L1, R1), m0, m1)
    as6 = lambda m: np.stack([m[k] for k in ("lx", "ly", "ltheta", "rx", "ry", "rtheta")], 1)
    o = oracle.temporal((L0, L0, R0), (L1, L1, R1), as6(m0), as6(m1))
    _compare("cluster", off, q, o.stages["cluster"])
    cnt = gpu_ctx.temporal_counters()
    assert cnt["grid_candidates"] == len(o.stages["grid"]["cf"]) and cnt["orient_survivors"] == len(o.stages["orient"]["cf"])
    assert cnt["gate_survivors"] == len(o.stages["bnb"]["cf"]) and len(q) > 30000


def test_empty_inputs_and_errors(gpu_ctx):
    img = np.full((64, 96), 100, np.uint8)
    none = np.zeros(0, _lib.MATE_DTYPE)
    one = _lib.mates_from_arrays([[30.0, 30.0, 0.3]], [[25.0, 30.0, 0.3]])
    for kf, cf in ((none, none), (one, none), (none, one)):
        off, q = gpu_ctx.temporal_quads((img,) * 3, (img,) * 3, kf, cf)
        assert len(q) == 0 and off[-1] == 0
    off, q = gpu_ctx.temporal_quads((img,) * 3, (img,) * 3, one, one, stage="orient")
    assert len(q) == 1 and q["cf_index"][0] == 0 and q["score_left"][0] == 1e6
    off, q = gpu_ctx.temporal_quads((img,) * 3, (img,) * 3, one, one, stage="ncc")     # flat patches: similarity -1
    assert len(q) == 0
    with pytest.raises(_lib.EbvoError):
        gpu_ctx.temporal_quads((img,) * 3, (img,) * 3, one, one, stage="grid", cap=0)


def test_sift_on_with_device_descriptors_end_to_end():
    """The reference's default flow entirely through the C ABI: stereo frames with sift_mode = 1, the mates' descriptor
    pairs from ebvo_sift_descriptors (left edges on the left view, right edges on the right view: Stereo_Matches.cpp:655-689,
    1627-1635), then the quad tracking SIFT-on - against the oracle fed with cv2 descriptors at the same keypoints.
    Device descriptors equal cv2's on 99.97 % of the entries (off by one elsewhere), so quads whose SIFT distance lies
    within a few units of the 200 gate may flip: <= 0.5 % of the lists."""
    cv2 = pytest.importorskip("cv2")
    cal = synth.kitti_calib(480, 300)
    calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
    prm = _lib.default_params(); prm.sift_mode = 1
    ctx = _lib.Context(0, 480, 300, max_batch=1, max_edges=65536, params=prm)
    sift = cv2.SIFT_create()

    def cv_desc(img, xyt):
        kps = []
        for x, y, t in xyt:
            for s in (1, -1):
                kps.append(cv2.KeyPoint(float(x + s * 8 * np.sin(t)), float(y - s * 8 * np.cos(t)), 1, float(180 / np.pi * t)))
        return sift.compute(img, kps)[1].reshape(len(xyt), 2, 128).astype(np.float32)

    fr = []
    for k in (0, 1):
        L, R, _ = synth.stereo_sequence_pair(cal, k, scene_seed=31)
        m = ctx.stereo_frame(calib, L, R, want_edges=False)
        l3 = np.stack([m["lx"], m["ly"], m["ltheta"]], 1); r3 = np.stack([m["rx"], m["ry"], m["rtheta"]], 1)
        dev = (ctx.sift_descriptors(L, _lib.edges_from_xyt(l3)), ctx.sift_descriptors(R, _lib.edges_from_xyt(r3)))
        ref = (cv_desc(L, l3), cv_desc(R, r3))
        assert np.mean(dev[0] == ref[0]) > 0.999 and np.abs(dev[0] - ref[0]).max() <= 1
        fr.append((L, R, m, np.concatenate([l3, r3], 1), dev, ref))
    (L0, R0, m0, a0, d0, c0), (L1, R1, m1, a1, d1, c1) = fr
    off, q = ctx.temporal_quads((L0, L0, R0), (L1, L1, R1), m0, m1, desc=(d0[0], d0[1], d1[0], d1[1]))
    ctx.close()
    o = oracle.temporal((L0, L0, R0), (L1, L1, R1), a0, a1, desc=(c0[0], c0[1], c1[0], c1[1])).stages["cluster"]
    assert len(o["cf"]) > 300
    same = np.diff(off) == np.diff(o["off"])
    assert same.mean() >= 0.995, same.mean()
    # on the keyframe mates whose lists have the same length: same quads, same places
    og, oo = np.repeat(same, np.diff(off)), np.repeat(same, np.diff(o["off"]))
    assert np.mean(q["cf_index"][og] == o["cf"][oo]) >= 0.995
    ok = q["cf_index"][og] == o["cf"][oo]
    assert np.abs(q["lx"][og][ok] - o["left"][oo][ok, 0]).max() < 1e-3 and np.abs(q["ly"][og][ok] - o["left"][oo][ok, 1]).max() < 1e-3

"""CPU: the stereo restatement (oracle/stereo_oracle.cpp): OpenCV primitives vs cv2, known answers, golden pin."""
import numpy as np
import pytest

import oracle
from edge_based_visual_odometry_b200 import synth

cv2 = pytest.importorskip("cv2")


def test_fundamental_matrix_formula():
    for name in ("kitti", "euroc", "eth3d"):
        cal = synth.CALIBS[name]()
        F21, F12 = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
        F21n, F12n = synth.fundamental_matrices(cal)
        assert np.abs(F21 - F21n).max() < 1e-15 and np.abs(F12 - F12n).max() < 1e-15
        assert np.abs(F21 - F12.T).max() < 1e-6 * np.abs(F21).max()   # F12 = F21^T up to the non-orthonormality of the YAML R21


def test_sobel_equals_cv2():
    """util_compute_Img_Gradients (utility.h:131-141): cv::Sobel CV_32F ksize 3 scale 1/8, default border."""
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (67, 91), dtype=np.uint8)
    gx, gy = oracle.sobel(img)
    f = img.astype(np.float32)
    assert np.array_equal(gx, cv2.Sobel(f, cv2.CV_32F, 1, 0, ksize=3, scale=1.0 / 8.0))
    assert np.array_equal(gy, cv2.Sobel(f, cv2.CV_32F, 0, 1, ksize=3, scale=1.0 / 8.0))


def _cv_patch_similarity(p1, p2):
    """get_patch_similarity (utility.cpp:163-180) spelled with the cv2 primitives OpenCV's MatExpr lowers to."""
    m1, m2 = cv2.mean(p1)[0], cv2.mean(p2)[0]
    d1, d2 = cv2.subtract(p1, (m1, 0, 0, 0)), cv2.subtract(p2, (m2, 0, 0, 0))
    s1, s2 = cv2.sumElems(cv2.multiply(d1, d1))[0], cv2.sumElems(cv2.multiply(d2, d2))[0]
    if s1 < 1e-10 or s2 < 1e-10:
        return -1.0
    n1 = cv2.addWeighted(p1, 1.0 / np.sqrt(s1), p1, 0.0, -m1 / np.sqrt(s1))
    n2 = cv2.addWeighted(p2, 1.0 / np.sqrt(s2), p2, 0.0, -m2 / np.sqrt(s2))
    return float(n1.ravel().astype(np.float64) @ n2.ravel().astype(np.float64))


def test_patch_similarity_vs_cv2_and_known_answers():
    rng = np.random.default_rng(1)
    for _ in range(50):
        a = (rng.random((7, 7)) * 255).astype(np.float32)
        b = (a + rng.normal(0, 20, (7, 7))).astype(np.float32)
        assert abs(oracle.patch_similarity(a, b) - _cv_patch_similarity(a, b)) < 2e-6
    a = (rng.random((7, 7)) * 255).astype(np.float32)
    assert abs(oracle.patch_similarity(a, a) - 1.0) < 1e-6
    assert abs(oracle.patch_similarity(a, 255 - a) + 1.0) < 1e-6
    assert oracle.patch_similarity(a, np.full((7, 7), 9.0, np.float32)) == -1.0      # flat patch sentinel
    n = a.copy(); n[3, 3] = np.nan
    assert np.isnan(oracle.patch_similarity(n, a))


def test_patch_sampling_nan_on_integer_coordinates_and_outside():
    img = np.arange(60 * 80, dtype=np.uint32).reshape(60, 80).astype(np.uint8)
    # theta = 0: "+" patch centre (x, y-5); cells at integer offsets => integer coordinates => NaN (utility.h:95-103)
    p, m = oracle.edge_patches(img, 30.0, 30.0, 0.0)
    assert np.isnan(p).all() and np.isnan(m).all()
    p, m = oracle.edge_patches(img, 30.25, 30.5, 0.3)
    assert np.isfinite(p).all() and np.isfinite(m).all()
    p, m = oracle.edge_patches(img, 2.25, 30.5, 0.3)       # leaves the image on the left
    assert np.isnan(p).any() or np.isnan(m).any()


def test_patch_layout_matches_reference_convention():
    """patch[i+3][j+3] = I(c + (cos*i - sin*j, sin*i + cos*j)), '+' side c = p + 5*(sin, -cos) (utility.cpp:82-93,146-159)."""
    yy, xx = np.mgrid[0:100, 0:120]
    img = np.clip(xx + 0 * yy, 0, 255).astype(np.uint8)   # I = x
    th = 0.4
    x, y = 50.3, 40.7
    p, m = oracle.edge_patches(img, x, y, th)
    c = np.array([x + 5 * np.sin(th), y - 5 * np.cos(th)])
    for i in range(-3, 4):
        for j in range(-3, 4):
            assert abs(p[i + 3, j + 3] - (c[0] + np.cos(th) * i - np.sin(th) * j)) < 1e-4


def test_cluster_hand_built_sets():
    # two well separated pairs -> two clusters, centres = plain means (equal distances => equal weights)
    pts = np.array([[10.0, 10.0, 0.5], [10.4, 10.0, 0.52], [20.0, 10.0, 0.5], [20.0, 10.6, 0.48]])
    cen, lab = oracle.cluster(pts, True)
    assert len(cen) == 2 and lab.tolist() == [0, 0, 1, 1]
    assert np.allclose(cen[0], [10.2, 10.0, 0.51]) and np.allclose(cen[1], [20.0, 10.3, 0.49])
    # orientation gate (20 deg, raw difference) keeps close points apart
    pts = np.array([[10.0, 10.0, 0.0], [10.3, 10.0, 1.0]])
    cen, lab = oracle.cluster(pts, True)
    assert len(cen) == 2
    cen, lab = oracle.cluster(pts, False)
    assert len(cen) == 1
    # MAX_CLUSTER_SIZE = 10: twelve coincident-ish points cannot all merge
    pts = np.array([[5 + 0.01 * k, 5.0, 0.1] for k in range(12)])
    cen, lab = oracle.cluster(pts, True)
    assert len(cen) == 2 and max(np.bincount(lab)) <= 10


def test_shift_to_epipolar_line_cases():
    line = np.array([0.0, 1.0, -50.0])     # y = 50
    # (1) closer than 0.4 px: perpendicular foot
    assert np.allclose(oracle.shift_to_line(line, 30.0, 50.3, 1.0), [30.0, 50.0, 1.0])
    # (2) farther: slide along the tangent if displacement < 3
    out = oracle.shift_to_line(line, 30.0, 51.0, np.pi / 2 - 0.2)
    assert abs(out[1] - 50.0) < 1e-9 and abs(out[2] - (np.pi / 2 - 0.2)) < 1e-15
    # (3) nearly parallel tangent: orientation perturbed by 10 deg, still too far => original returned
    out = oracle.shift_to_line(line, 30.0, 52.5, 0.01)
    assert np.allclose(out, [30.0, 52.5, 0.01])


def test_full_stereo_golden_pin(golden_stereo):
    g = golden_stereo
    res = oracle.stereo(g["L"], g["R"], g["eL"], g["eR"], g["F21"])
    for name, tot in zip(g["stage_names"], g["stage_totals"]):
        assert res.stages[str(name)]["off"][-1] == tot
    assert np.array_equal(res.mate_left, g["mate_left"])
    assert np.abs(res.mate_right - g["mate_right"]).max() < 1e-9
    assert np.abs(res.mate_score - g["mate_score"]).max() < 1e-12


def test_full_stereo_invariants(golden_stereo):
    g = golden_stereo
    res = oracle.stereo(g["L"], g["R"], g["eL"], g["eR"], g["F21"])
    st = res.stages
    # each gate only removes candidates, in order
    for a, b in (("epi", "disp"), ("disp", "orient"), ("orient", "ncc"), ("ncc", "bnb_ncc")):
        assert (np.diff(st[b]["off"]) <= np.diff(st[a]["off"])).all()
    for k in ("epi", "disp", "orient"):            # ascending right-edge index inside every list
        off, r = st[k]["off"], st[k]["ridx"]
        d = np.diff(r)
        starts = off[1:-1][np.diff(off)[:-1] > 0]
        mask = np.ones(len(d), bool); mask[starts[starts < len(r)] - 1] = False
        assert (d[mask] > 0).all()
    assert (st["ncc"]["score"] > 0.6).all() and (res.mate_score > 0.6).all()
    # final mates sit on the epipolar line of their left edge (second shift projects them)
    l = res.lines[res.mate_left]
    d = np.abs(l[:, 0] * res.mate_right[:, 0] + l[:, 1] * res.mate_right[:, 1] + l[:, 2]) / np.hypot(l[:, 0], l[:, 1])
    assert np.percentile(d, 99) < 1e-6
    disp = g["eL"][res.mate_left, 0] - res.mate_right[:, 0]
    assert np.percentile(np.abs(disp), 99) < 30      # GN may slide a few candidates far along the line (reference behaviour)


# ---- pin against the reference's own stereo code (Stereo_Matches.cpp + utility.cpp + EdgeClusterer.cpp compiled in
# ---- place against third_party_shim; golden fixture generated by tests/golden/make_golden.py) --------------------------
import os

GOLDEN_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stereo_ref_small.npz")
INDEX_STAGES = ("epi", "disp", "orient", "ncc", "bnb_ncc")


def _same_sets_per_left_edge(off_a, idx_a, off_b, idx_b):
    """Candidate index lists equal per left edge up to order (order may differ only between NCC near-ties)."""
    assert np.array_equal(off_a, off_b)
    bad = 0
    for i in range(len(off_a) - 1):
        a, b = idx_a[off_a[i]:off_a[i + 1]], idx_b[off_b[i]:off_b[i + 1]]
        if not np.array_equal(a, b) and not np.array_equal(np.sort(a), np.sort(b)):
            bad += 1
    return bad


def test_restatement_matches_reference_stereo_golden(golden_stereo):
    """The restatement against per-stage output of the reference's own compiled stereo code on the same pair."""
    g = golden_stereo
    ref = np.load(GOLDEN_REF)
    res = oracle.stereo(g["L"], g["R"], g["eL"], g["eR"], g["F21"])
    assert np.abs(ref["F21"] - g["F21"]).max() == 0.0
    for name in oracle.STAGES:
        assert np.array_equal(res.stages[name]["off"], ref[f"{name}_off"]), name
    for name in ("epi", "disp", "orient", "ncc"):                     # identical candidate lists, identical order
        assert np.array_equal(res.stages[name]["ridx"], ref[f"{name}_ridx"]), name
    assert _same_sets_per_left_edge(res.stages["bnb_ncc"]["off"], res.stages["bnb_ncc"]["ridx"], ref["bnb_ncc_off"], ref["bnb_ncc_ridx"]) == 0
    # NCC values: OpenCV's float type mix is restated twice (oracle: (p - m) * inv ; shim: p * inv - m * inv as
    # MatExpr lowers it); both are within a few 1e-6 of each other, hence the 1e-5 tie tolerance of the north star
    assert np.abs(res.stages["ncc"]["score"] - ref["ncc_score"]).max() < 1e-5
    for name in ("cluster", "ncc2", "best"):                          # order-insensitive from here on
        xyt = np.stack([res.stages[name]["x"], res.stages[name]["y"], res.stages[name]["th"]], 1)
        assert np.abs(xyt - ref[f"{name}_xyt"]).max() < 1e-9, name
    assert np.array_equal(res.mate_left, ref["mate_left"])
    assert np.abs(res.mate_right - ref["mate_right"]).max() < 1e-9
    assert np.abs(res.mate_score - ref["mate_score"]).max() < 1e-5


@pytest.mark.skipif(not oracle.have_stereo_ref(), reason="oracle/_ref/libstereo_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("name,shape,seed", [("kitti", (400, 240), 3), ("euroc", (376, 240), 5)])
def test_restatement_matches_compiled_reference_stereo(name, shape, seed):
    """Live run of the reference sources (rectified and general-F calibration) against the restatement."""
    cal = synth.CALIBS[name](*shape)
    L, R = synth.stereo_pair(cal, seed, density=1.5)
    eL, _ = oracle.toed(L)
    eR, _ = oracle.toed(R)
    F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
    res = oracle.stereo(L, R, eL, eR, F21)
    ref = oracle.stereo_reference(L, R, eL, eR, cal.Kl, cal.Kr, cal.R21, cal.T21)
    assert np.abs(ref.F21 - F21).max() < 1e-18 + 1e-12 * np.abs(F21).max()
    assert np.abs(ref.lines - res.lines).max() < 1e-12 * np.abs(res.lines).max()
    diff_edges = np.zeros(len(eL), bool)
    for st in oracle.STAGES:
        diff_edges |= np.diff(res.stages[st]["off"]) != np.diff(ref.stages[st]["off"])
    assert diff_edges.mean() <= 1e-3                                    # NCC near-ties at the 0.6 / 0.9 gates only
    if not diff_edges.any():
        for st in ("epi", "disp", "orient", "ncc"):
            assert np.array_equal(res.stages[st]["ridx"], ref.stages[st]["ridx"])
    common, io, ir = np.intersect1d(res.mate_left, ref.mate_left, return_indices=True)
    assert len(common) >= (1 - 1e-3) * len(ref.mate_left) > 100
    d = np.abs(res.mate_right[io] - ref.mate_right[ir]).max(axis=1)
    assert (d > 1e-9).mean() <= 1e-3


@pytest.mark.skipif(not oracle.have_stereo_ref(), reason="oracle/_ref/libstereo_ref.so not built (needs /root/reference)")
def test_reference_primitives_vs_restatement():
    """Utility::get_edge_patches / get_patch_similarity and EdgeClusterer, called directly in the reference library."""
    rng = np.random.default_rng(7)
    cal = synth.kitti_calib(200, 152)
    img, _ = synth.stereo_pair(cal, 7)
    for _ in range(30):
        x, y, th = rng.uniform(20, 180), rng.uniform(20, 130), rng.uniform(-np.pi, np.pi)
        p, m = oracle.edge_patches(img, x, y, th)
        rp, rm = oracle.ref_edge_patches(img, x, y, th)
        assert np.array_equal(p, rp, equal_nan=True) and np.array_equal(m, rm, equal_nan=True)
        q, _ = oracle.edge_patches(img, x + 3.3, y + 0.4, th + 0.05)
        assert abs(oracle.patch_similarity(p, q) - oracle.ref_patch_similarity(p, q)) < 1e-5
    assert np.isnan(oracle.ref_edge_patches(img, 30.0, 30.0, 0.0)[0]).all()
    for trial in range(30):
        n = int(rng.integers(1, 25))
        pts = np.stack([100 + np.cumsum(rng.choice([0.2, 0.6, 1.4], n)), 50 + rng.normal(0, 0.2, n), rng.choice([0.3, 0.35, 0.9], n)], 1)
        for by in (True, False):
            cen, lab = oracle.cluster(pts, by)
            rcen, rcnt = oracle.ref_cluster(pts, by)
            assert len(cen) == len(rcen) and np.abs(cen - rcen).max() < 1e-12
            assert np.array_equal(np.bincount(lab, minlength=len(cen)), rcnt)

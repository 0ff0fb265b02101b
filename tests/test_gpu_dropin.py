"""GPU: the C++ drop-in translation unit (dropin/cpu_toed_b200.cpp) compiled against the reference's own header,
driven exactly like Pipeline::ProcessEdges, must give the reference's edges."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from edge_based_visual_odometry_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dropin", "_build", "test_dropin_toed")


@pytest.mark.skipif(not os.path.exists(EXE), reason="dropin/_build not built (needs the reference headers at build time)")
def test_reference_class_backed_by_the_gpu(tmp_path):
    cal = synth.kitti_calib(640, 240)
    img, _ = synth.stereo_pair(cal, 4)
    raw = tmp_path / "img.raw"
    img.tofile(raw)
    out = subprocess.run([EXE, str(raw), str(img.shape[0]), str(img.shape[1])], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    n, nt = map(int, lines[0].split())
    got = np.array([[float(v) for v in l.split()] for l in lines[1:]])
    eo, nto = oracle.toed(img)
    assert n == len(eo) and nt == nto
    assert np.abs(got[:, :2] - eo[:, :2]).max() < 1e-3 and np.abs(got[:, 2] - eo[:, 2]).max() < 1e-4
    assert np.array_equal(got[:, 3].astype(int), np.arange(n))


EXE_STEREO = os.path.join(ROOT, "dropin", "_build", "test_dropin_stereo")
EXE_UNITS = os.path.join(ROOT, "dropin", "_build", "test_dropin_units")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _run_stereo_dropin(tmp_path, cal, L, R, eL, eR, writer_dir=None, sift=False):
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        np.array([L.shape[1], L.shape[0], len(eL), len(eR)], np.int32).tofile(f)
        np.concatenate([np.ravel(cal.Kl), np.ravel(cal.Kr), np.ravel(cal.R21), np.ravel(cal.T21)]).astype(np.float64).tofile(f)
        np.ascontiguousarray(L).tofile(f); np.ascontiguousarray(R).tofile(f)
        np.ascontiguousarray(eL[:, :3], np.float64).tofile(f); np.ascontiguousarray(eR[:, :3], np.float64).tofile(f)
    env = dict(os.environ, EBVO_DROPIN_SIFT="1" if sift else "0")
    out = subprocess.run([EXE_STEREO, str(inp), str(outp)] + ([str(writer_dir)] if writer_dir else []), capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, (out.returncode, out.stdout[-2000:], out.stderr[-2000:])
    raw = np.fromfile(outp, np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    return raw[4:].view(np.float64).reshape(n, 16)


@pytest.mark.skipif(not os.path.exists(EXE_STEREO), reason="dropin/_build not built (needs the reference headers at build time)")
def test_stereo_matches_members_backed_by_the_gpu_against_reference_output(tmp_path):
    """Pipeline::get_Stereo_Edge_Correspondences' call sequence on the reference's own classes, with
    get_Stereo_Edge_Pairs / finalize_stereo_edge_mates replaced by dropin/stereo_matches_b200.cpp, against the
    output of the reference's own CPU code on the same pair (tests/golden/stereo_ref_small.npz)."""
    g = np.load(os.path.join(GOLDEN, "stereo_small.npz"))
    ref = np.load(os.path.join(GOLDEN, "stereo_ref_small.npz"))
    cal = synth.kitti_calib(320, 200)
    rows = _run_stereo_dropin(tmp_path, cal, g["L"], g["R"], g["eL"], g["eR"], writer_dir=tmp_path)
    # on-disk format: the reference's own writer ran on the GPU mates - header + 16 columns per mate, the
    # first six being the left / right edges (default ostream precision: 6 significant digits)
    txt = (tmp_path / "finalized_stereo_edge_pairs_frame_0.txt").read_text().splitlines()
    assert txt[0].startswith("left_edge_location, left_edge_orientation, right_edge_location") and len(txt) == len(rows) + 1
    tab = np.array([[float(v) for v in l.split()] for l in txt[1:]])
    assert tab.shape == (len(rows), 16) and np.isfinite(tab[:, :6]).all()
    assert np.abs(tab[:, :6] - rows[:, 1:7]).max() < 5e-4 * max(1.0, np.abs(rows[:, 1:7]).max())
    assert np.array_equal(rows[:, 0].astype(int), ref["mate_left"])                       # focused_edge_indices after the stage
    assert np.array_equal(rows[:, 1:4], g["eL"][ref["mate_left"], :3])                     # left_edge copied from the frame
    assert np.abs(rows[:, 4:6] - ref["mate_right"][:, :2]).max() < 1e-3                    # right_edge location
    assert np.abs(rows[:, 6] - ref["mate_right"][:, 2]).max() < 1e-4                       # right_edge orientation
    assert np.abs(rows[:, 7] - ref["mate_score"]).max() < 1e-5                             # refine_final_scores[0]
    assert (rows[:, 12] == 0).all()                                                        # b_is_TP against the (-1,-1) no-GT placeholder
    # patches handed to finalisation: left from the raw left image, right from the undistorted right image (oracle = CPU restatement)
    for k in (0, len(rows) // 2, len(rows) - 1):
        pl, ml = oracle.edge_patches(g["L"], *rows[k, 1:4])
        pr, mr = oracle.edge_patches(g["R"], *rows[k, 4:7])
        for got, want in ((rows[k, 8], pl), (rows[k, 9], ml), (rows[k, 10], pr), (rows[k, 11], mr)):
            w = float(np.sum(want.astype(np.float64)))
            assert (np.isnan(got) and np.isnan(w)) or abs(got - w) < 1e-6 * max(1.0, abs(w))


@pytest.mark.skipif(not os.path.exists(EXE_STEREO), reason="dropin/_build not built (needs the reference headers at build time)")
def test_stereo_matches_dropin_general_calibration(tmp_path):
    """The same flow with the EuRoC calibration (R21 != I): Dataset's F21 and the library's agree through the drop-in."""
    cal = synth.CALIBS["euroc"]()
    L, R = synth.stereo_pair(cal, 1)
    eL, _ = oracle.toed(L)
    eR, _ = oracle.toed(R)
    F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
    res = oracle.stereo(L, R, eL, eR, F21, want_dumps=False)
    rows = _run_stereo_dropin(tmp_path, cal, L, R, eL, eR)
    assert np.array_equal(rows[:, 0].astype(int), res.mate_left) and len(rows) > 1000
    assert np.abs(rows[:, 4:6] - res.mate_right[:, :2]).max() < 1e-3 and np.abs(rows[:, 6] - res.mate_right[:, 2]).max() < 1e-4
    line_c = (F21 @ np.c_[eL[res.mate_left, :2], np.ones(len(rows))].T)[2]
    assert np.abs(rows[:, 13] - line_c).max() < 1e-9 * max(1.0, np.abs(line_c).max())     # epip_line_coeffs_of_left_edges


@pytest.mark.skipif(not os.path.exists(EXE_UNITS), reason="dropin/_build not built (needs the reference headers at build time)")
def test_utility_clusterer_and_matlab_ncc_adapters(tmp_path):
    """Utility::get_edge_patches / get_patch_similarity, MatlabNCCComputer::computeNCC and EdgeClusterer backed by the GPU."""
    cal = synth.kitti_calib(320, 200)
    img, _ = synth.stereo_pair(cal, 5)
    e, _ = oracle.toed(img)
    e = e[:: max(1, len(e) // 40)][:40, :3].copy()
    e[3] = (np.round(e[3, 0]), np.round(e[3, 1]), 0.0)  # integer sample coordinates: the reference's bilinear gives NaN (utility.h:95-103)
    rng = np.random.default_rng(3)
    sets = []
    for s in range(12):
        n = int(rng.integers(1, 16))
        base = rng.uniform(50, 150, 2)
        pts = np.c_[base[0] + rng.normal(0, 0.7, n), np.full(n, base[1]), rng.uniform(-0.4, 0.4, n)]
        sets.append(pts)
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        np.array([img.shape[1], img.shape[0], len(e), len(sets)], np.int32).tofile(f)
        np.ascontiguousarray(img).tofile(f); np.ascontiguousarray(e, np.float64).tofile(f)
        for p in sets:
            np.array([len(p)], np.int32).tofile(f); np.ascontiguousarray(p, np.float64).tofile(f)
    out = subprocess.run([EXE_UNITS, str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout[-2000:], out.stderr[-2000:])
    raw = np.fromfile(outp, np.uint8)
    nE = len(e)
    patches = raw[:nE * 98 * 4].view(np.float32).reshape(nE, 2, 49); o = nE * 98 * 4
    ncc = raw[o:o + (nE - 1) * 8].view(np.float64); o += (nE - 1) * 8
    mncc = raw[o:o + (nE - 1) * 8].view(np.float64); o += (nE - 1) * 8
    for k in range(nE):
        p, m = oracle.edge_patches(img, *e[k])
        assert np.array_equal(patches[k, 0], np.ravel(p), equal_nan=True) and np.array_equal(patches[k, 1], np.ravel(m), equal_nan=True)
    assert np.isnan(patches[3]).any()
    for k in range(nE - 1):
        for got, a, b in ((ncc[k], patches[k, 0], patches[k + 1, 0]), (mncc[k], patches[k, 1], patches[k + 1, 0])):
            want = oracle.patch_similarity(a.reshape(7, 7), b.reshape(7, 7))
            assert (np.isnan(got) and np.isnan(want)) or abs(got - want) < 1e-5
    for p in sets:
        ncl = int(raw[o:o + 4].view(np.int32)[0]); o += 4
        cen = raw[o:o + ncl * 32].view(np.float64).reshape(ncl, 4); o += ncl * 32
        lab = raw[o:o + 4 * len(p)].view(np.int32); o += 4 * len(p)
        want_c, want_lab = oracle.cluster(p, True)
        assert ncl == len(want_c) and np.array_equal(lab, want_lab)
        assert np.abs(cen[:, :3] - want_c).max() < 1e-9
        assert np.array_equal(cen[:, 3].astype(int), np.bincount(want_lab, minlength=ncl))       # contributing_edges per cluster


@pytest.mark.skipif(not os.path.exists(EXE_STEREO), reason="dropin/_build not built (needs the reference headers at build time)")
def test_stereo_matches_dropin_sift_on(tmp_path):
    """The reference's default flow (SIFT gate, BNB-SIFT, descriptors of the finalised mates) through the drop-in, against
    the oracle fed with cv2 descriptors (cv::SIFT is OpenCV code; tests/test_gpu_sift.py pins the descriptor kernel)."""
    cv2 = pytest.importorskip("cv2")

    def _cv2_descriptors(img, xyt):      # augment_Edge_Data keypoints (Stereo_Matches.cpp:668-677), one batched compute
        kps = [cv2.KeyPoint(float(x + s * 8 * np.sin(t)), float(y - s * 8 * np.cos(t)), 1, float(180 / np.pi * t)) for x, y, t in xyt for s in (1, -1)]
        k2, d = cv2.SIFT_create().compute(img, kps)
        assert len(k2) == len(kps)
        return d.reshape(len(xyt), 2, 128).astype(np.float32)

    cal = synth.kitti_calib(480, 200)
    L, R = synth.stereo_pair(cal, 6)
    eL, _ = oracle.toed(L)
    eR, _ = oracle.toed(R)
    F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
    dL, dR = _cv2_descriptors(L, eL[:, :3]), _cv2_descriptors(R, eR[:, :3])
    res = oracle.stereo(L, R, eL, eR, F21, descL=dL, descR=dR, want_dumps=False)
    rows = _run_stereo_dropin(tmp_path, cal, L, R, eL, eR, sift=True)
    left = rows[:, 0].astype(int)
    common = np.intersect1d(res.mate_left, left)
    assert len(common) >= (1 - 2e-3) * len(res.mate_left) and len(left) <= (1 + 2e-3) * len(res.mate_left)
    # descriptors handed to the caller: left ones equal cv2's on the same keypoints up to the off-by-one entries
    want = dL[left, 0].sum(1)
    dsum = np.abs(rows[:, 14] - want)      # sums of 128 entries: a handful of off-by-one entries per descriptor at most
    assert (rows[:, 14] >= 0).all() and (rows[:, 15] >= 0).all() and dsum.max() <= 16 and (dsum > 0).mean() < 0.05


EXE_TEMPORAL = os.path.join(ROOT, "dropin", "_build", "test_dropin_temporal")


@pytest.mark.skipif(not os.path.exists(EXE_TEMPORAL), reason="dropin/_build not built (needs the reference headers at build time)")
def test_temporal_matches_member_backed_by_the_gpu_against_reference_output(tmp_path):
    """Pipeline::get_Temporal_Edge_Correspondences' call sequence on the reference's own classes, with
    get_Temporal_Edge_Pairs_from_Quads replaced by dropin/temporal_matches_b200.cpp, against the output of the
    reference's own CPU code on the same sequence pair (tests/golden/temporal_ref_small.npz)."""
    g = np.load(os.path.join(GOLDEN, "temporal_ref_small.npz"))
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    H, W = g["kfL"].shape
    with open(inp, "wb") as f:
        np.array([W, H, len(g["kf"]), len(g["cf"])], np.int32).tofile(f)
        for k in ("kfL", "kfR", "cfL", "cfR"):
            np.ascontiguousarray(g[k], np.uint8).tofile(f)
        np.ascontiguousarray(g["kf"], np.float64).tofile(f)
        np.ascontiguousarray(g["cf"], np.float64).tofile(f)
        np.ascontiguousarray(g["mask"], np.uint8).tofile(f)
    out = subprocess.run([EXE_TEMPORAL, str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    raw = np.fromfile(outp, np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    rows = raw[4:].view(np.float64).reshape(n, 13)
    off = g["cluster_off"]
    assert n == off[-1] > 1000
    own = np.repeat(np.arange(len(off) - 1), np.diff(off))
    assert np.array_equal(rows[:, 0].astype(int), own) and np.array_equal(rows[:, 1].astype(int), g["cluster_cf"])
    assert np.abs(rows[:, 2:4] - g["cluster_left"][:, :2]).max() < 1e-3 and np.abs(rows[:, 4] - g["cluster_left"][:, 2]).max() < 1e-4
    assert np.abs(rows[:, 5:7] - g["cluster_right"][:, :2]).max() < 1e-3 and np.abs(rows[:, 7] - g["cluster_right"][:, 2]).max() < 1e-4
    assert np.abs(rows[:, 8:10] - g["cluster_ncc"]).max() < 1e-5
    assert np.abs(rows[:, 10:12] - g["cluster_score"]).max() < 1e-5
    assert np.array_equal(rows[:, 12].astype(int), g["cluster_valid"])

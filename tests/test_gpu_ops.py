"""GPU: the small reference-facing operators (patches, NCC, clusterer, Sobel) against the oracle."""
import numpy as np
import pytest

import oracle
from edge_based_visual_odometry_b200 import synth, _lib

pytestmark = pytest.mark.gpu


def test_sobel_bit_exact(gpu_ctx):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (67, 91), dtype=np.uint8)
    gx, gy = gpu_ctx.sobel(img)
    ox, oy = oracle.sobel(img)
    assert np.array_equal(gx, ox) and np.array_equal(gy, oy)


def test_edge_patches_bit_exact_including_nan_rules(gpu_ctx, golden_stereo):
    g = golden_stereo
    e = g["eL"][:400].copy()
    e[0] = [30.0, 30.0, 0.0]            # integer coordinates -> NaN patch (utility.h:95-103)
    e[1] = [3.3, 50.2, 0.7]             # leaves the image -> NaN cells
    p, m = gpu_ctx.edge_patches(g["L"], _lib.edges_from_xyt(e))
    for k in range(len(e)):
        op, om = oracle.edge_patches(g["L"], *e[k])
        assert np.array_equal(p[k], op, equal_nan=True) and np.array_equal(m[k], om, equal_nan=True)
    assert np.isnan(p[0]).all() and np.isnan(m[1]).any()      # "-" side of edge 1 crosses x < 0


def test_ncc_patch_pairs(gpu_ctx):
    """Utility::get_patch_similarity / MatlabNCCComputer::computeNCC semantics: 1, -1, flat sentinel, NaN."""
    rng = np.random.default_rng(3)
    a = (rng.random((200, 49)) * 255).astype(np.float32)
    b = (a + rng.normal(0, 25, a.shape)).astype(np.float32)
    b[0] = a[0]; b[1] = 255 - a[1]; b[2] = 7.0; a[3, 5] = np.nan
    got = gpu_ctx.ncc(a, b)
    want = np.array([oracle.patch_similarity(x, y) for x, y in zip(a, b)])
    assert abs(got[0] - 1) < 1e-6 and abs(got[1] + 1) < 1e-6 and got[2] == -1.0 and np.isnan(got[3])
    ok = ~np.isnan(want)
    assert np.abs(got[ok] - want[ok]).max() < 1e-6


def test_clusterer_vs_oracle(gpu_ctx):
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(1, 40))
        # points along a line with sub-pixel jitter so that merges, size limits and orientation gates all occur
        x = 100 + np.cumsum(rng.choice([0.2, 0.6, 1.4], n)) + rng.normal(0, 0.05, n)
        y = 50 + rng.normal(0, 0.2, n)
        th = rng.choice([0.3, 0.35, 0.9], n) + rng.normal(0, 0.02, n)
        pts = np.stack([x, y, th], 1)
        for by_orient in (True, False):
            cen_o, lab_o = oracle.cluster(pts, by_orient)
            cen_g, lab_g = gpu_ctx.cluster(_lib.edges_from_xyt(pts), by_orient)
            assert len(cen_g) == len(cen_o) and np.array_equal(lab_g, lab_o)
            assert np.abs(cen_g["x"] - cen_o[:, 0]).max() < 1e-9 and np.abs(cen_g["theta"] - cen_o[:, 2]).max() < 1e-9


@pytest.mark.parametrize("side", ["left", "right"])
def test_undistort_bit_identical_to_cv2(gpu_ctx, side):
    """Pipeline::prepare_Stereo_Images calls cv::undistort with the YAML's 4 coefficients (Pipeline.cpp:78-79); OpenCV's
    arithmetic (FP64 map, 5-bit fixed-point bilinear remap, stripes of 4096 / cols rows) restated in undistort.cu and
    compared with cv2 itself on the EuRoC calibration (config/euroc.yaml:12-18), every pixel."""
    cv2 = pytest.importorskip("cv2")
    cal = synth.CALIBS["euroc"]()
    img = synth.stereo_pair(cal, 1)[0 if side == "left" else 1]
    if side == "left":
        K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1.0]]); dist = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    else:
        K = np.array([[457.587, 0, 379.999], [0, 456.134, 255.238], [0, 0, 1.0]]); dist = np.array([-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05])
    got = gpu_ctx.undistort(img, K, dist)
    want = cv2.undistort(img, K, dist)
    assert np.array_equal(got, want)
    zero = gpu_ctx.undistort(img, K, np.zeros(4))            # KITTI / ETH3D: zero distortion is the identity
    assert np.array_equal(zero, img)

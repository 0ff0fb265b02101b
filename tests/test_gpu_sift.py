"""GPU: SIFT descriptors at the reference's edge keypoints (S4 / S7' / S13, "SIFT-on") against OpenCV itself.

cv::SIFT is third-party code (OpenCV, not in the reference repository; cv2 4.13 is in this image): the descriptor
kernel restates cv::SIFT::compute for provided keypoints (sift.cu) and is pinned here against cv2 on the same image.
OpenCV's SIMD summation orders (Gaussian blur, histogram votes) are not reproducible bit for bit, so the bar is:
>= 99.9 % of the 8-bit descriptor entries identical, the rest off by one; SIFT distances within 4 units."""
import numpy as np
import pytest

import oracle
from edge_based_visual_odometry_b200 import synth, _lib

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _cv2_descriptors(img, xyt):
    """augment_Edge_Data / apply_SIFT_filtering keypoints (Stereo_Matches.cpp:668-677,720-727), one batched compute."""
    kps = []
    for x, y, t in xyt:
        for s in (1, -1):
            kps.append(cv2.KeyPoint(float(x + s * 8 * np.sin(t)), float(y - s * 8 * np.cos(t)), 1, float(180 / np.pi * t)))
    k2, d = cv2.SIFT_create().compute(img, kps)
    assert len(k2) == len(kps)
    return d.reshape(len(xyt), 2, 128).astype(np.float32)


def _ctx(w, h, **kw):
    prm = _lib.default_params(); prm.sift_mode = 1
    return _lib.Context(0, w, h, max_batch=1, max_edges=65536, params=prm, **kw)


@pytest.mark.parametrize("shape", [(640, 240), (321, 203)])
def test_descriptors_against_cv2(shape):
    cal = synth.kitti_calib(*shape)
    img, _ = synth.stereo_pair(cal, 3)
    e, _ = oracle.toed(img)
    # add keypoints that leave the image (the border tests of calcSIFTDescriptor) and every orientation sign
    extra = np.array([[1.0, 1.0, 0.3], [shape[0] - 2.0, shape[1] - 2.0, -2.0], [5.5, shape[1] / 2, 3.1], [shape[0] / 2, 2.0, -3.1], [30.0, 30.0, 0.0]])
    xyt = np.vstack([e[:, :3], extra])
    ctx = _ctx(*shape)
    got = ctx.sift_descriptors(img, _lib.edges_from_xyt(xyt))
    got2 = ctx.sift_descriptors(img, _lib.edges_from_xyt(xyt))
    ctx.close()
    want = _cv2_descriptors(img, xyt)
    assert np.array_equal(got, got2)                                   # deterministic (integer vote accumulation)
    diff = np.abs(got - want)
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3, (diff.max(), (diff > 0).mean())
    d_got = np.sqrt(((got[:, 0] - got[:, 1]) ** 2).sum(1)); d_want = np.sqrt(((want[:, 0] - want[:, 1]) ** 2).sum(1))
    assert np.abs(d_got - d_want).max() < 4.0


def test_stereo_sift_on_against_oracle_with_cv2_descriptors():
    """The whole matcher with the SIFT gate (S4) and BNB-SIFT (S7') fed by device descriptors, against the oracle fed
    by cv2 descriptors.  Off-by-one descriptor entries move SIFT distances by < 2 units, so candidates within that
    distance of the 500 gate (or of the 0.4 ratio) may flip: the budget is the north star's 0.1 % of the left edges."""
    cal = synth.kitti_calib(640, 240)
    L, R = synth.stereo_pair(cal, 2)
    eL, _ = oracle.toed(L)
    eR, _ = oracle.toed(R)
    F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
    dL, dR = _cv2_descriptors(L, eL[:, :3]), _cv2_descriptors(R, eR[:, :3])
    res = oracle.stereo(L, R, eL, eR, F21, descL=dL, descR=dR)
    off = oracle.stereo(L, R, eL, eR, F21, want_dumps=False)
    ctx = _ctx(640, 240)
    ctx.set_stage_dumps(True)
    calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
    mates = ctx.stereo_match(calib, L, R, _lib.edges_from_xyt(eL), _lib.edges_from_xyt(eR))
    sg, so = ctx.stage("sift"), res.stages["sift"]
    ctx.close()
    assert len(res.mate_left) != len(off.mate_left)                    # the gate is not a no-op on this pair
    nL = len(eL)
    same = np.diff(sg["off"]) == np.diff(so["off"])
    assert (~same).mean() <= 1e-3                                      # S4 survivor lists
    if same.all():
        assert (sg["ridx"] != so["ridx"]).mean() <= 1e-3
    common = np.intersect1d(res.mate_left, mates["left_index"])
    assert len(common) >= (1 - 2e-3) * len(res.mate_left) and len(mates) <= (1 + 2e-3) * len(res.mate_left)
    io = np.searchsorted(res.mate_left, common); ig = np.searchsorted(mates["left_index"], common)
    d = np.hypot(res.mate_right[io, 0] - mates["rx"][ig], res.mate_right[io, 1] - mates["ry"][ig])
    assert (d > 1e-3).mean() <= 2e-3


def test_sift_on_batch_equals_single_frames():
    """sift_mode 1 through the pipelined batch call (descriptor buffers are per frame; sub-batch views shift them)."""
    cal = synth.kitti_calib(320, 200)
    pairs = [synth.stereo_pair(cal, f) for f in range(3)]
    Ls = [pairs[f % 3][0] for f in range(35)]
    Rs = [pairs[f % 3][1] for f in range(35)]
    prm = _lib.default_params(); prm.sift_mode = 1
    ctx = _lib.Context(0, 320, 200, max_batch=35, max_edges=16384, params=prm)
    calib = _lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21)
    out, n = ctx.stereo_batch(calib, Ls, Rs, cap=8000)
    singles = [ctx.stereo_frame(calib, Ls[f], Rs[f], want_edges=False) for f in range(3)]
    prm0 = _lib.default_params()
    ctx0 = _lib.Context(0, 320, 200, max_batch=1, max_edges=16384, params=prm0)
    off = ctx0.stereo_frame(calib, Ls[0], Rs[0], want_edges=False)
    ctx.close(); ctx0.close()
    for f in range(35):
        assert n[f] == len(singles[f % 3]) and np.array_equal(out[f, :n[f]], singles[f % 3])
    assert len(off) != len(singles[0])          # SIFT-on really ran (it prunes candidates)

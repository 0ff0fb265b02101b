import numpy as np

from edge_based_visual_odometry_b200 import synth


def test_deterministic_and_shapes():
    for name, (w, h) in {"kitti": (1241, 376), "euroc": (752, 480), "eth3d": (742, 464)}.items():
        cal = synth.CALIBS[name]()
        assert (cal.width, cal.height) == (w, h)
    cal = synth.kitti_calib(320, 200)
    a = synth.stereo_pair(cal, 3)
    b = synth.stereo_pair(cal, 3)
    c = synth.stereo_pair(cal, 4)
    assert a[0].dtype == np.uint8 and a[0].shape == (200, 320)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0], c[0])


def test_rectified_layers_shift_by_their_disparity():
    """x_R = x_L - d for the rectified KITTI calibration (Stereo_Matches.cpp:159 sign convention)."""
    cal = synth.kitti_calib(320, 200)
    layers, bg = synth.make_scene(cal, 5)
    L = layers[-1]
    H = synth._layer_homography(cal, L["Z"])
    p = H @ np.array([100.0, 80.0, 1.0])
    assert abs(p[0] / p[2] - (100.0 - L["disp"])) < 1e-9 and abs(p[1] / p[2] - 80.0) < 1e-9


def test_general_calibration_layers_satisfy_the_fundamental_matrix():
    cal = synth.euroc_calib()
    F21, _ = synth.fundamental_matrices(cal)
    layers, _ = synth.make_scene(cal, 2)
    for L in layers[:5]:
        H = synth._layer_homography(cal, L["Z"])
        pl = np.array([300.0, 200.0, 1.0])
        pr = H @ pl
        pr /= pr[2]
        l = F21 @ pl
        assert abs(pr @ l) / np.hypot(l[0], l[1]) < 1e-6

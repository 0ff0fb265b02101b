"""CPU: the TOED restatement (oracle/toed_oracle.c) against the unmodified reference and analytic cases."""
import numpy as np
import pytest

import oracle
from edge_based_visual_odometry_b200 import synth


def test_restatement_matches_golden_reference(golden_toed):
    """Golden vector produced by /root/reference/src/toed/cpu_toed.cpp compiled in place (tests/golden/make_golden.py)."""
    e, nt = oracle.toed(golden_toed["image"])
    ref = golden_toed["edges"]
    assert nt == int(golden_toed["n_total"])
    assert e.shape == ref.shape
    assert np.abs(e[:, :2] - ref[:, :2]).max() < 1e-9      # px
    assert np.abs(e[:, 2] - ref[:, 2]).max() < 1e-9        # rad


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape,seed", [((97, 61), 3), ((200, 152), 5), ((333, 129), 9)])
def test_restatement_matches_compiled_reference(shape, seed):
    cal = synth.kitti_calib(*shape)
    L, R = synth.stereo_pair(cal, seed, density=2.0)
    for img in (L, R):
        e, nt = oracle.toed(img)
        ref, ntr, _, _ = oracle.toed_reference(img)
        assert nt == ntr and e.shape == ref.shape
        if len(e):
            assert np.abs(e - ref).max() < 1e-9


def test_vertical_step_edge_subpixel():
    """A blurred vertical step at x = 60.5 (between pixel 60 and 61): edges at x ~ 60.5, tangent vertical."""
    H, W = 80, 120
    img = np.zeros((H, W), np.uint8)
    img[:, 61:] = 200
    e, _ = oracle.toed(img)
    assert len(e) > 50
    # reference convention: final x = (X-1)/2 on the 2x grid (cpu_toed.cpp:538) => -0.5 px offset w.r.t. the step
    assert np.abs(e[:, 0] - 60.0).max() < 0.05
    assert np.abs(np.abs(e[:, 2]) - np.pi / 2).max() < 1e-6
    assert e[:, 1].min() > 10 and e[:, 1].max() < H - 10


def test_constant_image_has_no_edges():
    e, nt = oracle.toed(np.full((64, 64), 77, np.uint8))
    assert len(e) == 0 and nt == 0


def test_edges_are_in_row_major_interp_order(golden_toed):
    e, _, maps = oracle.toed(golden_toed["image"], want_maps=True)
    spx = maps[4]
    ii, jj = np.nonzero(spx)
    keep = []
    H, W = golden_toed["image"].shape
    for i, j in zip(ii, jj):
        x, y = (maps[4][i, j] - 1) / 2, (maps[5][i, j] - 1) / 2
        if 10 < x < W - 10 and 10 < y < H - 10:
            keep.append((x, y))
    assert np.allclose(np.array(keep), e[:, :2], atol=0, rtol=0)

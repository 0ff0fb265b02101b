import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_toed():
    return np.load(os.path.join(GOLDEN, "toed_ref_small.npz"))


@pytest.fixture(scope="session")
def golden_stereo():
    return np.load(os.path.join(GOLDEN, "stereo_small.npz"))


@pytest.fixture(scope="session")
def kitti_case():
    """Full-size KITTI-shape pair + oracle edges + oracle stereo result (a few seconds of CPU)."""
    import oracle
    from edge_based_visual_odometry_b200 import synth
    cal = synth.kitti_calib()
    L, R = synth.stereo_pair(cal, 0)
    eL, ntL = oracle.toed(L)
    eR, ntR = oracle.toed(R)
    F21, _ = oracle.fundamental(cal.Kl, cal.Kr, cal.R21, cal.T21)
    res = oracle.stereo(L, R, eL, eR, F21)
    return dict(cal=cal, L=L, R=R, eL=eL, eR=eR, ntL=ntL, ntR=ntR, F21=F21, res=res)


@pytest.fixture(scope="session")
def gpu_ctx():
    from edge_based_visual_odometry_b200 import _lib
    ctx = _lib.Context(0, 1241, 480, max_batch=4, max_edges=131072)
    yield ctx
    ctx.close()

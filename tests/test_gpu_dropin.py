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

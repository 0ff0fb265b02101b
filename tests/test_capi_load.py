"""CPU: the C-ABI library loads, exports every symbol include/ebvo_b200.h declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from edge_based_visual_odometry_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "ebvo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ebvo_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_library_agree():
    names = _declared_functions()
    assert len(names) >= 20
    L = _lib.load()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/ebvo_b200.h but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == names


def test_struct_layouts():
    assert C.sizeof(_lib.Edge) == 32 and C.sizeof(_lib.Mate) == 64 and C.sizeof(_lib.Calib) == 30 * 8
    p = _lib.default_params()
    # reference defaults, include/definitions.h:17-36
    assert (p.epipolar_line_dist_thresh, p.max_disparity, p.ncc_thresh, p.bnb_ncc) == (0.5, 25.0, 0.6, 0.9)
    assert (p.location_perturbation, p.epip_tangency_displ_thresh, p.orient_perturbation) == (0.4, 3.0, 0.174533)
    assert (p.cluster_dist_thresh, p.cluster_orient_thresh_deg, p.max_cluster_size) == (1.0, 20.0, 10)
    assert (p.gn_max_iter, p.gn_tol, p.gn_huber_delta, p.gn_mode) == (20, 1e-3, 3.0, 0)


def test_fundamental_is_host_side_and_matches_reference_formula():
    from edge_based_visual_odometry_b200 import synth
    for name in ("kitti", "euroc", "eth3d"):
        cal = synth.CALIBS[name]()
        F21, F12 = _lib.fundamental(_lib.make_calib(cal.Kl, cal.Kr, cal.R21, cal.T21))
        F21n, F12n = synth.fundamental_matrices(cal)
        assert np.abs(F21 - F21n).max() < 1e-15 and np.abs(F12 - F12n).max() < 1e-15


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.EbvoError) as e:
        _lib.Context()
    assert e.value.code == -1      # EBVO_ERR_NO_DEVICE


def test_product_package_does_not_import_the_oracle():
    import subprocess, sys
    code = "import sys; import edge_based_visual_odometry_b200 as p; from edge_based_visual_odometry_b200 import _lib, synth, sharding; print('oracle' in sys.modules)"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.stdout.strip() == "False", out.stderr
    for root, _, files in os.walk(os.path.join(ROOT, "edge_based_visual_odometry_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                assert "oracle" not in open(os.path.join(root, f)).read(), f


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/ebvo_b200.h compiles as C99 (and as C++) on its own, without warnings."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "ebvo_b200.h"\nint main(void) { ebvo_params p; ebvo_mate m; ebvo_quad q; ebvo_calib c; (void)p; (void)m; (void)q; (void)c;\n'
                   '  return (int)sizeof(ebvo_edge) - 32; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only"], ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++"]):
        r = subprocess.run(cmd + ["-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr

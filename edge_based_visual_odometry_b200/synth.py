"""Seeded synthetic stereo pairs at the reference's dataset shapes.

Datasets (KITTI / EuRoC / ETH3D) are not available offline, so every test and bench
in this repo runs on synthetic pairs built here:

* a smooth low-frequency background at a far depth (about 2 px of disparity),
* ``n_obj = density * W * H / 2300`` filled, anti-aliased ellipses / convex polygons,
  each a fronto-parallel layer at its own depth, painted far-to-near in both views,
* Gaussian blur sigma = 1, i.i.d. N(0, 1) noise with separate left/right streams,
  rounding and clipping to uint8.

Every layer of depth Z is rendered into the right view through the plane-induced
homography ``K_r (R21 + T21 n^T / Z) K_l^-1`` with n = (0, 0, 1), so true
correspondences satisfy the fundamental matrix the reference derives from the same
calibration (``/root/reference/src/Dataset.cpp:102-112``) exactly.  For the rectified
KITTI calibration this reduces to ``x_R = x_L - d`` with d = fx * |Tx| / Z
(the sign convention of ``/root/reference/src/Stereo_Matches.cpp:159``).

Calibration numbers are copied from the reference YAML files
(``/root/reference/config/kitti.yaml:13-28``, ``euroc.yaml:11-28``,
``eth3d_cable_2.yaml:12-29``); only numbers, no code.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy import ndimage

SEED0 = 20260000


@dataclass
class Calibration:
    """Stereo calibration, the numeric content of a reference YAML file."""

    name: str
    width: int
    height: int
    Kl: np.ndarray
    Kr: np.ndarray
    R21: np.ndarray
    T21: np.ndarray
    # translation used to RENDER the right view (differs from T21 only for KITTI, whose
    # YAML lists +0.54 while the physical right camera sits at -0.54; F is identical up
    # to sign, so the epipolar geometry seen by the matcher is the same)
    T_render: np.ndarray = field(default=None)

    def __post_init__(self):
        if self.T_render is None:
            self.T_render = self.T21.copy()


def _K(fx, fy, cx, cy):
    return np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float64)


def kitti_calib(width=1241, height=376) -> Calibration:
    K = _K(718.856, 718.856, 607.1928, 185.2157)
    return Calibration("kitti", width, height, K, K.copy(), np.eye(3),
                       np.array([0.54, 0.0, 0.0]), np.array([-0.54, 0.0, 0.0]))


def euroc_calib(width=752, height=480) -> Calibration:
    Kl = _K(458.654, 457.296, 367.215, 248.375)
    Kr = _K(457.587, 456.134, 379.999, 255.238)
    R = np.array([[0.999997256477450, 0.002312067192420, 0.000376008102351],
                  [-0.002317135723285, 0.999898048506528, 0.014089835846697],
                  [-0.000343393120589, -0.014090668452670, 0.999900662638179]])
    T = np.array([-0.110073808127139, 0.000399121547014534, -0.000853702503351098])
    return Calibration("euroc", width, height, Kl, Kr, R, T)


def eth3d_cable2_calib(width=742, height=464) -> Calibration:
    Kl = _K(726.04388427734, 726.04388427734, 385.73248291016, 262.19641113281)
    Kr = _K(726.28741455078, 726.28741455078, 354.6496887207, 186.46566772461)
    R = np.array([[0.99992036819458, 0.012368063442409, -0.0024963859468699],
                  [-0.012378259561956, 0.99991494417191, -0.0041112052276731],
                  [0.0024453257210553, 0.0041417786851525, 0.99998843669891]])
    T = np.array([-0.089831538498402, -0.0001915143802762, 0.00019389642693568])
    return Calibration("eth3d_cable_2", width, height, Kl, Kr, R, T)


def kitti4k_calib(width=3840, height=2160) -> Calibration:
    s = 3840.0 / 1241.0  # = 3.094, KITTI intrinsics scaled to the 4K stress shape
    K = _K(718.856 * s, 718.856 * s, 607.1928 * s, 185.2157 * s)
    return Calibration("kitti4k", width, height, K, K.copy(), np.eye(3),
                       np.array([0.54, 0.0, 0.0]), np.array([-0.54, 0.0, 0.0]))


CALIBS = {"kitti": kitti_calib, "euroc": euroc_calib, "eth3d": eth3d_cable2_calib,
          "kitti4k": kitti4k_calib}


def _layer_homography(cal: Calibration, Z: float) -> np.ndarray:
    """Left pixel -> right pixel for the plane Z = const in the left camera frame."""
    n = np.array([0.0, 0.0, 1.0])
    return cal.Kr @ (cal.R21 + np.outer(cal.T_render, n) / Z) @ np.linalg.inv(cal.Kl)


def _coverage(kind, prm, u, v):
    """Anti-aliased coverage in [0,1] of one shape at left-view coordinates (u, v)."""
    cx, cy, a, b, phi = prm[:5]
    c, s = np.cos(phi), np.sin(phi)
    du, dv = u - cx, v - cy
    p = c * du + s * dv
    q = -s * du + c * dv
    if kind == 0:  # ellipse: first-order signed distance
        r = np.sqrt((p / a) ** 2 + (q / b) ** 2)
        g = np.sqrt((p / (a * a)) ** 2 + (q / (b * b)) ** 2) + 1e-12
        sd = (r * r - 1.0) / (2.0 * g)
        sd = np.where(r < 1e-6, -min(a, b), sd)
    else:  # convex polygon with `kind` sides inscribed in the ellipse (a, b)
        nside = kind
        sd = np.full(u.shape, -1e9)
        ang = 2.0 * np.pi * (np.arange(nside) + 0.5) / nside
        vx, vy = a * np.cos(2.0 * np.pi * np.arange(nside) / nside), b * np.sin(2.0 * np.pi * np.arange(nside) / nside)
        for k in range(nside):
            x0, y0 = vx[k], vy[k]
            x1, y1 = vx[(k + 1) % nside], vy[(k + 1) % nside]
            ex, ey = x1 - x0, y1 - y0
            ln = np.hypot(ex, ey)
            nx, ny = ey / ln, -ex / ln  # outward normal for counter-clockwise vertices
            sd = np.maximum(sd, (p - x0) * nx + (q - y0) * ny)
        del ang
    return np.clip(0.5 - sd, 0.0, 1.0)


def _render(cal, layers, bg, view):
    H, W = cal.height, cal.width
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)

    def to_left(Z):
        if view == "left":
            return xx, yy
        Hinv = np.linalg.inv(_layer_homography(cal, Z))
        w = Hinv[2, 0] * xx + Hinv[2, 1] * yy + Hinv[2, 2]
        return ((Hinv[0, 0] * xx + Hinv[0, 1] * yy + Hinv[0, 2]) / w,
                (Hinv[1, 0] * xx + Hinv[1, 1] * yy + Hinv[1, 2]) / w)

    u, v = to_left(bg["Z"])
    u, v = u + bg.get("ou", 0.0), v + bg.get("ov", 0.0)   # camera translation of a sequence frame (stereo_sequence_pair)
    img = (bg["base"] + bg["ax"] * np.sin(u * bg["fx"] + bg["px"]) + bg["ay"] * np.cos(v * bg["fy"] + bg["py"])
           + bg["gx"] * (u / W - 0.5) + bg["gy"] * (v / H - 0.5))
    for L in layers:  # far -> near
        Hm = _layer_homography(cal, L["Z"]) if view == "right" else np.eye(3)
        # bounding box of the shape in this view
        rad = max(L["prm"][2], L["prm"][3]) + 3.0
        c = Hm @ np.array([L["prm"][0], L["prm"][1], 1.0])
        cxv, cyv = c[0] / c[2], c[1] / c[2]
        x0, x1 = int(max(0, np.floor(cxv - rad - 2))), int(min(W, np.ceil(cxv + rad + 3)))
        y0, y1 = int(max(0, np.floor(cyv - rad - 2))), int(min(H, np.ceil(cyv + rad + 3)))
        if x0 >= x1 or y0 >= y1:
            continue
        sx, sy = xx[y0:y1, x0:x1], yy[y0:y1, x0:x1]
        if view == "left":
            uu, vv = sx, sy
        else:
            Hinv = np.linalg.inv(Hm)
            w = Hinv[2, 0] * sx + Hinv[2, 1] * sy + Hinv[2, 2]
            uu = (Hinv[0, 0] * sx + Hinv[0, 1] * sy + Hinv[0, 2]) / w
            vv = (Hinv[1, 0] * sx + Hinv[1, 1] * sy + Hinv[1, 2]) / w
        cov = _coverage(L["kind"], L["prm"], uu, vv)
        tone = L["gray"] + L["tx"] * (uu - L["prm"][0]) + L["ty"] * (vv - L["prm"][1])
        img[y0:y1, x0:x1] = img[y0:y1, x0:x1] * (1.0 - cov) + tone * cov
    return img


def make_scene(cal: Calibration, seed: int, density: float = 1.0, disp_range=(2.0, 22.0)):
    """Scene description (layers + background) for one frame; deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    W, H = cal.width, cal.height
    fxB = abs(cal.Kl[0, 0] * cal.T_render[0])
    n_obj = max(1, int(round(density * W * H / 2300.0)))
    layers = []
    for _ in range(n_obj):
        d = rng.uniform(*disp_range)
        kind = int(rng.choice([0, 0, 3, 4, 5, 6]))
        a = rng.uniform(6.0, 42.0) * (W / 1241.0) ** 0.5
        b = a * rng.uniform(0.35, 1.0)
        layers.append(dict(Z=fxB / d, disp=d, kind=kind,
                           prm=(rng.uniform(0, W), rng.uniform(0, H), a, b, rng.uniform(0, np.pi)),
                           gray=rng.uniform(25.0, 230.0), tx=rng.uniform(-0.6, 0.6), ty=rng.uniform(-0.6, 0.6)))
    layers.sort(key=lambda L: L["disp"])  # far (small disparity) first
    bg = dict(Z=fxB / 2.0, base=rng.uniform(90, 150), ax=rng.uniform(8, 20), ay=rng.uniform(8, 20),
              fx=rng.uniform(0.004, 0.012), fy=rng.uniform(0.006, 0.02), px=rng.uniform(0, 6.28),
              py=rng.uniform(0, 6.28), gx=rng.uniform(-30, 30), gy=rng.uniform(-30, 30))
    return layers, bg


def stereo_pair(cal: Calibration | str = "kitti", frame: int = 0, density: float = 1.0,
                seed0: int = SEED0, noise_sigma: float = 1.0, blur_sigma: float = 1.0):
    """Return (left_u8, right_u8) for frame `frame` (seed = seed0 + frame)."""
    if isinstance(cal, str):
        cal = CALIBS[cal]()
    seed = seed0 + frame
    layers, bg = make_scene(cal, seed, density)
    out = []
    for k, view in enumerate(("left", "right")):
        img = _render(cal, layers, bg, view)
        if blur_sigma > 0:
            img = ndimage.gaussian_filter(img, blur_sigma, mode="nearest")
        nrng = np.random.default_rng([seed, 7919 + k])
        img = img + noise_sigma * nrng.standard_normal(img.shape)
        out.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    return out[0], out[1]


def stereo_sequence_pair(cal: Calibration | str = "kitti", frame: int = 0, step=(0.35, 0.1), scene_seed: int = SEED0 + 4000,
                         density: float = 1.0, noise_sigma: float = 1.0, blur_sigma: float = 1.0):
    """Frame `frame` of a SEQUENCE: one layered scene (seed `scene_seed`) seen by a stereo rig that translates by
    `frame * step` baselines along (x, y) between frames (BASELINE.json configs[3]: keyframe -> current-frame edge
    tracking needs temporal correspondences, which independent per-frame scenes do not have).  A fronto-parallel
    layer of stereo disparity d moves by -d * frame * step pixels in both views; noise is fresh per frame.
    Returns (left_u8, right_u8, pose) with pose = camera translation in units of the baseline."""
    if isinstance(cal, str):
        cal = CALIBS[cal]()
    layers, bg = make_scene(cal, scene_seed, density)
    sx, sy = frame * step[0], frame * step[1]
    moved = []
    for L in layers:
        M = dict(L)
        cx, cy, a, b, phi = L["prm"]
        M["prm"] = (cx - L["disp"] * sx, cy - L["disp"] * sy, a, b, phi)
        moved.append(M)
    bg = dict(bg, ou=2.0 * sx, ov=2.0 * sy)    # the background plane sits at 2 px of disparity (make_scene)
    out = []
    for k, view in enumerate(("left", "right")):
        img = _render(cal, moved, bg, view)
        if blur_sigma > 0:
            img = ndimage.gaussian_filter(img, blur_sigma, mode="nearest")
        nrng = np.random.default_rng([scene_seed, frame, 7919 + k])
        img = img + noise_sigma * nrng.standard_normal(img.shape)
        out.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    return out[0], out[1], (sx, sy)


def position_descriptors(xyt, seed: int = 0) -> np.ndarray:
    """Deterministic stand-in for the (n, 2, 128) SIFT descriptor pairs of a set of edges: integer-valued float32 entries
    in [0, 255] that vary smoothly with the edge position and orientation, so that nearby, similarly oriented edges of two
    frames have close descriptors and distant ones do not.  Used where the stage logic around descriptors is under test
    (gates, best-nearly-best) and real cv::SIFT output would have to be stored."""
    xyt = np.asarray(xyt, np.float64).reshape(-1, 3)
    k = np.arange(128, dtype=np.float64)[None, :]
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, (2, 128))
    out = np.zeros((len(xyt), 2, 128), np.float32)
    for j in range(2):
        v = 110 + 60 * np.sin(xyt[:, :1] * (0.02 + 0.0007 * k) + ph[j]) + 50 * np.cos(xyt[:, 1:2] * (0.03 + 0.0005 * k) + 2 * ph[j]) \
            + 30 * np.sin(2 * xyt[:, 2:3] + 0.05 * k + j)
        out[:, j] = np.clip(np.rint(v), 0, 255)
    return out


def fundamental_matrices(cal: Calibration):
    """F21 and F12 exactly as the reference builds them (Dataset.cpp:102-112).

    Host-side float64 3x3 algebra only (used by the tests as an independent check)."""
    def skew(t):
        return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], dtype=np.float64)
    Kl_inv, Kr_inv = np.linalg.inv(cal.Kl), np.linalg.inv(cal.Kr)
    F21 = Kr_inv.T @ (skew(cal.T21) @ cal.R21) @ Kl_inv
    R12 = cal.R21.T
    T12 = -cal.R21.T @ cal.T21
    F12 = Kl_inv.T @ (skew(T12) @ R12) @ Kr_inv
    return F21, F12

"""TEST INFRASTRUCTURE ONLY: ctypes front-end of the CPU oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``edge_based_visual_odometry_b200`` never imports it.

Two libraries:

* ``oracle/libebvo_oracle.so``  - our restatement (``toed_oracle.c`` + ``stereo_oracle.cpp``)
* ``oracle/_ref/libtoed_ref.so`` - the UNMODIFIED reference TOED source compiled in place from
  ``/root/reference/src/toed/cpu_toed.cpp`` (built here, shipped to the GPU box as a binary)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

STAGES = ["epi", "disp", "orient", "sift", "ncc", "bnb_ncc", "bnb_sift", "shift", "gn", "cluster", "ncc2", "best"]


def build(force: bool = False) -> None:
    """Compile the restatement and, when /root/reference exists, oracle/_ref."""
    need = force or not os.path.exists(os.path.join(_DIR, "libebvo_oracle.so"))
    if not need:
        so = os.path.getmtime(os.path.join(_DIR, "libebvo_oracle.so"))
        need = any(os.path.getmtime(os.path.join(_DIR, f)) > so for f in ("toed_oracle.c", "stereo_oracle.cpp", "temporal_oracle.inl"))
    if need:
        subprocess.check_call(["make", "-C", _DIR, "libebvo_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/toed/cpu_toed.cpp") and (
            force or not os.path.exists(os.path.join(_DIR, "_ref", "libtoed_ref.so"))
            or not os.path.exists(os.path.join(_DIR, "_ref", "libstereo_ref.so"))
            or not os.path.exists(os.path.join(_DIR, "_ref", "libtemporal_ref.so"))):
        subprocess.check_call(["make", "-C", _DIR, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_DIR, "libebvo_oracle.so"))
        L.toed_oracle.restype = C.c_int
        L.so_run.restype = C.c_void_p
        L.so_patch_similarity.restype = C.c_double
        L.so_num_mates.restype = C.c_int
        L.so_stage_total.restype = C.c_int
        L.so_cluster.restype = C.c_int
        _LIB = L
    return _LIB


def have_ref() -> bool:
    return os.path.exists(os.path.join(_DIR, "_ref", "libtoed_ref.so"))


def have_stereo_ref() -> bool:
    return os.path.exists(os.path.join(_DIR, "_ref", "libstereo_ref.so"))


_SREF = None


def stereo_ref_lib():
    """oracle/_ref/libstereo_ref.so: the UNMODIFIED reference stereo sources compiled in place against third_party_shim."""
    global _SREF
    if _SREF is None:
        build()
        R = C.CDLL(os.path.join(_DIR, "_ref", "libstereo_ref.so"))
        R.rs_run.restype = C.c_void_p
        R.rs_num_mates.restype = C.c_int
        R.rs_stage_total.restype = C.c_int
        R.rs_patch_similarity.restype = C.c_double
        R.rs_cluster.restype = C.c_int
        _SREF = R
    return _SREF


def ref():
    global _REF
    if _REF is None:
        build()
        R = C.CDLL(os.path.join(_DIR, "_ref", "libtoed_ref.so"))
        R.toed_ref_run.restype = C.c_int
        R.toed_ref_create.restype = C.c_void_p
        R.toed_ref_detect.restype = C.c_int
        R.toed_ref_fetch.restype = C.c_int
        R.toed_ref_num_procs.restype = C.c_int
        _REF = R
    return _REF


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def toed_reference(img: np.ndarray, threads: int = 0):
    """Run the compiled reference detector.  Returns (edges[n,3], n_total, time_conv, time_nms)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    cap = 4 * H * W
    out = np.zeros((cap, 3))
    nt, tc, tn = C.c_int(), C.c_double(), C.c_double()
    n = ref().toed_ref_run(_p(img), H, W, W, _p(out), cap, C.byref(nt), None, 0, C.byref(tc), C.byref(tn), threads)
    return out[:n].copy(), nt.value, tc.value, tn.value


def toed(img: np.ndarray, want_maps: bool = False):
    """Run the restatement.  Returns (edges[n,3], n_total[, maps[7,2H,2W]])."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    cap = 4 * H * W
    out = np.zeros((cap, 3))
    nt = C.c_int()
    maps = np.zeros((7, 2 * H, 2 * W)) if want_maps else None
    n = lib().toed_oracle(_p(img), H, W, W, _p(out), cap, C.byref(nt), _p(maps), None, 0)
    if want_maps:
        return out[:n].copy(), nt.value, maps
    return out[:n].copy(), nt.value


def fundamental(Kl, Kr, R21, T21):
    a = [np.ascontiguousarray(m, dtype=np.float64) for m in (Kl, Kr, R21, T21)]
    F21, F12 = np.zeros((3, 3)), np.zeros((3, 3))
    lib().so_fundamental(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(F21), _p(F12))
    return F21, F12


def sobel(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    gx, gy = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    lib().so_sobel(_p(img), H, W, _p(gx), _p(gy))
    return gx, gy


def edge_patches(img, x, y, th):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    p, m = np.zeros(49, np.float32), np.zeros(49, np.float32)
    lib().so_edge_patches(_p(img), H, W, C.c_double(x), C.c_double(y), C.c_double(th), _p(p), _p(m))
    return p.reshape(7, 7), m.reshape(7, 7)


def patch_similarity(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32).ravel()
    b = np.ascontiguousarray(b, dtype=np.float32).ravel()
    return lib().so_patch_similarity(_p(a), _p(b))


def cluster(xyt, by_orientation=True):
    xyt = np.ascontiguousarray(xyt, dtype=np.float64).reshape(-1, 3)
    n = len(xyt)
    cen, lab = np.zeros((max(n, 1), 3)), np.zeros(max(n, 1), np.int32)
    k = lib().so_cluster(_p(xyt), n, int(by_orientation), _p(cen), _p(lab))
    return cen[:k].copy(), lab[:n].copy()


def shift_to_line(line, x, y, th):
    line = np.ascontiguousarray(line, dtype=np.float64)
    out = np.zeros(3)
    lib().so_shift_to_line(_p(line), C.c_double(x), C.c_double(y), C.c_double(th), _p(out))
    return out


class StereoResult:
    """Stage dumps + final mates of one oracle run."""

    def __init__(self, h, nL, want_dumps):
        L = lib()
        n = L.so_num_mates(C.c_void_p(h))
        self.mate_left = np.zeros(n, np.int32)
        self.mate_right = np.zeros((n, 3))
        rx, ry, rth = np.zeros(n), np.zeros(n), np.zeros(n)
        self.mate_score = np.zeros(n)
        L.so_get_mates(C.c_void_p(h), _p(self.mate_left), _p(rx), _p(ry), _p(rth), _p(self.mate_score))
        self.mate_right[:, 0], self.mate_right[:, 1], self.mate_right[:, 2] = rx, ry, rth
        self.lines = np.zeros((nL, 3))
        L.so_get_lines(C.c_void_p(h), _p(self.lines))
        t = np.zeros(len(STAGES))
        cnt = (C.c_long * 5)()
        L.so_get_stats(C.c_void_p(h), _p(t), cnt)
        self.stage_seconds = dict(zip(STAGES, t.tolist()))
        self.counts = dict(zip(["s1_total", "ncc_pairs1", "ncc_pairs2", "gn_pairs", "gn_iters"], list(cnt)))
        self.gn_iters = np.zeros(self.counts['gn_pairs'], np.int32)
        if len(self.gn_iters):
            L.so_get_gn_iters(C.c_void_p(h), _p(self.gn_iters))
        self.stages = {}
        if want_dumps:
            for k, name in enumerate(STAGES):
                tot = L.so_stage_total(C.c_void_p(h), k)
                if tot < 0:
                    continue
                off = np.zeros(nL + 1, np.int32)
                ridx = np.zeros(tot, np.int32)
                x, y, th, sc = (np.zeros(tot) for _ in range(4))
                L.so_get_stage(C.c_void_p(h), k, _p(off), _p(ridx), _p(x), _p(y), _p(th), _p(sc))
                self.stages[name] = dict(off=off, ridx=ridx, x=x, y=y, th=th, score=sc)
        L.so_free(C.c_void_p(h))


def stereo(Lraw, Rraw, Ledges, Redges, F21, Lund=None, Rund=None, descL=None, descR=None,
           want_dumps=True, threads=0) -> StereoResult:
    """Run the stereo restatement on one pair (no-GT branch)."""
    Lraw = np.ascontiguousarray(Lraw, dtype=np.uint8)
    Rraw = np.ascontiguousarray(Rraw, dtype=np.uint8)
    Lund = Lraw if Lund is None else np.ascontiguousarray(Lund, dtype=np.uint8)
    Rund = Rraw if Rund is None else np.ascontiguousarray(Rund, dtype=np.uint8)
    H, W = Lraw.shape
    Le = np.ascontiguousarray(Ledges, dtype=np.float64).reshape(-1, 3)
    Re = np.ascontiguousarray(Redges, dtype=np.float64).reshape(-1, 3)
    F = np.ascontiguousarray(F21, dtype=np.float64)
    mode = 0
    if descL is not None:
        mode = 1
        descL = np.ascontiguousarray(descL, dtype=np.float32)
        descR = np.ascontiguousarray(descR, dtype=np.float32)
    h = lib().so_run(_p(Lraw), _p(Rraw), _p(Lund), _p(Rund), H, W, _p(Le), len(Le), _p(Re), len(Re), _p(F),
                     mode, _p(descL), _p(descR), int(want_dumps), threads)
    return StereoResult(h, len(Le), want_dumps)


class ReferenceStereoResult:
    """Stage dumps + mates produced by the reference's own stereo code (oracle/ref_stereo_harness.cpp)."""

    def __init__(self, h, nL):
        R = stereo_ref_lib()
        hp = C.c_void_p(h)
        n = R.rs_num_mates(hp)
        self.mate_left = np.zeros(n, np.int32)
        rx, ry, rth = np.zeros(n), np.zeros(n), np.zeros(n)
        self.mate_score = np.zeros(n)
        R.rs_get_mates(hp, _p(self.mate_left), _p(rx), _p(ry), _p(rth), _p(self.mate_score))
        self.mate_right = np.stack([rx, ry, rth], 1) if n else np.zeros((0, 3))
        self.lines = np.zeros((nL, 3))
        self.F21 = np.zeros((3, 3))
        R.rs_get_lines(hp, _p(self.lines), _p(self.F21))
        self.stages = {}
        for k, name in enumerate(STAGES):
            tot = R.rs_stage_total(hp, k)
            if tot < 0:
                continue
            off = np.zeros(nL + 1, np.int32)
            ridx = np.zeros(tot, np.int32)
            x, y, th, sc = (np.zeros(tot) for _ in range(4))
            R.rs_get_stage(hp, k, _p(off), _p(ridx), _p(x), _p(y), _p(th), _p(sc))
            self.stages[name] = dict(off=off, ridx=ridx, x=x, y=y, th=th, score=sc)
        R.rs_free(hp)


def stereo_reference(Lraw, Rraw, Ledges, Redges, Kl, Kr, R21, T21, Lund=None, Rund=None) -> ReferenceStereoResult:
    """Run the reference's own stage functions (SIFT-off, see oracle/ref_stereo_harness.cpp) on one pair."""
    Lraw = np.ascontiguousarray(Lraw, dtype=np.uint8)
    Rraw = np.ascontiguousarray(Rraw, dtype=np.uint8)
    Lund = Lraw if Lund is None else np.ascontiguousarray(Lund, dtype=np.uint8)
    Rund = Rraw if Rund is None else np.ascontiguousarray(Rund, dtype=np.uint8)
    H, W = Lraw.shape
    Le = np.ascontiguousarray(Ledges, dtype=np.float64).reshape(-1, 3)
    Re = np.ascontiguousarray(Redges, dtype=np.float64).reshape(-1, 3)
    a = [np.ascontiguousarray(m, dtype=np.float64) for m in (Kl, Kr, R21, T21)]
    h = stereo_ref_lib().rs_run(_p(Lraw), _p(Rraw), _p(Lund), _p(Rund), H, W, _p(Le), len(Le), _p(Re), len(Re), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]))
    return ReferenceStereoResult(h, len(Le))


def ref_edge_patches(img, x, y, th):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    p, m = np.zeros(49, np.float32), np.zeros(49, np.float32)
    stereo_ref_lib().rs_edge_patches(_p(img), H, W, C.c_double(x), C.c_double(y), C.c_double(th), _p(p), _p(m))
    return p.reshape(7, 7), m.reshape(7, 7)


def ref_patch_similarity(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32).ravel()
    b = np.ascontiguousarray(b, dtype=np.float32).ravel()
    return stereo_ref_lib().rs_patch_similarity(_p(a), _p(b))


def ref_cluster(xyt, by_orientation=True):
    xyt = np.ascontiguousarray(xyt, dtype=np.float64).reshape(-1, 3)
    n = len(xyt)
    cen, cnt = np.zeros((max(n, 1), 3)), np.zeros(max(n, 1), np.int32)
    k = stereo_ref_lib().rs_cluster(_p(xyt), n, int(by_orientation), _p(cen), _p(cnt))
    return cen[:k].copy(), cnt[:k].copy()


# ---------------------------------------------------------------------------------------------------------------
# keyframe -> current-frame quad tracking (Temporal_Matches.cpp): restatement (temporal_oracle.inl) and the reference's
# own source compiled in place (ref_temporal_harness.cpp -> _ref/libtemporal_ref.so)
# ---------------------------------------------------------------------------------------------------------------
TQ_STAGES = ["grid", "orient", "ncc", "sift", "bnb", "bnb_sift", "gn", "cluster"]
_TREF = None


def have_temporal_ref() -> bool:
    return os.path.exists(os.path.join(_DIR, "_ref", "libtemporal_ref.so"))


def temporal_ref_lib():
    global _TREF
    if _TREF is None:
        build()
        R = C.CDLL(os.path.join(_DIR, "_ref", "libtemporal_ref.so"))
        R.rt_run.restype = C.c_void_p
        R.rt_stage_total.restype = C.c_int
        _TREF = R
    return _TREF


class QuadResult:
    """Per-stage ragged lists of candidate quads, one list per keyframe mate: off[n_kf + 1], cf (index of the
    current-frame mate), left / right (x, y, theta) of the cluster centres, ncc (left, right), score (left, right), valid."""

    def __init__(self, total, get, n_kf):
        self.stages = {}
        for k, name in enumerate(TQ_STAGES):
            tot = total(k)
            off, cf, valid = np.zeros(n_kf + 1, np.int32), np.zeros(tot, np.int32), np.zeros(tot, np.int32)
            l, r, ncc, sc, sift = np.zeros((tot, 3)), np.zeros((tot, 3)), np.zeros((tot, 2)), np.zeros((tot, 2)), np.zeros((tot, 2))
            get(k, _p(off), _p(cf), _p(l), _p(r), _p(ncc), _p(sc), _p(valid), _p(sift))
            self.stages[name] = dict(off=off, cf=cf, left=l, right=r, ncc=ncc, score=sc, valid=valid, sift=sift)
        self.seconds = 0.0
        self.counts = {}


def _tq_desc(desc, kf, cf):
    """desc = (kf_left, kf_right, cf_left, cf_right), each (n, 2, 128) float32 descriptor pairs, or None (SIFT-off)."""
    if desc is None:
        return [None] * 4
    out = [np.ascontiguousarray(d, dtype=np.float32).reshape(-1, 256) for d in desc]
    assert len(out[0]) == len(out[1]) == len(kf) and len(out[2]) == len(out[3]) == len(cf)
    return out


def _tq_args(kf_imgs, cf_imgs, kf, cf, kf_mask):
    imgs = [np.ascontiguousarray(a, dtype=np.uint8) for a in (*kf_imgs, *cf_imgs)]   # (L_raw, L_und, R_und) x 2
    H, W = imgs[0].shape
    kf = np.ascontiguousarray(kf, dtype=np.float64).reshape(-1, 6)
    cf = np.ascontiguousarray(cf, dtype=np.float64).reshape(-1, 6)
    mask = None if kf_mask is None else np.ascontiguousarray(kf_mask, dtype=np.uint8)
    return imgs, H, W, kf, cf, mask


def temporal(kf_imgs, cf_imgs, kf, cf, kf_mask=None, cell=15, radius=30.0, orient_deg=10.0, ncc_thresh=0.8, bnb=0.8,
             desc=None, sift_thresh=200.0) -> QuadResult:
    """Run the quad-tracking restatement.  kf_imgs / cf_imgs = (L_raw, L_und, R_und); kf / cf = n x 6 (left xyt, right xyt);
    desc = optional descriptor pairs (see _tq_desc): with them the SIFT gate and the SIFT best-nearly-best pass run."""
    imgs, H, W, kf, cf, mask = _tq_args(kf_imgs, cf_imgs, kf, cf, kf_mask)
    dd = _tq_desc(desc, kf, cf)
    L = lib()
    L.to_run.restype = C.c_void_p
    L.to_stage_total.restype = C.c_int
    h = C.c_void_p(L.to_run(*[_p(a) for a in imgs], H, W, _p(kf), len(kf), _p(mask), _p(cf), len(cf), cell, C.c_double(radius),
                            C.c_double(orient_deg), C.c_double(ncc_thresh), C.c_double(bnb), *[_p(a) for a in dd], C.c_double(sift_thresh)))
    res = QuadResult(lambda k: L.to_stage_total(h, k), lambda k, *a: L.to_get_stage(h, k, *a), len(kf))
    sec, cnt = C.c_double(), (C.c_long * 2)()
    L.to_get_stats(h, C.byref(sec), cnt)
    res.seconds, res.counts = sec.value, dict(gn_pairs=cnt[0], gn_iters=cnt[1])
    L.to_free(h)
    return res


def temporal_reference(kf_imgs, cf_imgs, kf, cf, kf_mask=None, desc=None) -> QuadResult:
    """Run the reference's own quad stages (oracle/ref_temporal_harness.cpp; thresholds of Temporal_Matches.cpp:185-213)."""
    imgs, H, W, kf, cf, mask = _tq_args(kf_imgs, cf_imgs, kf, cf, kf_mask)
    dd = _tq_desc(desc, kf, cf)
    R = temporal_ref_lib()
    h = C.c_void_p(R.rt_run(*[_p(a) for a in imgs], H, W, _p(kf), len(kf), _p(mask), _p(cf), len(cf), *[_p(a) for a in dd]))
    res = QuadResult(lambda k: R.rt_stage_total(h, k), lambda k, *a: R.rt_get_stage(h, k, *a), len(kf))
    R.rt_free(h)
    return res

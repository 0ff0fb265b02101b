"""ctypes binding of libebvo_b200.so (C ABI declared in include/ebvo_b200.h).

There is no CPU fallback: if the library is missing the import fails loudly, and without
a CUDA device ``Context()`` raises ``EbvoError`` (EBVO_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libebvo_b200.so")

STAGES = ["epi", "disp", "orient", "sift", "ncc", "bnb_ncc", "bnb_sift", "shift", "gn", "cluster", "ncc2", "best"]

EXPORTS = [
    "ebvo_params_default", "ebvo_create", "ebvo_destroy", "ebvo_last_error", "ebvo_fundamental", "ebvo_toed",
    "ebvo_stereo_match", "ebvo_stereo_match_full", "ebvo_stereo_frame", "ebvo_stereo_batch", "ebvo_stereo_batch_packed", "ebvo_stereo_batch_multi", "ebvo_batch_upload", "ebvo_batch_run",
    "ebvo_batch_sync", "ebvo_batch_download", "ebvo_batch_counts", "ebvo_batch_pack", "ebvo_edge_patches", "ebvo_ncc_patch_pair",
    "ebvo_cluster", "ebvo_sobel", "ebvo_sift_descriptors", "ebvo_undistort", "ebvo_launch_count", "ebvo_set_stage_dumps", "ebvo_stage_size", "ebvo_stage_fetch",
    "ebvo_set_profiling", "ebvo_get_kernel_times", "ebvo_stream", "ebvo_host_alloc", "ebvo_host_free",
    "ebvo_temporal_quads", "ebvo_temporal_quads_stage", "ebvo_temporal_counters",
]


class EbvoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ebvo error {code}: {msg}")
        self.code = code


class Edge(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("theta", C.c_double), ("index", C.c_int32), ("frame_source", C.c_int32)]


class Calib(C.Structure):
    _fields_ = [("Kl", C.c_double * 9), ("Kr", C.c_double * 9), ("R21", C.c_double * 9), ("T21", C.c_double * 3)]


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "epipolar_line_dist_thresh", "max_disparity", "orientation_thresh_deg", "orthogonal_shift_mag", "ncc_thresh",
        "bnb_ncc", "bnb_sift", "sift_threshold", "location_perturbation", "epip_tangency_displ_thresh",
        "orient_perturbation", "cluster_dist_thresh", "cluster_orient_thresh_deg", "cluster_orient_gauss_sigma")] + [
        ("max_cluster_size", C.c_int32), ("gn_max_iter", C.c_int32), ("gn_tol", C.c_double), ("gn_huber_delta", C.c_double),
        ("toed_mag_thresh", C.c_double), ("toed_border", C.c_int32), ("gn_mode", C.c_int32), ("sift_mode", C.c_int32)]


class Mate(C.Structure):
    _fields_ = [("left_index", C.c_int32), ("reserved", C.c_int32), ("lx", C.c_double), ("ly", C.c_double), ("ltheta", C.c_double),
                ("rx", C.c_double), ("ry", C.c_double), ("rtheta", C.c_double), ("score", C.c_double)]


EDGE_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("theta", "<f8"), ("index", "<i4"), ("frame_source", "<i4")])
MATE_DTYPE = np.dtype([("left_index", "<i4"), ("reserved", "<i4"), ("lx", "<f8"), ("ly", "<f8"), ("ltheta", "<f8"),
                       ("rx", "<f8"), ("ry", "<f8"), ("rtheta", "<f8"), ("score", "<f8")])
assert EDGE_DTYPE.itemsize == C.sizeof(Edge) == 32
assert MATE_DTYPE.itemsize == C.sizeof(Mate) == 64
# ebvo_quad (include/ebvo_b200.h): one surviving quad of the keyframe -> current-frame tracking
QUAD_DTYPE = np.dtype([("kf_index", "<i4"), ("cf_index", "<i4"), ("lx", "<f8"), ("ly", "<f8"), ("ltheta", "<f8"),
                       ("rx", "<f8"), ("ry", "<f8"), ("rtheta", "<f8"), ("ncc_left", "<f8"), ("ncc_right", "<f8"),
                       ("sift_left", "<f8"), ("sift_right", "<f8"), ("score_left", "<f8"), ("score_right", "<f8"),
                       ("valid", "<i4"), ("reserved", "<i4")])
assert QUAD_DTYPE.itemsize == 112
TQ_STAGES = ["grid", "orient", "ncc", "sift", "bnb", "bnb_sift", "gn", "cluster"]


class QuadParams(C.Structure):
    """ebvo_quad_params: knobs of Temporal_Matches::get_Temporal_Edge_Pairs_from_Quads (reference values)."""
    _fields_ = [("cell_size", C.c_int32), ("reserved", C.c_int32), ("grid_radius", C.c_double), ("orient_deg", C.c_double),
                ("ncc_thresh", C.c_double), ("bnb_thresh", C.c_double), ("sift_thresh", C.c_double)]


def mates_from_arrays(left_xyt, right_xyt) -> np.ndarray:
    """ebvo_mate records from (n, 3) left and right edges."""
    l = np.asarray(left_xyt, np.float64).reshape(-1, 3); r = np.asarray(right_xyt, np.float64).reshape(-1, 3)
    m = np.zeros(len(l), MATE_DTYPE)
    m["left_index"] = np.arange(len(l))
    m["lx"], m["ly"], m["ltheta"], m["rx"], m["ry"], m["rtheta"] = l[:, 0], l[:, 1], l[:, 2], r[:, 0], r[:, 1], r[:, 2]
    return m

_lib = None


def load():
    """Load the shared library (raises if it has not been built: python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                              "(no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.ebvo_last_error.restype = C.c_char_p
        L.ebvo_last_error.argtypes = [C.c_void_p]
        L.ebvo_host_alloc.restype = C.c_void_p
        L.ebvo_host_alloc.argtypes = [C.c_size_t]
        L.ebvo_host_free.restype = None
        L.ebvo_host_free.argtypes = [C.c_void_p]
        L.ebvo_stream.restype = C.c_void_p
        L.ebvo_stream.argtypes = [C.c_void_p]
        L.ebvo_destroy.argtypes = [C.c_void_p]
        L.ebvo_destroy.restype = None
        L.ebvo_launch_count.restype = C.c_longlong
        L.ebvo_launch_count.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_calib(Kl, Kr, R21, T21) -> Calib:
    c = Calib()
    c.Kl[:] = np.asarray(Kl, np.float64).ravel().tolist()
    c.Kr[:] = np.asarray(Kr, np.float64).ravel().tolist()
    c.R21[:] = np.asarray(R21, np.float64).ravel().tolist()
    c.T21[:] = np.asarray(T21, np.float64).ravel().tolist()
    return c


def edges_from_xyt(xyt) -> np.ndarray:
    xyt = np.asarray(xyt, np.float64).reshape(-1, 3)
    e = np.zeros(len(xyt), EDGE_DTYPE)
    e["x"], e["y"], e["theta"] = xyt[:, 0], xyt[:, 1], xyt[:, 2]
    e["index"] = np.arange(len(xyt))
    e["frame_source"] = -1
    return e


def default_params() -> Params:
    p = Params()
    load().ebvo_params_default(C.byref(p))
    return p


def fundamental(calib: Calib):
    F21, F12 = np.zeros((3, 3)), np.zeros((3, 3))
    load().ebvo_fundamental(C.byref(calib), _p(F21), _p(F12))
    return F21, F12


def stereo_batch_multi(contexts, calib, L_imgs, R_imgs, cap):
    """ebvo_stereo_batch_multi: one batch over several contexts (one per GPU), contiguous frame blocks, host threads."""
    F = len(L_imgs)
    h, w = L_imgs[0].shape
    out = np.zeros((F, cap), MATE_DTYPE)
    n_mates = np.zeros(F, np.int32)
    handles = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    c0 = contexts[0]
    rc = c0.L.ebvo_stereo_batch_multi(handles, len(contexts), C.byref(calib), F, c0._ptr_array(L_imgs), c0._ptr_array(R_imgs), w, h,
                                      L_imgs[0].strides[0], _p(out), cap, _p(n_mates))
    if rc != 0:
        msgs = [c.L.ebvo_last_error(c.h).decode() for c in contexts]
        raise RuntimeError(f"ebvo_stereo_batch_multi failed ({rc}): {msgs}")
    return out, n_mates


class Context:
    """One GPU context (ebvo_create / ebvo_destroy)."""

    def __init__(self, device=0, max_w=1241, max_h=376, max_batch=1, max_edges=65536, params: Params | None = None):
        self.L = load()
        self.h = C.c_void_p()
        self.max_edges = max_edges
        self.max_batch = max_batch
        rc = self.L.ebvo_create(C.byref(self.h), device, max_w, max_h, max_batch, max_edges,
                                C.byref(params) if params is not None else None)
        if rc != 0:
            msg = self.L.ebvo_last_error(self.h).decode() if self.h else "no CUDA device (there is no CPU fallback)"
            if self.h:
                self.L.ebvo_destroy(self.h)
                self.h = C.c_void_p()
            raise EbvoError(rc, msg)

    def close(self):
        if self.h:
            self.L.ebvo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise EbvoError(rc, self.L.ebvo_last_error(self.h).decode())

    @property
    def stream(self):
        return self.L.ebvo_stream(self.h)

    # ---- TOED ----
    def toed(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.zeros(self.max_edges, EDGE_DTYPE)
        n, nt = C.c_int(), C.c_int()
        self._ck(self.L.ebvo_toed(self.h, _p(img), w, h, img.strides[0], _p(out), len(out), C.byref(n), C.byref(nt)))
        return out[:n.value].copy(), nt.value

    # ---- stereo ----
    def stereo_match(self, calib, L_raw, R_raw, L_edges, R_edges, L_und=None, R_und=None, descL=None, descR=None):
        L_raw = np.ascontiguousarray(L_raw, np.uint8)
        R_raw = np.ascontiguousarray(R_raw, np.uint8)
        h, w = L_raw.shape
        L_und = None if L_und is None else np.ascontiguousarray(L_und, np.uint8)
        R_und = None if R_und is None else np.ascontiguousarray(R_und, np.uint8)
        Le = np.ascontiguousarray(L_edges, EDGE_DTYPE)
        Re = np.ascontiguousarray(R_edges, EDGE_DTYPE)
        if descL is not None:
            descL = np.ascontiguousarray(descL, np.float32)
            descR = np.ascontiguousarray(descR, np.float32)
        out = np.zeros(max(len(Le), 1), MATE_DTYPE)
        n = C.c_int()
        self._ck(self.L.ebvo_stereo_match(self.h, C.byref(calib), _p(L_raw), _p(R_raw), _p(L_und), _p(R_und), w, h, L_raw.strides[0],
                                          _p(Le), len(Le), _p(Re), len(Re), _p(descL), _p(descR), _p(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def stereo_frame(self, calib, L_img, R_img, want_edges=True):
        L_img = np.ascontiguousarray(L_img, np.uint8)
        R_img = np.ascontiguousarray(R_img, np.uint8)
        h, w = L_img.shape
        out = np.zeros(self.max_edges, MATE_DTYPE)
        n, nL, nR = C.c_int(), C.c_int(), C.c_int()
        Le = np.zeros(self.max_edges, EDGE_DTYPE) if want_edges else None
        Re = np.zeros(self.max_edges, EDGE_DTYPE) if want_edges else None
        self._ck(self.L.ebvo_stereo_frame(self.h, C.byref(calib), _p(L_img), _p(R_img), w, h, L_img.strides[0], _p(out), len(out),
                                          C.byref(n), _p(Le), C.byref(nL), _p(Re), C.byref(nR), self.max_edges))
        if want_edges:
            return out[:n.value].copy(), Le[:nL.value].copy(), Re[:nR.value].copy()
        return out[:n.value].copy()

    def _ptr_array(self, imgs):
        arr = (C.c_void_p * len(imgs))()
        for k, im in enumerate(imgs):
            arr[k] = im.__array_interface__["data"][0]      # (im.ctypes.data builds a ctypes object per image: 1 ms per 320 images)
        return arr

    def stereo_batch(self, calib, L_imgs, R_imgs, cap, out=None, n_mates=None):
        """Host buffers in, host buffers out.  L_imgs/R_imgs: lists of HxW uint8 arrays."""
        F = len(L_imgs)
        h, w = L_imgs[0].shape
        if out is None:
            out = np.zeros((F, cap), MATE_DTYPE)
        if n_mates is None:
            n_mates = np.zeros(F, np.int32)
        self._ck(self.L.ebvo_stereo_batch(self.h, C.byref(calib), F, self._ptr_array(L_imgs), self._ptr_array(R_imgs), w, h,
                                          L_imgs[0].strides[0], _p(out), cap, _p(n_mates)))
        return out, n_mates

    def stereo_batch_device(self, calib, L_imgs, R_imgs, n_mates=None):
        """ebvo_stereo_batch with out = NULL: host images in (pipelined H2D), results stay on the device for batch_pack()."""
        F = len(L_imgs)
        h, w = L_imgs[0].shape
        if n_mates is None:
            n_mates = np.zeros(F, np.int32)
        self._ck(self.L.ebvo_stereo_batch(self.h, C.byref(calib), F, self._ptr_array(L_imgs), self._ptr_array(R_imgs), w, h,
                                          L_imgs[0].strides[0], None, 0x7fffffff, _p(n_mates)))
        self._nframes = F
        return n_mates

    def stereo_batch_packed(self, calib, L_imgs, R_imgs, out_ptr: int, cap_records: int, n_mates=None):
        """ebvo_stereo_batch_packed: host images in, the batch's mates back to back at the HOST address out_ptr (streamed out
        sub-batch by sub-batch while the rest computes).  Returns (n_mates, records written)."""
        F = len(L_imgs)
        h, w = L_imgs[0].shape
        if n_mates is None:
            n_mates = np.zeros(F, np.int32)
        tot = C.c_longlong()
        self._ck(self.L.ebvo_stereo_batch_packed(self.h, C.byref(calib), F, self._ptr_array(L_imgs), self._ptr_array(R_imgs), w, h,
                                                 L_imgs[0].strides[0], C.c_void_p(out_ptr), C.c_longlong(cap_records), _p(n_mates), C.byref(tot)))
        self._nframes = F
        return n_mates, tot.value

    def batch_pack(self, dst_ptr: int, cap_records: int, offsets_ptr: int) -> int:
        """Pack the last batch's mates (device-resident) back to back into a caller-owned DEVICE buffer; returns the record count."""
        tot = C.c_longlong()
        self._ck(self.L.ebvo_batch_pack(self.h, C.c_void_p(dst_ptr), C.c_longlong(cap_records), C.c_void_p(offsets_ptr), C.byref(tot)))
        return tot.value

    def batch_upload(self, L_imgs, R_imgs):
        h, w = L_imgs[0].shape
        self._ck(self.L.ebvo_batch_upload(self.h, len(L_imgs), self._ptr_array(L_imgs), self._ptr_array(R_imgs), w, h, L_imgs[0].strides[0]))
        self._nframes = len(L_imgs)

    def batch_run(self, calib, do_match=True):
        self._ck(self.L.ebvo_batch_run(self.h, C.byref(calib), int(do_match)))

    def batch_sync(self):
        self._ck(self.L.ebvo_batch_sync(self.h))

    def batch_download(self, cap):
        F = self._nframes
        out = np.zeros((F, cap), MATE_DTYPE)
        n = np.zeros(F, np.int32)
        self._ck(self.L.ebvo_batch_download(self.h, _p(out), cap, _p(n)))
        return out, n

    def batch_counts(self):
        F = self._nframes
        nL, nR, nM = (np.zeros(F, np.int32) for _ in range(3))
        cnt = np.zeros((F, 8), np.int64)
        self._ck(self.L.ebvo_batch_counts(self.h, _p(nL), _p(nR), _p(nM), _p(cnt)))
        return nL, nR, nM, cnt

    # ---- helpers ----
    def edge_patches(self, img, edges):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        e = np.ascontiguousarray(edges, EDGE_DTYPE)
        p = np.zeros((len(e), 7, 7), np.float32)
        m = np.zeros((len(e), 7, 7), np.float32)
        self._ck(self.L.ebvo_edge_patches(self.h, _p(img), w, h, img.strides[0], _p(e), len(e), _p(p), _p(m)))
        return p, m

    def ncc(self, p1, p2):
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 49)
        p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 49)
        out = np.zeros(len(p1))
        self._ck(self.L.ebvo_ncc_patch_pair(self.h, _p(p1), _p(p2), len(p1), _p(out)))
        return out

    def cluster(self, edges, by_orientation=True):
        e = np.ascontiguousarray(edges, EDGE_DTYPE)
        cen = np.zeros(max(len(e), 1), EDGE_DTYPE)
        lab = np.zeros(max(len(e), 1), np.int32)
        n = C.c_int()
        self._ck(self.L.ebvo_cluster(self.h, _p(e), len(e), int(by_orientation), _p(cen), _p(lab), C.byref(n)))
        return cen[:n.value].copy(), lab[:len(e)].copy()

    def temporal_quads(self, kf_imgs, cf_imgs, kf, cf, kf_mask=None, stage="cluster", params=None, cap=None, desc=None):
        """Keyframe -> current-frame quad tracking (ebvo_temporal_quads_stage).  kf_imgs / cf_imgs = (L_raw, L_und, R_und)
        uint8 images; kf / cf = MATE_DTYPE arrays; desc = optional (kf_left, kf_right, cf_left, cf_right) descriptor pairs,
        each (n, 2, 128) float32 (SIFT-on).  Returns (off[n_kf + 1], quads) after `stage`."""
        dd = [None] * 4 if desc is None else [None if a is None else np.ascontiguousarray(a, np.float32).reshape(-1, 256) for a in desc]
        imgs = [np.ascontiguousarray(a, np.uint8) for a in (*kf_imgs, *cf_imgs)]
        h, w = imgs[0].shape
        kf = np.ascontiguousarray(kf, MATE_DTYPE); cf = np.ascontiguousarray(cf, MATE_DTYPE)
        mask = None if kf_mask is None else np.ascontiguousarray(kf_mask, np.uint8)
        k = TQ_STAGES.index(stage) if isinstance(stage, str) else int(stage)
        full = max(1, len(kf) * 128 if k >= 2 else min(len(kf) * max(len(cf), 1), 64 * 1024 * 1024))
        caps = [cap] if cap is not None else ([min(full, max(16 * len(kf), 1 << 16)), full] if k >= 2 else [full])
        off = np.zeros(len(kf) + 1, np.int32)
        n = C.c_int()
        qp = C.byref(params) if params is not None else None
        for c in caps:       # a modest buffer first, the worst case (128 quads per keyframe mate) only when it is needed
            out = np.empty(c, QUAD_DTYPE)
            rc = self.L.ebvo_temporal_quads_stage(self.h, *[_p(a) for a in imgs], w, h, imgs[0].strides[0], _p(kf), len(kf), _p(mask),
                                                  _p(cf), len(cf), *[_p(a) for a in dd], qp, k, _p(off), _p(out), c, C.byref(n))
            if rc != -4 or c == caps[-1]:
                break
        self._ck(rc)
        return off, out[:n.value].copy()

    def temporal_counters(self):
        c = (C.c_longlong * 8)()
        self._ck(self.L.ebvo_temporal_counters(self.h, c))
        return dict(gate_survivors=c[0], gn_problems=c[1], gn_iterations=c[2], grid_candidates=c[3], orient_survivors=c[4])

    def sobel(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        gx, gy = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
        self._ck(self.L.ebvo_sobel(self.h, _p(img), w, h, img.strides[0], _p(gx), _p(gy)))
        return gx, gy

    def undistort(self, img, K, dist):
        """cv::undistort(img, K, dist) for one 8-bit image (dist = k1, k2, p1, p2)."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        K = np.ascontiguousarray(K, np.float64).ravel(); dist = np.ascontiguousarray(dist, np.float64).ravel()
        out = np.zeros((h, w), np.uint8)
        self._ck(self.L.ebvo_undistort(self.h, _p(img), w, h, img.strides[0], _p(K), _p(dist), _p(out), out.strides[0]))
        return out

    def sift_descriptors(self, img, edges):
        """augment_Edge_Data: (n, 2, 128) float32 descriptors (needs a context created with params.sift_mode = 1)."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        e = np.ascontiguousarray(edges, EDGE_DTYPE)
        out = np.zeros((len(e), 2, 128), np.float32)
        self._ck(self.L.ebvo_sift_descriptors(self.h, _p(img), w, h, img.strides[0], _p(e), len(e), _p(out)))
        return out

    def launch_count(self):
        return int(self.L.ebvo_launch_count(self.h))

    def set_stage_dumps(self, enable=True):
        self._ck(self.L.ebvo_set_stage_dumps(self.h, int(enable)))

    def stage(self, name):
        k = STAGES.index(name)
        nl, tot = C.c_int(), C.c_int()
        self._ck(self.L.ebvo_stage_size(self.h, k, C.byref(nl), C.byref(tot)))
        off = np.zeros(nl.value + 1, np.int32)
        ridx = np.zeros(tot.value, np.int32)
        x, y, th, sc = (np.zeros(tot.value) for _ in range(4))
        self._ck(self.L.ebvo_stage_fetch(self.h, k, _p(off), _p(ridx), _p(x), _p(y), _p(th), _p(sc)))
        return dict(off=off, ridx=ridx, x=x, y=y, th=th, score=sc)

    def set_profiling(self, enable=True):
        self._ck(self.L.ebvo_set_profiling(self.h, int(enable)))

    def kernel_times(self):
        names = C.POINTER(C.c_char_p)()
        ms = C.POINTER(C.c_float)()
        ln = C.POINTER(C.c_int)()
        n = C.c_int()
        self._ck(self.L.ebvo_get_kernel_times(self.h, C.byref(names), C.byref(ms), C.byref(ln), C.byref(n)))
        return {names[k].decode(): (ms[k], ln[k]) for k in range(n.value)}

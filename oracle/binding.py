"""ctypes binding of the CPU oracle (oracle/liboracle_pcl.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.  PARITY UNPINNED: see
oracle/pcl_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_pcl.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _opt(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed with {rc}")


# ------------------------------------------------------------------ search
def radius_search(surf, q, radius, brute=False):
    surf, q = _f32(surf), _f32(q)
    n, nq = len(surf), len(q)
    counts = np.zeros(nq, np.int32)
    L = lib()
    _chk(L.orc_radius_count(_opt(surf), n, _opt(q), nq, C.c_double(radius), _opt(counts)), "radius_count")
    offsets = np.zeros(nq + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    idx = np.zeros(int(offsets[-1]), np.int32)
    d2 = np.zeros(int(offsets[-1]), np.float32)
    fn = L.orc_radius_search_brute if brute else L.orc_radius_search
    _chk(fn(_opt(surf), n, _opt(q), nq, C.c_double(radius), _opt(offsets), _opt(idx), _opt(d2)), "radius_search")
    return offsets, idx, d2


def radius_count(surf, q, radius):
    surf, q = _f32(surf), _f32(q)
    counts = np.zeros(len(q), np.int32)
    _chk(lib().orc_radius_count(_opt(surf), len(surf), _opt(q), len(q), C.c_double(radius), _opt(counts)), "radius_count")
    return counts.astype(np.int64)


def knn(surf, q, k, brute=False):
    surf, q = _f32(surf), _f32(q)
    idx = np.zeros((len(q), k), np.int32)
    d2 = np.zeros((len(q), k), np.float32)
    fn = lib().orc_knn_brute if brute else lib().orc_knn
    _chk(fn(_opt(surf), len(surf), _opt(q), len(q), k, _opt(idx), _opt(d2)), "knn")
    return idx, d2


def cloud_resolution(pts):
    pts = _f32(pts)
    r = C.c_double(0)
    _chk(lib().orc_cloud_resolution(_opt(pts), len(pts), C.byref(r)), "cloud_resolution")
    return r.value


# ------------------------------------------------------------------ normals
def normals(surf, q=None, radius=0.0, k=0, vp=(0, 0, 0), mode=0):
    surf = _f32(surf)
    q = surf if q is None else _f32(q)
    out = np.zeros((len(q), 4), np.float32)
    cnt = np.zeros(len(q), np.int32)
    gap = np.zeros(len(q), np.float32)
    vpa = (C.c_float * 3)(*vp)
    _chk(lib().orc_normals(_opt(surf), len(surf), _opt(q), len(q), C.c_double(radius), int(k), vpa,
                           int(mode), _opt(out), _opt(cnt), _opt(gap)), "normals")
    return out, cnt, gap


# ------------------------------------------------------------------ keypoints
def iss_saliency(pts, salient_radius, min_neighbors=5, g21=0.975, g32=0.975):
    pts = _f32(pts)
    sal = np.zeros(len(pts), np.float64)
    _chk(lib().orc_iss_saliency(_opt(pts), len(pts), C.c_double(salient_radius), min_neighbors,
                                C.c_double(g21), C.c_double(g32), _opt(sal)), "iss_saliency")
    return sal


def iss_nms(pts, saliency, nonmax_radius, min_neighbors=5):
    pts = _f32(pts)
    saliency = np.ascontiguousarray(saliency, np.float64)
    kp = np.zeros(len(pts), np.int32)
    nk = C.c_int(0)
    _chk(lib().orc_iss_nms(_opt(pts), len(pts), _opt(saliency), C.c_double(nonmax_radius), min_neighbors,
                           _opt(kp), C.byref(nk)), "iss_nms")
    return kp[: nk.value].copy()


def iss(pts, salient_radius, nonmax_radius, min_neighbors=5, g21=0.975, g32=0.975):
    sal = iss_saliency(pts, salient_radius, min_neighbors, g21, g32)
    return iss_nms(pts, sal, nonmax_radius, min_neighbors), sal


def harris_response(pts, normals4, radius):
    pts, normals4 = _f32(pts), _f32(normals4)
    r = np.zeros(len(pts), np.float32)
    _chk(lib().orc_harris_response(_opt(pts), _opt(normals4), len(pts), C.c_double(radius), _opt(r)), "harris_response")
    return r


def harris6d_response(pts, rgb, normals4, radius):
    """-> (response [n], gradients [n, 3] after the length rule, intensity [n]); rgb: uint32 0x00RRGGBB per point"""
    pts = _f32(pts)
    rgb = np.ascontiguousarray(rgb, np.uint32)
    nr = np.ascontiguousarray(normals4, np.float32)
    assert len(rgb) == len(pts) and nr.shape == (len(pts), 4)
    resp = np.zeros(len(pts), np.float32)
    grad = np.zeros((len(pts), 3), np.float32)
    inten = np.zeros(len(pts), np.float32)
    _chk(lib().orc_harris6d_response(_opt(pts), _opt(rgb), _opt(nr), len(pts), C.c_double(radius), _opt(resp), _opt(grad),
                                     _opt(inten)), "harris6d_response")
    return resp, grad, inten


def harris_nms(pts, response, radius, threshold):
    pts, response = _f32(pts), _f32(response)
    kp = np.zeros(len(pts), np.int32)
    nk = C.c_int(0)
    _chk(lib().orc_harris_nms(_opt(pts), _opt(response), len(pts), C.c_double(radius), C.c_float(threshold),
                              _opt(kp), C.byref(nk)), "harris_nms")
    return kp[: nk.value].copy()


def harris_refine(pts, normals4, radius, corners):
    pts, normals4 = _f32(pts), _f32(normals4)
    c = _f32(corners).copy()
    _chk(lib().orc_harris_refine(_opt(pts), _opt(normals4), len(pts), C.c_double(radius), _opt(c), len(c)), "harris_refine")
    return c


def snap_to_cloud(pts, q, max_d2=1e-4):
    pts, q = _f32(pts), _f32(q)
    out = np.zeros(len(q), np.int32)
    _chk(lib().orc_snap_to_cloud(_opt(pts), len(pts), _opt(q), len(q), C.c_float(max_d2), _opt(out)), "snap")
    return out


# ------------------------------------------------------------------ descriptors
def spfh(surf, normals4, pidx, radius=0.0, k=0):
    surf, normals4 = _f32(surf), _f32(normals4)
    pidx = np.ascontiguousarray(pidx, np.int32)
    out = np.zeros((len(pidx), 33), np.float32)
    _chk(lib().orc_spfh(_opt(surf), _opt(normals4), len(surf), _opt(pidx), len(pidx), C.c_double(radius), int(k),
                        _opt(out)), "spfh")
    return out


def fpfh(surf, normals4, q=None, radius=0.0, k=0):
    surf, normals4 = _f32(surf), _f32(normals4)
    q = surf if q is None else _f32(q)
    out = np.zeros((len(q), 33), np.float32)
    _chk(lib().orc_fpfh(_opt(surf), _opt(normals4), len(surf), _opt(q), len(q), C.c_double(radius), int(k),
                        _opt(out)), "fpfh")
    return out


def shot_lrf(surf, q, radius):
    surf = _f32(surf)
    q = surf if q is None else _f32(q)
    rf = np.zeros((len(q), 9), np.float32)
    gap = np.zeros((len(q), 2), np.float32)
    _chk(lib().orc_shot_lrf(_opt(surf), len(surf), _opt(q), len(q), C.c_double(radius), _opt(rf), _opt(gap)), "shot_lrf")
    return rf, gap


def shot352(surf, normals4, q, radius, lrf_in=None):
    surf, normals4 = _f32(surf), _f32(normals4)
    q = surf if q is None else _f32(q)
    out = np.zeros((len(q), 352), np.float32)
    rf = np.zeros((len(q), 9), np.float32)
    lrf = _f32(lrf_in) if lrf_in is not None else None
    _chk(lib().orc_shot352(_opt(surf), _opt(normals4), len(surf), _opt(q), len(q), C.c_double(radius), _opt(lrf),
                           _opt(out), _opt(rf)), "shot352")
    return out, rf


# ------------------------------------------------------------------ matching / ingest
def match_nn(a, b):
    a, b = _f32(a), _f32(b)
    idx = np.zeros(len(a), np.int32)
    d2 = np.zeros(len(a), np.float32)
    dim = a.shape[1] if a.ndim == 2 else b.shape[1]
    _chk(lib().orc_match_nn(_opt(a), len(a), _opt(b), len(b), dim, _opt(idx), _opt(d2)), "match_nn")
    return idx, d2


def match_reciprocal(a, b):
    a, b = _f32(a), _f32(b)
    qi = np.zeros(len(a), np.int32)
    mi = np.zeros(len(a), np.int32)
    dist = np.zeros(len(a), np.float32)
    m = C.c_int(0)
    _chk(lib().orc_match_reciprocal(_opt(a), len(a), _opt(b), len(b), a.shape[1], _opt(qi), _opt(mi), _opt(dist),
                                    C.byref(m)), "match_reciprocal")
    return qi[: m.value].copy(), mi[: m.value].copy(), dist[: m.value].copy()


def voxel_grid(pts, leaf):
    pts = _f32(pts)
    out = np.zeros((len(pts), 3), np.float32)
    m = C.c_int(0)
    _chk(lib().orc_voxel_grid(_opt(pts), len(pts), C.c_float(leaf), _opt(out), len(pts), C.byref(m)), "voxel_grid")
    return out[: m.value].copy()


# ------------------------------------------------------------------ range image / NARF
class RiDesc(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("planar", C.c_int), ("cx", C.c_float), ("cy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("ang_res", C.c_float), ("off_x", C.c_int), ("off_y", C.c_int)]


def range_image_planar(pts, width, height, cx, cy, fx, fy, min_range=0.0):
    """-> (img [h, w, 4] = x, y, z, range; RiDesc)"""
    pts = _f32(pts)
    img = np.zeros((height, width, 4), np.float32)
    _chk(lib().orc_range_image_planar(_opt(pts), len(pts), width, height, C.c_float(cx), C.c_float(cy), C.c_float(fx),
                                      C.c_float(fy), C.c_float(min_range), _opt(img)), "range_image_planar")
    return img, RiDesc(width, height, 1, cx, cy, fx, fy, 0.0, 0, 0)


def range_image_spherical(pts, ang_res, max_angle_w=2 * np.pi, max_angle_h=np.pi, min_range=0.0, border=0):
    pts = _f32(pts)
    cap = int(np.floor(max_angle_w / ang_res)) * int(np.floor(max_angle_h / ang_res)) + 16
    buf = np.zeros(cap * 4, np.float32)
    w, h, ox, oy = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    _chk(lib().orc_range_image_spherical(_opt(pts), len(pts), C.c_float(ang_res), C.c_float(max_angle_w),
                                         C.c_float(max_angle_h), C.c_float(min_range), int(border), _opt(buf), cap,
                                         C.byref(w), C.byref(h), C.byref(ox), C.byref(oy)), "range_image_spherical")
    img = buf[: w.value * h.value * 4].reshape(h.value, w.value, 4).copy()
    return img, RiDesc(w.value, h.value, 0, 0.0, 0.0, 1.0, 1.0, ang_res, ox.value, oy.value)


def narf_borders(img, desc):
    img = _f32(img)
    n = desc.width * desc.height
    traits = np.zeros(n, np.int32)
    scores = np.zeros((4, n), np.float32)
    sc = np.zeros(n, np.float32)
    sd = np.zeros((n, 3), np.float32)
    _chk(lib().orc_narf_borders(_opt(img), C.byref(desc), _opt(traits), _opt(scores), _opt(sc), _opt(sd)), "narf_borders")
    return traits, scores, sc, sd


def narf_keypoints(img, desc, support_size):
    img = _f32(img)
    n = desc.width * desc.height
    kp = np.zeros(n, np.int32)
    val = np.zeros(n, np.float32)
    interest = np.zeros(n, np.float32)
    m = C.c_int(0)
    _chk(lib().orc_narf_keypoints(_opt(img), C.byref(desc), C.c_float(support_size), _opt(kp), _opt(val), n, C.byref(m),
                                  _opt(interest)), "narf_keypoints")
    return kp[: m.value].copy(), val[: m.value].copy(), interest


def narf36(img, desc, kp_px, support_size, rotation_invariant=True):
    img = _f32(img)
    kp_px = np.ascontiguousarray(kp_px, np.int32)
    cap = max(1, 8 * len(kp_px))
    out = np.zeros((cap, 42), np.float32)
    m = C.c_int(0)
    _chk(lib().orc_narf36(_opt(img), C.byref(desc), _opt(kp_px), len(kp_px), C.c_float(support_size),
                          int(bool(rotation_invariant)), _opt(out), cap, C.byref(m)), "narf36")
    return out[: m.value].copy()


# ------------------------------------------------------------------ RANSAC rejection
def ransac_reject(src, tgt, corr_q, corr_m, threshold=0.015, max_iterations=1000, seed=12345):
    """-> (keep flags [n_corr], T [4, 4], iterations, best hypothesis index)"""
    src, tgt = _f32(src), _f32(tgt)
    q = np.ascontiguousarray(corr_q, np.int32)
    m = np.ascontiguousarray(corr_m, np.int32)
    keep = np.zeros(max(len(q), 1), np.int32)
    T = np.zeros(16, np.float32)
    ninl, it, bh = C.c_int(0), C.c_int(0), C.c_int(0)
    _chk(lib().orc_ransac_reject(_opt(src), len(src), _opt(tgt), len(tgt), _opt(q), _opt(m), len(q), C.c_double(threshold),
                                 int(max_iterations), C.c_uint64(seed), _opt(keep), _opt(T), C.byref(ninl), C.byref(it),
                                 C.byref(bh)), "ransac_reject")
    return keep[: len(q)].astype(bool), T.reshape(4, 4), it.value, bh.value


# ------------------------------------------------------------------ ICP
def icp(src, tgt, max_corr_dist=0.07, max_iterations=100, transformation_epsilon=1e-6,
        euclidean_fitness_epsilon=1e-4, guess=None):
    """-> dict(T [4, 4] float32, fitness, converged, iterations, state)"""
    src, tgt = _f32(src), _f32(tgt)
    T = np.zeros(16, np.float32)
    g = None if guess is None else np.ascontiguousarray(guess, np.float32).reshape(16)
    fit = C.c_double(0)
    conv, it, st = C.c_int(0), C.c_int(0), C.c_int(0)
    _chk(lib().orc_icp(_opt(src), len(src), _opt(tgt), len(tgt), C.c_double(max_corr_dist), int(max_iterations),
                       C.c_double(transformation_epsilon), C.c_double(euclidean_fitness_epsilon),
                       None if g is None else g.ctypes.data_as(C.c_void_p), _opt(T), C.byref(fit), C.byref(conv),
                       C.byref(it), C.byref(st)), "icp")
    return dict(T=T.reshape(4, 4), fitness=fit.value, converged=bool(conv.value), iterations=it.value, state=st.value)


# ------------------------------------------------------------------ PFH125 / PrincipalCurvatures
def pfh125(surf, normals4, q, radius=0.0, k=0, want_counts=False):
    surf, q = _f32(surf), _f32(q)
    nr = np.ascontiguousarray(normals4, np.float32)
    out = np.zeros((len(q), 125), np.float32)
    cnt = np.zeros((len(q), 125), np.int32) if want_counts else None
    _chk(lib().orc_pfh125(_opt(surf), _opt(nr), len(surf), _opt(q), len(q), C.c_double(radius), int(k), _opt(out),
                          None if cnt is None else _opt(cnt)), "pfh125")
    return (out, cnt) if want_counts else out


def principal_curvatures(surf, normals4, q, radius=0.0, k=0):
    """-> (rows [nq, 5], relative eigen gap [nq])"""
    surf, q = _f32(surf), _f32(q)
    nr = np.ascontiguousarray(normals4, np.float32)
    out = np.zeros((len(q), 5), np.float32)
    gap = np.zeros(len(q), np.float32)
    _chk(lib().orc_principal_curvatures(_opt(surf), _opt(nr), len(surf), _opt(q), len(q), C.c_double(radius), int(k),
                                        _opt(out), _opt(gap)), "principal_curvatures")
    return out, gap


# ------------------------------------------------------------------ SHOT1344 (shape + colour)
def shot1344(surf, rgb, normals4, q, qrgb, radius, lrf_in=None, want_lab=False):
    """-> (rows [nq, 1344], frames [nq, 9][, lab [n, 3]]); rgb / qrgb: packed 0x00RRGGBB uint32"""
    surf, q = _f32(surf), _f32(q)
    nr = np.ascontiguousarray(normals4, np.float32)
    rgb = np.ascontiguousarray(rgb, np.uint32)
    qrgb = np.ascontiguousarray(qrgb, np.uint32)
    out = np.zeros((len(q), 1344), np.float32)
    rf = np.zeros((len(q), 9), np.float32)
    lab = np.zeros((len(surf), 3), np.float32) if want_lab else None
    lrf = None if lrf_in is None else np.ascontiguousarray(lrf_in, np.float32)
    _chk(lib().orc_shot1344(_opt(surf), _opt(rgb), _opt(nr), len(surf), _opt(q), _opt(qrgb), len(q), C.c_double(radius),
                            None if lrf is None else _opt(lrf), _opt(out), _opt(rf), None if lab is None else _opt(lab)),
         "shot1344")
    return (out, rf, lab) if want_lab else (out, rf)


def moment_invariants(surf, q, radius=0.0, k=0):
    surf, q = _f32(surf), _f32(q)
    out = np.zeros((len(q), 3), np.float32)
    _chk(lib().orc_moment_invariants(_opt(surf), len(surf), _opt(q), len(q), C.c_double(radius), int(k), _opt(out)),
         "moment_invariants")
    return out


# ------------------------------------------------------------------ Unique Shape Context
def usc1980(surf, q, search_radius, min_radius=None, density_radius=None, local_radius=2.5, lrf_in=None):
    """-> (rows [nq, 1980], frames [nq, 9], density [n]); defaults = the reference's settings (r / 10, r / 5, 2.5)"""
    surf, q = _f32(surf), _f32(q)
    out = np.zeros((len(q), 1980), np.float32)
    rf = np.zeros((len(q), 9), np.float32)
    dens = np.zeros(max(len(surf), 1), np.int32)
    lrf = None if lrf_in is None else np.ascontiguousarray(lrf_in, np.float32)
    _chk(lib().orc_usc1980(_opt(surf), len(surf), _opt(q), len(q), C.c_double(search_radius),
                           C.c_double(min_radius if min_radius is not None else search_radius / 10.0),
                           C.c_double(density_radius if density_radius is not None else search_radius / 5.0),
                           C.c_double(local_radius), None if lrf is None else _opt(lrf), _opt(out), _opt(rf), _opt(dens)),
         "usc1980")
    return out, rf, dens[: len(surf)]


def sc3d1980(surf, normals4, q, search_radius, min_radius=None, density_radius=None, seed=12345):
    """3DSC -> (rows [nq, 1980], frames [nq, 9] the descriptors were computed in); defaults = the reference's
    settings (r / 10, r / 5)"""
    surf, q = _f32(surf), _f32(q)
    nr = np.ascontiguousarray(normals4, np.float32)
    assert nr.shape == (len(surf), 4)
    out = np.zeros((len(q), 1980), np.float32)
    rf = np.zeros((len(q), 9), np.float32)
    L = lib()
    L.orc_sc3d1980.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                               C.c_ulonglong, C.c_void_p, C.c_void_p]
    _chk(L.orc_sc3d1980(_opt(surf), _opt(nr), len(surf), _opt(q), len(q), C.c_double(search_radius),
                        C.c_double(min_radius if min_radius is not None else search_radius / 10.0),
                        C.c_double(density_radius if density_radius is not None else search_radius / 5.0),
                        C.c_ulonglong(seed), _opt(out), _opt(rf)), "sc3d1980")
    return out, rf


def spin_image153(surf, q, qnormals4, radius):
    surf, q = _f32(surf), _f32(q)
    nr = np.ascontiguousarray(qnormals4, np.float32)
    assert nr.shape == (len(q), 4)
    out = np.zeros((len(q), 153), np.float32)
    _chk(lib().orc_spin_image153(_opt(surf), len(surf), _opt(q), _opt(nr), len(q), C.c_double(radius), _opt(out)), "spin_image153")
    return out

"""ctypes binding of the C ABI in include/pfx_b200.h (lib/libpfx_b200.so).

The library is the product: there is no Python or CPU fallback.  Importing works without a GPU (so
that the symbol table can be checked on a CPU box); creating a Context needs an sm_100 device and
raises otherwise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpfx_b200.so")

HOST, DEVICE, HOST_ASYNC = 0, 1, 2
E_INVALID, E_PRECOND, E_CAPACITY, E_STATE = -1, -2, -3, -4

_lib = None


class PfxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pfx error {code}: {msg}")
        self.code = code


class Correspondence(C.Structure):
    _fields_ = [("index_query", C.c_int32), ("index_match", C.c_int32), ("distance", C.c_float)]


class RangeImageDesc(C.Structure):
    """pfx_range_image_desc (include/pfx_b200.h)"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("planar", C.c_int32), ("cx", C.c_float), ("cy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("ang_res", C.c_float), ("off_x", C.c_int32), ("off_y", C.c_int32)]


class IcpParams(C.Structure):
    """pfx_icp_params; defaults = the reference's settings (evaluation.cpp:863-885)"""
    _fields_ = [("max_correspondence_distance", C.c_double), ("max_iterations", C.c_int),
                ("transformation_epsilon", C.c_double), ("euclidean_fitness_epsilon", C.c_double)]


class IcpResult(C.Structure):
    _fields_ = [("transform", C.c_float * 16), ("fitness", C.c_double), ("converged", C.c_int), ("iterations", C.c_int),
                ("state", C.c_int), ("correspondences", C.c_int)]


CORR_DTYPE = np.dtype([("index_query", "<i4"), ("index_match", "<i4"), ("distance", "<f4")])

# name -> (restype, argtypes); mirrors include/pfx_b200.h one to one
_vp, _sz, _i, _d, _f = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_float
SIGNATURES = {
    "pfx_version": (_i, []),
    "pfx_create": (_i, [_i, C.POINTER(_vp)]),
    "pfx_destroy": (_i, [_vp]),
    "pfx_last_error": (C.c_char_p, [_vp]),
    "pfx_set_stream": (_i, [_vp, _vp]),
    "pfx_set_parity_mode": (_i, [_vp, _i]),
    "pfx_set_reuse": (_i, [_vp, _i]),
    "pfx_reuse_info": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "pfx_group_unique_id": (_i, [_vp]),
    "pfx_group_join": (_i, [_vp, _i, _i, _vp]),
    "pfx_group_leave": (_i, [_vp]),
    "pfx_group_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "pfx_group_allreduce": (_i, [_vp, C.POINTER(_d), _i, _i]),
    "pfx_slab_distribute": (_i, [_vp, _vp, _sz, _sz, _vp, _i, _d, C.POINTER(_sz), C.POINTER(_sz)]),
    "pfx_slab_owned_rows": (_i, [_vp, _vp, _i]),
    "pfx_slab_global_ids": (_i, [_vp, _vp, _i]),
    "pfx_slab_info": (_i, [_vp, C.POINTER(_d)]),
    "pfx_match_ring": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _i, _i, _vp, _vp, _i]),
    "pfx_sync": (_i, [_vp]),
    "pfx_launch_count": (C.c_uint64, [_vp]),
    "pfx_grid_info": (_i, [_vp, C.POINTER(_d)]),
    "pfx_set_knn_occupancy": (_i, [_vp, _f]),
    "pfx_profile_begin": (_i, [_vp, C.c_char_p]),
    "pfx_profile_end": (_i, [_vp, C.c_char_p, _sz]),
    "pfx_set_surface": (_i, [_vp, _vp, _sz, _sz, _i]),
    "pfx_set_queries": (_i, [_vp, _vp, _sz, _sz, _i]),
    "pfx_prepare_radius": (_i, [_vp, _d]),
    "pfx_set_surface_normals": (_i, [_vp, _vp, _sz, _sz, _i, _i]),
    "pfx_set_viewpoint": (_i, [_vp, _f, _f, _f]),
    "pfx_num_surface": (_sz, [_vp]),
    "pfx_get_surface": (_i, [_vp, _vp, _sz, _i]),
    "pfx_num_queries": (_sz, [_vp]),
    "pfx_knn": (_i, [_vp, _i, _vp, _vp, _i]),
    "pfx_radius_count": (_i, [_vp, _d, _vp, C.POINTER(C.c_int64), _i]),
    "pfx_radius_search": (_i, [_vp, _d, _i, _vp, _vp, _vp, _i]),
    "pfx_normals": (_i, [_vp, _d, _i, _vp, _sz, _i, _i]),
    "pfx_cloud_resolution": (_i, [_vp, C.POINTER(_d)]),
    "pfx_iss": (_i, [_vp, _d, _d, _i, _d, _d, _vp, _sz, C.POINTER(_sz), _vp, _i]),
    "pfx_iss_nms": (_i, [_vp, _vp, _d, _i, _vp, _sz, C.POINTER(_sz), _i]),
    "pfx_harris3d": (_i, [_vp, _d, _f, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, C.POINTER(_sz), _i]),
    "pfx_harris6d": (_i, [_vp, _d, _f, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, C.POINTER(_sz), _i]),
    "pfx_harris_nms": (_i, [_vp, _vp, _d, _f, _vp, _sz, C.POINTER(_sz), _i]),
    "pfx_fpfh": (_i, [_vp, _d, _i, _vp, _sz, _i]),
    "pfx_spfh": (_i, [_vp, _d, _i, _vp, _i]),
    "pfx_pfh125": (_i, [_vp, _d, _i, _vp, _sz, _i]),
    "pfx_principal_curvatures": (_i, [_vp, _d, _i, _vp, _sz, _i]),
    "pfx_moment_invariants": (_i, [_vp, _d, _i, _vp, _sz, _i]),
    "pfx_seq_float_sum": (_f, [_f, C.c_longlong]),
    "pfx_shot352": (_i, [_vp, _d, _vp, _vp, _sz, _i]),
    "pfx_shot_lrf": (_i, [_vp, _d, _vp, _i]),
    "pfx_set_surface_colors": (_i, [_vp, _vp, _sz, _sz, _i]),
    "pfx_set_query_colors": (_i, [_vp, _vp, _sz, _sz, _i]),
    "pfx_shot1344": (_i, [_vp, _d, _vp, _vp, _sz, _i]),
    "pfx_spin_image153": (_i, [_vp, _d, _vp, _sz, _sz, _vp, _sz, _i]),
    "pfx_usc1980": (_i, [_vp, _d, _d, _d, _d, _vp, _vp, _sz, _i]),
    "pfx_sc3d1980": (_i, [_vp, _d, _d, _d, C.c_uint64, _vp, _sz, _vp, _i]),
    "pfx_match": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _i, _i, _f, _vp, _sz, C.POINTER(_sz), _i]),
    "pfx_match_nn": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _i, _vp, _vp, _i]),
    "pfx_range_image_planar": (_i, [_vp, _i, _i, _f, _f, _f, _f, _f, _vp]),
    "pfx_range_image_spherical": (_i, [_vp, _f, _f, _f, _f, _i, _vp]),
    "pfx_range_image_set_pose": (_i, [_vp, _vp]),
    "pfx_range_image_set": (_i, [_vp, _vp, _vp, _i]),
    "pfx_range_image_get": (_i, [_vp, _vp, _vp, _i]),
    "pfx_narf_borders": (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    "pfx_narf_keypoints": (_i, [_vp, _f, _vp, _vp, _vp, _sz, C.POINTER(_sz), _vp, _i]),
    "pfx_narf36": (_i, [_vp, _vp, _sz, _f, _i, _vp, _sz, _sz, C.POINTER(_sz), _i]),
    "pfx_ransac_reject": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _d, _i, C.c_uint64, _vp, _sz, C.POINTER(_sz),
                                _vp, _vp, _vp, _i]),
    "pfx_icp_align": (_i, [_vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _sz, _i]),
    "pfx_set_match_engine": (_i, [_vp, _i]),
    "pfx_match_info": (_i, [_vp, _vp]),
    "pfx_voxel_grid": (_i, [_vp, _f, _vp, _sz, C.POINTER(_sz), _i]),
}


def load():
    """Loads lib/libpfx_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C pcl_feature_extraction_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):  # torch tensor
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Context:
    """One device context (pfx_ctx).  Thin: every method is one C-ABI call with numpy buffers."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.pfx_create(int(device), C.byref(h))
        if rc != 0:
            raise PfxError(rc, "pfx_create failed (an sm_100 CUDA device is required; there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.pfx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise PfxError(rc, self.lib.pfx_last_error(self.h).decode())

    # -- plumbing
    def set_stream(self, stream_ptr):
        self._chk(self.lib.pfx_set_stream(self.h, C.c_void_p(stream_ptr)))

    def sync(self):
        self._chk(self.lib.pfx_sync(self.h))

    @property
    def launches(self):
        return int(self.lib.pfx_launch_count(self.h))

    def grid_info(self):
        out = (C.c_double * 8)()
        self._chk(self.lib.pfx_grid_info(self.h, out))
        return dict(edge=out[0], dims=(int(out[1]), int(out[2]), int(out[3])), ncells=int(out[4]), n_valid=int(out[5]),
                    tile_flagged=int(out[6]))

    def set_knn_occupancy(self, frac):
        self._chk(self.lib.pfx_set_knn_occupancy(self.h, frac))

    def profile_begin(self, kernel_filter=None):
        self._chk(self.lib.pfx_profile_begin(self.h, kernel_filter.encode() if kernel_filter else None))

    def profile_end(self):
        """-> {kernel name: (launches, total ms)} measured with CUDA events on the launching stream"""
        buf = C.create_string_buffer(1 << 16)
        self._chk(self.lib.pfx_profile_end(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split("\t")
            out[name] = (int(cnt), float(ms))
        return out

    # -- inputs (numpy host arrays; *_dev variants take raw device pointers)
    def set_surface(self, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        self._surface_keepalive = pts
        self._chk(self.lib.pfx_set_surface(self.h, _ptr(pts), len(pts), pts.strides[0] if len(pts) else 12, HOST))

    def set_surface_dev(self, ptr, n, stride):
        self._chk(self.lib.pfx_set_surface(self.h, _ptr(ptr), n, stride, DEVICE))

    def set_queries(self, pts):
        if pts is None:
            self._chk(self.lib.pfx_set_queries(self.h, None, 0, 0, HOST))
            return
        pts = np.ascontiguousarray(pts, np.float32)
        self._chk(self.lib.pfx_set_queries(self.h, _ptr(pts), len(pts), pts.strides[0] if len(pts) else 12, HOST))

    def set_queries_dev(self, ptr, n, stride):
        self._chk(self.lib.pfx_set_queries(self.h, _ptr(ptr), n, stride, DEVICE))

    def prepare_radius(self, radius):
        """hint: build the radius index now on the auxiliary stream (overlaps the calls that follow)"""
        self._chk(self.lib.pfx_prepare_radius(self.h, radius))

    def set_surface_normals(self, normals4):
        nr = np.ascontiguousarray(normals4, np.float32)
        assert nr.ndim == 2 and nr.shape[1] in (4, 8)
        curv = 3 if nr.shape[1] == 4 else 4
        self._chk(self.lib.pfx_set_surface_normals(self.h, _ptr(nr), len(nr), nr.strides[0] if len(nr) else 16, curv, HOST))

    def set_viewpoint(self, x, y, z):
        self._chk(self.lib.pfx_set_viewpoint(self.h, x, y, z))

    @property
    def num_queries(self):
        return int(self.lib.pfx_num_queries(self.h))

    @property
    def num_surface(self):
        return int(self.lib.pfx_num_surface(self.h))

    # -- search
    def knn(self, k):
        nq = self.num_queries
        idx = np.zeros((nq, k), np.int32)
        d2 = np.zeros((nq, k), np.float32)
        self._chk(self.lib.pfx_knn(self.h, k, _ptr(idx), _ptr(d2), HOST))
        return idx, d2

    def radius_count(self, radius):
        counts = np.zeros(self.num_queries, np.int32)
        total = C.c_int64(0)
        self._chk(self.lib.pfx_radius_count(self.h, radius, _ptr(counts), C.byref(total), HOST))
        return counts

    def radius_search(self, radius, sorted=True):
        nq = self.num_queries
        counts = np.zeros(nq, np.int32)
        total = C.c_int64(0)
        self._chk(self.lib.pfx_radius_count(self.h, radius, _ptr(counts), C.byref(total), HOST))
        offsets = np.zeros(nq + 1, np.int64)
        np.cumsum(counts, out=offsets[1:])
        assert offsets[-1] == total.value
        idx = np.zeros(int(total.value), np.int32)
        d2 = np.zeros(int(total.value), np.float32)
        self._chk(self.lib.pfx_radius_search(self.h, radius, 1 if sorted else 0, _ptr(offsets), _ptr(idx), _ptr(d2), HOST))
        return offsets, idx, d2

    # -- normals
    def normals(self, radius=0.0, k=0, want_output=True):
        nq = self.num_queries
        out = np.zeros((nq, 4), np.float32) if want_output else None
        self._chk(self.lib.pfx_normals(self.h, radius, k, _ptr(out), 16, 3, HOST))
        return out

    def normals_dev(self, radius, k, out_ptr, stride=16, curv_off=3):
        self._chk(self.lib.pfx_normals(self.h, radius, k, _ptr(out_ptr), stride, curv_off, DEVICE))

    # -- keypoints
    def cloud_resolution(self):
        r = C.c_double(0)
        self._chk(self.lib.pfx_cloud_resolution(self.h, C.byref(r)))
        return r.value

    def iss(self, salient_radius, nonmax_radius, min_neighbors=5, g21=0.975, g32=0.975):
        n = self.num_surface
        kp = np.zeros(n, np.int32)
        sal = np.zeros(n, np.float64)
        nk = C.c_size_t(0)
        self._chk(self.lib.pfx_iss(self.h, salient_radius, nonmax_radius, min_neighbors, g21, g32, _ptr(kp), n,
                                   C.byref(nk), _ptr(sal), HOST))
        return kp[: nk.value].copy(), sal

    def iss_nms(self, saliency, nonmax_radius, min_neighbors=5):
        n = self.num_surface
        sal = np.ascontiguousarray(saliency, np.float64)
        kp = np.zeros(n, np.int32)
        nk = C.c_size_t(0)
        self._chk(self.lib.pfx_iss_nms(self.h, _ptr(sal), nonmax_radius, min_neighbors, _ptr(kp), n, C.byref(nk), HOST))
        return kp[: nk.value].copy()

    def harris3d(self, radius=0.01, threshold=1e-6, nonmax=True, refine=True, snap_max_d2=1e-4):
        n = self.num_surface
        resp = np.zeros(n, np.float32)
        kp = np.zeros(n, np.int32)
        xyz = np.zeros((n, 3), np.float32)
        snap = np.zeros(n, np.int32)
        nk = C.c_size_t(0)
        self._chk(self.lib.pfx_harris3d(self.h, radius, threshold, int(nonmax), int(refine), snap_max_d2, _ptr(resp),
                                        _ptr(kp), _ptr(xyz), _ptr(snap), n, C.byref(nk), HOST))
        m = nk.value
        return dict(response=resp, kp_idx=kp[:m].copy(), kp_xyz=xyz[:m].copy(), snapped_idx=snap[:m].copy())

    def harris6d(self, radius=0.01, threshold=1e-6, nonmax=True, refine=True, snap_max_d2=1e-4):
        """HarrisKeypoint6D; the surface colours must have been set (set_surface_colors)"""
        n = self.num_surface
        resp = np.zeros(n, np.float32)
        kp = np.zeros(n, np.int32)
        xyz = np.zeros((n, 3), np.float32)
        snap = np.zeros(n, np.int32)
        nk = C.c_size_t(0)
        self._chk(self.lib.pfx_harris6d(self.h, radius, threshold, int(nonmax), int(refine), snap_max_d2, _ptr(resp),
                                        _ptr(kp), _ptr(xyz), _ptr(snap), n, C.byref(nk), HOST))
        m = nk.value
        return dict(response=resp, kp_idx=kp[:m].copy(), kp_xyz=xyz[:m].copy(), snapped_idx=snap[:m].copy())

    def harris_nms(self, response, radius, threshold):
        n = self.num_surface
        resp = np.ascontiguousarray(response, np.float32)
        kp = np.zeros(n, np.int32)
        nk = C.c_size_t(0)
        self._chk(self.lib.pfx_harris_nms(self.h, _ptr(resp), radius, threshold, _ptr(kp), n, C.byref(nk), HOST))
        return kp[: nk.value].copy()

    # -- descriptors
    def fpfh(self, radius=0.0, k=0):
        out = np.zeros((self.num_queries, 33), np.float32)
        self._chk(self.lib.pfx_fpfh(self.h, radius, k, _ptr(out), 132, HOST))
        return out

    def fpfh_dev(self, radius, k, out_ptr, stride=132):
        self._chk(self.lib.pfx_fpfh(self.h, radius, k, _ptr(out_ptr), stride, DEVICE))

    def pfh125(self, radius=0.0, k=0):
        out = np.zeros((self.num_queries, 125), np.float32)
        self._chk(self.lib.pfx_pfh125(self.h, radius, k, _ptr(out), 500, HOST))
        return out

    def principal_curvatures(self, radius=0.0, k=0):
        """rows of (principal direction x, y, z, pc1, pc2)"""
        out = np.zeros((self.num_queries, 5), np.float32)
        self._chk(self.lib.pfx_principal_curvatures(self.h, radius, k, _ptr(out), 20, HOST))
        return out

    def moment_invariants(self, radius=0.0, k=0):
        out = np.zeros((self.num_queries, 3), np.float32)
        self._chk(self.lib.pfx_moment_invariants(self.h, radius, k, _ptr(out), 12, HOST))
        return out

    def spfh(self, radius=0.0, k=0):
        out = np.zeros((self.num_surface, 33), np.float32)
        self._chk(self.lib.pfx_spfh(self.h, radius, k, _ptr(out), HOST))
        return out

    def shot_lrf(self, radius):
        out = np.zeros((self.num_queries, 9), np.float32)
        self._chk(self.lib.pfx_shot_lrf(self.h, radius, _ptr(out), HOST))
        return out

    def shot352(self, radius, lrf_in=None):
        out = np.zeros((self.num_queries, 361), np.float32)
        lrf = np.ascontiguousarray(lrf_in, np.float32) if lrf_in is not None else None
        self._chk(self.lib.pfx_shot352(self.h, radius, _ptr(lrf), _ptr(out), 1444, HOST))
        return out[:, :352].copy(), out[:, 352:].copy()

    def set_surface_colors(self, rgb):
        """packed 0x00RRGGBB uint32 per surface point"""
        rgb = np.ascontiguousarray(rgb, np.uint32)
        self._chk(self.lib.pfx_set_surface_colors(self.h, _ptr(rgb), len(rgb), 4, HOST))

    def set_query_colors(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint32)
        self._chk(self.lib.pfx_set_query_colors(self.h, _ptr(rgb), len(rgb), 4, HOST))

    def shot1344(self, radius, lrf_in=None):
        out = np.zeros((self.num_queries, 1353), np.float32)
        lrf = np.ascontiguousarray(lrf_in, np.float32) if lrf_in is not None else None
        self._chk(self.lib.pfx_shot1344(self.h, radius, _ptr(lrf), _ptr(out), 5412, HOST))
        return out[:, :1344].copy(), out[:, 1344:].copy()

    def spin_image153(self, radius, query_normals):
        """query_normals: [nq, 3 or 4] normals of the QUERIES (the surface itself when no queries are set)"""
        nr = np.ascontiguousarray(query_normals, np.float32)
        out = np.zeros((self.num_queries, 153), np.float32)
        self._chk(self.lib.pfx_spin_image153(self.h, radius, _ptr(nr), len(nr), nr.strides[0] if len(nr) else 12, _ptr(out), 612, HOST))
        return out

    def usc1980(self, search_radius, min_radius=None, density_radius=None, local_radius=2.5, lrf_in=None):
        """-> (rows [nq, 1980], frames [nq, 9]); defaults = the reference's settings (r / 10, r / 5, 2.5)"""
        out = np.zeros((self.num_queries, 1989), np.float32)
        lrf = np.ascontiguousarray(lrf_in, np.float32) if lrf_in is not None else None
        self._chk(self.lib.pfx_usc1980(self.h, search_radius, min_radius if min_radius is not None else search_radius / 10.0,
                                       density_radius if density_radius is not None else search_radius / 5.0, local_radius,
                                       _ptr(lrf), _ptr(out), 7956, HOST))
        return out[:, :1980].copy(), out[:, 1980:].copy()

    def sc3d1980(self, search_radius, min_radius=None, density_radius=None, seed=12345):
        """3DSC -> (rows [nq, 1980], frames [nq, 9] the descriptors were computed in); defaults = the reference's
        settings (r / 10, r / 5).  Needs surface normals."""
        nq = self.num_queries
        out = np.zeros((nq, 1989), np.float32)
        fr = np.zeros((nq, 9), np.float32)
        self._chk(self.lib.pfx_sc3d1980(self.h, search_radius, min_radius if min_radius is not None else search_radius / 10.0,
                                        density_radius if density_radius is not None else search_radius / 5.0, int(seed),
                                        _ptr(out), 7956, _ptr(fr), HOST))
        return out[:, :1980].copy(), fr, out[:, 1980:].copy()

    def shot352_dev(self, radius, out_ptr, stride=1444):
        self._chk(self.lib.pfx_shot352(self.h, radius, None, _ptr(out_ptr), stride, DEVICE))

    # -- ingest
    def voxel_grid(self, leaf):
        """pcl::VoxelGrid centroids of the current surface, ascending voxel id"""
        n = max(self.num_surface, 1)
        out = np.zeros((n, 3), np.float32)
        m = C.c_size_t(0)
        self._chk(self.lib.pfx_voxel_grid(self.h, leaf, _ptr(out), n, C.byref(m), HOST))
        return out[: m.value].copy()

    # -- range image / NARF
    def range_image_planar(self, width, height, cx, cy, fx, fy, min_range=0.0):
        d = RangeImageDesc()
        self._chk(self.lib.pfx_range_image_planar(self.h, width, height, cx, cy, fx, fy, min_range, C.byref(d)))
        return d

    def range_image_set_pose(self, pose4x4=None):
        """sensor pose (world <- sensor, 4x4) of the range images built next; None = identity"""
        p = None if pose4x4 is None else np.ascontiguousarray(pose4x4, np.float32).reshape(16)
        self._chk(self.lib.pfx_range_image_set_pose(self.h, _ptr(p)))

    def range_image_spherical(self, ang_res, max_angle_w=2 * np.pi, max_angle_h=np.pi, min_range=0.0, border=0):
        d = RangeImageDesc()
        self._chk(self.lib.pfx_range_image_spherical(self.h, ang_res, max_angle_w, max_angle_h, min_range, border, C.byref(d)))
        return d

    def range_image_set(self, desc, img):
        img = np.ascontiguousarray(img, np.float32)
        self._chk(self.lib.pfx_range_image_set(self.h, C.byref(desc), _ptr(img), HOST))

    def range_image_get(self):
        d = RangeImageDesc()
        self._chk(self.lib.pfx_range_image_get(self.h, C.byref(d), None, HOST))
        img = np.zeros((d.height, d.width, 4), np.float32)
        if img.size:
            self._chk(self.lib.pfx_range_image_get(self.h, C.byref(d), _ptr(img), HOST))
        return d, img

    def narf_borders(self):
        d, _ = RangeImageDesc(), None
        self._chk(self.lib.pfx_range_image_get(self.h, C.byref(d), None, HOST))
        n = d.width * d.height
        traits = np.zeros(n, np.int32)
        scores = np.zeros((4, n), np.float32)
        cs = np.zeros(n, np.float32)
        cd = np.zeros((n, 3), np.float32)
        self._chk(self.lib.pfx_narf_borders(self.h, _ptr(traits), _ptr(scores), _ptr(cs), _ptr(cd), HOST))
        return traits, scores, cs, cd

    def narf_keypoints(self, support_size):
        d = RangeImageDesc()
        self._chk(self.lib.pfx_range_image_get(self.h, C.byref(d), None, HOST))
        n = max(d.width * d.height, 1)
        kp = np.zeros(n, np.int32)
        xyz = np.zeros((n, 3), np.float32)
        val = np.zeros(n, np.float32)
        interest = np.zeros(n, np.float32)
        m = C.c_size_t(0)
        self._chk(self.lib.pfx_narf_keypoints(self.h, support_size, _ptr(kp), _ptr(xyz), _ptr(val), n, C.byref(m),
                                              _ptr(interest), HOST))
        return kp[: m.value].copy(), xyz[: m.value].copy(), val[: m.value].copy(), interest

    def narf36(self, kp_px, support_size, rotation_invariant=True):
        kp_px = np.ascontiguousarray(kp_px, np.int32)
        cap = max(1, 8 * len(kp_px))
        out = np.zeros((cap, 42), np.float32)
        m = C.c_size_t(0)
        self._chk(self.lib.pfx_narf36(self.h, _ptr(kp_px), len(kp_px), support_size, int(bool(rotation_invariant)),
                                      _ptr(out), 168, cap, C.byref(m), HOST))
        return out[: m.value].copy()

    # -- matching
    def match_nn(self, a, b):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        dim = a.shape[1] if a.ndim == 2 and a.shape[1] else b.shape[1]
        idx = np.zeros(len(a), np.int32)
        d2 = np.zeros(len(a), np.float32)
        self._chk(self.lib.pfx_match_nn(self.h, _ptr(a), len(a), dim * 4, _ptr(b), len(b), dim * 4, dim, _ptr(idx),
                                        _ptr(d2), HOST))
        return idx, d2

    def match_nn_dev(self, a_ptr, na, b_ptr, nb, dim, idx_ptr, d2_ptr, stride_a=None, stride_b=None):
        """device-resident rows (row stride in bytes, default dim * 4); idx / d2: device buffers of na entries"""
        self._chk(self.lib.pfx_match_nn(self.h, _ptr(a_ptr), na, stride_a or dim * 4, _ptr(b_ptr), nb,
                                        stride_b or dim * 4, dim, _ptr(idx_ptr), _ptr(d2_ptr), DEVICE))

    def match(self, a, b, reciprocal=True, max_dist2=-1.0):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        dim = a.shape[1]
        out = np.zeros(max(len(a), 1), CORR_DTYPE)
        m = C.c_size_t(0)
        self._chk(self.lib.pfx_match(self.h, _ptr(a), len(a), dim * 4, _ptr(b), len(b), dim * 4, dim, int(reciprocal),
                                     max_dist2, _ptr(out), len(out), C.byref(m), HOST))
        return out[: m.value].copy()

    def ransac_reject(self, src, tgt, corr, threshold=0.015, max_iterations=1000, seed=12345):
        """-> (surviving correspondences, T [4, 4], iterations, best hypothesis)"""
        src = np.ascontiguousarray(src, np.float32)
        tgt = np.ascontiguousarray(tgt, np.float32)
        corr = np.ascontiguousarray(corr, CORR_DTYPE)
        out = np.zeros(max(len(corr), 1), CORR_DTYPE)
        T = np.zeros(16, np.float32)
        m = C.c_size_t(0)
        it, bh = C.c_int(0), C.c_int(0)
        self._chk(self.lib.pfx_ransac_reject(self.h, _ptr(src), len(src), src.strides[0] if len(src) else 12, _ptr(tgt), len(tgt),
                                             tgt.strides[0] if len(tgt) else 12, _ptr(corr), len(corr), threshold, max_iterations,
                                             seed, _ptr(out), len(out) if len(corr) else 0, C.byref(m), _ptr(T), C.byref(it),
                                             C.byref(bh), HOST))
        return out[: m.value].copy(), T.reshape(4, 4), it.value, bh.value

    def icp_align(self, src, max_corr_dist=0.07, max_iterations=100, transformation_epsilon=1e-6,
                  euclidean_fitness_epsilon=1e-4, guess=None, want_aligned=False):
        """ICP of `src` onto the current surface -> dict(T [4, 4], fitness, converged, iterations, state,
        correspondences[, aligned])"""
        src = np.ascontiguousarray(src, np.float32)
        prm = IcpParams(max_corr_dist, max_iterations, transformation_epsilon, euclidean_fitness_epsilon)
        res = IcpResult()
        g = None if guess is None else np.ascontiguousarray(guess, np.float32).reshape(16)
        al = np.zeros((len(src), 3), np.float32) if want_aligned else None
        self._chk(self.lib.pfx_icp_align(self.h, _ptr(src), len(src), src.strides[0] if len(src) else 12, C.byref(prm),
                                         None if g is None else _ptr(g), C.byref(res),
                                         None if al is None else _ptr(al), 12, HOST))
        out = dict(T=np.array(res.transform, np.float32).reshape(4, 4), fitness=res.fitness, converged=bool(res.converged),
                   iterations=res.iterations, state=res.state, correspondences=res.correspondences)
        if want_aligned:
            out["aligned"] = al
        return out

    # -- multi-GPU (group.cu): one rank per GPU, NCCL inside the library
    @staticmethod
    def group_unique_id():
        """128 bytes (ncclUniqueId) made on one rank; every rank passes them to group_join"""
        buf = (C.c_ubyte * 128)()
        rc = load().pfx_group_unique_id(buf)
        if rc != 0:
            raise PfxError(rc, "pfx_group_unique_id failed (NCCL not available)")
        return bytes(buf)

    def group_join(self, rank, world, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._chk(self.lib.pfx_group_join(self.h, rank, world, buf))

    def group_leave(self):
        self._chk(self.lib.pfx_group_leave(self.h))

    def group_allreduce(self, vals, op="sum"):
        a = (C.c_double * len(vals))(*[float(v) for v in vals])
        self._chk(self.lib.pfx_group_allreduce(self.h, a, len(vals), {"sum": 0, "max": 1, "min": 2}[op]))
        return [float(v) for v in a]

    def slab_distribute(self, part, halo, global_ids=None, mem=HOST, stride=None):
        """part: [n, >=3] float32 rows (numpy for HOST; a device pointer + explicit n via (ptr, n) for DEVICE).
        -> (n_owned, n_local); the context's surface becomes owned + halo points"""
        if mem == HOST:
            part = np.ascontiguousarray(part, np.float32)
            n, st = len(part), part.strides[0] if len(part) else 12
            ptr = _ptr(part)
            gid = None if global_ids is None else _ptr(np.ascontiguousarray(global_ids, np.int32))
        else:
            ptr_i, n = part
            ptr, st = _ptr(ptr_i), stride or 16
            gid = None if global_ids is None else _ptr(global_ids)
        no, nl = C.c_size_t(0), C.c_size_t(0)
        self._chk(self.lib.pfx_slab_distribute(self.h, ptr, n, st, gid, mem, float(halo), C.byref(no), C.byref(nl)))
        return no.value, nl.value

    def slab_owned_rows(self):
        """local rows (ascending) of this rank's own points"""
        info = self.slab_info()
        out = np.zeros(info["n_owned"], np.int32)
        self._chk(self.lib.pfx_slab_owned_rows(self.h, _ptr(out), HOST))
        return out

    def slab_global_ids(self):
        info = self.slab_info()
        out = np.zeros(info["n_local"], np.int32)
        self._chk(self.lib.pfx_slab_global_ids(self.h, _ptr(out), HOST))
        return out

    def slab_info(self):
        a = (C.c_double * 6)()
        self._chk(self.lib.pfx_slab_info(self.h, a))
        return dict(axis=int(a[0]), n_owned=int(a[1]), n_local=int(a[2]), n_total=int(a[3]), lo=a[4], hi=a[5])

    def match_ring_dev(self, a_ptr, na, b_ptr, nb, dim, b_offset, idx_ptr, d2_ptr, stride_a=None, stride_b=None):
        """both sides sharded over the group's ranks, device buffers; idx = GLOBAL target rows"""
        self._chk(self.lib.pfx_match_ring(self.h, _ptr(a_ptr), na, stride_a or dim * 4, _ptr(b_ptr), nb, stride_b or dim * 4,
                                          dim, int(b_offset), _ptr(idx_ptr), _ptr(d2_ptr), DEVICE))

    def set_reuse(self, enable):
        self._chk(self.lib.pfx_set_reuse(self.h, 1 if enable else 0))

    def reuse_info(self):
        a = (C.c_uint64 * 6)()
        self._chk(self.lib.pfx_reuse_info(self.h, a))
        return dict(surface_uploads=int(a[0]), surface_reused=int(a[1]), normals_passes=int(a[2]), normals_reused=int(a[3]),
                    normals_uploads=int(a[4]), normals_upload_skipped=int(a[5]))

    def set_parity_mode(self, strict):
        """True: reference-order arithmetic for normals / Harris3D / radius-search FPFH (bit-identical to the CPU
        path, include/pfx_b200.h); False: the throughput kernels (default)"""
        self._chk(self.lib.pfx_set_parity_mode(self.h, 1 if strict else 0))

    def set_match_engine(self, engine):
        self._chk(self.lib.pfx_set_match_engine(self.h, engine))

    def match_info(self):
        out = np.zeros(4, np.float64)
        self._chk(self.lib.pfx_match_info(self.h, _ptr(out)))
        return {"tc_passes": int(out[0]), "rows": int(out[1]), "redone_exact": int(out[2])}

"""ctypes front-end of the CPU ORACLE (oracle/o3d_oracle.c).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see the header of o3d_oracle.c): the
reference has no tests on this path and Open3D v0.15.1 is not available here.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libo3d_oracle.so")

CROP_NONE, CROP_MAX_RADIUS, CROP_MIN_RADIUS, CROP_MINMAX_RADIUS, CROP_CYLINDER = 0, 1, 2, 3, 4
_CROP_NAMES = {"None": 0, "MaxRadius": 1, "MinRadius": 2, "MinMaxRadius": 3, "Cylinder": 4}


class Cropper(C.Structure):
    _fields_ = [("kind", C.c_int32), ("invert", C.c_int32), ("rmin", C.c_double), ("rmax", C.c_double),
                ("zmin", C.c_double), ("zmax", C.c_double), ("center", C.c_double * 3)]


def cropper(kind="MinMaxRadius", rmin=0.0, rmax=20.0, zmin=-10.0, zmax=10.0, center=(0.0, 0.0, 0.0), invert=False) -> Cropper:
    c = Cropper()
    c.kind = _CROP_NAMES[kind] if isinstance(kind, str) else int(kind)
    c.invert = int(invert)
    c.rmin, c.rmax, c.zmin, c.zmax = float(rmin), float(rmax), float(zmin), float(zmax)
    c.center[0], c.center[1], c.center[2] = (float(v) for v in center)
    return c


class IcpResultC(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("fitness", C.c_double), ("inlier_rmse", C.c_double), ("n_corr", C.c_int32),
                ("iters", C.c_int32)]


@dataclass
class IcpResult:
    T: np.ndarray
    fitness: float
    inlier_rmse: float
    n_corr: int
    iters: int
    trace: np.ndarray | None = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O3 -fopenmp)."""
    src = os.path.join(_HERE, "o3d_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_crop.restype = C.c_size_t
        _lib.orc_voxel_down_sample.restype = C.c_size_t
        _lib.orc_random_down_sample.restype = C.c_size_t
        _lib.orc_transform.restype = C.c_size_t
        _lib.orc_voxelize_within_cropping_volume.restype = C.c_size_t
        _lib.orc_submap_insert_scan.restype = C.c_size_t
        _lib.orc_dense_create.restype = C.c_void_p
        _lib.orc_dense_to_cloud.restype = C.c_size_t
        _lib.orc_kdtree_build.restype = C.c_void_p
        _lib.orc_select_hash.restype = C.c_uint32
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def crop(c: Cropper, xyz, nrm=None):
    xyz = _f64(xyz).reshape(-1, 3)
    n = len(xyz)
    nrm = None if nrm is None else _f64(nrm).reshape(-1, 3)
    ox = np.empty((n, 3)); on = np.empty((n, 3)) if nrm is not None else None
    m = lib().orc_crop(C.byref(c), _p(xyz), _p(nrm), C.c_size_t(n), _p(ox), _p(on))
    return (ox[:m].copy(), None if on is None else on[:m].copy())


def voxel_down_sample(xyz, voxel, nrm=None, return_keys=False):
    xyz = _f64(xyz).reshape(-1, 3)
    n = len(xyz)
    nrm = None if nrm is None else _f64(nrm).reshape(-1, 3)
    ox = np.empty((n, 3)); on = np.empty((n, 3)) if nrm is not None else None
    keys = np.empty((n, 3), dtype=np.int32)
    m = lib().orc_voxel_down_sample(_p(xyz), _p(nrm), C.c_size_t(n), C.c_double(voxel), _p(ox), _p(on), _p(keys))
    out = (ox[:m].copy(), None if on is None else on[:m].copy())
    return out + (keys[:m].copy(),) if return_keys else out


def estimate_normals(xyz, knn, radius, return_cov=False):
    xyz = _f64(xyz).reshape(-1, 3)
    n = len(xyz)
    on = np.empty((n, 3)); cov = np.empty((n, 9)) if return_cov else None
    lib().orc_estimate_normals(_p(xyz), C.c_size_t(n), C.c_int(knn), C.c_double(radius), _p(on), _p(cov))
    return (on, cov.reshape(n, 3, 3)) if return_cov else on


def fast_eigen3x3(cov):
    cov = _f64(cov).reshape(9)
    out = np.empty(3)
    lib().orc_fast_eigen3x3(_p(cov), _p(out))
    return out


def random_down_sample(xyz, ratio, seed, nrm=None):
    xyz = _f64(xyz).reshape(-1, 3)
    n = len(xyz)
    nrm = None if nrm is None else _f64(nrm).reshape(-1, 3)
    ox = np.empty((n, 3)); on = np.empty((n, 3)) if nrm is not None else None
    m = lib().orc_random_down_sample(_p(xyz), _p(nrm), C.c_size_t(n), C.c_double(ratio), C.c_uint32(seed), _p(ox), _p(on))
    return (ox[:m].copy(), None if on is None else on[:m].copy())


def select_hash(seed, point):
    p = _f64(point).reshape(3)
    return int(lib().orc_select_hash(C.c_uint32(seed), _p(p)))


def registration_icp_p2plane(src, tgt, tgt_nrm, max_corr_dist, init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6,
                             trace=False) -> IcpResult:
    src = _f64(src).reshape(-1, 3); tgt = _f64(tgt).reshape(-1, 3); tgt_nrm = _f64(tgt_nrm).reshape(-1, 3)
    init = np.eye(4) if init is None else _f64(init).reshape(4, 4)
    res = IcpResultC()
    tr = np.zeros((max(max_iter, 1), 44)) if trace else None
    rc = lib().orc_registration_icp_p2plane(_p(src), C.c_size_t(len(src)), _p(tgt), _p(tgt_nrm), C.c_size_t(len(tgt)),
                                            C.c_double(max_corr_dist), _p(init), C.c_int(max_iter), C.c_double(rel_fitness),
                                            C.c_double(rel_rmse), C.byref(res), _p(tr), C.c_int(max_iter if trace else 0))
    if rc != 0:
        raise RuntimeError(f"orc_registration_icp_p2plane failed: {rc}")
    return IcpResult(np.array(res.T).reshape(4, 4), res.fitness, res.inlier_rmse, res.n_corr, res.iters,
                     None if tr is None else tr[:res.iters].copy())


def registration_icp_p2point(src, tgt, max_corr_dist, init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6) -> IcpResult:
    """RegistrationIcpPointToPoint::registerClouds (core/src/CloudRegistration.cpp:69-75): ICP with Eigen::umeyama updates."""
    src = _f64(src).reshape(-1, 3); tgt = _f64(tgt).reshape(-1, 3)
    init = np.eye(4) if init is None else _f64(init).reshape(4, 4)
    res = IcpResultC()
    rc = lib().orc_registration_icp_p2point(_p(src), C.c_size_t(len(src)), _p(tgt), C.c_size_t(len(tgt)), C.c_double(max_corr_dist), _p(init),
                                            C.c_int(max_iter), C.c_double(rel_fitness), C.c_double(rel_rmse), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"orc_registration_icp_p2point failed: {rc}")
    return IcpResult(np.array(res.T).reshape(4, 4), res.fitness, res.inlier_rmse, res.n_corr, res.iters, None)


def svd3(A):
    A = _f64(A).reshape(3, 3)
    U = np.empty((3, 3)); S = np.empty(3); V = np.empty((3, 3))
    lib().orc_svd3(_p(A), _p(U), _p(S), _p(V))
    return U, S, V


def registration_gicp(src, src_nrm, tgt, tgt_nrm, max_corr_dist, init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6, epsilon=1e-3) -> IcpResult:
    """RegistrationIcpGeneralized::registerClouds (core/src/CloudRegistration.cpp:15-20) for clouds that carry normals."""
    src = _f64(src).reshape(-1, 3); tgt = _f64(tgt).reshape(-1, 3); src_nrm = _f64(src_nrm).reshape(-1, 3); tgt_nrm = _f64(tgt_nrm).reshape(-1, 3)
    init = np.eye(4) if init is None else _f64(init).reshape(4, 4)
    res = IcpResultC()
    rc = lib().orc_registration_gicp(_p(src), _p(src_nrm), C.c_size_t(len(src)), _p(tgt), _p(tgt_nrm), C.c_size_t(len(tgt)), C.c_double(max_corr_dist),
                                     _p(init), C.c_int(max_iter), C.c_double(rel_fitness), C.c_double(rel_rmse), C.c_double(epsilon), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"orc_registration_gicp failed: {rc}")
    return IcpResult(np.array(res.T).reshape(4, 4), res.fitness, res.inlier_rmse, res.n_corr, res.iters, None)


def gicp_covariance_from_normal(n, epsilon=1e-3):
    n = _f64(n).reshape(3); Cm = np.empty((3, 3))
    lib().orc_gicp_covariance_from_normal(_p(n), C.c_double(epsilon), _p(Cm))
    return Cm


def icp_evaluate_bruteforce(src, tgt, tgt_nrm, r, T):
    src = _f64(src).reshape(-1, 3); tgt = _f64(tgt).reshape(-1, 3); tgt_nrm = _f64(tgt_nrm).reshape(-1, 3)
    T = _f64(T).reshape(4, 4)
    fit = C.c_double(); rm = C.c_double()
    JTJ = np.empty(36); JTr = np.empty(6); corr = np.empty(len(src), dtype=np.int32)
    lib().orc_icp_evaluate_bruteforce(_p(src), C.c_size_t(len(src)), _p(tgt), _p(tgt_nrm), C.c_size_t(len(tgt)), C.c_double(r),
                                      _p(T), C.byref(fit), C.byref(rm), _p(JTJ), _p(JTr), _p(corr))
    return fit.value, rm.value, JTJ.reshape(6, 6), JTr, corr


def ldlt6_solve(A, b):
    A = _f64(A).reshape(36); b = _f64(b).reshape(6)
    x = np.empty(6)
    lib().orc_ldlt6_solve(_p(A), _p(b), _p(x))
    return x


def vec6_to_mat4(x):
    x = _f64(x).reshape(6)
    T = np.empty(16)
    lib().orc_vec6_to_mat4(_p(x), _p(T))
    return T.reshape(4, 4)


def transform(T, xyz, nrm=None):
    T = _f64(T).reshape(4, 4)
    xyz = _f64(xyz).reshape(-1, 3)
    n = len(xyz)
    nrm = None if nrm is None else _f64(nrm).reshape(-1, 3)
    ox = np.empty((2 * n, 3)); on = np.empty((2 * n, 3)) if nrm is not None else None
    m = lib().orc_transform(_p(T), _p(xyz), _p(nrm), C.c_size_t(n), _p(ox), _p(on))
    return (ox[:m].copy(), None if on is None else on[:m].copy())


def voxelize_within_cropping_volume(voxel, c: Cropper, xyz, nrm=None, return_keys=False):
    xyz = _f64(xyz).reshape(-1, 3)
    n = len(xyz)
    nrm = None if nrm is None else _f64(nrm).reshape(-1, 3)
    ox = np.empty((n, 3)); on = np.empty((n, 3)) if nrm is not None else None
    keys = np.empty((n, 3), dtype=np.int32)
    m = lib().orc_voxelize_within_cropping_volume(C.c_double(voxel), C.byref(c), _p(xyz), _p(nrm), C.c_size_t(n), _p(ox), _p(on),
                                                  _p(keys))
    out = (ox[:m].copy(), None if on is None else on[:m].copy())
    return out + (keys[:m].copy(),) if return_keys else out


def submap_insert_scan(map_xyz, map_nrm, scan_xyz, scan_nrm, T, map_voxel, c: Cropper, return_keys=False):
    map_xyz = _f64(map_xyz).reshape(-1, 3); map_nrm = _f64(map_nrm).reshape(-1, 3)
    scan_xyz = _f64(scan_xyz).reshape(-1, 3); scan_nrm = _f64(scan_nrm).reshape(-1, 3)
    T = _f64(T).reshape(4, 4)
    cap = len(map_xyz) + 2 * len(scan_xyz) + 1
    ox = np.empty((cap, 3)); on = np.empty((cap, 3)); keys = np.empty((cap, 3), dtype=np.int32)
    m = lib().orc_submap_insert_scan(_p(map_xyz), _p(map_nrm), C.c_size_t(len(map_xyz)), _p(scan_xyz), _p(scan_nrm),
                                     C.c_size_t(len(scan_xyz)), _p(T), C.c_double(map_voxel), C.byref(c), _p(ox), _p(on), _p(keys))
    out = (ox[:m].copy(), on[:m].copy())
    return out + (keys[:m].copy(),) if return_keys else out


class DenseMap:
    """VoxelizedPointCloud restatement (core/src/Voxel.cpp:18-115)."""

    def __init__(self, voxel, max_voxels=1 << 20):
        self._h = C.c_void_p(lib().orc_dense_create(C.c_double(voxel), C.c_size_t(max_voxels)))
        self._cap = max_voxels

    def insert(self, xyz, nrm=None):
        xyz = _f64(xyz).reshape(-1, 3)
        nrm = None if nrm is None else _f64(nrm).reshape(-1, 3)
        if lib().orc_dense_insert(self._h, _p(xyz), _p(nrm), C.c_size_t(len(xyz))) != 0:
            raise RuntimeError("dense map capacity exceeded")

    def transform(self, T) -> None:
        """VoxelizedPointCloud::transform (core/src/Voxel.cpp:49-64)."""
        T = _f64(T).reshape(4, 4)
        lib().orc_dense_transform(self._h, _p(T))

    def carve(self, scan, sensor, voxel, radius=0.1, truncation=0.1, max_len=20.0) -> int:
        """Submap::carve on the dense map (core/src/Submap.cpp:125-136): returns the number of removed voxels."""
        scan = _f64(scan).reshape(-1, 3); s = _f64(sensor).reshape(3)
        lib().orc_dense_carve.restype = C.c_size_t
        return int(lib().orc_dense_carve(self._h, _p(scan), C.c_size_t(len(scan)), _p(s), C.c_double(voxel), C.c_double(radius),
                                         C.c_double(truncation), C.c_double(max_len)))

    def to_cloud(self):
        ox = np.empty((self._cap, 3)); on = np.empty((self._cap, 3)); keys = np.empty((self._cap, 3), dtype=np.int32)
        m = lib().orc_dense_to_cloud(self._h, _p(ox), _p(on), _p(keys))
        return ox[:m].copy(), on[:m].copy(), keys[:m].copy()

    def __del__(self):
        try:
            lib().orc_dense_destroy(self._h)
        except Exception:
            pass


def process_scan(raw_xyz, map_builder_cropper: Cropper, scan_matcher_cropper: Cropper, voxel, knn, knn_radius, ratio, seed):
    raw = _f64(raw_xyz).reshape(-1, 3)
    n = len(raw)
    mx = np.empty((n, 3)); mn = np.empty((n, 3)); ax = np.empty((n, 3)); an = np.empty((n, 3))
    nm = C.c_size_t(); na = C.c_size_t()
    rc = lib().orc_process_scan(_p(raw), C.c_size_t(n), C.byref(map_builder_cropper), C.byref(scan_matcher_cropper),
                                C.c_double(voxel), C.c_int(knn), C.c_double(knn_radius), C.c_double(ratio), C.c_uint32(seed),
                                _p(mx), _p(mn), C.byref(nm), _p(ax), _p(an), C.byref(na))
    if rc != 0:
        raise RuntimeError("ScanToMapIcp: cropped size is zero")  # core/src/ScanToMapRegistration.cpp:51-52
    return (mx[:nm.value].copy(), mn[:nm.value].copy()), (ax[:na.value].copy(), an[:na.value].copy())


def kdtree_search_hybrid(pts, q, radius, max_nn):
    pts = _f64(pts).reshape(-1, 3)
    t = C.c_void_p(lib().orc_kdtree_build(_p(pts), C.c_int(len(pts))))
    q = _f64(q).reshape(-1, 3)
    out = []
    d2 = np.empty(max_nn); idx = np.empty(max_nn, dtype=np.int32)
    for qi in q:
        qi = np.ascontiguousarray(qi)
        k = lib().orc_kdtree_search_hybrid(t, _p(qi), C.c_double(radius), C.c_int(max_nn), _p(d2), _p(idx))
        out.append((idx[:k].copy(), d2[:k].copy()))
    lib().orc_kdtree_free(t)
    return out


class CarvingParams(C.Structure):
    """SpaceCarvingParameters (core/include/open3d_slam/Parameters.hpp:85-92), the fields getIdxsOfCarvedPoints reads."""
    _fields_ = [("voxel_size", C.c_double), ("max_raytracing_length", C.c_double), ("truncation_distance", C.c_double),
                ("min_dot_product_with_normal", C.c_double)]


def carve(map_xyz, map_nrm, scan_xyz_map_frame, sensor, cropper_: Cropper, voxel_size=0.1, max_raytracing_length=20.0,
          truncation_distance=0.1, min_dot_product_with_normal=0.5):
    """Boolean mask of the map points space carving removes (Submap::carve -> getIdxsOfCarvedPoints)."""
    map_xyz = _f64(map_xyz).reshape(-1, 3)
    map_nrm = None if map_nrm is None else _f64(map_nrm).reshape(-1, 3)
    scan = _f64(scan_xyz_map_frame).reshape(-1, 3)
    s = _f64(sensor).reshape(3)
    prm = CarvingParams(voxel_size, max_raytracing_length, truncation_distance, min_dot_product_with_normal)
    removed = np.zeros(len(map_xyz), dtype=np.uint8)
    lib().orc_carve.restype = C.c_size_t
    lib().orc_carve(_p(map_xyz), _p(map_nrm), C.c_size_t(len(map_xyz)), _p(scan), C.c_size_t(len(scan)), _p(s), C.byref(cropper_),
                    C.byref(prm), removed.ctypes.data_as(C.c_void_p))
    return removed.astype(bool)


def overlap_flags(src, tgt, T, voxel, min_pts=1):
    """computeIndicesOfOverlappingPoints (core/src/helpers.cpp:307-332) as boolean masks over source and target."""
    src = _f64(src).reshape(-1, 3); tgt = _f64(tgt).reshape(-1, 3); T = _f64(T).reshape(4, 4)
    fs = np.zeros(len(src), dtype=np.uint8); ft = np.zeros(len(tgt), dtype=np.uint8)
    lib().orc_overlap_flags(_p(src), C.c_size_t(len(src)), _p(tgt), C.c_size_t(len(tgt)), _p(T), C.c_double(voxel), C.c_size_t(min_pts),
                            fs.ctypes.data_as(C.c_void_p), ft.ctypes.data_as(C.c_void_p))
    return fs.astype(bool), ft.astype(bool)


def information_matrix(src, tgt, max_corr_dist, T):
    """[O3D] GetInformationMatrixFromPointClouds (called at core/src/PlaceRecognition.cpp:148)."""
    src = _f64(src).reshape(-1, 3); tgt = _f64(tgt).reshape(-1, 3); T = _f64(T).reshape(4, 4)
    G = np.zeros((6, 6))
    if lib().orc_information_matrix(_p(src), C.c_size_t(len(src)), _p(tgt), C.c_size_t(len(tgt)), C.c_double(max_corr_dist), _p(T), _p(G)) != 0:
        raise RuntimeError("invalid max_correspondence_distance")
    return G


def undistort(xyz, linear_velocity, angular_velocity_rpy, scan_duration=0.1, spinning_clockwise=True):
    """ConstantVelocityMotionCompensation::undistortInputPointCloud (core/src/MotionCompensation.cpp:64-110) for given velocities."""
    xyz = _f64(xyz).reshape(-1, 3)
    lv = _f64(linear_velocity).reshape(3); av = _f64(angular_velocity_rpy).reshape(3)
    out = np.empty_like(xyz)
    lib().orc_undistort(_p(xyz), C.c_size_t(len(xyz)), _p(lv), _p(av), C.c_double(scan_duration), C.c_int(int(spinning_clockwise)), _p(out))
    return out


def pointcloud_transform(T, xyz, nrm=None):
    """[O3D] PointCloud::Transform (points and normals, no duplication quirk): what Submap::transform applies to mapCloud_."""
    T = _f64(T).reshape(4, 4); x = _f64(xyz).reshape(-1, 3).copy(); n = None if nrm is None else _f64(nrm).reshape(-1, 3).copy()
    lib().orc_pointcloud_transform(_p(T), _p(x), _p(n), C.c_size_t(len(x)))
    return x, n


def num_threads() -> int:
    return int(lib().orc_num_threads())

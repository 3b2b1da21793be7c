"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "object-oriented-slam_b200"))
from eslam_b200._structs import (EslamCamera, EslamEdgeBatch, EslamEllipsoidOut, EslamLmStats, EslamParams,  # noqa: E402
                                 EslamStageDump, default_camera, default_params)

_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_symmetry_error.restype = C.c_double
        _LIB.orc_estimate_ellipsoids.restype = C.c_longlong
        _LIB.orc_now_seconds.restype = C.c_double
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def se3_exp(v6):
    out = np.zeros(7); v = _d(v6); lib().orc_se3_exp(_dp(v), _dp(out)); return out


def se3_log(T7):
    out = np.zeros(6); v = _d(T7); lib().orc_se3_log(_dp(v), _dp(out)); return out


def se3_mul(a, b):
    out = np.zeros(7); a, b = _d(a), _d(b); lib().orc_se3_mul(_dp(a), _dp(b), _dp(out)); return out


def se3_inverse(a):
    out = np.zeros(7); a = _d(a); lib().orc_se3_inverse(_dp(a), _dp(out)); return out


def ellipsoid_rotate(e10, yaw):
    out = np.zeros(10); e = _d(e10); lib().orc_ellipsoid_rotate(_dp(e), C.c_double(yaw), _dp(out)); return out


def ellipsoid_exp_update(e10, u9):
    out = np.zeros(10); e, u = _d(e10), _d(u9); lib().orc_ellipsoid_exp_update(_dp(e), _dp(u), _dp(out)); return out


def ellipsoid_transform_from(e10, T7):
    out = np.zeros(10); e, t = _d(e10), _d(T7); lib().orc_ellipsoid_transform_from(_dp(e), _dp(t), _dp(out)); return out


def ellipsoid_from_minimal(v9):
    out = np.zeros(10); v = _d(v9); lib().orc_ellipsoid_from_minimal(_dp(v), _dp(out)); return out


def min_log_error_9dof(a10, b10):
    out = np.zeros(9); a, b = _d(a10), _d(b10); lib().orc_min_log_error_9dof(_dp(a), _dp(b), _dp(out)); return out


def ellipsoid_bbox(e10, Tcw7, K):
    out = np.zeros(4); e, t, k = _d(e10), _d(Tcw7), _d(K).reshape(-1); lib().orc_ellipsoid_bbox(_dp(e), _dp(t), _dp(k), _dp(out)); return out


def eigen3(A):
    A = _d(A).reshape(-1); ev = np.zeros(3); V = np.zeros(9); lib().orc_eigen3(_dp(A), _dp(ev), _dp(V)); return ev, V.reshape(3, 3)


def ldlt_solve(A, b):
    A = _d(A); b = _d(b); x = np.zeros_like(b); ok = lib().orc_ldlt_solve(C.c_int(len(b)), _dp(A.reshape(-1)), _dp(b), _dp(x)); return bool(ok), x


def plane_to_world(pl, Twc7):
    out = np.zeros(4); p, t = _d(pl), _d(Twc7); lib().orc_plane_to_world(_dp(p), _dp(t), _dp(out)); return out


def edge_error(etype, Tcw7, e10, meas, K):
    out = np.zeros(9); t, e, m, k = _d(Tcw7), _d(e10), _d(np.resize(meas, 10)), _d(K).reshape(-1)
    lib().orc_edge_error(C.c_int(etype), _dp(t), _dp(e), _dp(m), _dp(k), _dp(out)); return out


class Graph:
    """Keeps numpy arrays alive behind an EslamEdgeBatch."""

    def __init__(self, etype, frame, obj, meas, info, Tcw, n_objects, K):
        self.type = np.ascontiguousarray(etype, dtype=np.int32)
        self.frame = np.ascontiguousarray(frame, dtype=np.int32)
        self.object = np.ascontiguousarray(obj, dtype=np.int32)
        self.meas = _d(meas).reshape(-1, 10)
        self.info = _d(info)
        self.Tcw = _d(Tcw).reshape(-1, 7)
        self.n_objects = int(n_objects)
        self.K = _d(K).reshape(-1)
        b = EslamEdgeBatch()
        b.n_edges = len(self.type)
        b.type, b.frame, b.object = _ip(self.type), _ip(self.frame), _ip(self.object)
        b.meas, b.info = _dp(self.meas), _dp(self.info)
        b.n_frames = len(self.Tcw)
        b.Tcw = _dp(self.Tcw)
        b.n_objects = self.n_objects
        for i in range(9):
            b.K[i] = self.K[i]
        self.c = b


def edges_eval(g, ellipsoids, jac=True):
    e = _d(ellipsoids); n = g.c.n_edges
    r = np.zeros((n, 9)); J = np.zeros((n, 9, 9)); chi = np.zeros(n)
    lib().orc_edges_eval(C.byref(g.c), _dp(e), _dp(r), _dp(J) if jac else None, _dp(chi))
    return r, J, chi


def optimize_objects(g, ellipsoids, iters=10):
    e = _d(ellipsoids).copy(); st = EslamLmStats()
    lib().orc_optimize_objects(C.byref(g.c), C.c_int(iters), _dp(e), C.byref(st))
    return e, st


def optimize_slam(g, ellipsoids, iters=10, odom_meas=None, odom_info=1.0):
    """SLAM mode (free poses, odometry edges); returns (ellipsoids, Tcw, stats)."""
    e = _d(ellipsoids).copy(); st = EslamLmStats(); T = np.zeros((g.c.n_frames, 7))
    om = _d(odom_meas).reshape(-1, 7) if odom_meas is not None else None
    lib().orc_optimize_slam(C.byref(g.c), _dp(om) if om is not None else None, C.c_double(odom_info), C.c_int(iters), _dp(e), _dp(T), C.byref(st))
    return e, T, st


def odom_error(meas7, v1, v2):
    out = np.zeros(6); lib().orc_odom_error(_dp(_d(meas7)), _dp(_d(v1)), _dp(_d(v2)), _dp(out)); return out


class Dump:
    def __init__(self, cap):
        self.cap = cap
        self.crop = np.zeros((cap, 3)); self.voxel = np.zeros((cap, 3)); self.world = np.zeros((cap, 3))
        self.label = np.full(cap, -2, dtype=np.int32); self.sym = np.zeros((cap, 3)); self.final = np.zeros((cap, 3))

    def fill(self, d):
        d.cap_points = self.cap
        d.crop_xyz, d.voxel_xyz, d.world_xyz = _dp(self.crop), _dp(self.voxel), _dp(self.world)
        d.cluster_label = _ip(self.label)
        d.sym_xyz, d.final_xyz = _dp(self.sym), _dp(self.final)


def estimate_ellipsoids(depth, frame_of_obj, bbox, label, Twc, plane_world, prm=None, cam=None, dump_cap=0, unstable=0):
    prm = prm or default_params(); cam = cam or default_camera()
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    if depth.ndim == 2:
        depth = depth[None]
    fo = np.ascontiguousarray(frame_of_obj, dtype=np.int32); bb = _d(bbox).reshape(-1, 4)
    lb = np.ascontiguousarray(label, dtype=np.int32); tw = _d(Twc).reshape(-1, 7); pw = _d(plane_world)
    n = len(fo)
    out = (EslamEllipsoidOut * n)()
    dumps, keep = None, []
    if dump_cap:
        dumps = (EslamStageDump * n)()
        for i in range(n):
            k = Dump(dump_cap); k.fill(dumps[i]); keep.append(k)
    evals = lib().orc_estimate_ellipsoids(C.byref(prm), C.byref(cam), _dp(pw), C.c_int(unstable), C.c_int(depth.shape[0]),
                                          depth.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_size_t(depth.shape[2]), C.c_int(n), _ip(fo), _dp(bb), _ip(lb),
                                          _dp(tw), out, dumps)
    return out, dumps, keep, evals


def extract_ground_plane(depth, prm=None, cam=None):
    prm = prm or default_params(); cam = cam or default_camera()
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    if depth.ndim == 2:
        depth = depth[None]
    F = depth.shape[0]
    pl = np.zeros((F, 4)); found = np.zeros(F, dtype=np.int32); ninl = np.zeros(F, dtype=np.int32)
    lib().orc_extract_ground_plane(C.byref(prm), C.byref(cam), C.c_int(F), depth.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_size_t(depth.shape[2]),
                                   _dp(pl), _ip(found), _ip(ninl))
    return pl, found, ninl


def extract_planes_debug(depth, prm=None, cam=None, cap=64):
    prm = prm or default_params(); cam = cam or default_camera()
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    H, W = depth.shape
    normals = np.zeros((H, W, 3), dtype=np.float32); dm = np.zeros((H, W), dtype=np.float32); labels = np.zeros((H, W), dtype=np.int32)
    coef = np.zeros((cap, 4), dtype=np.float32); size = np.zeros(cap, dtype=np.int32); first = np.zeros(cap, dtype=np.int32)
    n = lib().orc_extract_planes_debug(C.byref(prm), C.byref(cam), depth.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_size_t(W),
                                       normals.ctypes.data_as(C.POINTER(C.c_float)), dm.ctypes.data_as(C.POINTER(C.c_float)), _ip(labels),
                                       coef.ctypes.data_as(C.POINTER(C.c_float)), _ip(size), _ip(first), C.c_int(cap))
    n = min(n, cap)
    return normals, dm, labels, coef[:n], size[:n], first[:n]


def voxel_grid(xyz_f32, leaf, unstable=0):
    x = np.ascontiguousarray(xyz_f32, dtype=np.float32); n = len(x); out = np.zeros((n, 3), dtype=np.float32)
    m = lib().orc_voxel_grid(x.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(n), C.c_double(leaf), C.c_int(unstable), out.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(n))
    return out[:m]


def euclidean_clusters(xyz_f32, tol, min_size):
    x = np.ascontiguousarray(xyz_f32, dtype=np.float32); n = len(x); lab = np.zeros(n, dtype=np.int32)
    k = lib().orc_euclidean_clusters(x.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(n), C.c_double(tol), C.c_int(min_size), _ip(lab))
    return k, lab


def symmetry_error(depth, bbox, T_oc7, cloud, plane4, dual_dis=0.0, sym_type=1, prm=None, cam=None):
    prm = prm or default_params(); cam = cam or default_camera()
    depth = np.ascontiguousarray(depth, dtype=np.uint16); cl = _d(cloud).reshape(-1, 3)
    bb, t, p = _d(bbox), _d(T_oc7), _d(plane4)
    return lib().orc_symmetry_error(C.byref(prm), C.byref(cam), depth.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_size_t(depth.shape[1]), _dp(bb), _dp(t),
                                    _dp(cl), C.c_int(len(cl)), _dp(p), C.c_double(dual_dis), C.c_int(sym_type))


def init_quadric(Twc, det5, K, rows=480, cols=640):
    tw, d, k = _d(Twc).reshape(-1, 7), _d(det5).reshape(-1, 5), _d(K).reshape(-1)
    out = np.zeros(10); ok = C.c_int(0)
    lib().orc_init_quadric(C.c_int(len(tw)), _dp(tw), _dp(d), _dp(k), C.c_int(rows), C.c_int(cols), _dp(out), C.byref(ok))
    return out, bool(ok.value)


def calibrate_measurement(bbox, prm=None, cam=None):
    prm = prm or default_params(); cam = cam or default_camera(); b = _d(bbox)
    return lib().orc_calibrate_measurement(C.byref(prm), C.byref(cam), _dp(b))

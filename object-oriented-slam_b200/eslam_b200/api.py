"""ctypes binding + host-side mirror of the reference's operator interface for the hot path.

Reference interface being mirrored (names, argument meaning, error behaviour):
  Frame                      include/core/Frame.h:12-41, src/core/Frame.cpp:7-27
  PlaneExtractor             src/plane/PlaneExtractor.h:41-70 (extractGroundPlane -> bool + g2o::plane)
  EllipsoidExtractor         src/pca/EllipsoidExtractor.h:37-131 (EstimateLocalEllipsoid, GetResult, GetSymmetryOutputData, SetSupportingPlane, OpenSymmetry)
  Initializer                include/core/Initializer.h:36-80 (initializeQuadric, getInitializeResult)
  Optimizer                  include/core/Optimizer.h:14-32 (GlobalObjectGraphOptimization, SetGroundPlane)
All compute goes through libeslam_b200.so; nothing here falls back to numpy for a hot-path stage.
"""
import ctypes as C
import os

import numpy as np

from ._structs import (EDGE_9DOF, EDGE_GRAVITY, EDGE_PROJ, EslamCamera, EslamEdgeBatch, EslamEllipsoidOut, EslamLmStats, EslamParams,
                       EslamStageDump, default_camera, default_params)

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "libeslam_b200.so")
_LIB = None

ABI_SYMBOLS = [
    "eslam_abi_version", "eslam_default_params", "eslam_default_camera", "eslam_create", "eslam_destroy", "eslam_last_error",
    "eslam_launch_count", "eslam_h2d_bytes", "eslam_extract_ground_plane_host", "eslam_extract_ground_plane_dev", "eslam_plane_to_world",
    "eslam_set_supporting_plane", "eslam_calibrate_measurement", "eslam_estimate_ellipsoids_host", "eslam_estimate_ellipsoids_dev",
    "eslam_fetch_ellipsoids", "eslam_estimate_ellipsoids_dev_submit", "eslam_estimate_ellipsoids_dev_collect", "eslam_copy_ellipsoids_dev", "eslam_batch_counters", "eslam_estimate_ellipsoids_submit", "eslam_estimate_ellipsoids_wait", "eslam_init_quadric", "eslam_edges_eval_host", "eslam_edges_upload", "eslam_edges_eval_dev",
    "eslam_edges_download", "eslam_optimize_objects_host", "eslam_optimize_slam_host", "eslam_pose_inverse", "eslam_pose_mul", "eslam_set_profiling", "eslam_get_stage_ms", "eslam_associate_frames_host",
]


class EslamError(RuntimeError):
    pass


def load_library(path=None):
    """Loads libeslam_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _LIB
    if _LIB is None:
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise EslamError(f"{p} not found: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        lib = C.CDLL(p)
        lib.eslam_last_error.restype = C.c_char_p
        lib.eslam_launch_count.restype = C.c_longlong
        lib.eslam_h2d_bytes.restype = C.c_ulonglong
        for n in ("eslam_sizeof_params", "eslam_sizeof_camera", "eslam_sizeof_ellipsoid_out", "eslam_sizeof_stage_dump", "eslam_sizeof_edge_batch", "eslam_sizeof_lm_stats"):
            getattr(lib, n).restype = C.c_size_t
        lib.eslam_plane_to_world.restype = None
        lib.eslam_pose_inverse.restype = None
        lib.eslam_pose_mul.restype = None
        _LIB = lib
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _u16p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


def pose_inverse(T7):
    out = np.zeros(7); t = _d(T7); load_library().eslam_pose_inverse(_dp(t), _dp(out)); return out


def pose_mul(A7, B7):
    """g2o::SE3Quat::operator* (se3quat.h:110-116)."""
    out = np.zeros(7); load_library().eslam_pose_mul(_dp(_d(A7)), _dp(_d(B7)), _dp(out)); return out


def plane_to_world(plane_cam, Twc7):
    out = np.zeros(4); p, t = _d(plane_cam), _d(Twc7); load_library().eslam_plane_to_world(_dp(p), _dp(t), _dp(out)); return out


def calibrate_measurement(bbox, prm=None, cam=None):
    prm = prm or default_params(); cam = cam or default_camera(); b = _d(bbox)
    return bool(load_library().eslam_calibrate_measurement(C.byref(prm), C.byref(cam), _dp(b)))


class Handle:
    """Owns an EslamHandle (device workspace + stream) -- the state the reference keeps in its extractor/optimizer objects."""

    def __init__(self, params=None, camera=None, device=0):
        self.lib = load_library()
        self.params = params or default_params()
        self.camera = camera or default_camera()
        self._h = C.c_void_p()
        rc = self.lib.eslam_create(C.byref(self.params), C.byref(self.camera), C.c_int(device), C.byref(self._h))
        if rc != 0:
            raise EslamError(f"eslam_create failed ({rc}): a CUDA device is required, there is no CPU fallback")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.eslam_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what):
        if rc != 0:
            raise EslamError(f"{what} failed ({rc}): {self.lib.eslam_last_error(self._h).decode()}")

    def set_profiling(self, on=True):
        self.check(self.lib.eslam_set_profiling(self._h, C.c_int(1 if on else 0)), "eslam_set_profiling")

    def stage_ms(self):
        ms = (C.c_float * 3)()
        self.check(self.lib.eslam_get_stage_ms(self._h, ms), "eslam_get_stage_ms")
        return [float(x) for x in ms]

    @property
    def launches(self):
        return int(self.lib.eslam_launch_count(self._h))

    @property
    def h2d_bytes(self):
        return int(self.lib.eslam_h2d_bytes(self._h))

    # ---- batched entry points (host buffers) ----
    def set_supporting_plane(self, plane_world):
        p = _d(plane_world)
        self.check(self.lib.eslam_set_supporting_plane(self._h, _dp(p)), "eslam_set_supporting_plane")

    def extract_ground_plane(self, depth):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        if depth.ndim == 2:
            depth = depth[None]
        F = depth.shape[0]
        pl = np.zeros((F, 4)); found = np.zeros(F, dtype=np.int32); ninl = np.zeros(F, dtype=np.int32)
        self.check(self.lib.eslam_extract_ground_plane_host(self._h, C.c_int(F), _u16p(depth), C.c_size_t(depth.shape[2]), _dp(pl), _ip(found), _ip(ninl)),
                   "eslam_extract_ground_plane_host")
        return pl, found, ninl

    def extract_ground_plane_dev(self, depth_ptr, n_frames, pitch, stream=0):
        pl = np.zeros((n_frames, 4)); found = np.zeros(n_frames, dtype=np.int32); ninl = np.zeros(n_frames, dtype=np.int32)
        self.check(self.lib.eslam_extract_ground_plane_dev(self._h, C.c_void_p(stream), C.c_int(n_frames), C.c_void_p(depth_ptr), C.c_size_t(pitch), _dp(pl),
                                                           _ip(found), _ip(ninl)), "eslam_extract_ground_plane_dev")
        return pl, found, ninl

    def debug_counters(self, obj):
        out = (C.c_ulonglong * 5)()
        self.check(self.lib.eslam_debug_counters(self._h, C.c_int(obj), out), "eslam_debug_counters")
        return [int(v) for v in out]

    def debug_plane_stage(self, depth, stage):
        """test hook: 1 distance map (float), 2 normals+plane_d (float4), 3 labels after CCL, 4 labels after refinement."""
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        H, W = depth.shape
        out = {1: np.zeros((H, W), np.float32), 2: np.zeros((H, W, 4), np.float32), 3: np.zeros((H, W), np.int32), 4: np.zeros((H, W), np.int32)}[stage]
        self.check(self.lib.eslam_debug_plane_stage(self._h, _u16p(depth), C.c_size_t(W), C.c_int(stage), out.ctypes.data_as(C.c_void_p)), "eslam_debug_plane_stage")
        return out

    def estimate_ellipsoids(self, depth, frame_of_obj, bbox, label, Twc, dump_cap=0):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        if depth.ndim == 2:
            depth = depth[None]
        fo = np.ascontiguousarray(frame_of_obj, dtype=np.int32); bb = _d(bbox).reshape(-1, 4)
        lb = np.ascontiguousarray(label, dtype=np.int32); tw = _d(Twc).reshape(-1, 7)
        n = len(fo)
        out = (EslamEllipsoidOut * n)()
        dumps, keep = None, []
        if dump_cap:
            dumps = (EslamStageDump * n)()
            for i in range(n):
                k = _DumpArrays(dump_cap); k.fill(dumps[i]); keep.append(k)
        self.check(self.lib.eslam_estimate_ellipsoids_host(self._h, C.c_int(depth.shape[0]), _u16p(depth), C.c_size_t(depth.shape[2]), C.c_int(n), _ip(fo), _dp(bb),
                                                           _ip(lb), _dp(tw), out, dumps), "eslam_estimate_ellipsoids_host")
        return out, dumps, keep

    def submit_ellipsoids(self, depth, frame_of_obj, bbox, label, Twc):
        """Pipelined form of estimate_ellipsoids (at most two batches in flight): returns a ticket for wait_ellipsoids.  The arrays
        are kept alive inside the ticket until the wait."""
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        if depth.ndim == 2:
            depth = depth[None]
        fo = np.ascontiguousarray(frame_of_obj, dtype=np.int32); bb = _d(bbox).reshape(-1, 4)
        lb = np.ascontiguousarray(label, dtype=np.int32); tw = _d(Twc).reshape(-1, 7)
        t = C.c_int(-1)
        self.check(self.lib.eslam_estimate_ellipsoids_submit(self._h, C.c_int(depth.shape[0]), _u16p(depth), C.c_size_t(depth.shape[2]), C.c_int(len(fo)), _ip(fo), _dp(bb),
                                                             _ip(lb), _dp(tw), C.byref(t)), "eslam_estimate_ellipsoids_submit")
        return (t.value, len(fo), (depth, fo, bb, lb, tw))

    def wait_ellipsoids(self, ticket):
        slot, n, _keep = ticket
        out = (EslamEllipsoidOut * n)()
        self.check(self.lib.eslam_estimate_ellipsoids_wait(self._h, C.c_int(slot), C.c_int(n), out), "eslam_estimate_ellipsoids_wait")
        return out

    def estimate_ellipsoids_dev(self, depth_ptr, n_frames, pitch, frame_of_obj, bbox, label, Twc, stream=0):
        fo = np.ascontiguousarray(frame_of_obj, dtype=np.int32); bb = _d(bbox).reshape(-1, 4)
        lb = np.ascontiguousarray(label, dtype=np.int32); tw = _d(Twc).reshape(-1, 7)
        self.check(self.lib.eslam_estimate_ellipsoids_dev(self._h, C.c_void_p(stream), C.c_int(n_frames), C.c_void_p(depth_ptr), C.c_size_t(pitch), C.c_int(len(fo)),
                                                          _ip(fo), _dp(bb), _ip(lb), _dp(tw)), "eslam_estimate_ellipsoids_dev")
        return len(fo)

    def fetch_ellipsoids(self, n_obj, stream=0):
        out = (EslamEllipsoidOut * n_obj)()
        self.check(self.lib.eslam_fetch_ellipsoids(self._h, C.c_void_p(stream), C.c_int(n_obj), out), "eslam_fetch_ellipsoids")
        return out

    def dev_submit(self, depth_ptr, n_frames, pitch, frame_of_obj, bbox, label, Twc, stream=0):
        """Pipelined device-resident form (two batches in flight): returns a ticket for dev_collect."""
        fo = np.ascontiguousarray(frame_of_obj, dtype=np.int32); bb = _d(bbox).reshape(-1, 4)
        lb = np.ascontiguousarray(label, dtype=np.int32); tw = _d(Twc).reshape(-1, 7)
        t = C.c_int(-1)
        self.check(self.lib.eslam_estimate_ellipsoids_dev_submit(self._h, C.c_void_p(stream), C.c_int(n_frames), C.c_void_p(depth_ptr), C.c_size_t(pitch), C.c_int(len(fo)),
                                                                 _ip(fo), _dp(bb), _ip(lb), _dp(tw), C.byref(t)), "eslam_estimate_ellipsoids_dev_submit")
        return (t.value, len(fo))

    def dev_collect(self, ticket, dst_ptr, stream=0):
        """`stream` waits for the ticket's batch, then its records are copied (D2D, asynchronous) to dst_ptr."""
        slot, n = ticket
        self.check(self.lib.eslam_estimate_ellipsoids_dev_collect(self._h, C.c_void_p(stream), C.c_int(slot), C.c_int(n), C.c_void_p(dst_ptr)),
                   "eslam_estimate_ellipsoids_dev_collect")

    def copy_ellipsoids_dev(self, n_obj, dst_ptr, stream=0):
        """records of the last *_dev batch -> caller-owned device memory (D2D on `stream`, asynchronous)."""
        self.check(self.lib.eslam_copy_ellipsoids_dev(self._h, C.c_void_p(stream), C.c_int(n_obj), C.c_void_p(dst_ptr)), "eslam_copy_ellipsoids_dev")

    def batch_counters(self):
        out = (C.c_ulonglong * 12)()
        self.check(self.lib.eslam_batch_counters(self._h, out), "eslam_batch_counters")
        keys = ("ref_evals", "ref_points", "evals", "points", "nn_scans", "nn_hits", "max_n_sym", "objects_with_symmetry", "verify_failures",
                "clk_scalar_phase", "clk_task_phase", "clk_cta_total")   # round pipeline: FD passes, FD fallbacks, -
        return {k: int(v) for k, v in zip(keys, out)}

    def associate_frames(self, det_start, Twc, det_pose7_cam, valid, map_center3, first_new_instance=0):
        """Batched DataAssociationSolver::Solve (src/core/DataAssociation.cpp:16-135) of n_frames frames against one map snapshot on the device.
        Returns (assoc [n_det] int32, n_new)."""
        ds = np.ascontiguousarray(det_start, dtype=np.int32); tw = _d(Twc).reshape(-1, 7)
        dp = _d(det_pose7_cam).reshape(-1, 7); va = np.ascontiguousarray(valid, dtype=np.int32); mc = _d(map_center3).reshape(-1, 3)
        n_det = len(dp)
        out = np.zeros(max(1, n_det), dtype=np.int32); n_new = C.c_int(0)
        self.check(self.lib.eslam_associate_frames_host(self._h, C.c_int(len(ds) - 1), _ip(ds), _dp(tw), C.c_int(n_det), _dp(dp) if n_det else None,
                                                        _ip(va) if n_det else None, C.c_int(len(mc)), _dp(mc) if len(mc) else None, C.c_int(first_new_instance),
                                                        _ip(out), C.byref(n_new)), "eslam_associate_frames_host")
        return out[:n_det], int(n_new.value)

    def init_quadric(self, Twc, det5, K, rows=None, cols=None):
        tw, d, k = _d(Twc).reshape(-1, 7), _d(det5).reshape(-1, 5), _d(K).reshape(-1)
        out = np.zeros(10); ok = C.c_int(0)
        self.check(self.lib.eslam_init_quadric(self._h, C.c_int(len(tw)), _dp(tw), _dp(d), _dp(k), C.c_int(rows or self.camera.rows), C.c_int(cols or self.camera.cols),
                                               _dp(out), C.byref(ok)), "eslam_init_quadric")
        return out, bool(ok.value)

    def edges_eval(self, graph, ellipsoids):
        e = _d(ellipsoids); n = graph.c.n_edges
        r = np.zeros((n, 9)); J = np.zeros((n, 9, 9)); chi = np.zeros(n)
        self.check(self.lib.eslam_edges_eval_host(self._h, C.byref(graph.c), _dp(e), _dp(r), _dp(J), _dp(chi)), "eslam_edges_eval_host")
        return r, J, chi

    def edges_upload(self, graph, ellipsoids):
        e = _d(ellipsoids)
        self.check(self.lib.eslam_edges_upload(self._h, C.byref(graph.c), _dp(e)), "eslam_edges_upload")

    def edges_eval_dev(self, stream=0):
        self.check(self.lib.eslam_edges_eval_dev(self._h, C.c_void_p(stream)), "eslam_edges_eval_dev")

    def optimize_objects(self, graph, ellipsoids, iters=10):
        e = _d(ellipsoids).copy(); st = EslamLmStats()
        self.check(self.lib.eslam_optimize_objects_host(self._h, C.byref(graph.c), C.c_int(iters), _dp(e), C.byref(st)), "eslam_optimize_objects_host")
        return e, st


    def optimize_slam(self, graph, ellipsoids, iters=10, odom_meas=None, odom_info=1.0):
        """SLAM mode (Optimizer.cpp:126-159, bSLAM_mode = true): free poses + odometry edges.  Returns (ellipsoids, Tcw, stats)."""
        e = _d(ellipsoids).copy(); st = EslamLmStats(); T = np.zeros((graph.c.n_frames, 7))
        om = _d(odom_meas).reshape(-1, 7) if odom_meas is not None else None
        self.check(self.lib.eslam_optimize_slam_host(self._h, C.byref(graph.c), _dp(om) if om is not None else None, C.c_double(odom_info), C.c_int(iters),
                                                     _dp(e), _dp(T), C.byref(st)), "eslam_optimize_slam_host")
        return e, T, st


class _DumpArrays:
    def __init__(self, cap):
        self.cap = cap
        self.crop = np.zeros((cap, 3)); self.voxel = np.zeros((cap, 3)); self.world = np.zeros((cap, 3))
        self.label = np.full(cap, -2, dtype=np.int32); self.sym = np.zeros((cap, 3)); self.final = np.zeros((cap, 3))

    def fill(self, d):
        d.cap_points = self.cap
        d.crop_xyz, d.voxel_xyz, d.world_xyz = _dp(self.crop), _dp(self.voxel), _dp(self.world)
        d.cluster_label = _ip(self.label)
        d.sym_xyz, d.final_xyz = _dp(self.sym), _dp(self.final)


class EdgeGraph:
    """numpy-backed EslamEdgeBatch (keeps the arrays alive)."""

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


# ----------------------------------------------------------------------------------------------------------------------
# reference-shaped classes
# ----------------------------------------------------------------------------------------------------------------------
class Frame:
    """include/core/Frame.h:12-41.  pose = [timestamp?] x y z qx qy qz qw (the last 7 entries are used, Frame.cpp:13)."""
    total_frame = 0

    def __init__(self, timestamp, pose, bboxMat, imDepth, imRGB=None):
        self.timestamp = timestamp
        self.rgb_img = None if imRGB is None else np.array(imRGB, copy=True)
        self.frame_img = np.array(imDepth, dtype=np.uint16, copy=True)
        pose = _d(pose).reshape(-1)
        self.cam_pose_Twc = pose[-7:].copy()
        self.cam_pose_Tcw = pose_inverse(self.cam_pose_Twc)
        self.frame_seq_id = Frame.total_frame
        Frame.total_frame += 1
        self.mmObservations = _d(bboxMat).reshape(-1, 8) if np.size(bboxMat) else np.zeros((0, 8))
        self.mvbOutlier = [False] * len(self.mmObservations)
        self.mpLocalObjects = []   # list of ellipsoid dicts or None, one per bbox row
        self.mbHaveLocalObject = False


class PlaneExtractor:
    """src/plane/PlaneExtractor.h:41-70.  extractGroundPlane(depth) -> (found: bool, plane: ndarray[4] in the camera frame)."""

    def __init__(self, handle):
        self.h = handle
        self.n_inliers = 0

    def extractGroundPlane(self, depth):
        pl, found, ninl = self.h.extract_ground_plane(depth)
        self.n_inliers = int(ninl[0])
        return bool(found[0]), pl[0].copy()


class EllipsoidExtractor:
    """src/pca/EllipsoidExtractor.h:37-131."""

    def __init__(self, handle):
        self.h = handle
        self.mResult = False
        self.mSymmetryOutputData = None
        self.miSystemState = 0
        self.mbSetPlane = False

    def OpenSymmetry(self):
        if not self.h.params.symmetry_open:
            raise EslamError("symmetry is a creation-time parameter (EslamParams.symmetry_open)")

    def SetSupportingPlane(self, plane_world):
        self.h.set_supporting_plane(plane_world)
        self.mbSetPlane = True

    def EstimateLocalEllipsoids(self, depth, frame_of_obj, bboxes, labels, Twc):
        """Batched form: returns the raw EslamEllipsoidOut array (one per bbox)."""
        out, _, _ = self.h.estimate_ellipsoids(depth, frame_of_obj, bboxes, labels, Twc)
        return out

    def EstimateLocalEllipsoid(self, depth, bbox, label, pose, camera=None):
        """Returns the camera-frame ellipsoid as a dict {vec10, prob, label}; GetResult() tells whether it is valid."""
        out, _, _ = self.h.estimate_ellipsoids(depth, [0], _d(bbox)[None], [int(label)], _d(pose).reshape(-1)[-7:][None])
        o = out[0]
        self.miSystemState = int(o.status)
        self.mResult = o.status == 0
        self.mSymmetryOutputData = dict(result=bool(o.symmetry_result), prob=float(o.prob), symmetryType=int(o.symmetry_type),
                                        planeVec=np.array(o.sym_plane_world), planeVec2=np.array(o.sym_plane2_world), center=np.array(o.sym_center_world))
        return dict(vec=np.array(o.ellipsoid), prob=float(o.prob), miLabel=int(label), miInstanceID=-1)

    def GetResult(self):
        return self.mResult

    def GetSymmetryOutputData(self):
        return self.mSymmetryOutputData


class Observation:
    """include/core/Initializer.h:14-24: label, bbox(4), rate, pFrame."""

    def __init__(self, label, bbox, rate, pFrame):
        self.label, self.bbox, self.rate, self.pFrame = int(label), _d(bbox), float(rate), pFrame


class Initializer:
    """include/core/Initializer.h:36-80."""

    def __init__(self, handle, rows, cols):
        self.h, self.rows, self.cols = handle, rows, cols
        self.mbResult = False

    def initializeQuadric(self, pose_mat_or_obs, detection_mat=None, calib=None):
        if detection_mat is None or isinstance(pose_mat_or_obs, (list, tuple)) and pose_mat_or_obs and isinstance(pose_mat_or_obs[0], Observation):
            obs, calib = pose_mat_or_obs, (calib if calib is not None else detection_mat)
            pose_mat = np.array([o.pFrame.cam_pose_Twc for o in obs])
            detection_mat = np.array([np.r_[o.bbox, o.rate] for o in obs])
        else:
            pose_mat = _d(pose_mat_or_obs).reshape(len(detection_mat), -1)[:, -7:]
        e, ok = self.h.init_quadric(pose_mat, detection_mat, calib, self.rows, self.cols)
        self.mbResult = ok
        return dict(vec=e, prob=1.0, miLabel=-1, miInstanceID=-1)

    def getInitializeResult(self):
        return self.mbResult


class Optimizer:
    """include/core/Optimizer.h:14-32; graph assembly of src/core/Optimizer.cpp:83-317 on the host, LM on the device."""

    def __init__(self, handle):
        self.h = handle
        self.mbGroundPlaneSet = False
        self.mGroundPlaneNormal = None
        self.last_stats = None

    def SetGroundPlane(self, normal4):
        self.mbGroundPlaneSet = True
        self.mGroundPlaneNormal = _d(normal4).copy()

    def build_graph(self, pFrames, ellipsoids_by_instance, K, objectObservations):
        prm = self.h.params
        inst_ids = [i for i in sorted(objectObservations.keys()) if i in ellipsoids_by_instance]   # std::map order, initialised only
        vid = {inst: k for k, inst in enumerate(inst_ids)}
        et, fr, ob, meas, info, is2d = [], [], [], [], [], []
        for inst in inst_ids:
            if self.mbGroundPlaneSet and prm.gravity_prior_open == 1:                      # Optimizer.cpp:183-196
                s = 1 * prm.gravity_prior_scale
                et.append(EDGE_GRAVITY); fr.append(0); ob.append(vid[inst]); meas.append(np.r_[self.mGroundPlaneNormal, np.zeros(6)]); info.append(s * s); is2d.append(False)
            obs = objectObservations[inst]
            if len(obs) > 2:                                                                 # :201-245
                for o in obs:
                    et.append(EDGE_PROJ); fr.append(o.pFrame.frame_seq_id); ob.append(vid[inst]); meas.append(np.r_[o.bbox, np.zeros(6)]); info.append((1.0 * 1.0) * o.rate); is2d.append(True)
        for fi, f in enumerate(pFrames):                                                     # :249-279
            for obj in f.mpLocalObjects:
                if obj is None or obj["miInstanceID"] not in vid:
                    continue
                s = 1.0 * np.sqrt(obj["prob"])
                et.append(EDGE_9DOF); fr.append(fi); ob.append(vid[obj["miInstanceID"]]); meas.append(obj["vec"]); info.append((s * s) * prm.edge3d_scale); is2d.append(False)
        Tcw = np.array([f.cam_pose_Tcw for f in pFrames])
        ell = np.array([ellipsoids_by_instance[i]["vec"] for i in inst_ids])
        return inst_ids, et, fr, ob, meas, info, is2d, Tcw, ell

    def GlobalObjectGraphOptimization(self, pFrames, ellipsoids_by_instance, rows, cols, K, objectObservations, iters=10, bSLAM_mode=False):
        inst_ids, et, fr, ob, meas, info, is2d, Tcw, ell = self.build_graph(pFrames, ellipsoids_by_instance, K, objectObservations)
        if not inst_ids or not et:
            return ellipsoids_by_instance
        g = EdgeGraph(et, fr, ob, np.array(meas), info, Tcw, len(inst_ids), K)
        if any(is2d):  # NaN check of the 2-D edges before insertion (Optimizer.cpp:234-243)
            _, _, chi = self.h.edges_eval(g, ell)
            keep = [not (d2 and np.isnan(c)) for d2, c in zip(is2d, chi)]
            if not all(keep):
                sel = np.nonzero(keep)[0]
                g = EdgeGraph(np.array(et)[sel], np.array(fr)[sel], np.array(ob)[sel], np.array(meas)[sel], np.array(info)[sel], Tcw, len(inst_ids), K)
        if bSLAM_mode:   # Optimizer.cpp:139-158: odometry measurement from the frames' Twc, information = I
            odo = []
            for k in range(1, len(pFrames)):
                prev_Tcw = pose_inverse(pFrames[k - 1].cam_pose_Twc); curr_Tcw = pose_inverse(pFrames[k].cam_pose_Twc)
                odo.append(pose_mul(curr_Tcw, pose_inverse(prev_Tcw)))
            out, self.last_Tcw, st = self.h.optimize_slam(g, ell, iters, odom_meas=np.array(odo) if odo else None, odom_info=1.0)
        else:
            out, st = self.h.optimize_objects(g, ell, iters)
        self.last_stats = st
        for k, inst in enumerate(inst_ids):
            ellipsoids_by_instance[inst]["vec"] = out[k].copy()
        return ellipsoids_by_instance

"""Shared builders for the parity tests (synthetic graphs, comparisons)."""
import numpy as np

import orc
from eslam_b200 import synth

K = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]])


def make_ba_problem(n_frames=40, n_objects=10, edges_per_frame=10, seed=4242, with_gravity=True, noise=1.0):
    """Config-4 shaped graph (BASELINE.json configs[3]); measurements come from the oracle's own projection."""
    g = synth.ba_graph(n_frames, n_objects, edges_per_frame, seed)
    objs, Twc = g["objs"], g["Twc"]
    g["noise2d"] = g["noise2d"] * noise; g["noise3d"] = g["noise3d"] * noise
    Tcw = np.array([orc.se3_inverse(T) for T in Twc])
    et, fr, ob, meas, info = [], [], [], [], []
    if with_gravity:
        for o in range(n_objects):
            et.append(2); fr.append(0); ob.append(o); meas.append(np.r_[0, 0, 1, 0, np.zeros(6)]); info.append(100.0 ** 2)
    for f in range(n_frames):
        for k, o in enumerate(g["pick2d"][f]):
            bb = orc.ellipsoid_bbox(objs[o], Tcw[f], K) + g["noise2d"][f, k]
            et.append(0); fr.append(f); ob.append(int(o)); meas.append(np.r_[bb, np.zeros(6)]); info.append(1.0)
    for f in range(n_frames):
        for k, o in enumerate(g["pick3d"][f]):
            local = orc.ellipsoid_transform_from(orc.ellipsoid_exp_update(objs[o], g["noise3d"][f, k]), Tcw[f])
            et.append(1); fr.append(f); ob.append(int(o)); meas.append(local); info.append(10000.0 * 0.9)
    init = np.array([orc.ellipsoid_exp_update(objs[o], g["init_noise"][o]) for o in range(n_objects)])
    return dict(type=et, frame=fr, object=ob, meas=np.array(meas), info=info, Tcw=Tcw, n_objects=n_objects, K=K, init=init, truth=objs)


def make_slam_problem(n_frames=40, n_objects=10, edges_per_frame=10, seed=4242, pose_noise=0.01, noise=1.0):
    """make_ba_problem + SLAM mode inputs: odometry measured on the true poses (Optimizer.cpp:141-143), initial poses perturbed
    on the left (exp(xi) * Tcw, the vertex' own oplus) for every frame but the fixed first one."""
    p = make_ba_problem(n_frames, n_objects, edges_per_frame, seed, noise=noise)
    rng = np.random.default_rng(seed + 1)
    Tcw = p["Tcw"]
    p["odom"] = np.array([orc.se3_mul(Tcw[k + 1], orc.se3_inverse(Tcw[k])) for k in range(n_frames - 1)])
    Tn = Tcw.copy()
    for f in range(1, n_frames):
        Tn[f] = orc.se3_mul(orc.se3_exp(rng.normal(0, pose_noise, 6)), Tcw[f])
    p["Tcw_true"], p["Tcw"] = Tcw, Tn
    return p


def pose_close(a, b, rtol=1e-4):
    a, b = np.array(a, dtype=float), np.array(b, dtype=float)
    if np.dot(a[3:7], b[3:7]) < 0:
        b = b.copy(); b[3:7] = -b[3:7]
    return np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b)))


def ellipsoid_close(a, b, rtol=1e-4):
    """pose/axes within rtol (relative to max(1,|value|)); quaternion sign-normalised."""
    a, b = np.array(a, dtype=float), np.array(b, dtype=float)
    if np.dot(a[3:7], b[3:7]) < 0:
        b = b.copy(); b[3:7] = -b[3:7]
    return np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b)))


def dump_points(keep, name, n):
    return getattr(keep, name)[:n]


def poison_outside_boxes(depth, frame_of_obj, bbox, seed=7):
    """Copy of the depth batch with every pixel OUTSIDE the int-truncated boxes of its frame replaced by noise.  The path only reads
    pixels inside a box (PointCloudFilter.cpp:31-52 crop, EllipsoidExtractor.cpp:583-643 centre samples, SymmetrySolver.cpp:75-84
    isInRange before the range-image read), so results must not change -- the property the rectangle-only upload relies on."""
    rng = np.random.default_rng(seed)
    noisy = rng.integers(0, 30000, size=depth.shape, dtype=np.uint16)
    keep = np.zeros(depth.shape, dtype=bool)
    for f, b in zip(frame_of_obj, bbox):
        x1, y1, x2, y2 = (int(v) for v in b)
        keep[int(f), max(y1, 0):max(y2, 0), max(x1, 0):max(x2, 0)] = True
    return np.where(keep, depth, noisy)

"""Synthetic TUM-shaped RGB-D input for the benchmark and the parity tests.

640x480 uint16 depth (scale 5000), intrinsics of Example/param/TUM3.yaml:62-70.  A scene is a ground plane
z=0 (world, z up), a camera ~1.1 m above it pitched down ~25 deg, and cuboids standing on the floor.  Depth is
ray-cast analytically, perturbed with Kinect-like noise sigma = 0.0015*z^2, quantised to raw units, and 15 % of
the pixels are dropped to zero in 8x8 blobs (the bundled fr3 frames have 15.0 % zeros).  numpy only.
"""
import numpy as np

FX, FY, CX, CY = 535.4, 539.2, 320.1, 247.6
ROWS, COLS, SCALE = 480, 640, 5000.0


def _rot_to_quat(R):
    # Shepperd, returns x y z w with w >= 0
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def camera_pose(height=1.1, pitch_deg=25.0, yaw_deg=0.0, xy=(0.0, 0.0)):
    """Twc as (R_wc 3x3, t_wc 3) and the 7-vector x y z qx qy qz qw."""
    p, y = np.deg2rad(pitch_deg), np.deg2rad(yaw_deg)
    fwd = np.array([np.cos(p) * np.cos(y), np.cos(p) * np.sin(y), -np.sin(p)])
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], axis=1)
    t = np.array([xy[0], xy[1], height])
    pose7 = np.concatenate([t, _rot_to_quat(R)])
    return R, t, pose7


_PIX = None


def _pixel_rays():
    global _PIX
    if _PIX is None:
        u, v = np.meshgrid(np.arange(COLS, dtype=np.float64), np.arange(ROWS, dtype=np.float64))
        _PIX = np.stack([(u - CX) / FX, (v - CY) / FY, np.ones_like(u)], axis=-1)  # camera-frame dir with z=1
    return _PIX


def render_depth(R, t, boxes, rng, noise=True, dropout=0.15, max_range=8.0):
    """boxes: list of (center_xy(2), size(3)=(lx,ly,lz), yaw).  Returns uint16 [480,640] and per-box bbox."""
    d_cam = _pixel_rays()
    d_w = d_cam @ R.T                      # world directions, scaled so camera z == 1
    zc = np.full((ROWS, COLS), np.inf)     # camera-frame depth (z) of the nearest hit
    # ground z = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        s = -t[2] / d_w[..., 2]
    s = np.where((d_w[..., 2] < 0) & (s > 0), s, np.inf)
    zc = np.minimum(zc, s)
    bboxes = []
    for (cxy, size, yaw) in boxes:
        c, s_ = np.cos(yaw), np.sin(yaw)
        Rb = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1.0]])
        cen = np.array([cxy[0], cxy[1], size[2] / 2.0])
        o = (t - cen) @ Rb                 # ray origin in box frame
        dd = d_w @ Rb                      # directions in box frame
        half = np.asarray(size) / 2.0
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-half - o) / dd
            t2 = (half - o) / dd
        tmin = np.minimum(t1, t2).max(axis=-1)
        tmax = np.maximum(t1, t2).min(axis=-1)
        hit = (tmax >= tmin) & (tmax > 0)
        sb = np.where(hit & (tmin > 0), tmin, np.inf)
        zc = np.minimum(zc, sb)
        # 2-D box from the projected corners (what a detector would output)
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * half
        cw = corners @ Rb.T + cen
        cc = (cw - t) @ R
        uv = np.stack([FX * cc[:, 0] / cc[:, 2] + CX, FY * cc[:, 1] / cc[:, 2] + CY], axis=1)
        x1, y1 = np.floor(uv.min(axis=0))
        x2, y2 = np.ceil(uv.max(axis=0))
        bboxes.append([max(11.0, x1), max(11.0, y1), min(COLS - 12.0, x2), min(ROWS - 12.0, y2)])
    z = zc.copy()
    if noise:
        z = z + rng.standard_normal(z.shape) * 0.0015 * np.where(np.isfinite(z), z, 0.0) ** 2
    z = np.where(np.isfinite(z) & (z < max_range) & (z > 0.3), z, 0.0)
    raw = np.clip(np.round(z * SCALE), 0, 65535).astype(np.uint16)
    if dropout > 0:
        blobs = rng.random((ROWS // 8, COLS // 8)) < dropout
        raw[np.kron(blobs, np.ones((8, 8), dtype=bool))] = 0
    return raw, np.array(bboxes, dtype=np.float64)


def scene_one_box(frame_idx, seed_base=1000):
    """Config 2 (BASELINE.json configs[1]): one cabinet-like box (0.9 x 0.45 x 1.0 m) 1.5-1.8 m in front of the camera, random yaw;
    the detector box is ~360 x 380 px, i.e. ~10 k valid stride-3 samples ("10k pts segmented")."""
    rng = np.random.default_rng(seed_base + frame_idx)
    yaw_cam = rng.uniform(-180, 180)
    R, t, pose7 = camera_pose(height=1.1 + rng.uniform(-0.05, 0.05), pitch_deg=25 + rng.uniform(-2, 2), yaw_deg=yaw_cam,
                              xy=rng.uniform(-1, 1, size=2))
    fwd = np.array([np.cos(np.deg2rad(yaw_cam)), np.sin(np.deg2rad(yaw_cam))])
    dist = rng.uniform(1.5, 1.8)
    cxy = t[:2] + fwd * dist + rng.uniform(-0.05, 0.05, size=2)
    boxes = [(cxy, (0.9, 0.45, 1.0), rng.uniform(0, 2 * np.pi))]
    depth, bboxes = render_depth(R, t, boxes, rng)
    labels = np.array([28], dtype=np.int32)
    return depth, bboxes, labels, pose7


def scene_multi_box(frame_idx, n_boxes=8, seed_base=3000):
    """Config 3: n small boxes spread on the floor; labels cycle 28/41/62."""
    rng = np.random.default_rng(seed_base + frame_idx)
    yaw_cam = rng.uniform(-180, 180)
    R, t, pose7 = camera_pose(height=1.3, pitch_deg=32 + rng.uniform(-2, 2), yaw_deg=yaw_cam, xy=rng.uniform(-1, 1, size=2))
    y = np.deg2rad(yaw_cam)
    fwd = np.array([np.cos(y), np.sin(y)])
    lat = np.array([-np.sin(y), np.cos(y)])
    boxes = []
    cols = (n_boxes + 1) // 2
    for b in range(n_boxes):
        row, col = divmod(b, cols)
        dist = 1.5 + 0.9 * row
        off = (col - (cols - 1) / 2.0) * (0.55 + 0.25 * row)
        cxy = t[:2] + fwd * dist + lat * off
        size = (rng.uniform(0.22, 0.32), rng.uniform(0.18, 0.26), rng.uniform(0.3, 0.45))
        boxes.append((cxy, size, rng.uniform(0, 2 * np.pi)))
    depth, bboxes = render_depth(R, t, boxes, rng)
    labels = np.array([(28, 41, 62)[b % 3] for b in range(n_boxes)], dtype=np.int32)
    return depth, bboxes, labels, pose7


GROUND_PLANE_WORLD = np.array([0.0, 0.0, 1.0, 0.0])


def make_batch(n_frames, kind="one_box", seed_base=None, n_boxes=8):
    """Returns depth [F,480,640] uint16, frame_of_obj [N], bbox [N,4], label [N], Twc [F,7]."""
    depths, fo, bb, lb, poses = [], [], [], [], []
    for f in range(n_frames):
        if kind == "one_box":
            d, b, l, p = scene_one_box(f, 1000 if seed_base is None else seed_base)
        else:
            d, b, l, p = scene_multi_box(f, n_boxes, 3000 if seed_base is None else seed_base)
        depths.append(d)
        poses.append(p)
        for k in range(len(b)):
            fo.append(f)
            bb.append(b[k])
            lb.append(l[k])
    return (np.ascontiguousarray(np.stack(depths)), np.asarray(fo, dtype=np.int32), np.ascontiguousarray(np.asarray(bb, dtype=np.float64)),
            np.asarray(lb, dtype=np.int32), np.ascontiguousarray(np.asarray(poses, dtype=np.float64)))


def ba_graph(n_frames=500, n_objects=50, edges_per_frame=50, seed=4242):
    """Config 4: cameras on a circle r=2.5 m looking at the origin, n_objects ellipsoids on the floor, per keyframe
    25 EdgeSE3EllipsoidProj + 25 EdgeSE3Ellipsoid9DOF, + one gravity prior per object.  Measurements are generated by
    the caller with the oracle's projection (tests/bench), here only geometry + noise draws are produced."""
    rng = np.random.default_rng(seed)
    objs = np.zeros((n_objects, 10))
    objs[:, 0:2] = rng.uniform(-1.5, 1.5, size=(n_objects, 2))
    axes = rng.uniform(0.1, 0.6, size=(n_objects, 3))
    objs[:, 2] = axes[:, 2]
    yaw = rng.uniform(-np.pi, np.pi, size=n_objects)
    objs[:, 5] = np.sin(yaw / 2)
    objs[:, 6] = np.cos(yaw / 2)
    objs[:, 7:10] = axes
    Twc = np.zeros((n_frames, 7))
    for f in range(n_frames):
        a = 2 * np.pi * f / n_frames
        pos = np.array([4.5 * np.cos(a), 4.5 * np.sin(a), 1.6])
        fwd = np.array([0, 0, 0.3]) - pos
        fwd /= np.linalg.norm(fwd)
        right = np.cross(fwd, [0, 0, 1.0])
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd], axis=1)
        Twc[f, :3] = pos
        Twc[f, 3:] = _rot_to_quat(R)
    half = edges_per_frame // 2
    pick2d = np.stack([rng.permutation(n_objects)[:half] for _ in range(n_frames)])
    pick3d = np.stack([rng.permutation(n_objects)[:edges_per_frame - half] for _ in range(n_frames)])
    noise2d = rng.normal(0, 2.0, size=(n_frames, half, 4))
    noise3d = rng.normal(0, 0.02, size=(n_frames, edges_per_frame - half, 9))
    init_noise = rng.normal(0, 0.03, size=(n_objects, 9))
    return dict(objs=objs, Twc=Twc, pick2d=pick2d, pick3d=pick3d, noise2d=noise2d, noise3d=noise3d, init_noise=init_noise)


# ---- oracle-free measurement generation for the BA graph (numpy only: bench.py's product arm must not touch oracle/) ----
def quat_mul(a, b):
    """Hamilton product of x y z w quaternions."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_to_rot(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def pose_inv(T7):
    R = quat_to_rot(T7[3:7])
    q = np.array([-T7[3], -T7[4], -T7[5], T7[6]])
    return np.concatenate([-R.T @ T7[:3], q])


def pose_compose(A7, B7):
    q = quat_mul(A7[3:7], B7[3:7])
    if q[3] < 0:
        q = -q
    return np.concatenate([A7[:3] + quat_to_rot(A7[3:7]) @ B7[:3], q / np.linalg.norm(q)])


def ellipsoid_bbox_numpy(e10, Tcw7, K):
    """Tangent-line bounding box of the projected ellipsoid (dual quadric Q* = T diag(a^2,b^2,c^2,-1) T^T, dual conic C* = P Q* P^T;
    the lines x = u tangent to the conic satisfy l^T C* l = 0 with l = (1, 0, -u)).  Returns x1 y1 x2 y2."""
    T = np.eye(4)
    T[:3, :3] = quat_to_rot(e10[3:7])
    T[:3, 3] = e10[:3]
    Q = T @ np.diag([e10[7] ** 2, e10[8] ** 2, e10[9] ** 2, -1.0]) @ T.T
    Tc = np.eye(4)
    Tc[:3, :3] = quat_to_rot(Tcw7[3:7])
    Tc[:3, 3] = Tcw7[:3]
    P = np.asarray(K, dtype=np.float64).reshape(3, 3) @ Tc[:3, :]
    Cs = P @ Q @ P.T
    du = np.sqrt(Cs[0, 2] ** 2 - Cs[0, 0] * Cs[2, 2])
    dv = np.sqrt(Cs[1, 2] ** 2 - Cs[1, 1] * Cs[2, 2])
    xs = sorted([(Cs[0, 2] - du) / Cs[2, 2], (Cs[0, 2] + du) / Cs[2, 2]])
    ys = sorted([(Cs[1, 2] - dv) / Cs[2, 2], (Cs[1, 2] + dv) / Cs[2, 2]])
    return np.array([xs[0], ys[0], xs[1], ys[1]])


def _perturb_ellipsoid(e10, n9):
    """Small perturbation of an ellipsoid: translation, rotation vector (left-multiplied), half-axes."""
    out = np.array(e10, dtype=np.float64)
    out[:3] += n9[:3]
    w = n9[3:6]
    ang = np.linalg.norm(w)
    dq = np.array([0, 0, 0, 1.0]) if ang < 1e-12 else np.concatenate([np.sin(ang / 2) * w / ang, [np.cos(ang / 2)]])
    q = quat_mul(dq, out[3:7])
    out[3:7] = q / np.linalg.norm(q) * (1.0 if q[3] >= 0 else -1.0)
    out[7:10] = np.abs(out[7:10] + n9[6:9])
    return out


def ba_problem(n_frames=500, n_objects=50, edges_per_frame=50, seed=4242, with_gravity=True, pose_noise=0.0):
    """BASELINE.json configs[3] as plain arrays (type/frame/object/meas/info/Tcw/init), measurements generated with numpy only.
    pose_noise > 0 also returns SLAM-mode inputs (odometry from the true poses, perturbed initial poses)."""
    Kmat = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1.0]])
    g = ba_graph(n_frames, n_objects, edges_per_frame, seed)
    objs, Twc = g["objs"], g["Twc"]
    Tcw = np.array([pose_inv(T) for T in Twc])
    et, fr, ob, meas, info = [], [], [], [], []
    if with_gravity:
        for o in range(n_objects):
            et.append(2); fr.append(0); ob.append(o); meas.append(np.r_[0, 0, 1, 0, np.zeros(6)]); info.append(100.0 ** 2)
    for f in range(n_frames):
        for k, o in enumerate(g["pick2d"][f]):
            bb = ellipsoid_bbox_numpy(objs[o], Tcw[f], Kmat) + g["noise2d"][f, k]
            et.append(0); fr.append(f); ob.append(int(o)); meas.append(np.r_[bb, np.zeros(6)]); info.append(1.0)
    for f in range(n_frames):
        for k, o in enumerate(g["pick3d"][f]):
            e = _perturb_ellipsoid(objs[o], g["noise3d"][f, k])
            local = np.concatenate([pose_compose(Tcw[f], e[:7]), e[7:10]])
            et.append(1); fr.append(f); ob.append(int(o)); meas.append(local); info.append(10000.0 * 0.9)
    init = np.array([_perturb_ellipsoid(objs[o], g["init_noise"][o]) for o in range(n_objects)])
    p = dict(type=np.asarray(et, np.int32), frame=np.asarray(fr, np.int32), object=np.asarray(ob, np.int32), meas=np.array(meas), info=np.asarray(info, np.float64),
             Tcw=Tcw, n_objects=n_objects, K=Kmat, init=init, truth=objs)
    if pose_noise > 0:
        rng = np.random.default_rng(seed + 1)
        p["odom"] = np.array([pose_compose(Tcw[k + 1], pose_inv(Tcw[k])) for k in range(n_frames - 1)])
        Tn = Tcw.copy()
        for f in range(1, n_frames):
            w = rng.normal(0, pose_noise, 6)
            ang = np.linalg.norm(w[3:])
            dq = np.array([0, 0, 0, 1.0]) if ang < 1e-12 else np.concatenate([np.sin(ang / 2) * w[3:] / ang, [np.cos(ang / 2)]])
            Tn[f] = pose_compose(np.concatenate([w[:3], dq]), Tcw[f])
        p["Tcw_true"], p["Tcw"] = Tcw, Tn
    return p

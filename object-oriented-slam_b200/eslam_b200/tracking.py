"""N1: per-frame orchestration of the hot path -- a thin host-side driver over the C-ABI that follows
src/core/Tracking.cpp (GrabPoseAndObjects :170-204, UpdateObjectObservation :493-564, UpdateDepthEllipsoidEstimation :290-381,
Update3DObservationDataAssociation :383-417, checkKeyFrameForInstances :421-475, JudgeInitialization :567-638,
ProcessGroundPlaneEstimation :716-799) and the result writer of System::SaveObjectsToFile (src/core/System.cpp:75-91).
All compute (plane, ellipsoids, initialiser, BA) goes through a `backend` with the Handle interface (eslam_b200.api.Handle on the GPU;
the tests plug in an oracle-backed object with the same methods to exercise this host logic without a GPU).  No viewer, no Map
visualisation lists, no interactive prompts."""
import numpy as np

from .api import Frame, Observation, calibrate_measurement, plane_to_world

MINIMUM_INITIALIZATION_FRAME = 15   # Tracking_MINIMUM_INITIALIZATION_FRAME (src/config/Config.cpp:29)


def se3_mul(a, b):
    """g2o::SE3Quat::operator* (se3quat.h:110-116) on 7-vectors x y z qx qy qz qw (normalises, w >= 0)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    qa, qb = a[3:], b[3:]
    qv = qa[:3]
    uv = 2.0 * np.cross(qv, b[:3])
    t = a[:3] + (b[:3] + qa[3] * uv + np.cross(qv, uv))
    x = qa[3] * qb[0] + qa[0] * qb[3] + qa[1] * qb[2] - qa[2] * qb[1]
    y = qa[3] * qb[1] + qa[1] * qb[3] + qa[2] * qb[0] - qa[0] * qb[2]
    z = qa[3] * qb[2] + qa[2] * qb[3] + qa[0] * qb[1] - qa[1] * qb[0]
    w = qa[3] * qb[3] - qa[0] * qb[0] - qa[1] * qb[1] - qa[2] * qb[2]
    q = np.array([x, y, z, w])
    if q[3] < 0:
        q = -q
    return np.concatenate([t, q / np.linalg.norm(q)])


def pose_inverse_np(T):
    T = np.asarray(T, float)
    qc = np.array([-T[3], -T[4], -T[5], T[6]])
    v = -T[:3]
    uv = 2.0 * np.cross(qc[:3], v)
    return np.concatenate([v + qc[3] * uv + np.cross(qc[:3], uv), qc])


def ellipsoid_to_minimal_vector(vec10):
    """ellipsoid::toMinimalVector (src/core/Ellipsoid.cpp:144-149): x y z roll pitch yaw a b c with SE3Quat::toXYZPRYVector's ZYX Euler
    angles of the quaternion (Thirdparty/g2o/g2o/types/se3quat.h:184-207)."""
    v = np.asarray(vec10, float)
    qx, qy, qz, qw = v[3:7]
    roll = np.arctan2(2.0 * (qw * qx + qy * qz), 1.0 - 2.0 * (qx * qx + qy * qy))
    pitch = np.arcsin(2.0 * (qw * qy - qz * qx))
    yaw = np.arctan2(2.0 * (qw * qz + qx * qy), 1.0 - 2.0 * (qy * qy + qz * qz))
    return np.concatenate([v[:3], [roll, pitch, yaw], v[7:10]])


class DataAssociationSolver:
    """N4: src/core/DataAssociation.cpp:16-135 -- greedy nearest-centre association (threshold 1 m) of the frame's single-frame
    ellipsoids with the map, used when withAssociation=False ("not steady" per Tracking.cpp:505-518).  Quirks kept: the value returned
    for a match is the COLUMN index in the instance-ordered map (DataAssociation.cpp:113-114), new instances count from 0."""

    def __init__(self):
        self.miInstanceNum = 0

    def CreateInstance(self):
        self.miInstanceNum += 1
        return self.miInstanceNum - 1

    def Solve(self, frame, map_ellipsoids):
        n = len(frame.mmObservations)
        assoc = [-1] * n
        valid = [i for i in range(n) if frame.mpLocalObjects[i] is not None]
        if not valid:
            return assoc
        insts = sorted(map_ellipsoids)
        cost = np.zeros((len(valid), len(insts)))
        for a, i in enumerate(valid):
            world = se3_mul(frame.cam_pose_Twc, np.asarray(frame.mpLocalObjects[i]["vec"], float)[:7])
            for b, inst in enumerate(insts):
                cost[a, b] = np.linalg.norm(np.asarray(map_ellipsoids[inst]["vec"])[:3] - world[:3])
        for a, i in enumerate(valid):                       # SolveCostMat :90-135
            if cost.shape[1] < 1:
                assoc[i] = self.CreateInstance()
                continue
            j = int(np.argmin(cost[a]))
            if cost[a, j] < 1.0:
                assoc[i] = j
                cost[:, j] = 999
            else:
                assoc[i] = self.CreateInstance()
        return assoc


    def SolveBatch(self, frames, map_ellipsoids, backend=None):
        """The frames of a batch against ONE map snapshot (the frame-parallel use of the path): what calling Solve frame by frame would
        return while the map does not change -- matches as column indices, new instances numbered in (frame, row) order.  With a
        `backend` that has associate_frames (eslam_b200.api.Handle) the batch runs on the device (csrc/associate.cu, one warp per frame)."""
        if backend is None or not hasattr(backend, "associate_frames"):
            return [self.Solve(f, map_ellipsoids) for f in frames]
        insts = sorted(map_ellipsoids)
        centres = np.array([np.asarray(map_ellipsoids[i]["vec"], float)[:3] for i in insts]).reshape(-1, 3)
        start, poses, valid = [0], [], []
        for f in frames:
            for obj in f.mpLocalObjects[:len(f.mmObservations)]:
                valid.append(0 if obj is None else 1)
                poses.append(np.zeros(7) if obj is None else np.asarray(obj["vec"], float)[:7])
            start.append(len(valid))
        assoc, n_new = backend.associate_frames(start, np.array([f.cam_pose_Twc for f in frames]), np.array(poses).reshape(-1, 7), valid, centres, self.miInstanceNum)
        self.miInstanceNum += n_new
        return [[int(v) for v in assoc[start[k]:start[k + 1]]] for k in range(len(frames))]


class Tracking:
    def __init__(self, backend, optimizer_cls, K, rows=480, cols=640, keyframe_check_close=True, min_init_frames=MINIMUM_INITIALIZATION_FRAME):
        self.h = backend
        self.optimizer = optimizer_cls(backend)
        self.K = np.asarray(K, float)
        self.rows, self.cols = rows, cols
        self.keyframe_check_close = keyframe_check_close
        self.min_init_frames = min_init_frames
        self.miGroundPlaneState = 1          # OpenGroundPlaneEstimation (:690)
        self.mGroundPlane = None
        self.mvpFrames = []
        self.mmObjectObservations = {}       # instance -> [Observation]
        self.mmObjectObservations3D = {}     # instance -> [(frame, ellipsoid dict)]
        self.map_ellipsoids = {}             # instance -> ellipsoid dict (world frame)  == Map::GetAllEllipsoidsMap()
        self.mmObjectHistory = {}            # instance -> rows [n_obs, valid, minimal vector(9)]   (Tracking.h mmObjectHistory)
        self.pDASolver = DataAssociationSolver()

    # ---- :716-799
    # ---- Tracking::RefreshObjectHistory (src/core/Tracking.cpp:810-852): one row per (instance, number of 2-D observations) --
    # [observation count, valid = 1, x y z roll pitch yaw a b c]; a new optimisation with the same observation count overwrites the last row
    def RefreshObjectHistory(self):
        for inst in sorted(self.map_ellipsoids):
            hist = self.mmObjectHistory.setdefault(inst, [])
            n_obs = len(self.mmObjectObservations.get(inst, []))     # the reference asserts that the instance has 2-D observations
            row = np.concatenate([[float(n_obs), 1.0], ellipsoid_to_minimal_vector(self.map_ellipsoids[inst]["vec"])])
            if hist and int(round(hist[-1][0])) == n_obs:
                hist[-1] = row
            else:
                hist.append(row)

    # ---- Tracking::SaveObjectHistory (src/core/Tracking.cpp:855-896): "TotalInstanceNum", then per instance "id historyNum" and its rows;
    # values go through a default-precision ostream (6 significant digits)
    def SaveObjectHistory(self, path):
        with open(path, "w") as f:
            f.write(f"{len(self.mmObjectHistory)}\n")
            for inst in sorted(self.mmObjectHistory):
                rows = self.mmObjectHistory[inst]
                f.write(f"{inst} {len(rows)}\n")
                for r in rows:
                    f.write(" ".join(f"{float(v):.6g}" for v in r) + "\n")

    def ProcessGroundPlaneEstimation(self, frame):
        pl, found, _ = self.h.extract_ground_plane(frame.frame_img)
        if not found[0]:
            return False
        self.mGroundPlane = plane_to_world(pl[0], frame.cam_pose_Twc)
        self.miGroundPlaneState = 2
        self.h.set_supporting_plane(self.mGroundPlane)
        self.optimizer.SetGroundPlane(self.mGroundPlane)
        return True

    # ---- :421-475
    def checkKeyFrameForInstances(self, frame, associations):
        dis_th, ang_th = (0.0, 0.0) if self.keyframe_check_close else (0.4, np.pi / 180.0 * 15)
        checks = []
        for inst in associations:
            if inst < 0:
                checks.append(False)
            elif inst not in self.mmObjectObservations:
                checks.append(True)
            else:
                last = self.mmObjectObservations[inst][-1].pFrame.cam_pose_Twc
                diff = se3_mul(pose_inverse_np(frame.cam_pose_Twc), last)
                dis = np.linalg.norm(diff[:3])
                n = np.linalg.norm(diff[3:6])
                angle = 2.0 * np.arctan2(n, abs(diff[6]))      # Eigen::AngleAxisd(quat).angle()
                checks.append(bool(dis > dis_th or angle > ang_th))
        return checks

    # ---- :170-204
    def GrabPoseAndObjects(self, timestamp, pose, bboxMat, imDepth, withAssociation=True, iters=10):
        frame = Frame(timestamp, pose, bboxMat, imDepth)
        frame.frame_seq_id = len(self.mvpFrames)        # Frame::total_frame is per-process in the reference; per-tracker here
        self.mvpFrames.append(frame)
        obs_mat = frame.mmObservations
        if self.miGroundPlaneState == 1:
            self.ProcessGroundPlaneEstimation(frame)
        # ---- UpdateDepthEllipsoidEstimation :290-381 (all eligible boxes of the frame in ONE batched call)
        frame.mpLocalObjects = [None] * len(obs_mat)
        eligible = []
        for i, det in enumerate(obs_mat):
            is_border = calibrate_measurement(det[1:5], self.h.params, self.h.camera)
            c2 = self.miGroundPlaneState == 2
            c3 = withAssociation and int(round(det[7])) < 0
            if (not is_border) and c2 and not c3:
                eligible.append(i)
        if eligible:
            out, _, _ = self.h.estimate_ellipsoids(frame.frame_img, [0] * len(eligible), obs_mat[eligible, 1:5], [int(round(obs_mat[i, 5])) for i in eligible],
                                                   frame.cam_pose_Twc[None])
            for k, i in enumerate(eligible):
                if out[k].status == 0:
                    frame.mpLocalObjects[i] = dict(vec=np.array(out[k].ellipsoid), prob=float(out[k].prob), miLabel=int(round(obs_mat[i, 5])), miInstanceID=-1)
        if withAssociation:
            associations = [int(round(det[7])) for det in obs_mat]      # GetMannualAssociation :477-490
        else:
            associations = self.pDASolver.Solve(frame, self.map_ellipsoids)   # :515-518
        key = self.checkKeyFrameForInstances(frame, associations)
        # ---- Update3DObservationDataAssociation :383-417
        for i, inst in enumerate(associations):
            obj = frame.mpLocalObjects[i]
            if obj is None or inst < 0:
                continue
            if not key[i]:
                frame.mpLocalObjects[i] = None
                continue
            obj["miInstanceID"] = inst
            self.mmObjectObservations3D.setdefault(inst, []).append((frame, obj))
        # ---- 2-D observations :532-563
        for i, det in enumerate(obs_mat):
            inst = associations[i]
            if inst < 0 or not key[i]:
                continue
            if calibrate_measurement(det[1:5], self.h.params, self.h.camera):
                continue
            self.mmObjectObservations.setdefault(inst, []).append(Observation(int(det[5]), det[1:5], det[6], frame))
        self.JudgeInitialization()
        # ---- ProcessCurrentFrame :219-226
        if self.map_ellipsoids:
            self.optimizer.GlobalObjectGraphOptimization(self.mvpFrames, self.map_ellipsoids, self.rows, self.cols, self.K, self.mmObjectObservations, iters=iters)
            self.RefreshObjectHistory()                                   # :228
        return frame

    # ---- :567-638
    def JudgeInitialization(self):
        for inst, obs in self.mmObjectObservations.items():
            if inst in self.map_ellipsoids or len(obs) < self.min_init_frames:
                continue
            Twc = np.array([o.pFrame.cam_pose_Twc for o in obs]); det = np.array([np.r_[o.bbox, o.rate] for o in obs])
            e, ok = self.h.init_quadric(Twc, det, self.K, self.rows, self.cols)
            if ok:
                self.map_ellipsoids[inst] = dict(vec=e, prob=1.0, miLabel=obs[0].label, miInstanceID=inst)
        for inst, lst in self.mmObjectObservations3D.items():
            if inst in self.map_ellipsoids:
                continue
            frame, obj = lst[-1]
            vec = np.array(obj["vec"], float)
            vec[:7] = se3_mul(frame.cam_pose_Twc, vec[:7])             # pE->transform_from(cam_pose_Twc)
            self.map_ellipsoids[inst] = dict(vec=vec, prob=obj["prob"], miLabel=obj["miLabel"], miInstanceID=inst)

    # ---- System::SaveObjectsToFile (System.cpp:75-91, dataprocess_utils.cpp:105-127): "instance x y z qx qy qz qw a b c", 12 significant digits
    def SaveObjectsToFile(self, path):
        with open(path, "w") as f:
            for inst in sorted(self.map_ellipsoids):
                row = [float(inst)] + [float(v) for v in self.map_ellipsoids[inst]["vec"]]
                f.write(" ".join(f"{v:.12g}" for v in row) + "\n")

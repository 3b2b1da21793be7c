"""Generates tests/golden/cabinet_golden.npz from the reference's bundled fr3_cabinet sequence (Example/dataset/cabinet).

Run in the authoring container only (the GPU box has no /root/reference):   python tests/golden/make_golden.py
Inputs kept: 4 depth frames (uint16, 640x480), their YOLO bbox row (bbox/<ts>.txt: id x1 y1 x2 y2 label prob instance) and the
ground-truth pose (associateGroundtruth.txt: ts rgb ts x y z qx qy qz qw).  Outputs stored: what the CPU oracle computes for them
(ground plane per frame, supporting plane = frame 0's plane in the world, single-frame ellipsoid per bbox, integer stage counts).
These pin the oracle against regressions and give the GPU path real-sensor inputs; they are NOT outputs of the reference binary
(which cannot be built here: no Eigen/PCL/OpenCV)."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "object-oriented-slam_b200")]
import orc  # noqa: E402

REF = "/root/reference/Example/dataset/cabinet"
lines = [l.split() for l in open(os.path.join(REF, "associateGroundtruth.txt")) if l.strip()]
picked, depth, bbox, label, pose = [], [], [], [], []
for idx in (0, 14, 30, 50):
    ln = lines[idx]
    ts = ln[1].split("/")[1][:-4]
    bfile = os.path.join(REF, "bbox", ts + ".txt")
    rows = [r.split() for r in open(bfile) if r.strip()] if os.path.exists(bfile) else []
    if not rows:
        continue
    d = cv2.imread(os.path.join(REF, "depth", ts + ".png"), cv2.IMREAD_UNCHANGED)
    assert d.dtype == np.uint16 and d.shape == (480, 640)
    picked.append(ts); depth.append(d)
    r = rows[0]
    bbox.append([float(v) for v in r[1:5]]); label.append(int(float(r[5])))
    pose.append([float(v) for v in ln[3:10]])
depth = np.stack(depth); bbox = np.array(bbox); label = np.array(label, np.int32); pose = np.array(pose)
planes, found, ninl = orc.extract_ground_plane(depth)
plane_world = orc.plane_to_world(planes[0], pose[0])
fo = np.arange(len(depth), dtype=np.int32)
out, _, _, evals = orc.estimate_ellipsoids(depth, fo, bbox, label, pose, plane_world)
ell = np.array([list(o.ellipsoid) for o in out]); prob = np.array([o.prob for o in out])
counts = np.array([[o.status, o.n_crop, o.n_voxel, o.n_plane, o.n_clusters, o.n_cluster, o.n_sym, o.n_final, o.best_seed] for o in out], np.int32)
np.savez_compressed(os.path.join(HERE, "cabinet_golden.npz"), timestamps=np.array(picked), depth=depth, bbox=bbox, label=label, pose=pose,
                    planes_cam=planes, plane_found=found, plane_inliers=ninl, plane_world=plane_world, ellipsoid=ell, prob=prob, counts=counts)
print("frames", picked); print("planes", planes, found, ninl); print("plane_world", plane_world); print("counts", counts); print("ellipsoids", ell)

"""N3: reader for the TUM-RGB-D-shaped directory the reference demo consumes (src/tum_rgbd/io.cpp:14-94,277-290;
README.md:66-69) and the bbox text format (`id x1 y1 x2 y2 label rate instance` per row, bbox/<rgb timestamp>.txt).
Only what feeds the hot path: depth PNG (uint16), ground-truth pose, detections."""
import os

import numpy as np


def load_sequence(root, max_frames=None):
    """Returns a list of dicts {timestamp, depth_path, pose7 (x y z qx qy qz qw), bbox (N x 8)} following associateGroundtruth.txt:
    `rgb_ts rgb_path gt_ts x y z qx qy qz qw` and associate.txt `rgb_ts rgb_path depth_ts depth_path`."""
    depth_of = {}
    with open(os.path.join(root, "associate.txt")) as f:
        for ln in f:
            p = ln.split()
            if len(p) >= 4:
                depth_of[p[0]] = p[3]
    frames = []
    with open(os.path.join(root, "associateGroundtruth.txt")) as f:
        for ln in f:
            p = ln.split()
            if len(p) < 10 or p[0] not in depth_of:
                continue
            ts_name = os.path.splitext(os.path.basename(p[1]))[0]
            bfile = os.path.join(root, "bbox", ts_name + ".txt")
            rows = []
            if os.path.exists(bfile):
                rows = [[float(v) for v in r.split()] for r in open(bfile) if r.strip()]
            bbox = np.array(rows, float).reshape(-1, 8) if rows else np.zeros((0, 8))
            frames.append(dict(timestamp=float(p[0]), depth_path=os.path.join(root, depth_of[p[0]]), pose7=np.array([float(v) for v in p[3:10]]), bbox=bbox))
            if max_frames and len(frames) >= max_frames:
                break
    return frames


def read_depth(path):
    import cv2
    d = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if d is None or d.dtype != np.uint16:
        raise IOError(f"{path}: expected a 16-bit depth PNG")
    return d

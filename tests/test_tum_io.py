"""N3: dataset reader, result writers / readers and the prefetching depth loader -- host code, no GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "object-oriented-slam_b200")]
from eslam_b200 import tum_io  # noqa: E402
from eslam_b200.tracking import Tracking, ellipsoid_to_minimal_vector  # noqa: E402

cv2 = pytest.importorskip("cv2")


def _write_sequence(root, n, rng):
    """A TUM-RGB-D-shaped directory as the reference demo reads it (src/tum_rgbd/io.cpp:14-94): associate.txt, associateGroundtruth.txt, depth/*.png, bbox/*.txt."""
    os.makedirs(os.path.join(root, "depth")); os.makedirs(os.path.join(root, "bbox")); os.makedirs(os.path.join(root, "rgb"))
    depths = []
    with open(os.path.join(root, "associate.txt"), "w") as fa, open(os.path.join(root, "associateGroundtruth.txt"), "w") as fg:
        for k in range(n):
            ts = f"{1341841278.0 + 0.1 * k:.4f}"
            d = rng.integers(0, 30000, size=(480, 640), dtype=np.uint16)
            depths.append(d)
            assert cv2.imwrite(os.path.join(root, "depth", ts + ".png"), d)
            fa.write(f"{ts} rgb/{ts}.png {ts} depth/{ts}.png\n")
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            fg.write(f"{ts} rgb/{ts}.png {ts} {0.1 * k} 0.2 0.3 {q[0]} {q[1]} {q[2]} {q[3]}\n")
            with open(os.path.join(root, "bbox", ts + ".txt"), "w") as fb:
                fb.write(f"0 100 120 300 340 28 0.9 {k % 2}\n1 10 20 50 60 59 0.5 -1\n")
    return depths


def test_load_sequence_and_prefetcher_deliver_the_frames_in_order(tmp_path):
    rng = np.random.default_rng(3)
    depths = _write_sequence(str(tmp_path), 7, rng)
    frames = tum_io.load_sequence(str(tmp_path))
    assert len(frames) == 7 and frames[3]["bbox"].shape == (2, 8) and frames[3]["bbox"][0, 5] == 28 and np.isclose(frames[3]["pose7"][0], 0.3)
    assert np.array_equal(tum_io.read_depth(frames[2]["depth_path"]), depths[2])
    seen = 0
    pf = tum_io.DepthPrefetcher(frames, batch=3, workers=2, pinned=False)
    batches = []
    for chunk, depth in pf:
        assert depth.dtype == np.uint16 and depth.shape == (len(chunk), 480, 640)
        batches.append(len(chunk))
        for k, fr in enumerate(chunk):
            assert fr is frames[seen + k] and np.array_equal(depth[k], depths[seen + k])
        seen += len(chunk)
    assert batches == [3, 3, 1] and seen == 7


def test_prefetcher_reports_a_bad_frame(tmp_path):
    rng = np.random.default_rng(4)
    _write_sequence(str(tmp_path), 2, rng)
    frames = tum_io.load_sequence(str(tmp_path))
    cv2.imwrite(frames[1]["depth_path"], np.zeros((10, 10), np.uint8))   # not a 16-bit 480 x 640 depth image
    with pytest.raises(IOError):
        for _ in tum_io.DepthPrefetcher(frames, batch=2, workers=1, pinned=False):
            pass


def test_minimal_vector_is_the_zyx_euler_form():
    # yaw 0.3 about z, then pitch -0.2 about y, then roll 0.1 about x (ZYX): q = qz * qy * qx
    def qmul(a, b):
        x1, y1, z1, w1 = a; x2, y2, z2, w2 = b
        return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    r, p, y = 0.1, -0.2, 0.3
    q = qmul(qmul([0, 0, np.sin(y / 2), np.cos(y / 2)], [0, np.sin(p / 2), 0, np.cos(p / 2)]), [np.sin(r / 2), 0, 0, np.cos(r / 2)])
    v = ellipsoid_to_minimal_vector(np.r_[1.0, 2.0, 3.0, q, 0.4, 0.5, 0.6])
    assert np.allclose(v, [1, 2, 3, r, p, y, 0.4, 0.5, 0.6], atol=1e-12)


def test_object_history_rows_and_file_format(tmp_path):
    """Tracking::RefreshObjectHistory / SaveObjectHistory (Tracking.cpp:810-896): a row per new observation count, the last row overwritten
    while the count stays, file = total, then "id n" + n rows of 11 values at 6 significant digits."""
    tr = Tracking.__new__(Tracking)
    tr.map_ellipsoids, tr.mmObjectObservations, tr.mmObjectHistory = {}, {}, {}
    e0 = np.r_[0.123456789, 2.0, 3.0, 0.0, 0.0, 0.0, 1.0, 0.4, 0.5, 0.6]
    tr.map_ellipsoids[2] = dict(vec=e0.copy()); tr.mmObjectObservations[2] = [None] * 3
    tr.RefreshObjectHistory()
    tr.map_ellipsoids[2]["vec"][0] = 0.5                       # re-optimised, same number of observations: the row is replaced
    tr.RefreshObjectHistory()
    tr.mmObjectObservations[2].append(None)                     # a new observation: a new row
    tr.map_ellipsoids[0] = dict(vec=e0.copy()); tr.mmObjectObservations[0] = [None]
    tr.RefreshObjectHistory()
    assert [len(tr.mmObjectHistory[k]) for k in (0, 2)] == [1, 2] and tr.mmObjectHistory[2][0][0] == 3 and tr.mmObjectHistory[2][1][0] == 4
    p = tmp_path / "object_history.txt"
    tr.SaveObjectHistory(str(p))
    lines = open(p).read().splitlines()
    assert lines[0] == "2" and lines[1] == "0 1" and lines[3] == "2 2" and len(lines) == 6
    assert lines[2].split()[:3] == ["1", "1", "0.123457"]      # default ostream precision
    h = tum_io.read_object_history(str(p))
    assert sorted(h) == [0, 2] and h[2].shape == (2, 11) and np.allclose(h[2][:, 2], 0.5) and np.allclose(h[0][0, 8:], [0.4, 0.5, 0.6])
    # objects.txt round trip
    tr.SaveObjectsToFile(str(tmp_path / "objects.txt"))
    o = tum_io.read_objects(str(tmp_path / "objects.txt"))
    assert sorted(o) == [0, 2] and np.allclose(o[2], tr.map_ellipsoids[2]["vec"], rtol=1e-11)

"""Smallest batch that takes the symmetry round pipeline (16 one-box objects) + the batched association: target for compute-sanitizer --tool racecheck."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np  # noqa: E402

import eslam_b200 as E  # noqa: E402
from eslam_b200 import synth  # noqa: E402

os.environ.setdefault("ESLAM_SYM_GRAPH", "0")
depth, fo, bb, lb, tw = synth.make_batch(16, "one_box")
h = E.Handle(); h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
out, _, _ = h.estimate_ellipsoids(depth, fo, bb, lb, tw)
print("status", [o.status for o in out], "launches", h.launches)
a, n = h.associate_frames([0, 2, 3], np.tile([0, 0, 0, 0, 0, 0, 1.0], (2, 1)), np.tile([0.1, 0, 1, 0, 0, 0, 1.0], (3, 1)), [1, 1, 1], [[0.1, 0, 1.0], [5, 5, 5]])
print("assoc", a, n)
h.close()

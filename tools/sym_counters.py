import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np
import eslam_b200 as E
from eslam_b200 import synth
import bench
depth, fo, bb, lb, tw = bench.make_frames("one_box", 64, 1000)
h = E.Handle(); h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
out, _, _ = h.estimate_ellipsoids(depth, fo, bb, lb, tw)
c = np.array([h.debug_counters(i) for i in range(64)], float)
ok = c[:, 0] > 0
print("objects with evals:", ok.sum(), "status", [o.status for o in out][:16])
c = c[ok]
print("per object: half-evals %.0f  NN scans/half-eval %.1f  cache hits/half-eval %.1f  points/half-eval %.1f  n_sym %.1f  hit rate %.3f" % (
    c[:,0].mean(), (c[:,1]/c[:,0]).mean(), (c[:,4]/c[:,0]).mean(), (c[:,2]/c[:,0]).mean(), c[:,3].mean(), (c[:,4].sum()/(c[:,4].sum()+c[:,1].sum()))))
print("evals min/max", c[:,0].min(), c[:,0].max(), " first rows", c[:4])

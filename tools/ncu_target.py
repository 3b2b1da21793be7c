"""Short ncu target: N one-box objects resident on the device, both extraction kernels launched back to back (profiling mode =
single stream, 2 launches per step).  usage: ncu ... python tools/ncu_target.py [n_frames] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np
import torch
import eslam_b200 as E
from eslam_b200 import synth
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 888
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
depth, fo, bb, lb, tw = bench.make_frames("one_box", n, 1000)
h = E.Handle(); h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
d = torch.from_numpy(depth.view(np.int16)).cuda()
h.set_profiling(True)
for _ in range(steps):
    h.estimate_ellipsoids_dev(d.data_ptr(), n, depth.shape[2], fo, bb, lb, tw)
    torch.cuda.synchronize()
print("stage ms", h.stage_ms())

"""Resident throughput vs number of object chunks / streams (env ESLAM_DEV_CHUNKS)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np, torch
import eslam_b200 as E
from eslam_b200 import synth
import bench
n = 1024
depth, fo, bb, lb, tw = bench.make_frames("one_box", n, 1000)
d = torch.from_numpy(depth.view(np.int16)).cuda()
for ch in (4, 8, 12, 16):
    os.environ["ESLAM_DEV_CHUNKS"] = str(ch)
    h = E.Handle(); h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
    run = lambda: h.estimate_ellipsoids_dev(d.data_ptr(), n, depth.shape[2], fo, bb, lb, tw)
    for _ in range(3): run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): run()
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 10 * 1e3
    print(f"dev chunks {ch:2d}: {ms:6.2f} ms/step  {n/ms*1e3:8.0f} frames/s", flush=True)
    h.close()

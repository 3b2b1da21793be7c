"""e2e (host buffers -> records) throughput of the extraction entry under different upload modes / pipeline depths.
Run on the GPU box: python tools/e2e_probe.py [frames] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np
import torch
import eslam_b200 as E
from eslam_b200 import synth
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
depth, fo, bb, lb, tw = bench.make_frames("one_box", n, 1000)
host = torch.from_numpy(depth.view(np.int16)).pin_memory().numpy().view(np.uint16)
modes = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [4, 3, 0]
chunk_list = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else [4, 8]
for mode in modes:
    for chunks in chunk_list:
        os.environ["ESLAM_H2D_MODE"] = str(mode); os.environ["ESLAM_HOST_CHUNKS"] = str(chunks)
        h = E.Handle(); h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
        for _ in range(2):
            h.estimate_ellipsoids(host, fo, bb, lb, tw)
        torch.cuda.synchronize(); b0 = h.h2d_bytes; t0 = time.perf_counter()
        for _ in range(steps):
            h.estimate_ellipsoids(host, fo, bb, lb, tw)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / steps
        print(f"h2d_mode {mode} chunks {chunks}: {dt*1e3:7.2f} ms/step  {n/dt:9.0f} frames/s  h2d {(h.h2d_bytes-b0)/steps/1e6:7.1f} MB/step", flush=True)
        os.environ["ESLAM_TRACE"] = "1"; h.estimate_ellipsoids(host, fo, bb, lb, tw); os.environ.pop("ESLAM_TRACE")
        h.close()

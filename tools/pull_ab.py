"""A/B of the host->device upload of the box rectangles in the e2e path: LDG/STG pull kernel vs the cp.async.bulk (TMA) pull kernel.
  python tools/pull_ab.py [--frames 1024] [--steps 10]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import eslam_b200 as E  # noqa: E402
from eslam_b200 import synth  # noqa: E402
from bench import make_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--settings", default="0:8:2,1:8:2", help="comma-separated pull_tma:stages:ctas_per_sm")
args = ap.parse_args()
depth, fo, bb, lb, tw = make_frames("one_box", args.frames, 1000)
host = torch.from_numpy(depth.view(np.int16)).pin_memory().numpy().view(np.uint16)
ref = None
for rep in range(2):
    for setting in args.settings.split(","):
        tma, stages, ctas = (setting.split(":") + ["8", "2"])[:3]
        os.environ["ESLAM_PULL_TMA"] = tma; os.environ["ESLAM_PULL_STAGES"] = stages; os.environ["ESLAM_PULL_CTAS"] = ctas
        h = E.Handle(); h.set_supporting_plane(synth.GROUND_PLANE_WORLD)

        def pipelined(steps):
            pending = h.submit_ellipsoids(host, fo, bb, lb, tw)
            for _ in range(steps - 1):
                nxt = h.submit_ellipsoids(host, fo, bb, lb, tw)
                h.wait_ellipsoids(pending)
                pending = nxt
            return h.wait_ellipsoids(pending)
        pipelined(3)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rec = bytes(pipelined(args.steps))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            h.estimate_ellipsoids(host, fo, bb, lb, tw)
        dtb = time.perf_counter() - t0
        if ref is None:
            ref = rec
        print(json.dumps({"pull_tma": tma, "stages": stages, "ctas_per_sm": ctas, "rep": rep, "e2e_pipelined_fps": args.frames * args.steps / dt, "e2e_blocking_fps": args.frames * args.steps / dtb, "records_equal": rec == ref}))
        h.close()

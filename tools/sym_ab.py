"""A/B of the symmetry kernel on the bench workload (1 GPU): round-1 kernel vs the current one at several occupancies; records must be
byte-identical; optional in-kernel verification of every shortcut decision.   python tools/sym_ab.py [--frames 1024] [--verify]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import eslam_b200 as E  # noqa: E402
from eslam_b200 import synth  # noqa: E402
from bench import make_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--seed", type=int, default=1000)
ap.add_argument("--verify", action="store_true")
ap.add_argument("--variants", default="2:3,3:0:8:1,3:0:8:0,3:0:4:1", help="impl:occ[:dev_chunks[:graph]], impl 1 = round-1 kernel, 2 = single launch, 3 = round pipeline")
args = ap.parse_args()
depth, fo, bb, lb, tw = make_frames("one_box", args.frames, args.seed)
dev = torch.from_numpy(depth.view(np.int16)).cuda()
ref = None
for var in args.variants.split(","):
    impl, occ, chunks, graph = (var.split(":") + ["8", "1"])[:4]
    os.environ["ESLAM_SYM_GRAPH"] = graph
    os.environ["ESLAM_SYM_IMPL"] = impl
    os.environ["ESLAM_SYM_OCC"] = occ
    os.environ["ESLAM_DEV_CHUNKS"] = chunks
    h = E.Handle()
    h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    for _ in range(3):
        h.estimate_ellipsoids_dev(dev.data_ptr(), args.frames, 640, fo, bb, lb, tw, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        h.estimate_ellipsoids_dev(dev.data_ptr(), args.frames, 640, fo, bb, lb, tw, stream=stream.cuda_stream)
    e1.record(stream); torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 10
    rec = bytes(h.fetch_ellipsoids(len(fo), stream=stream.cuda_stream))
    h.set_profiling(True)
    h.estimate_ellipsoids_dev(dev.data_ptr(), args.frames, 640, fo, bb, lb, tw, stream=stream.cuda_stream); torch.cuda.synchronize()
    k_ms = h.stage_ms()
    h.set_profiling(False)
    c = h.batch_counters()
    if ref is None:
        ref = rec
    print(json.dumps({"impl": impl, "occ": occ, "chunks": chunks, "graph": graph, "launches_per_call": h.launches // 14, "step_ms": step_ms, "extract_ms": k_ms[0], "symmetry_ms": k_ms[1], "records_equal_first": rec == ref, "counters": c}))
    h.close()
if args.verify:
    os.environ["ESLAM_SYM_IMPL"] = "3"; os.environ["ESLAM_SYM_OCC"] = "0"; os.environ["ESLAM_SYM_VERIFY"] = "1"
    h = E.Handle()
    h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
    h.estimate_ellipsoids_dev(dev.data_ptr(), args.frames, 640, fo, bb, lb, tw); torch.cuda.synchronize()
    rec = bytes(h.fetch_ellipsoids(len(fo)))
    c = h.batch_counters()
    print(json.dumps({"verify": True, "records_equal_first": rec == ref, "verify_failures": c["verify_failures"], "counters": c}))

"""A/B of whole-step settings on the bench workload through the resident submit/collect path (two batches in flight, like bench.py's `value`):
   python tools/step_ab.py "ESLAM_EXTRACT_UQ=1" "ESLAM_SYM_WINDOW=4" ...   (each argument: comma-separated NAME=VALUE settings; "" = defaults)"""
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

frames, steps = 1024, 12
depth, fo, bb, lb, tw = make_frames("one_box", frames, 1000)
dev = torch.from_numpy(depth.view(np.int16)).cuda()
ref = None
for setting in (sys.argv[1:] or [""]):
    env = dict(kv.split("=") for kv in setting.split(",") if kv)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    h = E.Handle(); h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    rec = torch.empty(len(fo) * E.api.C.sizeof(E.api.EslamEllipsoidOut), dtype=torch.uint8, device="cuda")

    def run(n):
        t = h.dev_submit(dev.data_ptr(), frames, 640, fo, bb, lb, tw, stream=stream.cuda_stream)
        for _ in range(n - 1):
            t2 = h.dev_submit(dev.data_ptr(), frames, 640, fo, bb, lb, tw, stream=stream.cuda_stream)
            h.dev_collect(t, rec.data_ptr(), stream=stream.cuda_stream)
            t = t2
        h.dev_collect(t, rec.data_ptr(), stream=stream.cuda_stream)
    run(4); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); run(steps); e1.record(stream); torch.cuda.synchronize()
    got = rec.cpu().numpy().tobytes()
    if ref is None:
        ref = got
    print(json.dumps({"setting": setting, "ms_per_step": e0.elapsed_time(e1) / steps, "frames_per_s": frames * steps / (e0.elapsed_time(e1) / 1e3), "records_equal_first": got == ref}))
    h.close()

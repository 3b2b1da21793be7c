"""Secondary measurements for profiles/: ground-plane extraction (BASELINE.json configs[2] plane part), batched edge evaluation and
object BA (configs[3]: 500 keyframes x 50 edges).  Prints one JSON line per stage.  Needs a GPU.
  python tools/bench_stages.py [--plane-frames 64] [--reps 5] [--skip-cpu]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import eslam_b200 as E  # noqa: E402
from eslam_b200.api import EdgeGraph  # noqa: E402
from bench import make_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--plane-frames", type=int, default=592, help="two chunks of 2 x 148 frames")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--skip-cpu", action="store_true")
ap.add_argument("--only", default="all")
args = ap.parse_args()
h = E.Handle()
peak = 6584.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

if args.only in ("all", "plane"):
    F = args.plane_frames
    depth, fo, bb, lb, tw = make_frames("multi_box", F, 3000)
    dev = torch.from_numpy(depth.view(np.int16)).cuda()
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    for _ in range(2):
        h.extract_ground_plane_dev(dev.data_ptr(), F, 640, stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = h.launches
    e0.record(stream)
    for _ in range(args.reps):
        pl, found, ninl = h.extract_ground_plane_dev(dev.data_ptr(), F, 640, stream.cuda_stream)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    t0 = time.perf_counter(); h.extract_ground_plane(depth); e2e_ms = (time.perf_counter() - t0) * 1e3
    line = {"stage": "ground_plane", "frames": F, "ms_per_batch": ms, "frames_per_s": F / ms * 1e3, "e2e_frames_per_s": F / e2e_ms * 1e3, "found": int(found.sum()),
            "launches_per_batch": (h.launches - l0) // args.reps,
            "roofline": {"bound": "hbm", "algorithmic_bytes": 614400 * F, "achieved_GBps": 614400 * F / (ms / 1e3) / 1e9, "peak_GBps": peak,
                         "frac": 614400 * F / (ms / 1e3) / 1e9 / peak}}
    if not args.skip_cpu:
        import orc
        n = min(F, 8)
        t0 = time.perf_counter(); po, fo_, no = orc.extract_ground_plane(depth[:n]); cpu = n / (time.perf_counter() - t0)
        line["cpu_oracle_frames_per_s_1thread"] = cpu
        line["max_abs_plane_diff_vs_oracle"] = float(np.abs(pl[:n] - po).max()); line["inliers_equal"] = bool(np.array_equal(ninl[:n], no))
    print(json.dumps(line))

if args.only in ("all", "config3"):
    # BASELINE.json configs[2]: 8 bbox/frame, ground-plane extraction forced EVERY frame, symmetry on -- the per-frame flow of
    # Tracking::GrabPoseAndObjects (plane -> camera-to-world -> supporting plane -> every box), batched: one ground-plane pass and one
    # extraction pass over the same resident batch on one stream (the supporting plane of the batch is the constant synthetic one;
    # per-frame planes are compared with it in the tests).
    from eslam_b200 import synth
    F = args.plane_frames
    depth, fo, bb, lb, tw = make_frames("multi_box", F, 3000)
    dev = torch.from_numpy(depth.view(np.int16)).cuda()
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
    def step():
        h.extract_ground_plane_dev(dev.data_ptr(), F, 640, stream.cuda_stream)
        h.estimate_ellipsoids_dev(dev.data_ptr(), F, 640, fo, bb, lb, tw, stream=stream.cuda_stream)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = h.launches
    e0.record(stream)
    for _ in range(args.reps):
        step()
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    out = h.fetch_ellipsoids(len(fo), stream=stream.cuda_stream)
    print(json.dumps({"stage": "config3: ground plane every frame + 8 boxes/frame with symmetry", "frames": F, "objects": int(len(fo)), "ms_per_batch": ms,
                      "frames_per_s": F / ms * 1e3, "ok_fraction": sum(1 for o in out if o.status == 0) / float(len(fo)), "launches_per_batch": (h.launches - l0) // args.reps}))

if args.only in ("all", "edges"):
    import orc
    from helpers import make_ba_problem, K
    p = make_ba_problem(500, 50, 50)
    gg = EdgeGraph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], K)
    E_ = gg.c.n_edges
    h.edges_upload(gg, p["init"])
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    for _ in range(3):
        h.edges_eval_dev(stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        h.edges_eval_dev(stream.cuda_stream)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nt = np.bincount(np.array(p["type"]), minlength=3)
    alg_bytes = 496 * nt[0] + 944 * nt[1] + (80 + 32 + 8 + 8 + 72) * nt[2]
    line = {"stage": "edges_eval (residual + numeric Jacobian, one launch)", "edges": int(E_), "ms": ms, "edges_per_s": E_ / ms * 1e3,
            "roofline": {"bound": "hbm (reported; the kernel is fp64-ALU bound)", "algorithmic_bytes": int(alg_bytes), "achieved_GBps": alg_bytes / (ms / 1e3) / 1e9,
                         "peak_GBps": peak, "frac": alg_bytes / (ms / 1e3) / 1e9 / peak}}
    t0 = time.perf_counter(); eg, sg = h.optimize_objects(gg, p["init"], 10); lm_ms = (time.perf_counter() - t0) * 1e3
    line["lm_optimize10_ms"] = lm_ms; line["lm_iterations"] = sg.iterations; line["lm_trials"] = sg.trials_total; line["chi2"] = [sg.chi2_initial, sg.chi2_final]
    if not args.skip_cpu:
        go = orc.Graph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], K)
        t0 = time.perf_counter(); orc.edges_eval(go, p["init"]); line["cpu_oracle_edges_eval_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); eo, so = orc.optimize_objects(go, p["init"], 10); line["cpu_oracle_lm_ms"] = (time.perf_counter() - t0) * 1e3
        line["cpu_lm_iterations"] = so.iterations; line["cpu_chi2_final"] = so.chi2_final
        line["max_rel_ellipsoid_diff"] = float(np.max(np.abs(eg - eo) / np.maximum(1.0, np.abs(eo))))
    print(json.dumps(line))

if args.only in ("all", "slam"):
    # SLAM mode on the configs[3] graph: 499 free poses + 50 ellipsoids = 3444 unknowns, 25 050 + 499 edges
    import orc
    from helpers import make_slam_problem, K
    p = make_slam_problem(500, 50, 50, pose_noise=0.005)
    gg = EdgeGraph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], K)
    h.optimize_slam(gg, p["init"], 2, odom_meas=p["odom"])          # allocations + first launches
    l0 = h.launches
    t0 = time.perf_counter(); eg, Tg, sg = h.optimize_slam(gg, p["init"], 10, odom_meas=p["odom"]); ms = (time.perf_counter() - t0) * 1e3
    line = {"stage": "slam_mode optimize(10) (free poses + odometry edges)", "unknowns": 6 * 499 + 9 * 50, "edges": int(gg.c.n_edges) + 499, "ms": ms,
            "iterations": sg.iterations, "trials": sg.trials_total, "chi2": [sg.chi2_initial, sg.chi2_final], "launches": int(h.launches - l0)}
    if not args.skip_cpu:
        go = orc.Graph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], K)
        t0 = time.perf_counter(); eo, To, so = orc.optimize_slam(go, p["init"], 10, odom_meas=p["odom"]); line["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        line["cpu_iterations"] = so.iterations; line["cpu_trials"] = so.trials_total; line["cpu_chi2_final"] = so.chi2_final
        line["max_rel_ellipsoid_diff"] = float(np.max(np.abs(eg - eo) / np.maximum(1.0, np.abs(eo))))
        sgn = np.sign(np.sum(Tg[:, 3:] * To[:, 3:], axis=1))[:, None]
        Tg2 = Tg.copy(); Tg2[:, 3:] *= sgn
        line["max_rel_pose_diff"] = float(np.max(np.abs(Tg2 - To) / np.maximum(1.0, np.abs(To))))
    print(json.dumps(line))

#!/usr/bin/env python
"""bench.py -- frames/sec of the per-frame quadric-estimation hot path (BASELINE.json metric) on synthetic TUM-shaped input.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl ours|reference] [--workload one_box|multi_box] [--no-stages]

A "step" is one pass of the hot path (EllipsoidExtractor::EstimateLocalEllipsoid for every bbox: crop -> voxel grid -> plane
filter -> clustering -> PCA -> 9-seed symmetry LM -> completion -> ellipsoid) over one batch of B synthetic 640x480 frames
(BASELINE.json configs[1]: 1 bbox/frame, label 28 => dual-reflection symmetry).  The batch (B x 614 400 B of depth) is larger than
the 126 MB L2, so every step re-reads its input from HBM.
  value          frames/s with the depth batch already resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e            frames/s through the reference-facing C-ABI call with HOST (pinned) buffers: H2D of the depth pixels the path reads
                 (the union rectangle of each frame's boxes), kernels and D2H of the EslamEllipsoidOut records inside the timed region
  roofline       dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline   the CPU oracle (oracle/liboracle.so, a single-threaded restatement of the reference: kind "port") on a bounded sample
--impl reference times the same oracle on all host threads (the reference cannot be built in this image: no Eigen/PCL/OpenCV).
  stages         (N = 1) the other BASELINE.json configs, each device-timed in the same run with its own cpu_baseline: ground plane
                 (P1/P2, own roofline), configs[2] (plane every frame + 8 boxes/frame), configs[3] (edge evaluation, optimize(10) in
                 mapping and SLAM mode); and the pageable-input e2e beside the pinned one
Multi-GPU: one process per GPU (torchrun), independent frames per rank (weak scaling, no data-path collective); per-rank step times
(min / median / max) and the largest symmetry cloud of every rank are reported so stragglers are visible.  configs[4] ("gather of
map"): a step variant whose timed region also all-gathers the per-bbox records device-to-device over NCCL (config.config5).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "object-oriented-slam_b200")]

import numpy as np  # noqa: E402


def _gen_one(args):
    from eslam_b200 import synth
    kind, f, seed_base = args
    if kind == "one_box":
        return synth.scene_one_box(f, seed_base)
    return synth.scene_multi_box(f, 8, seed_base)


def make_frames(kind, n_frames, seed_base):
    """Parallel synthetic-frame generation (numpy ray casting is ~0.2 s/frame single threaded)."""
    from concurrent.futures import ProcessPoolExecutor
    workers = max(1, min(os.cpu_count() or 1, 64, n_frames))
    jobs = [(kind, f, seed_base) for f in range(n_frames)]
    if workers == 1:
        res = [_gen_one(j) for j in jobs]
    else:
        with ProcessPoolExecutor(workers) as ex:
            res = list(ex.map(_gen_one, jobs, chunksize=max(1, n_frames // (workers * 4))))
    depth = np.ascontiguousarray(np.stack([r[0] for r in res]))
    fo, bb, lb = [], [], []
    for f, r in enumerate(res):
        for k in range(len(r[1])):
            fo.append(f); bb.append(r[1][k]); lb.append(r[2][k])
    tw = np.ascontiguousarray(np.stack([r[3] for r in res]))
    return depth, np.asarray(fo, np.int32), np.ascontiguousarray(np.asarray(bb, np.float64)), np.asarray(lb, np.int32), tw


class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region (B200_PROFILING.md clocks line).

    NVML is polled from a thread every 10 ms (the timed region of a default run is only a few hundred ms, shorter than
    nvidia-smi's start-up); if NVML cannot be opened the recipe's nvidia-smi loop is used instead."""

    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.sm, self.mx, self.reasons, self.stop_flag, self.thread, self.nv = [], 0, set(), False, None, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll(self):
        nv, hd = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(hd)
                for name, const in self.REASONS:
                    if mask & getattr(nv, const):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            self.nv = self._nvml_handle()
            self.mx = float(self.nv[0].nvmlDeviceGetMaxClockInfo(self.nv[1], self.nv[0].NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=1)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = list(self.sm), self.mx, set(self.reasons)
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx = max(mx, float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nv else "nvidia-smi"}


def oracle_fps(depth, fo, bb, lb, tw, plane, n_frames, threads):
    """frames/s of the CPU oracle on the first n_frames frames using `threads` host threads (ctypes drops the GIL)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    orc.lib()
    n_frames = min(n_frames, depth.shape[0])
    chunks = np.array_split(np.arange(n_frames), threads)

    def work(idx):
        if len(idx) == 0:
            return
        f0, f1 = idx[0], idx[-1] + 1
        sel = (fo >= f0) & (fo < f1)
        orc.estimate_ellipsoids(depth[f0:f1], fo[sel] - f0, bb[sel], lb[sel], tw[f0:f1], plane)

    t0 = time.perf_counter()
    if threads == 1:
        work(chunks[0])
    else:
        ts = [threading.Thread(target=work, args=(c,)) for c in chunks]
        [t.start() for t in ts]
        [t.join() for t in ts]
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def _threads_available():
    threads = os.cpu_count() or 1
    try:
        threads = min(threads, len(os.sched_getaffinity(0)))      # CPUs this process may actually run on
    except Exception:
        pass
    quota = "none"
    try:   # a cgroup CPU quota caps the useful parallelism below the thread count (reported, not enforced)
        mx, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if mx != "max":
            quota = "%.1f CPUs" % (float(mx) / float(per))
    except Exception:
        pass
    return threads, quota


def _peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def _ncu_traffic(kernel):
    """DRAM bytes per unit of `kernel` from the committed ncu --set full capture of the current round (None when absent)."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            k = tj["kernels"][kernel]
            return k, "profiles/" + name
        except Exception:
            continue
    return None, None


def _event_ms(torch, stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_stages(args, h, torch, stream, plane, peak, peak_src):
    """The other BASELINE.json configs, device-timed in this run (N = 1 only); every stage carries its own cpu_baseline."""
    import eslam_b200 as E
    from eslam_b200 import synth
    from eslam_b200.api import EdgeGraph
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sptr = stream.cuda_stream
    stages = []
    reps = max(3, min(args.steps, 5))
    # ---------------- ground plane (P1/P2) + configs[2]: plane every frame + 8 boxes/frame, symmetry on
    F = args.stage_frames
    depth, fo, bb, lb, tw = make_frames("multi_box", F, 3000)
    host = torch.from_numpy(depth.view(np.int16)).pin_memory()
    dev = host.to("cuda", non_blocking=True)
    torch.cuda.synchronize()
    host_np = host.numpy().view(np.uint16)
    res = {}

    def plane_step():
        res["plane"] = h.extract_ground_plane_dev(dev.data_ptr(), F, 640, sptr)
    for _ in range(2):
        plane_step()
    torch.cuda.synchronize()
    l0 = h.launches
    ms = _event_ms(torch, stream, plane_step, reps)
    pl, found, ninl = res["plane"]
    alg = 614400.0 * F
    tr, tr_src = _ncu_traffic("plane_stage")
    h.extract_ground_plane(host_np)                                      # first call: allocates the library's staging buffer
    e2e_runs = []
    for _ in range(3):
        t0 = time.perf_counter(); h.extract_ground_plane(host_np); e2e_runs.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = float(np.median(e2e_runs))
    import orc
    n_cpu = min(F, 16)
    t0 = time.perf_counter(); po, fo_, no = orc.extract_ground_plane(depth[:n_cpu]); cpu_dt = time.perf_counter() - t0
    stages.append({"stage": "ground_plane", "config": "PlaneExtractor::extractGroundPlane on every frame (P1/P2), %d frames per batch" % F,
                   "value": F / ms * 1e3, "unit": "frames/s", "ms_per_batch": ms, "gpu_launches_per_batch": (h.launches - l0) // reps,
                   "found_fraction": float(found.mean()),
                   "e2e": {"value": F / e2e_ms * 1e3, "unit": "frames/s", "h2d_bytes_per_batch": int(depth.nbytes), "d2h_bytes_per_batch": 40 * F,
                           "api": "eslam_extract_ground_plane_host (whole frames travel: the plane stage reads every pixel)"},
                   "roofline": {"kernel": "plane stage (13 launches per chunk of 296 frames)", "bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": alg / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_batch": alg,
                                "traffic": (tr["dram_bytes_per_frame"] * F if tr else None), "traffic_source": tr_src, "peak_source": peak_src},
                   "cpu_baseline": {"value": n_cpu / cpu_dt, "unit": "frames/s", "cores": 1, "kind": "port", "sample": "%d frames of the same batch, %.1f s" % (n_cpu, cpu_dt)},
                   "parity_in_run": {"max_abs_plane_diff_vs_oracle": float(np.abs(pl[:n_cpu] - po).max()), "inliers_equal": bool(np.array_equal(ninl[:n_cpu], no)),
                                     "found_equal": bool(np.array_equal(found[:n_cpu], fo_))}})
    # configs[2]
    h.set_supporting_plane(plane)

    def c2_step():
        h.extract_ground_plane_dev(dev.data_ptr(), F, 640, sptr)
        h.estimate_ellipsoids_dev(dev.data_ptr(), F, 640, fo, bb, lb, tw, stream=sptr)
    for _ in range(2):
        c2_step()
    torch.cuda.synchronize()
    l0 = h.launches
    ms = _event_ms(torch, stream, c2_step, reps)
    launches = (h.launches - l0) // reps
    out = h.fetch_ellipsoids(len(fo), stream=sptr)

    def c2_e2e():
        h.extract_ground_plane(host_np)
        return h.estimate_ellipsoids(host_np, fo, bb, lb, tw)[0]
    c2_e2e()
    t0 = time.perf_counter()
    for _ in range(reps):
        c2_e2e()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
    n_cpu = min(F, 8)
    sel = fo < n_cpu
    t0 = time.perf_counter()
    orc.extract_ground_plane(depth[:n_cpu]); orc.estimate_ellipsoids(depth[:n_cpu], fo[sel], bb[sel], lb[sel], tw[:n_cpu], plane)
    cpu_dt = time.perf_counter() - t0
    stages.append({"stage": "config2", "config": "BASELINE.json configs[2]: synthetic 640x480 depth, 8 bbox/frame, ground plane extracted for every frame + symmetry on, %d frames (%d objects) per batch" % (F, len(fo)),
                   "value": F / ms * 1e3, "unit": "frames/s", "ms_per_batch": ms, "gpu_launches_per_batch": launches,
                   "ok_fraction": sum(1 for o in out if o.status == 0) / float(len(fo)),
                   "e2e": {"value": F / e2e_ms * 1e3, "unit": "frames/s", "h2d_bytes_per_batch": int(depth.nbytes), "d2h_bytes_per_batch": 40 * F + int(C.sizeof(out[0])) * len(fo),
                           "api": "eslam_extract_ground_plane_host + eslam_estimate_ellipsoids_host (pinned host buffers)"},
                   "cpu_baseline": {"value": n_cpu / cpu_dt, "unit": "frames/s", "cores": 1, "kind": "port", "sample": "%d frames (%d boxes) of the same batch, %.1f s" % (n_cpu, int(sel.sum()), cpu_dt)}})
    del dev, host
    # ---------------- configs[3]: 500 keyframes x 50 ellipsoid edges
    p = synth.ba_problem(500, 50, 50)
    gg = EdgeGraph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], p["K"])
    E_ = int(gg.c.n_edges)
    h.edges_upload(gg, p["init"])
    for _ in range(3):
        h.edges_eval_dev(sptr)
    torch.cuda.synchronize()
    ms = _event_ms(torch, stream, lambda: h.edges_eval_dev(sptr), 20)
    nt = np.bincount(np.asarray(p["type"]), minlength=3)
    alg = float(496 * nt[0] + 944 * nt[1] + 200 * nt[2])
    t0 = time.perf_counter(); r_, J_, c_ = h.edges_eval(gg, p["init"]); ee_ms = (time.perf_counter() - t0) * 1e3
    go = orc.Graph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], p["K"])
    t0 = time.perf_counter(); ro, Jo, co = orc.edges_eval(go, p["init"]); cpu_ms = (time.perf_counter() - t0) * 1e3
    tr, tr_src = _ncu_traffic("edge_eval_kernel")
    stages.append({"stage": "config3_edge_eval", "config": "BASELINE.json configs[3]: 500 keyframes x 50 ellipsoid edges (+50 gravity priors), residuals + numeric 9-DoF Jacobians in one launch",
                   "value": E_ / ms * 1e3, "unit": "edges/s", "ms_per_batch": ms, "edges": E_, "gpu_launches_per_batch": 1,
                   "e2e": {"value": E_ / ee_ms * 1e3, "unit": "edges/s", "api": "eslam_edges_eval_host (graph up, residuals + Jacobians + chi2 down)",
                           "h2d_bytes_per_batch": int(E_ * (12 + 88) + p["Tcw"].nbytes + p["init"].nbytes), "d2h_bytes_per_batch": int(E_ * (72 + 648 + 8))},
                   "roofline": {"kernel": "edge_eval_kernel", "bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms / 1e3) / 1e9 / peak,
                                "algorithmic_bytes_per_batch": alg, "traffic": (tr["dram_bytes_per_launch"] if tr else None), "traffic_source": tr_src, "peak_source": peak_src,
                                "note": "fp64-ALU bound (19 residual evaluations per edge), not bandwidth bound"},
                   "cpu_baseline": {"value": E_ / cpu_ms * 1e3, "unit": "edges/s", "cores": 1, "kind": "port", "sample": "the same graph once, %.2f s" % (cpu_ms / 1e3)},
                   "parity_in_run": {"max_abs_residual_diff": float(np.abs(r_ - ro).max()), "max_rel_chi2_diff": float(np.max(np.abs(c_ - co) / np.maximum(1.0, np.abs(co))))}})
    t0 = time.perf_counter(); h.optimize_objects(gg, p["init"], 10); lm_first_ms = (time.perf_counter() - t0) * 1e3   # first call of this shape: plain launches
    h.optimize_objects(gg, p["init"], 10)                                                                                 # second call: captures the per-iteration graph
    l0 = h.launches
    t0 = time.perf_counter(); eg, sg = h.optimize_objects(gg, p["init"], 10); lm_ms = (time.perf_counter() - t0) * 1e3
    lm_launches = h.launches - l0
    t0 = time.perf_counter(); eo, so = orc.optimize_objects(go, p["init"], 10); cpu_ms = (time.perf_counter() - t0) * 1e3
    stages.append({"stage": "config3_optimize_mapping", "config": "Optimizer::GlobalObjectGraphOptimization, graph.optimize(10), fixed poses (mapping mode), same graph",
                   "value": lm_ms, "unit": "ms per optimize(10)", "higher_is_better": False, "first_call_ms": lm_first_ms, "iterations": sg.iterations, "trials": sg.trials_total, "gpu_launches_per_batch": int(lm_launches),
                   "chi2": [sg.chi2_initial, sg.chi2_final],
                   "e2e": {"value": lm_ms, "unit": "ms per optimize(10)", "api": "eslam_optimize_objects_host (host graph in, ellipsoids out; wall clock)"},
                   "cpu_baseline": {"value": cpu_ms, "unit": "ms per optimize(10)", "cores": 1, "kind": "port", "sample": "the same graph, optimize(10), %d iterations" % so.iterations},
                   "parity_in_run": {"max_rel_ellipsoid_diff": float(np.max(np.abs(eg - eo) / np.maximum(1.0, np.abs(eo)))), "iterations_equal": bool(sg.iterations == so.iterations)}})
    ps = synth.ba_problem(500, 50, 50, pose_noise=0.005)
    gs = EdgeGraph(ps["type"], ps["frame"], ps["object"], ps["meas"], ps["info"], ps["Tcw"], ps["n_objects"], ps["K"])
    h.optimize_slam(gs, ps["init"], 2, odom_meas=ps["odom"])
    l0 = h.launches
    t0 = time.perf_counter(); eg, Tg, sg = h.optimize_slam(gs, ps["init"], 10, odom_meas=ps["odom"]); sl_ms = (time.perf_counter() - t0) * 1e3
    sl_launches = h.launches - l0
    gso = orc.Graph(ps["type"], ps["frame"], ps["object"], ps["meas"], ps["info"], ps["Tcw"], ps["n_objects"], ps["K"])
    t0 = time.perf_counter(); eo, To, so = orc.optimize_slam(gso, ps["init"], 2, odom_meas=ps["odom"]); cpu_ms = (time.perf_counter() - t0) * 1e3
    eg2, Tg2, sg2 = h.optimize_slam(gs, ps["init"], 2, odom_meas=ps["odom"])
    stages.append({"stage": "config3_optimize_slam", "config": "same graph with bSLAM_mode: 499 free poses + odometry edges, 3444 unknowns, graph.optimize(10)",
                   "value": sl_ms / max(1, sg.trials_total), "unit": "ms per LM trial", "higher_is_better": False, "ms_per_optimize10": sl_ms, "iterations": sg.iterations, "trials": sg.trials_total,
                   "gpu_launches_per_batch": int(sl_launches), "chi2": [sg.chi2_initial, sg.chi2_final],
                   "e2e": {"value": sl_ms / max(1, sg.trials_total), "unit": "ms per LM trial", "api": "eslam_optimize_slam_host (wall clock)"},
                   "cpu_baseline": {"value": cpu_ms / max(1, so.trials_total), "unit": "ms per LM trial", "cores": 1, "kind": "port",
                                    "sample": "the same graph, optimize(2): %d trials in %.1f s (dense LDLT of the full system like the reference)" % (so.trials_total, cpu_ms / 1e3)},
                   "parity_in_run": {"max_rel_ellipsoid_diff_after_2_iterations": float(np.max(np.abs(eg2 - eo) / np.maximum(1.0, np.abs(eo)))),
                                     "trials_equal": bool(sg2.trials_total == so.trials_total)}})
    return stages


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=1024, help="frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="one_box", choices=["one_box", "multi_box"])
    ap.add_argument("--cpu-sample", type=int, default=1024, help="frames timed for cpu_baseline (1 thread)")
    ap.add_argument("--stage-frames", type=int, default=296, help="frames per batch of the ground-plane / configs[2] stages (one chunk of 2 x 148)")
    ap.add_argument("--no-stages", action="store_true", help="skip the secondary stages (ground plane, configs[2], configs[3])")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from eslam_b200 import synth
    plane = synth.GROUND_PLANE_WORLD
    workload_name = ("synthetic 640x480 depth, 1 bbox/frame, ~10k stride-3 points segmented per box, label 28 (dual symmetry)" if args.workload == "one_box"
                     else "synthetic 640x480 depth, 8 bbox/frame, labels 28/41/62")
    config = {"workload": workload_name, "frames_per_step_per_gpu": args.frames, "image": "640x480 uint16", "l2_policy": "inputs larger than L2 (batch >= 600 MB of depth)",
              "symmetry": "on", "plane": "supporting plane given (headline = configs[1]); ground plane, configs[2] and configs[3] are timed in `stages`"}
    metric = "frames/sec (640x480 depth, N objects)"

    # ------------------------------------------------------------------ reference arm: CPU oracle on all host threads
    if args.impl == "reference":
        if rank != 0:
            return
        threads, quota = _threads_available()
        sample = args.frames          # the same batch as our arm (same_config)
        depth, fo, bb, lb, tw = make_frames(args.workload, sample, 1000)
        for _ in range(args.warmup):
            oracle_fps(depth, fo, bb, lb, tw, plane, min(sample, 2 * threads), threads)
        tot = 0.0
        for _ in range(args.steps):
            _, dt = oracle_fps(depth, fo, bb, lb, tw, plane, sample, threads)
            tot += dt
        fps = sample * args.steps / tot
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config, frames_per_step_per_gpu=sample, warmup_note="warm-up steps run on the first %d frames" % min(sample, 2 * threads)),
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                                 "sample": f"{sample} frames per step, {threads} host threads (cgroup cpu quota: {quota}), oracle/liboracle.so (the reference itself cannot be built here: needs Eigen/PCL/OpenCV)"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import eslam_b200 as E
    from eslam_b200._structs import EslamEllipsoidOut
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line (rank 0).  NCCL prints its version banner (NCCL_DEBUG=VERSION is set on the GPU boxes) with
        # a plain write to fd 1 when the communicator is created, whatever NCCL_DEBUG_FILE says (measured), so fd 1 points at stderr
        # while the process group and its first collective are set up.
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    depth, fo, bb, lb, tw = make_frames(args.workload, args.frames, 1000 + 100000 * rank)
    n_obj = len(fo)
    h = E.Handle(device=local)
    h.set_supporting_plane(plane)
    stream = torch.cuda.Stream()          # explicit non-default stream: kernels AND the timing events live on it
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    host_depth = torch.from_numpy(depth.view(np.int16)).pin_memory()      # pinned host copy (e2e leg)
    dev_depth = host_depth.to("cuda", non_blocking=True)
    torch.cuda.synchronize()
    dptr = dev_depth.data_ptr()
    pitch = depth.shape[2]
    rec_bytes = C.sizeof(EslamEllipsoidOut)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        h.estimate_ellipsoids_dev(dptr, args.frames, pitch, fo, bb, lb, tw, stream=sptr)

    host_np = host_depth.numpy().view(np.uint16)

    def step_e2e(src=None):
        return h.estimate_ellipsoids(host_np if src is None else src, fo, bb, lb, tw)[0]

    def run_e2e_pipelined(steps):
        # the streaming form of the same call: batch k+1 is submitted before the records of batch k are collected (two in flight);
        # every step still uploads its own input from pinned host memory and reads its own records back
        pending = h.submit_ellipsoids(host_np, fo, bb, lb, tw)
        for _ in range(steps - 1):
            nxt = h.submit_ellipsoids(host_np, fo, bb, lb, tw)
            h.wait_ellipsoids(pending)
            pending = nxt
        return h.wait_ellipsoids(pending)

    # ---- resident (value): the depth batch stays in HBM, two batches in flight (eslam_estimate_ellipsoids_dev_submit / _collect: the
    # records of every step are collected device-to-device on the launching stream, which also carries the timing events)
    recbuf = torch.empty(n_obj * rec_bytes, dtype=torch.uint8, device="cuda")

    def run_resident(steps):
        pending = h.dev_submit(dptr, args.frames, pitch, fo, bb, lb, tw, stream=sptr)
        for _ in range(steps - 1):
            nxt = h.dev_submit(dptr, args.frames, pitch, fo, bb, lb, tw, stream=sptr)
            h.dev_collect(pending, recbuf.data_ptr(), stream=sptr)
            pending = nxt
        h.dev_collect(pending, recbuf.data_ptr(), stream=sptr)
    run_resident(args.warmup)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = h.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_resident(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = h.launches - l0
    clocks = sampler.stop()
    # the same steps one at a time (blocking form, one event per step): the spread between steps, visible per rank
    nb = max(3, min(args.steps, 10))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(nb + 1)]
    evs[0].record(stream)
    for k in range(nb):
        step_resident()
        evs[k + 1].record(stream)
    torch.cuda.synchronize()
    step_ms = np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(nb)])
    out = h.fetch_ellipsoids(n_obj, stream=sptr)
    if bytes(out) != recbuf.cpu().numpy().tobytes():
        raise SystemExit("bench: the pipelined device-resident records differ from the blocking call's")
    ok_frac = sum(1 for o in out if o.status == 0) / float(n_obj)
    counters = h.batch_counters()
    # per-kernel times (one extra profiled batch: single stream, kernels back to back)
    h.set_profiling(True)
    step_resident(); torch.cuda.synchronize()
    k_ms = h.stage_ms()
    h.set_profiling(False)
    # ---- e2e
    for _ in range(2):
        step_e2e()
    barrier()
    h2d0 = h.h2d_bytes
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d_per_step = (h.h2d_bytes - h2d0) // args.steps    # counted by the library from the copies it issued
    run_e2e_pipelined(2)
    barrier()
    t0 = time.perf_counter()
    pip = run_e2e_pipelined(args.steps)
    torch.cuda.synchronize()
    e2e_pip_s = time.perf_counter() - t0
    if bytes(pip) != bytes(step_e2e()):                   # the pipelined path must return the records of the blocking path
        raise SystemExit("bench: pipelined submit/wait records differ from the blocking call")
    # pageable host input (what a drop-in caller with a plain cv::Mat hands over): batched-DMA upload path
    pg_steps = max(2, min(args.steps, 5))
    step_e2e(depth)
    barrier()
    h2d0 = h.h2d_bytes
    t0 = time.perf_counter()
    for _ in range(pg_steps):
        pg = step_e2e(depth)
    torch.cuda.synchronize()
    e2e_pg_s = (time.perf_counter() - t0) * args.steps / pg_steps
    h2d_pg = (h.h2d_bytes - h2d0) // pg_steps
    if bytes(pg) != bytes(pip):
        raise SystemExit("bench: pageable-input records differ from the pinned-input records")
    # ---- configs[4]: the same resident step + device-to-device all-gather of the per-bbox records ("gather of map") inside the timed region
    c5_ms = None
    gathered_ok = None
    if world > 1:
        mine = torch.empty(n_obj * rec_bytes, dtype=torch.uint8, device="cuda")
        allrec = torch.empty(world * n_obj * rec_bytes, dtype=torch.uint8, device="cuda")

        from eslam_b200 import shard

        def step_c5():
            step_resident()
            shard.gather_records_dev(h, n_obj, stream_ptr=sptr, out=allrec)   # D2D copy + ncclAllGather on the launching stream, no host round trip
        for _ in range(3):
            step_c5()
        barrier()
        c5_ms = _event_ms(torch, stream, step_c5, args.steps) * args.steps
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record(stream)
        h.copy_ellipsoids_dev(n_obj, mine.data_ptr(), stream=sptr)
        dist.all_gather_into_tensor(allrec, mine)
        g1.record(stream)
        torch.cuda.synchronize()
        gather_only_ms = g0.elapsed_time(g1)
        blk = allrec.view(world, n_obj * rec_bytes)[rank].cpu().numpy().tobytes()
        gathered_ok = blk == bytes(h.fetch_ellipsoids(n_obj, stream=sptr))
    # ---- reduce over ranks
    t = torch.tensor([ms, e2e_s * 1e3, e2e_pip_s * 1e3, e2e_pg_s * 1e3, c5_ms or 0.0], device="cuda", dtype=torch.float64)
    mine_stats = torch.tensor([ms / args.steps, float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max()), float(counters["max_n_sym"]), k_ms[0], k_ms[1]],
                              device="cuda", dtype=torch.float64)
    all_stats = [mine_stats]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        all_stats = [torch.zeros_like(mine_stats) for _ in range(world)]
        dist.all_gather(all_stats, mine_stats)
    ms, e2e_ms, e2e_pip_ms, e2e_pg_ms, c5_ms_max = (float(x) for x in t)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    per_rank = [{"rank": r, "ms_per_step": float(s[0]), "blocking_step_ms_min": float(s[1]), "blocking_step_ms_median": float(s[2]), "blocking_step_ms_max": float(s[3]), "max_n_sym": int(s[4]),
                 "extract_kernel_ms": float(s[5]), "symmetry_kernel_ms": float(s[6])} for r, s in enumerate(all_stats)]
    frames_total = args.frames * world * args.steps
    value = frames_total / (ms / 1e3)
    e2e_sync = frames_total / (e2e_ms / 1e3)
    e2e = frames_total / (e2e_pip_ms / 1e3)
    e2e_pageable = frames_total / (e2e_pg_ms / 1e3)
    # ---- roofline of the dominant kernel (algorithmic bytes per SURVEY.md 8(d), counted by the library for this batch)
    peak, peak_src = _peak()
    crop_bytes = float(sum(2 * ((int(b[2]) - int(b[0]) + 2) // 3) * ((int(b[3]) - int(b[1]) + 2) // 3) for b in bb))   # 2 B per stride-3 sample
    sym_bytes = 14.0 * counters["ref_points"]   # 12 B float point + 2 B depth pixel per mirrored point of every cost evaluation of the reference's call sequence
    if k_ms[1] >= k_ms[0]:
        dom, alg, t_ms = "symmetry_stage", sym_bytes, k_ms[1]
    else:
        dom, alg, t_ms = "extract_kernel", crop_bytes, k_ms[0]
    achieved = alg / (t_ms / 1e3) / 1e9
    tr, tr_src = _ncu_traffic(dom)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (float(tr["dram_bytes_per_object"]) * n_obj if tr else None),
                "traffic_source": (tr_src + " (dram__bytes_read.sum + dram__bytes_write.sum per object x objects per launch; not re-measured in this run)") if tr else None,
                "peak_source": peak_src, "kernel_ms": {"extract_kernel": k_ms[0], "symmetry_stage": k_ms[1]},
                "kernels_of_stage": {"symmetry_stage": "sym_step_kernel(0), 56 x { sym_eval_kernel(r), sym_step_kernel(r + 1) }, sym_final_kernel: the round pipeline of the symmetry search "
                                                       "(launches past the last running seed exit at once), replayed as one CUDA graph per chunk"},
                "algorithmic_bytes_per_launch": alg, "work_counters": counters,
                "note": "kernel_ms: CUDA events on the launching stream of one profiled batch (one stream, launches back to back); algorithmic bytes of the symmetry stage = 14 B x mirrored points "
                        "visited by the reference's cost evaluations, counted by the kernels for this batch.  Both stages are latency / instruction-issue bound (per-object working set << L2), see DESIGN.md"}
    # ---- cpu baseline (1 thread, bounded sample): rank 0 at N = 1 only
    cpu_baseline = None
    stages = None
    if world == 1:
        cpu_fps, cpu_dt = oracle_fps(depth, fo, bb, lb, tw, plane, min(args.cpu_sample, args.frames), 1)
        cpu_baseline = {"value": cpu_fps, "unit": "frames/s", "cores": 1, "kind": "port",
                        "sample": f"{min(args.cpu_sample, args.frames)} frames of the same batch, {cpu_dt:.1f} s, oracle/liboracle.so single thread"}
        if not args.no_stages:
            del dev_depth
            stages = run_stages(args, h, torch, stream, plane, peak, peak_src)
    cfg = dict(config, objects_per_step_per_gpu=n_obj, ok_fraction=ok_frac, per_rank=per_rank)
    if world > 1:
        cfg["config5"] = {"what": "BASELINE.json configs[4]: the resident step + D2D copy of the records + ncclAllGather (all_gather_into_tensor) on the launching stream, inside the timed region",
                          "value": frames_total / (c5_ms_max / 1e3), "unit": "frames/s", "ms_per_step": c5_ms_max / args.steps, "gather_only_ms": gather_only_ms,
                          "gathered_bytes_per_step": world * n_obj * rec_bytes, "records_match_local_fetch": gathered_ok}
    line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(cfg, resident_api="eslam_estimate_ellipsoids_dev_submit / _collect, two batches in flight, records collected device-to-device every step"),
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_per_step) * world, "d2h_bytes_per_step": int(rec_bytes * n_obj) * world,
                    "api": "eslam_estimate_ellipsoids_submit / _wait, two batches in flight (host pinned buffers in, records out, every step)",
                    "single_blocking_call_value": e2e_sync,
                    "pageable_input": {"value": e2e_pageable, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_pg) * world,
                                       "api": "eslam_estimate_ellipsoids_host on a pageable numpy array (box rows staged into pinned memory by host threads, then pulled by the bulk-copy kernel)"}},
            "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline}
    if stages is not None:
        line["stages"] = stages
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

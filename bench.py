#!/usr/bin/env python
"""bench.py -- frames/sec of the per-frame quadric-estimation hot path (BASELINE.json metric) on synthetic TUM-shaped input.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl ours|reference] [--workload one_box|multi_box]

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
Multi-GPU: one process per GPU (torchrun), independent frames per rank (weak scaling, no data-path collective); after the timed
region the per-frame result records are all-gathered once over NCCL (BASELINE.json configs[4] "gather of map"), timed separately.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=1024, help="frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="one_box", choices=["one_box", "multi_box"])
    ap.add_argument("--cpu-sample", type=int, default=1024, help="frames timed for cpu_baseline (1 thread)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from eslam_b200 import synth
    plane = synth.GROUND_PLANE_WORLD
    workload_name = ("synthetic 640x480 depth, 1 bbox/frame, ~10k stride-3 points segmented per box, label 28 (dual symmetry)" if args.workload == "one_box"
                     else "synthetic 640x480 depth, 8 bbox/frame, labels 28/41/62")
    config = {"workload": workload_name, "frames_per_step_per_gpu": args.frames, "image": "640x480 uint16", "l2_policy": "inputs larger than L2 (batch >= 600 MB of depth)",
              "symmetry": "on", "plane": "supporting plane given (extraction path); ground-plane kernels benchmarked separately in profiles/"}
    metric = "frames/sec (640x480 depth, N objects)"

    # ------------------------------------------------------------------ reference arm: CPU oracle on all host threads
    if args.impl == "reference":
        if rank != 0:
            return
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
        sample = min(args.frames, 16 * threads, 2048)
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
                "config": dict(config, frames_per_step_per_gpu=sample),
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        h.estimate_ellipsoids_dev(dptr, args.frames, pitch, fo, bb, lb, tw, stream=sptr)

    host_np = host_depth.numpy().view(np.uint16)

    def step_e2e():
        return h.estimate_ellipsoids(host_np, fo, bb, lb, tw)[0]

    def run_e2e_pipelined(steps):
        # the streaming form of the same call: batch k+1 is submitted before the records of batch k are collected (two in flight);
        # every step still uploads its own input from pinned host memory and reads its own records back
        pending = h.submit_ellipsoids(host_np, fo, bb, lb, tw)
        res = None
        for _ in range(steps - 1):
            nxt = h.submit_ellipsoids(host_np, fo, bb, lb, tw)
            res = h.wait_ellipsoids(pending)
            pending = nxt
        return h.wait_ellipsoids(pending)

    # ---- resident (value)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = h.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = h.launches - l0
    clocks = sampler.stop()
    out = h.fetch_ellipsoids(n_obj, stream=sptr)
    ok_frac = sum(1 for o in out if o.status == 0) / float(n_obj)
    # per-kernel times (one extra profiled batch)
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
    pip_same = bytes(pip) == bytes(step_e2e())            # the pipelined path returns the same records (reported, not fatal)
    t = torch.tensor([ms, e2e_s * 1e3, e2e_pip_s * 1e3], device="cuda", dtype=torch.float64)
    gather_ms = None
    if world > 1:
        from eslam_b200 import shard
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize(); g0 = time.perf_counter()
        # the one exchange step of the path: fixed-size per-(frame,bbox) records, all-gathered over NCCL ("gather of map")
        allrec = shard.gather_records(out, np.arange(n_obj) + rank * n_obj, n_obj * world, device="cuda")
        torch.cuda.synchronize(); gather_ms = (time.perf_counter() - g0) * 1e3
        assert sum(1 for o in allrec if o.n_crop > 0) >= n_obj
    ms, e2e_ms, e2e_pip_ms = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    frames_total = args.frames * world * args.steps
    value = frames_total / (ms / 1e3)
    e2e_sync = frames_total / (e2e_ms / 1e3)
    e2e = frames_total / (e2e_pip_ms / 1e3)
    # ---- roofline of the dominant kernel (algorithmic bytes per SURVEY.md 8(d))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    crop_bytes = float(sum(2 * ((int(b[2]) - int(b[0]) + 2) // 3) * ((int(b[3]) - int(b[1]) + 2) // 3) for b in bb))   # 2 B per stride-3 sample
    n_sym = np.array([o.n_sym for o in out], dtype=np.float64)
    mult = np.array([2.0 if o.symmetry_type == 2 else 1.0 for o in out])
    sym_bytes = float((14.0 * n_sym * mult * 405.0).sum())   # ~14 B x M per cost evaluation x ~405 evaluations per object
    if k_ms[1] >= k_ms[0]:
        dom, alg, t_ms = "symmetry_finalize_kernel", sym_bytes, k_ms[1]
    else:
        dom, alg, t_ms = "extract_kernel", crop_bytes, k_ms[0]
    achieved = alg / (t_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    try:   # DRAM bytes per object of the same kernel from the committed ncu --set full capture, scaled to this launch's objects
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        traffic = float(tj["kernels"][dom]["dram_bytes_per_object"]) * n_obj
        traffic_src = "profiles/r01_ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per object x objects per launch)"
    except Exception:
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "kernel_ms": {"extract_kernel": k_ms[0], "symmetry_finalize_kernel": k_ms[1]},
                "algorithmic_bytes_per_launch": alg,
                "note": "both kernels are latency / instruction-issue bound at the reference's problem sizes (per-object working set << L2); see DESIGN.md"}
    # ---- cpu baseline (1 thread, bounded sample): rank 0 at N = 1 only
    cpu_baseline = None
    if world == 1:
        cpu_fps, cpu_dt = oracle_fps(depth, fo, bb, lb, tw, plane, min(args.cpu_sample, args.frames), 1)
        cpu_baseline = {"value": cpu_fps, "unit": "frames/s", "cores": 1, "kind": "port",
                        "sample": f"{min(args.cpu_sample, args.frames)} frames of the same batch, {cpu_dt:.1f} s, oracle/liboracle.so single thread"}
    line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config, objects_per_step_per_gpu=n_obj, ok_fraction=ok_frac, gather_ms=gather_ms, pipelined_records_equal_blocking=pip_same),
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_per_step) * world, "d2h_bytes_per_step": int(C.sizeof(EslamEllipsoidOut) * n_obj) * world,
                    "api": "eslam_estimate_ellipsoids_submit / _wait, two batches in flight (host pinned buffers in, records out, every step)",
                    "single_blocking_call_value": e2e_sync},
            "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""N3: reader for the TUM-RGB-D-shaped directory the reference demo consumes (src/tum_rgbd/io.cpp:14-94,277-290;
README.md:66-69) and the bbox text format (`id x1 y1 x2 y2 label rate instance` per row, bbox/<rgb timestamp>.txt).
Only what feeds the hot path: depth PNG (uint16), ground-truth pose, detections -- plus readers for the two result files the
reference writes (objects.txt: System::SaveObjectsToFile, src/core/System.cpp:75-91; object_history.txt: Tracking::SaveObjectHistory,
src/core/Tracking.cpp:855-896; the writers live on eslam_b200.tracking.Tracking) and a prefetching depth loader that keeps the
batched entry points fed: PNG decode on background threads straight into the page-locked batch buffers the pull-upload path wants."""
import os
import queue
import threading

import numpy as np


def load_sequence(root, max_frames=None):
    """Returns a list of dicts {timestamp, depth_path, pose7 (x y z qx qy qz qw), bbox (N x 8)} following associateGroundtruth.txt:
    `rgb_ts rgb_path gt_ts x y z qx qy qz qw` and associate.txt `rgb_ts rgb_path depth_ts depth_path`."""
    depth_of = {}
    with open(os.path.join(root, "associate.txt")) as f:
        for ln in f:
            p = ln.split()
            if len(p) >= 4:
                depth_of[p[0]] = p[3]
    frames = []
    with open(os.path.join(root, "associateGroundtruth.txt")) as f:
        for ln in f:
            p = ln.split()
            if len(p) < 10 or p[0] not in depth_of:
                continue
            ts_name = os.path.splitext(os.path.basename(p[1]))[0]
            bfile = os.path.join(root, "bbox", ts_name + ".txt")
            rows = []
            if os.path.exists(bfile):
                rows = [[float(v) for v in r.split()] for r in open(bfile) if r.strip()]
            bbox = np.array(rows, float).reshape(-1, 8) if rows else np.zeros((0, 8))
            frames.append(dict(timestamp=float(p[0]), depth_path=os.path.join(root, depth_of[p[0]]), pose7=np.array([float(v) for v in p[3:10]]), bbox=bbox))
            if max_frames and len(frames) >= max_frames:
                break
    return frames


def read_depth(path):
    import cv2
    d = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if d is None or d.dtype != np.uint16:
        raise IOError(f"{path}: expected a 16-bit depth PNG")
    return d


def read_objects(path):
    """objects.txt -> {instance: vec10 (x y z qx qy qz qw a b c)}."""
    out = {}
    for ln in open(path):
        p = ln.split()
        if len(p) == 11:
            out[int(round(float(p[0])))] = np.array([float(v) for v in p[1:]])
    return out


def read_object_history(path):
    """object_history.txt -> {instance: array [historyNum, 11] of (observation count, valid, x y z roll pitch yaw a b c)}."""
    tok = open(path).read().split()
    total, pos, out = int(tok[0]), 1, {}
    for _ in range(total):
        inst, n = int(tok[pos]), int(tok[pos + 1]); pos += 2
        out[inst] = np.array([float(v) for v in tok[pos:pos + 11 * n]]).reshape(n, 11); pos += 11 * n
    return out


class DepthPrefetcher:
    """Iterates over a frame list (load_sequence) in batches of `batch` frames: yields (frames, depth[B, rows, cols] uint16) with the depth
    batch already decoded into one of `depth_buffers` reusable buffers -- page-locked when CUDA is present (the *_host / _submit entry points
    then pull just the box rectangles over PCIe), plain numpy otherwise.  `workers` threads decode PNGs ahead of the consumer; a buffer is
    handed back for reuse when the consumer asks for the batch after the next one (two batches may be in flight in submit / wait)."""

    def __init__(self, frames, batch, rows=480, cols=640, workers=4, depth_buffers=3, pinned=None):
        self.frames, self.batch, self.rows, self.cols = list(frames), int(batch), rows, cols
        if pinned is None:
            try:
                import torch
                pinned = torch.cuda.is_available()
            except ImportError:
                pinned = False
        self._bufs, self._keep = [], []
        for _ in range(max(3, depth_buffers)):
            if pinned:
                import torch
                t = torch.empty((self.batch, rows, cols), dtype=torch.int16).pin_memory()
                self._keep.append(t)
                self._bufs.append(t.numpy().view(np.uint16))
            else:
                self._bufs.append(np.empty((self.batch, rows, cols), np.uint16))
        self.pinned = bool(pinned)
        self._free = queue.Queue()
        for i in range(len(self._bufs)):
            self._free.put(i)
        self._ready = queue.Queue(maxsize=len(self._bufs))
        self._workers = max(1, int(workers))
        self._thread = threading.Thread(target=self._produce, daemon=True)
        self._thread.start()

    def _decode_into(self, dst, frame):
        d = read_depth(frame["depth_path"])
        if d.shape != (self.rows, self.cols):
            raise IOError(f"{frame['depth_path']}: {d.shape} instead of {(self.rows, self.cols)}")
        dst[...] = d

    def _produce(self):
        from concurrent.futures import ThreadPoolExecutor
        try:
            with ThreadPoolExecutor(self._workers) as pool:
                for b0 in range(0, len(self.frames), self.batch):
                    chunk = self.frames[b0:b0 + self.batch]
                    i = self._free.get()
                    buf = self._bufs[i]
                    list(pool.map(lambda kf: self._decode_into(buf[kf[0]], kf[1]), enumerate(chunk)))   # cv2.imread releases the GIL
                    self._ready.put((i, chunk, None))
            self._ready.put((None, None, None))
        except Exception as e:   # surfaces in the consumer
            self._ready.put((None, None, e))

    def __iter__(self):
        held = []
        while True:
            i, chunk, err = self._ready.get()
            if err is not None:
                raise err
            if i is None:
                break
            held.append(i)
            if len(held) > 2:   # the batch before the previous one is no longer in flight
                self._free.put(held.pop(0))
            yield chunk, self._bufs[i][:len(chunk)]
        for i in held:
            self._free.put(i)

"""Frame-level data parallelism for the hot path: per-frame ellipsoid extraction depends only on
(depth, bboxes, Twc, supporting plane, params), so frames shard across ranks with NO data-path collective; the only exchange is
one all-gather of the fixed-size per-(frame,bbox) result records ("gather of map", BASELINE.json configs[4]).
Works with any torch.distributed backend: nccl on the GPU box (records on the device), gloo in the CPU tests."""
import ctypes as C

import numpy as np

from ._structs import EslamEllipsoidOut

RECORD_BYTES = C.sizeof(EslamEllipsoidOut)


def shard_frames(n_frames, rank, world):
    """rank r owns frames {i : i mod world == r} (round robin keeps the ranks balanced on a live stream)."""
    return np.arange(rank, n_frames, world, dtype=np.int64)


def shard_objects(frame_of_obj, my_frames):
    """indices of the objects living on my frames + their frame index re-based into my local frame list."""
    frame_of_obj = np.asarray(frame_of_obj)
    local = {int(f): k for k, f in enumerate(my_frames)}
    sel = np.array([i for i, f in enumerate(frame_of_obj) if int(f) in local], dtype=np.int64)
    return sel, np.array([local[int(frame_of_obj[i])] for i in sel], dtype=np.int32)


def records_to_bytes(out_array):
    return np.frombuffer(bytes(out_array), dtype=np.uint8).copy()


def bytes_to_records(buf, n):
    arr = (EslamEllipsoidOut * n)()
    C.memmove(arr, np.ascontiguousarray(buf).ctypes.data, n * RECORD_BYTES)
    return arr


def gather_records(local_records, local_obj_ids, n_obj_total, device="cpu", group=None):
    """All-gathers the per-object records of every rank and returns them in global object order (on every rank).
    Ranks may own different object counts: records are padded to the maximum count (fixed-size exchange)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_local = len(local_obj_ids)
    cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    n_max = int(max(int(c.item()) for c in cnts))
    rec = torch.zeros(n_max * RECORD_BYTES, dtype=torch.uint8, device=device)
    ids = torch.full((n_max,), -1, dtype=torch.int64, device=device)
    if n_local:
        rec[: n_local * RECORD_BYTES] = torch.from_numpy(records_to_bytes(local_records)).to(device)
        ids[:n_local] = torch.as_tensor(np.asarray(local_obj_ids, dtype=np.int64), device=device)
    all_rec = [torch.zeros_like(rec) for _ in range(world)]
    all_ids = [torch.zeros_like(ids) for _ in range(world)]
    dist.all_gather(all_rec, rec, group=group)
    dist.all_gather(all_ids, ids, group=group)
    out = (EslamEllipsoidOut * n_obj_total)()
    for r in range(world):
        buf = all_rec[r].cpu().numpy(); gid = all_ids[r].cpu().numpy()
        for k in range(int(cnts[r].item())):
            C.memmove(C.byref(out, int(gid[k]) * RECORD_BYTES), buf[k * RECORD_BYTES:(k + 1) * RECORD_BYTES].ctypes.data, RECORD_BYTES)
    return out


def gather_records_dev(handle, n_obj, stream_ptr=0, group=None, out=None):
    """Device-side "gather of map": the records of the handle's last *_dev batch go D2D into a torch buffer on the caller's stream and
    are all-gathered over NCCL (all_gather_into_tensor) on the same stream -- no host round trip, no per-record Python.  Every rank
    must own the same number of objects.  Returns a uint8 CUDA tensor [world, n_obj * RECORD_BYTES] in rank order; records_view()
    turns a rank's row (after .cpu()) into an EslamEllipsoidOut array."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.empty(n_obj * RECORD_BYTES, dtype=torch.uint8, device="cuda")
    if out is None:
        out = torch.empty(world * n_obj * RECORD_BYTES, dtype=torch.uint8, device="cuda")
    handle.copy_ellipsoids_dev(n_obj, mine.data_ptr(), stream=stream_ptr)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.view(world, n_obj * RECORD_BYTES)


def records_view(row_u8_cpu, n):
    return bytes_to_records(row_u8_cpu.numpy() if hasattr(row_u8_cpu, "numpy") else row_u8_cpu, n)

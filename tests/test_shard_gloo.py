"""world_size-2 gloo test of the multi-rank host logic (frame sharding + the one all-gather of result records).
The per-rank compute is done by the CPU oracle here (no GPU in this container); on the GPU box bench.py runs the same
shard/gather code over NCCL with the CUDA path."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "object-oriented-slam_b200")]
    import torch.distributed as dist
    import orc
    from eslam_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    depth, fo, bb, lb, tw = synth.make_batch(3, "multi_box", n_boxes=3)
    mine = shard.shard_frames(len(depth), rank, world)
    sel, local_f = shard.shard_objects(fo, mine)
    out, _, _, _ = orc.estimate_ellipsoids(depth[mine], local_f, bb[sel], lb[sel], tw[mine], synth.GROUND_PLANE_WORLD)
    allrec = shard.gather_records(out, sel, len(fo))
    if rank == 0:
        q.put([(o.status, o.n_crop, o.n_cluster, list(o.ellipsoid)) for o in allrec])
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_extraction_equals_single_process():
    sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "object-oriented-slam_b200")]
    import orc
    from eslam_b200 import shard, synth
    assert list(shard.shard_frames(5, 1, 2)) == [1, 3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    depth, fo, bb, lb, tw = synth.make_batch(3, "multi_box", n_boxes=3)
    ref, _, _, _ = orc.estimate_ellipsoids(depth, fo, bb, lb, tw, synth.GROUND_PLANE_WORLD)
    assert len(got) == len(ref) == 9
    for g, o in zip(got, ref):
        assert g[0] == o.status and g[1] == o.n_crop and g[2] == o.n_cluster
        assert np.array_equal(np.array(g[3]), np.array(o.ellipsoid))

"""N4: data association -- the host mirror of DataAssociationSolver (src/core/DataAssociation.cpp:16-135) against a
literal restatement of SolveCostMat, and the batched device version (csrc/associate.cu) against the host mirror."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "object-oriented-slam_b200")]
from eslam_b200.tracking import DataAssociationSolver, se3_mul  # noqa: E402


def _frames(rng, n_frames, n_map, max_det=8):
    centres = rng.uniform(-3, 3, size=(n_map, 3))
    m = {10 * k + 3: dict(vec=np.r_[centres[k], 0, 0, 0, 1, 0.3, 0.3, 0.3]) for k in range(n_map)}   # instance ids are not 0..n-1: the result is a COLUMN index
    frames = []
    for _ in range(n_frames):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        Twc = np.r_[rng.uniform(-1, 1, size=3), q if q[3] >= 0 else -q]
        n = int(rng.integers(0, max_det + 1))
        objs = []
        for _ in range(n):
            u = rng.uniform()
            if u < 0.2:
                objs.append(None)
                continue
            # an observation near a map object (in the camera frame), or somewhere else
            if n_map and u < 0.8:
                target = centres[int(rng.integers(0, n_map))] + rng.normal(scale=0.3, size=3)
            else:
                target = rng.uniform(-6, 6, size=3)
            Tcw = _inv(Twc)
            local = se3_mul(Tcw, np.r_[target, 0, 0, 0, 1])
            objs.append(dict(vec=np.r_[local, 0.2, 0.2, 0.2]))
        frames.append(SimpleNamespace(mmObservations=np.zeros((n, 8)), mpLocalObjects=objs, cam_pose_Twc=Twc))
    return frames, m


def _inv(T):
    from eslam_b200.tracking import pose_inverse_np
    return pose_inverse_np(T)


def _solve_cost_mat_literal(cost, counter):
    """SolveCostMat :90-135, statement by statement."""
    cost = cost.copy(); rows = cost.shape[0]; out = [0] * rows
    for i in range(rows):
        vec = cost[i]
        if vec.shape[0] < 1:
            out[i] = counter[0]; counter[0] += 1
            continue
        j = int(np.argmin(vec))
        if vec[j] < 1.0:
            out[i] = j
            cost[:, j] = 999
        else:
            out[i] = counter[0]; counter[0] += 1
    return out


def test_host_mirror_follows_solve_cost_mat():
    rng = np.random.default_rng(5)
    for n_map in (0, 1, 7):
        frames, m = _frames(rng, 20, n_map)
        solver = DataAssociationSolver(); counter = [0]
        for f in frames:
            got = solver.Solve(f, m)
            valid = [i for i, o in enumerate(f.mpLocalObjects) if o is not None]
            insts = sorted(m)
            cost = np.array([[np.linalg.norm(m[k]["vec"][:3] - se3_mul(f.cam_pose_Twc, f.mpLocalObjects[i]["vec"][:7])[:3]) for k in insts] for i in valid]).reshape(len(valid), len(insts))
            want = [-1] * len(f.mpLocalObjects)
            for i, v in zip(valid, _solve_cost_mat_literal(cost, counter) if valid else []):
                want[i] = v
            assert got == want
        assert solver.miInstanceNum == counter[0]


def test_solve_batch_without_a_device_equals_the_sequential_calls():
    rng = np.random.default_rng(6)
    frames, m = _frames(rng, 12, 5)
    a, b = DataAssociationSolver(), DataAssociationSolver()
    assert a.SolveBatch(frames, m) == [b.Solve(f, m) for f in frames] and a.miInstanceNum == b.miInstanceNum


@pytest.mark.gpu
@pytest.mark.parametrize("n_map", [0, 1, 40])
def test_device_batch_equals_the_host_mirror(n_map):
    import eslam_b200 as E
    rng = np.random.default_rng(7 + n_map)
    frames, m = _frames(rng, 300, n_map)
    h = E.Handle()
    dev, host = DataAssociationSolver(), DataAssociationSolver()
    dev.miInstanceNum = host.miInstanceNum = 5
    l0 = h.launches
    got = dev.SolveBatch(frames, m, backend=h)
    assert h.launches - l0 == 2
    want = [host.Solve(f, m) for f in frames]
    h.close()
    assert got == want and dev.miInstanceNum == host.miInstanceNum
    assert any(v >= 5 for row in got for v in row) and (n_map == 0 or any(0 <= v < n_map for row in got for v in row))

"""Host-side orchestration (N1) replayed on the golden fr3_cabinet frames with the oracle backend (CPU) and, on the GPU box, with the
CUDA backend -- both must produce the same map; the objects.txt writer is checked for format."""
import os

import numpy as np
import pytest

import orc
from eslam_b200 import Optimizer
from eslam_b200.tracking import Tracking, pose_inverse_np, se3_mul
from helpers import K, ellipsoid_close
from oracle_backend import OracleBackend

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cabinet_golden.npz"))


def _replay(backend, n_iter=10):
    tr = Tracking(backend, Optimizer, K)
    for f in range(len(G["depth"])):
        bbox = np.array([[0, *G["bbox"][f], G["label"][f], 0.5, 0]], float)   # id x1 y1 x2 y2 label rate instance
        tr.GrabPoseAndObjects(float(G["timestamps"][f]), G["pose"][f], bbox, G["depth"][f], withAssociation=True, iters=n_iter)
    return tr


def test_se3_helpers_match_oracle():
    a = orc.se3_exp([0.3, -0.2, 0.5, 1.0, 2.0, 3.0]); b = orc.se3_exp([-0.1, 0.4, 0.2, -1.0, 0.5, 0.3])
    assert np.allclose(se3_mul(a, b), orc.se3_mul(a, b), atol=1e-14)
    assert np.allclose(pose_inverse_np(a), orc.se3_inverse(a), atol=1e-14)


def test_tracking_replay_with_oracle_backend(tmp_path):
    tr = _replay(OracleBackend())
    assert tr.miGroundPlaneState == 2 and np.allclose(tr.mGroundPlane, G["plane_world"], atol=1e-12)
    assert list(tr.map_ellipsoids) == [0]
    assert len(tr.mmObjectObservations[0]) == 4 and len(tr.mmObjectObservations3D[0]) == 4
    e = tr.map_ellipsoids[0]["vec"]
    # the optimised object stays close to the single-frame estimates moved to the world frame
    first = np.array(G["ellipsoid"][0]); first[:7] = se3_mul(G["pose"][0], first[:7])
    assert np.linalg.norm(e[:3] - first[:3]) < 0.25 and np.all(e[7:] > 0.1) and np.all(e[7:] < 0.8)
    assert tr.optimizer.last_stats.chi2_final <= tr.optimizer.last_stats.chi2_initial
    p = tmp_path / "objects.txt"
    tr.SaveObjectsToFile(str(p))
    rows = [ln.split() for ln in open(p)]
    assert len(rows) == 1 and len(rows[0]) == 11 and float(rows[0][0]) == 0.0
    assert np.allclose([float(v) for v in rows[0][1:]], e, rtol=1e-11)


@pytest.mark.gpu
def test_tracking_replay_gpu_matches_oracle_backend():
    import eslam_b200 as E
    h = E.Handle()
    tg = _replay(h)
    to = _replay(OracleBackend())
    assert list(tg.map_ellipsoids) == list(to.map_ellipsoids) == [0]
    assert ellipsoid_close(tg.map_ellipsoids[0]["vec"], to.map_ellipsoids[0]["vec"], 1e-4)
    h.close()


REF_SEQ = "/root/reference/Example/dataset/cabinet"


@pytest.mark.skipif(not os.path.isdir(REF_SEQ), reason="reference dataset only exists in the authoring container")
def test_full_cabinet_sequence_smoke():
    """replay of the bundled sequence -> exactly one instance (id 0) in objects.txt (oracle backend)."""
    from eslam_b200 import tum_io
    frames = tum_io.load_sequence(REF_SEQ)
    assert len(frames) == 58 and sum(len(f["bbox"]) > 0 for f in frames) == 51
    tr = Tracking(OracleBackend(), Optimizer, K)
    n_extracted = 0
    for fr in frames[:20]:
        f = tr.GrabPoseAndObjects(fr["timestamp"], fr["pose7"], fr["bbox"], tum_io.read_depth(fr["depth_path"]), withAssociation=True, iters=10)
        n_extracted += sum(o is not None for o in f.mpLocalObjects)
    assert tr.miGroundPlaneState == 2 and list(tr.map_ellipsoids) == [0] and n_extracted >= 15
    assert np.isfinite(tr.map_ellipsoids[0]["vec"]).all()


def test_automatic_data_association_rules():
    from eslam_b200.tracking import DataAssociationSolver

    class F:
        pass
    f = F(); f.cam_pose_Twc = np.array([0, 0, 0, 0, 0, 0, 1.0]); f.mmObservations = np.zeros((3, 8))
    mk = lambda x: dict(vec=np.array([x, 0, 2, 0, 0, 0, 1, .3, .3, .3]), prob=1.0, miLabel=1, miInstanceID=-1)
    f.mpLocalObjects = [mk(0.1), None, mk(5.0)]
    s = DataAssociationSolver()
    assert s.Solve(f, {}) == [0, -1, 1]                                   # empty map: every valid observation creates an instance
    m = {7: mk(0.0), 9: mk(3.0)}
    assert DataAssociationSolver().Solve(f, m) == [0, -1, 0]              # match -> column index 0; 5.0 is > 1 m from both -> new instance 0

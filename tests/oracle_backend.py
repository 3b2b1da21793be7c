"""Oracle-backed stand-in for eslam_b200.api.Handle (same methods) so the host-side Tracking / Optimizer logic can be tested on CPU.
TEST INFRASTRUCTURE ONLY."""
import numpy as np

import orc
from eslam_b200._structs import default_camera, default_params


class OracleBackend:
    def __init__(self):
        self.params, self.camera = default_params(), default_camera()
        self.plane = None

    def set_supporting_plane(self, p):
        self.plane = np.array(p, float)

    def extract_ground_plane(self, depth):
        return orc.extract_ground_plane(depth, self.params, self.camera)

    def estimate_ellipsoids(self, depth, fo, bb, lb, tw, dump_cap=0):
        out, d, k, _ = orc.estimate_ellipsoids(depth, fo, bb, lb, tw, self.plane, self.params, self.camera, dump_cap)
        return out, d, k

    def init_quadric(self, Twc, det5, K, rows=None, cols=None):
        return orc.init_quadric(Twc, det5, K, rows or 480, cols or 640)

    def edges_eval(self, g, ell):
        go = orc.Graph(g.type, g.frame, g.object, g.meas, g.info, g.Tcw, g.n_objects, g.K)
        return orc.edges_eval(go, ell)

    def optimize_objects(self, g, ell, iters=10):
        go = orc.Graph(g.type, g.frame, g.object, g.meas, g.info, g.Tcw, g.n_objects, g.K)
        return orc.optimize_objects(go, ell, iters)

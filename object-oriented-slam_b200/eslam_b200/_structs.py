"""ctypes mirrors of the POD structs in include/eslam.h (the C-ABI of the hot path).

The layout here IS the ABI; tests/test_abi.py checks sizeof() against the library's own
answer (eslam_sizeof_*) so the two cannot drift silently.
"""
import ctypes as C

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class EslamCamera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("scale", C.c_double), ("rows", C.c_int), ("cols", C.c_int)]


class EslamParams(C.Structure):
    _fields_ = [("depth_range", C.c_double), ("voxel_leaf", C.c_double), ("plane_margin", C.c_double),
                ("cluster_tolerance", C.c_double), ("min_cluster_size", C.c_int), ("center_dis", C.c_double),
                ("symmetry_open", C.c_int), ("symmetry_grid", C.c_double), ("symmetry_sigma", C.c_double),
                ("symmetry_iters", C.c_int), ("plane_min_size", C.c_int), ("plane_angle_deg", C.c_double),
                ("plane_dist", C.c_double), ("edge3d_scale", C.c_double), ("gravity_prior_open", C.c_int),
                ("gravity_prior_scale", C.c_double), ("border_pixels", C.c_int), ("length_limit_pixels", C.c_int)]


class EslamEllipsoidOut(C.Structure):
    _fields_ = [("ellipsoid", C.c_double * 10), ("prob", C.c_double), ("status", C.c_int),
                ("symmetry_type", C.c_int), ("symmetry_result", C.c_int), ("best_seed", C.c_int),
                ("sym_plane_world", C.c_double * 4), ("sym_plane2_world", C.c_double * 4),
                ("sym_center_world", C.c_double * 3),
                ("n_crop", C.c_int), ("n_voxel", C.c_int), ("n_plane", C.c_int), ("n_clusters", C.c_int),
                ("n_cluster", C.c_int), ("n_sym", C.c_int), ("n_final", C.c_int), ("reserved", C.c_int)]


class EslamStageDump(C.Structure):
    _fields_ = [("cap_points", C.c_int), ("crop_xyz", c_double_p), ("voxel_xyz", c_double_p),
                ("world_xyz", c_double_p), ("cluster_label", c_int_p), ("sym_xyz", c_double_p),
                ("final_xyz", c_double_p), ("center_world", C.c_double * 3), ("pca_center", C.c_double * 3),
                ("pca_rot", C.c_double * 9), ("pca_eval", C.c_double * 3), ("T_wo", C.c_double * 7),
                ("seed_prob", C.c_double * 9), ("seed_init_err", C.c_double * 9),
                ("seed_final_err", C.c_double * 9), ("seed_plane", (C.c_double * 4) * 9),
                ("seed_dual_dis", C.c_double * 9)]


class EslamEdgeBatch(C.Structure):
    _fields_ = [("n_edges", C.c_int), ("type", c_int_p), ("frame", c_int_p), ("object", c_int_p),
                ("meas", c_double_p), ("info", c_double_p), ("n_frames", C.c_int), ("Tcw", c_double_p),
                ("n_objects", C.c_int), ("K", C.c_double * 9)]


class EslamLmStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("trials_total", C.c_int), ("terminated", C.c_int),
                ("chi2_initial", C.c_double), ("chi2_final", C.c_double), ("lambda_final", C.c_double)]


EDGE_PROJ, EDGE_9DOF, EDGE_GRAVITY = 0, 1, 2


def default_params():
    """Example/param/TUM3.yaml values + the hard-coded constants."""
    p = EslamParams()
    p.depth_range = 6.0
    p.voxel_leaf = 0.01
    p.plane_margin = 0.05
    p.cluster_tolerance = 0.02
    p.min_cluster_size = 100
    p.center_dis = 0.5
    p.symmetry_open = 1
    p.symmetry_grid = 0.1
    p.symmetry_sigma = 0.1
    p.symmetry_iters = 5
    p.plane_min_size = 200
    p.plane_angle_deg = 5.0
    p.plane_dist = 0.1
    p.edge3d_scale = 10000.0
    p.gravity_prior_open = 1
    p.gravity_prior_scale = 100.0
    p.border_pixels = 10
    p.length_limit_pixels = 0
    return p


def default_camera():
    """Example/param/TUM3.yaml:62-70."""
    c = EslamCamera()
    c.fx, c.fy, c.cx, c.cy = 535.4, 539.2, 320.1, 247.6
    c.scale = 5000.0
    c.rows, c.cols = 480, 640
    return c

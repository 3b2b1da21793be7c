"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/eslam.h declares, agrees with the
ctypes struct layouts, and refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import eslam_b200 as E
from eslam_b200 import _structs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "eslam.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(eslam_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = E.load_library()
    names = header_functions()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(E.api.ABI_SYMBOLS) <= set(names)
    assert lib.eslam_abi_version() == 1


def test_struct_layouts_match_the_library():
    lib = E.load_library()
    assert lib.eslam_sizeof_params() == C.sizeof(S.EslamParams)
    assert lib.eslam_sizeof_camera() == C.sizeof(S.EslamCamera)
    assert lib.eslam_sizeof_ellipsoid_out() == C.sizeof(S.EslamEllipsoidOut)
    assert lib.eslam_sizeof_stage_dump() == C.sizeof(S.EslamStageDump)
    assert lib.eslam_sizeof_edge_batch() == C.sizeof(S.EslamEdgeBatch)
    assert lib.eslam_sizeof_lm_stats() == C.sizeof(S.EslamLmStats)


def test_defaults_match_tum3_yaml():
    lib = E.load_library()
    p, c = S.EslamParams(), S.EslamCamera()
    lib.eslam_default_params(C.byref(p)); lib.eslam_default_camera(C.byref(c))
    d, dc = S.default_params(), S.default_camera()
    for f, _ in S.EslamParams._fields_:
        assert getattr(p, f) == getattr(d, f), f
    for f, _ in S.EslamCamera._fields_:
        assert getattr(c, f) == getattr(dc, f), f
    assert (p.cluster_tolerance, p.min_cluster_size, p.symmetry_grid, p.plane_min_size, p.edge3d_scale) == (0.02, 100, 0.1, 200, 10000.0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EslamError):
        E.Handle()


def test_host_helpers_match_oracle():
    import orc
    T = orc.se3_exp([0.3, -0.2, 0.5, 1.0, 2.0, 3.0])
    assert np.allclose(E.pose_inverse(T), orc.se3_inverse(T), atol=1e-15)
    Tn = T.copy(); Tn[3:] *= 1.0003          # non-unit quaternion (TUM ground truth has 4 decimals)
    pl = np.array([0.01, -0.74, -0.68, 1.09])
    assert np.allclose(E.plane_to_world(pl, Tn), orc.plane_to_world(pl, Tn), atol=1e-13)
    for bb in ([50, 50, 300, 300], [5, 50, 300, 300], [50, 50, 630, 300], [50, 50, 300, 470]):
        assert E.calibrate_measurement(bb) == bool(orc.calibrate_measurement(bb))


def test_entry_points_reject_a_null_handle_without_crashing():
    """every handle-taking entry point returns ESLAM_ERR_BAD_ARG (no dereference) for a NULL handle -- checked without a GPU."""
    lib = E.load_library()
    null = C.c_void_p(None)
    z = C.c_int(0)
    buf = (C.c_double * 16)()
    calls = [
        lambda: lib.eslam_set_supporting_plane(null, buf),
        lambda: lib.eslam_extract_ground_plane_host(null, 1, None, C.c_size_t(640), buf, C.byref(z), C.byref(z)),
        lambda: lib.eslam_estimate_ellipsoids_host(null, 1, None, C.c_size_t(640), 1, C.byref(z), buf, C.byref(z), buf, None, None),
        lambda: lib.eslam_estimate_ellipsoids_dev(null, None, 1, None, C.c_size_t(640), 1, C.byref(z), buf, C.byref(z), buf),
        lambda: lib.eslam_estimate_ellipsoids_submit(null, 1, None, C.c_size_t(640), 1, C.byref(z), buf, C.byref(z), buf, C.byref(z)),
        lambda: lib.eslam_estimate_ellipsoids_wait(null, 0, 1, None),
        lambda: lib.eslam_fetch_ellipsoids(null, None, 1, None),
        lambda: lib.eslam_init_quadric(null, 1, buf, buf, buf, 480, 640, buf, C.byref(z)),
        lambda: lib.eslam_edges_eval_dev(null, None),
        lambda: lib.eslam_set_profiling(null, 1),
    ]
    for f in calls:
        assert f() != 0
    lib.eslam_launch_count.restype = C.c_longlong
    assert lib.eslam_launch_count(null) == 0 and lib.eslam_h2d_bytes(null) == 0
    lib.eslam_destroy(null)       # no-op

"""Stage-by-stage GPU-vs-oracle diagnostics (run on the GPU box; prints, never asserts)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np  # noqa: E402

import orc  # noqa: E402
import eslam_b200 as E  # noqa: E402
from eslam_b200 import synth  # noqa: E402
from eslam_b200.api import EdgeGraph  # noqa: E402
from helpers import K, make_ba_problem  # noqa: E402
from test_gpu_parity import _compare_objects  # noqa: E402

h = E.Handle()
print("== edges")
p = make_ba_problem(40, 10, 10)
go = orc.Graph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], K)
gg = EdgeGraph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], K)
r0, J0, c0 = orc.edges_eval(go, p["init"])
r1, J1, c1 = h.edges_eval(gg, p["init"])
for t in range(3):
    m = np.array(p["type"]) == t
    print(f" type {t}: n={m.sum()} max|dr|={np.abs(r1[m]-r0[m]).max():.3e} (|r|max {np.abs(r0[m]).max():.3e})  max|dJ|={np.abs(J1[m]-J0[m]).max():.3e} (|J|max {np.abs(J0[m]).max():.3e})")
e0, s0 = orc.optimize_objects(go, p["init"], 10)
e1, s1 = h.optimize_objects(gg, p["init"], 10)
print(" LM oracle:", s0.iterations, s0.trials_total, s0.terminated, s0.chi2_initial, s0.chi2_final, s0.lambda_final)
print(" LM gpu   :", s1.iterations, s1.trials_total, s1.terminated, s1.chi2_initial, s1.chi2_final, s1.lambda_final)
print(" LM max|d ellipsoid| =", np.abs(e1 - e0).max())
print("== extraction")
for kind, nf in (("one_box", 6), ("multi_box", 3)):
    depth, fo, bb, lb, tw = synth.make_batch(nf, kind)
    h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
    t = time.time(); og, dg, kg = h.estimate_ellipsoids(depth, fo, bb, lb, tw, dump_cap=40000); tg = time.time() - t
    t = time.time(); oo, do, ko, ev = orc.estimate_ellipsoids(depth, fo, bb, lb, tw, synth.GROUND_PLANE_WORLD, dump_cap=40000); to = time.time() - t
    print(f" {kind}: {len(fo)} objects  gpu(with dump) {tg*1e3:.1f} ms  oracle {to*1e3:.1f} ms")
    for i, (g, o) in enumerate(zip(og, oo)):
        print(f"  obj {i}: status {g.status}/{o.status} counts gpu {g.n_crop, g.n_voxel, g.n_plane, g.n_clusters, g.n_cluster, g.n_sym, g.n_final} oracle {o.n_crop, o.n_voxel, o.n_plane, o.n_clusters, o.n_cluster, o.n_sym, o.n_final} seed {g.best_seed}/{o.best_seed} prob {g.prob:.9f}/{o.prob:.9f}")
        print(f"        max|d ellipsoid| = {np.abs(np.array(g.ellipsoid) - np.array(o.ellipsoid)).max():.3e}")
    for m in _compare_objects(og, dg, kg, oo, do, ko):
        print("  MISMATCH", m)
print("== ground plane")
depth, fo, bb, lb, tw = synth.make_batch(3, "one_box")
for f in range(3):
    normals, dm, labels, coef, size, first = orc.extract_planes_debug(depth[f])
    g_dm = h.debug_plane_stage(depth[f], 1)
    m = (dm <= 12) | (g_dm <= 12)
    print(f" frame {f}: dm mismatches (<=12 region): {(g_dm[m] != dm[m]).sum()} of {m.sum()}   max diff {np.abs(g_dm[m]-dm[m]).max() if m.any() else 0}")
    g_n = h.debug_plane_stage(depth[f], 2)
    nan_mis = (np.isnan(g_n[..., 0]) != np.isnan(normals[..., 0])).sum()
    ok = ~np.isnan(normals[..., 0]) & ~np.isnan(g_n[..., 0])
    print(f"   normals: nan-mask mismatches {nan_mis}; value mismatches {(g_n[..., :3][ok] != normals[ok]).any(axis=-1).sum()} of {ok.sum()}; max diff {np.abs(g_n[..., :3][ok]-normals[ok]).max():.3e}")
    from test_gpu_parity import _same_partition
    g_l3 = h.debug_plane_stage(depth[f], 3); g_l4 = h.debug_plane_stage(depth[f], 4)
    print(f"   labels after refine: same partition {_same_partition(g_l4, labels)}  (pre-refine vs refined differ in {(g_l3 != g_l4).sum()} px)")
t = time.time(); pg, fg, ng = h.extract_ground_plane(depth); tg = time.time() - t
t = time.time(); po, fo_, no = orc.extract_ground_plane(depth); to = time.time() - t
print(f" planes gpu {tg*1e3:.1f} ms oracle {to*1e3:.1f} ms"); print(pg, fg, ng); print(po, fo_, no); print(" max|d plane|", np.abs(pg - po).max())
print("launches:", h.launches)

"""Per-phase cycle counts of extract_kernel (clock64 stamps of thread 0, averaged over the objects of one resident batch)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "object-oriented-slam_b200")]
import numpy as np
import torch
import eslam_b200 as E
from eslam_b200 import synth
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 444
depth, fo, bb, lb, tw = bench.make_frames("one_box", n, 1000)
h = E.Handle(); h.set_supporting_plane(synth.GROUND_PLANE_WORLD)
d = torch.from_numpy(depth.view(np.int16)).cuda()
h.set_profiling(True)
for _ in range(3):
    h.estimate_ellipsoids_dev(d.data_ptr(), n, depth.shape[2], fo, bb, lb, tw); torch.cuda.synchronize()
clk = np.zeros((n, 9), np.int64); sclk = np.zeros((n, 4), np.int64)
for i in range(n):
    out = (C.c_longlong * 13)()
    assert h.lib.eslam_debug_phase_clocks(h._h, C.c_int(i), out) == 0
    clk[i] = list(out)[:9]; sclk[i] = list(out)[9:]
dt = np.diff(clk, axis=1).astype(float)
names = ["crop", "voxel sort", "voxel centroid + plane filter", "centre + cluster keys + sort", "neighbour search", "flatten + select + gather", "PCA + frame", "symmetry grid"]
tot = dt.sum(axis=1).mean()
print("stage ms", h.stage_ms(), " mean cycles per CTA %.0f (%.3f ms at 1.965 GHz)" % (tot, tot / 1.965e6))
for k, nm in enumerate(names):
    print("  %-32s %9.0f cycles  %5.1f %%" % (nm, dt[:, k].mean(), 100 * dt[:, k].mean() / tot))
# symmetry: the slots mean different things per implementation (ESLAM_SYM_IMPL)
impl = os.environ.get("ESLAM_SYM_IMPL", "3")
if impl == "2":     # single launch: cycles of the CTA's phases (thread 0)
    tot = sclk[:, 3].mean()
    print("symmetry_finalize_kernel (single launch): mean cycles per CTA %.0f (%.3f ms)" % (tot, tot / 1.965e6))
    for k, nm in enumerate(["scalar phases (LM step of the nine seeds, one warp)", "task phases (FD passes + classification)", "scan phases + ordered sums"]):
        print("  %-52s %9.0f cycles  %5.1f %%" % (nm, sclk[:, k].mean(), 100 * sclk[:, k].mean() / tot))
elif impl == "1":   # round-1 kernel: per-warp sums
    st = sclk.sum(axis=1).mean()
    print("round-1 symmetry kernel: mean cycles per seed warp %.0f (%.3f ms)" % (st / 9, st / 9 / 1.965e6))
    for k, nm in enumerate(["classification passes (mirror + project + caches)", "nearest-neighbour scans", "ordered ln P sums", "rest (LM scalar code, oplus, plane pairs, set-up)"]):
        print("  %-52s %9.0f cycles  %5.1f %%" % (nm, sclk[:, k].mean() / 9, 100 * sclk[:, k].mean() / st))
else:               # round pipeline: finite-difference passes and how many fell back to full evaluations
    print("symmetry round pipeline: finite-difference passes per object %.1f, fallbacks per object %.3f (per-round times: tools/ncu_rounds.py over an ncu launch list)" % (sclk[:, 0].mean(), sclk[:, 1].mean()))

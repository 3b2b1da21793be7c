// End-to-end run of the C++ host adapter on the device (tests/test_gpu_parity.py::test_cpp_adapter_end_to_end): ground plane -> local
// ellipsoid per frame -> quadric initialisation -> object graph optimisation, printed as numbers the test diffs with the ctypes path.
// input file: int32 n_frames, rows, cols; uint16 depth[n][rows][cols]; double bbox[n][4]; int32 label[n]; double Twc[n][7]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>
#include "../object-oriented-slam_b200/host/eslam_adapter.hpp"
using namespace EllipsoidSLAM;
int main(int argc, char **argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  int hdr[3]; in.read((char *)hdr, sizeof(hdr));
  const int n = hdr[0], rows = hdr[1], cols = hdr[2];
  std::vector<uint16_t> depth((size_t)n * rows * cols); in.read((char *)depth.data(), depth.size() * 2);
  std::vector<double> bbox((size_t)n * 4); in.read((char *)bbox.data(), bbox.size() * 8);
  std::vector<int> label(n); in.read((char *)label.data(), n * 4);
  std::vector<double> Twc((size_t)n * 7); in.read((char *)Twc.data(), Twc.size() * 8);
  if (!in) return 3;
  EslamParams p; EslamCamera c; eslam_default_params(&p); eslam_default_camera(&c);
  Context ctx(p, c, 0);
  PlaneExtractor pe(ctx); EllipsoidExtractor ex(ctx); Initializer ini(ctx, rows, cols); Optimizer op(ctx);
  Plane4 plane_cam{}, plane_world{};
  const bool found = pe.extractGroundPlane(depth.data(), cols, plane_cam);
  eslam_plane_to_world(plane_cam.data(), &Twc[0], plane_world.data());
  std::printf("plane %d %.17g %.17g %.17g %.17g\n", found ? 1 : 0, plane_world[0], plane_world[1], plane_world[2], plane_world[3]);
  ex.SetSupportingPlane(plane_world); op.SetGroundPlane(plane_world);
  std::vector<Frame *> frames; std::vector<ellipsoid_pod> locals(n); std::vector<Observation> obs_store(n); Observations obs;
  for (int f = 0; f < n; f++) {
    std::vector<std::array<double, 8>> bm{{0, bbox[4 * f], bbox[4 * f + 1], bbox[4 * f + 2], bbox[4 * f + 3], (double)label[f], 1.0, 0}};
    Frame *fr = new Frame(0.1 * f, &Twc[7 * f], bm, depth.data() + (size_t)f * rows * cols, rows, cols);
    Pose7 T; for (int i = 0; i < 7; i++) T[i] = Twc[7 * f + i];
    locals[f] = ex.EstimateLocalEllipsoid(fr->frame_img.data(), cols, {bbox[4 * f], bbox[4 * f + 1], bbox[4 * f + 2], bbox[4 * f + 3]}, label[f], T);
    locals[f].miInstanceID = 0;
    std::printf("local %d %d", f, ex.miSystemState);
    for (double v : locals[f].vec) std::printf(" %.17g", v);
    std::printf(" %.17g\n", locals[f].prob);
    fr->mpLocalObjects.push_back(ex.GetResult() ? &locals[f] : nullptr);
    frames.push_back(fr);
    obs_store[f] = Observation{label[f], {bbox[4 * f], bbox[4 * f + 1], bbox[4 * f + 2], bbox[4 * f + 3]}, 1.0, fr};
    obs.push_back(&obs_store[f]);
  }
  const double K[9] = {c.fx, 0, c.cx, 0, c.fy, c.cy, 0, 0, 1};
  ellipsoid_pod init = ini.initializeQuadric(obs, K);
  std::printf("init %d", ini.getInitializeResult() ? 1 : 0);
  for (double v : init.vec) std::printf(" %.17g", v);
  std::printf("\n");
  // global ellipsoid of instance 0: the first frame's local ellipsoid moved to the world (ellipsoid::transform_from)
  ellipsoid_pod glob = locals[0];
  { double out[7]; eslam_pose_mul(&Twc[0], glob.vec.data(), out); for (int i = 0; i < 7; i++) glob.vec[i] = out[i]; }
  std::map<int, ellipsoid_pod *> emap{{0, &glob}}; std::map<int, Observations> oo{{0, obs}};
  EslamLmStats st = op.GlobalObjectGraphOptimization(frames, emap, K, oo, 10);
  std::printf("opt %d %d", st.iterations, st.trials_total);
  for (double v : glob.vec) std::printf(" %.17g", v);
  std::printf(" %.17g\n", st.chi2_final);
  return 0;
}

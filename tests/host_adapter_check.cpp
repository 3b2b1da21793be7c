// Compile/link check of the C++ host adapter against libeslam_b200.so (run by tests/test_host_adapter.py; no GPU needed:
// it only exercises the ABI functions that do not touch the device and checks that Context refuses to run without CUDA).
#include <cmath>
#include <cstdio>
#include "../object-oriented-slam_b200/host/eslam_adapter.hpp"
int main() {
  EslamParams p; EslamCamera c;
  eslam_default_params(&p); eslam_default_camera(&c);
  if (eslam_abi_version() != ESLAM_ABI_VERSION) return 2;
  double T[7] = {1, 2, 3, 0, 0, std::sin(0.3), std::cos(0.3)}, Ti[7];
  eslam_pose_inverse(T, Ti);
  double bb[4] = {50, 50, 300, 300};
  if (eslam_calibrate_measurement(&p, &c, bb) != 0) return 3;
  bool threw = false;
  try { EllipsoidSLAM::Context ctx(p, c, 0); EllipsoidSLAM::EllipsoidExtractor ex(ctx); EllipsoidSLAM::PlaneExtractor pe(ctx); EllipsoidSLAM::Optimizer op(ctx); EllipsoidSLAM::Initializer in(ctx, 480, 640); EllipsoidSLAM::DataAssociationSolver da(ctx); (void)ex; (void)pe; (void)op; (void)in; (void)da; }
  catch (const std::exception &e) { threw = true; std::printf("no device: %s\n", e.what()); }
  std::printf("adapter ok (device %s)\n", threw ? "absent" : "present");
  return 0;
}

// Compile / instantiate check of the reference-signature overloads of host/eslam_adapter.hpp (ESLAM_HAVE_EIGEN_OPENCV) against mock
// Eigen / OpenCV headers and mock versions of the reference's own types (g2o::ellipsoid, g2o::plane, Frame, Map, Observation).
// Run by tests/test_host_adapter.py; without a GPU the Context constructor throws and the calls are only compiled, with one they run.
#include <cstdio>
#include <map>
#include "../object-oriented-slam_b200/host/eslam_adapter.hpp"
#ifndef ESLAM_HAVE_EIGEN_OPENCV
#error "the mock Eigen/OpenCV headers were not found: pass -Itests/mock_eigen_opencv"
#endif
namespace ref {   // shapes of the reference's types as the adapter templates use them
struct SE3Quat { Eigen::Matrix<double, 7, 1> v; Eigen::Matrix<double, 7, 1> toVector() const { return v; } };
struct ellipsoid {
  Eigen::Matrix<double, 10, 1> v; double prob = 1.0; int miLabel = -1, miInstanceID = -1;
  Eigen::Matrix<double, 10, 1> toVector() const { return v; }
  void fromVector(const Eigen::Matrix<double, 10, 1> &x) { v = x; }
};
struct plane { Eigen::Vector4d param; };
struct Frame { SE3Quat cam_pose_Twc, cam_pose_Tcw; int frame_seq_id = 0; double timestamp = 0; std::vector<ellipsoid *> mpLocalObjects; };
struct Observation { int label; Eigen::Vector4d bbox; double rate; Frame *pFrame; };
typedef std::vector<Observation *> Observations;
struct Map { std::map<int, ellipsoid *> m; std::map<int, ellipsoid *> GetAllEllipsoidsMap() { return m; } };
}  // namespace ref

int main() {
  EslamParams p; EslamCamera c;
  eslam_default_params(&p); eslam_default_camera(&c);
  try {
    EllipsoidSLAM::Context ctx(p, c, 0);
    EllipsoidSLAM::PlaneExtractor pe(ctx); EllipsoidSLAM::EllipsoidExtractor ex(ctx); EllipsoidSLAM::Initializer in(ctx, c.rows, c.cols); EllipsoidSLAM::Optimizer op(ctx);
    cv::Mat depth(c.rows, c.cols);
    for (auto &d : depth.data) d = 7000;
    Eigen::Vector4d plane_param; pe.extractGroundPlane(depth, plane_param);
    ref::plane gp; gp.param[0] = 0; gp.param[1] = 0; gp.param[2] = 1; gp.param[3] = 0;
    ex.SetSupportingPlane(&gp);
    Eigen::Vector4d bbox; bbox[0] = 100; bbox[1] = 100; bbox[2] = 300; bbox[3] = 300;
    Eigen::VectorXd pose(8); for (int i = 0; i < 8; i++) pose[i] = 0; pose[7] = 1;
    EllipsoidSLAM::camera_intrinsic cam{c.fx, c.fy, c.cx, c.cy, c.scale};
    EllipsoidSLAM::ellipsoid_pod e = ex.EstimateLocalEllipsoid(depth, bbox, 28, pose, cam);
    ref::ellipsoid ge = EllipsoidSLAM::to_reference<ref::ellipsoid>(e);
    Eigen::MatrixXd bboxMap(1, 8);
    EllipsoidSLAM::Frame fr(0.0, pose, bboxMap, depth);
    ref::Frame rf; for (int i = 0; i < 7; i++) { rf.cam_pose_Twc.v[i] = pose[1 + i]; rf.cam_pose_Tcw.v[i] = pose[1 + i]; }
    ref::Observation ob{28, bbox, 1.0, &rf}; ref::Observations obs{&ob, &ob, &ob};
    Eigen::Matrix3d K; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) K(i, j) = 0; K(0, 0) = c.fx; K(1, 1) = c.fy; K(0, 2) = c.cx; K(1, 2) = c.cy; K(2, 2) = 1;
    in.initializeQuadric(obs, K);
    Eigen::MatrixXd pm(3, 7), dm(3, 5); in.initializeQuadric(pm, dm, K);
    ref::Map map; map.m[0] = &ge; ge.miInstanceID = 0;
    std::vector<ref::Frame *> frames{&rf}; std::map<int, ref::Observations> oo; oo[0] = obs;
    Eigen::Vector4d gn = gp.param; op.SetGroundPlane(gn);
    op.GlobalObjectGraphOptimization(frames, &map, c.rows, c.cols, K, oo);
    std::printf("eigen adapter ran on the device (status %d)\n", ex.miSystemState);
  } catch (const std::exception &ex) { std::printf("no device: %s\n", ex.what()); }
  std::printf("eigen adapter ok\n");
  return 0;
}

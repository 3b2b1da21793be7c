// ref_shim.cpp -- C entry points over the REFERENCE's own g2o::ellipsoid / edge classes, compiled from the sources where they lie under
// /root/reference by oracle/ref_slice/build_ref.sh (needs Eigen3 headers; never copied into this repo).  Test infrastructure: the
// signatures mirror oracle/eslam_oracle.cpp's orc_* functions so tests/test_ref_slice.py can diff the restatement against the real code.
#include "include/core/BasicEllipsoidEdges.h"
#include "include/core/Ellipsoid.h"

using namespace g2o;

namespace {
ellipsoid from10(const double *v) { Vector10d x; for (int i = 0; i < 10; i++) x[i] = v[i]; ellipsoid e; e.fromVector(x); return e; }
SE3Quat pose7(const double *v) { Vector7d x; for (int i = 0; i < 7; i++) x[i] = v[i]; SE3Quat T; T.fromVector(x); return T; }
Matrix3d calib(const double *K) { Matrix3d m; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = K[3 * i + j]; return m; }
}  // namespace

extern "C" {
// ellipsoid::getBoundingBoxFromProjection (src/core/Ellipsoid.cpp:303-307)
void ref_ellipsoid_bbox(const double e10[10], const double Tcw7[7], const double K[9], double out4[4]) {
  Vector4d b = from10(e10).getBoundingBoxFromProjection(pose7(Tcw7), calib(K));
  for (int i = 0; i < 4; i++) out4[i] = b[i];
}
// ellipsoid::min_log_error_9dof (src/core/Ellipsoid.cpp:92-117)
void ref_min_log_error_9dof(const double self10[10], const double new10[10], double out9[9]) {
  Vector9d r = from10(self10).min_log_error_9dof(from10(new10));
  for (int i = 0; i < 9; i++) out9[i] = r[i];
}
// ellipsoid::exp_update (src/core/Ellipsoid.cpp:38-47)
void ref_ellipsoid_exp_update(const double e10[10], const double u9[9], double out10[10]) {
  Vector9d u; for (int i = 0; i < 9; i++) u[i] = u9[i];
  Vector10d r = from10(e10).exp_update(u).toVector();
  for (int i = 0; i < 10; i++) out10[i] = r[i];
}
// ellipsoid::transform_from (src/core/Ellipsoid.cpp:120-129)
void ref_ellipsoid_transform_from(const double e10[10], const double T7[7], double out10[10]) {
  Vector10d r = from10(e10).transform_from(pose7(T7)).toVector();
  for (int i = 0; i < 10; i++) out10[i] = r[i];
}
// the three edge residuals of src/core/BasicEllipsoidEdges.cpp through the reference's own vertices / computeError()
//   type 0: EdgeSE3EllipsoidProj (meas = bbox4), 1: EdgeSE3Ellipsoid9DOF (meas = ellipsoid10), 2: EdgeEllipsoidGravityPlanePrior (meas = plane4)
int ref_edge_error(int type, const double Tcw7[7], const double e10[10], const double *meas, const double K[9], double out9[9]) {
  VertexSE3Expmap vSE3; vSE3.setId(0); vSE3.setEstimate(pose7(Tcw7));
  VertexEllipsoid vE; vE.setId(1); vE.setEstimate(from10(e10));
  for (int i = 0; i < 9; i++) out9[i] = 0;
  if (type == 0) {
    EdgeSE3EllipsoidProj e; e.setVertex(0, &vSE3); e.setVertex(1, &vE);
    Vector4d m; for (int i = 0; i < 4; i++) m[i] = meas[i];
    Matrix3d Km = calib(K);
    e.setMeasurement(m); e.setKalib(Km); e.computeError();
    for (int i = 0; i < 4; i++) out9[i] = e.error()[i];
    return 4;
  }
  if (type == 1) {
    EdgeSE3Ellipsoid9DOF e; e.setVertex(0, &vSE3); e.setVertex(1, &vE);
    e.setMeasurement(from10(meas)); e.computeError();
    for (int i = 0; i < 9; i++) out9[i] = e.error()[i];
    return 9;
  }
  EdgeEllipsoidGravityPlanePrior e; e.setVertex(0, &vE);
  Vector4d m; for (int i = 0; i < 4; i++) m[i] = meas[i];
  e.setMeasurement(m); e.computeError();
  out9[0] = e.error()[0];
  return 1;
}
}

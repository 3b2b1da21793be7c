// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, load or call anything under oracle/.
//
// orc_math.h: dependency-free fp64 restatement of the Eigen / g2o value-type arithmetic the reference's hot
// path relies on.  Each function cites the reference file:line (relative to the reference tree) or, where the
// arithmetic lives in an absent third-party dependency (Eigen 3.3.x, unpinned: CMakeLists.txt:32 "Eigen3 3.1.0
// REQUIRED"), the published algorithm it restates.  PARITY UNPINNED at the Eigen boundary: the reference cannot
// be compiled here (no Eigen/PCL/OpenCV), and it ships no golden vectors; see DESIGN.md.
//
// Compile with -ffp-contract=off (the reference builds without -march=native, so no FMA contraction).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>
#include <vector>

namespace orc {

struct V3 { double x, y, z; };
struct Q4 { double x, y, z, w; };          // Eigen::Quaterniond coefficient order x y z w
struct M3 { double m[3][3]; };             // m[row][col]
struct M4 { double m[4][4]; };
struct SE3 { Q4 r; V3 t; };                // g2o::SE3Quat {_r,_t}  (se3quat.h:47-48)

inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { double n = norm(a); return V3{a.x / n, a.y / n, a.z / n}; }
inline double get(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
inline void set(V3 &a, int i, double v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }

inline M3 m3_identity() { M3 r{}; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
inline M3 m3_zero() { M3 r{}; return r; }
inline V3 col(const M3 &a, int j) { return V3{a.m[0][j], a.m[1][j], a.m[2][j]}; }
inline void setcol(M3 &a, int j, V3 v) { a.m[0][j] = v.x; a.m[1][j] = v.y; a.m[2][j] = v.z; }
inline M3 mul(const M3 &a, const M3 &b) {
  M3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double s = a.m[i][0] * b.m[0][j];
    s += a.m[i][1] * b.m[1][j];
    s += a.m[i][2] * b.m[2][j];
    r.m[i][j] = s;
  }
  return r;
}
inline V3 mul(const M3 &a, V3 v) {
  return V3{a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline M3 transpose(const M3 &a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
inline M3 add(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
inline M3 scale(const M3 &a, double s) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j]; return r; }

// Eigen 3x3 inverse: cofactor matrix times 1/det (Eigen/src/LU/InverseImpl.h compute_inverse_size3_helper)
inline double det3(const M3 &a) {
  return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
         a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
         a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}
inline M3 inverse3(const M3 &a) {
  M3 c;
  c.m[0][0] = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
  c.m[0][1] = a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2];
  c.m[0][2] = a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1];
  c.m[1][0] = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
  c.m[1][1] = a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0];
  c.m[1][2] = a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2];
  c.m[2][0] = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
  c.m[2][1] = a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1];
  c.m[2][2] = a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0];
  double det = a.m[0][0] * c.m[0][0] + a.m[0][1] * c.m[1][0] + a.m[0][2] * c.m[2][0];
  double invdet = 1.0 / det;
  return scale(c, invdet);
}

inline M4 m4_identity() { M4 r{}; for (int i = 0; i < 4; i++) r.m[i][i] = 1; return r; }
inline M4 mul(const M4 &a, const M4 &b) {
  M4 r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
    double s = a.m[i][0] * b.m[0][j];
    for (int k = 1; k < 4; k++) s += a.m[i][k] * b.m[k][j];
    r.m[i][j] = s;
  }
  return r;
}
inline M4 transpose(const M4 &a) { M4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = a.m[j][i]; return r; }
// general 4x4 inverse by Gauss-Jordan with partial pivoting (Eigen uses a cofactor formula; equal to rounding)
inline bool inverse4(const M4 &a, M4 &out, double *det_out = nullptr) {
  double w[4][8];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { w[i][j] = a.m[i][j]; w[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  double det = 1.0;
  for (int c = 0; c < 4; c++) {
    int p = c; double best = std::fabs(w[c][c]);
    for (int r = c + 1; r < 4; r++) if (std::fabs(w[r][c]) > best) { best = std::fabs(w[r][c]); p = r; }
    if (best == 0.0) { if (det_out) *det_out = 0; return false; }
    if (p != c) { for (int j = 0; j < 8; j++) std::swap(w[p][j], w[c][j]); det = -det; }
    double piv = w[c][c]; det *= piv;
    for (int j = 0; j < 8; j++) w[c][j] /= piv;
    for (int r = 0; r < 4; r++) if (r != c) { double f = w[r][c]; if (f != 0.0) for (int j = 0; j < 8; j++) w[r][j] -= f * w[c][j]; }
  }
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out.m[i][j] = w[i][4 + j];
  if (det_out) *det_out = det;
  return true;
}

// ---- Eigen::Quaterniond semantics (Eigen/src/Geometry/Quaternion.h) ----
inline Q4 q_mul(Q4 a, Q4 b) {  // Quaternion product (internal::quat_product generic)
  return Q4{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
            a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Q4 q_conj(Q4 a) { return Q4{-a.x, -a.y, -a.z, a.w}; }
inline V3 q_rot(Q4 q, V3 v) {  // QuaternionBase::_transformVector
  V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}
inline M3 q_to_R(Q4 q) {  // QuaternionBase::toRotationMatrix
  M3 R;
  double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R.m[0][0] = 1.0 - (tyy + tzz); R.m[0][1] = txy - twz; R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz; R.m[1][1] = 1.0 - (txx + tzz); R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy; R.m[2][1] = tyz + twx; R.m[2][2] = 1.0 - (txx + tyy);
  return R;
}
inline Q4 q_from_R(const M3 &m) {  // quaternionbase_assign_impl<Other,3,3>::run
  Q4 q;
  double t = m.m[0][0] + m.m[1][1] + m.m[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m.m[2][1] - m.m[1][2]) * t;
    q.y = (m.m[0][2] - m.m[2][0]) * t;
    q.z = (m.m[1][0] - m.m[0][1]) * t;
  } else {
    int i = 0;
    if (m.m[1][1] > m.m[0][0]) i = 1;
    if (m.m[2][2] > m.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m.m[i][i] - m.m[j][j] - m.m[k][k] + 1.0);
    double c[3];
    c[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m.m[k][j] - m.m[j][k]) * t;
    c[j] = (m.m[j][i] + m.m[i][j]) * t;
    c[k] = (m.m[k][i] + m.m[i][k]) * t;
    q.x = c[0]; q.y = c[1]; q.z = c[2];
  }
  return q;
}
inline Q4 q_from_angle_axis(double angle, V3 axis) {  // Quaternion& operator=(const AngleAxisType&)
  double ha = 0.5 * angle;
  double s = std::sin(ha);
  return Q4{s * axis.x, s * axis.y, s * axis.z, std::cos(ha)};
}
inline M3 angle_axis_to_R(double angle, V3 axis) {  // AngleAxis::toRotationMatrix
  M3 res;
  double s = std::sin(angle), c = std::cos(angle);
  V3 sin_axis = s * axis;
  V3 cos1_axis = (1.0 - c) * axis;
  double tmp;
  tmp = cos1_axis.x * axis.y; res.m[0][1] = tmp - sin_axis.z; res.m[1][0] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z; res.m[0][2] = tmp + sin_axis.y; res.m[2][0] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z; res.m[1][2] = tmp - sin_axis.x; res.m[2][1] = tmp + sin_axis.x;
  res.m[0][0] = cos1_axis.x * axis.x + c;
  res.m[1][1] = cos1_axis.y * axis.y + c;
  res.m[2][2] = cos1_axis.z * axis.z + c;
  return res;
}

// ---- g2o::SE3Quat (Thirdparty/g2o/g2o/types/se3quat.h) ----
inline void se3_normalize_rotation(SE3 &T) {  // se3quat.h:345-350
  if (T.r.w < 0) { T.r.x *= -1; T.r.y *= -1; T.r.z *= -1; T.r.w *= -1; }
  double n = std::sqrt(T.r.x * T.r.x + T.r.y * T.r.y + T.r.z * T.r.z + T.r.w * T.r.w);
  T.r.x /= n; T.r.y /= n; T.r.z /= n; T.r.w /= n;
}
inline SE3 se3_identity() { return SE3{Q4{0, 0, 0, 1}, V3{0, 0, 0}}; }
inline SE3 se3_from_vector(const double v[7]) {  // fromVector, se3quat.h:160-163 (NO normalisation)
  return SE3{Q4{v[3], v[4], v[5], v[6]}, V3{v[0], v[1], v[2]}};
}
inline SE3 se3_from_vector_ctor(const double v[7]) {  // SE3Quat(const Vector7d&), se3quat.h:63-66 (normalises)
  SE3 T = se3_from_vector(v); se3_normalize_rotation(T); return T;
}
inline void se3_to_vector(const SE3 &T, double v[7]) {  // toVector, se3quat.h:147-157
  v[0] = T.t.x; v[1] = T.t.y; v[2] = T.t.z; v[3] = T.r.x; v[4] = T.r.y; v[5] = T.r.z; v[6] = T.r.w;
}
inline SE3 se3_make(Q4 q, V3 t) { SE3 T{q, t}; se3_normalize_rotation(T); return T; }  // SE3Quat(q,t), :59-61
inline SE3 se3_mul(const SE3 &a, const SE3 &b) {  // operator*, se3quat.h:110-116
  SE3 r = a;
  r.t = r.t + q_rot(a.r, b.t);
  r.r = q_mul(a.r, b.r);
  se3_normalize_rotation(r);
  return r;
}
inline SE3 se3_inverse(const SE3 &a) {  // se3quat.h:129-134
  SE3 r;
  r.r = q_conj(a.r);
  r.t = q_rot(r.r, a.t * -1.0);
  return r;
}
inline M4 se3_to_hom(const SE3 &T) {  // to_homogeneous_matrix, se3quat.h:335-343
  M4 h = m4_identity();
  M3 R = q_to_R(T.r);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) h.m[i][j] = R.m[i][j];
  h.m[0][3] = T.t.x; h.m[1][3] = T.t.y; h.m[2][3] = T.t.z;
  return h;
}
inline M3 skew(V3 v) {  // se3_ops.hpp:28-39
  M3 m = m3_zero();
  m.m[0][1] = -v.z; m.m[0][2] = v.y; m.m[1][2] = -v.x; m.m[1][0] = v.z; m.m[2][0] = -v.y; m.m[2][1] = v.x;
  return m;
}
inline V3 deltaR(const M3 &R) {  // se3_ops.hpp:41-48
  return V3{R.m[2][1] - R.m[1][2], R.m[0][2] - R.m[2][0], R.m[1][0] - R.m[0][1]};
}
inline void se3_log(const SE3 &T, double res[6]) {  // SE3Quat::log, se3quat.h:229-266 ; returns (omega, upsilon)
  M3 R = q_to_R(T.r);
  double d = 0.5 * (R.m[0][0] + R.m[1][1] + R.m[2][2] - 1);
  V3 omega;
  V3 dR = deltaR(R);
  M3 V_inv;
  if (d > 0.99999) {
    omega = 0.5 * dR;
    M3 Omega = skew(omega);
    V_inv = add(add(m3_identity(), scale(Omega, -0.5)), scale(mul(Omega, Omega), 1. / 12.));
  } else {
    double theta = std::acos(d);
    omega = (theta / (2 * std::sqrt(1 - d * d))) * dR;
    M3 Omega = skew(omega);
    V_inv = add(add(m3_identity(), scale(Omega, -0.5)),
                scale(mul(Omega, Omega), (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta)));
  }
  V3 upsilon = mul(V_inv, T.t);
  res[0] = omega.x; res[1] = omega.y; res[2] = omega.z;
  res[3] = upsilon.x; res[4] = upsilon.y; res[5] = upsilon.z;
}
inline SE3 se3_exp(const double update[6]) {  // SE3Quat::exp, se3quat.h:274-322
  V3 omega{update[0], update[1], update[2]};
  V3 upsilon{update[3], update[4], update[5]};
  double theta = norm(omega);
  M3 Omega = skew(omega);
  M3 R, V;
  if (theta < 0.00001) {
    R = add(add(m3_identity(), Omega), mul(Omega, Omega));
    V = R;
  } else {
    M3 Omega2 = mul(Omega, Omega);
    R = add(add(m3_identity(), scale(Omega, std::sin(theta) / theta)), scale(Omega2, (1 - std::cos(theta)) / (theta * theta)));
    V = add(add(m3_identity(), scale(Omega, (1 - std::cos(theta)) / (theta * theta))),
            scale(Omega2, (theta - std::sin(theta)) / (std::pow(theta, 3))));
  }
  return se3_make(q_from_R(R), mul(V, upsilon));
}
inline void quat_to_euler_zyx(Q4 q, double &roll, double &pitch, double &yaw) {  // se3quat.h:184-195
  roll = std::atan2(2 * (q.w * q.x + q.y * q.z), 1 - 2 * (q.x * q.x + q.y * q.y));
  pitch = std::asin(2 * (q.w * q.y - q.z * q.x));
  yaw = std::atan2(2 * (q.w * q.z + q.x * q.y), 1 - 2 * (q.y * q.y + q.z * q.z));
}
inline Q4 zyx_euler_to_quat(double roll, double pitch, double yaw) {  // matrix_utils.cpp:40-54
  double sy = std::sin(yaw * 0.5), cy = std::cos(yaw * 0.5);
  double sp = std::sin(pitch * 0.5), cp = std::cos(pitch * 0.5);
  double sr = std::sin(roll * 0.5), cr = std::cos(roll * 0.5);
  Q4 q;
  q.w = cr * cp * cy + sr * sp * sy;
  q.x = sr * cp * cy - cr * sp * sy;
  q.y = cr * sp * cy + sr * cp * sy;
  q.z = cr * cp * sy - sr * sp * cy;
  return q;
}
inline void rot_to_euler_zyx(const M3 &R, double &roll, double &pitch, double &yaw) {  // matrix_utils.cpp:75-95
  pitch = std::asin(-R.m[2][0]);
  if (std::fabs(pitch - M_PI / 2.0) < 1.0e-3) {
    roll = 0.0;
    yaw = std::atan2(R.m[1][2] - R.m[0][1], R.m[0][2] + R.m[1][1]) + roll;
  } else if (std::fabs(pitch + M_PI / 2.0) < 1.0e-3) {
    roll = 0.0;
    yaw = std::atan2(R.m[1][2] - R.m[0][1], R.m[0][2] + R.m[1][1]) - roll;
  } else {
    roll = std::atan2(R.m[2][1], R.m[2][2]);
    yaw = std::atan2(R.m[1][0], R.m[0][0]);
  }
}
// TransformPoint (Geometry.cpp:160-166): homogeneous multiply then divide by w
inline V3 transform_point_hom(const M4 &T, V3 p) {
  double X[4] = {p.x, p.y, p.z, 1.0};
  double Y[4];
  for (int i = 0; i < 4; i++) {
    double s = T.m[i][0] * X[0];
    s += T.m[i][1] * X[1]; s += T.m[i][2] * X[2]; s += T.m[i][3] * X[3];
    Y[i] = s;
  }
  return V3{Y[0] / Y[3], Y[1] / Y[3], Y[2] / Y[3]};
}

// ---- Eigen::SelfAdjointEigenSolver<Matrix3d>::compute (Eigen 3.3.x SelfAdjointEigenSolver.h: scaling,
// tridiagonalization_inplace_selector<.,3,false>, tridiagonal_qr_step with Wilkinson shift, ascending sort).
// Used at EllipsoidExtractor.cpp:179 and Initializer.cpp:233; eigenvector SIGNS follow this algorithm. ----
inline void givens(double p, double q, double &c, double &s) {  // JacobiRotation::makeGivens (real)
  if (q == 0.0) { c = p < 0.0 ? -1.0 : 1.0; s = 0.0; }
  else if (p == 0.0) { c = 0.0; s = q < 0.0 ? 1.0 : -1.0; }
  else if (std::fabs(p) > std::fabs(q)) {
    double t = q / p; double u = std::sqrt(1.0 + t * t); if (p < 0.0) u = -u;
    c = 1.0 / u; s = -t * c;
  } else {
    double t = p / q; double u = std::sqrt(1.0 + t * t); if (q < 0.0) u = -u;
    s = -1.0 / u; c = -t * s;
  }
}
inline double eig_hypot(double x, double y) {  // numext::hypot (positive_real_hypot)
  double ax = std::fabs(x), ay = std::fabs(y);
  double p = std::max(ax, ay);
  if (p == 0.0) return 0.0;
  double q = std::min(ax, ay);
  double qp = q / p;
  return p * std::sqrt(1.0 + qp * qp);
}
template <int N>
inline void tridiagonal_qr_step(double *diag, double *subdiag, int start, int end, double (*Q)[N]) {
  double td = (diag[end - 1] - diag[end]) * 0.5;
  double e = subdiag[end - 1];
  double mu = diag[end];
  if (td == 0.0) mu -= std::fabs(e);
  else {
    double e2 = subdiag[end - 1] * subdiag[end - 1];
    double h = eig_hypot(td, e);
    if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
    else mu -= e2 / (td + (td > 0.0 ? h : -h));
  }
  double x = diag[start] - mu;
  double z = subdiag[start];
  for (int k = start; k < end; ++k) {
    double c, s;
    givens(x, z, c, s);
    double sdk = s * diag[k] + c * subdiag[k];
    double dkp1 = s * subdiag[k] + c * diag[k + 1];
    diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
    diag[k + 1] = s * sdk + c * dkp1;
    subdiag[k] = c * sdk - s * dkp1;
    if (k > start) subdiag[k - 1] = c * subdiag[k - 1] - s * z;
    x = subdiag[k];
    if (k < end - 1) { z = -s * subdiag[k + 1]; subdiag[k + 1] = c * subdiag[k + 1]; }
    if (Q) {  // q.applyOnTheRight(k,k+1,rot): x_i' = c x_i - s y_i ; y_i' = s x_i + c y_i
      for (int i = 0; i < N; i++) {
        double xi = Q[i][k], yi = Q[i][k + 1];
        Q[i][k] = c * xi - s * yi;
        Q[i][k + 1] = s * xi + c * yi;
      }
    }
  }
}
template <int N>
inline bool tridiag_eigen(double *diag, double *subdiag, double (*Q)[N]) {  // computeFromTridiagonal_impl
  const int n = N;
  int end = n - 1, start = 0, iter = 0;
  const int maxIterations = 30;
  const double considerAsZero = std::numeric_limits<double>::min();
  const double precision = 2.0 * std::numeric_limits<double>::epsilon();
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (std::fabs(subdiag[i]) <= (std::fabs(diag[i]) + std::fabs(diag[i + 1])) * precision || std::fabs(subdiag[i]) <= considerAsZero)
        subdiag[i] = 0;
    while (end > 0 && subdiag[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && subdiag[start - 1] != 0) start--;
    tridiagonal_qr_step<N>(diag, subdiag, start, end, Q);
  }
  bool ok = iter <= maxIterations * n;
  if (ok) {
    for (int i = 0; i < n - 1; ++i) {
      int k = 0; double mn = diag[i];
      for (int j = 1; j < n - i; j++) if (diag[i + j] < mn) { mn = diag[i + j]; k = j; }
      if (k > 0) {
        std::swap(diag[i], diag[k + i]);
        if (Q) for (int r = 0; r < n; r++) std::swap(Q[r][i], Q[r][k + i]);
      }
    }
  }
  return ok;
}
// eigenvalues ascending in eval[3]; eigenvectors = columns of evec
inline void self_adjoint_eigen3(const M3 &A, double eval[3], M3 &evec) {
  double mat[3][3];
  // lower triangle, scaled to [-1,1]
  double scale = 0.0;
  for (int i = 0; i < 3; i++) for (int j = 0; j <= i; j++) scale = std::max(scale, std::fabs(A.m[i][j]));
  if (scale == 0.0) scale = 1.0;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) mat[i][j] = (j <= i) ? A.m[i][j] / scale : 0.0;
  double diag[3], subdiag[2];
  double Q[3][3];
  const double tol = std::numeric_limits<double>::min();
  diag[0] = mat[0][0];
  double v1norm2 = mat[2][0] * mat[2][0];
  if (v1norm2 <= tol) {
    diag[1] = mat[1][1]; diag[2] = mat[2][2];
    subdiag[0] = mat[1][0]; subdiag[1] = mat[2][1];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q[i][j] = (i == j) ? 1.0 : 0.0;
  } else {
    double beta = std::sqrt(mat[1][0] * mat[1][0] + v1norm2);
    double invBeta = 1.0 / beta;
    double m01 = mat[1][0] * invBeta;
    double m02 = mat[2][0] * invBeta;
    double q = 2.0 * m01 * mat[2][1] + m02 * (mat[2][2] - mat[1][1]);
    diag[1] = mat[1][1] + m02 * q;
    diag[2] = mat[2][2] - m02 * q;
    subdiag[0] = beta;
    subdiag[1] = mat[2][1] - m01 * q;
    double Qi[3][3] = {{1, 0, 0}, {0, m01, m02}, {0, m02, -m01}};
    std::memcpy(Q, Qi, sizeof(Q));
  }
  tridiag_eigen<3>(diag, subdiag, Q);
  for (int i = 0; i < 3; i++) eval[i] = diag[i] * scale;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) evec.m[i][j] = Q[i][j];
}

// cyclic Jacobi for small symmetric matrices (eigenvalues only need be accurate; used for the 4x4 sign test at
// Initializer.cpp:191-205 where only eigenvalue signs/values are consumed). ascending order.
inline void jacobi_eigen_sym(int n, std::vector<double> &A, std::vector<double> &eval) {
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      double apq = A[p * n + q];
      if (apq == 0.0) continue;
      double app = A[p * n + p], aqq = A[q * n + q];
      double tau = (aqq - app) / (2.0 * apq);
      double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
      double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
      for (int k = 0; k < n; k++) {
        double akp = A[k * n + p], akq = A[k * n + q];
        A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
      }
      for (int k = 0; k < n; k++) {
        double apk = A[p * n + k], aqk = A[q * n + k];
        A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
      }
    }
  }
  eval.resize(n);
  for (int i = 0; i < n; i++) eval[i] = A[i * n + i];
  std::sort(eval.begin(), eval.end());
}

// one-sided (Hestenes) Jacobi SVD of A (m x n, row-major): returns the right singular vector of the smallest
// singular value.  Stands in for Eigen::JacobiSVD thin-V last column (Initializer.cpp:170-173); the sign of the
// vector is arbitrary there as well and is cancelled downstream (Initializer.cpp:207-216).
inline void svd_null_vector(int m, int n, const std::vector<double> &Ain, std::vector<double> &v_out) {
  std::vector<double> A(Ain), V(n * n, 0.0);
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      double alpha = 0, beta = 0, gamma = 0;
      for (int k = 0; k < m; k++) { double ap = A[k * n + p], aq = A[k * n + q]; alpha += ap * ap; beta += aq * aq; gamma += ap * aq; }
      if (std::fabs(gamma) <= 1e-15 * std::sqrt(alpha * beta) || gamma == 0.0) continue;
      rotated = true;
      double zeta = (beta - alpha) / (2.0 * gamma);
      double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
      double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
      for (int k = 0; k < m; k++) { double ap = A[k * n + p], aq = A[k * n + q]; A[k * n + p] = c * ap - s * aq; A[k * n + q] = s * ap + c * aq; }
      for (int k = 0; k < n; k++) { double vp = V[k * n + p], vq = V[k * n + q]; V[k * n + p] = c * vp - s * vq; V[k * n + q] = s * vp + c * vq; }
    }
    if (!rotated) break;
  }
  int best = 0; double bestn = std::numeric_limits<double>::max();
  for (int j = 0; j < n; j++) { double s2 = 0; for (int k = 0; k < m; k++) s2 += A[k * n + j] * A[k * n + j]; if (s2 < bestn) { bestn = s2; best = j; } }
  v_out.resize(n);
  for (int k = 0; k < n; k++) v_out[k] = V[k * n + best];
}

// ---- Eigen::LDLT<MatrixXd> (Eigen 3.3.x Cholesky/LDLT.h: ldlt_inplace<Lower>::unblocked + _solve_impl), used by
// g2o LinearSolverDense (linear_solver_dense.h:103-113): returns false when !isPositive(). A is n x n row-major
// symmetric (full); x = A^-1 b. ----
inline bool ldlt_solve(int n, std::vector<double> mat, const double *b, double *x) {
  auto M = [&](int i, int j) -> double & { return mat[(size_t)i * n + j]; };
  std::vector<int> transp(n);
  std::vector<double> temp(n);
  int sign = 0;  // 0 ZeroSign, 1 PositiveSemiDef, -1 NegativeSemiDef, 2 Indefinite
  bool found_zero_pivot = false;
  bool ret = true;
  if (n <= 1) {
    if (n == 1) { transp[0] = 0; sign = M(0, 0) > 0 ? 1 : (M(0, 0) < 0 ? -1 : 0); }
  } else {
    for (int k = 0; k < n; ++k) {
      int big = k; double bv = std::fabs(M(k, k));
      for (int i = k + 1; i < n; i++) if (std::fabs(M(i, i)) > bv) { bv = std::fabs(M(i, i)); big = i; }
      transp[k] = big;
      if (k != big) {
        int s = n - big - 1;
        for (int j = 0; j < k; j++) std::swap(M(k, j), M(big, j));
        for (int i = 0; i < s; i++) std::swap(M(big + 1 + i, k), M(big + 1 + i, big));
        std::swap(M(k, k), M(big, big));
        for (int i = k + 1; i < big; ++i) { double tmp = M(i, k); M(i, k) = M(big, i); M(big, i) = tmp; }
      }
      int rs = n - k - 1;
      if (k > 0) {
        for (int j = 0; j < k; j++) temp[j] = M(j, j) * M(k, j);
        double acc = 0; bool first = true;
        for (int j = 0; j < k; j++) { double pr = M(k, j) * temp[j]; if (first) { acc = pr; first = false; } else acc += pr; }
        M(k, k) -= acc;
        for (int i = 0; i < rs; i++) {
          double a = 0; bool f = true;
          for (int j = 0; j < k; j++) { double pr = M(k + 1 + i, j) * temp[j]; if (f) { a = pr; f = false; } else a += pr; }
          M(k + 1 + i, k) -= a;
        }
      }
      double realAkk = M(k, k);
      bool pivot_is_valid = std::fabs(realAkk) > 0.0;
      if (k == 0 && !pivot_is_valid) {
        sign = 0;
        for (int j = 0; j < n; ++j) transp[j] = j;
        break;
      }
      if (rs > 0 && pivot_is_valid) for (int i = 0; i < rs; i++) M(k + 1 + i, k) /= realAkk;
      if (found_zero_pivot && pivot_is_valid) ret = false;
      else if (!pivot_is_valid) found_zero_pivot = true;
      if (sign == 1) { if (realAkk < 0) sign = 2; }
      else if (sign == -1) { if (realAkk > 0) sign = 2; }
      else if (sign == 0) { if (realAkk > 0) sign = 1; else if (realAkk < 0) sign = -1; }
    }
  }
  (void)ret;
  bool positive = (sign == 1 || sign == 0);
  if (!positive) return false;
  // solve: dst = P b ; L^-1 ; D^-1 (pseudo) ; L^-T ; P^-1
  std::vector<double> d(b, b + n);
  for (int k = 0; k < n; k++) if (transp[k] != k) std::swap(d[k], d[transp[k]]);
  for (int i = 0; i < n; i++) { double s = d[i]; for (int j = 0; j < i; j++) s -= M(i, j) * d[j]; d[i] = s; }
  const double tolerance = 1.0 / std::numeric_limits<double>::max();
  for (int i = 0; i < n; i++) { if (std::fabs(M(i, i)) > tolerance) d[i] /= M(i, i); else d[i] = 0; }
  for (int i = n - 1; i >= 0; i--) { double s = d[i]; for (int j = i + 1; j < n; j++) s -= M(j, i) * d[j]; d[i] = s; }
  for (int k = n - 1; k >= 0; k--) if (transp[k] != k) std::swap(d[k], d[transp[k]]);
  for (int i = 0; i < n; i++) x[i] = d[i];
  return true;
}

}  // namespace orc

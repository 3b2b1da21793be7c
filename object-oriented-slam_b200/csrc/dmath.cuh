// dmath.cuh -- fp64 device math for the quadric hot path (sm_100a).
//
// Value types and operation ORDER follow the reference's Eigen/g2o arithmetic so that, compiled with -fmad=false,
// every +,-,*,/,sqrt below rounds exactly like the reference's x86-64 build (no FMA there: CMakeLists.txt:4-12 has
// no -march flag).  Only libm calls (sin/cos/atan2/acos/tan/exp) may differ from glibc by <= 1-2 ulp.
//   g2o::SE3Quat      Thirdparty/g2o/g2o/types/se3quat.h:40-361
//   g2o::ellipsoid    src/core/Ellipsoid.cpp:16-307
//   g2o::plane        include/core/Plane.h:46-109, src/core/Plane.cpp:97-144
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>

#define DEV __device__ __forceinline__

namespace eslam {

struct d3 { double x, y, z; };
struct quat { double x, y, z, w; };
struct mat3 { double a[9]; };  // row-major
struct se3 { quat r; d3 t; };
struct ellipsoid { se3 pose; double s[3]; };
struct dplane { double p[4]; double dual; };

DEV d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
DEV d3 add3(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
DEV d3 sub3(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
DEV d3 scl3(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
DEV double dot3(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DEV d3 cross3(d3 a, d3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DEV double norm3(d3 a) { return sqrt(dot3(a, a)); }
DEV d3 unit3(d3 a) { double n = norm3(a); return mk3(a.x / n, a.y / n, a.z / n); }

DEV mat3 m3_eye() { mat3 m; m.a[0] = 1; m.a[1] = 0; m.a[2] = 0; m.a[3] = 0; m.a[4] = 1; m.a[5] = 0; m.a[6] = 0; m.a[7] = 0; m.a[8] = 1; return m; }
DEV mat3 m3_mul(const mat3 &A, const mat3 &B) {
  mat3 C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = A.a[i * 3] * B.a[j];
      s += A.a[i * 3 + 1] * B.a[3 + j];
      s += A.a[i * 3 + 2] * B.a[6 + j];
      C.a[i * 3 + j] = s;
    }
  return C;
}
DEV d3 m3_vec(const mat3 &A, d3 v) {
  return mk3(A.a[0] * v.x + A.a[1] * v.y + A.a[2] * v.z, A.a[3] * v.x + A.a[4] * v.y + A.a[5] * v.z, A.a[6] * v.x + A.a[7] * v.y + A.a[8] * v.z);
}
DEV d3 m3_col(const mat3 &A, int j) { return mk3(A.a[j], A.a[3 + j], A.a[6 + j]); }
DEV void m3_setcol(mat3 &A, int j, d3 v) { A.a[j] = v.x; A.a[3 + j] = v.y; A.a[6 + j] = v.z; }
DEV mat3 m3_skew(d3 v) {  // se3_ops.hpp:28-39
  mat3 m;
  m.a[0] = 0; m.a[1] = -v.z; m.a[2] = v.y; m.a[3] = v.z; m.a[4] = 0; m.a[5] = -v.x; m.a[6] = -v.y; m.a[7] = v.x; m.a[8] = 0;
  return m;
}
// a*I + b*W + c*W2  evaluated as (I + b*W) + c*W2 with the identity contributing first (Eigen expression order)
DEV mat3 m3_lin3(const mat3 &W, double b, const mat3 &W2, double c) {
  mat3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) { double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0; r.a[i] = (id + b * W.a[i]) + c * W2.a[i]; }
  return r;
}
DEV mat3 m3_inverse(const mat3 &M) {  // Eigen compute_inverse_size3: cofactors * (1/det)
  mat3 c;
  c.a[0] = M.a[4] * M.a[8] - M.a[5] * M.a[7];
  c.a[1] = M.a[2] * M.a[7] - M.a[1] * M.a[8];
  c.a[2] = M.a[1] * M.a[5] - M.a[2] * M.a[4];
  c.a[3] = M.a[5] * M.a[6] - M.a[3] * M.a[8];
  c.a[4] = M.a[0] * M.a[8] - M.a[2] * M.a[6];
  c.a[5] = M.a[2] * M.a[3] - M.a[0] * M.a[5];
  c.a[6] = M.a[3] * M.a[7] - M.a[4] * M.a[6];
  c.a[7] = M.a[1] * M.a[6] - M.a[0] * M.a[7];
  c.a[8] = M.a[0] * M.a[4] - M.a[1] * M.a[3];
  double det = M.a[0] * c.a[0] + M.a[1] * c.a[3] + M.a[2] * c.a[6];
  double inv = 1.0 / det;
#pragma unroll
  for (int i = 0; i < 9; i++) c.a[i] = inv * c.a[i];
  return c;
}

// ---- Eigen::Quaterniond ----
DEV quat q_mul(quat a, quat b) {
  quat r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r;
}
DEV quat q_conj(quat a) { quat r; r.x = -a.x; r.y = -a.y; r.z = -a.z; r.w = a.w; return r; }
DEV d3 q_rot(quat q, d3 v) {
  d3 qv = mk3(q.x, q.y, q.z);
  d3 uv = cross3(qv, v);
  uv = add3(uv, uv);
  return add3(add3(v, scl3(q.w, uv)), cross3(qv, uv));
}
DEV mat3 q_to_R(quat q) {
  mat3 R;
  double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R.a[0] = 1.0 - (tyy + tzz); R.a[1] = txy - twz; R.a[2] = txz + twy;
  R.a[3] = txy + twz; R.a[4] = 1.0 - (txx + tzz); R.a[5] = tyz - twx;
  R.a[6] = txz - twy; R.a[7] = tyz + twx; R.a[8] = 1.0 - (txx + tyy);
  return R;
}
DEV quat q_from_R(const mat3 &m) {
  quat q;
  double t = m.a[0] + m.a[4] + m.a[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m.a[7] - m.a[5]) * t;
    q.y = (m.a[2] - m.a[6]) * t;
    q.z = (m.a[3] - m.a[1]) * t;
  } else {
    int i = 0;
    if (m.a[4] > m.a[0]) i = 1;
    if (m.a[8] > m.a[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m.a[i * 4] - m.a[j * 4] - m.a[k * 4] + 1.0);
    double c[3];
    c[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m.a[k * 3 + j] - m.a[j * 3 + k]) * t;
    c[j] = (m.a[j * 3 + i] + m.a[i * 3 + j]) * t;
    c[k] = (m.a[k * 3 + i] + m.a[i * 3 + k]) * t;
    q.x = c[0]; q.y = c[1]; q.z = c[2];
  }
  return q;
}
DEV quat q_axis_angle(double angle, d3 axis) {
  double ha = 0.5 * angle;
  // sin(+-0) = +-0 and cos(+-0) = 1 exactly: the symmetry planes are vertical, so half of these calls see a zero elevation
  const bool zero = (ha == 0.0);
  double s = zero ? ha : sin(ha);
  quat q; q.x = s * axis.x; q.y = s * axis.y; q.z = s * axis.z; q.w = zero ? 1.0 : cos(ha);
  return q;
}
DEV mat3 R_axis_angle(double angle, d3 axis) {  // Eigen::AngleAxis::toRotationMatrix
  mat3 r;
  double s = sin(angle), c = cos(angle);
  d3 sa = scl3(s, axis);
  d3 ca = scl3(1.0 - c, axis);
  double tmp;
  tmp = ca.x * axis.y; r.a[1] = tmp - sa.z; r.a[3] = tmp + sa.z;
  tmp = ca.x * axis.z; r.a[2] = tmp + sa.y; r.a[6] = tmp - sa.y;
  tmp = ca.y * axis.z; r.a[5] = tmp - sa.x; r.a[7] = tmp + sa.x;
  r.a[0] = ca.x * axis.x + c; r.a[4] = ca.y * axis.y + c; r.a[8] = ca.z * axis.z + c;
  return r;
}

// ---- g2o::SE3Quat ----
DEV void se3_normalize(se3 &T) {
  if (T.r.w < 0) { T.r.x *= -1; T.r.y *= -1; T.r.z *= -1; T.r.w *= -1; }
  double n = sqrt(T.r.x * T.r.x + T.r.y * T.r.y + T.r.z * T.r.z + T.r.w * T.r.w);
  T.r.x /= n; T.r.y /= n; T.r.z /= n; T.r.w /= n;
}
DEV se3 se3_load(const double *v) { se3 T; T.t = mk3(v[0], v[1], v[2]); T.r.x = v[3]; T.r.y = v[4]; T.r.z = v[5]; T.r.w = v[6]; return T; }
DEV void se3_store(const se3 &T, double *v) { v[0] = T.t.x; v[1] = T.t.y; v[2] = T.t.z; v[3] = T.r.x; v[4] = T.r.y; v[5] = T.r.z; v[6] = T.r.w; }
DEV se3 se3_mul(const se3 &a, const se3 &b) {
  se3 r;
  r.t = add3(a.t, q_rot(a.r, b.t));
  r.r = q_mul(a.r, b.r);
  se3_normalize(r);
  return r;
}
DEV se3 se3_inv(const se3 &a) {
  se3 r;
  r.r = q_conj(a.r);
  r.t = q_rot(r.r, scl3(-1.0, mk3(a.t.x, a.t.y, a.t.z)));
  return r;
}
// note: (_t * -1.) in the reference is t*(-1) -- identical to (-1)*t in IEEE
DEV void se3_log(const se3 &T, double *res) {
  mat3 R = q_to_R(T.r);
  double d = 0.5 * (R.a[0] + R.a[4] + R.a[8] - 1);
  d3 dR = mk3(R.a[7] - R.a[5], R.a[2] - R.a[6], R.a[3] - R.a[1]);
  d3 omega;
  mat3 Vinv;
  if (d > 0.99999) {
    omega = scl3(0.5, dR);
    mat3 W = m3_skew(omega);
    Vinv = m3_lin3(W, -0.5, m3_mul(W, W), 1. / 12.);
  } else {
    double theta = acos(d);
    omega = scl3(theta / (2 * sqrt(1 - d * d)), dR);
    mat3 W = m3_skew(omega);
    Vinv = m3_lin3(W, -0.5, m3_mul(W, W), (1 - theta / (2 * tan(theta / 2))) / (theta * theta));
  }
  d3 ups = m3_vec(Vinv, T.t);
  res[0] = omega.x; res[1] = omega.y; res[2] = omega.z; res[3] = ups.x; res[4] = ups.y; res[5] = ups.z;
}
DEV se3 se3_exp(const double *u) {
  d3 omega = mk3(u[0], u[1], u[2]);
  d3 ups = mk3(u[3], u[4], u[5]);
  double theta = norm3(omega);
  mat3 W = m3_skew(omega);
  mat3 W2 = m3_mul(W, W);
  mat3 R, V;
  if (theta < 0.00001) {
    R = m3_lin3(W, 1.0, W2, 1.0);
    V = R;
  } else {
    double st = sin(theta), ct = cos(theta);
    R = m3_lin3(W, st / theta, W2, (1 - ct) / (theta * theta));
    V = m3_lin3(W, (1 - ct) / (theta * theta), W2, (theta - st) / (pow(theta, 3.0)));
  }
  se3 T;
  T.r = q_from_R(R);
  T.t = m3_vec(V, ups);
  se3_normalize(T);
  return T;
}
// TransformPoint (Geometry.cpp:160-166): 4x4 homogeneous multiply, then divide by w (w == 1 exactly)
struct hom34 { double m[12]; };
DEV hom34 se3_hom(const se3 &T) {
  mat3 R = q_to_R(T.r);
  hom34 h;
  h.m[0] = R.a[0]; h.m[1] = R.a[1]; h.m[2] = R.a[2]; h.m[3] = T.t.x;
  h.m[4] = R.a[3]; h.m[5] = R.a[4]; h.m[6] = R.a[5]; h.m[7] = T.t.y;
  h.m[8] = R.a[6]; h.m[9] = R.a[7]; h.m[10] = R.a[8]; h.m[11] = T.t.z;
  return h;
}
DEV d3 hom_apply(const hom34 &h, d3 p) {
  // rows 0..2 of T*[p;1]; last row gives w = 0*x+0*y+0*z+1*1 = 1 and x/1 == x exactly
  double X = h.m[0] * p.x; X += h.m[1] * p.y; X += h.m[2] * p.z; X += h.m[3] * 1.0;
  double Y = h.m[4] * p.x; Y += h.m[5] * p.y; Y += h.m[6] * p.z; Y += h.m[7] * 1.0;
  double Z = h.m[8] * p.x; Z += h.m[9] * p.y; Z += h.m[10] * p.z; Z += h.m[11] * 1.0;
  return mk3(X, Y, Z);
}

// ---- g2o::ellipsoid ----
DEV ellipsoid ell_load(const double *v) { ellipsoid e; e.pose = se3_load(v); e.s[0] = v[7]; e.s[1] = v[8]; e.s[2] = v[9]; return e; }
DEV void ell_store(const ellipsoid &e, double *v) { se3_store(e.pose, v); v[7] = e.s[0]; v[8] = e.s[1]; v[9] = e.s[2]; }
DEV ellipsoid ell_oplus(const ellipsoid &e, const double *u) {  // exp_update Ellipsoid.cpp:38-47
  ellipsoid r;
  r.pose = se3_mul(e.pose, se3_exp(u));
  r.s[0] = e.s[0] + u[6]; r.s[1] = e.s[1] + u[7]; r.s[2] = e.s[2] + u[8];
  return r;
}
// P = K[I|0] * hom(Tcw) : 3x4 row-major (Ellipsoid.cpp:276-287)
DEV void proj_matrix(const se3 &Tcw, const double *K, double *P) {
  hom34 T = se3_hom(Tcw);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double s = K[i * 3] * T.m[j];
      s += K[i * 3 + 1] * T.m[4 + j];
      s += K[i * 3 + 2] * T.m[8 + j];
      s += 0.0 * (j == 3 ? 1.0 : 0.0);
      P[i * 4 + j] = s;
    }
}
// getBoundingBoxFromProjection (Ellipsoid.cpp:209-273,303-307)
DEV void ell_bbox(const ellipsoid &e, const se3 &Tcw, const double *K, double *out) {
  // Q* = T diag(1/(1/a^2),...,-1) T^T, T = hom(pose) 4x4 with last row (0,0,0,1)
  hom34 T = se3_hom(e.pose);
  double dg[4];
  dg[0] = 1.0 / (1 / (e.s[0] * e.s[0])); dg[1] = 1.0 / (1 / (e.s[1] * e.s[1])); dg[2] = 1.0 / (1 / (e.s[2] * e.s[2])); dg[3] = 1.0 / (-1.0);
  double TQ[16];  // (T * Qc)  4x4
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double tij = (i < 3) ? T.m[i * 4 + j] : (j == 3 ? 1.0 : 0.0);
      TQ[i * 4 + j] = tij * dg[j];  // sum over k of T[i][k]*Qc[k][j] has a single non-zero term (others add exact zeros)
    }
  double Qs[16];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double tjk = (j < 3) ? T.m[j * 4 + k] : (k == 3 ? 1.0 : 0.0);
        double pr = TQ[i * 4 + k] * tjk;
        s = (k == 0) ? pr : s + pr;
      }
      Qs[i * 4 + j] = s;
    }
  double P[12];
  proj_matrix(Tcw, K, P);
  double PQ[12];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double s = P[i * 4] * Qs[j];
      s += P[i * 4 + 1] * Qs[4 + j]; s += P[i * 4 + 2] * Qs[8 + j]; s += P[i * 4 + 3] * Qs[12 + j];
      PQ[i * 4 + j] = s;
    }
  mat3 Cs;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = PQ[i * 4] * P[j * 4];
      s += PQ[i * 4 + 1] * P[j * 4 + 1]; s += PQ[i * 4 + 2] * P[j * 4 + 2]; s += PQ[i * 4 + 3] * P[j * 4 + 3];
      Cs.a[i * 3 + j] = s;
    }
  mat3 C = m3_inverse(Cs);
  double c22 = C.a[8];
#pragma unroll
  for (int i = 0; i < 9; i++) C.a[i] = C.a[i] / c22;
  double a = C.a[0], b = C.a[1] * 2, c = C.a[4], d = C.a[2] * 2, ee = C.a[7] * 2;
  double theta = 1 / 2.0 * atan2(b, (a - c));
  double x_c = (b * ee - 2 * c * d) / (4 * a * c - b * b);
  double y_c = (b * d - 2 * a * ee) / (4 * a * c - b * b);
  double a_2 = 2 * (a * x_c * x_c + c * y_c * y_c + b * x_c * y_c - 1) / (a + c + sqrt((a - c) * (a - c) + b * b));
  double b_2 = 2 * (a * x_c * x_c + c * y_c * y_c + b * x_c * y_c - 1) / (a + c - sqrt((a - c) * (a - c) + b * b));
  double ax1 = sqrt(a_2), ax2 = sqrt(b_2);
  double ct2 = cos(theta) * cos(theta);
  double st2 = 1 - ct2;
  double xl = sqrt(ax1 * ax1 * ct2 + ax2 * ax2 * st2);
  double yl = sqrt(ax1 * ax1 * st2 + ax2 * ax2 * ct2);
  out[0] = x_c - xl; out[1] = y_c - yl; out[2] = x_c + xl; out[3] = y_c + yl;
}
// min_log_error_9dof (Ellipsoid.cpp:63-117): self = object vertex, other = measurement in world
DEV void ell_min_log_error(const ellipsoid &self, const ellipsoid &other, double *out) {
  double best = 0; bool have = false;
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    double yaw = (double)(i - 1) * M_PI / 2.0;
    ellipsoid r = other;
    se3 rot; rot.t = mk3(0, 0, 0); rot.r.x = 0; rot.r.y = 0; rot.r.z = sin(yaw * 0.5); rot.r.w = cos(yaw * 0.5);
    se3_normalize(rot);
    r.pose = se3_mul(other.pose, rot);
    if (i == 0 || i == 2) { double t = r.s[0]; r.s[0] = r.s[1]; r.s[1] = t; }  // |yaw -/+ pi/2| < 1e-6  (3pi/2 never occurs)
    double e[9];
    se3_log(se3_mul(se3_inv(r.pose), self.pose), e);
    e[6] = self.s[0] - r.s[0]; e[7] = self.s[1] - r.s[1]; e[8] = self.s[2] - r.s[2];
    double s = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) s += e[k] * e[k];
    double nrm = sqrt(s);
    // Eigen minCoeff visitor: init with coefficient 0, then strict <  (NaN in slot 0 poisons the result)
    if (!have || nrm < best) {
      if (!have || nrm < best) { best = nrm; have = true;
#pragma unroll
        for (int k = 0; k < 9; k++) out[k] = e[k]; }
    }
  }
}

// ---- g2o::plane ----
DEV double plane_signed_dist(const double *p, d3 pt) {  // Plane.cpp:108-115 (keep_flag = true)
  double fenzi = p[0] * pt.x + p[1] * pt.y + p[2] * pt.z + p[3];
  double fenmu = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  return fenzi / fenmu;
}
DEV mat3 plane_rotation(d3 v) {  // Plane.h:88-92
  double az = atan2(v.y, v.x);
  const double rxy = sqrt(v.x * v.x + v.y * v.y);
  double el = (v.z == 0.0 && rxy > 0.0) ? v.z : atan2(v.z, rxy);   // atan2(+-0, r > 0) = +-0 exactly
  quat q = q_mul(q_axis_angle(az, mk3(0, 0, 1)), q_axis_angle(-el, mk3(0, 1, 0)));
  return q_to_R(q);
}
DEV void plane_oplus_common(dplane &pl, double az, double el, double dd) {
  const bool el0 = (el == 0.0);
  double s = el0 ? el : sin(el), c = el0 ? 1.0 : cos(el);   // exact for +-0 (VertexPlane / VertexDualPlane always pass el = 0)
  d3 n = mk3(c * cos(az), c * sin(az), s);
  mat3 R = plane_rotation(mk3(pl.p[0], pl.p[1], pl.p[2]));
  double d = -pl.p[3] + dd;
  d3 rn = m3_vec(R, n);
  pl.p[0] = rn.x; pl.p[1] = rn.y; pl.p[2] = rn.z; pl.p[3] = -d;
  double nn = sqrt(pl.p[0] * pl.p[0] + pl.p[1] * pl.p[1] + pl.p[2] * pl.p[2]);
  double inv = 1. / nn;
  pl.p[0] = pl.p[0] * inv; pl.p[1] = pl.p[1] * inv; pl.p[2] = pl.p[2] * inv; pl.p[3] = pl.p[3] * inv;
}
DEV void plane_another(const dplane &pl, double *out) {  // Plane.cpp:137-144 + :97-105
  double ang = atan2(pl.p[1], pl.p[0]);
  out[0] = sin(ang); out[1] = -cos(ang); out[2] = 0; out[3] = -pl.dual;
}
DEV d3 mirror_point(d3 pt, const double *pl) {  // SymmetrySolver.cpp:193-214
  double down = sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]);
  d3 normal = mk3(pl[0] / down, pl[1] / down, pl[2] / down);
  double val = pl[0] * pt.x + pl[1] * pt.y + pl[2] * pt.z + pl[3] * 1.0;
  double dis = fabs(val) / down;
  double symbol = (val > 0) ? -1.0 : 1.0;
  double k = 2 * symbol * dis;
  return mk3(pt.x + k * normal.x, pt.y + k * normal.y, pt.z + k * normal.z);
}
// (T^T)^-1 applied to a plane for rigid T = [R t]: n' = R n, d' = d - t.n'   (Plane.cpp:117-122, closed form)
DEV void plane_transform_rigid(const se3 &T, const double *p, double *out) {
  mat3 R = q_to_R(T.r);
  d3 n = m3_vec(R, mk3(p[0], p[1], p[2]));
  out[0] = n.x; out[1] = n.y; out[2] = n.z; out[3] = p[3] - dot3(T.t, n);
}

// ---- Eigen::SelfAdjointEigenSolver<Matrix3d>::compute (3.3.x): scale, closed-form tridiagonalisation, implicit QR ----
DEV void givens(double p, double q, double &c, double &s) {
  if (q == 0.0) { c = p < 0.0 ? -1.0 : 1.0; s = 0.0; }
  else if (p == 0.0) { c = 0.0; s = q < 0.0 ? 1.0 : -1.0; }
  else if (fabs(p) > fabs(q)) { double t = q / p; double u = sqrt(1.0 + t * t); if (p < 0.0) u = -u; c = 1.0 / u; s = -t * c; }
  else { double t = p / q; double u = sqrt(1.0 + t * t); if (q < 0.0) u = -u; s = -1.0 / u; c = -t * s; }
}
DEV void eig3_sym(const mat3 &A, double *ev, mat3 &V) {
  double scale = 0.0;
  for (int i = 0; i < 3; i++) for (int j = 0; j <= i; j++) scale = fmax(scale, fabs(A.a[i * 3 + j]));
  if (scale == 0.0) scale = 1.0;
  double m00 = A.a[0] / scale, m10 = A.a[3] / scale, m11 = A.a[4] / scale, m20 = A.a[6] / scale, m21 = A.a[7] / scale, m22 = A.a[8] / scale;
  double diag[3], sub[2];
  double Q[9];
  diag[0] = m00;
  double v1norm2 = m20 * m20;
  if (v1norm2 <= DBL_MIN) {
    diag[1] = m11; diag[2] = m22; sub[0] = m10; sub[1] = m21;
    for (int i = 0; i < 9; i++) Q[i] = (i % 4 == 0) ? 1.0 : 0.0;
  } else {
    double beta = sqrt(m10 * m10 + v1norm2);
    double invBeta = 1.0 / beta;
    double m01 = m10 * invBeta, m02 = m20 * invBeta;
    double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q; diag[2] = m22 - m02 * q;
    sub[0] = beta; sub[1] = m21 - m01 * q;
    Q[0] = 1; Q[1] = 0; Q[2] = 0; Q[3] = 0; Q[4] = m01; Q[5] = m02; Q[6] = 0; Q[7] = m02; Q[8] = -m01;
  }
  int end = 2, start = 0, iter = 0;
  const double prec = 2.0 * DBL_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (fabs(sub[i]) <= (fabs(diag[i]) + fabs(diag[i + 1])) * prec || fabs(sub[i]) <= DBL_MIN) sub[i] = 0;
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 90) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0) start--;
    // tridiagonal_qr_step
    double td = (diag[end - 1] - diag[end]) * 0.5;
    double e = sub[end - 1];
    double mu = diag[end];
    if (td == 0.0) mu -= fabs(e);
    else {
      double e2 = sub[end - 1] * sub[end - 1];
      double ax = fabs(td), ay = fabs(e);
      double pp = fmax(ax, ay), qq = fmin(ax, ay);
      double h = 0.0;
      if (pp != 0.0) { double qp = qq / pp; h = pp * sqrt(1.0 + qp * qp); }
      if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu;
    double z = sub[start];
    for (int k = start; k < end; ++k) {
      double c, s;
      givens(x, z, c, s);
      double sdk = s * diag[k] + c * sub[k];
      double dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
      for (int i = 0; i < 3; i++) {
        double xi = Q[i * 3 + k], yi = Q[i * 3 + k + 1];
        Q[i * 3 + k] = c * xi - s * yi;
        Q[i * 3 + k + 1] = s * xi + c * yi;
      }
    }
  }
  for (int i = 0; i < 2; ++i) {
    int k = 0; double mn = diag[i];
    for (int j = 1; j < 3 - i; j++) if (diag[i + j] < mn) { mn = diag[i + j]; k = j; }
    if (k > 0) {
      double t = diag[i]; diag[i] = diag[k + i]; diag[k + i] = t;
      for (int r = 0; r < 3; r++) { double u = Q[r * 3 + i]; Q[r * 3 + i] = Q[r * 3 + k + i]; Q[r * 3 + k + i] = u; }
    }
  }
  for (int i = 0; i < 3; i++) ev[i] = diag[i] * scale;
  for (int i = 0; i < 9; i++) V.a[i] = Q[i];
}

// ---- Eigen::LDLT (3.3.x) for small dense SPD systems, N <= 9: returns false when !isPositive() ----
template <int N>
DEV bool ldlt_solve_small(double *M /*N*N row-major, destroyed*/, const double *b, double *x) {
  int tr[N];
  double temp[N];
  int sign = 0;
  if (N == 1) { tr[0] = 0; sign = M[0] > 0 ? 1 : (M[0] < 0 ? -1 : 0); }
  else {
    for (int k = 0; k < N; ++k) {
      int big = k; double bv = fabs(M[k * N + k]);
      for (int i = k + 1; i < N; i++) { double v = fabs(M[i * N + i]); if (v > bv) { bv = v; big = i; } }
      tr[k] = big;
      if (k != big) {
        int s = N - big - 1;
        for (int j = 0; j < k; j++) { double t = M[k * N + j]; M[k * N + j] = M[big * N + j]; M[big * N + j] = t; }
        for (int i = 0; i < s; i++) { double t = M[(big + 1 + i) * N + k]; M[(big + 1 + i) * N + k] = M[(big + 1 + i) * N + big]; M[(big + 1 + i) * N + big] = t; }
        { double t = M[k * N + k]; M[k * N + k] = M[big * N + big]; M[big * N + big] = t; }
        for (int i = k + 1; i < big; ++i) { double t = M[i * N + k]; M[i * N + k] = M[big * N + i]; M[big * N + i] = t; }
      }
      int rs = N - k - 1;
      if (k > 0) {
        for (int j = 0; j < k; j++) temp[j] = M[j * N + j] * M[k * N + j];
        double acc = 0;
        for (int j = 0; j < k; j++) { double pr = M[k * N + j] * temp[j]; acc = (j == 0) ? pr : acc + pr; }
        M[k * N + k] -= acc;
        for (int i = 0; i < rs; i++) {
          double a = 0;
          for (int j = 0; j < k; j++) { double pr = M[(k + 1 + i) * N + j] * temp[j]; a = (j == 0) ? pr : a + pr; }
          M[(k + 1 + i) * N + k] -= a;
        }
      }
      double akk = M[k * N + k];
      bool valid = fabs(akk) > 0.0;
      if (k == 0 && !valid) { sign = 0; for (int j = 0; j < N; ++j) tr[j] = j; break; }
      if (rs > 0 && valid) for (int i = 0; i < rs; i++) M[(k + 1 + i) * N + k] /= akk;
      if (sign == 1) { if (akk < 0) sign = 2; }
      else if (sign == -1) { if (akk > 0) sign = 2; }
      else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
    }
  }
  if (!(sign == 1 || sign == 0)) return false;
  double d[N];
  for (int i = 0; i < N; i++) d[i] = b[i];
  for (int k = 0; k < N; k++) if (tr[k] != k) { double t = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t; }
  for (int i = 0; i < N; i++) { double s = d[i]; for (int j = 0; j < i; j++) s -= M[i * N + j] * d[j]; d[i] = s; }
  const double tol = 1.0 / DBL_MAX;
  for (int i = 0; i < N; i++) { if (fabs(M[i * N + i]) > tol) d[i] /= M[i * N + i]; else d[i] = 0; }
  for (int i = N - 1; i >= 0; i--) { double s = d[i]; for (int j = i + 1; j < N; j++) s -= M[j * N + i] * d[j]; d[i] = s; }
  for (int k = N - 1; k >= 0; k--) if (tr[k] != k) { double t = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t; }
  for (int i = 0; i < N; i++) x[i] = d[i];
  return true;
}

}  // namespace eslam

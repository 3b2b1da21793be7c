// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
// orc_graph.h: g2o::ellipsoid value type, the three ellipsoid edges, g2o numeric Jacobians and the
// Levenberg-Marquardt driver, restated from the reference (file:line cited per function).
#pragma once
#include "orc_math.h"
#include <functional>

namespace orc {

struct Ellipsoid {  // g2o::ellipsoid (include/core/Ellipsoid.h:26-136): pose + scale (+ prob, ids)
  SE3 pose;
  double s[3];
};
inline Ellipsoid ell_from_vector(const double v[10]) {  // Ellipsoid.cpp:24-29 (fromVector: no normalisation)
  Ellipsoid e; e.pose = se3_from_vector(v); e.s[0] = v[7]; e.s[1] = v[8]; e.s[2] = v[9]; return e;
}
inline void ell_to_vector(const Ellipsoid &e, double v[10]) {  // Ellipsoid.cpp:152-157
  se3_to_vector(e.pose, v); v[7] = e.s[0]; v[8] = e.s[1]; v[9] = e.s[2];
}
inline Ellipsoid ell_from_minimal(const double v[9]) {  // Ellipsoid.cpp:16-22
  Ellipsoid e; e.pose = se3_make(zyx_euler_to_quat(v[3], v[4], v[5]), V3{v[0], v[1], v[2]});
  e.s[0] = v[6]; e.s[1] = v[7]; e.s[2] = v[8]; return e;
}
inline void ell_to_minimal(const Ellipsoid &e, double v[9]) {  // Ellipsoid.cpp:144-149, se3quat.h:197-208
  double r, p, y; quat_to_euler_zyx(e.pose.r, r, p, y);
  v[0] = e.pose.t.x; v[1] = e.pose.t.y; v[2] = e.pose.t.z; v[3] = r; v[4] = p; v[5] = y;
  v[6] = e.s[0]; v[7] = e.s[1]; v[8] = e.s[2];
}
inline Ellipsoid ell_exp_update(const Ellipsoid &e, const double u[9]) {  // Ellipsoid.cpp:38-47
  Ellipsoid r; r.pose = se3_mul(e.pose, se3_exp(u));
  for (int i = 0; i < 3; i++) r.s[i] = e.s[i] + u[6 + i];
  return r;
}
inline Ellipsoid ell_transform_from(const Ellipsoid &e, const SE3 &Twc) {  // Ellipsoid.cpp:120-129
  Ellipsoid r = e; r.pose = se3_mul(Twc, e.pose); return r;
}
inline Ellipsoid ell_rotate(const Ellipsoid &e, double yaw) {  // rotate_ellipsoid, Ellipsoid.cpp:75-90
  Ellipsoid r = e;
  SE3 rot = se3_make(Q4{0, 0, std::sin(yaw * 0.5), std::cos(yaw * 0.5)}, V3{0, 0, 0});
  r.pose = se3_mul(e.pose, rot);
  const double eps = 1e-6;
  if ((std::fabs(yaw - M_PI / 2.0) < eps) || (std::fabs(yaw + M_PI / 2.0) < eps) || (std::fabs(yaw - 3 * M_PI / 2.0) < eps))
    std::swap(r.s[0], r.s[1]);
  return r;
}
inline void ell_log_error_9dof(const Ellipsoid &self, const Ellipsoid &newone, double res[9]) {  // Ellipsoid.cpp:63-71
  SE3 diff = se3_mul(se3_inverse(newone.pose), self.pose);
  se3_log(diff, res);
  for (int i = 0; i < 3; i++) res[6 + i] = self.s[i] - newone.s[i];
}
inline void ell_min_log_error_9dof(const Ellipsoid &self, const Ellipsoid &newone, double out[9]) {  // Ellipsoid.cpp:92-117
  double errs[4][9], norms[4];
  const double angles[4] = {-1, 0, 1, 2};
  for (int i = 0; i < 4; i++) {
    Ellipsoid rot = ell_rotate(newone, angles[i] * M_PI / 2.0);
    ell_log_error_9dof(self, rot, errs[i]);
    double s = 0; for (int k = 0; k < 9; k++) s += errs[i][k] * errs[i][k];
    norms[i] = std::sqrt(s);
  }
  int mi = 0;  // minCoeff: first minimum
  for (int i = 1; i < 4; i++) if (norms[i] < norms[mi]) mi = i;
  for (int k = 0; k < 9; k++) out[k] = errs[mi][k];
}
// generateProjectionMatrix (Ellipsoid.cpp:276-287): P = (K [I|0]) * hom(Tcw), 3x4
inline void projection_matrix(const SE3 &Tcw, const double K[9], double P[3][4]) {
  double KI[3][4];
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) KI[i][j] = K[i * 3 + j]; KI[i][3] = 0.0; }
  // Kalib * identity_lefttop is an exact copy of K with a zero column (products with 0/1)
  M4 T = se3_to_hom(Tcw);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) {
    double s = KI[i][0] * T.m[0][j];
    for (int k = 1; k < 4; k++) s += KI[i][k] * T.m[k][j];
    P[i][j] = s;
  }
}
inline M4 ell_generate_quadric(const Ellipsoid &e) {  // Ellipsoid.cpp:290-300
  double axis[4] = {1 / (e.s[0] * e.s[0]), 1 / (e.s[1] * e.s[1]), 1 / (e.s[2] * e.s[2]), -1};
  M4 Qc{};  // Q_c.inverse() of the diagonal matrix
  for (int i = 0; i < 4; i++) Qc.m[i][i] = 1.0 / axis[i];
  M4 T = se3_to_hom(e.pose);
  return mul(mul(T, Qc), transpose(T));
}
inline bool ell_project_ellipse(const Ellipsoid &e, const SE3 &Tcw, const double K[9], double out5[5]) {  // Ellipsoid.cpp:209-249
  M4 Qs = ell_generate_quadric(e);
  double P[3][4]; projection_matrix(Tcw, K, P);
  double PQ[3][4];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) { double s = P[i][0] * Qs.m[0][j]; for (int k = 1; k < 4; k++) s += P[i][k] * Qs.m[k][j]; PQ[i][j] = s; }
  M3 Cs;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = PQ[i][0] * P[j][0]; for (int k = 1; k < 4; k++) s += PQ[i][k] * P[j][k]; Cs.m[i][j] = s; }
  M3 C = inverse3(Cs);
  double c22 = C.m[2][2];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C.m[i][j] = C.m[i][j] / c22;
  // (the SelfAdjointEigenSolver at :217-223 and delta,k at :234-235 are unused by the output)
  double a = C.m[0][0], b = C.m[0][1] * 2, c = C.m[1][1], d = C.m[0][2] * 2, ee = C.m[2][1] * 2;
  double theta = 1 / 2.0 * std::atan2(b, (a - c));
  double x_c = (b * ee - 2 * c * d) / (4 * a * c - b * b);
  double y_c = (b * d - 2 * a * ee) / (4 * a * c - b * b);
  double a_2 = 2 * (a * x_c * x_c + c * y_c * y_c + b * x_c * y_c - 1) / (a + c + std::sqrt((a - c) * (a - c) + b * b));
  double b_2 = 2 * (a * x_c * x_c + c * y_c * y_c + b * x_c * y_c - 1) / (a + c - std::sqrt((a - c) * (a - c) + b * b));
  out5[0] = x_c; out5[1] = y_c; out5[2] = theta; out5[3] = std::sqrt(a_2); out5[4] = std::sqrt(b_2);
  return true;
}
inline void ellipse_bbox(const double el[5], double out[4]) {  // getBoundingBoxFromEllipse, Ellipsoid.cpp:252-273
  double a = el[3], b = el[4], theta = el[2], x = el[0], y = el[1];
  double cos_theta_2 = std::cos(theta) * std::cos(theta);
  double sin_theta_2 = 1 - cos_theta_2;
  double x_limit = std::sqrt(a * a * cos_theta_2 + b * b * sin_theta_2);
  double y_limit = std::sqrt(a * a * sin_theta_2 + b * b * cos_theta_2);
  out[0] = x - x_limit; out[1] = y - y_limit; out[2] = x + x_limit; out[3] = y + y_limit;
}
inline void ell_bbox_from_projection(const Ellipsoid &e, const SE3 &Tcw, const double K[9], double out[4]) {  // Ellipsoid.cpp:303-307
  double el[5]; ell_project_ellipse(e, Tcw, K, el); ellipse_bbox(el, out);
}

// ---- edges (src/core/BasicEllipsoidEdges.cpp) ----
enum { EDGE_PROJ = 0, EDGE_9DOF = 1, EDGE_GRAVITY = 2 };
inline int edge_dim(int type) { return type == EDGE_PROJ ? 4 : (type == EDGE_9DOF ? 9 : 1); }
// returns residual (dim values) for one edge
inline void edge_compute_error(int type, const SE3 &Tcw, const Ellipsoid &obj, const double *meas, const double K[9], double *err) {
  if (type == EDGE_PROJ) {  // EdgeSE3EllipsoidProj::computeError :102-112
    double proj[4]; ell_bbox_from_projection(obj, Tcw, K, proj);
    for (int i = 0; i < 4; i++) { err[i] = 0; if (meas[i] >= 5) err[i] = proj[i] - meas[i]; }
  } else if (type == EDGE_9DOF) {  // EdgeSE3Ellipsoid9DOF::computeError :68-77
    SE3 Twc = se3_inverse(Tcw);
    Ellipsoid m = ell_from_vector(meas);
    Ellipsoid esti_global = ell_transform_from(m, Twc);
    ell_min_log_error_9dof(obj, esti_global, err);
  } else {  // EdgeEllipsoidGravityPlanePrior::computeError :129-152
    M3 R = q_to_R(obj.pose.r);
    V3 zAxis = col(R, 2);
    V3 n{meas[0], meas[1], meas[2]};
    double cos_angle = dot(zAxis, n) / norm(zAxis) / norm(n);
    if (cos_angle > 1) cos_angle = cos_angle - 0.0001;
    else if (cos_angle < -1) cos_angle = cos_angle + 0.0001;
    err[0] = std::acos(cos_angle) - 0;
  }
}

struct EdgeBatch {
  int n_edges; const int *type; const int *frame; const int *object; const double *meas; const double *info;
  int n_frames; const double *Tcw; int n_objects; const double *K;
};

// g2o numeric Jacobian w.r.t. the (free) ellipsoid vertex: base_binary_edge.hpp:147-197 / base_unary_edge.hpp:94-117
inline void edge_numeric_jacobian(int type, const SE3 &Tcw, const Ellipsoid &obj, const double *meas, const double K[9], double *J /*dim x 9 row-major*/) {
  const double delta = 1e-9;
  const double scalar = 1.0 / (2 * delta);
  int D = edge_dim(type);
  double add[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int d = 0; d < 9; d++) {
    double e1[9], e2[9];
    add[d] = delta;
    edge_compute_error(type, Tcw, ell_exp_update(obj, add), meas, K, e1);
    add[d] = -delta;
    edge_compute_error(type, Tcw, ell_exp_update(obj, add), meas, K, e2);
    add[d] = 0.0;
    for (int r = 0; r < D; r++) J[r * 9 + d] = scalar * (e1[r] - e2[r]);
  }
}

// ---- generic g2o Levenberg driver (optimization_algorithm_levenberg.cpp:61-164 + sparse_optimizer.cpp:354-419) ----
struct LmProblem {
  int dim;                                              // total free dimension
  std::function<double()> compute_errors;               // computeActiveErrors + activeRobustChi2 (no robust kernels here)
  std::function<void(std::vector<double> &H, std::vector<double> &b)> build_system;  // dense H (dim x dim), b
  std::function<void()> push, pop, discard_top;
  std::function<void(const double *x)> update;          // SparseOptimizer::update: oplus on every vertex
};
struct LmStats { int iterations = 0, trials_total = 0, terminated = 0; double chi2_initial = 0, chi2_final = 0, lambda_final = 0; };

inline LmStats lm_optimize(LmProblem &P, int iterations) {
  LmStats st;
  const double tau = 1e-5, goodUpper = 2. / 3., goodLower = 1. / 3.;
  const int maxTrials = 10;
  double lambda = -1, ni = 2;
  int nBad = 0;
  std::vector<double> H, b, x(P.dim), Hl;
  bool ok = true;
  for (int it = 0; it < iterations && ok; it++) {
    // ---- solve(it)
    double currentChi = P.compute_errors();
    if (it == 0) st.chi2_initial = currentChi;
    double tempChi = currentChi;
    double iniChi = currentChi;
    P.build_system(H, b);
    if (it == 0) {  // computeLambdaInit :166-180
      double maxDiagonal = 0;
      for (int j = 0; j < P.dim; j++) maxDiagonal = std::max(std::fabs(H[(size_t)j * P.dim + j]), maxDiagonal);
      lambda = tau * maxDiagonal; ni = 2; nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      P.push();
      Hl = H;
      for (int j = 0; j < P.dim; j++) Hl[(size_t)j * P.dim + j] += lambda;  // setLambda
      bool ok2 = ldlt_solve(P.dim, Hl, b.data(), x.data());
      if (!ok2) { /* g2o leaves _x untouched from the previous solve */ }
      P.update(x.data());
      tempChi = P.compute_errors();
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = (currentChi - tempChi);
      double scale = 0;  // computeScale :182-189
      for (int j = 0; j < P.dim; j++) scale += x[j] * (lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, goodUpper);
        double scaleFactor = std::max(goodLower, alpha);
        lambda *= scaleFactor; ni = 2; currentChi = tempChi;
        P.discard_top();
      } else {
        lambda *= ni; ni *= 2;
        P.pop();
      }
      qmax++;
    } while (rho < 0 && qmax < maxTrials);
    st.iterations++; st.trials_total += qmax;
    bool terminate = false;
    if (qmax == maxTrials || rho == 0) terminate = true;
    else {
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
      if (nBad >= 3) terminate = true;
    }
    st.chi2_final = currentChi;
    if (terminate) { st.terminated = 1; ok = false; }
  }
  st.lambda_final = lambda;
  return st;
}

// Object BA in mapping mode (Optimizer.cpp:83-317): all SE3 fixed, one free 9-DoF vertex per ellipsoid.
inline LmStats optimize_objects(const EdgeBatch &g, int iters, double *ellipsoids10) {
  int No = g.n_objects;
  std::vector<Ellipsoid> est(No), backup;
  for (int i = 0; i < No; i++) est[i] = ell_from_vector(ellipsoids10 + 10 * i);
  std::vector<SE3> Tcw(g.n_frames);
  for (int f = 0; f < g.n_frames; f++) Tcw[f] = se3_from_vector(g.Tcw + 7 * f);
  std::vector<double> err((size_t)g.n_edges * 9, 0.0);
  std::vector<std::vector<Ellipsoid>> stack;
  LmProblem P;
  P.dim = 9 * No;
  P.compute_errors = [&]() {
    double chi = 0;
    for (int k = 0; k < g.n_edges; k++) {
      int t = g.type[k]; int D = edge_dim(t);
      const SE3 &T = (t == EDGE_GRAVITY) ? Tcw[0] : Tcw[g.frame[k]];
      edge_compute_error(t, T, est[g.object[k]], g.meas + 10 * k, g.K, &err[(size_t)k * 9]);
      double c = 0;  // chi2 = e^T (s I) e  (base_edge.h:58-61)
      for (int r = 0; r < D; r++) c += err[(size_t)k * 9 + r] * (g.info[k] * err[(size_t)k * 9 + r]);
      chi += c;
    }
    return chi;
  };
  P.build_system = [&](std::vector<double> &H, std::vector<double> &b) {
    int n = 9 * No;
    H.assign((size_t)n * n, 0.0); b.assign(n, 0.0);
    double J[81];
    for (int k = 0; k < g.n_edges; k++) {
      int t = g.type[k]; int D = edge_dim(t); int o = g.object[k];
      const SE3 &T = (t == EDGE_GRAVITY) ? Tcw[0] : Tcw[g.frame[k]];
      edge_numeric_jacobian(t, T, est[o], g.meas + 10 * k, g.K, J);
      const double *e = &err[(size_t)k * 9];
      double w = g.info[k];
      // binary: omega_r = -omega*e; b += B^T omega_r ; A += B^T omega B   (base_binary_edge.hpp:73-90)
      // unary:  b -= A^T omega e ; A += A^T omega A                      (base_unary_edge.hpp:65-66)
      for (int i = 0; i < 9; i++) {
        double bi = 0;
        for (int r = 0; r < D; r++) {
          // unary: (A^T omega) e ; binary: B^T (-(omega e))
          double pr = (t == EDGE_GRAVITY) ? ((J[r * 9 + i] * w) * e[r]) : (J[r * 9 + i] * (-(w * e[r])));
          if (r == 0) bi = pr; else bi += pr;
        }
        if (t == EDGE_GRAVITY) b[9 * o + i] -= bi; else b[9 * o + i] += bi;
        for (int j = 0; j < 9; j++) {
          double h = 0;
          for (int r = 0; r < D; r++) { double pr = (J[r * 9 + i] * w) * J[r * 9 + j]; if (r == 0) h = pr; else h += pr; }
          H[(size_t)(9 * o + i) * n + 9 * o + j] += h;
        }
      }
    }
  };
  P.push = [&]() { stack.push_back(est); };
  P.pop = [&]() { est = stack.back(); stack.pop_back(); };
  P.discard_top = [&]() { stack.pop_back(); };
  P.update = [&](const double *x) { for (int i = 0; i < No; i++) est[i] = ell_exp_update(est[i], x + 9 * i); };
  LmStats st = lm_optimize(P, iters);
  for (int i = 0; i < No; i++) ell_to_vector(est[i], ellipsoids10 + 10 * i);
  return st;
}

// ---- SLAM mode (Optimizer.cpp:126-159 with bSLAM_mode = true): camera poses are free 6-DoF vertices (frame 0 fixed),
// consecutive frames are tied by EdgeSE3Expmap odometry edges (types_six_dof_expmap.h:83-99), every Jacobian is numeric
// (base_binary_edge.hpp:131-205), the system is solved densely (LinearSolverDense, no marginalisation).
// Hessian order = vertex id order (sparse_optimizer.cpp sortVectorContainers): poses 1..F-1 then ellipsoids.
// Edge order = insertion order: F-1 odometry edges first (added inside the pose loop), then the ellipsoid edges of `g`.
inline SE3 pose_oplus(const SE3 &T, const double u[6]) { return se3_mul(se3_exp(u), T); }  // VertexSE3Expmap::oplusImpl, types_six_dof_expmap.h:73-76
inline void odom_error(const SE3 &meas, const SE3 &v1, const SE3 &v2, double err[6]) {       // EdgeSE3Expmap::computeError :89-98
  SE3 e = se3_mul(se3_mul(meas, v1), se3_inverse(v2));
  se3_log(e, err);
}
inline LmStats optimize_slam(const EdgeBatch &g, const double *odom_meas7, double odom_info, int iters, double *ellipsoids10, double *Tcw_out7) {
  const int No = g.n_objects, F = g.n_frames, NP = F - 1;
  std::vector<Ellipsoid> est(No);
  for (int i = 0; i < No; i++) est[i] = ell_from_vector(ellipsoids10 + 10 * i);
  std::vector<SE3> Tcw(F), odo(NP > 0 ? NP : 0);
  for (int f = 0; f < F; f++) Tcw[f] = se3_from_vector(g.Tcw + 7 * f);
  for (int k = 0; k < NP; k++)
    odo[k] = odom_meas7 ? se3_from_vector(odom_meas7 + 7 * k) : se3_mul(Tcw[k + 1], se3_inverse(Tcw[k]));  // Optimizer.cpp:141-143
  const int n = 6 * NP + 9 * No;
  auto pidx = [&](int f) { return 6 * (f - 1); };          // f >= 1
  auto oidx = [&](int o) { return 6 * NP + 9 * o; };
  std::vector<double> err((size_t)g.n_edges * 9, 0.0), oerr((size_t)(NP > 0 ? NP : 0) * 6, 0.0);
  struct Snap { std::vector<Ellipsoid> e; std::vector<SE3> T; };
  std::vector<Snap> stack;
  LmProblem P;
  P.dim = n;
  P.compute_errors = [&]() {
    double chi = 0;
    for (int k = 0; k < NP; k++) {
      odom_error(odo[k], Tcw[k], Tcw[k + 1], &oerr[(size_t)k * 6]);
      double c = 0;
      for (int r = 0; r < 6; r++) c += oerr[(size_t)k * 6 + r] * (odom_info * oerr[(size_t)k * 6 + r]);
      chi += c;
    }
    for (int k = 0; k < g.n_edges; k++) {
      int t = g.type[k]; int D = edge_dim(t);
      const SE3 &T = (t == EDGE_GRAVITY) ? Tcw[0] : Tcw[g.frame[k]];
      edge_compute_error(t, T, est[g.object[k]], g.meas + 10 * k, g.K, &err[(size_t)k * 9]);
      double c = 0;
      for (int r = 0; r < D; r++) c += err[(size_t)k * 9 + r] * (g.info[k] * err[(size_t)k * 9 + r]);
      chi += c;
    }
    return chi;
  };
  // H(ia.., ib..) += A^T (w I) B with the reference's association (A(r,i)*w)*B(r,j)
  auto add_block = [&](std::vector<double> &H, int ia, int da, const double *A, int ib, int db, const double *B, int D, double w, int lda, int ldb) {
    for (int i = 0; i < da; i++)
      for (int j = 0; j < db; j++) {
        double h = 0;
        for (int r = 0; r < D; r++) { double pr = (A[r * lda + i] * w) * B[r * ldb + j]; if (r == 0) h = pr; else h += pr; }
        H[(size_t)(ia + i) * n + ib + j] += h;
        if (ia != ib) H[(size_t)(ib + j) * n + ia + i] += h;   // off-diagonal block: the lower copy is the transpose (linear_solver_dense.h:88-91)
      }
  };
  auto add_b = [&](std::vector<double> &b, int ia, int da, const double *A, const double *e, int D, double w, int lda) {
    for (int i = 0; i < da; i++) {
      double bi = 0;
      for (int r = 0; r < D; r++) { double pr = A[r * lda + i] * (-(w * e[r])); if (r == 0) bi = pr; else bi += pr; }
      b[ia + i] += bi;
    }
  };
  P.build_system = [&](std::vector<double> &H, std::vector<double> &b) {
    H.assign((size_t)n * n, 0.0); b.assign(n, 0.0);
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    for (int k = 0; k < NP; k++) {  // odometry edges: vertex i = frame k, vertex j = frame k+1
      double Ji[36], Jj[36], e1[6], e2[6], add[6] = {0, 0, 0, 0, 0, 0};
      bool iFree = k >= 1;
      if (iFree) for (int d = 0; d < 6; d++) {
        add[d] = delta; odom_error(odo[k], pose_oplus(Tcw[k], add), Tcw[k + 1], e1);
        add[d] = -delta; odom_error(odo[k], pose_oplus(Tcw[k], add), Tcw[k + 1], e2);
        add[d] = 0.0;
        for (int r = 0; r < 6; r++) Ji[r * 6 + d] = scalar * (e1[r] - e2[r]);
      }
      for (int d = 0; d < 6; d++) {
        add[d] = delta; odom_error(odo[k], Tcw[k], pose_oplus(Tcw[k + 1], add), e1);
        add[d] = -delta; odom_error(odo[k], Tcw[k], pose_oplus(Tcw[k + 1], add), e2);
        add[d] = 0.0;
        for (int r = 0; r < 6; r++) Jj[r * 6 + d] = scalar * (e1[r] - e2[r]);
      }
      const double *e = &oerr[(size_t)k * 6];
      if (iFree) {
        add_b(b, pidx(k), 6, Ji, e, 6, odom_info, 6);
        add_block(H, pidx(k), 6, Ji, pidx(k), 6, Ji, 6, odom_info, 6, 6);
        add_block(H, pidx(k), 6, Ji, pidx(k + 1), 6, Jj, 6, odom_info, 6, 6);
      }
      add_b(b, pidx(k + 1), 6, Jj, e, 6, odom_info, 6);
      add_block(H, pidx(k + 1), 6, Jj, pidx(k + 1), 6, Jj, 6, odom_info, 6, 6);
    }
    double Jo[81], Jp[54];
    for (int k = 0; k < g.n_edges; k++) {
      int t = g.type[k]; int D = edge_dim(t); int o = g.object[k];
      int f = (t == EDGE_GRAVITY) ? 0 : g.frame[k];
      const SE3 &T = Tcw[f];
      const double *e = &err[(size_t)k * 9];
      double w = g.info[k];
      edge_numeric_jacobian(t, T, est[o], g.meas + 10 * k, g.K, Jo);
      if (t == EDGE_GRAVITY) {  // unary (base_unary_edge.hpp:65-66)
        for (int i = 0; i < 9; i++) {
          double bi = 0;
          for (int r = 0; r < D; r++) { double pr = (Jo[r * 9 + i] * w) * e[r]; if (r == 0) bi = pr; else bi += pr; }
          b[oidx(o) + i] -= bi;
        }
        add_block(H, oidx(o), 9, Jo, oidx(o), 9, Jo, D, w, 9, 9);
        continue;
      }
      bool pFree = f >= 1;
      if (pFree) {
        double add[6] = {0, 0, 0, 0, 0, 0}, e1[9], e2[9];
        for (int d = 0; d < 6; d++) {
          add[d] = delta; edge_compute_error(t, pose_oplus(T, add), est[o], g.meas + 10 * k, g.K, e1);
          add[d] = -delta; edge_compute_error(t, pose_oplus(T, add), est[o], g.meas + 10 * k, g.K, e2);
          add[d] = 0.0;
          for (int r = 0; r < D; r++) Jp[r * 6 + d] = scalar * (e1[r] - e2[r]);
        }
        add_b(b, pidx(f), 6, Jp, e, D, w, 6);
        add_block(H, pidx(f), 6, Jp, pidx(f), 6, Jp, D, w, 6, 6);
        add_block(H, pidx(f), 6, Jp, oidx(o), 9, Jo, D, w, 6, 9);
      }
      add_b(b, oidx(o), 9, Jo, e, D, w, 9);
      add_block(H, oidx(o), 9, Jo, oidx(o), 9, Jo, D, w, 9, 9);
    }
  };
  P.push = [&]() { stack.push_back(Snap{est, Tcw}); };
  P.pop = [&]() { est = stack.back().e; Tcw = stack.back().T; stack.pop_back(); };
  P.discard_top = [&]() { stack.pop_back(); };
  P.update = [&](const double *x) {
    for (int f = 1; f < F; f++) Tcw[f] = pose_oplus(Tcw[f], x + pidx(f));
    for (int i = 0; i < No; i++) est[i] = ell_exp_update(est[i], x + oidx(i));
  };
  LmStats st = lm_optimize(P, iters);
  for (int i = 0; i < No; i++) ell_to_vector(est[i], ellipsoids10 + 10 * i);
  if (Tcw_out7) for (int f = 0; f < F; f++) se3_to_vector(Tcw[f], Tcw_out7 + 7 * f);
  return st;
}

}  // namespace orc

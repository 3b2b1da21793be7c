// ORACLE -- TEST INFRASTRUCTURE ONLY.
// CPU restatement (one thread, fp64 / fp32 exactly where the reference uses them) of the reference's per-frame
// quadric-estimation hot path, exported as `orc_*` C symbols for ctypes.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the product (libeslam_b200.so) never
// links or calls it.  PARITY UNPINNED: the reference itself cannot be built here (needs Eigen/PCL/OpenCV/Pangolin,
// none installed, no network) and ships no tests or golden vectors; the oracle is pinned instead against the
// analytic known-answer tests of SURVEY.md Appendix C (tests/test_oracle_kat.py).
#include "orc_math.h"
#include "orc_graph.h"
#include "orc_extract.h"
#include "orc_plane.h"
#include <chrono>

using namespace orc;

namespace {
// Initializer::initializeQuadric (src/core/Initializer.cpp:24-56,58-248)
bool init_quadric(int n_obs, const double *Twc7, const double *det5, const double K[9], int rows, int cols, double ell10[10]) {
  std::vector<double> planes;  // 4 x N column list
  for (int i = 0; i < n_obs; i++) {
    const double *det = det5 + 5 * i;
    if (det[0] < 1 && det[1] < 1 && det[2] < 1 && det[3] < 1) continue;   // :70-71
    SE3 wc = se3_from_vector_ctor(Twc7 + 7 * i);
    double P[3][4]; projection_matrix(se3_inverse(wc), K, P);
    double x1 = det[0], y1 = det[1], x2 = det[2], y2 = det[3];
    double lines[4][3] = {{1, 0, -x1}, {0, 1, -y1}, {1, 0, -x2}, {0, 1, -y2}};
    bool keep[4] = {x1 > 0 && x1 < cols - 1, y1 > 0 && y1 < rows - 1, x2 > 0 && x2 < cols - 1, y2 > 0 && y2 < rows - 1};  // :122-141
    for (int l = 0; l < 4; l++) if (keep[l]) {
      for (int r = 0; r < 4; r++) { double s = P[0][r] * lines[l][0]; s += P[1][r] * lines[l][1]; s += P[2][r] * lines[l][2]; planes.push_back(s); }
    }
  }
  int N = (int)planes.size() / 4;
  if (N < 9) return false;
  std::vector<double> A((size_t)N * 10);
  for (int i = 0; i < N; i++) {  // getVectorFromPlanesHomo :147-164
    const double *p = &planes[4 * i];
    double v[10] = {p[0] * p[0], 2 * p[0] * p[1], 2 * p[0] * p[2], 2 * p[0] * p[3], p[1] * p[1], 2 * p[1] * p[2], 2 * p[1] * p[3], p[2] * p[2], 2 * p[2] * p[3], p[3] * p[3]};
    for (int k = 0; k < 10; k++) A[(size_t)i * 10 + k] = v[k];
  }
  std::vector<double> q; svd_null_vector(N, 10, A, q);   // getQStarFromVectors :166-184
  M4 QStar = {{{q[0], q[1], q[2], q[3]}, {q[1], q[4], q[5], q[6]}, {q[2], q[5], q[7], q[8]}, {q[3], q[6], q[8], q[9]}}};
  // getEllipsoidFromQStar :186-248
  M4 Qi; double detQs; if (!inverse4(QStar, Qi, &detQs)) return false;
  double cb = std::cbrt(detQs);
  M4 Q; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Q.m[i][j] = Qi.m[i][j] * cb;
  auto eig4 = [](const M4 &M, double ev[4]) { std::vector<double> a(16), e; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) a[i * 4 + j] = 0.5 * (M.m[i][j] + M.m[j][i]); jacobi_eigen_sym(4, a, e); for (int i = 0; i < 4; i++) ev[i] = e[i]; };
  double eigens[4]; eig4(Q, eigens);
  int num_pos = int(eigens[0] > 0) + int(eigens[1] > 0) + int(eigens[2] > 0) + int(eigens[3] > 0);
  int num_neg = int(eigens[0] < 0) + int(eigens[1] < 0) + int(eigens[2] < 0) + int(eigens[3] < 0);
  if (!(num_pos == 3 && num_neg == 1) && !(num_pos == 1 && num_neg == 3)) return false;
  if (eigens[3] > 0) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Q.m[i][j] = -Q.m[i][j]; eig4(Q, eigens); }
  double lambda[3] = {1.0 / eigens[0], 1.0 / eigens[1], 1.0 / eigens[2]};
  M3 Q33; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q33.m[i][j] = Q.m[i][j];
  M4 tmp; double detQ; inverse4(Q, tmp, &detQ);
  double k = detQ / det3(Q33);
  double s[3]; for (int i = 0; i < 3; i++) s[i] = std::sqrt(std::fabs(-k * lambda[i]));
  double t[4] = {QStar.m[0][3], QStar.m[1][3], QStar.m[2][3], QStar.m[3][3]};
  for (int i = 0; i < 3; i++) t[i] = t[i] / t[3];
  double ev3[3]; M3 rot; self_adjoint_eigen3(Q33, ev3, rot);
  double r, p, y; rot_to_euler_zyx(rot, r, p, y);
  double mv[9] = {t[0], t[1], t[2], r, p, y, s[0], s[1], s[2]};
  Ellipsoid e = ell_from_minimal(mv);
  ell_to_vector(e, ell10);
  return true;
}
}  // namespace

extern "C" {

// ---- math KATs
void orc_se3_exp(const double v6[6], double out7[7]) { se3_to_vector(se3_exp(v6), out7); }
void orc_se3_log(const double T7[7], double out6[6]) { se3_log(se3_from_vector(T7), out6); }
void orc_se3_mul(const double a7[7], const double b7[7], double out7[7]) { se3_to_vector(se3_mul(se3_from_vector(a7), se3_from_vector(b7)), out7); }
void orc_se3_inverse(const double a7[7], double out7[7]) { se3_to_vector(se3_inverse(se3_from_vector(a7)), out7); }
void orc_ellipsoid_exp_update(const double e10[10], const double u9[9], double out10[10]) { ell_to_vector(ell_exp_update(ell_from_vector(e10), u9), out10); }
void orc_ellipsoid_rotate(const double e10[10], double yaw, double out10[10]) { ell_to_vector(ell_rotate(ell_from_vector(e10), yaw), out10); }
void orc_ellipsoid_transform_from(const double e10[10], const double T7[7], double out10[10]) { ell_to_vector(ell_transform_from(ell_from_vector(e10), se3_from_vector(T7)), out10); }
void orc_ellipsoid_from_minimal(const double v9[9], double out10[10]) { ell_to_vector(ell_from_minimal(v9), out10); }
void orc_ellipsoid_to_minimal(const double e10[10], double out9[9]) { ell_to_minimal(ell_from_vector(e10), out9); }
void orc_min_log_error_9dof(const double self10[10], const double new10[10], double out9[9]) { ell_min_log_error_9dof(ell_from_vector(self10), ell_from_vector(new10), out9); }
void orc_ellipsoid_bbox(const double e10[10], const double Tcw7[7], const double K[9], double out4[4]) { ell_bbox_from_projection(ell_from_vector(e10), se3_from_vector(Tcw7), K, out4); }
void orc_eigen3(const double A9[9], double eval[3], double evec9[9]) {
  M3 A, V; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A.m[i][j] = A9[i * 3 + j];
  self_adjoint_eigen3(A, eval, V);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) evec9[i * 3 + j] = V.m[i][j];
}
int orc_ldlt_solve(int n, const double *A, const double *b, double *x) { std::vector<double> m(A, A + (size_t)n * n); return ldlt_solve(n, m, b, x) ? 1 : 0; }
void orc_plane_to_world(const double plane_cam[4], const double Twc7[7], double out[4]) { for (int i = 0; i < 4; i++) out[i] = plane_cam[i]; SE3 T = se3_from_vector(Twc7); plane_transform(out, T); }
void orc_plane_oplus(double p4[4], double *dual_dis, const double v3[3], int dual) {
  Plane pl; for (int i = 0; i < 4; i++) pl.p[i] = p4[i]; pl.dual_dis = *dual_dis;
  if (dual) plane_oplus_dual(pl, v3); else plane_oplus(pl, v3);
  for (int i = 0; i < 4; i++) p4[i] = pl.p[i];
  *dual_dis = pl.dual_dis;
}
void orc_pose_inverse(const double T7[7], double out7[7]) { se3_to_vector(se3_inverse(se3_from_vector(T7)), out7); }

// ---- edges / BA
void orc_edge_error(int type, const double Tcw7[7], const double e10[10], const double *meas, const double K[9], double *err) {
  edge_compute_error(type, se3_from_vector(Tcw7), ell_from_vector(e10), meas, K, err);
}
// residuals [n_edges*9], jacobians [n_edges*81] (dim x 9 row-major), chi2 [n_edges]
void orc_edges_eval(const EslamEdgeBatch *g, const double *ellipsoids10, double *residuals, double *jacobians, double *chi2) {
  for (int k = 0; k < g->n_edges; k++) {
    int t = g->type[k]; int D = edge_dim(t);
    SE3 T = se3_from_vector(g->Tcw + 7 * (t == EDGE_GRAVITY ? 0 : g->frame[k]));
    Ellipsoid e = ell_from_vector(ellipsoids10 + 10 * g->object[k]);
    double r[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, J[81];
    for (int i = 0; i < 81; i++) J[i] = 0;
    edge_compute_error(t, T, e, g->meas + 10 * k, g->K, r);
    if (jacobians) edge_numeric_jacobian(t, T, e, g->meas + 10 * k, g->K, J);
    for (int i = 0; i < 9; i++) residuals[(size_t)k * 9 + i] = i < D ? r[i] : 0.0;
    if (jacobians) for (int i = 0; i < 81; i++) jacobians[(size_t)k * 81 + i] = (i / 9) < D ? J[i] : 0.0;
    if (chi2) { double c = 0; for (int i = 0; i < D; i++) c += r[i] * (g->info[k] * r[i]); chi2[k] = c; }
  }
}
int orc_optimize_objects(const EslamEdgeBatch *g, int iters, double *ellipsoids10, EslamLmStats *stats) {
  EdgeBatch b{g->n_edges, g->type, g->frame, g->object, g->meas, g->info, g->n_frames, g->Tcw, g->n_objects, g->K};
  LmStats st = optimize_objects(b, iters, ellipsoids10);
  if (stats) { stats->iterations = st.iterations; stats->trials_total = st.trials_total; stats->terminated = st.terminated; stats->chi2_initial = st.chi2_initial; stats->chi2_final = st.chi2_final; stats->lambda_final = st.lambda_final; }
  return 0;
}

int orc_optimize_slam(const EslamEdgeBatch *g, const double *odom_meas7, double odom_info, int iters, double *ellipsoids10, double *Tcw_out7, EslamLmStats *stats) {
  EdgeBatch b{g->n_edges, g->type, g->frame, g->object, g->meas, g->info, g->n_frames, g->Tcw, g->n_objects, g->K};
  LmStats st = optimize_slam(b, odom_meas7, odom_info, iters, ellipsoids10, Tcw_out7);
  if (stats) { stats->iterations = st.iterations; stats->trials_total = st.trials_total; stats->terminated = st.terminated; stats->chi2_initial = st.chi2_initial; stats->chi2_final = st.chi2_final; stats->lambda_final = st.lambda_final; }
  return 0;
}
// odometry residual (for tests): err6 = log(meas * v1 * v2^-1)
void orc_odom_error(const double meas7[7], const double v1[7], const double v2[7], double err6[6]) {
  odom_error(se3_from_vector(meas7), se3_from_vector(v1), se3_from_vector(v2), err6);
}

// ---- extraction
int orc_calibrate_measurement(const EslamParams *p, const EslamCamera *c, const double m[4]) {  // dataprocess_utils.cpp:150-190
  int x_length = m[2] - m[0], y_length = m[3] - m[1];
  if (x_length < p->length_limit_pixels || y_length < p->length_limit_pixels) return 1;
  int correct = 0; int b = p->border_pixels;
  if (m[0] > b && m[0] < c->cols - 1 - b) correct++;
  if (m[2] > b && m[2] < c->cols - 1 - b) correct++;
  if (m[1] > b && m[1] < c->rows - 1 - b) correct++;
  if (m[3] > b && m[3] < c->rows - 1 - b) correct++;
  return correct != 4;
}
// returns total symmetry cost evaluations (for reporting); out[n_obj]
long long orc_estimate_ellipsoids(const EslamParams *prm, const EslamCamera *cam, const double plane_world[4], int unstable_voxel_sort,
                                  int n_frames, const uint16_t *depth, size_t pitch, int n_obj, const int *frame_of_obj, const double *bbox4,
                                  const int *label, const double *Twc7, EslamEllipsoidOut *out, EslamStageDump *dump) {
  ExtractorState st; st.plane_set = true; for (int i = 0; i < 4; i++) st.plane[i] = plane_world[i]; st.unstable_voxel_sort = unstable_voxel_sort != 0;
  long long evals = 0;
  for (int i = 0; i < n_obj; i++) {
    int f = frame_of_obj[i];
    if (f < 0 || f >= n_frames) { std::memset(&out[i], 0, sizeof(out[i])); out[i].status = ESLAM_OBJ_BBOX_OUTSIDE; continue; }
    EslamStageDump *d = dump ? &dump[i] : nullptr;
    if (d) { for (int s = 0; s < 9; s++) { d->seed_prob[s] = d->seed_init_err[s] = d->seed_final_err[s] = d->seed_dual_dis[s] = 0; for (int k = 0; k < 4; k++) d->seed_plane[s][k] = 0; } }
    estimate_local_ellipsoid(st, *prm, *cam, depth + (size_t)f * cam->rows * pitch, pitch, bbox4 + 4 * i, label[i], Twc7 + 7 * f, out[i], d, &evals);
  }
  return evals;
}
void orc_last_seed_evals(long long *out9) { for (int s = 0; s < 9; s++) out9[s] = g_last_seed_evals[s]; }
// stage helpers for stage-isolated parity tests
int orc_voxel_grid(const float *xyz, int n, double leaf, int unstable, float *out_xyz, int cap) {
  std::vector<P3f> in(n), o; for (int i = 0; i < n; i++) in[i] = P3f{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  voxel_grid(in, leaf, o, unstable != 0);
  int m = std::min<int>(cap, (int)o.size());
  for (int i = 0; i < m; i++) { out_xyz[3 * i] = o[i].x; out_xyz[3 * i + 1] = o[i].y; out_xyz[3 * i + 2] = o[i].z; }
  return (int)o.size();
}
int orc_euclidean_clusters(const float *xyz, int n, double tol, int min_size, int *label_out) {
  std::vector<P3f> in(n); for (int i = 0; i < n; i++) in[i] = P3f{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  std::vector<std::vector<int>> cl; euclidean_clusters(in, tol, min_size, n, cl);
  for (int i = 0; i < n; i++) label_out[i] = -1;
  for (size_t c = 0; c < cl.size(); c++) for (int i : cl[c]) label_out[i] = (int)c;
  return (int)cl.size();
}
// symmetry cost for an explicit plane (KAT 7): returns edge error (= -mean ln P)
double orc_symmetry_error(const EslamParams *prm, const EslamCamera *cam, const uint16_t *depth, size_t pitch, const double bbox[4], const double T_oc7[7],
                          const double *cloud_xyz, int n, const double plane4[4], double dual_dis, int symType) {
  std::vector<P3d> cl(n); for (int i = 0; i < n; i++) cl[i] = P3d{cloud_xyz[3 * i], cloud_xyz[3 * i + 1], cloud_xyz[3 * i + 2]};
  SymContext ctx; ctx.depth = depth; ctx.pitch = pitch; ctx.cam = *cam; for (int i = 0; i < 4; i++) ctx.bbox[i] = bbox[i];
  ctx.T_oc = se3_from_vector(T_oc7);
  double K[9] = {cam->fx, 0, cam->cx, 0, cam->fy, cam->cy, 0, 0, 1}; for (int i = 0; i < 9; i++) ctx.K[i] = K[i];
  ctx.cloud = &cl; ctx.sigma = prm->symmetry_sigma; ctx.cloud_f.resize(n);
  for (int i = 0; i < n; i++) ctx.cloud_f[i] = P3f{(float)cl[i].x, (float)cl[i].y, (float)cl[i].z};
  Plane pl; for (int i = 0; i < 4; i++) pl.p[i] = plane4[i]; pl.dual_dis = dual_dis;
  return symmetry_edge_error(ctx, pl, symType);
}

// ---- ground plane
int orc_extract_ground_plane(const EslamParams *prm, const EslamCamera *cam, int n_frames, const uint16_t *depth, size_t pitch, double *plane_cam, int *found, int *n_inliers) {
  for (int f = 0; f < n_frames; f++) {
    int ni = 0; double pl[4] = {0, 0, 0, 0};
    bool ok = extract_ground_plane(*prm, *cam, depth + (size_t)f * cam->rows * pitch, pitch, pl, &ni);
    for (int k = 0; k < 4; k++) plane_cam[4 * f + k] = ok ? pl[k] : 0.0;
    found[f] = ok ? 1 : 0; if (n_inliers) n_inliers[f] = ok ? ni : 0;
  }
  return 0;
}
// debug: normals [rows*cols*3] float, dist_map [rows*cols] float, labels [rows*cols] int; regions: up to cap x (coef4, size, first_pixel)
int orc_extract_planes_debug(const EslamParams *prm, const EslamCamera *cam, const uint16_t *depth, size_t pitch, float *normals, float *dist_map, int *labels,
                             float *region_coef, int *region_size, int *region_first, int cap) {
  PlaneDebug dbg; extract_planes(*prm, *cam, depth, pitch, dbg);
  size_t NP = (size_t)cam->rows * cam->cols;
  if (normals) std::memcpy(normals, dbg.normals.data(), NP * 3 * sizeof(float));
  if (dist_map) std::memcpy(dist_map, dbg.dist_map.data(), NP * sizeof(float));
  if (labels) std::memcpy(labels, dbg.labels.data(), NP * sizeof(int));
  int n = std::min<int>(cap, (int)dbg.regions.size());
  for (int i = 0; i < n; i++) { for (int k = 0; k < 4; k++) region_coef[4 * i + k] = dbg.regions[i].coef[k]; region_size[i] = dbg.regions[i].size; region_first[i] = dbg.regions[i].first_pixel; }
  return (int)dbg.regions.size();
}

// ---- initializer
int orc_init_quadric(int n_obs, const double *Twc7, const double *det5, const double K[9], int rows, int cols, double ell10[10], int *ok) {
  for (int i = 0; i < 10; i++) ell10[i] = 0;
  bool r = init_quadric(n_obs, Twc7, det5, K, rows, cols, ell10);
  if (ok) *ok = r ? 1 : 0;
  return 0;
}

double orc_now_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // extern "C"

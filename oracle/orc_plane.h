// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
// orc_plane.h: ground-plane extraction, src/plane/PlaneExtractor.cpp:22-183.
// The arithmetic lives in PCL (absent from the reference tree, version unpinned, README.md:36): this file restates
// the published upstream algorithms (PCL 1.8-1.10) -- PARITY UNPINNED:
//   pcl::IntegralImageNormalEstimation (AVERAGE_3D_GRADIENT, maxDepthChangeFactor 0.05, smoothing 10, border IGNORE)
//   pcl::OrganizedMultiPlaneSegmentation::segmentAndRefine (PlaneCoefficientComparator depth-dependent distance,
//     OrganizedConnectedComponentSegmentation, computeMeanAndCovarianceMatrix<float>, pcl::eigen33,
//     PlaneRefinementComparator two-pass refinement)
#pragma once
#include "orc_extract.h"

namespace orc {

struct PlaneRegion { float coef[4]; int size; int first_pixel; };
struct PlaneDebug {
  std::vector<float> normals;      // [rows*cols*3], NaN where undefined
  std::vector<float> dist_map;     // chamfer distance map
  std::vector<int> labels;         // after refinement
  std::vector<PlaneRegion> regions;  // kept planes (>= Plane.MinSize), label order
};

inline void computeRoots2f(float b, float c, float roots[3]) {
  roots[0] = 0.0f;
  float d = (float)(b * b - 4.0 * c);
  if (d < 0.0) d = 0.0;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
inline void computeRoots3f(const float m[3][3], float roots[3]) {  // pcl::computeRoots (common/eigen.hpp)
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) computeRoots2f(c2, c1, roots);
  else {
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = std::sqrt(-a_over_3);
    float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
    float cos_theta = std::cos(theta), sin_theta = std::sin(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) { std::swap(roots[1], roots[2]); if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]); }
    if (roots[0] <= 0) computeRoots2f(c2, c1, roots);
  }
}
inline void eigen33f(const float mat[3][3], float &eigenvalue, float evec[3]) {  // pcl::eigen33(mat, eigenvalue, eigenvector)
  float scale = 0;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) scale = std::max(scale, std::fabs(mat[i][j]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float sm[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) sm[i][j] = mat[i][j] / scale;
  float ev[3]; computeRoots3f(sm, ev);
  eigenvalue = ev[0] * scale;
  for (int i = 0; i < 3; i++) sm[i][i] -= ev[0];
  auto crossf = [](const float a[3], const float b[3], float o[3]) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
  float v1[3], v2[3], v3[3];
  crossf(sm[0], sm[1], v1); crossf(sm[0], sm[2], v2); crossf(sm[1], sm[2], v3);
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float *v; float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; } else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; } else { v = v3; l = l3; }
  float s = std::sqrt(l);
  evec[0] = v[0] / s; evec[1] = v[1] / s; evec[2] = v[2] / s;
}

// PlaneExtractor::extractPlanes (PlaneExtractor.cpp:22-107)
inline void extract_planes(const EslamParams &prm, const EslamCamera &cam, const uint16_t *depth, size_t pitch, PlaneDebug &dbg) {
  const int W = cam.cols, H = cam.rows;
  const size_t NP = (size_t)W * H;
  std::vector<P3f> pts(NP);
  for (int m = 0; m < H; m++) for (int n = 0; n < W; n++) {  // :31-51 (zero depth kept as (0,0,0))
    uint16_t dv = depth[(size_t)m * pitch + n];
    double d = dv / cam.scale;
    P3f p; p.z = (float)d;
    p.x = (float)((n - cam.cx) * p.z / cam.fx);
    p.y = (float)((m - cam.cy) * p.z / cam.fy);
    pts[(size_t)m * W + n] = p;
  }
  // ---- IntegralImageNormalEstimation::initAverage3DGradientMethod
  std::vector<float> diff_x(NP * 4, 0.0f), diff_y(NP * 4, 0.0f);
  for (int ri = 1; ri < H - 1; ++ri) for (int ci = 1; ci < W - 1; ++ci) {
    size_t o = ((size_t)ri * W + ci) * 4;
    const P3f &rg = pts[(size_t)ri * W + ci + 1], &lf = pts[(size_t)ri * W + ci - 1], &dn = pts[(size_t)(ri + 1) * W + ci], &up = pts[(size_t)(ri - 1) * W + ci];
    diff_x[o + 0] = rg.x - lf.x; diff_x[o + 1] = rg.y - lf.y; diff_x[o + 2] = rg.z - lf.z;
    diff_y[o + 0] = dn.x - up.x; diff_y[o + 1] = dn.y - up.y; diff_y[o + 2] = dn.z - up.z;
  }
  // IntegralImage2D<float,3>::computeIntegralImages (first order, double accumulators)
  const int IW = W + 1;
  std::vector<double> iiX((size_t)IW * (H + 1) * 3, 0.0), iiY((size_t)IW * (H + 1) * 3, 0.0);
  auto integral = [&](const std::vector<float> &data, std::vector<double> &ii) {
    for (int r = 0; r < H; r++) {
      double *prev = &ii[(size_t)r * IW * 3], *cur = &ii[(size_t)(r + 1) * IW * 3];
      cur[0] = cur[1] = cur[2] = 0;
      for (int c = 0; c < W; c++) for (int k = 0; k < 3; k++) {
        double v = prev[(c + 1) * 3 + k] + cur[c * 3 + k] - prev[c * 3 + k];
        v += (double)data[((size_t)r * W + c) * 4 + k];
        cur[(c + 1) * 3 + k] = v;
      }
    }
  };
  integral(diff_x, iiX); integral(diff_y, iiY);
  auto first_order_sum = [&](const std::vector<double> &ii, int sx, int sy, int w, int h, double out[3]) {
    size_t ul = (size_t)sy * IW + sx, ur = ul + w, ll = (size_t)(sy + h) * IW + sx, lr = ll + w;
    for (int k = 0; k < 3; k++) out[k] = ii[lr * 3 + k] + ii[ul * 3 + k] - ii[ur * 3 + k] - ii[ll * 3 + k];
  };
  // ---- computeFeature: depth change map + chamfer distance map
  const float max_depth_change_factor = 0.05f, normal_smoothing_size = 10.0f;
  std::vector<unsigned char> dcm(NP, 255);
  for (int ri = 0; ri < H - 1; ++ri) for (int ci = 0; ci < W - 1; ++ci) {
    size_t index = (size_t)ri * W + ci;
    const float depth_ = pts[index].z, depthR = pts[index + 1].z, depthD = pts[index + W].z;
    const float ddc = (max_depth_change_factor * (std::fabs(depth_) + 1.0f) * 2.0f);
    if (std::fabs(depth_ - depthR) > ddc || !std::isfinite(depth_) || !std::isfinite(depthR)) { dcm[index] = 0; dcm[index + 1] = 0; }
    if (std::fabs(depth_ - depthD) > ddc || !std::isfinite(depth_) || !std::isfinite(depthD)) { dcm[index] = 0; dcm[index + W] = 0; }
  }
  std::vector<float> &dm = dbg.dist_map; dm.assign(NP, 0.0f);
  for (size_t i = 0; i < NP; i++) dm[i] = dcm[i] == 0 ? 0.0f : (float)(W + H);
  {  // first pass (note previous_row[ci+1] at ci==W-1 reads current_row[0]: flat indexing reproduces that)
    for (int ri = 1; ri < H; ++ri) {
      float *previous_row = &dm[(size_t)(ri - 1) * W], *current_row = &dm[(size_t)ri * W];
      for (int ci = 1; ci < W; ++ci) {
        const float upLeft = previous_row[ci - 1] + 1.4f, up = previous_row[ci] + 1.0f, upRight = previous_row[ci + 1] + 1.4f;
        const float left = current_row[ci - 1] + 1.0f, center = current_row[ci];
        const float minValue = std::min(std::min(upLeft, up), std::min(left, upRight));
        if (minValue < center) current_row[ci] = minValue;
      }
    }
    for (int ri = H - 2; ri >= 0; --ri) {  // second pass (next_row[ci-1] at ci==0 reads current_row[W-1])
      float *next_row = &dm[(size_t)(ri + 1) * W], *current_row = &dm[(size_t)ri * W];
      for (int ci = W - 2; ci >= 0; --ci) {
        const float lowerLeft = next_row[ci - 1] + 1.4f, lower = next_row[ci] + 1.0f, lowerRight = next_row[ci + 1] + 1.4f;
        const float right = current_row[ci + 1] + 1.0f, center = current_row[ci];
        const float minValue = std::min(std::min(lowerLeft, right), std::min(lower, lowerRight));
        if (minValue < center) current_row[ci] = minValue;
      }
    }
  }
  // ---- computeFeatureFull (BORDER_POLICY_IGNORE, no depth dependent smoothing)
  const float NaN = std::numeric_limits<float>::quiet_NaN();
  std::vector<float> &nrm = dbg.normals; nrm.assign(NP * 3, NaN);
  const int border = (int)normal_smoothing_size;
  for (int ri = border; ri < H - border; ++ri) for (int ci = border; ci < W - border; ++ci) {
    size_t index = (size_t)ri * W + ci;
    if (!std::isfinite(pts[index].z)) continue;
    float smoothing = std::min(dm[index], normal_smoothing_size);
    if (smoothing > 2.0f) {
      int rw = (int)smoothing, rh = (int)smoothing, rw2 = rw / 2, rh2 = rh / 2;  // setRectSize
      double gx[3], gy[3];
      first_order_sum(iiX, ci - rw2, ri - rh2, rw, rh, gx);
      first_order_sum(iiY, ci - rw2, ri - rh2, rw, rh, gy);
      double nv[3] = {gy[1] * gx[2] - gy[2] * gx[1], gy[2] * gx[0] - gy[0] * gx[2], gy[0] * gx[1] - gy[1] * gx[0]};  // gradient_y.cross(gradient_x)
      double nl = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
      if (nl == 0.0) continue;
      double s = std::sqrt(nl);
      float nx = (float)(nv[0] / s), ny = (float)(nv[1] / s), nz = (float)(nv[2] / s);
      // flipNormalTowardsViewpoint (vp = 0)
      float vx = 0.0f - pts[index].x, vy = 0.0f - pts[index].y, vz = 0.0f - pts[index].z;
      float cos_theta = (vx * nx + vy * ny + vz * nz);
      if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
      nrm[index * 3] = nx; nrm[index * 3 + 1] = ny; nrm[index * 3 + 2] = nz;
    }
  }
  // ---- OrganizedMultiPlaneSegmentation::segment
  std::vector<float> plane_d(NP);
  for (size_t i = 0; i < NP; i++) plane_d[i] = pts[i].x * nrm[i * 3] + pts[i].y * nrm[i * 3 + 1] + pts[i].z * nrm[i * 3 + 2];
  const float AngTh = (float)prm.plane_angle_deg, DisTh = (float)prm.plane_dist;
  const float ang_thresh = std::cos((float)(0.017453 * AngTh));     // PlaneCoefficientComparator::setAngularThreshold: cosf
  auto compare = [&](size_t idx1, size_t idx2) -> bool {           // PlaneCoefficientComparator::compare, depth dependent
    float threshold = DisTh;
    float z = pts[idx1].z;
    threshold *= z * z;
    float dotn = nrm[idx1 * 3] * nrm[idx2 * 3] + nrm[idx1 * 3 + 1] * nrm[idx2 * 3 + 1] + nrm[idx1 * 3 + 2] * nrm[idx2 * 3 + 2];
    return (std::fabs(plane_d[idx1] - plane_d[idx2]) < threshold) && (dotn > ang_thresh);
  };
  // OrganizedConnectedComponentSegmentation::segment (all points finite here)
  const unsigned invalid_label = std::numeric_limits<unsigned>::max();
  std::vector<unsigned> labels(NP, invalid_label), run_ids;
  auto findRoot = [&](unsigned i) { while (run_ids[i] != i) i = run_ids[i]; return i; };
  unsigned clust_id = 0;
  labels[0] = clust_id++; run_ids.push_back(labels[0]);
  for (int c = 1; c < W; ++c) {
    if (compare(c, c - 1)) labels[c] = labels[c - 1];
    else { labels[c] = clust_id++; run_ids.push_back(labels[c]); }
  }
  for (int r = 1; r < H; ++r) {
    size_t cur = (size_t)r * W, prev = (size_t)(r - 1) * W;
    if (compare(cur, prev)) labels[cur] = labels[prev];
    else { labels[cur] = clust_id++; run_ids.push_back(labels[cur]); }
    for (int c = 1; c < W; ++c) {
      if (compare(cur + c, cur + c - 1)) labels[cur + c] = labels[cur + c - 1];
      if (compare(cur + c, prev + c)) {
        if (labels[cur + c] == invalid_label) labels[cur + c] = labels[prev + c];
        else if (labels[prev + c] != invalid_label) {
          unsigned root1 = findRoot(labels[cur + c]), root2 = findRoot(labels[prev + c]);
          if (root1 < root2) run_ids[root2] = root1; else run_ids[root1] = root2;
        }
      }
      if (labels[cur + c] == invalid_label) { labels[cur + c] = clust_id++; run_ids.push_back(labels[cur + c]); }
    }
  }
  std::vector<unsigned> map(clust_id);
  unsigned max_id = 0;
  for (unsigned runIdx = 0; runIdx < run_ids.size(); ++runIdx) {
    if (run_ids[runIdx] == runIdx) map[runIdx] = max_id++;
    else map[runIdx] = map[findRoot(runIdx)];
  }
  std::vector<std::vector<int>> label_indices(max_id + 1);
  for (size_t idx = 0; idx < NP; idx++) { labels[idx] = map[labels[idx]]; label_indices[labels[idx]].push_back((int)idx); }
  // fit planes
  struct Model { float v[4]; };
  std::vector<Model> models; std::vector<std::vector<int>> inliers;
  float vp[4] = {0, 0, 0, 0};
  const unsigned min_inliers = 100; const float maximum_curvature = 0.001f;
  for (size_t i = 0; i < label_indices.size(); i++) {
    if ((unsigned)label_indices[i].size() > min_inliers) {
      float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // computeMeanAndCovarianceMatrix<PointT,float>
      for (int id : label_indices[i]) {
        const P3f &p = pts[id];
        accu[0] += p.x * p.x; accu[1] += p.x * p.y; accu[2] += p.x * p.z; accu[3] += p.y * p.y; accu[4] += p.y * p.z; accu[5] += p.z * p.z;
        accu[6] += p.x; accu[7] += p.y; accu[8] += p.z;
      }
      float cnt = (float)label_indices[i].size();
      for (int k = 0; k < 9; k++) accu[k] /= cnt;
      float cen[4] = {accu[6], accu[7], accu[8], 1.0f};
      float cov[3][3];
      cov[0][0] = accu[0] - accu[6] * accu[6]; cov[0][1] = accu[1] - accu[6] * accu[7]; cov[0][2] = accu[2] - accu[6] * accu[8];
      cov[1][1] = accu[3] - accu[7] * accu[7]; cov[1][2] = accu[4] - accu[7] * accu[8]; cov[2][2] = accu[5] - accu[8] * accu[8];
      cov[1][0] = cov[0][1]; cov[2][0] = cov[0][2]; cov[2][1] = cov[1][2];
      float eigen_value, ev[3];
      eigen33f(cov, eigen_value, ev);
      float pp[4] = {ev[0], ev[1], ev[2], 0};
      pp[3] = -1 * (pp[0] * cen[0] + pp[1] * cen[1] + pp[2] * cen[2] + pp[3] * cen[3]);
      for (int k = 0; k < 4; k++) vp[k] -= cen[k];   // (sic) vp accumulates across clusters in PCL
      float cos_theta = vp[0] * pp[0] + vp[1] * pp[1] + vp[2] * pp[2] + vp[3] * pp[3];
      if (cos_theta < 0) {
        for (int k = 0; k < 4; k++) pp[k] *= -1;
        pp[3] = 0;
        pp[3] = -1 * (pp[0] * cen[0] + pp[1] * cen[1] + pp[2] * cen[2] + pp[3] * cen[3]);
      }
      float curvature;
      float eig_sum = cov[0][0] + cov[1][1] + cov[2][2];
      if (eig_sum != 0) curvature = std::fabs(eigen_value / eig_sum); else curvature = 0;
      if (curvature < maximum_curvature) {
        Model m; for (int k = 0; k < 4; k++) m.v[k] = pp[k];
        models.push_back(m); inliers.push_back(label_indices[i]);
      }
    }
  }
  // ---- refine (PlaneRefinementComparator: distance_threshold_ 0.02 default, not depth dependent)
  std::vector<char> grow(label_indices.size(), 0); std::vector<int> label_to_model(label_indices.size(), 0);
  for (size_t i = 0; i < models.size(); i++) { int ml = (int)labels[inliers[i][0]]; label_to_model[ml] = (int)i; grow[ml] = 1; }
  const float refine_thresh = 0.02f;
  auto rcompare = [&](size_t idx1, size_t idx2) -> bool {
    int cl = (int)labels[idx1], nl = (int)labels[idx2];
    if (!(grow[cl] && !grow[nl])) return false;
    const Model &mc = models[label_to_model[cl]];
    const P3f &pt = pts[idx2];
    double ptp = std::fabs(mc.v[0] * pt.x + mc.v[1] * pt.y + mc.v[2] * pt.z + mc.v[3]);
    return ptp < refine_thresh;
  };
  for (int r = 0; r < H - 1; ++r) {
    size_t cur = (size_t)r * W, nxt = cur + W;
    for (int c = 0; c < W - 1; ++c) {
      int current_label = (int)labels[cur + c];
      if (rcompare(cur + c, cur + c + 1)) { labels[cur + c + 1] = current_label; inliers[label_to_model[current_label]].push_back((int)(cur + c + 1)); }
      if (rcompare(cur + c, nxt + c)) { labels[nxt + c] = current_label; inliers[label_to_model[current_label]].push_back((int)(nxt + c)); }
    }
  }
  {
    size_t cur = (size_t)W * (H - 1);
    for (int r = 0; r < H - 1; ++r, cur -= W) {
      size_t prv = cur - W;
      for (int c = W - 1; c >= 0; --c) {
        int current_label = (int)labels[cur + c];
        if (rcompare(cur + c, cur + c - 1)) { labels[cur + c - 1] = current_label; inliers[label_to_model[current_label]].push_back((int)(cur + c - 1)); }
        if (rcompare(cur + c, prv + c)) { labels[prv + c] = current_label; inliers[label_to_model[current_label]].push_back((int)(prv + c)); }
      }
    }
  }
  dbg.labels.assign(NP, -1);
  for (size_t i = 0; i < NP; i++) dbg.labels[i] = (int)labels[i];
  // :89-106
  dbg.regions.clear();
  for (size_t i = 0; i < inliers.size(); ++i) {
    int planeSize = (int)inliers[i].size();
    if (planeSize < prm.plane_min_size) continue;
    PlaneRegion R; for (int k = 0; k < 4; k++) R.coef[k] = models[i].v[k];
    if (R.coef[3] < 0) for (int k = 0; k < 4; k++) R.coef[k] = -R.coef[k];
    R.size = planeSize; R.first_pixel = inliers[i][0];
    dbg.regions.push_back(R);
  }
}

// PlaneExtractor::extractGroundPlane (PlaneExtractor.cpp:126-183)
inline bool extract_ground_plane(const EslamParams &prm, const EslamCamera &cam, const uint16_t *depth, size_t pitch, double plane[4], int *n_inliers, PlaneDebug *dbg_out = nullptr) {
  PlaneDebug local; PlaneDebug &dbg = dbg_out ? *dbg_out : local;
  extract_planes(prm, cam, depth, pitch, dbg);
  if (dbg.regions.size() < 1) return false;
  struct Cand { double v[4]; int size; };
  std::vector<Cand> cands;
  for (const PlaneRegion &R : dbg.regions) {
    double v[4] = {(double)R.coef[0], (double)R.coef[1], (double)R.coef[2], (double)R.coef[3]};
    double cos_theta = v[0] * 0.0 + v[1] * 1.0 + v[2] * 0.0;
    cos_theta = cos_theta / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / 1.0;
    double theta = std::acos(cos_theta);
    if (theta > M_PI / 4 && theta < 3 * M_PI / 4) continue;
    Cand c; for (int k = 0; k < 4; k++) c.v[k] = v[k]; c.size = R.size; cands.push_back(c);
  }
  if (cands.size() < 1) return false;
  std::stable_sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) { return a.size > b.size; });
  double *g = cands[0].v;
  double value = plane_distance_to_point(g, V3{0, 0, 0}, true);
  if (value < 0) for (int k = 0; k < 4; k++) g[k] = -g[k];
  for (int k = 0; k < 4; k++) plane[k] = g[k];
  if (n_inliers) *n_inliers = cands[0].size;
  return true;
}

}  // namespace orc

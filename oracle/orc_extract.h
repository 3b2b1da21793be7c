// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
// orc_extract.h: single-frame ellipsoid extraction (src/pca/EllipsoidExtractor.cpp), point-cloud helpers
// (src/symmetry/PointCloudFilter.cpp, src/core/Geometry.cpp), g2o::plane (include/core/Plane.h, src/core/Plane.cpp)
// and the symmetry solver (src/symmetry/Symmetry.cpp, SymmetrySolver.cpp), restated for one CPU thread.
//
// PCL is NOT vendored in the reference (README.md:36 "PCL 1.7+", unpinned).  The PCL pieces restate the published
// upstream algorithms (PCL 1.8-1.10 sources as remembered) -- PARITY UNPINNED there:
//   pcl::VoxelGrid<PointXYZRGB>::applyFilter      -> voxel_grid()
//   pcl::EuclideanClusterExtraction<PointXYZ>     -> euclidean_clusters()
//   pcl::compute3DCentroid / computeCovarianceMatrixNormalized -> inside process_pca()
//   pcl::KdTreeFLANN::nearestKSearch(K=1)         -> nearest_float()   (exact NN, flann::L2_Simple<float>)
// Documented deviation: PCL sorts voxel (idx,point) pairs with the UNSTABLE std::sort, so the float summation order
// inside a voxel is implementation-defined in the reference; the oracle's canonical order is ascending point index
// (stable).  `unstable_voxel_sort=1` switches to std::sort on the PCL struct to quantify the (<= few ulp) effect.
#pragma once
#include "orc_graph.h"
#include "../include/eslam.h"
#include <cstdio>

namespace orc {

struct P3f { float x, y, z; };
struct P3d { double x, y, z; };

// ---------------- g2o::plane ----------------
struct Plane { double p[4]; double dual_dis; };
inline Plane plane_from_dis_angle_trans(double dis, double angle, double trans) {  // Plane.cpp:97-105
  Plane pl; pl.p[0] = std::sin(angle); pl.p[1] = -std::cos(angle); pl.p[2] = 0; pl.p[3] = -dis; pl.dual_dis = trans; return pl;
}
inline double plane_distance_to_point(const double p[4], V3 pt, bool keep_flag) {  // Plane.cpp:108-115
  double fenzi = p[0] * pt.x + p[1] * pt.y + p[2] * pt.z + p[3];
  double fenmu = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  double value = fenzi / fenmu;
  return keep_flag ? value : std::fabs(value);
}
inline void plane_transform(double p[4], const SE3 &Twc) {  // Plane.cpp:117-122: param = (Twc^T)^-1 param
  M4 T = se3_to_hom(Twc), Tt = transpose(T), inv;
  inverse4(Tt, inv);
  double r[4];
  for (int i = 0; i < 4; i++) { double s = inv.m[i][0] * p[0]; for (int k = 1; k < 4; k++) s += inv.m[i][k] * p[k]; r[i] = s; }
  for (int i = 0; i < 4; i++) p[i] = r[i];
}
inline M3 plane_rotation(V3 v) {  // Plane.h:88-92: (AngleAxis(az,Z)*AngleAxis(-el,Y)).toRotationMatrix()
  double az = std::atan2(v.y, v.x);
  double el = std::atan2(v.z, std::sqrt(v.x * v.x + v.y * v.y));
  Q4 q = q_mul(q_from_angle_axis(az, V3{0, 0, 1}), q_from_angle_axis(-el, V3{0, 1, 0}));
  return q_to_R(q);
}
inline void plane_normalize(double p[4]) {  // Plane.h:79-82
  double n = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  double inv = 1. / n;
  for (int i = 0; i < 4; i++) p[i] = p[i] * inv;
}
inline void plane_oplus(Plane &pl, const double v[3]) {  // Plane.h:46-59
  double az = v[0], el = v[1];
  double s = std::sin(el), c = std::cos(el);
  V3 n{c * std::cos(az), c * std::sin(az), s};
  M3 R = plane_rotation(V3{pl.p[0], pl.p[1], pl.p[2]});
  double d = -pl.p[3] + v[2];
  V3 rn = mul(R, n);
  pl.p[0] = rn.x; pl.p[1] = rn.y; pl.p[2] = rn.z; pl.p[3] = -d;
  plane_normalize(pl.p);
}
inline void plane_oplus_dual(Plane &pl, const double v[3]) {  // Plane.h:62-77
  double az = v[0], el = 0;
  double s = std::sin(el), c = std::cos(el);
  V3 n{c * std::cos(az), c * std::sin(az), s};
  M3 R = plane_rotation(V3{pl.p[0], pl.p[1], pl.p[2]});
  double d = -pl.p[3] + v[1];
  V3 rn = mul(R, n);
  pl.p[0] = rn.x; pl.p[1] = rn.y; pl.p[2] = rn.z; pl.p[3] = -d;
  plane_normalize(pl.p);
  pl.dual_dis += v[2];
}
inline void plane_another_vec(const Plane &pl, double out[4]) {  // GenerateAnotherPlaneVec, Plane.cpp:137-144
  Plane p2 = plane_from_dis_angle_trans(pl.dual_dis, std::atan2(pl.p[1], pl.p[0]), 0);
  for (int i = 0; i < 4; i++) out[i] = p2.p[i];
}

// ---------------- S1 getPointCloudInRect (PointCloudFilter.cpp:21-56) ----------------
inline bool bbox_inside(const double bbox[4], const EslamCamera &cam) {
  int x1 = int(bbox[0]), y1 = int(bbox[1]), x2 = int(bbox[2]), y2 = int(bbox[3]);
  return x1 >= 0 && y1 >= 0 && x2 <= cam.cols && y2 <= cam.rows;
}
inline void get_pointcloud_in_rect(const uint16_t *depth, size_t pitch, const double bbox[4], const EslamCamera &cam, double range, std::vector<P3d> &cloud) {
  int x1 = int(bbox[0]), y1 = int(bbox[1]), x2 = int(bbox[2]), y2 = int(bbox[3]);
  cloud.clear();
  for (int y = y1; y < y2; y = y + 3)
    for (int x = x1; x < x2; x = x + 3) {
      uint16_t d = depth[(size_t)y * pitch + x];
      double z = d / cam.scale;
      if (z <= 0.1 || z > range) continue;
      P3d p; p.z = z;
      p.x = (x - cam.cx) * p.z / cam.fx;
      p.y = (y - cam.cy) * p.z / cam.fy;
      cloud.push_back(p);
    }
}

// ---------------- S2 DownSamplePointCloudOnly (PointCloudFilter.cpp:249-261) = pcl::VoxelGrid ----------------
struct VoxIdx { unsigned idx; unsigned cloud_point_index; bool operator<(const VoxIdx &o) const { return idx < o.idx; } };
// returns false when PCL would return the input unfiltered (index overflow warning)
inline bool voxel_grid(const std::vector<P3f> &in, double leaf_d, std::vector<P3f> &out, bool unstable_sort = false) {
  out.clear();
  if (in.empty()) return true;
  float leaf = (float)leaf_d;            // setLeafSize(float,float,float)
  float inv = 1.0f / leaf;               // inverse_leaf_size_ = Array4f::Ones()/leaf_size_.array()
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  for (const P3f &p : in) {  // getMinMax3D
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if ((dx * dy * dz) > (int64_t)std::numeric_limits<int32_t>::max()) { out = in; return false; }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; a++) { min_b[a] = (int)std::floor(mn[a] * inv); max_b[a] = (int)std::floor(mx[a] * inv); div_b[a] = max_b[a] - min_b[a] + 1; }
  int mul0 = 1, mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
  std::vector<VoxIdx> iv; iv.reserve(in.size());
  for (unsigned i = 0; i < in.size(); i++) {
    int ijk0 = (int)(std::floor(in[i].x * inv) - (float)min_b[0]);
    int ijk1 = (int)(std::floor(in[i].y * inv) - (float)min_b[1]);
    int ijk2 = (int)(std::floor(in[i].z * inv) - (float)min_b[2]);
    int idx = ijk0 * mul0 + ijk1 * mul1 + ijk2 * mul2;
    iv.push_back(VoxIdx{(unsigned)idx, i});
  }
  if (unstable_sort) std::sort(iv.begin(), iv.end());
  else std::stable_sort(iv.begin(), iv.end());
  size_t index = 0;
  while (index < iv.size()) {
    size_t i = index + 1;
    while (i < iv.size() && iv[i].idx == iv[index].idx) ++i;
    float sx = 0, sy = 0, sz = 0;  // AccumulatorXYZ: Vector3f
    for (size_t li = index; li < i; ++li) { const P3f &p = in[iv[li].cloud_point_index]; sx += p.x; sy += p.y; sz += p.z; }
    float n = (float)(i - index);
    out.push_back(P3f{sx / n, sy / n, sz / n});
    index = i;
  }
  return true;
}
inline void downsample_only(const std::vector<P3d> &in, std::vector<P3d> &out, double grid, bool unstable_sort = false) {
  std::vector<P3f> f(in.size()), o;
  for (size_t i = 0; i < in.size(); i++) f[i] = P3f{(float)in[i].x, (float)in[i].y, (float)in[i].z};  // QuadricPointCloudToPcl :133-153
  voxel_grid(f, grid, o, unstable_sort);
  out.resize(o.size());
  for (size_t i = 0; i < o.size(); i++) out[i] = P3d{(double)o[i].x, (double)o[i].y, (double)o[i].z};   // pclToQuadricPointCloud :92-110
}

// ---------------- S6 pcl::EuclideanClusterExtraction ----------------
inline float l2_simple(const P3f &a, const P3f &b) {  // flann::L2_Simple<float>
  float result = 0.0f, diff;
  diff = a.x - b.x; result += diff * diff;
  diff = a.y - b.y; result += diff * diff;
  diff = a.z - b.z; result += diff * diff;
  return result;
}
// clusters: sorted by size descending, ties in discovery order; each cluster's indices ascending
inline void euclidean_clusters(const std::vector<P3f> &pts, double tolerance, int min_size, int max_size, std::vector<std::vector<int>> &clusters) {
  clusters.clear();
  const int n = (int)pts.size();
  if (n == 0) return;
  const float r2 = (float)(tolerance * tolerance);
  // uniform grid with cell = tolerance for the radius queries (exact: every pair within r lies in adjacent cells)
  float mn[3] = {pts[0].x, pts[0].y, pts[0].z};
  for (const P3f &p : pts) { mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z); }
  const double cell = tolerance * 1.0001;
  auto cellof = [&](const P3f &p, int c[3]) {
    c[0] = (int)std::floor(((double)p.x - mn[0]) / cell); c[1] = (int)std::floor(((double)p.y - mn[1]) / cell); c[2] = (int)std::floor(((double)p.z - mn[2]) / cell);
  };
  struct Key { int64_t k; int i; };
  auto mk = [](const int c[3]) { return ((int64_t)c[2] << 42) | ((int64_t)c[1] << 21) | (int64_t)c[0]; };
  std::vector<Key> keys(n);
  for (int i = 0; i < n; i++) { int c[3]; cellof(pts[i], c); keys[i] = Key{mk(c), i}; }
  std::sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) { return a.k < b.k || (a.k == b.k && a.i < b.i); });
  auto range_of = [&](int64_t k, int &lo, int &hi) {
    lo = (int)(std::lower_bound(keys.begin(), keys.end(), k, [](const Key &a, int64_t kk) { return a.k < kk; }) - keys.begin());
    hi = lo; while (hi < n && keys[hi].k == k) hi++;
  };
  std::vector<char> processed(n, 0);
  std::vector<int> queue;
  for (int i = 0; i < n; i++) {
    if (processed[i]) continue;
    queue.clear(); queue.push_back(i); processed[i] = 1;
    size_t sq = 0;
    while (sq < queue.size()) {
      int q = queue[sq];
      int c[3]; cellof(pts[q], c);
      for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
        int cc[3] = {c[0] + dx, c[1] + dy, c[2] + dz};
        if (cc[0] < 0 || cc[1] < 0 || cc[2] < 0) continue;
        int lo, hi; range_of(mk(cc), lo, hi);
        for (int t = lo; t < hi; t++) {
          int j = keys[t].i;
          if (processed[j]) continue;
          if (l2_simple(pts[q], pts[j]) < r2) { queue.push_back(j); processed[j] = 1; }  // RadiusResultSet::addPoint: dist < radius
        }
      }
      sq++;
    }
    if ((int)queue.size() >= min_size && (int)queue.size() <= max_size) {
      std::vector<int> r(queue); std::sort(r.begin(), r.end());
      clusters.push_back(r);
    }
  }
  std::stable_sort(clusters.begin(), clusters.end(), [](const std::vector<int> &a, const std::vector<int> &b) { return a.size() > b.size(); });
}

// ---------------- S7/S8 PCA + frame ----------------
struct PCAResult { V3 center; M3 rot; V3 cov; V3 scale; };
inline PCAResult process_pca(const std::vector<P3f> &cloud) {  // EllipsoidExtractor.cpp:170-196
  double c[3] = {0, 0, 0};
  for (const P3f &p : cloud) { c[0] += p.x; c[1] += p.y; c[2] += p.z; }   // compute3DCentroid(Vector4d)
  double N = (double)cloud.size();
  c[0] /= N; c[1] /= N; c[2] /= N;
  M3 cov = m3_zero();
  for (const P3f &p : cloud) {  // computeCovarianceMatrix (dense branch)
    double px = p.x - c[0], py = p.y - c[1], pz = p.z - c[2];
    cov.m[1][1] += py * py; cov.m[1][2] += py * pz; cov.m[2][2] += pz * pz;
    double sx = px;
    px = px * sx; py = py * sx; pz = pz * sx;
    cov.m[0][0] += px; cov.m[0][1] += py; cov.m[0][2] += pz;
  }
  cov.m[1][0] = cov.m[0][1]; cov.m[2][0] = cov.m[0][2]; cov.m[2][1] = cov.m[1][2];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) cov.m[i][j] /= N;
  PCAResult out;
  double ev[3];
  self_adjoint_eigen3(cov, ev, out.rot);
  // PointType c (float!) :184-191
  out.center = V3{(double)(float)c[0], (double)(float)c[1], (double)(float)c[2]};
  out.cov = V3{ev[0], ev[1], ev[2]};
  out.scale = V3{0, 0, 0};
  return out;
}
inline void adjust_chirality(PCAResult &d) { setcol(d.rot, 2, cross(col(d.rot, 0), col(d.rot, 1))); }  // :760-764
inline void align_z_axis_to_gravity(PCAResult &d, const double plane[4]) {  // :207-267
  double max_abs = 0; bool max_pos = false; int max_id = -1;
  V3 z_axis = normalized(V3{plane[0], plane[1], plane[2]});
  for (int i = 0; i < 3; i++) {
    double cos_theta = dot(col(d.rot, i), z_axis);
    bool flag_pos = cos_theta > 0;
    double a = std::fabs(cos_theta);
    if (a > max_abs) { max_abs = a; max_pos = flag_pos; max_id = i; }
  }
  if (max_id < 0) max_id = 0;  // assert compiled out in Release; keep deterministic
  M3 sw; V3 cs;
  V3 zv = max_pos ? col(d.rot, max_id) : (col(d.rot, max_id) * -1.0);
  setcol(sw, 2, zv); cs.z = get(d.cov, max_id);
  int xi = (max_id + 1) % 3;
  setcol(sw, 0, col(d.rot, xi)); cs.x = get(d.cov, xi);
  int yi = (max_id + 2) % 3;
  setcol(sw, 1, cross(col(sw, 2), col(sw, 0))); cs.y = get(d.cov, yi);
  d.rot = sw; d.cov = cs;
}
inline M3 calib_rot_to_ground(const M3 &rot, V3 normal) {  // :766-788
  V3 zA = col(rot, 2);
  V3 axis = normalized(cross(zA, normal));
  double norm1 = norm(normal), norm2 = norm(zA);
  double vec_dot = dot(normal, zA);
  double cos_theta = vec_dot / norm1 / norm2;
  double theta = std::acos(cos_theta);
  return mul(angle_axis_to_R(theta, axis), rot);
}

// ---------------- S5 GetCenter (:583-643) ----------------
inline bool get_center(const uint16_t *depth, size_t pitch, const double bbox[4], const SE3 &Twc, const EslamCamera &cam, double depth_range, V3 &center) {
  int x = int((bbox[0] + bbox[2]) / 2.0);
  int y = int((bbox[1] + bbox[3]) / 2.0);
  int point_num = 10;
  int x_delta = std::fabs(bbox[0] - bbox[2]) / 4.0 / point_num;
  int y_delta = std::fabs(bbox[1] - bbox[3]) / 4.0 / point_num;
  std::vector<P3f> cloud;
  for (int x_id = -point_num / 2; x_id < point_num / 2; x_id++)
    for (int y_id = -point_num / 2; y_id < point_num / 2; y_id++) {
      int x_ = x + x_id * x_delta, y_ = y + y_id * y_delta;
      if (x_ < 0 || y_ < 0 || x_ >= cam.cols || y_ >= cam.rows) continue;  // (reference would read out of bounds)
      uint16_t d = depth[(size_t)y_ * pitch + x_];
      P3f p;
      p.z = d / cam.scale;
      if (p.z <= 0.1 || p.z > depth_range) continue;
      p.x = (x_ - cam.cx) * p.z / cam.fx;
      p.y = (y_ - cam.cy) * p.z / cam.fy;
      cloud.push_back(p);
    }
  if (cloud.size() < 2) return false;
  double c[3] = {0, 0, 0};
  for (const P3f &p : cloud) { c[0] += p.x; c[1] += p.y; c[2] += p.z; }
  double N = (double)cloud.size();
  c[0] /= N; c[1] /= N; c[2] /= N;
  center = transform_point_hom(se3_to_hom(Twc), V3{c[0], c[1], c[2]});
  return true;
}

// ---------------- symmetry (S9-S13) ----------------
struct SymContext {
  const uint16_t *depth; size_t pitch; EslamCamera cam; double bbox[4];
  SE3 T_oc;                 // camera pose in the object-normalised frame (poseNormalized, :373-375)
  double K[9];
  const std::vector<P3d> *cloud;  // object-normalised cloud
  std::vector<P3f> cloud_f;       // kd-tree input (initializeKDTree, SymmetrySolver.cpp:411-433)
  double sigma;
  long long n_cost_evals = 0;
};
// Symmetry::getProjDepthMat (Symmetry.cpp:141-161), evaluated lazily per pixel
inline uint16_t proj_depth_at(const SymContext &c, int x, int y) {
  uint16_t d = c.depth[(size_t)y * c.pitch + x];
  double xi = (x - c.cam.cx), yi = (y - c.cam.cy), f = c.cam.fx;
  double realz = double(d) * std::sqrt(xi * xi + f * f + yi * yi) / f;   // calculateProjZ :137-139
  return (uint16_t)(int32_t)realz;  // ushort(realz): x86 cvttsd2si then truncation to 16 bit
}
inline V3 symmetry_point_of_plane(V3 pt, const double pl[4]) {  // SymmetrySolver.cpp:193-214
  V3 normal = normalized(V3{pl[0], pl[1], pl[2]});
  double val = pl[0] * pt.x + pl[1] * pt.y + pl[2] * pt.z + pl[3] * 1.0;
  double up = std::fabs(val);
  double down = std::sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]);
  double dis = up / down;
  double symbol = (val > 0) ? -1 : 1;
  // point_3 + 2*symbol*dis*normal ; then real_to_homo/homo_to_real (divide by 1)
  double k = 2 * symbol * dis;
  return V3{pt.x + k * normal.x, pt.y + k * normal.y, pt.z + k * normal.z};
}
inline int nearest_float(const std::vector<P3f> &cloud, const P3f &q) {
  int best = 0; float bd = std::numeric_limits<float>::max();
  for (int i = 0; i < (int)cloud.size(); i++) { float d = l2_simple(q, cloud[i]); if (d < bd) { bd = d; best = i; } }
  return best;
}
inline double find_minimal_distance(const SymContext &c, V3 p) {  // findMinimalDistanceWithKdTree :140-166
  P3f q{(float)p.x, (float)p.y, (float)p.z};
  int id = nearest_float(c.cloud_f, q);
  const P3f &n = c.cloud_f[id];
  double dx = p.x - (double)n.x, dy = p.y - (double)n.y, dz = p.z - (double)n.z;
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}
// GetPointCloudProb (SymmetrySolver.cpp:49-138) on an explicit mirrored cloud
inline double pointcloud_prob(SymContext &c, const std::vector<V3> &sym) {
  c.n_cost_evals++;
  SE3 wc = c.T_oc, cw = se3_inverse(wc);
  double P[3][4]; projection_matrix(cw, c.K, P);
  int num = (int)sym.size();
  double ln_total_P = 0; int num_invalid = 0;
  int bx0 = (int)c.bbox[0], bx2 = (int)c.bbox[2], by1 = (int)c.bbox[1], by3 = (int)c.bbox[3];
  for (int i = 0; i < num; i++) {
    V3 p = sym[i];
    double uvh[3];
    for (int r = 0; r < 3; r++) { double s = P[r][0] * p.x; s += P[r][1] * p.y; s += P[r][2] * p.z; s += P[r][3] * 1.0; uvh[r] = s; }
    double u = uvh[0] / uvh[2], v = uvh[1] / uvh[2];
    // int(double): x86 cvttsd2si; NaN/out-of-range -> INT_MIN
    auto toint = [](double a) -> int { if (!(a > -2147483649.0 && a < 2147483648.0)) return std::numeric_limits<int>::min(); return (int)a; };
    int x = toint(u), y = toint(v);
    double dis_diff = 0;
    bool valid = std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z);
    bool inrange = (bx0 < x && x < bx2) && (by1 < y && y < by3);
    if (inrange && (x < 0 || y < 0 || x >= c.cam.cols || y >= c.cam.rows)) inrange = false;  // (reference would read out of bounds)
    if (inrange) {
      uint16_t d = proj_depth_at(c, x, y);
      if (d == 0) dis_diff = 0;
      else {
        double depth = d / c.cam.scale;
        double ddx = wc.t.x - p.x, ddy = wc.t.y - p.y, ddz = wc.t.z - p.z;
        double dis_cam_point = std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        if (dis_cam_point > depth) dis_diff = 0;
        else { if (!valid) { dis_diff = 0; num_invalid++; } else dis_diff = find_minimal_distance(c, p); }
      }
    } else {
      if (!valid) { dis_diff = 0; num_invalid++; } else dis_diff = find_minimal_distance(c, p);
    }
    double Sigma_inv = 1.0 / c.sigma;
    double ln_P = -0.5 * Sigma_inv * Sigma_inv * dis_diff * dis_diff;
    ln_total_P += ln_P;
  }
  int num_valid = num - num_invalid;
  if (num_valid > 0) return ln_total_P / double(num_valid);
  return -INFINITY;
}
inline void mirror_cloud(const std::vector<P3d> &cloud, const double pl[4], std::vector<V3> &out, bool append) {
  if (!append) out.clear();
  for (const P3d &p : cloud) out.push_back(symmetry_point_of_plane(V3{p.x, p.y, p.z}, pl));
}
// edge error for the (dual) symmetry plane vertex: EdgeSymmetryPlane::computeError :375-396 / Dual :458-481
inline double symmetry_edge_error(SymContext &c, const Plane &pl, int symType) {
  std::vector<V3> sym;
  if (symType == 1) {
    mirror_cloud(*c.cloud, pl.p, sym, false);
  } else {
    double p2[4]; plane_another_vec(pl, p2);
    mirror_cloud(*c.cloud, pl.p, sym, false);
    mirror_cloud(*c.cloud, p2, sym, true);
  }
  return -pointcloud_prob(c, sym);
}
struct SymResult { double prob, init_error, final_error; Plane plane; };
// OptimizeSymmetryPlane / OptimizeSymmetryDualPlane (SymmetrySolver.cpp:217-333): 1 vertex, 1 unary edge, LM(iters)
inline SymResult optimize_symmetry(SymContext &c, const Plane &init, int symType, int iters) {
  Plane est = init;
  std::vector<Plane> stack;
  const int dim = (symType == 1) ? 2 : 3;
  double err = 0;  // e->error()
  auto oplus = [&](Plane &p, const double *u) {
    if (symType == 1) { double v3[3] = {u[0], 0, u[1]}; plane_oplus(p, v3); }   // VertexPlane::oplusImpl :527-532
    else plane_oplus_dual(p, u);                                                 // VertexDualPlane::oplusImpl :544-547
  };
  SymResult R;
  err = symmetry_edge_error(c, est, symType);
  R.init_error = err;
  LmProblem P;
  P.dim = dim;
  P.compute_errors = [&]() { err = symmetry_edge_error(c, est, symType); return err * (1.0 * err); };  // chi2, information = 1
  P.build_system = [&](std::vector<double> &H, std::vector<double> &b) {
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    double J[3]; double add[3] = {0, 0, 0};
    double errorBeforeNumeric = err;
    for (int d = 0; d < dim; d++) {  // base_unary_edge.hpp:94-117
      Plane pp = est; add[d] = delta; oplus(pp, add);
      double e1 = symmetry_edge_error(c, pp, symType);
      Plane pm = est; add[d] = -delta; oplus(pm, add);
      double e2 = symmetry_edge_error(c, pm, symType);
      add[d] = 0.0;
      J[d] = scalar * (e1 - e2);
    }
    err = errorBeforeNumeric;
    H.assign((size_t)dim * dim, 0.0); b.assign(dim, 0.0);
    for (int i = 0; i < dim; i++) {
      b[i] -= (J[i] * 1.0) * err;
      for (int j = 0; j < dim; j++) H[(size_t)i * dim + j] += (J[i] * 1.0) * J[j];
    }
  };
  P.push = [&]() { stack.push_back(est); };
  P.pop = [&]() { est = stack.back(); stack.pop_back(); };
  P.discard_top = [&]() { stack.pop_back(); };
  P.update = [&](const double *x) { oplus(est, x); };
  lm_optimize(P, iters);
  // NOTE: e->error() is the value of the LAST computeError call, which after a rejected LM trial belongs to the
  // rejected (popped) state -- the reference reads it anyway (SymmetrySolver.cpp:265-268,321-323).
  R.final_error = err;
  R.prob = std::exp(-err);
  R.plane = est;
  return R;
}

// ---------------- the extractor ----------------
static long long g_last_seed_evals[9];
struct ExtractorState { bool plane_set = false; double plane[4] = {0, 0, 1, 0}; bool unstable_voxel_sort = false; };
inline int symmetry_type_of_label(int label) {  // LoadSymmetryPrior :52-79 ; -1 = label unknown
  switch (label) { case 58: return 0; case 59: case 62: case 57: case 66: case 63: case 64: case 41: return 1; case 28: return 2; default: return -1; }
}
inline void copy_pts(double *dst, int cap, const std::vector<P3d> &v) { if (!dst) return; int n = std::min<int>(cap, (int)v.size()); for (int i = 0; i < n; i++) { dst[3 * i] = v[i].x; dst[3 * i + 1] = v[i].y; dst[3 * i + 2] = v[i].z; } }

// EllipsoidExtractor::EstimateLocalEllipsoid (EllipsoidExtractor.cpp:292-493)
inline void estimate_local_ellipsoid(const ExtractorState &st, const EslamParams &prm, const EslamCamera &cam, const uint16_t *depth, size_t pitch,
                                     const double bbox[4], int label, const double Twc7[7], EslamEllipsoidOut &out, EslamStageDump *dump,
                                     long long *n_cost_evals = nullptr) {
  std::memset(&out, 0, sizeof(out));
  out.best_seed = -1;
  const int cap = dump ? dump->cap_points : 0;
  if (!bbox_inside(bbox, cam)) { out.status = ESLAM_OBJ_BBOX_OUTSIDE; return; }
  SE3 Twc = se3_from_vector(Twc7);
  // ---- ExtractPointCloud :86-168
  std::vector<P3d> local; get_pointcloud_in_rect(depth, pitch, bbox, cam, prm.depth_range, local);
  out.n_crop = (int)local.size();
  if (dump) copy_pts(dump->crop_xyz, cap, local);
  std::vector<P3d> local_ds; downsample_only(local, local_ds, prm.voxel_leaf, st.unstable_voxel_sort);
  out.n_voxel = (int)local_ds.size();
  if (dump) copy_pts(dump->voxel_xyz, cap, local_ds);
  M4 Twc_h = se3_to_hom(Twc);
  std::vector<P3d> filtered;  // transformPointCloud + ApplySupportingPlaneFilter :562-578
  for (const P3d &p : local_ds) {
    V3 w = transform_point_hom(Twc_h, V3{p.x, p.y, p.z});
    double dis = plane_distance_to_point(st.plane, w, true);
    if (dis > prm.plane_margin) filtered.push_back(P3d{w.x, w.y, w.z});
  }
  out.n_plane = (int)filtered.size();
  if (dump) copy_pts(dump->world_xyz, cap, filtered);
  if (filtered.size() < 1) { out.status = ESLAM_OBJ_NO_POINTS_AFTER_PLANE; return; }
  V3 center;
  if (!get_center(depth, pitch, bbox, Twc, cam, prm.depth_range, center)) { out.status = ESLAM_OBJ_NO_CENTER; return; }
  if (dump) { dump->center_world[0] = center.x; dump->center_world[1] = center.y; dump->center_world[2] = center.z; }
  // ---- ApplyEuclideanFilter :646-737
  std::vector<P3f> pf(filtered.size());
  for (size_t i = 0; i < filtered.size(); i++) pf[i] = P3f{(float)filtered[i].x, (float)filtered[i].y, (float)filtered[i].z};
  std::vector<std::vector<int>> clusters;
  euclidean_clusters(pf, prm.cluster_tolerance, prm.min_cluster_size, (int)pf.size(), clusters);
  out.n_clusters = (int)clusters.size();
  if (dump && dump->cluster_label) { for (int i = 0; i < std::min<int>(cap, (int)pf.size()); i++) dump->cluster_label[i] = -1; for (size_t c = 0; c < clusters.size(); c++) for (int i : clusters[c]) if (i < cap) dump->cluster_label[i] = (int)c; }
  int sel = -1;
  for (int i = 0; i < (int)clusters.size(); i++) {
    if (clusters.size() == 1) sel = i;
    double mini = 999999;  // getDistanceFromPointToCloud :739-752
    for (int id : clusters[i]) { double dx = center.x - (double)pf[id].x, dy = center.y - (double)pf[id].y, dz = center.z - (double)pf[id].z; double dis = std::sqrt(dx * dx + dy * dy + dz * dz); if (dis < mini) mini = dis; }
    if (mini < prm.center_dis) { sel = i; break; }
  }
  if (sel < 0) { out.status = ESLAM_OBJ_NO_CLUSTER; return; }
  std::vector<P3f> cl(clusters[sel].size());
  for (size_t i = 0; i < cl.size(); i++) cl[i] = pf[clusters[sel][i]];
  out.n_cluster = (int)cl.size();
  // ---- PCA + alignment :308-315
  PCAResult data = process_pca(cl);
  adjust_chirality(data);
  align_z_axis_to_gravity(data, st.plane);
  data.rot = calib_rot_to_ground(data.rot, V3{st.plane[0], st.plane[1], st.plane[2]});
  V3 centerPCA = data.center;
  if (dump) { dump->pca_center[0] = centerPCA.x; dump->pca_center[1] = centerPCA.y; dump->pca_center[2] = centerPCA.z; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dump->pca_rot[i * 3 + j] = data.rot.m[i][j]; dump->pca_eval[0] = data.cov.x; dump->pca_eval[1] = data.cov.y; dump->pca_eval[2] = data.cov.z; }
  // ---- downsample for symmetry :322-324 (cluster floats -> double -> float again: exact)
  std::vector<P3f> objf; voxel_grid(cl, prm.symmetry_grid, objf, st.unstable_voxel_sort);
  std::vector<P3d> obj(objf.size());
  for (size_t i = 0; i < objf.size(); i++) obj[i] = P3d{(double)objf[i].x, (double)objf[i].y, (double)objf[i].z};
  out.n_sym = (int)obj.size();
  // ---- normalised frame :327-347
  V3 rz = normalized(V3{st.plane[0], st.plane[1], st.plane[2]});
  V3 rx = normalized(col(data.rot, 0));
  V3 ry = cross(rz, rx);
  M3 Rwo; setcol(Rwo, 0, rx); setcol(Rwo, 1, ry); setcol(Rwo, 2, rz);
  SE3 Two; Two.r = q_from_R(Rwo); Two.t = centerPCA;   // setRotation/setTranslation: no normalisation
  SE3 Tow = se3_inverse(Two);
  M4 Tow_h = se3_to_hom(Tow);
  std::vector<P3d> objn(obj.size());
  for (size_t i = 0; i < obj.size(); i++) { V3 q = transform_point_hom(Tow_h, V3{obj[i].x, obj[i].y, obj[i].z}); objn[i] = P3d{q.x, q.y, q.z}; }
  if (dump) { copy_pts(dump->sym_xyz, cap, objn); se3_to_vector(Two, dump->T_wo); }
  // ---- symmetry :352-456
  bool runSymmetry = false; int symmetryType = -1; double sym_prob = 1.0;
  if (prm.symmetry_open) { symmetryType = symmetry_type_of_label(label); if (symmetryType > 0) runSymmetry = true; }
  out.symmetry_type = symmetryType > 0 ? symmetryType : 0;
  if (runSymmetry && objn.size() >= 1) {
    SymContext ctx; ctx.depth = depth; ctx.pitch = pitch; ctx.cam = cam; for (int i = 0; i < 4; i++) ctx.bbox[i] = bbox[i];
    ctx.T_oc = se3_mul(Tow, se3_from_vector(Twc7));
    double K[9] = {cam.fx, 0, cam.cx, 0, cam.fy, cam.cy, 0, 0, 1};
    for (int i = 0; i < 9; i++) ctx.K[i] = K[i];
    ctx.cloud = &objn; ctx.sigma = prm.symmetry_sigma;
    ctx.cloud_f.resize(objn.size());
    for (size_t i = 0; i < objn.size(); i++) ctx.cloud_f[i] = P3f{(float)objn[i].x, (float)objn[i].y, (float)objn[i].z};
    // GenerateInitPlanes Symmetry.cpp:44-86
    const double diff_dis = 0.2, diff_angle = M_PI / 180.0 * 5;
    const double start_dis = -diff_dis * 1, start_angle = -diff_angle * 1;
    SymResult best{}; int best_i = -1;
    for (int i = 0; i < 3; i++) for (int m = 0; m < 3; m++) {
      double dis = start_dis + diff_dis * i, angle = start_angle + diff_angle * m;
      Plane init = plane_from_dis_angle_trans(dis, angle, 0);
      const long long evals_before = ctx.n_cost_evals;
      SymResult r = optimize_symmetry(ctx, init, symmetryType, prm.symmetry_iters);
      int s = i * 3 + m;
      g_last_seed_evals[s] = ctx.n_cost_evals - evals_before;   // diagnostics: work per seed of the last object (tools/seed_balance.py)
      if (dump) { dump->seed_prob[s] = r.prob; dump->seed_init_err[s] = r.init_error; dump->seed_final_err[s] = r.final_error; for (int k = 0; k < 4; k++) dump->seed_plane[s][k] = r.plane.p[k]; dump->seed_dual_dis[s] = r.plane.dual_dis; }
      // sort(desc by prob) of 9 elements = insertion sort = stable: first maximum wins (Symmetry.cpp:122-126)
      if (best_i < 0 || r.prob > best.prob) { best = r; best_i = s; }
    }
    if (n_cost_evals) *n_cost_evals += ctx.n_cost_evals;
    out.best_seed = best_i; out.symmetry_result = 1; sym_prob = best.prob;
    double p1[4], p2[4];
    for (int k = 0; k < 4; k++) p1[k] = best.plane.p[k];
    plane_another_vec(best.plane, p2);
    { double w1[4]; for (int k = 0; k < 4; k++) w1[k] = p1[k]; plane_transform(w1, Two); for (int k = 0; k < 4; k++) out.sym_plane_world[k] = w1[k]; }
    if (symmetryType == 2) { double w2[4]; for (int k = 0; k < 4; k++) w2[k] = p2[k]; plane_transform(w2, Two); for (int k = 0; k < 4; k++) out.sym_plane2_world[k] = w2[k]; }
    out.sym_center_world[0] = centerPCA.x; out.sym_center_world[1] = centerPCA.y; out.sym_center_world[2] = centerPCA.z;
    // completion :399-446
    std::vector<V3> sym1; mirror_cloud(objn, p1, sym1, false);
    if (symmetryType == 2) {
      std::vector<V3> s21; mirror_cloud(objn, p2, s21, false);
      std::vector<P3d> sym1d(sym1.size()); for (size_t i = 0; i < sym1.size(); i++) sym1d[i] = P3d{sym1[i].x, sym1[i].y, sym1[i].z};
      std::vector<V3> s22; mirror_cloud(sym1d, p2, s22, false);
      for (V3 &v : s21) sym1.push_back(v);
      for (V3 &v : s22) sym1.push_back(v);
    }
    for (V3 &v : sym1) objn.push_back(P3d{v.x, v.y, v.z});
    V3 total{0, 0, 0};  // GetPointcloudCenter Geometry.cpp:143-158
    for (const P3d &p : objn) { total.x += p.x; total.y += p.y; total.z += p.z; }
    double nn = double(objn.size());
    V3 centerCombined{total.x / nn, total.y / nn, total.z / nn};
    V3 cw = transform_point_hom(se3_to_hom(Two), centerCombined);
    out.sym_center_world[0] = cw.x; out.sym_center_world[1] = cw.y; out.sym_center_world[2] = cw.z;
    M3 Rom; V3 c0{p1[0], p1[1], p1[2]}; double n0 = norm(c0); c0 = V3{c0.x / n0, c0.y / n0, c0.z / n0};
    setcol(Rom, 0, c0); setcol(Rom, 2, V3{0, 0, 1}); setcol(Rom, 1, cross(V3{0, 0, 1}, c0));
    SE3 Tom; Tom.t = centerCombined; Tom.r = q_from_R(Rom);
    SE3 Tmo = se3_inverse(Tom);
    M4 Tmo_h = se3_to_hom(Tmo);
    for (P3d &p : objn) { V3 q = transform_point_hom(Tmo_h, V3{p.x, p.y, p.z}); p = P3d{q.x, q.y, q.z}; }
    Two = se3_mul(Two, Tom);
  }
  out.n_final = (int)objn.size();
  if (dump) copy_pts(dump->final_xyz, cap, objn);
  // ---- ProcessPCANormalized :495-531 + ConstructEllipsoid :270-290
  double mxx = 0, mxy = 0, mxz = 0;
  for (const P3d &p : objn) { if (std::fabs(p.x) > mxx) mxx = std::fabs(p.x); if (std::fabs(p.y) > mxy) mxy = std::fabs(p.y); if (std::fabs(p.z) > mxz) mxz = std::fabs(p.z); }
  Q4 qI = q_from_R(m3_identity());
  double ev[10] = {0, 0, 0, qI.x, qI.y, qI.z, qI.w, mxx, mxy, mxz};
  Ellipsoid e0 = ell_from_vector(ev);
  Ellipsoid eg = ell_transform_from(e0, Two);                       // :465
  Ellipsoid el = ell_transform_from(eg, se3_inverse(se3_from_vector(Twc7)));  // :468-469
  ell_to_vector(el, out.ellipsoid);
  out.prob = runSymmetry ? sym_prob : 1.0;
  out.status = ESLAM_OBJ_OK;
}

}  // namespace orc

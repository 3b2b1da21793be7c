// graph_dev.cuh -- device code and workspace shared by the mapping-mode (edges.cu) and SLAM-mode (slam.cu) graph optimisers.
#pragma once
#include "common.cuh"
#include "dmath.cuh"

namespace eslam {

struct GraphWorkspace {
  int n_edges = 0, n_frames = 0, n_objects = 0;
  DevBuf type, frame, object, meas, info, Tcw, K, ell, ell_trial, ell_backup;
  DevBuf res, jac, chi2;
  DevBuf obj_start, obj_edges;   // CSR: edges of each object in graph order
  DevBuf H, b, x, xtmp, okflag, lm;  // H [No*81], b [No*9], x [No*9]
  PinBuf pin;
};

struct LmState {  // lives in device memory, mirrored to pinned host memory after every decision
  double lambda, ni, currentChi, tempChi, rho, iniChi;
  int accepted, all_ok, pad0, pad1;
};

__device__ __forceinline__ int edge_dim(int t) { return t == ESLAM_EDGE_PROJ ? 4 : (t == ESLAM_EDGE_9DOF ? 9 : 1); }

static __device__ void edge_error(int type, const se3 &Tcw, const ellipsoid &obj, const double *meas, const double *K, double *err) {
  if (type == ESLAM_EDGE_PROJ) {
    double proj[4];
    ell_bbox(obj, Tcw, K, proj);
#pragma unroll
    for (int i = 0; i < 4; i++) { err[i] = 0; if (meas[i] >= 5) err[i] = proj[i] - meas[i]; }
  } else if (type == ESLAM_EDGE_9DOF) {
    se3 Twc = se3_inv(Tcw);
    ellipsoid m = ell_load(meas);
    ellipsoid g = m;
    g.pose = se3_mul(Twc, m.pose);  // transform_from
    ell_min_log_error(obj, g, err);
  } else {
    mat3 R = q_to_R(obj.pose.r);
    d3 z = m3_col(R, 2);
    d3 n = mk3(meas[0], meas[1], meas[2]);
    double c = dot3(z, n) / norm3(z) / norm3(n);
    if (c > 1) c = c - 0.0001;
    else if (c < -1) c = c + 0.0001;
    err[0] = acos(c) - 0;
  }
}

// sequential (graph-order) sum of the per-edge chi2 by one warp: lanes fetch 32 values, lane 0 adds them in order.
static __device__ double warp_ordered_sum(const double *v, int n) {
  int lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int base = 0; base < n; base += 32) {
    double x = (base + lane < n) ? v[base + lane] : 0.0;
    int cnt = min(32, n - base);
    for (int k = 0; k < cnt; k++) {
      double xk = __shfl_sync(0xffffffffu, x, k);
      acc += xk;
    }
  }
  return acc;  // every lane holds the same value
}

// one CTA (96 threads) per object: thread t < 81 owns H(i,j), t in [81,90) owns b(i); edges visited in graph order so the
// fp64 accumulation order equals g2o's (activeEdges sorted by internalId, optimizable_graph.cpp:273).
static __global__ void __launch_bounds__(96) build_system_kernel(const int *__restrict__ obj_start, const int *__restrict__ obj_edges, const int *__restrict__ type,
                                                          const double *__restrict__ info, const double *__restrict__ res, const double *__restrict__ jac,
                                                          double *__restrict__ H, double *__restrict__ b) {
  int o = blockIdx.x, t = threadIdx.x;
  if (t >= 90) return;
  int s = obj_start[o], epos = obj_start[o + 1];
  double acc = 0.0;
  if (t < 81) {
    int i = t / 9, j = t % 9;
    for (int k = s; k < epos; k++) {
      int e = obj_edges[k];
      int D = edge_dim(type[e]);
      double w = info[e];
      const double *J = jac + (size_t)e * 81;
      double h = 0;
      for (int r = 0; r < D; r++) { double pr = (J[r * 9 + i] * w) * J[r * 9 + j]; h = (r == 0) ? pr : h + pr; }
      acc += h;
    }
    H[(size_t)o * 81 + t] = acc;
  } else {
    int i = t - 81;
    for (int k = s; k < epos; k++) {
      int e = obj_edges[k];
      int ty = type[e];
      int D = edge_dim(ty);
      double w = info[e];
      const double *J = jac + (size_t)e * 81;
      const double *r_ = res + (size_t)e * 9;
      double bi = 0;
      for (int r = 0; r < D; r++) {
        double pr = (ty == ESLAM_EDGE_GRAVITY) ? ((J[r * 9 + i] * w) * r_[r]) : (J[r * 9 + i] * (-(w * r_[r])));
        bi = (r == 0) ? pr : bi + pr;
      }
      if (ty == ESLAM_EDGE_GRAVITY) acc -= bi; else acc += bi;
    }
    b[(size_t)o * 9 + i] = acc;
  }
}

}  // namespace eslam

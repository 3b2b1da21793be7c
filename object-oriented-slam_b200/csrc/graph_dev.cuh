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
  // mapping-mode optimize(iters) as one replayable launch sequence (device-resident control flow): valid for this key only
  cudaGraphExec_t lm_exec = nullptr;
  struct LmKey { int iters, n_edges, n_objects; const void *p[14]; } lm_key = {};
  const void *lm_more[5] = {};
};

struct LmState {  // lives in device memory; mirrored to pinned host memory (SLAM mode: after every decision, mapping mode: once at the end)
  double lambda, ni, currentChi, tempChi, rho, iniChi;
  int accepted, all_ok, pad0, pad1;
  // device-resident control flow of the mapping-mode LM (edges.cu): every launch of the fixed sequence looks at these and exits when it has nothing to do
  int trial_active, terminated, qmax, nBad, iterations, trials_total;
  double chi2_initial, chi2_final;
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

// sequential (graph-order) sum acc + v[0] + v[1] + ... by one warp: the lanes stage 256 values in shared memory (coalesced, the
// next chunk already in flight), lane 0 adds them strictly in order.  Every lane returns the same value.
static __device__ double warp_ordered_sum(const double *v, int n, double acc = 0.0) {
  __shared__ double stage[256];
  int lane = threadIdx.x & 31;
  double nxt[8];
#pragma unroll
  for (int j = 0; j < 8; j++) nxt[j] = (j * 32 + lane < n) ? v[j * 32 + lane] : 0.0;
  for (int base = 0; base < n; base += 256) {
#pragma unroll
    for (int j = 0; j < 8; j++) stage[j * 32 + lane] = nxt[j];
    __syncwarp();
    int nb = base + 256;
#pragma unroll
    for (int j = 0; j < 8; j++) nxt[j] = (nb + j * 32 + lane < n) ? v[nb + j * 32 + lane] : 0.0;
    if (lane == 0) {
      int cnt = min(256, n - base);
      for (int k = 0; k < cnt; k++) acc += stage[k];
    }
    __syncwarp();
  }
  return __shfl_sync(0xffffffffu, acc, 0);
}

// one CTA (96 threads) per object: thread t < 81 owns H(i,j), t in [81,90) owns b(i); edges visited in graph order so the
// fp64 accumulation order equals g2o's (activeEdges sorted by internalId, optimizable_graph.cpp:273).  The Jacobians / residuals of
// 16 edges at a time are staged in shared memory with coalesced, independent loads (the scattered per-edge reads were the whole cost);
// rows >= dim of every Jacobian are stored as zeros, so all 9 rows are summed (adding +0.0 is exact).
static __global__ void __launch_bounds__(96) build_system_kernel(const int *__restrict__ obj_start, const int *__restrict__ obj_edges, const int *__restrict__ type,
                                                          const double *__restrict__ info, const double *__restrict__ res, const double *__restrict__ jac,
                                                          double *__restrict__ H, double *__restrict__ b, const LmState *ctl) {
  if (ctl && ctl->terminated) return;   // mapping-mode sequence: the optimisation has already stopped
  constexpr int CH = 16;
  __shared__ double Js[CH][81], rs[CH][9], ws[CH];
  __shared__ int ts[CH];
  int o = blockIdx.x, t = threadIdx.x;
  int s = obj_start[o], epos = obj_start[o + 1];
  double acc = 0.0;
  const int i = (t < 81) ? t / 9 : t - 81, j = t % 9;
  for (int c0 = s; c0 < epos; c0 += CH) {
    int cnt = min(CH, epos - c0);
    for (int q = t; q < cnt * 81; q += 96) { int k = q / 81, m = q % 81; Js[k][m] = jac[(size_t)obj_edges[c0 + k] * 81 + m]; }
    for (int q = t; q < cnt * 9; q += 96) { int k = q / 9, m = q % 9; rs[k][m] = res[(size_t)obj_edges[c0 + k] * 9 + m]; }
    if (t < cnt) { int e = obj_edges[c0 + t]; ws[t] = info[e]; ts[t] = type[e]; }
    __syncthreads();
    if (t < 81) {
      for (int k = 0; k < cnt; k++) {
        double w = ws[k];
        double h = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) { double pr = (Js[k][r * 9 + i] * w) * Js[k][r * 9 + j]; h = (r == 0) ? pr : h + pr; }
        acc += h;
      }
    } else if (t < 90) {
      for (int k = 0; k < cnt; k++) {
        double w = ws[k];
        bool grav = ts[k] == ESLAM_EDGE_GRAVITY;
        double bi = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) {
          double pr = grav ? ((Js[k][r * 9 + i] * w) * rs[k][r]) : (Js[k][r * 9 + i] * (-(w * rs[k][r])));
          bi = (r == 0) ? pr : bi + pr;
        }
        if (grav) acc -= bi; else acc += bi;
      }
    }
    __syncthreads();
  }
  if (t < 81) H[(size_t)o * 81 + t] = acc;
  else if (t < 90) b[(size_t)o * 9 + i] = acc;
}

}  // namespace eslam

// slam.cu -- SLAM mode of the object graph optimisation: free camera poses + odometry edges.
//
// Replaces, on the device:
//   Optimizer::GlobalObjectGraphOptimization with bSLAM_mode = true      src/core/Optimizer.cpp:126-159,290-291
//   VertexSE3Expmap::oplusImpl  (T <- exp(dx) * T)                       Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:73-76
//   EdgeSE3Expmap::computeError (log(C * v1 * v2^-1))                    types_six_dof_expmap.h:89-98
//   numeric Jacobians of BOTH vertices of every binary edge              Thirdparty/g2o/g2o/core/base_binary_edge.hpp:131-205
//   quadratic form incl. the pose/object off-diagonal block              base_binary_edge.hpp:55-120
//   OptimizationAlgorithmLevenberg::solve + LinearSolverDense            optimization_algorithm_levenberg.cpp:61-164, linear_solver_dense.h:65-113
//
// The reference solves the full (6(F-1)+9No)^2 system with a dense pivoted LDLT.  Here the structure is used instead:
//   H = [ Hpp  W  ]   Hpp block tridiagonal (odometry chain, 6x6 blocks), Hoo block diagonal (9x9), W = Hpo dense-ish.
//       [ W^T  Hoo]
// (Hpp + lambda I) is factorised by a block-tridiagonal Cholesky, X = (Hpp + lambda I)^-1 [W | b_p] by per-column
// substitution, the reduced object system S = Hoo + lambda I - W^T X (9No x 9No, dense) by an in-place Cholesky, then
// x_p is back-substituted.  Mathematically the same solution as the reference's dense LDLT (both are exact direct
// solves of an SPD system); accumulation of every H/b entry follows g2o's edge order.
//
// Layout (HBM): Jp f64[E*54] (dim x 6), Jo f64[E*81] (dim x 9); odometry ores f64[NP*6], Ji/Jj f64[NP*36];
// Dp/Op f64[NP*36] (diagonal / (k,k+1) blocks of Hpp), W/X f64[6NP][ncols] row-major with ncols = 9No+1 (last column
// b_p), Lf/Mf f64[NP*36] (factor), S f64[n_o*n_o] column-major lower triangle, NP = F-1 free poses.
#include "common.cuh"
#include "dmath.cuh"
#include "internal.h"
#include "graph_dev.cuh"

namespace eslam {

struct SlamWorkspace {
  int NP = 0, No = 0, E = 0, F = 0, ncols = 0;
  double odom_info = 1.0;
  DevBuf T, T_trial, odo, ores, oJi, oJj, ochi2, Jp;
  DevBuf frm_start, frm_edges;
  DevBuf Dp, Op, bp, W, X, Lf, Mf, S, rhs, xo_tmp, xp_tmp, xp, xo;
};

// ---------------------------------------------------------------------------------------------------------------- edges
// thread = (edge, slot): slots 0..5 pose column, 6..14 object column, 15 residual + chi2.  with_jac == 0: slot 15 only.
__global__ void __launch_bounds__(256) slam_edge_eval_kernel(int n_edges, const int *__restrict__ type, const int *__restrict__ frame,
                                                             const int *__restrict__ object, const double *__restrict__ meas,
                                                             const double *__restrict__ info, const double *__restrict__ Tcw,
                                                             const double *__restrict__ Kc, const double *__restrict__ ell, double *__restrict__ res,
                                                             double *__restrict__ Jp, double *__restrict__ Jo, double *__restrict__ chi2, int with_jac) {
  const int slots = with_jac ? 16 : 1;
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int e = (int)(gid / slots);
  int slot = with_jac ? (int)(gid % slots) : 15;
  if (e >= n_edges) return;
  int t = type[e];
  int D = edge_dim(t);
  double K[9];
#pragma unroll
  for (int i = 0; i < 9; i++) K[i] = Kc[i];
  int f = (t == ESLAM_EDGE_GRAVITY) ? 0 : frame[e];
  se3 T = se3_load(Tcw + 7 * f);
  ellipsoid obj = ell_load(ell + 10 * object[e]);
  const double *m = meas + 10 * (size_t)e;
  const double delta = 1e-9;
  const double scalar = 1.0 / (2 * delta);
  if (slot == 15) {
    double r[9];
#pragma unroll
    for (int i = 0; i < 9; i++) r[i] = 0;
    edge_error(t, T, obj, m, K, r);
    double w = info[e];
    double c = 0;
    for (int i = 0; i < D; i++) c += r[i] * (w * r[i]);
    for (int i = 0; i < 9; i++) res[(size_t)e * 9 + i] = i < D ? r[i] : 0.0;
    chi2[e] = c;
  } else if (slot >= 6) {
    int c = slot - 6;
    double add[9];
#pragma unroll
    for (int i = 0; i < 9; i++) add[i] = 0;
    double e1[9], e2[9];
    add[c] = delta;
    edge_error(t, T, ell_oplus(obj, add), m, K, e1);
    add[c] = -delta;
    edge_error(t, T, ell_oplus(obj, add), m, K, e2);
    for (int r = 0; r < 9; r++) Jo[(size_t)e * 81 + r * 9 + c] = r < D ? scalar * (e1[r] - e2[r]) : 0.0;
  } else {
    int c = slot;
    if (t == ESLAM_EDGE_GRAVITY || f == 0) {  // unary edge / fixed first pose: no pose column
      for (int r = 0; r < 9; r++) Jp[(size_t)e * 54 + r * 6 + c] = 0.0;
      return;
    }
    double add[6];
#pragma unroll
    for (int i = 0; i < 6; i++) add[i] = 0;
    double e1[9], e2[9];
    add[c] = delta;
    edge_error(t, se3_mul(se3_exp(add), T), obj, m, K, e1);
    add[c] = -delta;
    edge_error(t, se3_mul(se3_exp(add), T), obj, m, K, e2);
    for (int r = 0; r < 9; r++) Jp[(size_t)e * 54 + r * 6 + c] = r < D ? scalar * (e1[r] - e2[r]) : 0.0;
  }
}

__device__ __forceinline__ void odom_err(const se3 &C, const se3 &v1, const se3 &v2, double *err) {
  se3 e = se3_mul(se3_mul(C, v1), se3_inv(v2));
  se3_log(e, err);
}

// thread = (odometry edge k, slot): slots 0..5 column of vertex i (frame k), 6..11 column of vertex j (frame k+1), 12 residual
__global__ void __launch_bounds__(208) slam_odom_eval_kernel(int NP, const double *__restrict__ odo, const double *__restrict__ Tcw, double w,
                                                             double *__restrict__ ores, double *__restrict__ Ji, double *__restrict__ Jj,
                                                             double *__restrict__ ochi2, int with_jac) {
  const int slots = with_jac ? 13 : 1;
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int k = (int)(gid / slots);
  int slot = with_jac ? (int)(gid % slots) : 12;
  if (k >= NP) return;
  se3 C = se3_load(odo + 7 * (size_t)k);
  se3 v1 = se3_load(Tcw + 7 * (size_t)k), v2 = se3_load(Tcw + 7 * (size_t)(k + 1));
  const double delta = 1e-9;
  const double scalar = 1.0 / (2 * delta);
  if (slot == 12) {
    double r[6];
    odom_err(C, v1, v2, r);
    double c = 0;
    for (int i = 0; i < 6; i++) { c += r[i] * (w * r[i]); ores[(size_t)k * 6 + i] = r[i]; }
    ochi2[k] = c;
    return;
  }
  double add[6];
#pragma unroll
  for (int i = 0; i < 6; i++) add[i] = 0;
  double e1[6], e2[6];
  if (slot < 6) {
    if (k == 0) { for (int r = 0; r < 6; r++) Ji[(size_t)k * 36 + r * 6 + slot] = 0.0; return; }  // frame 0 is fixed
    add[slot] = delta;  odom_err(C, se3_mul(se3_exp(add), v1), v2, e1);
    add[slot] = -delta; odom_err(C, se3_mul(se3_exp(add), v1), v2, e2);
    for (int r = 0; r < 6; r++) Ji[(size_t)k * 36 + r * 6 + slot] = scalar * (e1[r] - e2[r]);
  } else {
    int c = slot - 6;
    add[c] = delta;  odom_err(C, v1, se3_mul(se3_exp(add), v2), e1);
    add[c] = -delta; odom_err(C, v1, se3_mul(se3_exp(add), v2), e2);
    for (int r = 0; r < 6; r++) Jj[(size_t)k * 36 + r * 6 + c] = scalar * (e1[r] - e2[r]);
  }
}

// ------------------------------------------------------------------------------------------------------- system assembly
// one CTA per free pose p (frame f = p+1).  Contributions are added in g2o's edge order: odometry edge f-1 (as vertex j),
// odometry edge f (as vertex i), then the frame's ellipsoid edges in graph order.  Rows [6p, 6p+6) of W belong to this CTA.
__global__ void __launch_bounds__(64) slam_build_pose_kernel(int NP, int ncols, const int *__restrict__ frm_start, const int *__restrict__ frm_edges,
                                                             const int *__restrict__ type, const int *__restrict__ object, const double *__restrict__ info,
                                                             const double *__restrict__ res, const double *__restrict__ Jp, const double *__restrict__ Jo,
                                                             double w_odo, const double *__restrict__ ores, const double *__restrict__ Ji,
                                                             const double *__restrict__ Jj, double *__restrict__ Dp, double *__restrict__ Op,
                                                             double *__restrict__ bp, double *__restrict__ W) {
  int p = blockIdx.x, t = threadIdx.x;
  int f = p + 1;
  int s = frm_start[f], epos = frm_start[f + 1];
  bool has_next = (p + 1 < NP);
  if (t < 36) {
    int i = t / 6, j = t % 6;
    double acc = 0.0, h = 0;
    const double *A = Jj + (size_t)p * 36;
    for (int r = 0; r < 6; r++) { double pr = (A[r * 6 + i] * w_odo) * A[r * 6 + j]; h = (r == 0) ? pr : h + pr; }
    acc += h;
    double off = 0.0;
    if (has_next) {
      const double *Ai = Ji + (size_t)(p + 1) * 36, *Bj = Jj + (size_t)(p + 1) * 36;
      for (int r = 0; r < 6; r++) { double pr = (Ai[r * 6 + i] * w_odo) * Ai[r * 6 + j]; h = (r == 0) ? pr : h + pr; }
      acc += h;
      for (int r = 0; r < 6; r++) { double pr = (Ai[r * 6 + i] * w_odo) * Bj[r * 6 + j]; h = (r == 0) ? pr : h + pr; }
      off += h;
    }
    for (int k = s; k < epos; k++) {
      int e = frm_edges[k];
      int D = edge_dim(type[e]);
      double w = info[e];
      const double *J = Jp + (size_t)e * 54;
      for (int r = 0; r < D; r++) { double pr = (J[r * 6 + i] * w) * J[r * 6 + j]; h = (r == 0) ? pr : h + pr; }
      acc += h;
    }
    Dp[(size_t)p * 36 + t] = acc;
    Op[(size_t)p * 36 + t] = off;
  } else if (t < 42) {
    int i = t - 36;
    double acc = 0.0, bi = 0;
    {
      const double *A = Jj + (size_t)p * 36; const double *e = ores + (size_t)p * 6;
      for (int r = 0; r < 6; r++) { double pr = A[r * 6 + i] * (-(w_odo * e[r])); bi = (r == 0) ? pr : bi + pr; }
      acc += bi;
    }
    if (has_next) {
      const double *A = Ji + (size_t)(p + 1) * 36; const double *e = ores + (size_t)(p + 1) * 6;
      for (int r = 0; r < 6; r++) { double pr = A[r * 6 + i] * (-(w_odo * e[r])); bi = (r == 0) ? pr : bi + pr; }
      acc += bi;
    }
    for (int k = s; k < epos; k++) {
      int e = frm_edges[k];
      int D = edge_dim(type[e]);
      double w = info[e];
      const double *J = Jp + (size_t)e * 54; const double *r_ = res + (size_t)e * 9;
      for (int r = 0; r < D; r++) { double pr = J[r * 6 + i] * (-(w * r_[r])); bi = (r == 0) ? pr : bi + pr; }
      acc += bi;
    }
    bp[(size_t)p * 6 + i] = acc;
    W[(size_t)(6 * p + i) * ncols + (ncols - 1)] = acc;
  }
  if (t < 54) {  // pose/object off-diagonal blocks: W(6p+i, 9o+j) += sum_r (Jp(r,i) w) Jo(r,j)
    int i = t / 9, j = t % 9;
    for (int k = s; k < epos; k++) {
      int e = frm_edges[k];
      int D = edge_dim(type[e]);
      double w = info[e];
      const double *A = Jp + (size_t)e * 54, *B = Jo + (size_t)e * 81;
      double h = 0;
      for (int r = 0; r < D; r++) { double pr = (A[r * 6 + i] * w) * B[r * 9 + j]; h = (r == 0) ? pr : h + pr; }
      W[(size_t)(6 * p + i) * ncols + 9 * object[e] + j] += h;
    }
  }
}

// ordered sum of two arrays (odometry edges first, then the ellipsoid edges: g2o's active edge order)
__device__ double warp_ordered_sum2(const double *v1, int n1, const double *v2, int n2) {
  return warp_ordered_sum(v2, n2, warp_ordered_sum(v1, n1));
}

__global__ void slam_lm_begin_kernel(LmState *st, const double *ochi2, int NP, const double *chi2, int n_edges, const double *Dp, const double *Hoo, int n_obj,
                                     int iteration) {
  double chi = warp_ordered_sum2(ochi2, NP, chi2, n_edges);
  if (threadIdx.x == 0) {
    st->currentChi = chi; st->tempChi = chi; st->iniChi = chi; st->rho = 0;
    if (iteration == 0) {
      double mx = 0;
      for (int p = 0; p < NP; p++) for (int j = 0; j < 6; j++) mx = fmax(fabs(Dp[(size_t)p * 36 + j * 7]), mx);
      for (int o = 0; o < n_obj; o++) for (int j = 0; j < 9; j++) mx = fmax(fabs(Hoo[(size_t)o * 81 + j * 10]), mx);
      st->lambda = 1e-5 * mx; st->ni = 2;
    }
    st->all_ok = 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------- solver
// block-tridiagonal Cholesky of Hpp + lambda I, one warp.  L_k L_k^T = A_k - M_k M_k^T,  M_k = C_k L_{k-1}^-T,
// C_k = block(k, k-1) = Op[k-1]^T.  The chain is sequential by nature (499 dependent 6x6 steps): the next step's blocks are
// prefetched into registers while the current one is factorised, the 6x6 Cholesky runs in registers of lane 0.
__global__ void __launch_bounds__(32) slam_factor_kernel(LmState *st, int NP, const double *__restrict__ Dp, const double *__restrict__ Op,
                                                         double *__restrict__ Lf, double *__restrict__ Mf) {
  __shared__ double L[36], M[36], A[36];
  int lane = threadIdx.x;
  double lambda = st->lambda;
  int bad = 0;
  for (int t = lane; t < 36; t += 32) { L[t] = 0; M[t] = 0; }
  // lane t holds entries t and t+32 (t < 4) of the diagonal block, and of the transposed off-diagonal block, of the next step
  int t0 = lane, t1 = lane + 32;
  double d0 = Dp[t0], d1 = (t1 < 36) ? Dp[t1] : 0.0, c0 = 0, c1 = 0;
  __syncwarp();
  for (int k = 0; k < NP; k++) {
    double a0 = d0, a1 = d1, cc0 = c0, cc1 = c1;
    if (k + 1 < NP) {  // prefetch step k+1: D_{k+1} and C_{k+1} = Op[k]^T  (C(i,j) = Op[k](j,i))
      d0 = Dp[(size_t)(k + 1) * 36 + t0]; if (t1 < 36) d1 = Dp[(size_t)(k + 1) * 36 + t1];
      c0 = Op[(size_t)k * 36 + (t0 % 6) * 6 + t0 / 6]; if (t1 < 36) c1 = Op[(size_t)k * 36 + (t1 % 6) * 6 + t1 / 6];
    }
    if (k > 0) {  // stage C_k in A[], then rows of M = C L^-T by lanes 0..5
      A[t0] = cc0; if (t1 < 36) A[t1] = cc1;
      __syncwarp();
      if (lane < 6) {
        double y[6];
#pragma unroll
        for (int j = 0; j < 6; j++) {
          double v = A[lane * 6 + j];
#pragma unroll
          for (int m = 0; m < j; m++) v -= L[j * 6 + m] * y[m];
          y[j] = v / L[j * 6 + j];
        }
#pragma unroll
        for (int j = 0; j < 6; j++) M[lane * 6 + j] = y[j];
      }
      __syncwarp();
    }
    {
      int i = t0 / 6, j = t0 % 6;
      double v = a0; if (i == j) v += lambda;
      if (k > 0) {
#pragma unroll
        for (int m = 0; m < 6; m++) v -= M[i * 6 + m] * M[j * 6 + m];
      }
      A[t0] = v;
      if (t1 < 36) {
        i = t1 / 6; j = t1 % 6;
        v = a1; if (i == j) v += lambda;
        if (k > 0) {
#pragma unroll
          for (int m = 0; m < 6; m++) v -= M[i * 6 + m] * M[j * 6 + m];
        }
        A[t1] = v;
      }
    }
    __syncwarp();
    if (lane == 0) {  // 6x6 Cholesky (lower) in registers
      double l[6][6];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
#pragma unroll
        for (int m = 0; m < j; m++) d -= l[j][m] * l[j][m];
        if (!(d > 0)) { bad = 1; d = 1; }
        double ljj = sqrt(d);
        l[j][j] = ljj;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
          double v = A[i * 6 + j];
#pragma unroll
          for (int m = 0; m < j; m++) v -= l[i][m] * l[j][m];
          l[i][j] = v / ljj;
        }
      }
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) L[i * 6 + j] = (j <= i) ? l[i][j] : 0.0;
    }
    __syncwarp();
    Lf[(size_t)k * 36 + t0] = L[t0]; Mf[(size_t)k * 36 + t0] = M[t0];
    if (t1 < 36) { Lf[(size_t)k * 36 + t1] = L[t1]; Mf[(size_t)k * 36 + t1] = M[t1]; }
    __syncwarp();
  }
  if (lane == 0 && bad) st->all_ok = 0;
}

// X = (Hpp + lambda I)^-1 W, one thread per column (coalesced across columns), sequential along the chain.  The factor blocks of
// the next step are staged in shared memory (double buffer) and the next W/X rows are prefetched while the current step computes.
__global__ void __launch_bounds__(64) slam_subst_kernel(int NP, int ncols, const double *__restrict__ Lf, const double *__restrict__ Mf,
                                                        const double *__restrict__ W, double *__restrict__ X) {
  __shared__ double Ls[2][36], Ms[2][36];
  int tid = threadIdx.x;
  int c = blockIdx.x * blockDim.x + tid;
  bool act = c < ncols;
  int cc = act ? c : ncols - 1;
  double y[6], prev[6], w[6], wn[6];
  if (tid < 36) { Ls[0][tid] = Lf[tid]; Ms[0][tid] = Mf[tid]; }
#pragma unroll
  for (int i = 0; i < 6; i++) wn[i] = W[(size_t)i * ncols + cc];
  __syncthreads();
  for (int k = 0; k < NP; k++) {
    const double *L = Ls[k & 1], *M = Ms[k & 1];
#pragma unroll
    for (int i = 0; i < 6; i++) w[i] = wn[i];
    if (k + 1 < NP) {
      if (tid < 36) { Ls[(k + 1) & 1][tid] = Lf[(size_t)(k + 1) * 36 + tid]; Ms[(k + 1) & 1][tid] = Mf[(size_t)(k + 1) * 36 + tid]; }
#pragma unroll
      for (int i = 0; i < 6; i++) wn[i] = W[(size_t)(6 * (k + 1) + i) * ncols + cc];
    }
    if (k > 0) {
#pragma unroll
      for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int m = 0; m < 6; m++) w[i] -= M[i * 6 + m] * prev[m];
      }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      double v = w[i];
#pragma unroll
      for (int m = 0; m < i; m++) v -= L[i * 6 + m] * y[m];
      y[i] = v / L[i * 6 + i];
    }
    if (act) {
#pragma unroll
      for (int i = 0; i < 6; i++) X[(size_t)(6 * k + i) * ncols + c] = y[i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) prev[i] = y[i];
    __syncthreads();
  }
  // backward sweep; buffers: L_k in Ls[k&1], M_{k+1} in Ms[(k+1)&1] (left there by the forward sweep for k = NP-1)
  if (NP > 0) {
#pragma unroll
    for (int i = 0; i < 6; i++) wn[i] = prev[i];   // y_{NP-1} is still in registers
  }
  for (int k = NP - 1; k >= 0; k--) {
    const double *L = Ls[k & 1], *M = Ms[(k + 1) & 1];
#pragma unroll
    for (int i = 0; i < 6; i++) w[i] = wn[i];
    if (k > 0) {  // stage what step k-1 needs: L_{k-1} (slot of L_{k+1}) and M_k (slot of M_{k+2}); neither slot is read by step k
      if (tid < 36) { Ls[(k - 1) & 1][tid] = Lf[(size_t)(k - 1) * 36 + tid]; Ms[k & 1][tid] = Mf[(size_t)k * 36 + tid]; }
#pragma unroll
      for (int i = 0; i < 6; i++) wn[i] = X[(size_t)(6 * (k - 1) + i) * ncols + cc];
    }
    if (k < NP - 1) {
#pragma unroll
      for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int m = 0; m < 6; m++) w[i] -= M[m * 6 + i] * prev[m];
      }
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
      double v = w[i];
#pragma unroll
      for (int m = i + 1; m < 6; m++) v -= L[m * 6 + i] * y[m];
      y[i] = v / L[i * 6 + i];
    }
    if (act) {
#pragma unroll
      for (int i = 0; i < 6; i++) X[(size_t)(6 * k + i) * ncols + c] = y[i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) prev[i] = y[i];
    __syncthreads();
  }
}

// S = Hoo + lambda I - W^T X (n_o x n_o, column-major with leading dimension n_o+1, lower triangle used) and
// rhs = b_o - W^T X[:, last], stored as row n_o of S.  32x32 tile per CTA, 2x2 per thread, K swept in order (deterministic).
__global__ void __launch_bounds__(256) slam_schur_kernel(const LmState *st, int rows, int n_o, int ncols, const double *__restrict__ W, const double *__restrict__ X,
                                                         const double *__restrict__ Hoo, const double *__restrict__ bo, double *__restrict__ S) {
  __shared__ double Ws[16][33], Xs[16][33];
  int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int r0 = 0; r0 < rows; r0 += 16) {
    for (int t = threadIdx.x; t < 16 * 32; t += 256) {
      int r = t / 32, c = t % 32;
      Ws[r][c] = (r0 + r < rows && a0 + c < n_o) ? W[(size_t)(r0 + r) * ncols + a0 + c] : 0.0;
      Xs[r][c] = (r0 + r < rows && b0 + c < ncols) ? X[(size_t)(r0 + r) * ncols + b0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) {
      double wa0 = Ws[r][ty], wa1 = Ws[r][ty + 16], xb0 = Xs[r][tx], xb1 = Xs[r][tx + 16];
      acc[0][0] += wa0 * xb0; acc[0][1] += wa0 * xb1; acc[1][0] += wa1 * xb0; acc[1][1] += wa1 * xb1;
    }
    __syncthreads();
  }
  double lambda = st->lambda;
  for (int u = 0; u < 2; u++)
    for (int v = 0; v < 2; v++) {
      int a = a0 + ty + 16 * u, b = b0 + tx + 16 * v;
      if (a >= n_o || b >= ncols) continue;
      if (b == n_o) { S[(size_t)a * (n_o + 1) + n_o] = bo[a] - acc[u][v]; continue; }
      double hv = 0.0;
      if (a / 9 == b / 9) hv = Hoo[(size_t)(a / 9) * 81 + (a % 9) * 9 + (b % 9)];
      if (a == b) hv += lambda;
      S[(size_t)b * (n_o + 1) + a] = hv - acc[u][v];   // element (a, b) of S, column-major
    }
}

// in-place LDL^T of the lower triangle of S + solve S x = rhs, one CTA of 1024 threads.  S is column-major with leading dimension
// n+1: row n of every column is the right-hand side, so the forward substitution happens inside the elimination.  Columns are left
// unscaled (c_ik = L_ik d_k).  Panels of 8 columns are factorised in shared memory, then applied to the trailing matrix as ONE
// rank-8 update (8 FMAs per global read-modify-write instead of 1; same operation order as eight rank-1 updates).  The backward
// substitution x_k = (w_k - sum_{i>k} c_ik x_i) / d_k is done by a single warp with shuffles.
__global__ void __launch_bounds__(1024) slam_dense_solve_kernel(LmState *st, int n, double *__restrict__ S, double *__restrict__ x) {
  extern __shared__ double sm[];
  const int ld = n + 1, ldp = n + 1;
  double *P = sm, *F = sm + 8 * ldp, *xs = sm + 16 * ldp;
  __shared__ int bad;
  int tid = threadIdx.x;
  int tx = tid % 32, ty = tid / 32;
  if (tid == 0) bad = 0;
  __syncthreads();
  for (int kb = 0; kb < n; kb += 8) {
    const int nb = min(8, n - kb);
    const int m = n + 1 - kb;  // rows kb..n of the panel
    for (int t = tid; t < nb * m; t += 1024) { int kk = t / m, r = t % m; P[kk * ldp + r] = S[(size_t)(kb + kk) * ld + kb + r]; }
    __syncthreads();
    for (int kk = 0; kk < nb; kk++) {
      double d = P[kk * ldp + kk];
      if (!(d > 0)) { if (tid == 0) bad = 1; d = 1; }
      int ncol = nb - 1 - kk;
      for (int t = tid; t < ncol * m; t += 1024) {
        int jj = kk + 1 + t / m, r = t % m;
        if (r >= jj) P[jj * ldp + r] -= P[kk * ldp + r] * (P[kk * ldp + jj] / d);
      }
      __syncthreads();
    }
    for (int t = tid; t < nb * m; t += 1024) {
      int kk = t / m, r = t % m;
      double c = P[kk * ldp + r], d = P[kk * ldp + kk];
      if (!(d > 0)) d = 1;
      if (r >= kk) S[(size_t)(kb + kk) * ld + kb + r] = c;
      F[kk * ldp + r] = c / d;
    }
    __syncthreads();
    for (int j = kb + nb + ty; j < n; j += 32) {
      double f[8];
#pragma unroll
      for (int kk = 0; kk < 8; kk++) f[kk] = (kk < nb) ? F[kk * ldp + (j - kb)] : 0.0;
      double *cj = S + (size_t)j * ld;
      for (int i = j + tx; i <= n; i += 32) {
        double acc = cj[i];
#pragma unroll
        for (int kk = 0; kk < 8; kk++) if (kk < nb) acc -= P[kk * ldp + (i - kb)] * f[kk];
        cj[i] = acc;
      }
    }
    __syncthreads();
  }
  if (ty == 0) {
    for (int k = n - 1; k >= 0; k--) {
      const double *ck = S + (size_t)k * ld;
      double s = 0;
      for (int i = k + 1 + tx; i < n; i += 32) s += ck[i] * xs[i];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tx == 0) {
        double d = ck[k];
        if (!(d > 0)) d = 1;
        xs[k] = (ck[n] - s) / d;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += 1024) x[i] = xs[i];
  if (tid == 0 && bad) st->all_ok = 0;
}

// x_p = X[:, last] - X[:, :n_o] x_o ; one warp per row
__global__ void __launch_bounds__(128) slam_xp_kernel(int rows, int n_o, int ncols, const double *__restrict__ X, const double *__restrict__ xo, double *__restrict__ xp) {
  int r = blockIdx.x * 4 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (r >= rows) return;
  double s = 0;
  for (int c = lane; c < n_o; c += 32) s += X[(size_t)r * ncols + c] * xo[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) xp[r] = X[(size_t)r * ncols + n_o] - s;
}

// commit the solution (only when the factorisations succeeded: a failed solve leaves g2o's x untouched) and apply it
__global__ void slam_update_kernel(const LmState *st, int NP, int n_obj, const double *xp_tmp, const double *xo_tmp, double *xp, double *xo, const double *T,
                                   double *T_trial, const double *ell, double *ell_trial) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = st->all_ok != 0;
  if (i < NP + 1) {
    if (i == 0) { for (int k = 0; k < 7; k++) T_trial[k] = T[k]; }
    else {
      int p = i - 1;
      double u[6];
      for (int k = 0; k < 6; k++) { if (ok) xp[(size_t)p * 6 + k] = xp_tmp[(size_t)p * 6 + k]; u[k] = xp[(size_t)p * 6 + k]; }
      se3 r = se3_mul(se3_exp(u), se3_load(T + 7 * (size_t)i));
      se3_store(r, T_trial + 7 * (size_t)i);
    }
  } else if (i < NP + 1 + n_obj) {
    int o = i - NP - 1;
    double u[9];
    for (int k = 0; k < 9; k++) { if (ok) xo[(size_t)o * 9 + k] = xo_tmp[(size_t)o * 9 + k]; u[k] = xo[(size_t)o * 9 + k]; }
    ellipsoid r = ell_oplus(ell_load(ell + 10 * (size_t)o), u);
    ell_store(r, ell_trial + 10 * (size_t)o);
  }
}

__global__ void slam_decide_kernel(LmState *st, const double *ochi2, int NP, const double *chi2, int n_edges, const double *xp, const double *bp, const double *xo,
                                   const double *bo, int n_obj, double *T, const double *T_trial, double *ell, const double *ell_trial, LmState *host_copy) {
  double tempChi = warp_ordered_sum2(ochi2, NP, chi2, n_edges);
  if (threadIdx.x == 0) {
    if (!st->all_ok) tempChi = DBL_MAX;
    double lambda = st->lambda;
    double scale = 0;
    for (int j = 0; j < 6 * NP; j++) scale += xp[j] * (lambda * xp[j] + bp[j]);
    for (int j = 0; j < 9 * n_obj; j++) scale += xo[j] * (lambda * xo[j] + bo[j]);
    scale += 1e-3;
    double rho = (st->currentChi - tempChi);
    rho /= scale;
    int accepted = 0;
    if (rho > 0 && isfinite(tempChi)) {
      double alpha = 1. - pow((2 * rho - 1), 3.0);
      alpha = fmin(alpha, 2. / 3.);
      double sf = fmax(1. / 3., alpha);
      st->lambda = lambda * sf; st->ni = 2; st->currentChi = tempChi; accepted = 1;
    } else {
      st->lambda = lambda * st->ni; st->ni = st->ni * 2;
    }
    st->tempChi = tempChi; st->rho = rho; st->accepted = accepted; st->all_ok = 1;
    *host_copy = *st;
  }
  __syncwarp();
  int accepted = __shfl_sync(0xffffffffu, (threadIdx.x == 0) ? st->accepted : 0, 0);
  if (accepted) {
    for (int i = threadIdx.x; i < 10 * n_obj; i += 32) ell[i] = ell_trial[i];
    for (int i = threadIdx.x; i < 7 * (NP + 1); i += 32) T[i] = T_trial[i];
  }
}

// ------------------------------------------------------------------------------------------------------------------ host
static int slam_eval(EslamHandle *h, cudaStream_t s, const double *T, const double *ell, bool jac) {
  GraphWorkspace &w = *h->gr;
  SlamWorkspace &q = *h->sl;
  if (w.n_edges > 0) {
    long long threads = (long long)w.n_edges * (jac ? 16 : 1);
    int grid = (int)((threads + 255) / 256);
    slam_edge_eval_kernel<<<grid, 256, 0, s>>>(w.n_edges, w.type.as<int>(), w.frame.as<int>(), w.object.as<int>(), w.meas.as<double>(), w.info.as<double>(), T,
                                               w.K.as<double>(), ell, w.res.as<double>(), q.Jp.as<double>(), w.jac.as<double>(), w.chi2.as<double>(), jac ? 1 : 0);
    ESLAM_LAUNCHED(h);
  }
  if (q.NP > 0) {
    long long threads = (long long)q.NP * (jac ? 13 : 1);
    int grid = (int)((threads + 207) / 208);
    slam_odom_eval_kernel<<<grid, 208, 0, s>>>(q.NP, q.odo.as<double>(), T, q.odom_info, q.ores.as<double>(), q.oJi.as<double>(), q.oJj.as<double>(),
                                               q.ochi2.as<double>(), jac ? 1 : 0);
    ESLAM_LAUNCHED(h);
  }
  ESLAM_CUDA_OK(h, cudaGetLastError());
  return ESLAM_OK;
}

int slam_optimize(EslamHandle *h, const EslamEdgeBatch *g, const double *odom_meas7, double odom_info, int iters, double *ell_inout, double *Tcw_out,
                  EslamLmStats *stats, cudaStream_t s) {
  int rc;
  if (g->n_frames < 1) return ESLAM_ERR_BAD_ARG;
  if ((rc = graph_upload(h, g, ell_inout, s))) return rc;
  if (!h->sl) h->sl = new SlamWorkspace();
  GraphWorkspace &w = *h->gr;
  SlamWorkspace &q = *h->sl;
  const int F = g->n_frames, NP = F - 1, No = g->n_objects, E = g->n_edges;
  const int n_o = 9 * No, ncols = n_o + 1, rows = 6 * NP;
  const size_t solve_smem = (size_t)(17 * (n_o + 1)) * 8;   // panel P[8][n_o+1], F[8][n_o+1], x[n_o]
  if (solve_smem > 220 * 1024) { h->err = "slam mode: more than 183 ellipsoids are not supported by the single-CTA reduced solve"; return ESLAM_ERR_BAD_ARG; }
  ESLAM_CUDA_OK(h, cudaFuncSetAttribute(slam_dense_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_smem));
  q.NP = NP; q.No = No; q.E = E; q.F = F; q.ncols = ncols; q.odom_info = odom_info;
  // odometry measurements: given, or Tcw[k+1] * Tcw[k]^-1 (Optimizer.cpp:141-143)
  std::vector<double> odo((size_t)(NP > 0 ? NP : 1) * 7, 0.0);
  for (int k = 0; k < NP; k++) {
    if (odom_meas7) { for (int i = 0; i < 7; i++) odo[(size_t)k * 7 + i] = odom_meas7[(size_t)k * 7 + i]; }
    else {
      double inv[7];
      eslam_pose_inverse(g->Tcw + 7 * (size_t)k, inv);
      eslam_pose_mul(g->Tcw + 7 * (size_t)(k + 1), inv, &odo[(size_t)k * 7]);
    }
  }
  // CSR of the binary edges per frame, graph order
  std::vector<int> start(F + 2, 0), list((size_t)E + 1);
  for (int k = 0; k < E; k++) if (g->type[k] != ESLAM_EDGE_GRAVITY) start[g->frame[k] + 1]++;
  for (int f = 0; f < F; f++) start[f + 1] += start[f];
  start[F + 1] = start[F];
  {
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int k = 0; k < E; k++) if (g->type[k] != ESLAM_EDGE_GRAVITY) list[fill[g->frame[k]]++] = k;
  }
  ESLAM_CUDA_OK(h, q.T.reserve((size_t)F * 56)); ESLAM_CUDA_OK(h, q.T_trial.reserve((size_t)F * 56));
  ESLAM_CUDA_OK(h, q.odo.reserve(odo.size() * 8));
  ESLAM_CUDA_OK(h, q.ores.reserve((size_t)NP * 48 + 8)); ESLAM_CUDA_OK(h, q.oJi.reserve((size_t)NP * 288 + 8)); ESLAM_CUDA_OK(h, q.oJj.reserve((size_t)NP * 288 + 8));
  ESLAM_CUDA_OK(h, q.ochi2.reserve((size_t)NP * 8 + 8)); ESLAM_CUDA_OK(h, q.Jp.reserve((size_t)E * 432 + 8));
  ESLAM_CUDA_OK(h, q.frm_start.reserve((size_t)(F + 2) * 4)); ESLAM_CUDA_OK(h, q.frm_edges.reserve((size_t)(E + 1) * 4));
  ESLAM_CUDA_OK(h, q.Dp.reserve((size_t)NP * 288 + 8)); ESLAM_CUDA_OK(h, q.Op.reserve((size_t)NP * 288 + 8)); ESLAM_CUDA_OK(h, q.bp.reserve((size_t)NP * 48 + 8));
  ESLAM_CUDA_OK(h, q.W.reserve((size_t)rows * ncols * 8 + 8)); ESLAM_CUDA_OK(h, q.X.reserve((size_t)rows * ncols * 8 + 8));
  ESLAM_CUDA_OK(h, q.Lf.reserve((size_t)NP * 288 + 8)); ESLAM_CUDA_OK(h, q.Mf.reserve((size_t)NP * 288 + 8));
  ESLAM_CUDA_OK(h, q.S.reserve((size_t)(n_o + 1) * n_o * 8)); ESLAM_CUDA_OK(h, q.rhs.reserve((size_t)n_o * 8));
  ESLAM_CUDA_OK(h, q.xo_tmp.reserve((size_t)n_o * 8)); ESLAM_CUDA_OK(h, q.xo.reserve((size_t)n_o * 8));
  ESLAM_CUDA_OK(h, q.xp_tmp.reserve((size_t)rows * 8 + 8)); ESLAM_CUDA_OK(h, q.xp.reserve((size_t)rows * 8 + 8));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(q.T.p, g->Tcw, (size_t)F * 56, cudaMemcpyHostToDevice, s));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(q.odo.p, odo.data(), odo.size() * 8, cudaMemcpyHostToDevice, s));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(q.frm_start.p, start.data(), (size_t)(F + 2) * 4, cudaMemcpyHostToDevice, s));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(q.frm_edges.p, list.data(), (size_t)(E + 1) * 4, cudaMemcpyHostToDevice, s));
  ESLAM_CUDA_OK(h, cudaMemsetAsync(q.xo.p, 0, (size_t)n_o * 8, s));
  ESLAM_CUDA_OK(h, cudaMemsetAsync(q.xp.p, 0, (size_t)rows * 8 + 8, s));
  ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));

  LmState *st = w.lm.as<LmState>();
  LmState *hs = w.pin.as<LmState>();
  EslamLmStats out; memset(&out, 0, sizeof(out));
  int nBad = 0;
  const int maxTrials = 10;
  for (int it = 0; it < iters; it++) {
    if ((rc = slam_eval(h, s, q.T.as<double>(), w.ell.as<double>(), true))) return rc;
    build_system_kernel<<<No, 96, 0, s>>>(w.obj_start.as<int>(), w.obj_edges.as<int>(), w.type.as<int>(), w.info.as<double>(), w.res.as<double>(),
                                          w.jac.as<double>(), w.H.as<double>(), w.b.as<double>(), nullptr);
    ESLAM_LAUNCHED(h);
    if (NP > 0) {
      ESLAM_CUDA_OK(h, cudaMemsetAsync(q.W.p, 0, (size_t)rows * ncols * 8, s));
      slam_build_pose_kernel<<<NP, 64, 0, s>>>(NP, ncols, q.frm_start.as<int>(), q.frm_edges.as<int>(), w.type.as<int>(), w.object.as<int>(), w.info.as<double>(),
                                               w.res.as<double>(), q.Jp.as<double>(), w.jac.as<double>(), odom_info, q.ores.as<double>(), q.oJi.as<double>(),
                                               q.oJj.as<double>(), q.Dp.as<double>(), q.Op.as<double>(), q.bp.as<double>(), q.W.as<double>());
      ESLAM_LAUNCHED(h);
    }
    slam_lm_begin_kernel<<<1, 32, 0, s>>>(st, q.ochi2.as<double>(), NP, w.chi2.as<double>(), E, q.Dp.as<double>(), w.H.as<double>(), No, it);
    ESLAM_LAUNCHED(h);
    double rho = 0; int qmax = 0;
    double iniChi = 0, currentChi = 0;
    do {
      if (NP > 0) {
        slam_factor_kernel<<<1, 32, 0, s>>>(st, NP, q.Dp.as<double>(), q.Op.as<double>(), q.Lf.as<double>(), q.Mf.as<double>());
        ESLAM_LAUNCHED(h);
        slam_subst_kernel<<<(ncols + 63) / 64, 64, 0, s>>>(NP, ncols, q.Lf.as<double>(), q.Mf.as<double>(), q.W.as<double>(), q.X.as<double>());
        ESLAM_LAUNCHED(h);
      }
      dim3 grid((ncols + 31) / 32, (n_o + 31) / 32);
      slam_schur_kernel<<<grid, 256, 0, s>>>(st, rows, n_o, ncols, q.W.as<double>(), q.X.as<double>(), w.H.as<double>(), w.b.as<double>(), q.S.as<double>());
      ESLAM_LAUNCHED(h);
      slam_dense_solve_kernel<<<1, 1024, solve_smem, s>>>(st, n_o, q.S.as<double>(), q.xo_tmp.as<double>());
      ESLAM_LAUNCHED(h);
      if (NP > 0) {
        slam_xp_kernel<<<(rows + 3) / 4, 128, 0, s>>>(rows, n_o, ncols, q.X.as<double>(), q.xo_tmp.as<double>(), q.xp_tmp.as<double>());
        ESLAM_LAUNCHED(h);
      }
      slam_update_kernel<<<(NP + 1 + No + 63) / 64, 64, 0, s>>>(st, NP, No, q.xp_tmp.as<double>(), q.xo_tmp.as<double>(), q.xp.as<double>(), q.xo.as<double>(),
                                                                q.T.as<double>(), q.T_trial.as<double>(), w.ell.as<double>(), w.ell_trial.as<double>());
      ESLAM_LAUNCHED(h);
      if ((rc = slam_eval(h, s, q.T_trial.as<double>(), w.ell_trial.as<double>(), false))) return rc;
      slam_decide_kernel<<<1, 32, 0, s>>>(st, q.ochi2.as<double>(), NP, w.chi2.as<double>(), E, q.xp.as<double>(), q.bp.as<double>(), q.xo.as<double>(),
                                          w.b.as<double>(), No, q.T.as<double>(), q.T_trial.as<double>(), w.ell.as<double>(), w.ell_trial.as<double>(), hs);
      ESLAM_LAUNCHED(h);
      ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));
      rho = hs->rho; iniChi = hs->iniChi; currentChi = hs->currentChi;
      if (it == 0 && qmax == 0) out.chi2_initial = iniChi;
      qmax++;
    } while (rho < 0 && qmax < maxTrials);
    out.iterations++; out.trials_total += qmax; out.chi2_final = currentChi; out.lambda_final = hs->lambda;
    bool terminate = false;
    if (qmax == maxTrials || rho == 0) terminate = true;
    else {
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
      if (nBad >= 3) terminate = true;
    }
    if (terminate) { out.terminated = 1; break; }
  }
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(ell_inout, w.ell.p, (size_t)No * 80, cudaMemcpyDeviceToHost, s));
  if (Tcw_out) ESLAM_CUDA_OK(h, cudaMemcpyAsync(Tcw_out, q.T.p, (size_t)F * 56, cudaMemcpyDeviceToHost, s));
  ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));
  if (stats) *stats = out;
  return ESLAM_OK;
}

void slam_free(EslamHandle *h) {
  if (!h->sl) return;
  SlamWorkspace &q = *h->sl;
  DevBuf *bufs[] = {&q.T, &q.T_trial, &q.odo, &q.ores, &q.oJi, &q.oJj, &q.ochi2, &q.Jp, &q.frm_start, &q.frm_edges, &q.Dp, &q.Op, &q.bp, &q.W, &q.X,
                    &q.Lf, &q.Mf, &q.S, &q.rhs, &q.xo_tmp, &q.xp_tmp, &q.xp, &q.xo};
  for (DevBuf *b : bufs) b->release();
  delete h->sl; h->sl = nullptr;
}

}  // namespace eslam

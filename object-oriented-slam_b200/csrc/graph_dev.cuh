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

// Sequential (graph-order) sum acc + v[0] + v[1] + ... by one warp, 256 (or 32) terms per step, BIT-IDENTICAL to the one-by-one loop
// (g2o's activeChi2: `chi += e->chi2()` over the active edges) -- every lane returns the same value.
// While the running sum s stays inside one binade [2^e, 2^(e+1)) its ulp u is fixed and s = S u with an integer S; a term
// c >= 0 is c = (q + f) u with an integer q and 0 <= f < 1, and RN(s + c) = (S + q + [f > 1/2]) u: the rounding of each step depends
// on the term alone (ties f = 1/2 excepted, where it depends on the parity of S + q).  So the increments d = q + [f > 1/2] of a block of terms are
// computed independently, added up with an integer warp reduction, and the result is exact provided no term is a tie, negative,
// subnormal, non-finite or above the binade, and S + sum d stays below 2^53 (no carry into the next binade at any prefix, because
// the increments are non-negative).  A step that fails any of these runs the plain 32-step loop instead -- in a chi-square sum of
// 25 000 edges that is the ~20 steps in which the sum crosses a power of two.
// increment of the integer mantissa S of a running sum in binade `es` (biased exponent) by the term c; ok = false where the term does
// not allow the shortcut (negative, subnormal, non-finite, above the binade, or a tie)
__device__ __forceinline__ unsigned long long ordered_sum_inc(int es, double c, bool &ok) {
  const unsigned long long kMant = (1ull << 52) - 1ull;
  if (c == 0.0) return 0ull;                                     // +-0 adds nothing
  const unsigned long long cb = (unsigned long long)__double_as_longlong(c);
  const int ec = (int)((cb >> 52) & 0x7ffull);
  const int shift = es - ec;
  if ((cb >> 63) != 0ull || ec == 0 || ec == 0x7ff || shift < 0) { ok = false; return 0ull; }
  if (shift >= 54) return 0ull;                                  // c < u / 2: the sum does not move
  const unsigned long long mc = (cb & kMant) | (1ull << 52);
  if (shift == 0) return mc;
  const unsigned long long q = mc >> shift, f = mc & ((1ull << shift) - 1ull), half = 1ull << (shift - 1);
  if (f == half) { ok = false; return 0ull; }                    // tie: round-to-even needs the running parity
  return q + (f > half ? 1ull : 0ull);
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
// s <- s + (terms whose increments add up to `total`) when that stays inside the binade; false otherwise (s untouched)
__device__ __forceinline__ bool ordered_sum_commit(double &s, unsigned long long total) {
  const unsigned long long kMant = (1ull << 52) - 1ull;
  const unsigned long long sb = (unsigned long long)__double_as_longlong(s);
  const unsigned long long ms = (sb & kMant) | (1ull << 52);
  if (total >= (1ull << 53) || ms + total >= (1ull << 53)) return false;
  s = __longlong_as_double((long long)((sb & ~kMant) | ((ms + total) & kMant)));
  return true;
}
static __device__ double warp_ordered_sum(const double *v, int n, double acc = 0.0) {
  const int lane = threadIdx.x & 31;
  double s = acc;
  double nxt[8];
#pragma unroll
  for (int j = 0; j < 8; j++) nxt[j] = (j * 32 + lane < n) ? v[j * 32 + lane] : 0.0;
  for (int base = 0; base < n; base += 256) {
    double cur[8];
#pragma unroll
    for (int j = 0; j < 8; j++) cur[j] = nxt[j];
    const int nb = base + 256;
#pragma unroll
    for (int j = 0; j < 8; j++) nxt[j] = (nb + j * 32 + lane < n) ? v[nb + j * 32 + lane] : 0.0;   // the next 256 terms are in flight during this block
    // (1) the whole block of 256 terms at once (term j * 32 + lane of the block sits in cur[j] of this lane; missing terms are +0)
    {
      const unsigned long long sb = (unsigned long long)__double_as_longlong(s);
      const int es = (int)((sb >> 52) & 0x7ffull);
      bool ok = (sb >> 63) == 0ull && es > 0 && es < 0x7ff;   // s positive and normal
      unsigned long long dsum = 0ull;
      if (ok) {
#pragma unroll
        for (int j = 0; j < 8; j++) dsum += ordered_sum_inc(es, cur[j], ok);
      }
      if (__all_sync(0xffffffffu, ok) && ordered_sum_commit(s, warp_sum_u64(dsum))) continue;
    }
    // (2) 32 terms at a time, (3) one by one
#pragma unroll 1
    for (int j = 0; j < 8; j++) {
      const int cnt = min(32, n - (base + j * 32));
      if (cnt <= 0) break;
      const double c = cur[j];
      const unsigned long long sb = (unsigned long long)__double_as_longlong(s);
      const int es = (int)((sb >> 52) & 0x7ffull);
      bool ok = (sb >> 63) == 0ull && es > 0 && es < 0x7ff;
      unsigned long long d = 0ull;
      if (ok) d = ordered_sum_inc(es, c, ok);
      if (__all_sync(0xffffffffu, ok) && ordered_sum_commit(s, warp_sum_u64(d))) continue;
      for (int k = 0; k < cnt; k++) s += __shfl_sync(0xffffffffu, c, k);
    }
  }
  return s;
}
// the plain loop (A/B baseline and checker of the routine above)
static __device__ double warp_ordered_sum_plain(const double *v, int n, double acc = 0.0) {
  if ((threadIdx.x & 31) == 0) for (int k = 0; k < n; k++) acc += v[k];
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

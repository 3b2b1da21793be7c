// symmetry.cu -- symmetry-plane estimation (9 seeds x Levenberg-Marquardt on the symmetry cost) + point-cloud completion +
// final ellipsoid.  The nine seeds of an object advance in ROUNDS (one cost evaluation per seed and round, then the scalar LM step);
// two drivers share the device functions below:
//   * the ROUND PIPELINE (default for batches; second half of this file): phase-pure launches per chunk of objects --
//     sym_step_kernel (LM step of all seeds), sym_eval_kernel (one CTA per object, one warp per seed, any warp takes any task),
//     sym_final_kernel -- replayed as one CUDA graph, seed state in HBM / L2 between launches;
//   * the SINGLE LAUNCH (symmetry_finalize_kernel: one CTA per object, the rounds inside the CTA, state in shared memory): handfuls
//     of objects and clouds of 257..512 points (cached path) or more (generic global-memory path).
//
// Replaces
//   Symmetry::estimateSymmetry / GenerateInitPlanes / getProjDepthMat     src/symmetry/Symmetry.cpp:44-161
//   SymmetrySolver::GetPointCloudProb / findMinimalDistanceWithKdTree /
//   GetSymmetryPointOfPlane / OptimizeSymmetry(Dual)Plane                  src/symmetry/SymmetrySolver.cpp:49-333
//   EdgeSymmetry(Dual)Plane::computeError, Vertex(Dual)Plane::oplusImpl    src/symmetry/SymmetrySolver.cpp:375-396,458-481,527-547
//   the g2o LM loop + numeric Jacobian + dense LDLT they run on            (see edges.cu header for citations)
//   completion, re-centring, extents, ellipsoid assembly                   src/pca/EllipsoidExtractor.cpp:399-492,495-531
//
// The cost of a plane is  -mean_i ln P_i  over the mirrored cloud: each mirrored point is projected into the depth image, classified
// (occluded / no depth: cost 0; otherwise its distance to the nearest cloud point) and the ln P are summed IN POINT ORDER in fp64.  The
// numeric Jacobian (delta 1e-9, amplification 5e8) makes every rounding of that sum matter, so the arithmetic that produces a VALUE
// (mirroring, the fp64 distance to the float neighbour, the ordered sum) is the reference's operation for operation, while every
// DECISION (which pixel, occluded or not, which neighbour) is taken by the cheapest test that provably gives the reference's answer:
//   * pixel:      u = uvh0 / uvh2 is only truncated to an int, so a reciprocal-multiply estimate decides it unless it lies within a
//                 margin of an integer (then the exact division runs);
//   * occlusion:  sqrt(s) > depth is decided on s - depth^2 unless the two are within a margin (then the exact sqrt runs);
//   * neighbour:  every full scan of the float cloud also yields the second-best distance; the scanned query q0 is cached with the
//                 radius rho = (d2 - d1) / 2 (rounded down): any later query within rho of q0 has the same unique nearest neighbour
//                 (triangle inequality, slack for the float rounding of the reference's own scan), so the scan is skipped;
//   * the 2*dim finite-difference planes of a linearisation move every mirrored point by < 4e-6 m (checked per linearisation from the
//     plane parameters): a point whose decisions at the base plane hold with that margin ("robust") needs no projection at all for
//     the finite-difference planes -- only the mirrored point and its distance to the known neighbour, for all 2*dim planes in ONE
//     pass over the points (one point load, 2*dim independent fp64 chains per lane); the few non-robust points take the full path.
//   The range image of getProjDepthMat is never materialised (lazy per pixel, cached per point); the kd-tree is an exact brute-force
//   float scan over the <= 512-point cloud in shared memory (same argmin as FLANN's exact search, first index on ties).
// The ordered sums of the 2*dim planes run on 2*dim lanes side by side (values staged through shared memory).
// Structure: the kernel is bound by instruction issue / fetch, not by memory, so what matters is how many instructions run and how
// well they stay in the instruction caches.  The nine seeds of an object are the nine warps of one CTA and advance in rounds: in a
// round every live seed runs its next cost evaluation (a finite-difference pass or a trial plane), then -- one barrier later -- warp 0
// runs the scalar Levenberg-Marquardt step of all nine seeds in nine LANES (solve, oplus, plane constants: ~2 k scalar fp64
// instructions per step, sin / cos / atan2 shared as non-inlined routines), instead of nine warps each running it in one lane.
#include "common.cuh"
#include "dmath.cuh"
#include "block_prims.cuh"
#include "extract_types.cuh"
#include "internal.h"

namespace eslam {
namespace {

constexpr double kFdDelta = 4.0e-6;        // verified bound on |pm_fd - pm_base| per coordinate for the finite-difference planes [m]
constexpr double kPixMargin = 4.04e-6;     // kFdDelta * 1.01 (denominator shift), scaled per point to pixels
constexpr double kDepthMargin = 5.0e-6;    // |s - t| >= kDepthMargin * (2 + s + t)  =>  |sqrt s - sqrt t| >= 1e-5 > sqrt(3) * kFdDelta
constexpr float kNnSafe = 1.0e-5f;         // a cached neighbour is final for the finite-difference planes when rho - |q - q0| exceeds this
constexpr double kQMax = 14.0;             // float rounding of a query (2^-23 |q|) stays below kNnSafe - sqrt(3) kFdDelta for |q| <= kQMax
constexpr int kSlowMax = 48;               // non-robust points of one half handled by the finite-difference pass (more: generic evaluations)
constexpr int kSeeds = 9;
constexpr unsigned kIdxMask = 1023u;       // neighbour index packed into the low mantissa bits of rho^2 (rounded down: conservative)

struct PlanePair {   // per-evaluation constants of GetSymmetryPointOfPlane for the plane and its dual (hoisted, same rounding)
  double p1[4], p2[4], down1, down2, n1[3], n2[3];
};
struct SeedResult { double prob, init_err, final_err; dplane plane; };

// Shared memory of a CTA: the object (constants + float cloud) and nine seeds (LM state + evaluation scratch).
template <int CAP>
struct SymObj {
  const uint16_t *img; size_t pitch;
  const double *objn; const float4 *nn_g; ObjState *st;
  int M, symtype, npts, in_smem;
  int bx0, bx2, by1, by3, rows, cols;
  double P[12], cam[3], fx, cx, cy, scale, coef, A0, A1, A2, R2;
  int n_live, slotbit, sm, iters;
  // work list of the current phase (any warp takes any task): task t < n_fd is the FD pass of fd_seed[t]; then every trial seed
  // contributes nhalves * nchunks classification tasks; scan phase: seed tr_seed[k] owns tasks [scan_first[k], scan_first[k + 1])
  int n_tasks, next_task, n_fd, n_tr;
  unsigned char fd_seed[kSeeds + 3], tr_seed[kSeeds + 3];
  int scan_first[kSeeds + 1];
  // float cloud, pair-interleaved SoA (odd tail padded far away)
  unsigned long long X2[CAP / 2], Y2[CAP / 2], Z2[CAP / 2];
  double PX[CAP], PY[CAP], PZ[CAP];   // the object-normalised cloud itself (objn), SoA
};
enum { kActDone = 0, kActInit = 1, kActFd = 2, kActTrial = 3 };
// Levenberg-Marquardt state of one seed: everything that survives from one cost evaluation to the next (no per-evaluation scratch).
// The single-launch kernel keeps it in shared memory; the round pipeline (sym_eval_kernel / sym_step_kernel) keeps it in HBM / L2
// between launches and stages it through shared memory inside an evaluation launch.
struct SeedCore {
  // ---- A: read (and partly written) by every cost evaluation
  float4 *nnc[2];          // NN cache, two banks: [bank][point] = (q0.x, q0.y, q0.z, rho^2 | neighbour index)
  uint4 *pxc;              // pixel cache: [point] = (x | y << 16, ushort range value, depth = value / scale as a double)
  PlanePair pp;            // the trial plane being evaluated
  double J[3];
  double err_last;         // value of the last computeError call
  int nBad, ok, qmax, ok2, it, cur, fd_ok, action;
  int wbank, qn, invalid;  // the running single-plane evaluation: bank it writes, points queued for a scan, invalid points
  int stale;               // round pipeline (one NN bank): the last trial was rejected, so the bank no longer holds the evaluation of `est`
  unsigned n_evals, n_queries, n_points, n_hits, n_ref, n_verify_fail;
  int inv_s[6];
  // ---- B: finite-difference passes only
  alignas(16) PlanePair base;
  PlanePair fd[6];
  // ---- C: the scalar Levenberg-Marquardt step only
  alignas(16) dplane est;
  dplane backup;
  double H[9], b[3], x[3];
  double lambda, ni, currentChi, iniChi, err_est;   // err_est: error of `est`
  SeedResult res;
};
template <int CAP>
struct SymSeed : SeedCore {
  unsigned mA[2][2][CAP / 32];        // [bank][half][word]: point needs work for a finite-difference plane (cost != 0, or not robust)
  unsigned mF[2][2][CAP / 32];        //                     point is robust, carries a distance and its cached neighbour is final
  // ---- per-evaluation scratch (everything from here on is dead between two evaluations)
  alignas(16) double lnp[2 * CAP];    // ln P of every mirrored point (-0.0 where it carries no distance), point order; FD passes stage their tiles here
  alignas(16) unsigned short queue[2 * CAP];   // points that need a scan: li | half << 9 | robust << 15 (FD passes keep their side buffer here)
};
// scratch of a finite-difference pass (the seed runs no other evaluation meanwhile): staging tile [6][33] and the slow points' values
// [6][kSlowMax] in lnp, the active-point list [CAP] and the slow-point list [kSlowMax] in queue
constexpr int kFdSideOff = 200;
static_assert(6 * 33 <= kFdSideOff && kFdSideOff + 6 * kSlowMax <= 2 * 256 && 256 + kSlowMax <= 2 * 256, "FD scratch aliases lnp / queue");

__device__ __forceinline__ int x86_int(double a) { return (a > -2147483649.0 && a < 2147483648.0) ? (int)a : (int)0x80000000; }
__device__ __forceinline__ double rcp_fast(double a) {   // ~1 ulp reciprocal without the special-case path (guarded use only)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = __fma_rn(-a, r, 1.0); r = __fma_rn(r, e, r);
  e = __fma_rn(-a, r, 1.0); r = __fma_rn(r, e, r);
  return r;
}

// sin / cos / atan2 exist ONCE in the kernel (the scalar LM step calls them ~11 times; inlined they were 40 KB of cold code)
__device__ __noinline__ double n_sin(double a) { return sin(a); }
__device__ __noinline__ double n_cos(double a) { return cos(a); }
__device__ __noinline__ double n_atan2(double y, double x) { return atan2(y, x); }
// dmath.cuh's q_axis_angle / plane_rotation / plane_oplus_common / plane_another on the shared trig routines (same arithmetic)
__device__ __forceinline__ quat s_axis_angle(double angle, d3 axis) {
  const double ha = 0.5 * angle;
  const bool zero = (ha == 0.0);   // sin(+-0) = +-0 and cos(+-0) = 1 exactly
  const double sn = zero ? ha : n_sin(ha);
  quat q; q.x = sn * axis.x; q.y = sn * axis.y; q.z = sn * axis.z; q.w = zero ? 1.0 : n_cos(ha);
  return q;
}
__device__ __forceinline__ mat3 s_plane_rotation(d3 v) {  // Plane.h:88-92
  const double az = n_atan2(v.y, v.x);
  const double rxy = sqrt(v.x * v.x + v.y * v.y);
  const double el = (v.z == 0.0 && rxy > 0.0) ? v.z : n_atan2(v.z, rxy);   // atan2(+-0, r > 0) = +-0 exactly
  return q_to_R(q_mul(s_axis_angle(az, mk3(0, 0, 1)), s_axis_angle(-el, mk3(0, 1, 0))));
}
// VertexPlane (yaw, 0, dis) :527-532 / VertexDualPlane (+ dual) :544-547 on Plane.h:46-77 with elevation 0
__device__ __noinline__ void sym_oplus(dplane &pl, const double *u, int symtype) {
  const double az = u[0], dd = u[1];
  const d3 n = mk3(1.0 * n_cos(az), 1.0 * n_sin(az), 0.0);   // c = cos(0) = 1, s = sin(0) = 0
  const mat3 R = s_plane_rotation(mk3(pl.p[0], pl.p[1], pl.p[2]));
  const double d = -pl.p[3] + dd;
  const d3 rn = m3_vec(R, n);
  pl.p[0] = rn.x; pl.p[1] = rn.y; pl.p[2] = rn.z; pl.p[3] = -d;
  const double nn = sqrt(pl.p[0] * pl.p[0] + pl.p[1] * pl.p[1] + pl.p[2] * pl.p[2]);
  const double inv = 1. / nn;
  pl.p[0] = pl.p[0] * inv; pl.p[1] = pl.p[1] * inv; pl.p[2] = pl.p[2] * inv; pl.p[3] = pl.p[3] * inv;
  if (symtype != 1) pl.dual += u[2];
}
// a / b with the zero numerators of the vertical planes (p[2] = 0, dual = 0) answered without the division's slow path: +-0 / b = +-0 for b > 0
__device__ __forceinline__ double div_z(double a, double b) { return (a == 0.0 && b > 0.0) ? a : a / b; }
__device__ __noinline__ void make_pair(PlanePair &P, const dplane &pl) {
  for (int k = 0; k < 4; k++) P.p1[k] = pl.p[k];
  const double ang = n_atan2(pl.p[1], pl.p[0]);                              // plane_another: Plane.cpp:137-144 + :97-105
  P.p2[0] = n_sin(ang); P.p2[1] = -n_cos(ang); P.p2[2] = 0; P.p2[3] = -pl.dual;
  P.down1 = sqrt(P.p1[0] * P.p1[0] + P.p1[1] * P.p1[1] + P.p1[2] * P.p1[2]);
  P.down2 = sqrt(P.p2[0] * P.p2[0] + P.p2[1] * P.p2[1] + P.p2[2] * P.p2[2]);
  for (int k = 0; k < 3; k++) { P.n1[k] = div_z(P.p1[k], P.down1); P.n2[k] = div_z(P.p2[k], P.down2); }
}
// GetSymmetryPointOfPlane :193-214 with q = plane, n = its unit normal, down = |q.xyz|
__device__ __forceinline__ d3 mirror_q(d3 pt, const double *q, const double *n, double down) {
  const double val = q[0] * pt.x + q[1] * pt.y + q[2] * pt.z + q[3] * 1.0;
  const double dis_pl = fabs(val) / down;
  const double kk = 2 * ((val > 0) ? -1.0 : 1.0) * dis_pl;
  return mk3(pt.x + kk * n[0], pt.y + kk * n[1], pt.z + kk * n[2]);
}
template <int CAP>
__device__ __forceinline__ d3 load_pt(const SymObj<CAP> &O, int j) {
  if (O.in_smem) return mk3(O.PX[j], O.PY[j], O.PZ[j]);
  return mk3(O.objn[3 * j], O.objn[3 * j + 1], O.objn[3 * j + 2]);
}

// Classification of a mirrored point (SymmetrySolver.cpp:62-116): 0 = cost 0 (occluded / no depth), 1 = needs the nearest-neighbour
// distance, 2 = invalid (not counted).  `rob` = every decision holds for any point within kFdDelta (per coordinate) of pm.
// `pc` is the point's pixel-cache entry: the range value of a pixel is a function of the pixel alone.
template <int CAP>
__device__ __forceinline__ int classify(const SymObj<CAP> &O, d3 pm, uint4 pc, uint4 *slot, bool &rob) {
  double uvh[3];
#pragma unroll
  for (int r = 0; r < 3; r++) { double s = O.P[r * 4] * pm.x; s += O.P[r * 4 + 1] * pm.y; s += O.P[r * 4 + 2] * pm.z; s += O.P[r * 4 + 3] * 1.0; uvh[r] = s; }
  // pixel: the reference truncates u = uvh0 / uvh2 (cvttsd2si); the estimate decides unless it is within the margin of an integer
  const double rw = rcp_fast(uvh[2]);
  const double ua = uvh[0] * rw, va = uvh[1] * rw;
  const double arw = fabs(rw) * kPixMargin;
  const double mu = (O.A0 + fabs(ua) * O.A2) * arw + 1e-6, mv = (O.A1 + fabs(va) * O.A2) * arw + 1e-6;   // 1e-6 px also covers the estimate's own error
  const double fu = ua - floor(ua), fv = va - floor(va);
  int x, y;
  rob = fabs(uvh[2]) > 1e-3 && fabs(ua) < 1e5 && fabs(va) < 1e5 && fu > mu && fu < 1.0 - mu && fv > mv && fv < 1.0 - mv;
  if (rob) { x = (int)ua; y = (int)va; }
  else { const double u = uvh[0] / uvh[2], v = uvh[1] / uvh[2]; x = x86_int(u); y = x86_int(v); }
  const bool valid = isfinite(pm.x) && isfinite(pm.y) && isfinite(pm.z);
  rob = rob && valid;
  bool inrange = (O.bx0 < x && x < O.bx2) && (O.by1 < y && y < O.by3);
  if (inrange && (x < 0 || y < 0 || x >= O.cols || y >= O.rows)) inrange = false;
  if (inrange) {
    const unsigned key = (unsigned)x | ((unsigned)y << 16);
    unsigned d; double depth;
    if (pc.x == key) {
      d = pc.y; depth = __hiloint2double((int)pc.w, (int)pc.z);
    } else {
      const unsigned short d0 = O.img[(size_t)y * O.pitch + x];
      if (d0 == 0) { d = 0; depth = 0.0; }   // 0 * sqrt(..) / f = +0 -> ushort 0 (and 0 / x is the division's slow path)
      else {
        const double xi = (x - O.cx), yi = (y - O.cy), f = O.fx;
        const double realz = (double)d0 * sqrt(xi * xi + f * f + yi * yi) / f;   // calculateProjZ, Symmetry.cpp:137-139 (lazy range image)
        d = (unsigned short)(int)realz;                                            // ushort(realz)
        depth = d / O.scale;
      }
      *slot = make_uint4(key, d, (unsigned)__double2loint(depth), (unsigned)__double2hiint(depth));
    }
    if (d == 0) return 0;
    const double ax = O.cam[0] - pm.x, ay = O.cam[1] - pm.y, az = O.cam[2] - pm.z;
    const double s = ax * ax + ay * ay + az * az, t = depth * depth;
    bool occluded;
    if (fabs(s - t) >= kDepthMargin * (2.0 + s + t)) occluded = s > t;   // sqrt(s) > depth with >= 1e-5 to spare
    else { occluded = sqrt(s) > depth; rob = false; }
    if (occluded) return 0;
  }
  return valid ? 1 : 2;
}

// packed fp32x2 arithmetic (Blackwell FADD2/FMUL2): two cloud points per instruction, each half rounded (rn) exactly like the scalar op
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) { return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32); }

// fp64 distance from the query to the FLOAT neighbour `best` of the shared-memory cloud (findMinimalDistanceWithKdTree :159-163)
template <int CAP>
__device__ __forceinline__ double dist_to_neighbour(const SymObj<CAP> &O, d3 pm, int best) {
  const unsigned long long X = O.X2[best >> 1], Y = O.Y2[best >> 1], Z = O.Z2[best >> 1];
  const int sh = (best & 1) * 32;
  const double bx = (double)__uint_as_float((unsigned)(X >> sh)), by = (double)__uint_as_float((unsigned)(Y >> sh)), bz = (double)__uint_as_float((unsigned)(Z >> sh));
  const double dx = pm.x - bx, dy = pm.y - by, dz = pm.z - bz;
  return sqrt(dx * dx + dy * dy + dz * dz);
}

struct NN2 { float r1, r2; int best; };   // smallest and second smallest float distance (flann::L2_Simple order), first index of the smallest
// Exact float NN over pairs [j0, j1) of the shared-memory cloud: minimum of r = ((dx^2 + dy^2) + dz^2), first index on ties -- the
// reference's linear scan -- plus the runner-up distance (equal to r1 on a tie).  Branch free; all lanes scan the same pair at the
// same time (broadcast LDS.64).  Products packed (FMUL2), sums scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even
// under -fmad=false, which would change the rounding; scalar add.rn.f32 is never contracted.
template <int CAP>
__device__ __forceinline__ NN2 nn_scan_range(const SymObj<CAP> &O, float qx, float qy, float qz, int j0, int j1) {
  const unsigned long long qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
  float bd = FLT_MAX, sd = FLT_MAX; int best = 0x7fffffff;
#pragma unroll 2
  for (int j = j0; j < j1; j++) {
    const unsigned long long dx = f2_sub(qx2, O.X2[j]), dy = f2_sub(qy2, O.Y2[j]), dz = f2_sub(qz2, O.Z2[j]);
    const unsigned long long sx = f2_mul(dx, dx), sy = f2_mul(dy, dy), sz = f2_mul(dz, dz);
    const float r0 = __fadd_rn(__fadd_rn(__uint_as_float((unsigned)sx), __uint_as_float((unsigned)sy)), __uint_as_float((unsigned)sz));
    const float r1 = __fadd_rn(__fadd_rn(__uint_as_float((unsigned)(sx >> 32)), __uint_as_float((unsigned)(sy >> 32))), __uint_as_float((unsigned)(sz >> 32)));
    sd = fminf(sd, fmaxf(r0, bd));
    if (r0 < bd) { bd = r0; best = 2 * j; }
    sd = fminf(sd, fmaxf(r1, bd));
    if (r1 < bd) { bd = r1; best = 2 * j + 1; }
  }
  NN2 r; r.r1 = bd; r.r2 = sd; r.best = best;
  return r;
}
template <int CAP>
__device__ __noinline__ NN2 nn_scan(const SymObj<CAP> &O, float qx, float qy, float qz) {
  NN2 r = nn_scan_range(O, qx, qy, qz, 0, (O.M + 1) >> 1);
  if (r.best == 0x7fffffff) r.best = 0;   // every distance is +inf / NaN: the linear scan keeps index 0
  return r;
}
// Warp-cooperative variant for a PARTIAL chunk of cq <= 16 queries: the 32 lanes are split into L = 32 / slots parts, part p scans the
// p-th slice of the cloud for query (lane % slots); the parts are merged with an (r, index) lexicographic butterfly -- slices are
// index-ordered, so "smallest r, then smallest index" is the linear scan's first-minimum rule.
template <int CAP>
__device__ __noinline__ NN2 nn_scan_split(const SymObj<CAP> &O, float qx, float qy, float qz, int slots) {
  const int lane = threadIdx.x & 31;
  const int L = 32 / slots, part = lane / slots;
  const int pairs = (O.M + 1) >> 1, pp = (pairs + L - 1) / L;
  const int j0 = part * pp, j1 = min(pairs, j0 + pp);
  NN2 r = nn_scan_range(O, qx, qy, qz, j0, j1);
  for (int o = slots; o < 32; o <<= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, r.r1, o), os = __shfl_xor_sync(0xffffffffu, r.r2, o);
    const int oi = __shfl_xor_sync(0xffffffffu, r.best, o);
    r.r2 = fminf(fminf(r.r2, os), fmaxf(r.r1, ob));
    if (ob < r.r1 || (ob == r.r1 && oi < r.best)) { r.r1 = ob; r.best = oi; }
  }
  if (r.best == 0x7fffffff) r.best = 0;
  return r;
}
// cache word of a scanned query: rho^2 (rounded down, low mantissa bits cleared) | neighbour index.  rho = (d2 - d1) / 2 with both
// distances moved 8e-6 relative against us (float rounding of the reference's r, of this scan and of the later |q - q0|^2).
__device__ __forceinline__ float nn_pack(const NN2 &r) {
  const float a = sqrtf(r.r1), b = sqrtf(r.r2);
  const float gap = b * (1.0f - 8e-6f) - a * (1.0f + 8e-6f);
  float rho2 = 0.0f;
  if (gap > 0.0f) { const float hh = 0.5f * gap; rho2 = fminf(hh * hh * (1.0f - 1e-4f), 1e30f); }
  return __uint_as_float((__float_as_uint(rho2) & ~kIdxMask) | (unsigned)r.best);
}
// does the cached entry c decide the neighbour of float query (qx, qy, qz)?  `safe`: also for any query within kNnSafe of it
__device__ __forceinline__ bool nn_cached(float4 c, float qx, float qy, float qz, bool &safe) {
  const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
  const float s2 = dx * dx + dy * dy + dz * dz;
  const float rho2 = __uint_as_float(__float_as_uint(c.w) & ~kIdxMask);
  const bool same = __float_as_uint(qx) == __float_as_uint(c.x) && __float_as_uint(qy) == __float_as_uint(c.y) && __float_as_uint(qz) == __float_as_uint(c.z);
  const bool hit = (s2 < rho2) || same;
  // safe <=> sqrt(s2) (1 + 1e-6) + kNnSafe < sqrt(rho2) (1 - 1e-6), decided on squares: (a + k)^2 <= 2 a^2 + 2 k^2
  safe = hit && (2.0f * s2 * (1.0f + 4e-6f) + 2.0f * kNnSafe * kNnSafe < rho2 * (1.0f - 4e-6f));
  return hit;
}

// ---------------------------------------------------------------------------------------------------------------- generic path
// M > CAP (never seen at the reference's 0.1 m grid; kept for completeness): global-memory cloud, no caches, point by point.
template <int CAP>
__device__ __noinline__ double eval_generic(const SymObj<CAP> &O, SymSeed<CAP> &S, const PlanePair &PP) {
  const int lane = threadIdx.x & 31;
  const int npts = O.npts, M = O.M;
  double ln_total = 0.0; int invalid = 0;
  for (int base = 0; base < npts; base += 32) {
    const int idx = base + lane;
    double lnP = 0.0; int cls = 0;
    if (idx < npts) {
      const bool second = idx >= M;
      const d3 pm = mirror_q(load_pt(O, idx - (second ? M : 0)), second ? PP.p2 : PP.p1, second ? PP.n2 : PP.n1, second ? PP.down2 : PP.down1);
      uint4 scratch; bool rob;
      cls = classify(O, pm, make_uint4(0xffffffffu, 0u, 0u, 0u), &scratch, rob);
      double dis = 0.0;
      if (cls == 1) {
        const float qx = (float)pm.x, qy = (float)pm.y, qz = (float)pm.z;
        float bd = FLT_MAX; int best = 0;
        for (int k = 0; k < M; k++) {
          const float4 q = O.nn_g[k];
          float dd = __fsub_rn(qx, q.x);
          float r = __fmul_rn(dd, dd);
          dd = __fsub_rn(qy, q.y); r = __fadd_rn(r, __fmul_rn(dd, dd));
          dd = __fsub_rn(qz, q.z); r = __fadd_rn(r, __fmul_rn(dd, dd));
          if (r < bd) { bd = r; best = k; }
        }
        const float4 q = O.nn_g[best];
        const double dx = pm.x - (double)q.x, dy = pm.y - (double)q.y, dz = pm.z - (double)q.z;
        dis = sqrt(dx * dx + dy * dy + dz * dz);
      }
      lnP = O.coef * dis * dis;
    }
    const int cnt = min(32, npts - base);
    for (int k = 0; k < cnt; k++) ln_total += __shfl_sync(0xffffffffu, lnP, k);
    invalid += __popc(__ballot_sync(0xffffffffu, cls == 2));
  }
  if (lane == 0) { S.n_evals++; S.n_points += npts; }
  const int num_valid = npts - invalid;
  const double aver = (num_valid > 0) ? (ln_total / (double)num_valid) : -INFINITY;
  return -aver;
}

// ---------------------------------------------------------------------------------------------------------------- one plane, one warp
// (fallback of the finite-difference pass; the trial planes go through classify_chunk / scan_chunk below)
// Edge error (= -mean ln P) of the prepared plane pair PP; warp-collective, every lane returns the same value.
// wbank >= 0: the evaluation becomes the state of bank `wbank` (NN cache entries + robust masks), read from bank S.cur;
// wbank < 0:  read-only (finite-difference fallback).
// (1) classify all mirrored points of a half; a point that needs a distance either hits the NN cache or is queued, (2) the queued
// points are scanned with full lanes, (3) ln P accumulated in point order over the points that carry a distance (every other point
// contributes -0.0, which never changes the sum).
template <int CAP, bool VERIFY>
__device__ __noinline__ double eval_single(const SymObj<CAP> &O, SymSeed<CAP> &S, const PlanePair &PP, int wbank) {
  if (CAP > 256 && !O.in_smem) return eval_generic(O, S, PP);
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int npts = O.npts, M = O.M;
  const double coef = O.coef;
  const float4 *rc = S.nnc[S.cur];
  unsigned *queue = reinterpret_cast<unsigned *>(S.lnp + CAP);   // warp-local evaluation: compact ln P list in the first half, queue in the second
  float4 *wc = wbank >= 0 ? S.nnc[wbank] : nullptr;
  double ln_total = 0.0;
  int invalid = 0;
  for (int half = 0; half * M < npts; half++) {
    const int h0 = half * M;
    const double *q = half ? PP.p2 : PP.p1, *n = half ? PP.n2 : PP.n1;
    const double down = half ? PP.down2 : PP.down1;
    int qn = 0, nzn = 0, hits = 0;
    for (int base = 0; base < M; base += 32) {
      const int li = base + lane;
      int cls = 0; bool rob = true, safe = false, hit = false;
      double dis = 0.0;
      if (li < M) {
        const float4 c = rc[h0 + li];   // both cache entries are requested before the fp64 classification chain
        const uint4 pc = S.pxc[h0 + li];
        const d3 pm = mirror_q(load_pt(O, li), q, n, down);
        cls = classify(O, pm, pc, S.pxc + h0 + li, rob);
        if (cls == 1) {
          hit = nn_cached(c, (float)pm.x, (float)pm.y, (float)pm.z, safe);
          if (hit) {
            const int best = (int)(__float_as_uint(c.w) & kIdxMask);
            dis = dist_to_neighbour(O, pm, best);
            if (VERIFY) { const NN2 r = nn_scan(O, (float)pm.x, (float)pm.y, (float)pm.z); if (r.best != best) atomicAdd(&S.n_verify_fail, 1u); }
          }
        }
        if (wc && !(cls == 1 && !hit)) wc[h0 + li] = c;   // carried forward (the scanned points write their new entry below)
      }
      const unsigned need = __ballot_sync(0xffffffffu, cls == 1 && !hit);
      const unsigned hitm = __ballot_sync(0xffffffffu, hit);
      const int pos = nzn + __popc((need | hitm) & lt);    // slot in the ordered list of points that carry a distance
      if (hit) S.lnp[pos] = coef * dis * dis;
      if (cls == 1 && !hit) queue[qn + __popc(need & lt)] = (unsigned)li | ((unsigned)pos << 9) | (rob ? 0x80000000u : 0u);
      qn += __popc(need); nzn += __popc(need | hitm); hits += __popc(hitm);
      invalid += __popc(__ballot_sync(0xffffffffu, cls == 2));
      if (wbank >= 0) {
        const unsigned a = __ballot_sync(0xffffffffu, cls == 1 || !rob), f = __ballot_sync(0xffffffffu, cls == 1 && rob && safe);
        if (lane == 0) { S.mA[wbank][half][base >> 5] = a; S.mF[wbank][half][base >> 5] = f; }
      }
    }
    __syncwarp();
    if (lane == 0) { S.n_evals++; S.n_queries += qn; S.n_points += M; S.n_hits += hits; }
    for (int qb = 0; qb < qn; qb += 32) {
      const int cq = min(32, qn - qb);
      int slots = 32;
      if (cq <= 16) { slots = 1; while (slots < cq) slots <<= 1; }
      const int ql = lane % slots;
      const bool active = ql < cq;
      const unsigned e = queue[qb + (active ? ql : 0)];
      const int li = (int)(e & 511u), pos = (int)((e >> 9) & 511u);
      const d3 pm = mirror_q(load_pt(O, li), q, n, down);
      const float qx = (float)pm.x, qy = (float)pm.y, qz = (float)pm.z;
      NN2 r;
      if (slots == 32) { if (active) r = nn_scan(O, qx, qy, qz); }
      else r = nn_scan_split(O, qx, qy, qz, slots);
      if (active && lane < slots) {
        const double dis = dist_to_neighbour(O, pm, r.best);
        S.lnp[pos] = coef * dis * dis;
        if (wbank >= 0) {
          const float w = nn_pack(r);
          wc[h0 + li] = make_float4(qx, qy, qz, w);
          const float rho2 = __uint_as_float(__float_as_uint(w) & ~kIdxMask);
          if ((e & 0x80000000u) && rho2 * (1.0f - 4e-6f) > kNnSafe * kNnSafe) atomicOr(&S.mF[wbank][half][li >> 5], 1u << (li & 31));
        }
      }
    }
    __syncwarp();
    {   // ordered sum: every lane runs the same chain on broadcast loads
      int k = 0;
      for (; k + 4 <= nzn; k += 4) {
        const double2 a = *reinterpret_cast<const double2 *>(&S.lnp[k]), b = *reinterpret_cast<const double2 *>(&S.lnp[k + 2]);
        ln_total += a.x; ln_total += a.y; ln_total += b.x; ln_total += b.y;
      }
      for (; k < nzn; k++) ln_total += S.lnp[k];
    }
    __syncwarp();
  }
  const int num_valid = npts - invalid;
  const double aver = (num_valid > 0) ? (ln_total / (double)num_valid) : -INFINITY;
  return -aver;
}

// ---------------------------------------------------------------------------------------------------------------- one plane, any warps
// The evaluation of a trial plane (S.pp, result = state of bank S.wbank) cut into independent tasks: classify_chunk takes 32 points of
// one half (ln P of the cache hits and -0.0 of the cost-free points into S.lnp[point], scan candidates into S.queue), scan_chunk
// takes 32 queued points; the ordered sum over S.lnp runs in the scalar phase.
template <int CAP, bool VERIFY>
__device__ __noinline__ void classify_chunk(const SymObj<CAP> &O, SymSeed<CAP> &S, int half, int chunk) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int M = O.M, h0 = half * M, wbank = S.wbank;
  const PlanePair &PP = S.pp;
  const double *q = half ? PP.p2 : PP.p1, *n = half ? PP.n2 : PP.n1;
  const double down = half ? PP.down2 : PP.down1;
  const int li = chunk * 32 + lane;
  int cls = 0; bool rob = true, safe = false, hit = false;
  if (li < M) {
    const bool two_banks = S.nnc[0] != S.nnc[1];   // (the round pipeline keeps ONE bank: both names point at it)
    float4 c = S.nnc[S.cur][h0 + li];   // the cache entries are requested before the fp64 classification chain
    const float4 c2 = two_banks ? S.nnc[wbank][h0 + li] : c;   // the other bank: the entry of an earlier (rejected) trial -- any entry ever written is valid
    const uint4 pc = S.pxc[h0 + li];
    const d3 pm = mirror_q(load_pt(O, li), q, n, down);
    cls = classify(O, pm, pc, S.pxc + h0 + li, rob);
    double lnp = -0.0;
    if (cls == 1) {
      hit = nn_cached(c, (float)pm.x, (float)pm.y, (float)pm.z, safe);
      if (!hit && two_banks && wbank != S.cur) {   // a shorter step after a rejected trial often lands near that trial's query
        bool safe2;
        if (nn_cached(c2, (float)pm.x, (float)pm.y, (float)pm.z, safe2)) { hit = true; safe = safe2; c = c2; }
      }
      if (hit) {
        const int best = (int)(__float_as_uint(c.w) & kIdxMask);
        const double dis = dist_to_neighbour(O, pm, best);
        lnp = O.coef * dis * dis;
        if (VERIFY) { const NN2 r = nn_scan(O, (float)pm.x, (float)pm.y, (float)pm.z); if (r.best != best) atomicAdd(&S.n_verify_fail, 1u); }
      }
    }
    S.lnp[h0 + li] = lnp;
    if (two_banks && !(cls == 1 && !hit)) S.nnc[wbank][h0 + li] = c;   // carried forward (the scanned points write their new entry later)
  }
  const unsigned need = __ballot_sync(0xffffffffu, cls == 1 && !hit);
  const unsigned a = __ballot_sync(0xffffffffu, cls == 1 || !rob), f = __ballot_sync(0xffffffffu, cls == 1 && rob && safe);
  const int ninv = __popc(__ballot_sync(0xffffffffu, cls == 2));
  const int nhit = __popc(__ballot_sync(0xffffffffu, hit));
  int qbase = 0;
  if (lane == 0) {
    S.mA[wbank][half][chunk] = a; S.mF[wbank][half][chunk] = f;
    if (need) qbase = atomicAdd(&S.qn, __popc(need));
    if (ninv) atomicAdd(&S.invalid, ninv);
    if (nhit) atomicAdd(&S.n_hits, (unsigned)nhit);
  }
  qbase = __shfl_sync(0xffffffffu, qbase, 0);
  if (cls == 1 && !hit) S.queue[qbase + __popc(need & lt)] = (unsigned short)(li | (half << 9) | (rob ? 0x8000 : 0));
}
template <int CAP, bool VERIFY>
__device__ __noinline__ void scan_chunk(const SymObj<CAP> &O, SymSeed<CAP> &S, int qb) {
  const int lane = threadIdx.x & 31;
  const int M = O.M, wbank = S.wbank;
  const PlanePair &PP = S.pp;
  const int cq = min(32, S.qn - qb);
  int slots = 32;
  if (cq <= 16) { slots = 1; while (slots < cq) slots <<= 1; }
  const int ql = lane % slots;
  const bool active = ql < cq;
  const unsigned e = S.queue[qb + (active ? ql : 0)];
  const int li = (int)(e & 511u), half = (int)((e >> 9) & 1u), h0 = half * M;
  const d3 pm = mirror_q(load_pt(O, li), half ? PP.p2 : PP.p1, half ? PP.n2 : PP.n1, half ? PP.down2 : PP.down1);
  const float qx = (float)pm.x, qy = (float)pm.y, qz = (float)pm.z;
  NN2 r;
  if (slots == 32) { if (active) r = nn_scan(O, qx, qy, qz); }
  else r = nn_scan_split(O, qx, qy, qz, slots);
  if (active && lane < slots) {
    const double dis = dist_to_neighbour(O, pm, r.best);
    S.lnp[h0 + li] = O.coef * dis * dis;
    const float w = nn_pack(r);
    S.nnc[wbank][h0 + li] = make_float4(qx, qy, qz, w);
    const float rho2 = __uint_as_float(__float_as_uint(w) & ~kIdxMask);
    if ((e & 0x8000u) && rho2 * (1.0f - 4e-6f) > kNnSafe * kNnSafe) atomicOr(&S.mF[wbank][half][li >> 5], 1u << (li & 31));
  }
}

// ---------------------------------------------------------------------------------------------------------------- 2*dim planes
// All finite-difference planes of one linearisation in ONE pass over the points of bank S.cur (the evaluation of S.est).  Writes
// S.J[0..DIM).  Returns false (nothing written) when a half has more than kSlowMax non-robust points: the caller falls back to
// read-only single-plane evaluations.
template <int CAP, int DIM, bool VERIFY>
__device__ __noinline__ bool eval_fd(const SymObj<CAP> &O, SymSeed<CAP> &S) {
  constexpr int NFD = 2 * DIM;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int npts = O.npts, M = O.M, cur = S.cur;
  const double coef = O.coef;
  const float4 *rc = S.nnc[cur];
  double (*stage)[33] = reinterpret_cast<double (*)[33]>(S.lnp);
  double (*side)[kSlowMax] = reinterpret_cast<double (*)[kSlowMax]>(S.lnp + kFdSideOff);
  unsigned short *list = S.queue, *slow = S.queue + CAP;
  double ln_tot = 0.0;   // lanes < NFD: ordered sum of plane `lane`
  if (lane < 6) S.inv_s[lane] = 0;
  unsigned visited = 0;
  for (int half = 0; half * M < npts; half++) {
    const int h0 = half * M;
    // compact lists of the active and of the slow (active, not fast) points, in point order
    int n_act = 0, n_slow = 0;
    for (int w = 0; w * 32 < M; w++) {
      const unsigned A = S.mA[cur][half][w], F = S.mF[cur][half][w], Sl = A & ~F;
      if (A >> lane & 1u) list[n_act + __popc(A & lt)] = (unsigned short)(w * 32 + lane);
      if (Sl >> lane & 1u) { const int k = n_slow + __popc(Sl & lt); if (k < kSlowMax) slow[k] = (unsigned short)(w * 32 + lane); }
      n_act += __popc(A); n_slow += __popc(Sl);
    }
    if (n_slow > kSlowMax) return false;   // warp uniform
    __syncwarp();
    visited += (unsigned)n_act;
    // slow points: one lane per (point, plane), full classification, cached or scanned neighbour
    for (int it = lane; it < n_slow * NFD; it += 32) {
      const int s = it / NFD, k = it - s * NFD;
      const int li = slow[s];
      const PlanePair &Q = S.fd[k];
      const d3 pm = mirror_q(load_pt(O, li), half ? Q.p2 : Q.p1, half ? Q.n2 : Q.n1, half ? Q.down2 : Q.down1);
      bool rob, safe;
      uint4 scratch;   // several lanes may hold the same point: the pixel cache is read, not written
      const int cls = classify(O, pm, S.pxc[h0 + li], &scratch, rob);
      double val = -0.0;
      if (cls == 1) {
        const float qx = (float)pm.x, qy = (float)pm.y, qz = (float)pm.z;
        const float4 c = rc[h0 + li];
        int best;
        if (nn_cached(c, qx, qy, qz, safe)) best = (int)(__float_as_uint(c.w) & kIdxMask);
        else best = nn_scan(O, qx, qy, qz).best;
        const double dis = dist_to_neighbour(O, pm, best);
        val = coef * dis * dis;
      }
      if (cls == 2) atomicAdd(&S.inv_s[k], 1);
      side[k][s] = val;
    }
    __syncwarp();
    for (int base = 0; base < n_act; base += 32) {
      const bool e = base + lane < n_act;
      const int li = e ? (int)list[base + lane] : 0;
      const bool fast = e && ((S.mF[cur][half][li >> 5] >> (li & 31)) & 1u);
      d3 pt = mk3(0, 0, 0);
      double bx = 0, by = 0, bz = 0;
      int best = 0, sslot = 0;
      if (fast) {
        best = (int)(__float_as_uint(rc[h0 + li].w) & kIdxMask);
        pt = load_pt(O, li);
        const unsigned long long X = O.X2[best >> 1], Y = O.Y2[best >> 1], Z = O.Z2[best >> 1];
        const int sh = (best & 1) * 32;
        bx = (double)__uint_as_float((unsigned)(X >> sh)); by = (double)__uint_as_float((unsigned)(Y >> sh)); bz = (double)__uint_as_float((unsigned)(Z >> sh));
      } else if (e) {
        for (int s = 0; s < n_slow; s++) if ((int)slow[s] == li) sslot = s;
      }
#pragma unroll 2
      for (int k = 0; k < NFD; k++) {
        double val = -0.0;
        if (fast) {
          const PlanePair &Q = S.fd[k];
          const d3 pm = mirror_q(pt, half ? Q.p2 : Q.p1, half ? Q.n2 : Q.n1, half ? Q.down2 : Q.down1);
          const double dx = pm.x - bx, dy = pm.y - by, dz = pm.z - bz;
          const double dis = sqrt(dx * dx + dy * dy + dz * dz);
          val = coef * dis * dis;
          if (VERIFY) {   // the full path must agree: same class, same neighbour
            bool rob2; uint4 scratch;
            const int cls = classify(O, pm, make_uint4(0xffffffffu, 0u, 0u, 0u), &scratch, rob2);
            const NN2 r = nn_scan(O, (float)pm.x, (float)pm.y, (float)pm.z);
            if (cls != 1 || r.best != best) atomicAdd(&S.n_verify_fail, 1u);
          }
        } else if (e) {
          val = side[k][sslot];
        }
        stage[k][lane] = val;
      }
      __syncwarp();
      if (lane < NFD) {
        const int cnt = min(32, n_act - base);
        const double *row = stage[lane];
#pragma unroll 4
        for (int i = 0; i < cnt; i++) ln_tot += row[i];
      }
      __syncwarp();
    }
  }
  double err = 0.0;
  if (lane < NFD) {
    const int num_valid = npts - S.inv_s[lane];
    const double aver = (num_valid > 0) ? (ln_tot / (double)num_valid) : -INFINITY;
    err = -aver;
  }
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const double e1 = __shfl_sync(0xffffffffu, err, 2 * d), e2 = __shfl_sync(0xffffffffu, err, 2 * d + 1);
    if (lane == 0) S.J[d] = (1.0 / (2 * 1e-9)) * (e1 - e2);
  }
  if (lane == 0) { S.n_evals += NFD; S.n_points += visited * NFD; }
  __syncwarp();
  return true;
}

template <int DIM>
__device__ __noinline__ bool lm_solve_dim(const double *H, const double *b, double lambda, double *x) {
  double Mx[DIM * DIM];
  for (int i = 0; i < DIM * DIM; i++) Mx[i] = H[i];
  for (int i = 0; i < DIM; i++) Mx[i * DIM + i] += lambda;
  return ldlt_solve_small<DIM>(Mx, b, x);
}
// Upper bound of the per-coordinate displacement of any mirrored point between plane pairs A (base) and B:
// pm = pt - 2 (n.pt + c) n with c = p3 / |p|;  |d pm| <= 2 [ (|dn| R + |dc|) |n'| + (R + |c|) |dn| ].
__device__ __forceinline__ bool fd_pair_ok(const PlanePair &A, const PlanePair &B, double R) {
  bool ok = true;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const double *na = h ? A.n2 : A.n1, *nb = h ? B.n2 : B.n1;
    const double ca = div_z(h ? A.p2[3] : A.p1[3], h ? A.down2 : A.down1), cb = div_z(h ? B.p2[3] : B.p1[3], h ? B.down2 : B.down1);
    const double dn = 1.7320508075688774 * fmax(fmax(fabs(nb[0] - na[0]), fabs(nb[1] - na[1])), fabs(nb[2] - na[2]));
    const double dc = fabs(cb - ca), cabs = fabs(ca);
    const double dpm = 2.0 * ((dn * R + dc) * 1.001 + (R + cabs) * dn) + 1e-12;
    ok = ok && (dpm <= kFdDelta) && (3.0 * R + 2.0 * cabs + 1e-3 <= kQMax);   // NaN anywhere -> false
  }
  return ok;
}

// One scalar Levenberg-Marquardt step of a seed (OptimizeSymmetryPlane / OptimizeSymmetryDualPlane: one vertex, one unary edge with
// information 1, optimize(iters); g2o optimization_algorithm_levenberg.cpp:61-164), run between two evaluation rounds by ONE LANE.
// It consumes the result of the seed's last action (S.err_last / S.J) and prepares the next one.  Exact economy over the literal call
// sequence: computeActiveErrors() at the top of an iteration re-evaluates the plane whose error is already known (the accepted trial,
// or the restored backup) -- the cost is a pure function, so the value is reused.
__device__ __noinline__ void lm_step(const int symtype, const int iters, SeedCore &S) {
  const int dim = (symtype == 1) ? 2 : 3;
  bool start_trial = false, begin_iter = false;
  if (S.action == kActInit) {
    S.n_ref = 1; S.res.init_err = S.err_last; S.err_est = S.err_last; S.stale = 0;
    S.lambda = -1; S.ni = 2; S.nBad = 0; S.x[0] = S.x[1] = S.x[2] = 0; S.ok = 1; S.it = 0;
    begin_iter = true;
  } else if (S.action == kActFd) {
    const double err = S.err_est;   // constructQuadraticForm with the error of `est`
    for (int i = 0; i < dim; i++) {
      S.b[i] = 0.0 - (S.J[i] * 1.0) * err;
      for (int j = 0; j < dim; j++) S.H[i * dim + j] = 0.0 + (S.J[i] * 1.0) * S.J[j];
    }
    if (S.it == 0) {
      double mx = 0;
      for (int j = 0; j < dim; j++) mx = fmax(fabs(S.H[j * dim + j]), mx);
      S.lambda = 1e-5 * mx; S.ni = 2; S.nBad = 0;
    }
    S.qmax = 0;
    start_trial = true;
  } else if (S.action == kActTrial) {
    const double err = S.err_last;
    S.n_ref++;
    double tempChi = err * (1.0 * err);
    if (!S.ok2) tempChi = DBL_MAX;
    double rho = (S.currentChi - tempChi);
    double scale = 0;
    for (int j = 0; j < dim; j++) scale += S.x[j] * (S.lambda * S.x[j] + S.b[j]);
    scale += 1e-3;
    rho /= scale;
    if (rho > 0 && isfinite(tempChi)) {
      double alpha = 1. - pow((2 * rho - 1), 3.0);
      alpha = fmin(alpha, 2. / 3.);
      const double sf = fmax(1. / 3., alpha);
      S.lambda *= sf; S.ni = 2; S.currentChi = tempChi;
      S.err_est = err; S.cur ^= 1;   // the trial's evaluation is the state of `est` from now on
      S.stale = 0;
    } else {
      S.lambda *= S.ni; S.ni *= 2; S.est = S.backup;
      S.stale = 1;
    }
    S.qmax++;
    if (rho < 0 && S.qmax < 10) start_trial = true;
    else {
      if (S.qmax == 10 || rho == 0) S.ok = 0;
      else {
        if ((S.iniChi - S.currentChi) * 1e3 < S.iniChi) S.nBad++; else S.nBad = 0;
        if (S.nBad >= 3) S.ok = 0;
      }
      S.it++;
      begin_iter = true;
    }
  }
  if (begin_iter) {
    if (S.it >= iters || !S.ok) {
      S.res.final_err = S.err_last;   // e->error() after optimize(): value of the LAST computeError call (possibly a rejected trial)
      S.res.prob = exp(-S.err_last);
      S.res.plane = S.est;
      S.action = kActDone;
    } else {
      S.err_last = S.err_est;         // computeActiveErrors() at `est`
      S.n_ref += 1 + 2 * dim;
      S.currentChi = S.err_est * (1.0 * S.err_est); S.iniChi = S.currentChi;
      S.action = kActFd;
    }
  }
  if (start_trial) {
    S.backup = S.est;
    double x[3] = {S.x[0], S.x[1], S.x[2]};
    const bool ok2 = (dim == 2) ? lm_solve_dim<2>(S.H, S.b, S.lambda, x) : lm_solve_dim<3>(S.H, S.b, S.lambda, x);
    S.x[0] = x[0]; S.x[1] = x[1]; S.x[2] = x[2];
    S.ok2 = ok2 ? 1 : 0;
    sym_oplus(S.est, x, symtype);
    make_pair(S.pp, S.est);
    S.action = kActTrial;
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// completion (EllipsoidExtractor.cpp:399-446), extents (:495-531) and ellipsoid assembly (:461-492) by ONE warp
__device__ __noinline__ void finalize_object(const ObjMeta &m, ObjState &st, const ExtractParamsDev &prm, const ExtractPtrs &W, EslamEllipsoidOut &out) {
  const int lane = threadIdx.x & 31;
  const int M = st.n_sym;
  const int run_sym = st.run_sym;
  const double *objn = W.objn + 3 * (size_t)m.off;
  se3 Two = se3_load(st.T_wo);
  const se3 Twc = se3_load(m.Twc);
  double sym_prob = 1.0;
  dplane bestp; bestp.p[0] = 1; bestp.p[1] = 0; bestp.p[2] = 0; bestp.p[3] = 0; bestp.dual = 0;
  int best = -1;
  if (run_sym) {   // sort(desc by prob) of 9 entries == insertion sort == stable: first maximum wins (Symmetry.cpp:122-126)
    const volatile double *sp = st.seed_prob;
    best = 0;
    for (int s = 1; s < 9; s++) if (sp[s] > sp[best]) best = s;
    for (int k = 0; k < 4; k++) bestp.p[k] = ((const volatile double *)st.seed_plane[best])[k];
    bestp.dual = ((const volatile double *)st.seed_dual)[best];
    sym_prob = sp[best];
  }
  double p1[4] = {bestp.p[0], bestp.p[1], bestp.p[2], bestp.p[3]}, p2[4];
  plane_another(bestp, p2);
  const int copies = run_sym ? ((m.symtype == 2) ? 4 : 2) : 1;
  const int nfin = copies * M;
  auto gen_point = [&](int idx) -> d3 {   // mirrored points generated on the fly, never stored
    int part = idx / M, j = idx - part * M;
    d3 p = mk3(objn[3 * j], objn[3 * j + 1], objn[3 * j + 2]);
    if (part == 1) p = mirror_point(p, p1);                         // pSymCloudNormalized
    else if (part == 2) p = mirror_point(p, p2);                    // pSymCloudNormalized2_1
    else if (part == 3) p = mirror_point(mirror_point(p, p1), p2);  // pSymCloudNormalized2_2
    return p;
  };
  hom34 Tmo_h;
  bool retransform = false;
  if (run_sym) {
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = lane; i < nfin; i += 32) { d3 p = gen_point(i); s0 += p.x; s1 += p.y; s2 += p.z; }
    s0 = warp_sum_d(s0); s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
    const double nn = (double)nfin;
    d3 cc = mk3(s0 / nn, s1 / nn, s2 / nn);   // GetPointcloudCenter
    d3 c0 = mk3(p1[0], p1[1], p1[2]);
    const double n0 = norm3(c0);
    c0 = mk3(c0.x / n0, c0.y / n0, c0.z / n0);
    mat3 Rom; m3_setcol(Rom, 0, c0); m3_setcol(Rom, 2, mk3(0, 0, 1)); m3_setcol(Rom, 1, cross3(mk3(0, 0, 1), c0));
    se3 Tom; Tom.t = cc; Tom.r = q_from_R(Rom);
    Tmo_h = se3_hom(se3_inv(Tom));
    retransform = true;
    if (lane == 0) {
      d3 cw = hom_apply(se3_hom(Two), cc);
      out.sym_center_world[0] = cw.x; out.sym_center_world[1] = cw.y; out.sym_center_world[2] = cw.z;
      double w1[4]; plane_transform_rigid(Two, p1, w1);
      for (int k = 0; k < 4; k++) out.sym_plane_world[k] = w1[k];
      if (m.symtype == 2) { double w2[4]; plane_transform_rigid(Two, p2, w2); for (int k = 0; k < 4; k++) out.sym_plane2_world[k] = w2[k]; }
      out.symmetry_result = 1; out.best_seed = best; st.best_seed = best;
    }
    Two = se3_mul(Two, Tom);
  }
  double e0x = 0, e0y = 0, e0z = 0;   // ProcessPCANormalized: axis-aligned extents
  for (int i = lane; i < nfin; i += 32) {
    d3 p = gen_point(i);
    if (retransform) p = hom_apply(Tmo_h, p);
    e0x = fmax(e0x, fabs(p.x)); e0y = fmax(e0y, fabs(p.y)); e0z = fmax(e0z, fabs(p.z));
    if (W.d_final) { double *q = W.d_final + 3 * ((size_t)4 * m.off + i); q[0] = p.x; q[1] = p.y; q[2] = p.z; }
  }
  e0x = warp_max_d(e0x); e0y = warp_max_d(e0y); e0z = warp_max_d(e0z);
  if (lane == 0) {
    ellipsoid e0;
    mat3 I = m3_eye();
    e0.pose.r = q_from_R(I); e0.pose.t = mk3(0, 0, 0);
    e0.s[0] = e0x; e0.s[1] = e0y; e0.s[2] = e0z;
    ellipsoid eg = e0; eg.pose = se3_mul(Two, e0.pose);                 // transform_from(*pSE3Two)       :465
    ellipsoid el = eg; el.pose = se3_mul(se3_inv(Twc), eg.pose);        // transform_from(campose_wc.inverse()) :468-469
    ell_store(el, out.ellipsoid);
    out.prob = run_sym ? sym_prob : 1.0;
    out.n_final = nfin; st.n_final = nfin;
    out.status = ESLAM_OBJ_OK;
  }
}

// dense index of this SM (first touch): %smid is not dense on a B200 (148 SMs of a larger die), the cache pool is sized for sm_count
__device__ __forceinline__ int dense_sm(int *map) {
  unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if (smid >= (unsigned)kMaxSmId) __trap();
  volatile int *slot = map + smid;
  int idx = *slot;
  if (idx < 0) {
    if (atomicCAS(map + smid, -1, -2) == -1) { idx = atomicAdd(map + kMaxSmId, 1); *slot = idx; __threadfence(); }
    else { while ((idx = *slot) < 0) {} }
  }
  return idx;
}

template <int CAP> constexpr size_t sym_smem_bytes() { return sizeof(SymObj<CAP>) + kSeeds * sizeof(SymSeed<CAP>); }

// grid = n_obj CTAs of nine warps: warp w owns seed w of the CTA's object.  Two instantiations share the launch: CAP = 256 takes the
// objects with n_sym <= 256, CAP = 512 the rest (n_sym <= 512 on the cached path, larger clouds on the generic path); a CTA whose
// object belongs to the other instantiation exits at once.
template <int CAP, int MINB, bool VERIFY>
__global__ void __launch_bounds__(32 * kSeeds, MINB) symmetry_finalize_kernel(const uint16_t *__restrict__ depth, size_t pitch, size_t frame_stride,
                                                                              const ObjMeta *__restrict__ metas, ObjState *__restrict__ states, ExtractParamsDev prm,
                                                                              ExtractPtrs W, EslamEllipsoidOut *__restrict__ outs) {
  extern __shared__ __align__(16) unsigned char sym_smem_raw[];
  SymObj<CAP> &O = *reinterpret_cast<SymObj<CAP> *>(sym_smem_raw);
  SymSeed<CAP> *seeds = reinterpret_cast<SymSeed<CAP> *>(sym_smem_raw + sizeof(SymObj<CAP>));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int obj = blockIdx.x;
  const ObjMeta &m = metas[obj];
  ObjState &st = states[obj];
  EslamEllipsoidOut &out = outs[obj];
  const int status = st.status;
  const int run_sym = st.run_sym;
  const int M = st.n_sym;
  const bool small = (CAP == 256);
  if (status == 0 && run_sym) { if (small != (M <= 256)) return; }   // the other instantiation's object
  else if (!small) return;                                           // objects without symmetry are closed by the CAP = 256 launch
  if (threadIdx.x == 0) {
    for (int i = 0; i < 10; i++) out.ellipsoid[i] = 0;
    out.prob = 0; out.status = status; out.symmetry_type = m.symtype > 0 ? m.symtype : 0; out.symmetry_result = 0; out.best_seed = -1;
    for (int i = 0; i < 4; i++) { out.sym_plane_world[i] = 0; out.sym_plane2_world[i] = 0; }
    for (int i = 0; i < 3; i++) out.sym_center_world[i] = 0;
    out.n_crop = st.n_crop; out.n_voxel = st.n_voxel; out.n_plane = st.n_plane; out.n_clusters = st.n_clusters; out.n_cluster = st.n_cluster;
    out.n_sym = st.n_sym; out.n_final = 0; out.reserved = 0;
    if (!prm.symmetry_open) out.symmetry_type = 0;
  }
  if (status != 0) return;
  if (!run_sym) {   // no symmetry for this label: warp 0 finalises directly
    if (wid == 0) { __syncwarp(); finalize_object(m, st, prm, W, out); }
    return;
  }
  const double *objn = W.objn + 3 * (size_t)m.off;
  const float4 *objnf = W.objnf + m.off;
  const bool in_smem = M <= CAP;
  const int npts = (m.symtype == 2) ? 2 * M : M;
  // cache slot of this resident CTA: a free bit of this SM's mask
  if (threadIdx.x == 0) {
    const int sm = dense_sm(W.smid_map);
    if (sm >= W.sm_count) __trap();
    unsigned *slotword = W.slotmask2 + (small ? 0 : W.sm_count) + sm;
    int slotbit = -1;
    while (slotbit < 0) {
      const unsigned freeb = ~*((volatile unsigned *)slotword);
      if (freeb) {
        const int b = __ffs(freeb) - 1;
        if (!(atomicOr(slotword, 1u << b) & (1u << b))) slotbit = b;
      }
    }
    O.sm = sm; O.slotbit = slotbit;
    O.img = depth + (size_t)m.frame * frame_stride; O.pitch = pitch; O.st = &st;
    O.objn = objn; O.nn_g = objnf; O.M = M; O.symtype = m.symtype; O.npts = npts; O.in_smem = in_smem ? 1 : 0;
    for (int i = 0; i < 12; i++) O.P[i] = st.P[i];
    O.A0 = fabs(O.P[0]) + fabs(O.P[1]) + fabs(O.P[2]); O.A1 = fabs(O.P[4]) + fabs(O.P[5]) + fabs(O.P[6]); O.A2 = fabs(O.P[8]) + fabs(O.P[9]) + fabs(O.P[10]);
    O.cam[0] = st.T_oc[0]; O.cam[1] = st.T_oc[1]; O.cam[2] = st.T_oc[2];
    O.bx0 = st.bbox_i[0]; O.by1 = st.bbox_i[1]; O.bx2 = st.bbox_i[2]; O.by3 = st.bbox_i[3];
    O.rows = prm.rows; O.cols = prm.cols; O.fx = prm.fx; O.cx = prm.cx; O.cy = prm.cy; O.scale = prm.scale;
    const double sinv = 1.0 / prm.sigma;
    O.coef = -0.5 * sinv * sinv;   // ln P = -0.5 * Sigma_inv * Sigma_inv * d * d, evaluated left to right (:117-118)
    O.iters = prm.symmetry_iters; O.n_live = kSeeds; O.R2 = 0.0;
  }
  __syncthreads();
  SymSeed<CAP> &S = seeds[wid];
  {
    constexpr int kSlots = small ? kSymSlots256 : kSymSlots512;
    const size_t slot = (size_t)O.sm * kSlots + O.slotbit;
    unsigned char *pool = (small ? W.sympool256 : W.sympool512) + (slot * kSeeds + wid) * ((size_t)CAP * 96);
    float4 *nn0 = reinterpret_cast<float4 *>(pool), *nn1 = nn0 + 2 * CAP;
    uint4 *pxc = reinterpret_cast<uint4 *>(nn1 + 2 * CAP);
    if (in_smem) {
      // invalidate this seed's caches: a NaN x never equals / is never within rho of a float query; no pixel has x = y = 65535
      for (int j = lane; j < npts; j += 32) {
        nn0[j] = make_float4(__uint_as_float(0x7fffffffu), 0.0f, 0.0f, 0.0f);
        nn1[j] = make_float4(__uint_as_float(0x7fffffffu), 0.0f, 0.0f, 0.0f);
        pxc[j] = make_uint4(0xffffffffu, 0u, 0u, 0u);
      }
      for (int j = threadIdx.x; j < (M + 1) / 2; j += blockDim.x) {
        const float4 a = objnf[2 * j];
        const float far_ = 1.0e30f;   // (q - 1e30)^2 overflows to +inf: never < FLT_MAX, so the pad can never win
        const float4 b = (2 * j + 1 < M) ? objnf[2 * j + 1] : make_float4(far_, far_, far_, 0.0f);
        O.X2[j] = f2_pack(a.x, b.x); O.Y2[j] = f2_pack(a.y, b.y); O.Z2[j] = f2_pack(a.z, b.z);
      }
      for (int j = threadIdx.x; j < M; j += blockDim.x) { O.PX[j] = objn[3 * j]; O.PY[j] = objn[3 * j + 1]; O.PZ[j] = objn[3 * j + 2]; }
      if (wid == 0) {
        double r2 = 0.0;
        for (int j = lane; j < M; j += 32) { const double x = objn[3 * j], y = objn[3 * j + 1], z = objn[3 * j + 2]; r2 = fmax(r2, x * x + y * y + z * z); }
        r2 = warp_max_d(r2);
        if (lane == 0) O.R2 = sqrt(r2) * 1.000001 + 1e-9;
      }
    }
    if (lane == 0) {
      S.nnc[0] = nn0; S.nnc[1] = nn1; S.pxc = pxc;
      S.cur = 0; S.fd_ok = 0; S.action = kActInit;
      S.n_evals = S.n_queries = S.n_points = S.n_hits = S.n_ref = S.n_verify_fail = 0;
      // GenerateInitPlanes (Symmetry.cpp:44-86): seed = i*3+m, dis = -0.2 + 0.2 i, angle = -5deg + 5deg m
      const double diff_dis = 0.2, diff_angle = M_PI / 180.0 * 5;
      const double start_dis = -diff_dis * 1, start_angle = -diff_angle * 1;
      const int si = wid / 3, sm_ = wid % 3;
      const double dis = start_dis + diff_dis * si, angle = start_angle + diff_angle * sm_;
      S.est.p[0] = sin(angle); S.est.p[1] = -cos(angle); S.est.p[2] = 0; S.est.p[3] = -dis; S.est.dual = 0;   // fromDisAngleTrans
      make_pair(S.pp, S.est);
    }
  }
  __syncthreads();
  // Rounds.  Scalar phase (warp 0, one lane per seed): close the seed's running single-plane evaluation (ordered sum over S.lnp) and
  // run its LM step; then the work of all live seeds is cut into tasks that ANY warp takes: a finite-difference pass per linearising
  // seed (the longest tasks, handed out first), a classification task per 32 points of every trial plane, and -- one barrier
  // later -- a scan task per 32 queued points.  Round 0 is the cost of the initial plane (bank 0), entered as a trial.
  const int dim = (O.symtype == 1) ? 2 : 3;
  const int nchunks = (M + 31) >> 5, nhalves = (npts > M) ? 2 : 1;
  bool first = true;
  // phase clocks of the CTA (thread 0): [0] scalar phases, [1] task phases, [2] scan phases, [3] whole CTA (tools/sym_counters.py)
  long long clk_scalar = 0, clk_task = 0, clk_scan = 0, clk_t = 0;
  const long long clk_begin = clock64();
  while (true) {
    if (threadIdx.x == 0) clk_t = clock64();
    if (wid == 0) {
      if (lane < kSeeds) {
        SymSeed<CAP> &T = seeds[lane];
        if (first) { T.wbank = 0; T.qn = 0; T.invalid = 0; }
        else if (T.action != kActDone) {
          if (T.action != kActFd) {   // kActInit / kActTrial: the evaluation of T.pp is complete, sum ln P in point order
            double ln_total = 0.0;
            if (O.in_smem) {
              int k = 0;
              for (; k + 4 <= npts; k += 4) {
                const double2 u = *reinterpret_cast<const double2 *>(&T.lnp[k]), v = *reinterpret_cast<const double2 *>(&T.lnp[k + 2]);
                ln_total += u.x; ln_total += u.y; ln_total += v.x; ln_total += v.y;
              }
              for (; k < npts; k++) ln_total += T.lnp[k];
              const int num_valid = npts - T.invalid;
              const double aver = (num_valid > 0) ? (ln_total / (double)num_valid) : -INFINITY;
              T.err_last = -aver;
              T.n_evals += nhalves; T.n_queries += (unsigned)T.qn; T.n_points += (unsigned)npts;
            }
          }
          lm_step(O.symtype, O.iters, T);
          T.wbank = T.cur ^ 1; T.qn = 0; T.invalid = 0;
        }
      }
      __syncwarp();
      {   // task list of the round: FD passes first (the longest tasks), then the classification chunks of the trial planes
        const int act = lane < kSeeds ? seeds[lane].action : kActDone;
        const unsigned mfd = __ballot_sync(0xffffffffu, act == kActFd), mtr = __ballot_sync(0xffffffffu, act == kActInit || act == kActTrial);
        const unsigned lt = (1u << lane) - 1u;
        if (act == kActFd) O.fd_seed[__popc(mfd & lt)] = (unsigned char)lane;
        if (act == kActInit || act == kActTrial) O.tr_seed[__popc(mtr & lt)] = (unsigned char)lane;
        if (lane == 0) {
          O.n_fd = __popc(mfd); O.n_tr = __popc(mtr);
          O.n_tasks = O.n_fd + O.n_tr * (O.in_smem ? nhalves * nchunks : 1); O.next_task = 0; O.n_live = O.n_fd + O.n_tr;
        }
      }
    }
    first = false;
    __syncthreads();
    if (threadIdx.x == 0) { const long long t = clock64(); clk_scalar += t - clk_t; clk_t = t; }
    if (O.n_live == 0) break;
    while (true) {
      int t = 0;
      if (lane == 0) t = atomicAdd(&O.next_task, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= O.n_tasks) break;
      if (t >= O.n_fd) {
        const int u = t - O.n_fd;
        if (O.in_smem) {
          const int per = nhalves * nchunks, k = u / per, r = u - k * per, hf = r / nchunks;
          classify_chunk<CAP, VERIFY>(O, seeds[O.tr_seed[k]], hf, r - hf * nchunks);
        } else {   // generic path: the whole evaluation is one task
          SymSeed<CAP> &T = seeds[O.tr_seed[u]];
          const double err = eval_generic(O, T, T.pp);
          if (lane == 0) T.err_last = err;
        }
      } else {
        SymSeed<CAP> &T = seeds[O.fd_seed[t]];
        // linearizeOplus (numeric, delta 1e-9): the 2*dim perturbed planes prepared by 2*dim lanes, the base pair by lane 6
        if (lane == 0) T.fd_ok = 1;
        __syncwarp();
        if (lane < 2 * dim) {
          const int d = lane >> 1;
          const double del = (lane & 1) ? -1e-9 : 1e-9;
          double add[3];
          add[0] = (d == 0) ? del : 0.0; add[1] = (d == 1) ? del : 0.0; add[2] = (d == 2) ? del : 0.0;
          dplane tp = T.est;
          sym_oplus(tp, add, O.symtype);
          make_pair(T.fd[lane], tp);
        } else if (lane == 6) {
          make_pair(T.base, T.est);
        }
        __syncwarp();
        if (lane < 2 * dim && !fd_pair_ok(T.base, T.fd[lane], O.R2)) T.fd_ok = 0;
        __syncwarp();
        bool done = false;
        if (T.fd_ok && O.in_smem) done = (dim == 2) ? eval_fd<CAP, 2, VERIFY>(O, T) : eval_fd<CAP, 3, VERIFY>(O, T);
        if (!done) {
          for (int d = 0; d < dim; d++) {
            const double e1 = eval_single<CAP, VERIFY>(O, T, T.fd[2 * d], -1);
            const double e2 = eval_single<CAP, VERIFY>(O, T, T.fd[2 * d + 1], -1);
            if (lane == 0) T.J[d] = (1.0 / (2 * 1e-9)) * (e1 - e2);
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { const long long t = clock64(); clk_task += t - clk_t; clk_t = t; }
    if (O.in_smem) {   // scan tasks: 32 queued points of one trial plane each
      if (wid == 0) {
        int cnt = (lane < O.n_tr) ? (seeds[O.tr_seed[lane]].qn + 31) >> 5 : 0, incl = cnt;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane <= kSeeds) O.scan_first[lane] = incl - cnt;   // exclusive prefix; entry n_tr (and beyond) = total
        if (lane == 0) O.next_task = 0;
        const int total = __shfl_sync(0xffffffffu, incl, 15);   // lanes 0..15 are covered by the 4-step scan (only 9 lanes are non-zero)
        if (lane == 0) O.n_tasks = total;
      }
      __syncthreads();
      while (true) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&O.next_task, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= O.n_tasks) break;
        int k = 0;
        while (k + 1 < O.n_tr && O.scan_first[k + 1] <= t) k++;
        scan_chunk<CAP, VERIFY>(O, seeds[O.tr_seed[k]], (t - O.scan_first[k]) << 5);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) clk_scan += clock64() - clk_t;
  }
  if (threadIdx.x == 0) { st.sym_clk[0] = clk_scalar; st.sym_clk[1] = clk_task; st.sym_clk[2] = clk_scan; st.sym_clk[3] = clock64() - clk_begin; }
  // all seeds done: results to the object state, the slot back to the SM, completion by warp 0
  if (lane == 0) {
    const SeedResult &r = S.res;
    st.seed_prob[wid] = r.prob; st.seed_init[wid] = r.init_err; st.seed_final[wid] = r.final_err; st.seed_dual[wid] = r.plane.dual;
    for (int k = 0; k < 4; k++) st.seed_plane[wid][k] = r.plane.p[k];
    atomicAdd(&st.n_evals, (unsigned long long)S.n_evals); atomicAdd(&st.n_queries, (unsigned long long)S.n_queries);
    atomicAdd(&st.n_points, (unsigned long long)S.n_points); atomicAdd(&st.n_hits, (unsigned long long)S.n_hits);
    atomicAdd(&st.n_ref_evals, (unsigned long long)S.n_ref); atomicAdd(&st.n_ref_points, (unsigned long long)S.n_ref * (unsigned long long)O.npts);
    if (VERIFY) atomicAdd(&st.n_verify_fail, (unsigned long long)S.n_verify_fail);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();   // the cache stores of this CTA are ordered before the next owner's
    atomicAnd(W.slotmask2 + (small ? 0 : W.sm_count) + O.sm, ~(1u << O.slotbit));
  }
  if (wid == 0) finalize_object(m, st, prm, W, out);
}

// ================================================================================================================ round pipeline
// The same search as symmetry_finalize_kernel, cut into PHASE-PURE launches (default; ESLAM_SYM_IMPL=2 selects the single launch):
//   sym_step_kernel   one thread per seed of the whole chunk (plus seven helper lanes): the scalar Levenberg-Marquardt step, the
//                     finite-difference planes of the next linearisation, results of finished seeds
//   sym_eval_kernel   one CTA per object: the cost evaluations the seeds are waiting for (finite-difference passes, trial planes)
//   sym_final_kernel  one warp per object: completion, extents, ellipsoid record
// issued as step(0), { eval(r), step(r + 1) } for r < 1 + 11 * iters (the longest possible LM run: every iteration 1 linearisation
// + 10 trials); a launch whose chunk has no running seed left exits at once.  Why: the single launch spends a quarter of every
// CTA's life in its scalar phase -- one warp streaming ~35 KB of cold fp64 code (oplus, trigonometry, LDLT, divisions) through a
// 32 KB instruction cache that the evaluation code of the co-resident CTAs keeps flushing -- while its eight other warps wait at
// the barrier.  Here that code runs once per round for ALL seeds of a chunk with every group busy, and the evaluation launches hold
// nothing but evaluation code.  The seeds' state (SeedCore + robust masks, kSeedStride bytes) and their NN / pixel caches live in
// HBM / L2 between launches; an evaluation CTA stages its nine seeds through shared memory.
constexpr int kRoundCap = 256;                                        // clouds of up to 256 points; larger ones take the single-launch kernel
constexpr int kSeedPersist = (int)((sizeof(SeedCore) + 2 * sizeof(unsigned[2][2][kRoundCap / 32]) + 15) & ~(size_t)15);
constexpr int kSeedStride = (kSeedPersist + 127) / 128 * 128;          // whole 128-byte lines per seed: an object's state can be dropped from L2 line by line
constexpr size_t kRoundCacheBytes = (size_t)kRoundCap * 64;           // per seed: one NN bank + the pixel cache, 2 * CAP entries of 16 bytes each
constexpr int kGraphAtFirstSight = 64;                                // chunks of this many objects are captured into a graph on their first call
constexpr int kRoundsMinObjects = 16;                                  // smaller launches take the single-launch kernel
constexpr int kStepThreadsDefault = 128;                              // eight lanes per seed: 16 seeds per CTA (one-warp CTAs that slip in beside three resident compute CTAs were measured: no gain)
static_assert(sizeof(SeedCore) % 16 == 0, "SeedCore is followed by the masks without padding");
#pragma nv_diag_suppress 1427
constexpr int kSeedA16 = (int)(__builtin_offsetof(SeedCore, base) / 16);    // part A in 16-byte units
constexpr int kSeedC16 = (int)(__builtin_offsetof(SeedCore, est) / 16);     // parts A + B
#pragma nv_diag_default 1427
constexpr int kSeedCore16 = (int)(sizeof(SeedCore) / 16);
static_assert(kSeedA16 <= 32 && __builtin_offsetof(SeedCore, base) % 16 == 0 && __builtin_offsetof(SeedCore, est) % 16 == 0, "part A is one 16-byte load per lane");
static_assert(sizeof(SymSeed<kRoundCap>) >= (size_t)kSeedPersist, "persistent prefix");
#pragma nv_diag_suppress 1427
static_assert(__builtin_offsetof(SymSeed<kRoundCap>, mA) == sizeof(SeedCore) && __builtin_offsetof(SymSeed<kRoundCap>, lnp) == (size_t)kSeedPersist,
              "the persistent prefix is SeedCore + the robust masks; the scratch starts right behind it");
#pragma nv_diag_default 1427

struct SymChunk {   // descriptor of a chunk's launches (device memory, fixed address per chunk: the launches can be replayed as a graph)
  const uint16_t *depth; size_t pitch, frame_stride;
  const ObjMeta *metas; ObjState *states; EslamEllipsoidOut *outs;
  unsigned char *seedp;   // [n_obj][kSeeds][kSeedStride]
  unsigned char *cache;   // [n_obj][kSeeds][kRoundCacheBytes]
  int *live;              // [rounds + 2]: seeds of the chunk waiting for evaluation r (zeroed before step(0))
  int n_obj, rounds;
  ExtractParamsDev prm; ExtractPtrs W;
};
__device__ __forceinline__ SeedCore *seed_core(const SymChunk &C, int obj, int sd) {
  return reinterpret_cast<SeedCore *>(C.seedp + ((size_t)obj * kSeeds + sd) * kSeedStride);
}
__device__ __forceinline__ bool round_owned(const ObjState &st) { return st.status == 0 && st.run_sym && st.n_sym <= kRoundCap; }

// copies of a seed's state between HBM / L2 and shared memory with all loads of a batch in flight before the first store
template <int NLANES>
__device__ __forceinline__ void copy_words8(unsigned long long *dst, const unsigned long long *src, int n, int l) {
  for (int base = 0; base < n; base += 8 * NLANES) {
    unsigned long long t[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { const int i = base + u * NLANES + l; t[u] = (i < n) ? src[i] : 0ull; }
#pragma unroll
    for (int u = 0; u < 8; u++) { const int i = base + u * NLANES + l; if (i < n) dst[i] = t[u]; }
  }
}
__device__ __forceinline__ void copy_quads4(uint4 *dst, const uint4 *src, int n, int lane) {
  for (int base = 0; base < n; base += 128) {
    uint4 t[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = base + u * 32 + lane; t[u] = (i < n) ? src[i] : make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = base + u * 32 + lane; if (i < n) dst[i] = t[u]; }
  }
}

// step(0): seed planes (GenerateInitPlanes, Symmetry.cpp:44-86).  step(k > 0): consumes evaluation k - 1.
template <int kStepThreads>
__global__ void __launch_bounds__(kStepThreads) sym_step_kernel(const SymChunk *__restrict__ Cp, int k) {
  const SymChunk &C = *Cp;
  if (k > 0 && *((volatile const int *)(C.live + k - 1)) == 0) return;   // chunk uniform
  const int lane = threadIdx.x & 31, sub = threadIdx.x & 7;
  const int g = (blockIdx.x * kStepThreads + threadIdx.x) >> 3;
  const int obj = g / kSeeds, sd = g - obj * kSeeds;
  bool valid = obj < C.n_obj;
  ObjState *st = valid ? C.states + obj : nullptr;
  if (valid) valid = round_owned(*st);
  const int symtype = valid ? C.metas[obj].symtype : 2;
  const int dim = (symtype == 1) ? 2 : 3;
  // the seed's state is staged through shared memory by its eight lanes (the scalar code touches ~150 fields: one coalesced copy
  // instead of that many dependent L2 round trips)
  __shared__ __align__(16) unsigned long long s_seed[kStepThreads / 8][sizeof(SeedCore) / 8];
  SeedCore *G = valid ? seed_core(C, obj, sd) : nullptr;
  SeedCore *S = reinterpret_cast<SeedCore *>(s_seed[threadIdx.x >> 3]);
  if (valid && k > 0) {
    copy_words8<8>(s_seed[threadIdx.x >> 3], reinterpret_cast<const unsigned long long *>(G), (int)(sizeof(SeedCore) / 8), sub);
  }
  __syncwarp();
  const bool was_live = valid && (k == 0 || S->action != kActDone);
  int action = kActDone;
  if (valid && sub == 0) {
    if (k == 0) {
      unsigned char *cb = C.cache + ((size_t)obj * kSeeds + sd) * kRoundCacheBytes;
      // ONE NN bank (both names point at it): a linearisation always follows an accepted trial, whose evaluation is the last writer
      // of the bank; any entry ever written stays a valid (query, neighbour, radius) triple for the trials in between.  The one
      // exception -- a rejected trial followed by a linearisation (non-finite rho) -- is flagged `stale` and takes full evaluations.
      float4 *nn0 = reinterpret_cast<float4 *>(cb);
      S->nnc[0] = nn0; S->nnc[1] = nn0; S->pxc = reinterpret_cast<uint4 *>(nn0 + 2 * kRoundCap);
      S->cur = 0; S->fd_ok = 0; S->action = kActInit;
      S->n_evals = S->n_queries = S->n_points = S->n_hits = S->n_ref = S->n_verify_fail = 0;
      const double diff_dis = 0.2, diff_angle = M_PI / 180.0 * 5;
      const double start_dis = -diff_dis * 1, start_angle = -diff_angle * 1;
      const int si = sd / 3, sm_ = sd % 3;
      const double dis = start_dis + diff_dis * si, angle = start_angle + diff_angle * sm_;
      S->est.p[0] = sin(angle); S->est.p[1] = -cos(angle); S->est.p[2] = 0; S->est.p[3] = -dis; S->est.dual = 0;   // fromDisAngleTrans
      make_pair(S->pp, S->est);
      S->wbank = 0; S->qn = 0; S->invalid = 0;
      action = kActInit;
    } else if (S->action != kActDone) {
      lm_step(symtype, C.prm.symmetry_iters, *S);
      S->wbank = S->cur ^ 1; S->qn = 0; S->invalid = 0;
      action = S->action;
      if (action == kActDone) {   // results to the object state; the last seed of an object is seen by sym_final_kernel through st->done
        const SeedResult &r = S->res;
        st->seed_prob[sd] = r.prob; st->seed_init[sd] = r.init_err; st->seed_final[sd] = r.final_err; st->seed_dual[sd] = r.plane.dual;
        for (int q = 0; q < 4; q++) st->seed_plane[sd][q] = r.plane.p[q];
        atomicAdd(&st->n_evals, (unsigned long long)S->n_evals); atomicAdd(&st->n_queries, (unsigned long long)S->n_queries);
        atomicAdd(&st->n_points, (unsigned long long)S->n_points); atomicAdd(&st->n_hits, (unsigned long long)S->n_hits);
        const unsigned long long npts = (unsigned long long)((symtype == 2) ? 2 * st->n_sym : st->n_sym);
        atomicAdd(&st->n_ref_evals, (unsigned long long)S->n_ref); atomicAdd(&st->n_ref_points, (unsigned long long)S->n_ref * npts);
        atomicAdd(&st->n_verify_fail, (unsigned long long)S->n_verify_fail);
        __threadfence();
        atomicAdd(&st->done, 1);
      }
    }
  }
  __syncwarp();
  action = __shfl_sync(0xffffffffu, action, lane & ~7);
  // linearizeOplus (numeric, delta 1e-9): the 2*dim perturbed planes by 2*dim lanes of the seed's group, the base pair by lane 6
  if (action == kActFd) {
    if (sub < 2 * dim) {
      const int d = sub >> 1;
      const double del = (sub & 1) ? -1e-9 : 1e-9;
      double add[3];
      add[0] = (d == 0) ? del : 0.0; add[1] = (d == 1) ? del : 0.0; add[2] = (d == 2) ? del : 0.0;
      dplane tp = S->est;
      sym_oplus(tp, add, symtype);
      make_pair(S->fd[sub], tp);
    } else if (sub == 6) {
      make_pair(S->base, S->est);
    }
  }
  __syncwarp();
  bool pair_ok = true;
  if (action == kActFd && sub < 2 * dim) pair_ok = fd_pair_ok(S->base, S->fd[sub], st->sym_R2);
  const unsigned bad = __ballot_sync(0xffffffffu, !pair_ok) & (0xffu << (lane & ~7));
  if (action == kActFd && sub == 0) S->fd_ok = (bad || S->stale) ? 0 : 1;
  __syncwarp();
  if (was_live) copy_words8<8>(reinterpret_cast<unsigned long long *>(G), s_seed[threadIdx.x >> 3], (int)(sizeof(SeedCore) / 8), sub);
  const unsigned lv = __ballot_sync(0xffffffffu, sub == 0 && action != kActDone);
  if (lane == 0 && lv) atomicAdd(C.live + k, __popc(lv));
}

// eval(r): the evaluations the seeds of one object wait for -- the round body of symmetry_finalize_kernel without its scalar phase.
template <bool VERIFY, int MINB>
__global__ void __launch_bounds__(32 * kSeeds, MINB) sym_eval_kernel(const SymChunk *__restrict__ Cp, int r) {
  constexpr int CAP = kRoundCap;
  extern __shared__ __align__(16) unsigned char sym_smem_raw[];
  SymObj<CAP> &O = *reinterpret_cast<SymObj<CAP> *>(sym_smem_raw);
  SymSeed<CAP> *seeds = reinterpret_cast<SymSeed<CAP> *>(sym_smem_raw + sizeof(SymObj<CAP>));
  const SymChunk &C = *Cp;
  if (*((volatile const int *)(C.live + r)) == 0) return;   // chunk uniform
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int obj = blockIdx.x;
  ObjState &st = C.states[obj];
  if (!round_owned(st) || *((volatile int *)&st.done) >= kSeeds) return;   // CTA uniform
  const ObjMeta &m = C.metas[obj];
  const int M = st.n_sym;
  const int npts = (m.symtype == 2) ? 2 * M : M;
  const double *objn = C.W.objn + 3 * (size_t)m.off;
  const float4 *objnf = C.W.objnf + m.off;
  if (threadIdx.x == 0) {
    const ExtractParamsDev &prm = C.prm;
    O.sm = 0; O.slotbit = 0;
    O.img = C.depth + (size_t)m.frame * C.frame_stride; O.pitch = C.pitch; O.st = &st;
    O.objn = objn; O.nn_g = objnf; O.M = M; O.symtype = m.symtype; O.npts = npts; O.in_smem = 1;
    for (int i = 0; i < 12; i++) O.P[i] = st.P[i];
    O.A0 = fabs(O.P[0]) + fabs(O.P[1]) + fabs(O.P[2]); O.A1 = fabs(O.P[4]) + fabs(O.P[5]) + fabs(O.P[6]); O.A2 = fabs(O.P[8]) + fabs(O.P[9]) + fabs(O.P[10]);
    O.cam[0] = st.T_oc[0]; O.cam[1] = st.T_oc[1]; O.cam[2] = st.T_oc[2];
    O.bx0 = st.bbox_i[0]; O.by1 = st.bbox_i[1]; O.bx2 = st.bbox_i[2]; O.by3 = st.bbox_i[3];
    O.rows = prm.rows; O.cols = prm.cols; O.fx = prm.fx; O.cx = prm.cx; O.cy = prm.cy; O.scale = prm.scale;
    const double sinv = 1.0 / prm.sigma;
    O.coef = -0.5 * sinv * sinv;
    O.iters = prm.symmetry_iters; O.n_live = 0; O.R2 = 0.0;
  }
  SymSeed<CAP> &S = seeds[wid];
  uint4 *gseed = reinterpret_cast<uint4 *>(C.seedp + ((size_t)obj * kSeeds + wid) * kSeedStride);
  {   // stage this warp's seed: part A and the robust masks always, the finite-difference planes (part B) when a pass is due; the
      // scalar LM state (part C) never comes here
    uint4 *dst = reinterpret_cast<uint4 *>(&S);
    if (lane < kSeedA16) dst[lane] = gseed[lane];
    for (int i = kSeedCore16 + lane; i < kSeedPersist / 16; i += 32) dst[i] = gseed[i];
    __syncwarp();
    if (S.action == kActFd) copy_quads4(dst + kSeedA16, gseed + kSeedA16, kSeedC16 - kSeedA16, lane);
  }
  for (int j = threadIdx.x; j < (M + 1) / 2; j += blockDim.x) {
    const float4 a = objnf[2 * j];
    const float far_ = 1.0e30f;   // (q - 1e30)^2 overflows to +inf: never < FLT_MAX, so the pad can never win
    const float4 b = (2 * j + 1 < M) ? objnf[2 * j + 1] : make_float4(far_, far_, far_, 0.0f);
    O.X2[j] = f2_pack(a.x, b.x); O.Y2[j] = f2_pack(a.y, b.y); O.Z2[j] = f2_pack(a.z, b.z);
  }
  for (int j = threadIdx.x; j < M; j += blockDim.x) { O.PX[j] = objn[3 * j]; O.PY[j] = objn[3 * j + 1]; O.PZ[j] = objn[3 * j + 2]; }
  __syncwarp();
  if (r == 0) {   // invalidate this seed's caches: a NaN x never equals / is never within rho of a float query; no pixel has x = y = 65535
    float4 *nn0 = S.nnc[0], *nn1 = S.nnc[1]; uint4 *pxc = S.pxc;
    for (int j = lane; j < npts; j += 32) {
      nn0[j] = make_float4(__uint_as_float(0x7fffffffu), 0.0f, 0.0f, 0.0f);
      nn1[j] = make_float4(__uint_as_float(0x7fffffffu), 0.0f, 0.0f, 0.0f);
      pxc[j] = make_uint4(0xffffffffu, 0u, 0u, 0u);
    }
    if (wid == 0) {
      double r2 = 0.0;
      for (int j = lane; j < M; j += 32) { const double x = objn[3 * j], y = objn[3 * j + 1], z = objn[3 * j + 2]; r2 = fmax(r2, x * x + y * y + z * z); }
      r2 = warp_max_d(r2);
      if (lane == 0) st.sym_R2 = sqrt(r2) * 1.000001 + 1e-9;
    }
  }
  __syncthreads();
  const int dim = (O.symtype == 1) ? 2 : 3;
  const int nchunks = (M + 31) >> 5, nhalves = (npts > M) ? 2 : 1;
  if (wid == 0) {   // task list: FD passes first (the longest tasks), then the classification chunks of the trial planes
    const int act = lane < kSeeds ? seeds[lane].action : kActDone;
    const unsigned mfd = __ballot_sync(0xffffffffu, act == kActFd), mtr = __ballot_sync(0xffffffffu, act == kActInit || act == kActTrial);
    const unsigned lt = (1u << lane) - 1u;
    if (act == kActFd) O.fd_seed[__popc(mfd & lt)] = (unsigned char)lane;
    if (act == kActInit || act == kActTrial) O.tr_seed[__popc(mtr & lt)] = (unsigned char)lane;
    if (lane == 0) { O.n_fd = __popc(mfd); O.n_tr = __popc(mtr); O.n_tasks = O.n_fd + O.n_tr * nhalves * nchunks; O.next_task = 0; }
  }
  __syncthreads();
  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&O.next_task, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= O.n_tasks) break;
    if (t >= O.n_fd) {
      const int u = t - O.n_fd, per = nhalves * nchunks, q = u / per, rr = u - q * per, hf = rr / nchunks;
      classify_chunk<CAP, VERIFY>(O, seeds[O.tr_seed[q]], hf, rr - hf * nchunks);
    } else {
      SymSeed<CAP> &T = seeds[O.fd_seed[t]];
      bool done = false;
      if (T.fd_ok) done = (dim == 2) ? eval_fd<CAP, 2, VERIFY>(O, T) : eval_fd<CAP, 3, VERIFY>(O, T);
      if (lane == 0) { atomicAdd((unsigned long long *)&st.sym_clk[0], 1ull); if (!done) atomicAdd((unsigned long long *)&st.sym_clk[1], 1ull); }
      if (!done) {
        for (int d = 0; d < dim; d++) {
          const double e1 = eval_single<CAP, VERIFY>(O, T, T.fd[2 * d], -1);
          const double e2 = eval_single<CAP, VERIFY>(O, T, T.fd[2 * d + 1], -1);
          if (lane == 0) T.J[d] = (1.0 / (2 * 1e-9)) * (e1 - e2);
        }
      }
    }
  }
  __syncthreads();
  if (O.n_tr > 0) {   // scan tasks: 32 queued points of one trial plane each (block uniform)
    if (wid == 0) {
      int cnt = (lane < O.n_tr) ? (seeds[O.tr_seed[lane]].qn + 31) >> 5 : 0, incl = cnt;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (lane <= kSeeds) O.scan_first[lane] = incl - cnt;
      if (lane == 0) O.next_task = 0;
      const int total = __shfl_sync(0xffffffffu, incl, 15);
      if (lane == 0) O.n_tasks = total;
    }
    __syncthreads();
    while (true) {
      int t = 0;
      if (lane == 0) t = atomicAdd(&O.next_task, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= O.n_tasks) break;
      int q = 0;
      while (q + 1 < O.n_tr && O.scan_first[q + 1] <= t) q++;
      scan_chunk<CAP, VERIFY>(O, seeds[O.tr_seed[q]], (t - O.scan_first[q]) << 5);
    }
    __syncthreads();
    // the evaluation of S.pp is complete: sum ln P in point order (every lane runs the same chain on broadcast loads)
    if (S.action == kActInit || S.action == kActTrial) {
      double ln_total = 0.0;
      int q = 0;
      for (; q + 4 <= npts; q += 4) {
        const double2 u = *reinterpret_cast<const double2 *>(&S.lnp[q]), v = *reinterpret_cast<const double2 *>(&S.lnp[q + 2]);
        ln_total += u.x; ln_total += u.y; ln_total += v.x; ln_total += v.y;
      }
      for (; q < npts; q++) ln_total += S.lnp[q];
      const int num_valid = npts - S.invalid;
      const double aver = (num_valid > 0) ? (ln_total / (double)num_valid) : -INFINITY;
      __syncwarp();
      if (lane == 0) { S.err_last = -aver; S.n_evals += nhalves; S.n_queries += (unsigned)S.qn; S.n_points += (unsigned)npts; }
      __syncwarp();
    }
  }
  {   // persistent state back (only a live seed changed)
    if (S.action != kActDone) {
      const uint4 *src = reinterpret_cast<const uint4 *>(&S);
      if (lane < kSeedA16) gseed[lane] = src[lane];
      for (int i = kSeedCore16 + lane; i < kSeedPersist / 16; i += 32) gseed[i] = src[i];
    }
  }
}

// completion (and the records of the objects that never reach the symmetry search)
__global__ void __launch_bounds__(128) sym_final_kernel(const SymChunk *__restrict__ Cp) {
  const SymChunk &C = *Cp;
  const int obj = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (obj >= C.n_obj) return;   // warp uniform
  const ObjMeta &m = C.metas[obj];
  ObjState &st = C.states[obj];
  EslamEllipsoidOut &out = C.outs[obj];
  const int status = st.status;
  if (status == 0 && st.run_sym && st.n_sym > kRoundCap) return;   // owned by the single-launch kernel (CAP = 512 / generic path)
  if (lane == 0) {
    for (int i = 0; i < 10; i++) out.ellipsoid[i] = 0;
    out.prob = 0; out.status = status; out.symmetry_type = m.symtype > 0 ? m.symtype : 0; out.symmetry_result = 0; out.best_seed = -1;
    for (int i = 0; i < 4; i++) { out.sym_plane_world[i] = 0; out.sym_plane2_world[i] = 0; }
    for (int i = 0; i < 3; i++) out.sym_center_world[i] = 0;
    out.n_crop = st.n_crop; out.n_voxel = st.n_voxel; out.n_plane = st.n_plane; out.n_clusters = st.n_clusters; out.n_cluster = st.n_cluster;
    out.n_sym = st.n_sym; out.n_final = 0; out.reserved = 0;
    if (!C.prm.symmetry_open) out.symmetry_type = 0;
  }
  __syncwarp();
  if (status != 0) return;
  if (st.run_sym && *((volatile int *)&st.done) != kSeeds) {   // cannot happen (the round count covers the longest LM run); never emit a half-finished record
    if (lane == 0) out.status = ESLAM_OBJ_INTERNAL;
    return;
  }
  finalize_object(m, st, C.prm, C.W, out);
  // the seeds of this object are done: their state and the touched part of their caches are dead -- dropped from L2 without write-back
  if (C.W.l2_discard && st.run_sym) {
    __syncwarp();
    const int npts = (m.symtype == 2) ? 2 * st.n_sym : st.n_sym;
    l2_discard_lines(C.seedp + (size_t)obj * kSeeds * kSeedStride, (size_t)kSeeds * kSeedStride, lane, 32);
    const size_t touched = (((size_t)npts * 16) + 127) & ~(size_t)127;
    for (int sd = 0; sd < kSeeds; sd++) {
      const unsigned char *cb = C.cache + ((size_t)obj * kSeeds + sd) * kRoundCacheBytes;
      l2_discard_lines(cb, touched, lane, 32);                                   // NN bank
      l2_discard_lines(cb + (size_t)2 * kRoundCap * 16, touched, lane, 32);      // pixel cache
    }
  }
}

template <int CAP, int MINB, bool VERIFY>
cudaError_t launch_variant(cudaStream_t s, int n_obj, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, const ObjMeta *metas, ObjState *states,
                           const ExtractParamsDev &prm, const ExtractPtrs &ptrs, EslamEllipsoidOut *outs) {
  static bool attr[16] = {};   // per device
  int dev = 0; cudaGetDevice(&dev);
  const size_t smem = sym_smem_bytes<CAP>();
  if (dev < 16 && !attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(symmetry_finalize_kernel<CAP, MINB, VERIFY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(symmetry_finalize_kernel<CAP, MINB, VERIFY>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr[dev] = true;
  }
  symmetry_finalize_kernel<CAP, MINB, VERIFY><<<n_obj, 32 * kSeeds, smem, s>>>(depth_dev, pitch, frame_stride, metas, states, prm, ptrs, outs);
  return cudaGetLastError();
}

}  // namespace

int symmetry_v1_launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj, const ExtractParamsDev &prm,
                       const ExtractPtrs &ptrs);

size_t symmetry_pool_bytes(int cap) { return (size_t)kSeeds * (size_t)cap * 96; }   // cache bytes of one resident CTA

size_t symmetry_round_seed_bytes() { return (size_t)kSeeds * kSeedStride; }
size_t symmetry_round_cache_bytes() { return (size_t)kSeeds * kRoundCacheBytes; }
size_t symmetry_round_desc_bytes() { return sizeof(SymChunk); }

// Round pipeline of one chunk on stream s: descriptor upload, live counters, step(0), { eval(r), step(r + 1) }, completion, and the
// single-launch kernel for the (rare) clouds of more than kRoundCap points.
static int symmetry_rounds_launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj,
                                  const ExtractParamsDev &prm, const ExtractPtrs &ptrs, int chunk) {
  ExtractWorkspace &w = *h->ex;
  const int rounds = 1 + 11 * (prm.symmetry_iters > 0 ? prm.symmetry_iters : 0);   // 1 initial cost + per iteration 1 linearisation + at most 10 trials
  SymChunk d;
  d.depth = depth_dev; d.pitch = pitch; d.frame_stride = frame_stride;
  d.metas = w.meta.as<ObjMeta>() + obj0; d.states = w.state.as<ObjState>() + obj0; d.outs = w.out.as<EslamEllipsoidOut>() + obj0;
  d.seedp = w.symseed.as<unsigned char>() + (size_t)obj0 * symmetry_round_seed_bytes();
  d.cache = w.symcache.as<unsigned char>() + (size_t)obj0 * symmetry_round_cache_bytes();
  d.live = w.symlive.as<int>() + (size_t)chunk * (kSymMaxRounds + 2);
  d.n_obj = n_obj; d.rounds = rounds; d.prm = prm; d.W = ptrs;
  SymChunk *pin = w.pin_symchunk2[w.pin_cur].as<SymChunk>() + chunk, *dev = w.symchunk.as<SymChunk>() + chunk;
  memcpy(pin, &d, sizeof(d));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(dev, pin, sizeof(d), cudaMemcpyHostToDevice, s));
  const size_t smem = sym_smem_bytes<kRoundCap>();
  int dev_id = 0; cudaGetDevice(&dev_id);
  static bool attr[16][3] = {};
  const int vi = h->sym_verify ? 1 : 0;
  // evaluation kernel variant: 0 = three resident CTAs per SM (72 registers), 1 = verification build, 2 = two CTAs per SM (env ESLAM_SYM_OCC=2, tuning)
  const int variant = vi ? 1 : (h->sym_occ == 2 ? 2 : 0);
  void (*eval_fn)(const SymChunk *, int) = variant == 1 ? sym_eval_kernel<true, 1> : (variant == 2 ? sym_eval_kernel<false, 2> : sym_eval_kernel<false, kSymCtas256>);
  if (dev_id < 16 && !attr[dev_id][variant]) {
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(eval_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(eval_fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr[dev_id][variant] = true;
  }
  // The clouds of more than kRoundCap points (rare: 1 in ~8000 objects of the bench stream) take the single-launch kernel.  It runs on
  // a side stream next to the rounds: alone, one such CTA lives ~2 ms, and behind the rounds it was the straggler of its whole rank.
  cudaStream_t big = s;
  if (!h->profiling) {
    if (!h->chunk_aux[chunk]) {
      ESLAM_CUDA_OK(h, cudaStreamCreateWithFlags(&h->chunk_aux[chunk], cudaStreamNonBlocking));
      ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->aux_fork[chunk], cudaEventDisableTiming));
      ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->aux_join[chunk], cudaEventDisableTiming));
    }
    big = h->chunk_aux[chunk];
    ESLAM_CUDA_OK(h, cudaEventRecord(h->aux_fork[chunk], s));
    ESLAM_CUDA_OK(h, cudaStreamWaitEvent(big, h->aux_fork[chunk], 0));
  }
  if (big != s) {
    cudaError_t e = vi ? launch_variant<512, 1, true>(big, n_obj, depth_dev, pitch, frame_stride, d.metas, d.states, prm, ptrs, d.outs)
                       : launch_variant<512, kSymCtas512, false>(big, n_obj, depth_dev, pitch, frame_stride, d.metas, d.states, prm, ptrs, d.outs);
    ESLAM_CUDA_OK(h, e);
    ESLAM_LAUNCHED(h);
    ESLAM_CUDA_OK(h, cudaEventRecord(h->aux_join[chunk], big));
  }
  const int step_threads = (h->step_threads == 32 || h->step_threads == 64) ? h->step_threads : kStepThreadsDefault;   // env ESLAM_STEP_THREADS (tuning)
  void (*step_fn)(const SymChunk *, int) = step_threads == 128 ? sym_step_kernel<128> : (step_threads == 64 ? sym_step_kernel<64> : sym_step_kernel<32>);
  const int step_grid = (n_obj * kSeeds * 8 + step_threads - 1) / step_threads;
  auto issue = [&](cudaStream_t q) -> cudaError_t {
    cudaError_t e = cudaMemsetAsync(d.live, 0, sizeof(int) * (size_t)(rounds + 2), q);
    if (e != cudaSuccess) return e;
    step_fn<<<step_grid, step_threads, 0, q>>>(dev, 0);
    for (int r = 0; r < rounds; r++) {
      eval_fn<<<n_obj, 32 * kSeeds, smem, q>>>(dev, r);
      step_fn<<<step_grid, step_threads, 0, q>>>(dev, r + 1);
    }
    sym_final_kernel<<<(n_obj + 3) / 4, 128, 0, q>>>(dev);
    return cudaGetLastError();
  };
  bool replayed = false;
  if (h->sym_graph) {   // the sequence depends on (n_obj, rounds) and two fixed addresses only: capture once per chunk, replay afterwards
    // small launches are captured on the second consecutive call with the same shape; a caller whose small batches change from call
    // to call keeps plain launches instead of paying an instantiation per call (the per-frame tracking driver, 1..8 boxes, does not
    // come here at all: fewer than 16 objects take the single-launch kernel)
    ExtractWorkspace::SymGraph &g = w.symgraph[chunk];
    const bool same = g.n_obj == n_obj && g.rounds == rounds && g.verify == variant + 8 * step_threads && g.desc == dev && g.live == d.live;
    if (g.exec && !same) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    if (!same) { g.n_obj = n_obj; g.rounds = rounds; g.verify = variant + 8 * step_threads; g.desc = dev; g.live = d.live; }
    // (a profiled batch is timed: it never pays for plain launches of a sequence this long; a chunk of a streamed batch -- 64 objects
    // and more -- is expected to come back with the same shape, so its graph is instantiated at first sight: with two workspaces in
    // flight the "second call" of the second workspace would otherwise land ~10 ms of instantiation in the caller's steady state)
    if ((same || h->profiling || n_obj >= kGraphAtFirstSight) && !g.exec) {
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const cudaError_t e1 = issue(s);
        const cudaError_t e2 = cudaStreamEndCapture(s, &graph);
        if (!(e1 == cudaSuccess && e2 == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess)) g.exec = nullptr;
        if (graph) cudaGraphDestroy(graph);
      }
      cudaGetLastError();
    }
    if (g.exec) { ESLAM_CUDA_OK(h, cudaGraphLaunch(g.exec, s)); replayed = true; }
  }
  if (!replayed) ESLAM_CUDA_OK(h, issue(s));
  h->launches += 2 * rounds + 2;
  if (big != s) {
    ESLAM_CUDA_OK(h, cudaStreamWaitEvent(s, h->aux_join[chunk], 0));
  } else {
    cudaError_t e = vi ? launch_variant<512, 1, true>(s, n_obj, depth_dev, pitch, frame_stride, d.metas, d.states, prm, ptrs, d.outs)
                       : launch_variant<512, kSymCtas512, false>(s, n_obj, depth_dev, pitch, frame_stride, d.metas, d.states, prm, ptrs, d.outs);
    ESLAM_CUDA_OK(h, e);
    ESLAM_LAUNCHED(h);
  }
  return ESLAM_OK;
}

int symmetry_finalize_launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj,
                             const ExtractParamsDev &prm, const ExtractPtrs &ptrs, int chunk) {
  ExtractWorkspace &w = *h->ex;
  if (n_obj <= 0) return ESLAM_OK;
  if (h->sym_impl == 1) return symmetry_v1_launch(h, s, depth_dev, pitch, frame_stride, obj0, n_obj, prm, ptrs);
  // round pipeline for batches; a handful of objects (the per-frame tracking driver) is latency bound either way and takes the single
  // launch: one CTA per object and two launches instead of ~115 mostly idle ones
  if (h->sym_impl == 3 && n_obj >= kRoundsMinObjects && 1 + 11 * prm.symmetry_iters <= kSymMaxRounds && chunk >= 0 && chunk < kMaxChunks && w.symseed.p)
    return symmetry_rounds_launch(h, s, depth_dev, pitch, frame_stride, obj0, n_obj, prm, ptrs, chunk);
  const ObjMeta *metas = w.meta.as<ObjMeta>() + obj0;
  ObjState *states = w.state.as<ObjState>() + obj0;
  EslamEllipsoidOut *outs = w.out.as<EslamEllipsoidOut>() + obj0;
  cudaError_t e;
  if (h->sym_verify) {
    e = launch_variant<256, 1, true>(s, n_obj, depth_dev, pitch, frame_stride, metas, states, prm, ptrs, outs);
    if (e == cudaSuccess) e = launch_variant<512, 1, true>(s, n_obj, depth_dev, pitch, frame_stride, metas, states, prm, ptrs, outs);
  } else {
    switch (h->sym_occ) {   // env ESLAM_SYM_OCC: resident CTAs per SM of the CAP = 256 instantiation (register budget), tuning hook
      case 1: e = launch_variant<256, 1, false>(s, n_obj, depth_dev, pitch, frame_stride, metas, states, prm, ptrs, outs); break;
      case 2: e = launch_variant<256, 2, false>(s, n_obj, depth_dev, pitch, frame_stride, metas, states, prm, ptrs, outs); break;
      case 4: e = launch_variant<256, 4, false>(s, n_obj, depth_dev, pitch, frame_stride, metas, states, prm, ptrs, outs); break;
      default: e = launch_variant<256, kSymCtas256, false>(s, n_obj, depth_dev, pitch, frame_stride, metas, states, prm, ptrs, outs); break;
    }
    if (e == cudaSuccess) e = launch_variant<512, kSymCtas512, false>(s, n_obj, depth_dev, pitch, frame_stride, metas, states, prm, ptrs, outs);
  }
  ESLAM_CUDA_OK(h, e);
  ESLAM_LAUNCHED(h); ESLAM_LAUNCHED(h);
  return ESLAM_OK;
}

}  // namespace eslam

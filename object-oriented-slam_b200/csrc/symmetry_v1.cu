// symmetry_v1.cu -- ROUND-1 symmetry kernel, kept only as the A/B baseline of symmetry.cu (env ESLAM_SYM_IMPL=1): same results bit for bit,
// ~3x the instructions.  Not used by default.
// symmetry-plane estimation (9 seeds x in-register Levenberg-Marquardt) + point-cloud completion +
// final ellipsoid, one CTA per object, ONE WARP PER SEED.
//
// Replaces
//   Symmetry::estimateSymmetry / GenerateInitPlanes / getProjDepthMat     src/symmetry/Symmetry.cpp:44-161
//   SymmetrySolver::GetPointCloudProb / findMinimalDistanceWithKdTree /
//   GetSymmetryPointOfPlane / OptimizeSymmetry(Dual)Plane                  src/symmetry/SymmetrySolver.cpp:49-333
//   EdgeSymmetry(Dual)Plane::computeError, Vertex(Dual)Plane::oplusImpl    src/symmetry/SymmetrySolver.cpp:375-396,458-481,527-547
//   the g2o LM loop + numeric Jacobian + dense LDLT they run on            (see edges.cu header for citations)
//   completion, re-centring, extents, ellipsoid assembly                   src/pca/EllipsoidExtractor.cpp:399-492,495-531
// B200 design notes
//   * the range image of getProjDepthMat (307 200 sqrt per object in the reference) is never materialised: the
//     ushort range value is computed lazily for the <= 2M pixels a cost evaluation actually reads;
//   * the kd-tree is replaced by an exact brute-force float NN over the <= ~500-point cloud staged in shared memory
//     (same argmin as FLANN's exact search; distance recomputed in fp64 from the float neighbour, SymmetrySolver.cpp:159-163);
//   * each warp owns one seed end-to-end: the LM control flow (fp64 scalars) is warp-uniform, every cost evaluation
//     streams the mirrored points 32 at a time and accumulates ln P in the reference's sequential order via shuffles,
//     so no block barrier and no scratch memory is needed inside the ~45 sequential evaluations of a seed.
#include "common.cuh"
#include "dmath.cuh"
#include "block_prims.cuh"
#include "extract_types.cuh"
#include "internal.h"

namespace eslam {
namespace symv1 {

constexpr int kSymSmemPts = 256;   // cloud staged in shared memory (pair-interleaved SoA) up to this many points
constexpr int kSymQueue = 256;     // per-warp query queue / distance slots / NN cache (fast path when #mirrored points <= 256)

struct PlanePair {   // per-evaluation constants of GetSymmetryPointOfPlane for the plane and its dual (hoisted, same rounding)
  double p1[4], p2[4], down1, down2, n1[3], n2[3];
};
// Everything that is uniform across a block (object) or a warp (seed) lives in SHARED memory, not in per-lane registers:
// the kernel is latency bound (long fp64 div/sqrt chains), so occupancy -- registers per thread -- is what buys throughput.
struct SymBlock {
  const uint16_t *img; size_t pitch;
  const double *objn; const float4 *nn_g;
  int M, symtype, npts, in_smem;
  int bx0, bx2, by1, by3, rows, cols;
  ObjState *st;
  double P[12], cam[3], fx, cx, cy, scale, sinv;
  unsigned long long X2[kSymSmemPts / 2], Y2[kSymSmemPts / 2], Z2[kSymSmemPts / 2];   // pair-interleaved SoA float cloud (odd tail padded far away)
};
struct SeedResult { double prob, init_err, final_err; dplane plane; };
struct SymWarp {
  float4 *cache;   // per-warp NN cache in HBM/L2: slot i = (float query x, y, z, nearest index) of mirrored point i at its last full scan
  uint4 *pcache;   // per-warp pixel cache: slot i = (x | y << 16, ushort range value, depth = value / scale as a double) of the last pixel it read
  PlanePair pp;
  PlanePair fd[6];   // the 2*dim finite-difference planes of one linearisation, prepared by 2*dim lanes in parallel
  dplane est, backup, trial;
  double H[9], b[3], x[3], J[3];
  double lambda, ni, currentChi, tempChi, iniChi, rho, e1;
  int nBad, ok, qmax, again, accepted;
  unsigned n_evals, n_queries, n_points, n_hits, n_ref;   // work counters, flushed once per seed (n_ref: cost evaluations of the reference's call sequence)
  long long clk[4];   // cycles in: classification passes, NN scans, ordered sums, everything else of the seed (clock64, lane 0)
  SeedResult res;
  unsigned short queue[kSymQueue];
  double dis[kSymQueue];
};

__device__ __forceinline__ int x86_int(double a) { return (a > -2147483649.0 && a < 2147483648.0) ? (int)a : (int)0x80000000; }

__device__ __noinline__ void make_pair(PlanePair &P, const dplane &pl) {
  for (int k = 0; k < 4; k++) P.p1[k] = pl.p[k];
  plane_another(pl, P.p2);
  P.down1 = sqrt(P.p1[0] * P.p1[0] + P.p1[1] * P.p1[1] + P.p1[2] * P.p1[2]);
  P.down2 = sqrt(P.p2[0] * P.p2[0] + P.p2[1] * P.p2[1] + P.p2[2] * P.p2[2]);
  for (int k = 0; k < 3; k++) { P.n1[k] = P.p1[k] / P.down1; P.n2[k] = P.p2[k] / P.down2; }
}
__device__ __forceinline__ d3 mirrored(const SymBlock &B, const PlanePair &P, int idx) {   // GetSymmetryPointOfPlane :193-214
  const int second = idx >= B.M;
  const int j = idx - (second ? B.M : 0);
  const d3 pt = mk3(B.objn[3 * j], B.objn[3 * j + 1], B.objn[3 * j + 2]);
  const double *q = second ? P.p2 : P.p1;
  const double *n = second ? P.n2 : P.n1;
  const double val = q[0] * pt.x + q[1] * pt.y + q[2] * pt.z + q[3] * 1.0;
  const double dis_pl = fabs(val) / (second ? P.down2 : P.down1);
  const double kk = 2 * ((val > 0) ? -1.0 : 1.0) * dis_pl;
  return mk3(pt.x + kk * n[0], pt.y + kk * n[1], pt.z + kk * n[2]);
}
// 0 = cost 0 (occluded / no depth), 1 = needs the nearest-neighbour distance, 2 = invalid (not counted)   SymmetrySolver.cpp:62-116
// `pc` is the point's pixel-cache entry: when the mirrored point projects to the same pixel as last time (always, for the 1e-9
// finite-difference perturbations) the range value of that pixel -- a function of the pixel alone -- is reused instead of being
// reloaded and recomputed (1 gather, 1 sqrt, 2 divisions).
__device__ __forceinline__ int classify(const SymBlock &B, d3 pm, uint4 pc, uint4 *slot) {
  double uvh[3];
#pragma unroll
  for (int r = 0; r < 3; r++) { double s = B.P[r * 4] * pm.x; s += B.P[r * 4 + 1] * pm.y; s += B.P[r * 4 + 2] * pm.z; s += B.P[r * 4 + 3] * 1.0; uvh[r] = s; }
  const double u = uvh[0] / uvh[2], v = uvh[1] / uvh[2];
  const int x = x86_int(u), y = x86_int(v);
  const bool valid = isfinite(pm.x) && isfinite(pm.y) && isfinite(pm.z);
  bool inrange = (B.bx0 < x && x < B.bx2) && (B.by1 < y && y < B.by3);
  if (inrange && (x < 0 || y < 0 || x >= B.cols || y >= B.rows)) inrange = false;
  if (inrange) {
    const unsigned key = (unsigned)x | ((unsigned)y << 16);
    unsigned d; double depth;
    if (pc.x == key) {
      d = pc.y; depth = __hiloint2double((int)pc.w, (int)pc.z);
    } else {
      const unsigned short d0 = B.img[(size_t)y * B.pitch + x];
      const double xi = (x - B.cx), yi = (y - B.cy), f = B.fx;
      const double realz = (double)d0 * sqrt(xi * xi + f * f + yi * yi) / f;   // calculateProjZ, Symmetry.cpp:137-139 (lazy range image)
      d = (unsigned short)(int)realz;                                            // ushort(realz)
      depth = d / B.scale;
      *slot = make_uint4(key, d, (unsigned)__double2loint(depth), (unsigned)__double2hiint(depth));
    }
    if (d == 0) return 0;
    const double ax = B.cam[0] - pm.x, ay = B.cam[1] - pm.y, az = B.cam[2] - pm.z;
    const double dis_cam = sqrt(ax * ax + ay * ay + az * az);
    if (dis_cam > depth) return 0;
  }
  return valid ? 1 : 2;
}
// packed fp32x2 arithmetic (Blackwell FADD2/FMUL2): two cloud points per instruction, each half rounded (rn) exactly like the scalar op
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) { return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32); }

struct NNRes { double dis; int best; };
// fp64 distance from the query to the FLOAT neighbour `best` of the shared-memory cloud (findMinimalDistanceWithKdTree :159-163)
__device__ __forceinline__ double dist_to_neighbour(const SymBlock &B, d3 pm, int best) {
  const unsigned long long X = B.X2[best >> 1], Y = B.Y2[best >> 1], Z = B.Z2[best >> 1];
  const int sh = (best & 1) * 32;
  const double bx = (double)__uint_as_float((unsigned)(X >> sh)), by = (double)__uint_as_float((unsigned)(Y >> sh)), bz = (double)__uint_as_float((unsigned)(Z >> sh));
  const double dx = pm.x - bx, dy = pm.y - by, dz = pm.z - bz;
  return sqrt(dx * dx + dy * dy + dz * dz);
}
// Exact float NN (flann::L2_Simple operation order; minimum r = ((dx^2 + dy^2) + dz^2), first index on ties -- the oracle's linear
// scan) + fp64 distance to the float neighbour (findMinimalDistanceWithKdTree :140-166).  All lanes scan the same pair of cloud
// points at the same time (broadcast LDS.64 of the pair-interleaved SoA copy); the loop body is branch free.
__device__ __noinline__ NNRes nn_distance(const SymBlock &B, d3 pm) {
  const float qx = (float)pm.x, qy = (float)pm.y, qz = (float)pm.z;
  float bd = FLT_MAX; int best = 0;
  const int M = B.M;
  float bx, by, bz;
  if (B.in_smem) {
    const unsigned long long qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
    const int pairs = (M + 1) >> 1;
#pragma unroll 2
    for (int j = 0; j < pairs; j++) {
      unsigned long long dx = f2_sub(qx2, B.X2[j]), dy = f2_sub(qy2, B.Y2[j]), dz = f2_sub(qz2, B.Z2[j]);
      // products packed (FMUL2), sums scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false,
      // which would change the rounding; scalar add.rn.f32 is never contracted.
      const unsigned long long sx = f2_mul(dx, dx), sy = f2_mul(dy, dy), sz = f2_mul(dz, dz);
      const float r0 = __fadd_rn(__fadd_rn(__uint_as_float((unsigned)sx), __uint_as_float((unsigned)sy)), __uint_as_float((unsigned)sz));
      const float r1 = __fadd_rn(__fadd_rn(__uint_as_float((unsigned)(sx >> 32)), __uint_as_float((unsigned)(sy >> 32))), __uint_as_float((unsigned)(sz >> 32)));
      if (r0 < bd) { bd = r0; best = 2 * j; }
      if (r1 < bd) { bd = r1; best = 2 * j + 1; }
    }
    const unsigned long long X = B.X2[best >> 1], Y = B.Y2[best >> 1], Z = B.Z2[best >> 1];
    const int sh = (best & 1) * 32;
    bx = __uint_as_float((unsigned)(X >> sh)); by = __uint_as_float((unsigned)(Y >> sh)); bz = __uint_as_float((unsigned)(Z >> sh));
  } else {
    for (int k = 0; k < M; k++) {
      const float4 q = B.nn_g[k];
      float dd = __fsub_rn(qx, q.x);
      float r = __fmul_rn(dd, dd);
      dd = __fsub_rn(qy, q.y); r = __fadd_rn(r, __fmul_rn(dd, dd));
      dd = __fsub_rn(qz, q.z); r = __fadd_rn(r, __fmul_rn(dd, dd));
      if (r < bd) { bd = r; best = k; }
    }
    const float4 q = B.nn_g[best];
    bx = q.x; by = q.y; bz = q.z;
  }
  const double dx = pm.x - (double)bx, dy = pm.y - (double)by, dz = pm.z - (double)bz;
  NNRes r; r.dis = sqrt(dx * dx + dy * dy + dz * dz); r.best = best;
  return r;
}

// Warp-cooperative variant for a PARTIAL chunk of cq <= 16 queries (shared-memory cloud): the 32 lanes are split into
// L = 32 / pow2ceil(cq) parts, part p scans the p-th slice of the cloud for query (lane % slots); the parts are merged with an
// (r, index) lexicographic butterfly -- slices are index-ordered, so "smallest r, then smallest index" is exactly the linear scan's
// first-minimum rule.  Returns the distance for lanes with part == 0 && q < cq (other lanes: 0).
__device__ __noinline__ NNRes nn_distance_split(const SymBlock &B, d3 pm, int slots, bool active) {
  const int lane = threadIdx.x & 31;
  const int L = 32 / slots;
  const int part = lane / slots;
  const float qx = (float)pm.x, qy = (float)pm.y, qz = (float)pm.z;
  const unsigned long long qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
  const int pairs = (B.M + 1) >> 1;
  const int pp = (pairs + L - 1) / L;
  const int j0 = part * pp, j1 = min(pairs, j0 + pp);
  float bd = FLT_MAX; int best = 0x7fffffff;
  for (int j = j0; j < j1; j++) {
    unsigned long long dx = f2_sub(qx2, B.X2[j]), dy = f2_sub(qy2, B.Y2[j]), dz = f2_sub(qz2, B.Z2[j]);
    const unsigned long long sx = f2_mul(dx, dx), sy = f2_mul(dy, dy), sz = f2_mul(dz, dz);
    const float r0 = __fadd_rn(__fadd_rn(__uint_as_float((unsigned)sx), __uint_as_float((unsigned)sy)), __uint_as_float((unsigned)sz));
    const float r1 = __fadd_rn(__fadd_rn(__uint_as_float((unsigned)(sx >> 32)), __uint_as_float((unsigned)(sy >> 32))), __uint_as_float((unsigned)(sz >> 32)));
    if (r0 < bd) { bd = r0; best = 2 * j; }
    if (r1 < bd) { bd = r1; best = 2 * j + 1; }
  }
  for (int o = slots; o < 32; o <<= 1) {   // merge the parts of the same query (lanes q, q+slots, q+2*slots, ...)
    const float obd = __shfl_xor_sync(0xffffffffu, bd, o);
    const int obest = __shfl_xor_sync(0xffffffffu, best, o);
    if (obd < bd || (obd == bd && obest < best)) { bd = obd; best = obest; }
  }
  NNRes r; r.dis = 0.0; r.best = -1;
  if (!active || part != 0 || best == 0x7fffffff) return r;
  r.dis = dist_to_neighbour(B, pm, best); r.best = best;
  return r;
}

// edge error (= -mean ln P) for the prepared plane pair PP (shared memory); warp-collective, every lane returns the same value.
// Fast path (#points <= kSymQueue): (1) classify all mirrored points; points that need a nearest-neighbour distance either hit the
// NN cache or are compacted into a per-warp queue so that (2) the brute-force searches run with full lanes; (3) ln P accumulated in
// the reference's point order -- over the points with a distance only: every other point contributes -0.0, which never changes the sum.
// (__noinline__: the kernel is instruction-fetch bound when the ~1k-instruction evaluation body is replicated at its call sites)
__device__ __noinline__ double sym_eval(const SymBlock &B, SymWarp &W, const PlanePair &PP, bool update_cache = true) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int npts = B.npts;
  double ln_total = 0.0;
  int invalid = 0;
  if (B.M <= kSymQueue) {
    // the mirrored cloud is processed in halves of M points (plane 1, then the dual plane), so the per-warp queue / distance slots
    // only need M entries; ln P keeps accumulating in the reference's point order across the halves
    const double sinv = B.sinv;
    for (int h0 = 0; h0 < npts; h0 += B.M) {
      const int cntp = min(B.M, npts - h0);
      int qn = 0, nzn = 0, hits = 0;
      const long long t_a = clock64();
      for (int base = 0; base < cntp; base += 32) {
        const int li = base + lane;
        int cls = 0;
        double dis = 0.0;
        if (li < cntp) {
          const float4 c = W.cache[h0 + li];   // both cache entries are requested before the fp64 classification chain
          const uint4 pc = W.pcache[h0 + li];
          const d3 pm = mirrored(B, PP, h0 + li);
          cls = classify(B, pm, pc, W.pcache + h0 + li);
          // NN cache: a bit-identical float query has the same exact nearest neighbour (the +-1e-9 finite-difference perturbations
          // of the plane often leave the float-rounded mirrored point unchanged), so the brute-force scan is skipped.
          if (cls == 1 && __float_as_uint(c.x) == __float_as_uint((float)pm.x) && __float_as_uint(c.y) == __float_as_uint((float)pm.y) &&
              __float_as_uint(c.z) == __float_as_uint((float)pm.z)) {
            dis = dist_to_neighbour(B, pm, __float_as_int(c.w));
            cls = 3;
          }
        }
        const unsigned need = __ballot_sync(0xffffffffu, cls == 1);
        const unsigned hit = __ballot_sync(0xffffffffu, cls == 3);
        const int pos = nzn + __popc((need | hit) & lt);    // slot in the ordered list of points that carry a distance
        if (cls == 3) W.dis[pos] = dis;
        if (cls == 1) W.queue[qn + __popc(need & lt)] = (unsigned short)(li | (pos << 8));   // li, pos < 256
        qn += __popc(need); nzn += __popc(need | hit); hits += __popc(hit);
        invalid += __popc(__ballot_sync(0xffffffffu, cls == 2));
      }
      __syncwarp();
      const long long t_b = clock64();
      if (lane == 0) { W.n_evals++; W.n_queries += qn; W.n_points += cntp; W.n_hits += hits; }
      for (int qb = 0; qb < qn; qb += 32) {
        const int cq = min(32, qn - qb);
        if (cq > 16) {
          if (lane < cq) {
            const int e = W.queue[qb + lane], li = e & 255;
            const d3 pm = mirrored(B, PP, h0 + li);
            const NNRes r = nn_distance(B, pm);
            W.dis[e >> 8] = r.dis;
            if (update_cache) W.cache[h0 + li] = make_float4((float)pm.x, (float)pm.y, (float)pm.z, __int_as_float(r.best));
          }
        } else {   // partial chunk: split every query's scan over several lanes
          int slots = 1;
          while (slots < cq) slots <<= 1;
          const int q = lane % slots;
          const bool active = q < cq;
          const int e = W.queue[qb + (active ? q : 0)], li = e & 255;
          const d3 pm = mirrored(B, PP, h0 + li);
          const NNRes r = nn_distance_split(B, pm, slots, active);
          if (active && lane < slots) {
            W.dis[e >> 8] = r.dis;
            if (update_cache) W.cache[h0 + li] = make_float4((float)pm.x, (float)pm.y, (float)pm.z, __int_as_float(r.best));
          }
        }
      }
      __syncwarp();
      const long long t_c = clock64();
      for (int base = 0; base < nzn; base += 32) {
        const double dis = (base + lane < nzn) ? W.dis[base + lane] : 0.0;
        const double lnP = -0.5 * sinv * sinv * dis * dis;
        const int cnt = min(32, nzn - base);
        for (int k = 0; k < cnt; k++) ln_total += __shfl_sync(0xffffffffu, lnP, k);
      }
      __syncwarp();
      if (lane == 0) { const long long t_d = clock64(); W.clk[0] += t_b - t_a; W.clk[1] += t_c - t_b; W.clk[2] += t_d - t_c; }
    }
  } else {
    const double sinv = B.sinv;
    for (int base = 0; base < npts; base += 32) {
      const int idx = base + lane;
      double lnP = 0.0; int cls = 0;
      if (idx < npts) {
        const d3 pm = mirrored(B, PP, idx);
        uint4 scratch;
        cls = classify(B, pm, make_uint4(0xffffffffu, 0u, 0u, 0u), &scratch);
        const double dis = (cls == 1) ? nn_distance(B, pm).dis : 0.0;
        lnP = -0.5 * sinv * sinv * dis * dis;
      }
      const int cnt = min(32, npts - base);
      for (int k = 0; k < cnt; k++) ln_total += __shfl_sync(0xffffffffu, lnP, k);
      invalid += __popc(__ballot_sync(0xffffffffu, cls == 2));
    }
  }
  const int num_valid = npts - invalid;
  const double aver = (num_valid > 0) ? (ln_total / (double)num_valid) : -INFINITY;
  return -aver;
}
// evaluation of a single plane: lane 0 prepares the pair
__device__ __forceinline__ double sym_eval_plane(const SymBlock &B, SymWarp &W, const dplane *pl) {
  if ((threadIdx.x & 31) == 0) make_pair(W.pp, *pl);
  __syncwarp();
  return sym_eval(B, W, W.pp);
}

__device__ __noinline__ void sym_oplus(dplane &p, const double *u, int symtype) {
  if (symtype == 1) plane_oplus_common(p, u[0], 0.0, u[1]);           // VertexPlane: (yaw, 0, dis)   :527-532, Plane.h:46-59
  else { plane_oplus_common(p, u[0], 0.0, u[1]); p.dual += u[2]; }    // VertexDualPlane             :544-547, Plane.h:62-77
}

template <int DIM>
__device__ __noinline__ bool lm_solve_dim(const double *H, const double *b, double lambda, double *x) {
  double M[DIM * DIM];
  for (int i = 0; i < DIM * DIM; i++) M[i] = H[i];
  for (int i = 0; i < DIM; i++) M[i * DIM + i] += lambda;
  return ldlt_solve_small<DIM>(M, b, x);
}

// OptimizeSymmetryPlane / OptimizeSymmetryDualPlane: one vertex, one unary edge (information 1), optimize(iters).
// Warp-collective: lane 0 runs the scalar Levenberg-Marquardt bookkeeping on the shared-memory state, all lanes evaluate costs.
// Two exact economies over the literal g2o call sequence: (1) computeActiveErrors() at the top of an iteration re-evaluates the plane
// whose error is already known (the accepted trial, or the restored backup) -- the cost is a pure function, so the value is reused;
// (2) the 2*dim finite-difference planes of a linearisation are prepared (oplus + mirror constants) by 2*dim lanes at once.
__device__ void sym_optimize(const SymBlock &B, SymWarp &W, int iters) {
  const int lane = threadIdx.x & 31;
  const int symtype = B.symtype;
  const int dim = (symtype == 1) ? 2 : 3;
  double err = sym_eval_plane(B, W, &W.est);   // value of the last computeError call
  double err_est = err;                        // error of the plane currently held in W.est
  if (lane == 0) { W.n_ref = 1; W.res.init_err = err; W.lambda = -1; W.ni = 2; W.nBad = 0; W.x[0] = W.x[1] = W.x[2] = 0; W.ok = 1; }
  __syncwarp();
  for (int it = 0; it < iters; it++) {
    if (!W.ok) break;
    err = err_est;                              // computeActiveErrors() at W.est
    if (lane == 0) { W.n_ref += 1 + 2 * dim; W.currentChi = err * (1.0 * err); W.tempChi = W.currentChi; W.iniChi = W.currentChi; }
    // linearizeOplus (numeric, delta 1e-9) + constructQuadraticForm
    if (lane < 2 * dim) {
      const int d = lane >> 1;
      const double del = (lane & 1) ? -1e-9 : 1e-9;
      double add[3];
      add[0] = (d == 0) ? del : 0.0; add[1] = (d == 1) ? del : 0.0; add[2] = (d == 2) ? del : 0.0;
      dplane t = W.est;
      sym_oplus(t, add, symtype);
      make_pair(W.fd[lane], t);
    }
    __syncwarp();
    for (int d = 0; d < dim; d++) {
      // the finite-difference evaluations leave the NN cache on the queries of the base plane: their own queries are the base query or
      // one float ulp beside it, so comparing against the base hits more often than comparing against the previous perturbation
      const double e1 = sym_eval(B, W, W.fd[2 * d], false);
      const double e2 = sym_eval(B, W, W.fd[2 * d + 1], false);
      if (lane == 0) W.J[d] = (1.0 / (2 * 1e-9)) * (e1 - e2);
    }
    if (lane == 0) {
      for (int i = 0; i < dim; i++) {
        W.b[i] = 0.0 - (W.J[i] * 1.0) * err;
        for (int j = 0; j < dim; j++) W.H[i * dim + j] = 0.0 + (W.J[i] * 1.0) * W.J[j];
      }
      if (it == 0) {
        double mx = 0;
        for (int j = 0; j < dim; j++) mx = fmax(fabs(W.H[j * dim + j]), mx);
        W.lambda = 1e-5 * mx; W.ni = 2; W.nBad = 0;
      }
      W.rho = 0; W.qmax = 0;
    }
    __syncwarp();
    while (true) {
      __syncwarp();   // every lane has read W.again / W.accepted of the previous trial before lane 0 overwrites them (racecheck: WAR)
      if (lane == 0) {
        W.backup = W.est;
        double x[3] = {W.x[0], W.x[1], W.x[2]};
        bool ok2 = (dim == 2) ? lm_solve_dim<2>(W.H, W.b, W.lambda, x) : lm_solve_dim<3>(W.H, W.b, W.lambda, x);
        W.x[0] = x[0]; W.x[1] = x[1]; W.x[2] = x[2];
        W.again = ok2 ? 1 : 0;   // reused as "ok2" until the decision below
        sym_oplus(W.est, x, symtype);
      }
      __syncwarp();
      err = sym_eval_plane(B, W, &W.est);
      if (lane == 0) {
        const bool ok2 = W.again != 0;
        W.n_ref++;
        double tempChi = err * (1.0 * err);
        if (!ok2) tempChi = DBL_MAX;
        double rho = (W.currentChi - tempChi);
        double scale = 0;
        for (int j = 0; j < dim; j++) scale += W.x[j] * (W.lambda * W.x[j] + W.b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && isfinite(tempChi)) {
          double alpha = 1. - pow((2 * rho - 1), 3.0);
          alpha = fmin(alpha, 2. / 3.);
          const double sf = fmax(1. / 3., alpha);
          W.lambda *= sf; W.ni = 2; W.currentChi = tempChi; W.accepted = 1;
        } else {
          W.lambda *= W.ni; W.ni *= 2; W.est = W.backup; W.accepted = 0;
        }
        W.tempChi = tempChi; W.rho = rho; W.qmax++;
        W.again = (rho < 0 && W.qmax < 10) ? 1 : 0;
        if (!W.again) {
          if (W.qmax == 10 || rho == 0) W.ok = 0;
          else {
            if ((W.iniChi - W.currentChi) * 1e3 < W.iniChi) W.nBad++; else W.nBad = 0;
            if (W.nBad >= 3) W.ok = 0;
          }
        }
      }
      __syncwarp();
      if (W.accepted) err_est = err;   // W.est now holds the evaluated trial; otherwise it was restored to the backup (err_est unchanged)
      if (!W.again) break;
    }
  }
  if (lane == 0) {
    W.res.final_err = err;          // e->error() after optimize(): value of the LAST computeError call (possibly a rejected trial)
    W.res.prob = exp(-err);
    W.res.plane = W.est;
    atomicAdd(&B.st->n_evals, (unsigned long long)W.n_evals); atomicAdd(&B.st->n_queries, (unsigned long long)W.n_queries);
    atomicAdd(&B.st->n_points, (unsigned long long)W.n_points); atomicAdd(&B.st->n_hits, (unsigned long long)W.n_hits);
    atomicAdd(&B.st->n_ref_evals, (unsigned long long)W.n_ref); atomicAdd(&B.st->n_ref_points, (unsigned long long)W.n_ref * (unsigned long long)B.npts);
  }
  __syncwarp();
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

// grid = 9 * n_obj one-warp CTAs: CTA b handles seed (b % 9) of object (b / 9).  Seeds are independent until the very end; the warp
// that finishes LAST for an object (atomic counter) performs the completion/finalisation, so no warp ever waits at a barrier for a
// slower seed and the SM keeps ~28 independent warps in flight (the kernel is latency / issue bound, not bandwidth bound).
// S = seed warps per CTA (1 or 3; the warps of a CTA work on the same object and share its shared-memory cloud), MINB = CTAs per SM.
template <int S, int MINB>
__global__ void __launch_bounds__(32 * S, MINB) symmetry_finalize_kernel(const uint16_t *__restrict__ depth, size_t pitch, size_t frame_stride,
                                                                         const ObjMeta *__restrict__ metas, ObjState *__restrict__ states, ExtractParamsDev prm,
                                                                         ExtractPtrs W, EslamEllipsoidOut *__restrict__ outs, int obj0) {
  static_assert(9 % S == 0, "the seed warps of a CTA must belong to one object");
  __shared__ SymBlock B;
  __shared__ SymWarp Wps[S];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  SymWarp &Wp = Wps[wid];
  const int gseed = blockIdx.x * S + wid;
  const int obj = gseed / 9, seed = gseed - obj * 9;
  const ObjMeta &m = metas[obj];
  ObjState &st = states[obj];
  EslamEllipsoidOut &out = outs[obj];
  const int status = st.status;
  const int run_sym = st.run_sym;
  if (seed == 0 && lane == 0) {
    for (int i = 0; i < 10; i++) out.ellipsoid[i] = 0;
    out.prob = 0; out.status = status; out.symmetry_type = m.symtype > 0 ? m.symtype : 0; out.symmetry_result = 0; out.best_seed = -1;
    for (int i = 0; i < 4; i++) { out.sym_plane_world[i] = 0; out.sym_plane2_world[i] = 0; }
    for (int i = 0; i < 3; i++) out.sym_center_world[i] = 0;
    out.n_crop = st.n_crop; out.n_voxel = st.n_voxel; out.n_plane = st.n_plane; out.n_clusters = st.n_clusters; out.n_cluster = st.n_cluster;
    out.n_sym = st.n_sym; out.n_final = 0; out.reserved = 0;
    if (!prm.symmetry_open) out.symmetry_type = 0;
  }
  if (status != 0) return;
  if (!run_sym) {   // no symmetry for this label: seed 0 finalises directly
    if (seed == 0) { __syncwarp(); finalize_object(m, st, prm, W, out); }
    return;
  }
  const int M = st.n_sym;
  const double *objn = W.objn + 3 * (size_t)m.off;
  const float4 *objnf = W.objnf + m.off;
  // cache slot of this resident warp: a free bit of this SM's mask (at most 32 CTAs are resident per SM)
  unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if (smid >= (unsigned)kMaxSmId) __trap();
  unsigned *slotword = W.slotmask + smid;
  int slotbit = -1;
  while (slotbit < 0) {   // warp-uniform control flow: lane 0 reads and claims, every lane follows the broadcast outcome
    unsigned cur = 0;
    if (lane == 0) cur = *((volatile unsigned *)slotword);
    cur = __shfl_sync(0xffffffffu, cur, 0);
    const unsigned freeb = ~cur;
    if (freeb) {
      const int b = __ffs(freeb) - 1;
      unsigned old = 0;
      if (lane == 0) old = atomicOr(slotword, 1u << b);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (!(old & (1u << b))) slotbit = b;
    }
  }
  const size_t slot = (size_t)smid * 32 + slotbit;
  float4 *cache = W.nncache + slot * (2 * kSymQueue);
  uint4 *pcache = W.pixcache + slot * (2 * kSymQueue);
  {
    const bool in_smem = M <= kSymSmemPts;
    // invalidate this warp's NN cache: a NaN x never equals a float-converted finite query
    if (M <= kSymQueue) for (int j = lane; j < 2 * M; j += 32) {
      cache[j] = make_float4(__uint_as_float(0x7fffffffu), 0.0f, 0.0f, 0.0f);
      pcache[j] = make_uint4(0xffffffffu, 0u, 0u, 0u);   // no pixel has x = y = 65535
    }
    if (in_smem) for (int j = threadIdx.x; j < (M + 1) / 2; j += 32 * S) {
      const float4 a = objnf[2 * j];
      const float far_ = 1.0e30f;   // (q - 1e30)^2 overflows to +inf: never < FLT_MAX, so the pad can never win
      const float4 b = (2 * j + 1 < M) ? objnf[2 * j + 1] : make_float4(far_, far_, far_, 0.0f);
      B.X2[j] = f2_pack(a.x, b.x); B.Y2[j] = f2_pack(a.y, b.y); B.Z2[j] = f2_pack(a.z, b.z);
    }
    if (lane == 0) { Wp.cache = cache; Wp.pcache = pcache; Wp.n_evals = Wp.n_queries = Wp.n_points = Wp.n_hits = 0; Wp.accepted = 0; Wp.clk[0] = Wp.clk[1] = Wp.clk[2] = Wp.clk[3] = 0; }
    if (threadIdx.x == 0) {
      B.img = depth + (size_t)m.frame * frame_stride; B.pitch = pitch; B.st = &st;
      B.objn = objn; B.nn_g = objnf; B.M = M; B.symtype = m.symtype; B.npts = (m.symtype == 2) ? 2 * M : M; B.in_smem = in_smem ? 1 : 0;
      for (int i = 0; i < 12; i++) B.P[i] = st.P[i];
      B.cam[0] = st.T_oc[0]; B.cam[1] = st.T_oc[1]; B.cam[2] = st.T_oc[2];
      B.bx0 = st.bbox_i[0]; B.by1 = st.bbox_i[1]; B.bx2 = st.bbox_i[2]; B.by3 = st.bbox_i[3];
      B.rows = prm.rows; B.cols = prm.cols; B.fx = prm.fx; B.cx = prm.cx; B.cy = prm.cy; B.scale = prm.scale;
      B.sinv = 1.0 / prm.sigma;
    }
    if (lane == 0) {
      // GenerateInitPlanes (Symmetry.cpp:44-86): seed = i*3+m, dis = -0.2 + 0.2 i, angle = -5deg + 5deg m
      const double diff_dis = 0.2, diff_angle = M_PI / 180.0 * 5;
      const double start_dis = -diff_dis * 1, start_angle = -diff_angle * 1;
      const int si = seed / 3, sm = seed % 3;
      const double dis = start_dis + diff_dis * si, angle = start_angle + diff_angle * sm;
      Wp.est.p[0] = sin(angle); Wp.est.p[1] = -cos(angle); Wp.est.p[2] = 0; Wp.est.p[3] = -dis; Wp.est.dual = 0;   // fromDisAngleTrans
    }
    if (S > 1) __syncthreads(); else __syncwarp();
  }
  const long long t_seed0 = clock64();
  sym_optimize(B, Wp, prm.symmetry_iters);
  int prev = 0;
  if (lane == 0) {
    __threadfence();                          // this warp's cache stores are ordered before the next owner's
    atomicAnd(slotword, ~(1u << slotbit));    // hand the cache slot back
    Wp.clk[3] = clock64() - t_seed0 - Wp.clk[0] - Wp.clk[1] - Wp.clk[2];
    for (int k = 0; k < 4; k++) atomicAdd((unsigned long long *)&st.sym_clk[k], (unsigned long long)Wp.clk[k]);
    const SeedResult &r = Wp.res;
    st.seed_prob[seed] = r.prob; st.seed_init[seed] = r.init_err; st.seed_final[seed] = r.final_err; st.seed_dual[seed] = r.plane.dual;
    for (int k = 0; k < 4; k++) st.seed_plane[seed][k] = r.plane.p[k];
    __threadfence();
    prev = atomicAdd(&st.done, 1);
  }
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev == 8) {   // last seed of this object: every other seed's results are visible (fence + atomic)
    __threadfence();
    finalize_object(m, st, prm, W, out);
  }
}

int launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj,
                             const ExtractParamsDev &prm, const ExtractPtrs &ptrs) {
  ExtractWorkspace &w = *h->ex;
  if (n_obj <= 0) return ESLAM_OK;
  // Shape: one seed warp per CTA, 28 CTAs/SM (72 registers).  Measured alternatives (profiles/r01_kernel_shape_ab.txt): 3 seed warps per
  // CTA sharing the cloud at 27 warps/SM (+13 % time) and at 36 warps/SM with a 56-register cap (+18 %, spills) are both slower.
  if (!h->attr_symmetry_v1) { ESLAM_CUDA_OK(h, cudaFuncSetAttribute(symmetry_finalize_kernel<1, 28>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)); h->attr_symmetry_v1 = true; }
  symmetry_finalize_kernel<1, 28><<<9 * n_obj, 32, 0, s>>>(depth_dev, pitch, frame_stride, w.meta.as<ObjMeta>() + obj0, w.state.as<ObjState>() + obj0, prm, ptrs,
                                                           w.out.as<EslamEllipsoidOut>() + obj0, obj0);
  ESLAM_LAUNCHED(h);
  ESLAM_CUDA_OK(h, cudaGetLastError());
  return ESLAM_OK;
}

}  // namespace symv1
int symmetry_v1_launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj, const ExtractParamsDev &prm,
                       const ExtractPtrs &ptrs) {
  return symv1::launch(h, s, depth_dev, pitch, frame_stride, obj0, n_obj, prm, ptrs);
}
}  // namespace eslam

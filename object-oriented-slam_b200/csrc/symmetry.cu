// symmetry.cu -- symmetry-plane estimation (9 seeds x in-register Levenberg-Marquardt) + point-cloud completion +
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

constexpr int kSymSmemPts = 1024;

struct SymCtx {
  const uint16_t *img; size_t pitch;
  const double *objn; const float4 *nn;   // nn: float cloud (shared memory when M <= kSymSmemPts)
  int M, symtype;
  double P[12];
  d3 cam_t;
  int bx0, bx2, by1, by3;
  int rows, cols;
  double fx, cx, cy, scale, sinv;
};

__device__ __forceinline__ int x86_int(double a) { return (a > -2147483649.0 && a < 2147483648.0) ? (int)a : (int)0x80000000; }

// edge error (= -mean ln P) of plane `pl`; warp-collective, every lane returns the same value.
__device__ double sym_eval(const SymCtx &c, const dplane &pl) {
  const int lane = threadIdx.x & 31;
  double p2[4];
  plane_another(pl, p2);
  const int M = c.M;
  const int npts = (c.symtype == 2) ? 2 * M : M;
  // per-plane constants of GetSymmetryPointOfPlane hoisted out of the point loop (same values, same rounding)
  const double down1 = sqrt(pl.p[0] * pl.p[0] + pl.p[1] * pl.p[1] + pl.p[2] * pl.p[2]);
  const double n1x = pl.p[0] / down1, n1y = pl.p[1] / down1, n1z = pl.p[2] / down1;
  const double down2 = sqrt(p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]);
  const double n2x = p2[0] / down2, n2y = p2[1] / down2, n2z = p2[2] / down2;
  double ln_total = 0.0;
  int invalid = 0;
  for (int base = 0; base < npts; base += 32) {
    int idx = base + lane;
    double lnP = 0.0; int inv = 0;
    if (idx < npts) {
      int second = idx >= M;
      int j = idx - (second ? M : 0);
      d3 pt = mk3(c.objn[3 * j], c.objn[3 * j + 1], c.objn[3 * j + 2]);
      const double q0 = second ? p2[0] : pl.p[0], q1 = second ? p2[1] : pl.p[1], q2 = second ? p2[2] : pl.p[2], q3 = second ? p2[3] : pl.p[3];
      const double val = q0 * pt.x + q1 * pt.y + q2 * pt.z + q3 * 1.0;
      const double dis_pl = fabs(val) / (second ? down2 : down1);
      const double kk = 2 * ((val > 0) ? -1.0 : 1.0) * dis_pl;
      d3 pm = mk3(pt.x + kk * (second ? n2x : n1x), pt.y + kk * (second ? n2y : n1y), pt.z + kk * (second ? n2z : n1z));
      double uvh[3];
#pragma unroll
      for (int r = 0; r < 3; r++) { double s = c.P[r * 4] * pm.x; s += c.P[r * 4 + 1] * pm.y; s += c.P[r * 4 + 2] * pm.z; s += c.P[r * 4 + 3] * 1.0; uvh[r] = s; }
      double u = uvh[0] / uvh[2], v = uvh[1] / uvh[2];
      int x = x86_int(u), y = x86_int(v);
      bool valid = isfinite(pm.x) && isfinite(pm.y) && isfinite(pm.z);
      bool inrange = (c.bx0 < x && x < c.bx2) && (c.by1 < y && y < c.by3);
      if (inrange && (x < 0 || y < 0 || x >= c.cols || y >= c.rows)) inrange = false;
      bool need_nn = false;
      if (inrange) {
        unsigned short d0 = c.img[(size_t)y * c.pitch + x];
        double xi = (x - c.cx), yi = (y - c.cy), f = c.fx;
        double realz = (double)d0 * sqrt(xi * xi + f * f + yi * yi) / f;   // calculateProjZ, Symmetry.cpp:137-139
        unsigned short d = (unsigned short)(int)realz;                      // ushort(realz)
        if (d != 0) {
          double depth = d / c.scale;
          double ax = c.cam_t.x - pm.x, ay = c.cam_t.y - pm.y, az = c.cam_t.z - pm.z;
          double dis_cam = sqrt(ax * ax + ay * ay + az * az);
          if (!(dis_cam > depth)) { if (valid) need_nn = true; else inv = 1; }
        }
      } else {
        if (valid) need_nn = true; else inv = 1;
      }
      double dis = 0.0;
      if (need_nn) {
        float qx = (float)pm.x, qy = (float)pm.y, qz = (float)pm.z;
        float bd = FLT_MAX; int best = 0;
        for (int k = 0; k < M; k++) {
          float4 q = c.nn[k];
          float r = 0.0f, dd;
          dd = __fsub_rn(qx, q.x); r = __fadd_rn(r, __fmul_rn(dd, dd));
          dd = __fsub_rn(qy, q.y); r = __fadd_rn(r, __fmul_rn(dd, dd));
          dd = __fsub_rn(qz, q.z); r = __fadd_rn(r, __fmul_rn(dd, dd));
          if (r < bd) { bd = r; best = k; }
        }
        float4 q = c.nn[best];
        double dx = pm.x - (double)q.x, dy = pm.y - (double)q.y, dz = pm.z - (double)q.z;
        dis = sqrt(dx * dx + dy * dy + dz * dz);
      }
      lnP = -0.5 * c.sinv * c.sinv * dis * dis;
    }
    int cnt = min(32, npts - base);
    for (int k = 0; k < cnt; k++) ln_total += __shfl_sync(0xffffffffu, lnP, k);
    invalid += __popc(__ballot_sync(0xffffffffu, inv != 0));
  }
  int num_valid = npts - invalid;
  double aver = (num_valid > 0) ? (ln_total / (double)num_valid) : -INFINITY;
  return -aver;
}

__device__ __forceinline__ void sym_oplus(dplane &p, const double *u, int symtype) {
  if (symtype == 1) plane_oplus_common(p, u[0], 0.0, u[1]);           // VertexPlane: (yaw, 0, dis)   :527-532, Plane.h:46-59
  else { plane_oplus_common(p, u[0], 0.0, u[1]); p.dual += u[2]; }    // VertexDualPlane             :544-547, Plane.h:62-77
}

template <int DIM>
__device__ bool lm_solve_dim(const double *H, const double *b, double lambda, double *x) {
  double M[DIM * DIM];
  for (int i = 0; i < DIM * DIM; i++) M[i] = H[i];
  for (int i = 0; i < DIM; i++) M[i * DIM + i] += lambda;
  return ldlt_solve_small<DIM>(M, b, x);
}

struct SeedResult { double prob, init_err, final_err; dplane plane; };

// OptimizeSymmetryPlane / OptimizeSymmetryDualPlane: one vertex, one unary edge (information 1), optimize(iters)
__device__ SeedResult sym_optimize(const SymCtx &c, dplane est, int iters) {
  const int dim = (c.symtype == 1) ? 2 : 3;
  SeedResult R;
  double err = sym_eval(c, est);
  R.init_err = err;
  double lambda = -1, ni = 2; int nBad = 0;
  double H[9], b[3], x[3] = {0, 0, 0}, J[3];
  bool ok = true;
  for (int it = 0; it < iters && ok; it++) {
    err = sym_eval(c, est);
    double currentChi = err * (1.0 * err);
    double tempChi = currentChi;
    const double iniChi = currentChi;
    {  // linearizeOplus (numeric) + constructQuadraticForm
      const double delta = 1e-9, scalar = 1.0 / (2 * delta);
      double add[3] = {0, 0, 0};
      for (int d = 0; d < dim; d++) {
        dplane pp = est; add[d] = delta; sym_oplus(pp, add, c.symtype);
        double e1 = sym_eval(c, pp);
        dplane pm = est; add[d] = -delta; sym_oplus(pm, add, c.symtype);
        double e2 = sym_eval(c, pm);
        add[d] = 0.0;
        J[d] = scalar * (e1 - e2);
      }
      for (int i = 0; i < dim; i++) {
        b[i] = 0.0 - (J[i] * 1.0) * err;
        for (int j = 0; j < dim; j++) H[i * dim + j] = 0.0 + (J[i] * 1.0) * J[j];
      }
    }
    if (it == 0) {
      double mx = 0;
      for (int j = 0; j < dim; j++) mx = fmax(fabs(H[j * dim + j]), mx);
      lambda = 1e-5 * mx; ni = 2; nBad = 0;
    }
    double rho = 0; int qmax = 0;
    do {
      dplane backup = est;
      bool ok2 = (dim == 2) ? lm_solve_dim<2>(H, b, lambda, x) : lm_solve_dim<3>(H, b, lambda, x);
      sym_oplus(est, x, c.symtype);
      err = sym_eval(c, est);
      tempChi = err * (1.0 * err);
      if (!ok2) tempChi = DBL_MAX;
      rho = (currentChi - tempChi);
      double scale = 0;
      for (int j = 0; j < dim; j++) scale += x[j] * (lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3.0);
        alpha = fmin(alpha, 2. / 3.);
        double sf = fmax(1. / 3., alpha);
        lambda *= sf; ni = 2; currentChi = tempChi;
      } else {
        lambda *= ni; ni *= 2; est = backup;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0) ok = false;
    else {
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
      if (nBad >= 3) ok = false;
    }
  }
  R.final_err = err;          // e->error() after optimize(): value of the LAST computeError call (possibly a rejected trial)
  R.prob = exp(-err);
  R.plane = est;
  return R;
}

// grid = n_obj, block = 288 (9 warps = 9 seeds)
__global__ void __launch_bounds__(288, 2) symmetry_finalize_kernel(const uint16_t *__restrict__ depth, size_t pitch, size_t frame_stride,
                                                                const ObjMeta *__restrict__ metas, ObjState *__restrict__ states, ExtractParamsDev prm,
                                                                ExtractPtrs W, EslamEllipsoidOut *__restrict__ outs) {
  __shared__ BlockScratch S;
  __shared__ float4 s_nn[kSymSmemPts];
  __shared__ SeedResult s_seed[9];
  __shared__ int s_best;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ObjMeta m = metas[blockIdx.x];
  ObjState &st = states[blockIdx.x];
  EslamEllipsoidOut &out = outs[blockIdx.x];
  const int status = st.status;
  if (tid == 0) {
    for (int i = 0; i < 10; i++) out.ellipsoid[i] = 0;
    out.prob = 0; out.status = status; out.symmetry_type = m.symtype > 0 ? m.symtype : 0; out.symmetry_result = 0; out.best_seed = -1;
    for (int i = 0; i < 4; i++) { out.sym_plane_world[i] = 0; out.sym_plane2_world[i] = 0; }
    for (int i = 0; i < 3; i++) out.sym_center_world[i] = 0;
    out.n_crop = st.n_crop; out.n_voxel = st.n_voxel; out.n_plane = st.n_plane; out.n_clusters = st.n_clusters; out.n_cluster = st.n_cluster;
    out.n_sym = st.n_sym; out.n_final = 0; out.reserved = 0;
    if (!prm.symmetry_open) out.symmetry_type = 0;
  }
  if (status != 0) return;
  const int M = st.n_sym;
  const int run_sym = st.run_sym;
  const double *objn = W.objn + 3 * (size_t)m.off;
  const float4 *objnf = W.objnf + m.off;
  se3 Two = se3_load(st.T_wo);
  const se3 Twc = se3_load(m.Twc);
  double sym_prob = 1.0;
  dplane bestp; bestp.p[0] = 1; bestp.p[1] = 0; bestp.p[2] = 0; bestp.p[3] = 0; bestp.dual = 0;

  if (run_sym) {
    const bool in_smem = M <= kSymSmemPts;
    if (in_smem) for (int i = tid; i < M; i += blockDim.x) s_nn[i] = objnf[i];
    __syncthreads();
    SymCtx c;
    c.img = depth + (size_t)m.frame * frame_stride; c.pitch = pitch;
    c.objn = objn; c.nn = in_smem ? s_nn : objnf; c.M = M; c.symtype = m.symtype;
    for (int i = 0; i < 12; i++) c.P[i] = st.P[i];
    c.cam_t = mk3(st.T_oc[0], st.T_oc[1], st.T_oc[2]);
    c.bx0 = st.bbox_i[0]; c.by1 = st.bbox_i[1]; c.bx2 = st.bbox_i[2]; c.by3 = st.bbox_i[3];
    c.rows = prm.rows; c.cols = prm.cols; c.fx = prm.fx; c.cx = prm.cx; c.cy = prm.cy; c.scale = prm.scale;
    c.sinv = 1.0 / prm.sigma;
    // GenerateInitPlanes (Symmetry.cpp:44-86): seed = i*3+m, dis = -0.2 + 0.2 i, angle = -5deg + 5deg m
    {
      const double diff_dis = 0.2, diff_angle = M_PI / 180.0 * 5;
      const double start_dis = -diff_dis * 1, start_angle = -diff_angle * 1;
      int si = warp / 3, sm = warp % 3;
      double dis = start_dis + diff_dis * si, angle = start_angle + diff_angle * sm;
      dplane init; init.p[0] = sin(angle); init.p[1] = -cos(angle); init.p[2] = 0; init.p[3] = -dis; init.dual = 0;   // fromDisAngleTrans
      SeedResult r = sym_optimize(c, init, prm.symmetry_iters);
      if (lane == 0) {
        s_seed[warp] = r;
        st.seed_prob[warp] = r.prob; st.seed_init[warp] = r.init_err; st.seed_final[warp] = r.final_err; st.seed_dual[warp] = r.plane.dual;
        for (int k = 0; k < 4; k++) st.seed_plane[warp][k] = r.plane.p[k];
      }
    }
    __syncthreads();
    if (tid == 0) {  // sort(desc by prob) of 9 entries == insertion sort == stable: first maximum wins (Symmetry.cpp:122-126)
      int best = 0;
      for (int s = 1; s < 9; s++) if (s_seed[s].prob > s_seed[best].prob) best = s;
      s_best = best; st.best_seed = best;
    }
    __syncthreads();
    bestp = s_seed[s_best].plane;
    sym_prob = s_seed[s_best].prob;
  }

  // ---- completion :399-446 (mirrored points generated on the fly, never stored) ----
  double p1[4] = {bestp.p[0], bestp.p[1], bestp.p[2], bestp.p[3]}, p2[4];
  plane_another(bestp, p2);
  const int copies = run_sym ? ((m.symtype == 2) ? 4 : 2) : 1;
  const int nfin = copies * M;
  auto gen_point = [&](int idx) -> d3 {
    int part = idx / M, j = idx - part * M;
    d3 p = mk3(objn[3 * j], objn[3 * j + 1], objn[3 * j + 2]);
    if (part == 1) p = mirror_point(p, p1);                       // pSymCloudNormalized
    else if (part == 2) p = mirror_point(p, p2);                  // pSymCloudNormalized2_1
    else if (part == 3) p = mirror_point(mirror_point(p, p1), p2);  // pSymCloudNormalized2_2
    return p;
  };
  hom34 Tmo_h;
  bool retransform = false;
  if (run_sym) {
    double s3[3] = {0, 0, 0};
    for (int i = tid; i < nfin; i += blockDim.x) { d3 p = gen_point(i); s3[0] += p.x; s3[1] += p.y; s3[2] += p.z; }
    block_sum_double<3>(s3, S);
    double nn = (double)nfin;
    d3 cc = mk3(s3[0] / nn, s3[1] / nn, s3[2] / nn);   // GetPointcloudCenter
    d3 c0 = mk3(p1[0], p1[1], p1[2]);
    double n0 = norm3(c0);
    c0 = mk3(c0.x / n0, c0.y / n0, c0.z / n0);
    mat3 Rom; m3_setcol(Rom, 0, c0); m3_setcol(Rom, 2, mk3(0, 0, 1)); m3_setcol(Rom, 1, cross3(mk3(0, 0, 1), c0));
    se3 Tom; Tom.t = cc; Tom.r = q_from_R(Rom);
    Tmo_h = se3_hom(se3_inv(Tom));
    retransform = true;
    if (tid == 0) {
      d3 cw = hom_apply(se3_hom(Two), cc);
      out.sym_center_world[0] = cw.x; out.sym_center_world[1] = cw.y; out.sym_center_world[2] = cw.z;
      double w1[4]; plane_transform_rigid(Two, p1, w1);
      for (int k = 0; k < 4; k++) out.sym_plane_world[k] = w1[k];
      if (m.symtype == 2) { double w2[4]; plane_transform_rigid(Two, p2, w2); for (int k = 0; k < 4; k++) out.sym_plane2_world[k] = w2[k]; }
      out.symmetry_result = 1; out.best_seed = s_best;
    }
    Two = se3_mul(Two, Tom);
  }
  // ---- ProcessPCANormalized :495-531 : axis-aligned extents
  double ext[3] = {0, 0, 0};
  for (int i = tid; i < nfin; i += blockDim.x) {
    d3 p = gen_point(i);
    if (retransform) p = hom_apply(Tmo_h, p);
    ext[0] = fmax(ext[0], fabs(p.x)); ext[1] = fmax(ext[1], fabs(p.y)); ext[2] = fmax(ext[2], fabs(p.z));
    if (W.d_final) { double *q = W.d_final + 3 * ((size_t)4 * m.off + i); q[0] = p.x; q[1] = p.y; q[2] = p.z; }
  }
  for (int k = 0; k < 3; k++) ext[k] = -block_min_double(-ext[k], S);
  if (tid == 0) {
    ellipsoid e0;
    mat3 I = m3_eye();
    e0.pose.r = q_from_R(I); e0.pose.t = mk3(0, 0, 0);
    e0.s[0] = ext[0]; e0.s[1] = ext[1]; e0.s[2] = ext[2];
    ellipsoid eg = e0; eg.pose = se3_mul(Two, e0.pose);                 // transform_from(*pSE3Two)       :465
    ellipsoid el = eg; el.pose = se3_mul(se3_inv(Twc), eg.pose);        // transform_from(campose_wc.inverse()) :468-469
    ell_store(el, out.ellipsoid);
    out.prob = run_sym ? sym_prob : 1.0;
    out.n_final = nfin; st.n_final = nfin;
    out.status = ESLAM_OBJ_OK;
  }
}

int symmetry_finalize_launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int n_obj,
                             const ExtractParamsDev &prm, const ExtractPtrs &ptrs) {
  ExtractWorkspace &w = *h->ex;
  symmetry_finalize_kernel<<<n_obj, 288, 0, s>>>(depth_dev, pitch, frame_stride, w.meta.as<ObjMeta>(), w.state.as<ObjState>(), prm, ptrs,
                                                 w.out.as<EslamEllipsoidOut>());
  ESLAM_LAUNCHED(h);
  ESLAM_CUDA_OK(h, cudaGetLastError());
  return ESLAM_OK;
}

}  // namespace eslam

// plane.cu -- ground-plane extraction from full 640x480 depth frames, batched over frames.
//
// Replaces PlaneExtractor::extractPlanes / extractGroundPlane (src/plane/PlaneExtractor.cpp:22-107, :126-183) and the
// PCL algorithms behind them (PCL is an external, unpinned dependency of the reference -- see oracle/orc_plane.h for the
// restated upstream algorithms):
//   pcl::IntegralImageNormalEstimation (AVERAGE_3D_GRADIENT)      -> points_kernel, integral_kernel, distmap_kernel, normals_kernel
//   pcl::OrganizedConnectedComponentSegmentation + PlaneCoefficientComparator -> ccl_strip_kernel (shared-memory union-find per 16-row strip),
//                                                                    ccl_seam_kernel (strip boundaries, lock-free global union-find), ccl_flatten_kernel
//   computeMeanAndCovarianceMatrix<float> + pcl::eigen33          -> fit_kernel (one warp per region, sequential float sums in raster order)
//   OrganizedMultiPlaneSegmentation::refine (PlaneRefinementComparator) -> refine_kernel (row-sequential, scan-parallel inside a row)
//   selection rules of extractGroundPlane                         -> select_kernel
//
// Exactness notes (why some stages look sequential): the reference's result depends on (a) the exact recurrence order of PCL's
// double integral images, (b) float chamfer distances truncated with int(), (c) FLOAT accumulation of 9 moments over up to
// ~250k pixels per region (rel. error ~1e-3 in the reference itself!) and (d) scan-order label propagation.  Each of these is
// reproduced operation-for-operation, parallelised across frames / rows / regions instead of inside the dependent chain:
//   (a) anti-diagonal wavefront inside a warp: lane = image row of a 32-row band, 2 columns behind the lane above (shuffles),
//   (b) row-sequential chamfer passes; the in-row recurrence is evaluated exactly with a 13-step window,
//   (c) one warp per region walks the label image in raster order; lanes 0..8 own one float accumulator each,
//   (d) per row: two block-wide max-scans resolve the left-to-right (right-to-left) propagation, vertical checks are parallel.
// HBM layout per frame: pts float4[NP] (x,y,z,-), nrm float4[NP] (nx,ny,nz,plane_d), iiX/iiY double[3][(H+1)][pitch(W+1)] (planar),
// dm float[NP], parent/count/maxidx int[NP].  Algorithmic bytes: 2 B/pixel of depth in, 32 B out per frame.
#include "common.cuh"
#include "block_prims.cuh"
#include "internal.h"
#include <math.h>

namespace eslam {

constexpr int kMaxRegions = 4096;

struct PlaneParamsDev {
  double fx, fy, cx, cy, scale;
  int rows, cols;
  float dis_th;        // Plane.DistanceThreshold (float)
  float ang_cos;       // cosf(0.017453 * AngTh)
  int min_size;        // Plane.MinSize
};

struct PlaneFramePtrs {
  float4 *pts, *nrm;
  double *iiX, *iiY;
  float *dm;
  int *parent, *count, *maxidx, *cmin, *cmax;   // union-find parents; per-root size, last pixel, column range (bounding box of the region)
  int *label;          // flattened region label of every pixel (= smallest pixel index of the region), relabelled by the refinement
  int *cand_root;      // [F * kMaxRegions]
  int *cand_n;         // [F]
  float4 *cand_coef;   // [F * kMaxRegions]
  int *cand_valid;     // [F * kMaxRegions]
  int *cand_size;      // [F * kMaxRegions]  inlier count (grows during refinement)
  double *out_plane;   // [F*4]
  int *out_found, *out_inl;
  unsigned *cpart;     // [F * blocks * 6]: per-block partial ranges written by points_kernel
  unsigned *crange;    // [F * 6]: float bits of the smallest non-zero |x|, |y|, |z| and of the largest |x|, |y|, |z| of the frame's points
  int exact_mode;      // 1: frames whose integral images are provably exact take the fused tile kernel, 0: every frame takes the recurrence
};

struct PlaneWorkspace {
  DevBuf pts, nrm, iiX, iiY, dm, parent, label, count, maxidx, cmin, cmax, cand_root, cand_n, cand_coef, cand_valid, cand_size, out_plane, out_found, out_inl, crange, cpart;
  PinBuf pin;
  int chunk = 0;
};

// ---------------------------------------------------------------------------------------------- P1: back-projection
__global__ void points_kernel(const uint16_t *__restrict__ depth, size_t pitch, size_t frame_stride, PlaneParamsDev prm, PlaneFramePtrs W) {
  const int NP = prm.rows * prm.cols;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int f = blockIdx.y;
  if (idx >= NP) idx = NP - 1;   // the tail threads redo the last pixel (same values): the warp reductions below need full warps
  int m = idx / prm.cols, n = idx - m * prm.cols;
  unsigned short dv = depth[(size_t)f * frame_stride + (size_t)m * pitch + n];
  double d = dv / prm.scale;
  float z = (float)d;                                   // PointT p; p.z = d   (zero depth kept as (0,0,0), PlaneExtractor.cpp:36-47)
  float x = (float)((n - prm.cx) * (double)z / prm.fx);
  float y = (float)((m - prm.cy) * (double)z / prm.fy);
  W.pts[(size_t)f * NP + idx] = make_float4(x, y, z, 0.0f);
  {   // per-block range of the coordinates -> W.cpart[f][block][6] (crange_kernel folds the blocks): decides frame_ii_exact()
    __shared__ unsigned s_rng[8][6];
    unsigned mn[3], mx[3];
    const float c3[3] = {fabsf(x), fabsf(y), fabsf(z)};
#pragma unroll
    for (int k = 0; k < 3; k++) { const unsigned b = __float_as_uint(c3[k]); mx[k] = b; mn[k] = b ? b : 0xffffffffu; }   // positive floats order like their bits
#pragma unroll
    for (int k = 0; k < 3; k++) { mn[k] = __reduce_min_sync(0xffffffffu, mn[k]); mx[k] = __reduce_max_sync(0xffffffffu, mx[k]); }
    const int wp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) { s_rng[wp][k] = mn[k]; s_rng[wp][3 + k] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      unsigned v = s_rng[0][threadIdx.x];
      for (int q = 1; q < (int)(blockDim.x >> 5); q++) v = threadIdx.x < 3 ? min(v, s_rng[q][threadIdx.x]) : max(v, s_rng[q][threadIdx.x]);
      W.cpart[((size_t)f * gridDim.x + blockIdx.x) * 6 + threadIdx.x] = v;
    }
  }
  W.count[(size_t)f * NP + idx] = 0;
  W.maxidx[(size_t)f * NP + idx] = 0;
  W.cmin[(size_t)f * NP + idx] = 0x7fffffff;
  W.cmax[(size_t)f * NP + idx] = -1;
}

// folds the per-block coordinate ranges of a frame: 6 warps, warp q owns quantity q (min |x| |y| |z|, max |x| |y| |z|)
__global__ void crange_kernel(PlaneFramePtrs W, int nblocks) {
  const int f = blockIdx.x, q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned v = q < 3 ? 0xffffffffu : 0u;
  for (int b = lane; b < nblocks; b += 32) { const unsigned u = W.cpart[((size_t)f * nblocks + b) * 6 + q]; v = q < 3 ? min(v, u) : max(v, u); }
  v = q < 3 ? __reduce_min_sync(0xffffffffu, v) : __reduce_max_sync(0xffffffffu, v);
  if (lane == 0) W.crange[6 * f + q] = v;
}

// ---------------------------------------------------------------------------------------------- exactness of the integral images
// Every addend of an integral image is a float difference of two coordinates of the same component, hence an integer multiple of
// u = ulp(smallest non-zero |coordinate|) of that component; every entry and every intermediate of PCL's recurrence
// ((II(r,c+1) + II(r+1,c)) - II(r,c)) + v is bounded by 2 B + V with B <= max(H, W) * 4 * Cmax (the row / column sums telescope up to
// the float rounding of the differences) and V <= 2 Cmax.  When (2 B + V) / u < 2^52 for x, y and z every one of those doubles is
// exactly representable: NO rounding happens anywhere in the recurrence, the integral images are the exact sums, and a box sum
// ((lr + ul) - ur) - ll equals the exact sum of the addends inside the box -- whatever the summation order and whatever the origin of
// the integral image.  Such frames (every realistic frame: depths of 8 cm and more with the TUM intrinsics) take normals_exact_kernel,
// which builds TILE-LOCAL integral images in shared memory; the others keep the wavefront recurrence + normals_kernel.
__device__ __forceinline__ bool frame_ii_exact(const unsigned *crange, int f, int H, int Wd) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const unsigned mnb = crange[6 * f + k], mxb = crange[6 * f + 3 + k];
    if (mnb == 0xffffffffu) continue;                          // the component is zero everywhere
    if ((mxb >> 23) >= 255u || (mnb >> 23) == 0u) { ok = false; continue; }   // inf / NaN / subnormal coordinates
    const double cmax = (double)__uint_as_float(mxb);
    const double bound = 2.0 * (double)max(H, Wd) * 4.0 * cmax * 1.01 + 2.0 * cmax;
    const double ulp = scalbn(1.0, (int)(mnb >> 23) - 127 - 23);
    ok = ok && (bound < ulp * 4503599627370496.0);            // 2^52
  }
  return ok;
}
// One CTA per 32 x 32 output tile: the six gradient channels of the (32 + 10)^2 pixel window as tile-local integral images in
// shared memory (exact doubles), then pcl::IntegralImageNormalEstimation's AVERAGE_3D_GRADIENT normal per pixel -- the integral
// images of these frames never exist in HBM.
constexpr int kNxT = 32, kNxHalo = 5, kNxW = kNxT + 2 * kNxHalo, kNxP = kNxW + 1;   // window 42, integral image 43 x 43 (odd pitch: no bank conflicts)
constexpr int kNxPW = kNxW + 3;                                                       // coordinate tile: window + 1 pixel around it, pitch 45 (odd)
constexpr int kNxSmemBytes = 6 * kNxP * kNxP * (int)sizeof(double) + 3 * (kNxW + 2) * kNxPW * (int)sizeof(float);
__global__ void __launch_bounds__(512, 2) normals_exact_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  extern __shared__ __align__(16) unsigned char nx_smem[];
  double *II = reinterpret_cast<double *>(nx_smem);                                   // [6][43][43]
  float *P = reinterpret_cast<float *>(nx_smem + 6 * kNxP * kNxP * sizeof(double));   // [3][44][45]
  const int H = prm.rows, Wd = prm.cols, NP = H * Wd;
  const int f = blockIdx.z;
  if (!W.exact_mode || !frame_ii_exact(W.crange, f, H, Wd)) return;
  const int R0 = blockIdx.y * kNxT - kNxHalo, C0 = blockIdx.x * kNxT - kNxHalo;   // first pixel of the window
  const float4 *pts = W.pts + (size_t)f * NP;
  constexpr int PW = kNxPW, PH = kNxW + 2, PS = PH * PW;
  for (int i = threadIdx.x; i < PH * (kNxW + 2); i += blockDim.x) {
    const int rr = i / (kNxW + 2), cc = i - rr * (kNxW + 2);
    const int pr = R0 - 1 + rr, pc = C0 - 1 + cc;
    float4 p = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (pr >= 0 && pr < H && pc >= 0 && pc < Wd) p = pts[(size_t)pr * Wd + pc];
    const int o = rr * PW + cc;
    P[o] = p.x; P[PS + o] = p.y; P[2 * PS + o] = p.z;
  }
  for (int t = threadIdx.x; t < 6 * kNxP; t += blockDim.x) II[(size_t)(t / kNxP) * kNxP * kNxP + (t % kNxP)] = 0.0;   // row 0 of every channel
  __syncthreads();
  // row prefix sums of the addends: thread = (channel, window row), lanes = rows; channel k < 3: d/dx of component k, else d/dy of k - 3.
  // Addends are zero outside image rows 1 .. H-2 / columns 1 .. W-2 (IntegralImageNormalEstimation::initAverage3DGradientMethod).
  for (int t = threadIdx.x; t < 6 * kNxW; t += blockDim.x) {
    const int ch = t / kNxW, i = t - ch * kNxW, k = ch % 3;
    const int r = R0 + i;
    const float *Pk = P + k * PS + (i + 1) * PW + 1;   // Pk[j] = component k of window pixel (i, j)
    double *row = II + ((size_t)ch * kNxP + (i + 1)) * kNxP;
    const bool rowok = r >= 1 && r <= H - 2;
    const int jlo = rowok ? max(0, 1 - C0) : kNxW, jhi = rowok ? min(kNxW, Wd - 1 - C0) : kNxW;   // columns with a non-zero addend: [jlo, jhi)
    double acc = 0.0;
    row[0] = 0.0;
    int j = 0;
    for (; j < jlo && j < kNxW; j++) row[j + 1] = 0.0;
    if (ch < 3) {
      float left = (j < kNxW) ? Pk[j - 1] : 0.0f, mid = (j < kNxW) ? Pk[j] : 0.0f;
#pragma unroll 6
      for (; j < jhi; j++) {
        const float right = Pk[j + 1];
        acc = __dadd_rn(acc, (double)__fsub_rn(right, left));
        row[j + 1] = acc;
        left = mid; mid = right;
      }
    } else {
#pragma unroll 6
      for (; j < jhi; j++) {
        acc = __dadd_rn(acc, (double)__fsub_rn(Pk[j + PW], Pk[j - PW]));
        row[j + 1] = acc;
      }
    }
    for (; j < kNxW; j++) row[j + 1] = acc;
  }
  __syncthreads();
  // column prefix sums: thread = (channel, column), lanes = columns
  for (int t = threadIdx.x; t < 6 * kNxW; t += blockDim.x) {
    const int ch = t / kNxW, j = t - ch * kNxW;
    double *col = II + (size_t)ch * kNxP * kNxP + kNxP + (j + 1);
    double acc = 0.0;
#pragma unroll 6
    for (int i = 0; i < kNxW; i++) { acc = __dadd_rn(acc, col[(size_t)i * kNxP]); col[(size_t)i * kNxP] = acc; }
  }
  __syncthreads();
  const float nanf_ = __int_as_float(0x7fc00000);
  const int border = 10;   // int(normal_smoothing_size_)
  for (int t = threadIdx.x; t < kNxT * kNxT; t += blockDim.x) {
    const int ti = t / kNxT, tj = t - ti * kNxT;
    const int ri = blockIdx.y * kNxT + ti, ci = blockIdx.x * kNxT + tj;
    if (ri >= H || ci >= Wd) continue;
    const size_t idx = (size_t)ri * Wd + ci;
    const int pi = (ti + kNxHalo + 1) * PW + (tj + kNxHalo + 1);
    const float px = P[pi], py = P[PS + pi], pz = P[2 * PS + pi];
    float nx = nanf_, ny = nanf_, nz = nanf_;
    if (ri >= border && ri < H - border && ci >= border && ci < Wd - border && isfinite(pz)) {
      const float smoothing = fminf(W.dm[(size_t)f * NP + idx], 10.0f);
      if (smoothing > 2.0f) {
        const int rw = (int)smoothing, rw2 = rw / 2;
        const int sy = ti + kNxHalo - rw2, sx = tj + kNxHalo - rw2;   // window coordinates of the box
        const int ul = sy * kNxP + sx, ur = ul + rw, ll = (sy + rw) * kNxP + sx, lr = ll + rw;
        double g[6];
#pragma unroll
        for (int ch = 0; ch < 6; ch++) {
          const double *q = II + (size_t)ch * kNxP * kNxP;
          g[ch] = __dsub_rn(__dsub_rn(__dadd_rn(q[lr], q[ul]), q[ur]), q[ll]);
        }
        const double *gx = g, *gy = g + 3;
        const double n0 = __dsub_rn(__dmul_rn(gy[1], gx[2]), __dmul_rn(gy[2], gx[1]));   // gradient_y.cross(gradient_x)
        const double n1 = __dsub_rn(__dmul_rn(gy[2], gx[0]), __dmul_rn(gy[0], gx[2]));
        const double n2 = __dsub_rn(__dmul_rn(gy[0], gx[1]), __dmul_rn(gy[1], gx[0]));
        const double nl = __dadd_rn(__dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1)), __dmul_rn(n2, n2));
        if (nl != 0.0) {
          const double s = sqrt(nl);
          nx = (float)(n0 / s); ny = (float)(n1 / s); nz = (float)(n2 / s);
          const float vx = __fsub_rn(0.0f, px), vy = __fsub_rn(0.0f, py), vz = __fsub_rn(0.0f, pz);   // flipNormalTowardsViewpoint, vp = 0
          const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
          if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
        }
      }
    }
    const float pd = __fadd_rn(__fadd_rn(__fmul_rn(px, nx), __fmul_rn(py, ny)), __fmul_rn(pz, nz));   // plane_d = p . n
    W.nrm[(size_t)f * NP + idx] = make_float4(nx, ny, nz, pd);
  }
}

// ---------------------------------------------------------------------------------------------- P2: integral images (exact recurrence)
// II(r+1,c+1) = ((II(r,c+1) + II(r+1,c)) - II(r,c)) + x(r,c)   in double, IntegralImage2D<float,3>::computeIntegralImages.
// The rounding of every entry depends on the recurrence order, so the dependent chain is kept and parallelism comes from elsewhere:
//   * the 6 channels (d/dx and d/dy of x, y, z) are independent: one warp per (frame, channel), 6 warps per CTA share the frame in L1;
//   * inside a warp lane l owns image row r0 + l of a 32-row band and runs TWO columns behind lane l-1, so the value it needs from the
//     row above is exactly lane l-1's previous result: one shuffle, issued a step ahead (off the 3-DADD critical path), no shared memory,
//     no barrier; bands are processed top to bottom, lane 0 takes the last row of the band above from the image just written
//     (coalesced 32-column chunks fetched one block ahead);
//   * the channel's point-difference input is prefetched 16 steps ahead into registers.
// Layout: planar, row pitch padded to 4 doubles: channel k of frame f at ii + ((f*3 + k) * (H+1)) * IWp; results leave as 16-byte pairs.
__device__ __host__ __forceinline__ int ii_pitch(int Wd) { return (Wd + 1 + 3) & ~3; }
__device__ __forceinline__ void st_global_256(double *p, double a, double b, double c, double d) {   // one full 32-byte sector (STG.256)
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
constexpr int kIntRing = 128;            // columns of the per-warp input ring (4 blocks of 32: lanes span 62 columns + the block in flight)
constexpr int kIntRingPitch = kIntRing + 1;   // odd pitch: lane l reads column (t - 2l) of row l -> bank (t - l) mod 32, conflict free
constexpr int kIntWarps = 6;
__global__ void __launch_bounds__(32 * kIntWarps) integral_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  extern __shared__ float s_ring[];      // [kIntWarps][32 rows][kIntRingPitch]: this channel's point differences x(r, c), staged coalesced
  const int H = prm.rows, Wd = prm.cols, NP = H * Wd, IWp = ii_pitch(Wd);
  const int f = blockIdx.x, ch = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (W.exact_mode && frame_ii_exact(W.crange, f, H, Wd)) return;   // this frame's integral images stay in shared memory (normals_exact_kernel)
  const int k = ch % 3;
  const bool isX = ch < 3;
  const float *ptsf = reinterpret_cast<const float *>(W.pts + (size_t)f * NP) + k;   // component k of pixel i at ptsf[4 * i]
  double *ii = (isX ? W.iiX : W.iiY) + ((size_t)f * 3 + k) * (size_t)(H + 1) * IWp;
  float *ring = s_ring + (size_t)ch * 32 * kIntRingPitch;
  for (int j = lane; j < IWp; j += 32) ii[j] = 0.0;   // row 0 of the integral image
  const int steps = Wd + 62;                          // lane 31 starts 62 steps after lane 0
  const int offA = isX ? 1 : 0, offB = isX ? -1 : 0, rA = isX ? 0 : 1, rB = isX ? 0 : -1;
  // x(row, c) of this channel for 8 rows of the band at column c = cb + lane (lane = column: 512-byte spans, 16 sectors per load)
  auto load_rows8 = [&](int row0, int cb, float (&vt)[8]) {
    const int c = cb + lane;
    const bool colok = c >= 1 && c <= Wd - 2;
    const int cc = min(max(c, 1), Wd - 2);            // clamped coordinates: all 16 loads are unconditional and issue back to back
    float a[8], b[8];
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
      const int rowc = min(max(row0 + rr, 1), H - 2);
      a[rr] = ptsf[4 * ((size_t)(rowc + rA) * Wd + cc + offA)];
      b[rr] = ptsf[4 * ((size_t)(rowc + rB) * Wd + cc + offB)];
    }
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
      const int row = row0 + rr;
      const float d = __fsub_rn(a[rr], b[rr]);
      vt[rr] = (colok && row >= 1 && row <= H - 2) ? d : 0.0f;
    }
  };
  // L2 prefetch of the same 16 sectors-per-row two column blocks ahead (a full 32-step period before load_rows8 asks for them): the
  // point image was just streamed out by points_kernel and is far larger than L2, so without it every 8-step piece waits on DRAM
  auto prefetch_rows8 = [&](int row0, int cb) {
    const int cc = min(max(cb + lane, 1), Wd - 2);
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
      const int rowc = min(max(row0 + rr, 1), H - 2);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptsf + 4 * ((size_t)(rowc + rA) * Wd + cc + offA)));
      if (!isX) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptsf + 4 * ((size_t)(rowc + rB) * Wd + cc + offB)));
    }
  };
  auto store_rows8 = [&](int rr0, int cb, const float (&vt)[8]) {
    const int col = (cb + lane) & (kIntRing - 1);
#pragma unroll
    for (int rr = 0; rr < 8; rr++) ring[(rr0 + rr) * kIntRingPitch + col] = vt[rr];
  };
  for (int r0 = 0; r0 < H; r0 += 32) {
    __syncwarp();                                     // the band above is complete and visible (same warp wrote it)
    const int r = r0 + lane;
    const bool rowok = r < H;
    const double *uprev = ii + (size_t)r0 * IWp;      // II row r0 = sums down to pixel row r0 - 1 (zeros for the first band)
    double *orow = ii + (size_t)(r + 1) * IWp;
    const float *myrow = ring + lane * kIntRingPitch;
    double last = 0.0, upleft = 0.0, h0 = 0.0, h1 = 0.0, h2 = 0.0;
    double up = (lane == 0 && r0 > 0) ? __ldcg(uprev + 1) : 0.0;
    double chunk_cur = 0.0, chunk_next = (r0 > 0 && 2 + lane <= Wd) ? __ldcg(uprev + 2 + lane) : 0.0;
#pragma unroll 1
    for (int seg = 0; seg < 4; seg++) { prefetch_rows8(r0 + 8 * seg, 32); float vt[8]; load_rows8(r0 + 8 * seg, 0, vt); store_rows8(8 * seg, 0, vt); }   // column block 0
    __syncwarp();
#pragma unroll 1
    for (int t0 = 0; t0 < steps; t0 += 32) {
      chunk_cur = chunk_next;                         // chunk_cur[i] = II(r0, t0 + 2 + i)
      { const int j = t0 + 32 + 2 + lane; chunk_next = (r0 > 0 && j <= Wd) ? __ldcg(uprev + j) : 0.0; }
#pragma unroll
      for (int seg = 0; seg < 4; seg++) {
        // the next column block arrives in four 8-row pieces: loads in flight during 8 steps, then into the ring slot nobody reads now
        float vt[8];
        load_rows8(r0 + 8 * seg, t0 + 32, vt);
        if (t0 + 64 < Wd) prefetch_rows8(r0 + 8 * seg, t0 + 64);
#pragma unroll
        for (int i8 = 0; i8 < 8; i8++) {
          const int i = seg * 8 + i8;
          const int t = t0 + i;
          const int c = t - 2 * lane;
          // value of the row above for the NEXT step: lane l-1's latest result (column c+1 of its row), or the chunk for lane 0
          double upn = __shfl_up_sync(0xffffffffu, last, 1);
          const double cv = __shfl_sync(0xffffffffu, chunk_cur, i);
          if (lane == 0) upn = cv;
          const float v = myrow[c & (kIntRing - 1)];
          if (rowok && c >= 0 && c < Wd) {
            const double val = __dadd_rn(__dsub_rn(__dadd_rn(up, last), upleft), (double)v);
            last = val;
            if ((c & 3) == 2) st_global_256(orow + (c - 2), h0, h1, h2, val);   // integral-image columns c-2 .. c+1 (h0 = column 0 = zero at c = 2)
            h0 = h1; h1 = h2; h2 = val;
          }
          upleft = up; up = upn;
        }
        store_rows8(8 * seg, t0 + 32, vt);
      }
      __syncwarp();                                   // the new block is visible before the next 32 steps read it
    }
    // columns after the last complete group of four: Wd & ~3 .. Wd (+ padding)
    if (rowok) {
      const int rem = Wd & 3;                         // number of results not yet stored: (Wd + 1) % 4 entries remain, the first may be pending in h*
      const int jb = (Wd + 1) & ~3;                   // first column of the incomplete group
      const int pend = Wd + 1 - jb;                   // 0..3 pending columns: values for columns jb .. Wd
      if (pend == 1) st_global_256(orow + jb, h2, 0.0, 0.0, 0.0);
      else if (pend == 2) st_global_256(orow + jb, h1, h2, 0.0, 0.0);
      else if (pend == 3) st_global_256(orow + jb, h0, h1, h2, 0.0);
      (void)rem;
    }
  }
}

// ---------------------------------------------------------------------------------------------- P3: depth-change + chamfer distance map
__device__ __forceinline__ bool dc_viol(float d0, float d1) {   // depth-change test of computeFeature (max_depth_change_factor 0.05)
  float ddc = __fmul_rn(__fmul_rn(0.05f, __fadd_rn(fabsf(d0), 1.0f)), 2.0f);
  return fabsf(__fsub_rn(d0, d1)) > ddc || !isfinite(d0) || !isfinite(d1);
}
constexpr int kChamferWin = 13;
// blockDim.x = cols rounded up to 32 (<= 1024): thread = column
__global__ void distmap_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  __shared__ float s_prev[1024 + 2];
  __shared__ float s_a[1024 + 2];
  const int H = prm.rows, Wd = prm.cols, NP = H * Wd;
  const int f = blockIdx.x, ci = threadIdx.x;
  const float4 *pts = W.pts + (size_t)f * NP;
  float *dm = W.dm + (size_t)f * NP;
  const float big = (float)(Wd + H);
  auto init_val = [&](int ri, int c) -> float {
    float z = pts[(size_t)ri * Wd + c].z;
    bool zero = false;
    if (ri <= H - 2) {
      if (c <= Wd - 2) { if (dc_viol(z, pts[(size_t)ri * Wd + c + 1].z)) zero = true; if (dc_viol(z, pts[(size_t)(ri + 1) * Wd + c].z)) zero = true; }
      if (c >= 1) { if (dc_viol(pts[(size_t)ri * Wd + c - 1].z, z)) zero = true; }
    }
    if (ri >= 1 && c <= Wd - 2) { if (dc_viol(pts[(size_t)(ri - 1) * Wd + c].z, z)) zero = true; }
    return zero ? 0.0f : big;
  };
  // ---- first pass (top-left to bottom-right)
  float cur = 0.0f;
  if (ci < Wd) { cur = init_val(0, ci); dm[ci] = cur; s_prev[ci] = cur; }
  __syncthreads();
  // The per-row recurrences are short; what would dominate a row step is the latency of the point loads behind init_val.  They only
  // depend on static data, so the RAW depths a row needs (left / right neighbour, the pixel two rows down) are requested one row step
  // before they are used and the column's own depths stream through registers (up <- centre <- down).
  const int cl = max(ci - 1, 0), cr = min(ci + 1, Wd - 1), cc = min(ci, Wd - 1);
  auto zat = [&](int r_, int c_) -> float { return pts[(size_t)min(r_, H - 1) * Wd + c_].z; };
  float zu = 0.0f, zc = zat(0, cc), zd = zat(1, cc);                     // column stream, positioned on row 0
  float zl_n = zat(1, cl), zr_n = zat(1, cr), zd_n = zat(2, cc);          // row 1's neighbours and the pixel below it
  // column 0 of the same row, followed by every thread (same addresses: one broadcast transaction per warp): the last column's
  // "upper right" neighbour is current_row[0], whose value in this pass is just its initial one -- no exchange between threads needed
  float yu = 0.0f, yc = zat(0, 0), yd = zat(1, 0), yr_n = zat(1, min(1, Wd - 1)), yd_n = zat(2, 0);
  auto init_regs = [&](int ri, int c_, float u, float cz, float d, float l, float r) -> float {   // init_val(ri, c_) from registers
    bool zero = false;
    if (ri <= H - 2) {
      if (c_ <= Wd - 2) { if (dc_viol(cz, r)) zero = true; if (dc_viol(cz, d)) zero = true; }
      if (c_ >= 1) { if (dc_viol(l, cz)) zero = true; }
    }
    if (ri >= 1 && c_ <= Wd - 2) { if (dc_viol(u, cz)) zero = true; }
    return zero ? 0.0f : big;
  };
  // two barriers per row: (1) the row's pre-scan values are complete, (2) the row's final values are complete
  for (int ri = 1; ri < H; ri++) {
    zu = zc; zc = zd; zd = zd_n; yu = yc; yc = yd; yd = yd_n;
    const float zl = zl_n, zr = zr_n, yr = yr_n;
    zl_n = zat(ri + 1, cl); zr_n = zat(ri + 1, cr); zd_n = zat(ri + 2, cc);   // for the next row step
    yr_n = zat(ri + 1, min(1, Wd - 1)); yd_n = zat(ri + 2, 0);
    const float center = init_regs(ri, ci, zu, zc, zd, zl, zr);
    float a = center;
    if (ci < Wd) {
      if (ci >= 1) {
        float upLeft = __fadd_rn(s_prev[ci - 1], 1.4f), up = __fadd_rn(s_prev[ci], 1.0f);
        float upRight = (ci + 1 < Wd) ? __fadd_rn(s_prev[ci + 1], 1.4f) : 0.0f;
        a = fminf(a, fminf(upLeft, up));
        if (ci + 1 < Wd) a = fminf(a, upRight);
        // previous_row[ci+1] runs into current_row[0] (never updated in this pass: its initial value)
        if (ci == Wd - 1) a = fminf(a, __fadd_rn(init_regs(ri, 0, yu, yc, yd, 0.0f, yr), 1.4f));
      }
      s_a[ci] = a;
    }
    __syncthreads();
    if (ci < Wd) {
      float v;
      if (ci == 0) v = a;
      else {
        int k0 = max(0, ci - kChamferWin);
        v = s_a[k0];
        for (int k = k0 + 1; k <= ci; k++) v = fminf(s_a[k], __fadd_rn(v, 1.0f));
      }
      cur = v;
      s_prev[ci] = cur; dm[(size_t)ri * Wd + ci] = cur;
    }
    __syncthreads();
  }
  // ---- second pass (bottom-right to top-left); s_prev holds the final row H-1.  The forward values this pass reads (own column, and
  // the row's last column for the ci == 0 wrap: next_row[-1] == current_row[W-1]) are requested one row step ahead; the thread of the
  // last column overwrites its value only after the barrier that follows their use.
  float fwd_next = (H > 1) ? dm[(size_t)(H - 2) * Wd + cc] : 0.0f, wrap_next = (H > 1) ? dm[(size_t)(H - 2) * Wd + Wd - 1] : 0.0f;
  for (int ri = H - 2; ri >= 0; ri--) {
    const float fwd = fwd_next, wrapf = wrap_next;
    if (ri >= 1) { fwd_next = dm[(size_t)(ri - 1) * Wd + cc]; wrap_next = dm[(size_t)(ri - 1) * Wd + Wd - 1]; }
    float a = fwd;
    if (ci < Wd) {
      if (ci <= Wd - 2) {
        float lowerLeft = (ci >= 1) ? __fadd_rn(s_prev[ci - 1], 1.4f) : __fadd_rn(wrapf, 1.4f);
        float lower = __fadd_rn(s_prev[ci], 1.0f), lowerRight = __fadd_rn(s_prev[ci + 1], 1.4f);
        a = fminf(a, fminf(lowerLeft, fminf(lower, lowerRight)));
      }
      s_a[ci] = a;
    }
    __syncthreads();
    if (ci < Wd) {
      float v;
      if (ci == Wd - 1) v = a;
      else {
        int k0 = min(Wd - 1, ci + kChamferWin);
        v = s_a[k0];
        for (int k = k0 - 1; k >= ci; k--) v = fminf(s_a[k], __fadd_rn(v, 1.0f));
      }
      cur = v;
      s_prev[ci] = cur; dm[(size_t)ri * Wd + ci] = cur;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- P4: normals + plane_d
__global__ void normals_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  const int H = prm.rows, Wd = prm.cols, NP = H * Wd, IW = ii_pitch(Wd);
  int f = blockIdx.y;
  if (W.exact_mode && frame_ii_exact(W.crange, f, H, Wd)) return;   // done by normals_exact_kernel
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NP; idx += gridDim.x * blockDim.x) {
  int ri = idx / Wd, ci = idx - ri * Wd;
  const float4 p = W.pts[(size_t)f * NP + idx];
  const float nanf_ = __int_as_float(0x7fc00000);
  float nx = nanf_, ny = nanf_, nz = nanf_;
  const int border = 10;   // int(normal_smoothing_size_)
  if (ri >= border && ri < H - border && ci >= border && ci < Wd - border && isfinite(p.z)) {
    float smoothing = fminf(W.dm[(size_t)f * NP + idx], 10.0f);
    if (smoothing > 2.0f) {
      int rw = (int)smoothing, rw2 = rw / 2;
      const size_t plane = (size_t)(H + 1) * IW;   // planar integral images, padded pitch (integral_kernel)
      const double *iiX = W.iiX + (size_t)f * 3 * plane;
      const double *iiY = W.iiY + (size_t)f * 3 * plane;
      int sx = ci - rw2, sy = ri - rw2;
      size_t ul = (size_t)sy * IW + sx, ur = ul + rw, ll = (size_t)(sy + rw) * IW + sx, lr = ll + rw;
      double gx[3], gy[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double *px = iiX + k * plane, *py = iiY + k * plane;
        gx[k] = __dsub_rn(__dsub_rn(__dadd_rn(px[lr], px[ul]), px[ur]), px[ll]);
        gy[k] = __dsub_rn(__dsub_rn(__dadd_rn(py[lr], py[ul]), py[ur]), py[ll]);
      }
      double n0 = __dsub_rn(__dmul_rn(gy[1], gx[2]), __dmul_rn(gy[2], gx[1]));   // gradient_y.cross(gradient_x)
      double n1 = __dsub_rn(__dmul_rn(gy[2], gx[0]), __dmul_rn(gy[0], gx[2]));
      double n2 = __dsub_rn(__dmul_rn(gy[0], gx[1]), __dmul_rn(gy[1], gx[0]));
      double nl = __dadd_rn(__dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1)), __dmul_rn(n2, n2));
      if (nl != 0.0) {
        double s = sqrt(nl);
        nx = (float)(n0 / s); ny = (float)(n1 / s); nz = (float)(n2 / s);
        float vx = __fsub_rn(0.0f, p.x), vy = __fsub_rn(0.0f, p.y), vz = __fsub_rn(0.0f, p.z);   // flipNormalTowardsViewpoint, vp = 0
        float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
        if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
      }
    }
  }
  float pd = __fadd_rn(__fadd_rn(__fmul_rn(p.x, nx), __fmul_rn(p.y, ny)), __fmul_rn(p.z, nz));   // plane_d = p . n
  W.nrm[(size_t)f * NP + idx] = make_float4(nx, ny, nz, pd);
  }
}

// ---------------------------------------------------------------------------------------------- P5: connected components
// relaxed GPU-scope accesses (the CTAs of a frame share its parent array through L2); `volatile` would compile to system-scope strong
// accesses (LDG.E.STRONG.SYS) that stall the dispatcher -- 80-95 % of the stall samples of the two union-find kernels
__device__ __forceinline__ int pvload(const int *p) { int v; asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void pvstore(int *p, int v) { asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ int p_find(int *parent, int i) {   // path halving (benign races: only ancestors are ever written)
  int p = pvload(parent + i);
  while (p != i) {
    int gp = pvload(parent + p);
    if (gp != p) pvstore(parent + i, gp);
    i = p; p = gp;
  }
  return i;
}
__device__ void p_union(int *parent, int a, int b) {
  while (true) {
    a = p_find(parent, a); b = p_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    if (atomicCAS(parent + a, a, b) == a) return;
  }
}
__device__ __forceinline__ bool plane_compare(const float4 *pts, const float4 *nrm, int i1, int i2, float dis_th, float ang_cos) {
  float z = pts[i1].z;                          // depth dependent threshold (setDistanceThreshold(.., true)), z_axis = (0,0,1)
  float threshold = __fmul_rn(dis_th, __fmul_rn(z, z));
  float4 a = nrm[i1], b = nrm[i2];
  float dotn = __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
  return (fabsf(__fsub_rn(a.w, b.w)) < threshold) && (dotn > ang_cos);
}
__device__ __forceinline__ bool plane_compare_vals(float z, float4 a, float4 b, float dis_th, float ang_cos) {   // plane_compare on loaded values
  const float threshold = __fmul_rn(dis_th, __fmul_rn(z, z));
  const float dotn = __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
  return (fabsf(__fsub_rn(a.w, b.w)) < threshold) && (dotn > ang_cos);
}
// shared-memory union-find of a strip (volatile on shared memory is a plain LDS/STS)
__device__ __forceinline__ int s_find(int *par, int i) {
  int p = *((volatile int *)(par + i));
  while (p != i) {
    const int gp = *((volatile int *)(par + p));
    if (gp != p) *((volatile int *)(par + i)) = gp;
    i = p; p = gp;
  }
  return i;
}
__device__ __forceinline__ void s_union(int *par, int a, int b) {
  while (true) {
    a = s_find(par, a); b = s_find(par, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    if (atomicCAS(par + a, a, b) == a) return;
  }
}
// Phase 1: one CTA labels a strip of kStripRows full image rows entirely in shared memory.  Thread = column; per row the horizontal
// runs inside a warp get their first pixel as parent (plain stores), the links that are left -- runs continuing across a warp seam and
// the non-redundant vertical links (a vertical link is redundant when the pixel and its left neighbour are in the same run, both link
// upwards and the two upper pixels are linked to each other) -- go through the shared-memory union-find.  Every pixel then gets the
// global index of its strip-local root (the smallest index of its component inside the strip) as parent.
constexpr int kStripRows = 16;
__global__ void ccl_strip_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  extern __shared__ int s_par[];   // [kStripRows * cols]
  const int H = prm.rows, Wd = prm.cols, NP = H * Wd;
  const int f = blockIdx.y, row0 = blockIdx.x * kStripRows;
  const int nrows = min(kStripRows, H - row0);
  const int c = threadIdx.x, lane = c & 31;
  const bool colok = c < Wd;
  const float4 *pts = W.pts + (size_t)f * NP + (size_t)row0 * Wd, *nrm = W.nrm + (size_t)f * NP + (size_t)row0 * Wd;
  unsigned hbits = 0, vbits = 0;   // bit r: run continues across the warp seam / non-redundant vertical link of local row r
  float4 n_up = make_float4(0, 0, 0, 0);
  bool h_up = false;
  for (int r = 0; r < nrows; r++) {
    const int li = r * Wd + c;
    float z = 0.0f; float4 n = make_float4(0, 0, 0, 0);
    if (colok) { z = pts[li].z; n = nrm[li]; }
    float4 nl;
    nl.x = __shfl_up_sync(0xffffffffu, n.x, 1); nl.y = __shfl_up_sync(0xffffffffu, n.y, 1);
    nl.z = __shfl_up_sync(0xffffffffu, n.z, 1); nl.w = __shfl_up_sync(0xffffffffu, n.w, 1);
    if (lane == 0 && colok && c >= 1) nl = nrm[li - 1];
    const bool h = colok && c >= 1 && plane_compare_vals(z, n, nl, prm.dis_th, prm.ang_cos);
    const bool v = colok && r >= 1 && plane_compare_vals(z, n, n_up, prm.dis_th, prm.ang_cos);
    const unsigned starts = __ballot_sync(0xffffffffu, !h || lane == 0);
    if (colok) {
      const int head_lane = 31 - __clz(starts & (0xffffffffu >> (31 - lane)));
      s_par[li] = li - (lane - head_lane);
    }
    const bool v_left = __shfl_up_sync(0xffffffffu, v ? 1 : 0, 1) != 0;
    const bool need_v = v && !(h && lane >= 1 && v_left && h_up);
    if (h && lane == 0) hbits |= 1u << r;
    if (need_v) vbits |= 1u << r;
    n_up = n; h_up = h;
  }
  __syncthreads();
  for (int r = 0; r < nrows; r++) {
    if ((hbits >> r) & 1u) s_union(s_par, r * Wd + c, r * Wd + c - 1);
    if ((vbits >> r) & 1u) s_union(s_par, r * Wd + c, (r - 1) * Wd + c);
  }
  __syncthreads();
  // Parents out; region statistics (size, last pixel, column range) go to the strip-local root's slots: each thread walks its column
  // and issues one set of reductions per vertical run, so the counters of a big region see a few hundred updates per strip instead
  // of one per warp of the whole frame on four global addresses; ccl_flatten_kernel folds the local roots into the global ones.
  if (colok) {
    const size_t gbase = (size_t)f * NP + (size_t)row0 * Wd;
    int cur = -1, cnt = 0, lastidx = 0;
    for (int r = 0; r <= nrows; r++) {
      const int root = r < nrows ? s_find(s_par, r * Wd + c) : -2;
      if (root != cur) {
        if (cnt > 0) {
          atomicAdd(W.count + gbase + cur, cnt); atomicMax(W.maxidx + gbase + cur, lastidx);
          atomicMin(W.cmin + gbase + cur, c); atomicMax(W.cmax + gbase + cur, c);
        }
        cur = root; cnt = 0;
      }
      if (r < nrows) { W.parent[gbase + r * Wd + c] = row0 * Wd + root; cnt++; lastidx = (row0 + r) * Wd + c; }
    }
  }
}
// Phase 2: the vertical links across strip boundaries (first row of every strip but the first) on the global parents.
__global__ void ccl_seam_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  const int Wd = prm.cols, NP = prm.rows * prm.cols;
  const int c = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const int f = blockIdx.z, row = (blockIdx.y + 1) * kStripRows;
  const bool act = c < Wd && row < prm.rows;
  const float4 *pts = W.pts + (size_t)f * NP, *nrm = W.nrm + (size_t)f * NP;
  const int idx = row * Wd + c;
  const bool v = act && plane_compare(pts, nrm, idx, idx - Wd, prm.dis_th, prm.ang_cos);
  const bool h = act && c >= 1 && plane_compare(pts, nrm, idx, idx - 1, prm.dis_th, prm.ang_cos);
  const bool v_left = __shfl_up_sync(0xffffffffu, v ? 1 : 0, 1) != 0;
  bool need_v = v;
  if (v && h && lane >= 1 && v_left && plane_compare(pts, nrm, idx - Wd, idx - Wd - 1, prm.dis_th, prm.ang_cos)) need_v = false;
  if (need_v) p_union(W.parent + (size_t)f * NP, idx, idx - Wd);
}
// root of pixel i once all unions are done (the parents are static now): weak L2 loads, no path compression -- the chains are a strip
// root plus a few seam hops, and strong (ld.relaxed.gpu) loads stall the dispatcher (80 % of this kernel's stall samples before)
__device__ __forceinline__ int p_find_static(const int *parent, int i) {
  int p = __ldcg(parent + i);
  while (p != i) { i = p; p = __ldcg(parent + i); }
  return i;
}
__global__ void ccl_flatten_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  const int NP = prm.rows * prm.cols;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int f = blockIdx.y;
  const bool act = idx < NP;
  const int *parent = W.parent + (size_t)f * NP;
  if (!act) return;
  const int r = p_find_static(parent, idx);
  W.label[(size_t)f * NP + idx] = r;
  // strip-local roots (the only pixels with a non-zero size) that were merged into another root hand their statistics on
  const int cn = W.count[(size_t)f * NP + idx];
  if (cn > 0 && r != idx) {
    atomicAdd(W.count + (size_t)f * NP + r, cn);
    atomicMax(W.maxidx + (size_t)f * NP + r, W.maxidx[(size_t)f * NP + idx]);
    atomicMin(W.cmin + (size_t)f * NP + r, W.cmin[(size_t)f * NP + idx]);
    atomicMax(W.cmax + (size_t)f * NP + r, W.cmax[(size_t)f * NP + idx]);
  }
}

// ---------------------------------------------------------------------------------------------- P6: candidate regions (> 100 px), label order
__global__ void candidates_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  __shared__ BlockScratch S;
  const int NP = prm.rows * prm.cols;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int *parent = W.label + (size_t)f * NP, *count = W.count + (size_t)f * NP;
  int K = 0;
  for (int tile = 0; tile < NP; tile += 4 * blockDim.x) {
    bool acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {   // four independent label loads in flight
      const int i = tile + u * blockDim.x + tid;
      acc[u] = (i < NP) && parent[i] == i;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = tile + u * blockDim.x + tid; acc[u] = acc[u] && (unsigned)count[i] > 100u; }   // min_inliers_ = 100, strict
    // almost every tile holds no region root: one barrier per four tiles instead of a scan per tile
    if (!__syncthreads_or((acc[0] || acc[1] || acc[2] || acc[3]) ? 1 : 0)) continue;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = tile + u * blockDim.x + tid;
      int tot;
      int ex = block_excl_scan(acc[u] ? 1 : 0, S, tot);
      if (acc[u] && K + ex < kMaxRegions) { W.cand_root[f * kMaxRegions + K + ex] = i; W.cand_size[f * kMaxRegions + K + ex] = count[i]; }
      K += tot;
    }
  }
  if (tid == 0) W.cand_n[f] = min(K, kMaxRegions);
}

// ---------------------------------------------------------------------------------------------- P7: plane fit (float, sequential raster order)
__device__ void roots2f(float b, float c, float *roots) {
  roots[0] = 0.0f;
  float d = (float)((double)__fmul_rn(b, b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = sqrtf(d);
  roots[2] = __fmul_rn(0.5f, __fadd_rn(b, sd));
  roots[1] = __fmul_rn(0.5f, __fsub_rn(b, sd));
}
__device__ void roots3f(const float *m /*3x3*/, float *roots) {   // pcl::computeRoots
  float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
  float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  float c2 = m[0] + m[4] + m[8];
  if (fabsf(c0) < 1.1920929e-07f) roots2f(c2, c1, roots);
  else {
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = sqrtf(-a_over_3);
    float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
    float cos_theta = cosf(theta), sin_theta = sinf(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    float t;
    if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[1] >= roots[2]) { t = roots[1]; roots[1] = roots[2]; roots[2] = t; if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; } }
    if (roots[0] <= 0) roots2f(c2, c1, roots);
  }
}
// The float moment sums must be formed in raster order (bit parity with pcl::computeMeanAndCovarianceMatrix<float>), so one dependent
// FADD chain per region (~4 cycles per member point) is the floor; everything else is taken off it with a producer/consumer split
// inside a 256-thread CTA that works through the frame's candidate regions largest first (CTA b takes size ranks b, b+8, ...):
//   warps 1..7 (producers) each own one shared-memory tile slot: producer p takes tiles p-1, p-1+7, ... of the region's bounding box
//                          (8 groups of 32 pixels in raster order), keeps 16 coalesced loads in flight, ranks the members with ballots
//                          and writes their 9 addends, order preserved, into its slot -- seven tiles in flight cover the DRAM latency;
//   warp 0 (consumer)      walks the tiles in order; lanes 0..8 own one accumulator each (xx xy xz yy yz zz x y z): per point one LDS
//                          and the FADD.
// Slots are handed over with sequence counters in shared memory (produced / consumed per slot, cumulative over regions).
constexpr int kFitWarps = 8;
constexpr int kFitProducers = kFitWarps - 1;
constexpr int kFitThreads = 32 * kFitWarps;
constexpr int kFitGroups = 8;                                   // 32-pixel groups per tile
constexpr int kFitTilePts = kFitGroups * 32;                    // 256 points per tile
constexpr int kFitRankMax = 256;
constexpr int kFitSmemBytes = kFitProducers * kFitTilePts * 9 * (int)sizeof(float);
__device__ __forceinline__ void fit_wait_eq(volatile int *flag, int want) {   // warp-collective: lane 0 polls, the warp follows
  if ((threadIdx.x & 31) == 0) while (*flag != want) __nanosleep(20);
  __syncwarp();
  __threadfence_block();
}
__global__ void __launch_bounds__(kFitThreads) fit_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  extern __shared__ float s_v[];               // [kFitProducers][kFitTilePts * 9]: the 9 addends of every member point of a tile
  __shared__ volatile int s_full[kFitProducers], s_done[kFitProducers];
  __shared__ int s_cnt[kFitProducers];
  __shared__ int s_size[kFitRankMax], s_order[kFitRankMax];
  const int NP = prm.rows * prm.cols, Wd = prm.cols;
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int *parent = W.label + (size_t)f * NP;
  const float4 *pts = W.pts + (size_t)f * NP;
  const int ncand = W.cand_n[f];
  const bool ranked = ncand <= kFitRankMax;
  if (threadIdx.x < kFitProducers) { s_full[threadIdx.x] = 0; s_done[threadIdx.x] = 0; }
  if (ranked) {   // size rank (descending, ties by index): the big regions start first, on different CTAs
    for (int k = threadIdx.x; k < ncand; k += kFitThreads) s_size[k] = W.cand_size[f * kMaxRegions + k];
    __syncthreads();
    for (int k = threadIdx.x; k < ncand; k += kFitThreads) {
      const int sk = s_size[k];
      int rank = 0;
      for (int j = 0; j < ncand; j++) { const int sj = s_size[j]; rank += (sj > sk || (sj == sk && j < k)) ? 1 : 0; }
      s_order[rank] = k;
    }
  }
  __syncthreads();
  int prev = 0;                                // producer: code of the tile this warp's slot carried last
  int it = 0;                                  // region counter: slot hand-over codes are (it << 20) | (tile + 1), unique per CTA
  for (int rk = blockIdx.x; rk < ncand; rk += gridDim.x) {
  ++it;
  const int k = ranked ? s_order[rk] : rk;
  const int root = W.cand_root[f * kMaxRegions + k];
  const int last = W.maxidx[(size_t)f * NP + root];
  const int r0 = root / Wd, r1 = last / Wd;                       // bounding box rows / columns (32-aligned)
  const int c0 = W.cmin[(size_t)f * NP + root] & ~31, c1 = W.cmax[(size_t)f * NP + root];
  const int gpr = (c1 - c0) / 32 + 1;                             // groups per row
  const int ngroups = (r1 - r0 + 1) * gpr;
  const int ntiles = (ngroups + kFitGroups - 1) / kFitGroups;
  float acc = 0.0f;
  if (warp >= 1) {
    const int p = warp - 1;
    float *slot = s_v + (size_t)p * kFitTilePts * 9;
    for (int t = p; t < ntiles; t += kFitProducers) {
      float4 mine[kFitGroups]; int labs[kFitGroups];
      int gr = (t * kFitGroups) / gpr, gc = (t * kFitGroups) - gr * gpr;   // row / group-in-row of the tile's first group
#pragma unroll
      for (int q = 0; q < kFitGroups; q++) {        // 16 independent coalesced loads
        const int g = t * kFitGroups + q;
        const int cc = c0 + gc * 32 + lane;
        labs[q] = -1; mine[q] = make_float4(0, 0, 0, 0);
        if (g < ngroups && cc < Wd) { const int i = (r0 + gr) * Wd + cc; labs[q] = parent[i]; mine[q] = pts[i]; }
        if (++gc == gpr) { gc = 0; gr++; }
      }
      fit_wait_eq(&s_done[p], prev);                // the consumer is done with the tile this slot carried before
      int off = 0;
#pragma unroll
      for (int q = 0; q < kFitGroups; q++) {
        const unsigned mask = __ballot_sync(0xffffffffu, labs[q] == root);
        if ((mask >> lane) & 1u) {
          float *o = slot + (off + __popc(mask & ((1u << lane) - 1u))) * 9;
          const float4 pt = mine[q];
          o[0] = __fmul_rn(pt.x, pt.x); o[1] = __fmul_rn(pt.x, pt.y); o[2] = __fmul_rn(pt.x, pt.z);
          o[3] = __fmul_rn(pt.y, pt.y); o[4] = __fmul_rn(pt.y, pt.z); o[5] = __fmul_rn(pt.z, pt.z);
          o[6] = pt.x; o[7] = pt.y; o[8] = pt.z;
        }
        off += __popc(mask);
      }
      __syncwarp();
      prev = (it << 20) | (t + 1);
      if (lane == 0) { s_cnt[p] = off; __threadfence_block(); s_full[p] = prev; }
    }
  } else {
    for (int t = 0; t < ntiles; t++) {
      const int p = t % kFitProducers;
      const int want = (it << 20) | (t + 1);
      fit_wait_eq(&s_full[p], want);
      const int n = s_cnt[p];
      const float *buf = s_v + (size_t)p * kFitTilePts * 9 + (lane < 9 ? lane : 0);
      int j = 0;
      for (; j + 8 <= n; j += 8) {   // 8 independent LDS, then the 8-long dependent FADD chain
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = buf[(j + u) * 9];
#pragma unroll
        for (int u = 0; u < 8; u++) acc = __fadd_rn(acc, v[u]);
      }
      for (; j < n; j++) acc = __fadd_rn(acc, buf[j * 9]);
      __syncwarp();
      if (lane == 0) { __threadfence_block(); s_done[p] = want; }
    }
  }
  if (warp == 0) {
  float accu[9];
#pragma unroll
  for (int j = 0; j < 9; j++) accu[j] = __shfl_sync(0xffffffffu, acc, j);
  if (lane == 0) {
  float cnt = (float)W.count[(size_t)f * NP + root];
#pragma unroll
  for (int j = 0; j < 9; j++) accu[j] = __fdiv_rn(accu[j], cnt);
  float cov[9];
  cov[0] = __fsub_rn(accu[0], __fmul_rn(accu[6], accu[6])); cov[1] = __fsub_rn(accu[1], __fmul_rn(accu[6], accu[7])); cov[2] = __fsub_rn(accu[2], __fmul_rn(accu[6], accu[8]));
  cov[4] = __fsub_rn(accu[3], __fmul_rn(accu[7], accu[7])); cov[5] = __fsub_rn(accu[4], __fmul_rn(accu[7], accu[8])); cov[8] = __fsub_rn(accu[5], __fmul_rn(accu[8], accu[8]));
  cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
  // pcl::eigen33(mat, eigenvalue, eigenvector)
  float scale = 0.0f;
  for (int j = 0; j < 9; j++) scale = fmaxf(scale, fabsf(cov[j]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;
  float sm[9];
  for (int j = 0; j < 9; j++) sm[j] = __fdiv_rn(cov[j], scale);
  float ev[3];
  roots3f(sm, ev);
  float eigen_value = ev[0] * scale;
  sm[0] -= ev[0]; sm[4] -= ev[0]; sm[8] -= ev[0];
  float v1[3] = {sm[1] * sm[5] - sm[2] * sm[4], sm[2] * sm[3] - sm[0] * sm[5], sm[0] * sm[4] - sm[1] * sm[3]};      // row0 x row1
  float v2[3] = {sm[1] * sm[8] - sm[2] * sm[7], sm[2] * sm[6] - sm[0] * sm[8], sm[0] * sm[7] - sm[1] * sm[6]};      // row0 x row2
  float v3[3] = {sm[4] * sm[8] - sm[5] * sm[7], sm[5] * sm[6] - sm[3] * sm[8], sm[3] * sm[7] - sm[4] * sm[6]};      // row1 x row2
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2], l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float *v; float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; } else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; } else { v = v3; l = l3; }
  float s = sqrtf(l);
  float pp[4] = {__fdiv_rn(v[0], s), __fdiv_rn(v[1], s), __fdiv_rn(v[2], s), 0.0f};
  pp[3] = -1 * (pp[0] * accu[6] + pp[1] * accu[7] + pp[2] * accu[8] + pp[3] * 1.0f);
  // (PCL flips the sign towards an accumulated view point here; the reference re-normalises the sign afterwards -- PlaneExtractor.cpp:98-99,176-178)
  float eig_sum = cov[0] + cov[4] + cov[8];
  float curvature = (eig_sum != 0) ? fabsf(eigen_value / eig_sum) : 0.0f;
  W.cand_coef[f * kMaxRegions + k] = make_float4(pp[0], pp[1], pp[2], pp[3]);
  int valid = curvature < 0.001f;   // maximum_curvature_
  W.cand_valid[f * kMaxRegions + k] = valid;
  if (valid) W.count[(size_t)f * NP + root] = -(k + 1);   // marks the root as a model label (refine_labels_)
  }
  }
  }
}

// ---------------------------------------------------------------------------------------------- P8: refinement
struct MaxScanScratch { int tot[2][32]; };
// inclusive max scan in thread order (or reverse thread order): warp scan by shuffles, the warp totals cross through a double-buffered
// shared table (one barrier per call; `phase` flips in every thread), each warp folds the totals of the warps before (after) it with
// one load per lane and a shuffle reduction
__device__ __forceinline__ int block_incl_max_scan(int v, MaxScanScratch &S, int &phase, bool reverse) {
  const int lane = lane_id(), w = warp_id(), nw = n_warps();
  const int ninf = -0x7fffffff;
  int x = v;
  if (!reverse) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x = max(x, t); }
    if (lane == 31) S.tot[phase][w] = x;
  } else {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_down_sync(0xffffffffu, x, o); if (lane + o < 32) x = max(x, t); }
    if (lane == 0) S.tot[phase][w] = x;
  }
  __syncthreads();
  int y = ninf;
  if (lane < nw && (reverse ? (lane > w) : (lane < w))) y = S.tot[phase][lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) y = max(y, __shfl_xor_sync(0xffffffffu, y, o));
  phase ^= 1;
  return max(x, y);
}
__device__ __forceinline__ bool refine_ok(const float4 c, const float4 p) {   // PlaneRefinementComparator: |a x + b y + c z + d| < 0.02
  float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.x, p.x), __fmul_rn(c.y, p.y)), __fmul_rn(c.z, p.z)), c.w);
  return fabsf(d) < 0.02f;
}
// Per-pixel model image for the refinement: index of the valid plane model the pixel's region carries, or -1 (what the chained lookup
// parent -> count gave before).  Stored in the distance-map buffer, which is free after the connected components.
__global__ void model_image_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  const int NP = prm.rows * prm.cols;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (idx >= NP) return;
  const int r = W.label[(size_t)f * NP + idx];
  const int cn = W.count[(size_t)f * NP + r];
  reinterpret_cast<int *>(W.dm)[(size_t)f * NP + idx] = cn < 0 ? (-cn - 1) : -1;
}
// blockDim.x = cols rounded up to 32 (<= 1024); thread = column.  The scan-order label propagation of
// OrganizedMultiPlaneSegmentation::refine is row-sequential (958 dependent row steps per frame), so nothing global may sit on the
// per-step critical path: every thread carries its column's model label of the current and the next row in registers (the vertical
// check only ever relabels the thread's own column), the labels and points of the row after are prefetched one step ahead, and the
// plane coefficients of the first 64 models sit in shared memory.
constexpr int kRefineCoefSmem = 64;
__global__ void refine_kernel(PlaneParamsDev prm, PlaneFramePtrs W) {
  __shared__ MaxScanScratch S;
  int phase = 0;
  __shared__ int s_m[2 * 1024];
  __shared__ float4 s_coef[kRefineCoefSmem];
  __shared__ volatile int s_post_m, s_post_seq;   // label of column 0 after its row step, handed to the thread of column W-1 (backward pass)
  __shared__ int s_add[kRefineCoefSmem];   // inliers gained by the first 64 models (flushed once)
  const int H = prm.rows, Wd = prm.cols, NP = H * Wd;
  const int f = blockIdx.x, c = threadIdx.x, lane = c & 31, w = c >> 5, nw = (blockDim.x + 31) >> 5;
  const bool colok = c < Wd;
  int *parent = W.label + (size_t)f * NP;
  int *mimg = reinterpret_cast<int *>(W.dm) + (size_t)f * NP;
  const float4 *pts = W.pts + (size_t)f * NP;
  const float4 *coef_g = W.cand_coef + f * kMaxRegions;
  const int *croot = W.cand_root + f * kMaxRegions;
  int *csize = W.cand_size + f * kMaxRegions;
  const int ncand = W.cand_n[f];
  for (int i = threadIdx.x; i < kRefineCoefSmem; i += blockDim.x) { s_add[i] = 0; if (i < ncand) s_coef[i] = coef_g[i]; }
  if (threadIdx.x == 0) { s_post_m = -1; s_post_seq = 0; }
  __syncthreads();
  auto coef = [&](int L) -> float4 { return L < kRefineCoefSmem ? s_coef[L] : coef_g[L]; };
  // inlier counters: every thread keeps (model, count) of its current streak in registers and flushes on a change of model
  int acc_m = -1, acc_n = 0;
  auto flush_inliers = [&]() {
    if (acc_n > 0) { if (acc_m < kRefineCoefSmem) atomicAdd(&s_add[acc_m], acc_n); else atomicAdd(&csize[acc_m], acc_n); }
    acc_n = 0;
  };
  auto add_inlier = [&](int L) { if (L != acc_m) { flush_inliers(); acc_m = L; } acc_n++; };
  // nearest flagged column at or before c in scan direction (-1: none).  The flags of a warp are one ballot, the nearest flagged lane
  // a count-leading-zeros; the warps' own nearest flags cross through shared memory (one barrier, double-buffered) and are folded
  // with one redux instruction.  Replaces two 10-step shuffle scans per call.
  auto nearest_flag = [&](bool flag, bool fwd) -> int {
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int wbase = c - lane;
    int x, tot;
    if (fwd) {
      const unsigned lo = bal & (0xffffffffu >> (31 - lane));
      x = lo ? wbase + 31 - __clz(lo) : -1; tot = bal ? wbase + 31 - __clz(bal) : -1;
    } else {
      const unsigned hi = bal & (0xffffffffu << lane);
      x = hi ? wbase + __ffs(hi) - 1 : -1; tot = bal ? wbase + __ffs(bal) - 1 : -1;
    }
    if (lane == 0) S.tot[phase][w] = tot;
    __syncthreads();
    int r;
    if (fwd) {
      const int y = __reduce_max_sync(0xffffffffu, (lane < w) ? S.tot[phase][lane] : -1);
      r = max(x, y);
    } else {
      int t = 0x7fffffff;
      if (lane > w && lane < nw) { const int v = S.tot[phase][lane]; if (v >= 0) t = v; }
      const int y = __reduce_min_sync(0xffffffffu, t);
      r = x >= 0 ? x : (y == 0x7fffffff ? -1 : y);
    }
    phase ^= 1;
    return r;
  };
  const float4 zero4 = make_float4(0, 0, 0, 0);
  for (int pass = 0; pass < 2; pass++) {
    const bool fwd = pass == 0;
    const int dr = fwd ? 1 : -1;
    const int rfirst = fwd ? 0 : H - 1;
    __syncthreads();   // the first pass's relabelling is visible to the whole CTA
    int m_cur = colok ? mimg[rfirst * Wd + c] : -1;
    float4 p_cur = colok ? pts[rfirst * Wd + c] : zero4;
    int m_nxt = colok ? mimg[(rfirst + dr) * Wd + c] : -1;
    float4 p_nxt = colok ? pts[(rfirst + dr) * Wd + c] : zero4;
    ptrdiff_t off2 = (ptrdiff_t)(rfirst + 2 * dr) * Wd + c;      // pixel two rows ahead in this column
    const ptrdiff_t doff = (ptrdiff_t)dr * Wd;
    for (int step = 0; step < H - 1; step++) {
      const int r = rfirst + dr * step;                   // current row
      const int rn = r + dr;                              // row hit by the vertical check
      int m_n2 = -1; float4 p_n2 = zero4;                 // labels / points of the row after rn: prefetched, nobody writes them this step
      if (colok && step + 1 < H - 1) { m_n2 = mimg[off2]; p_n2 = pts[off2]; }
      off2 += doff;
      int m = m_cur;
      // one barrier publishes the row's labels and tells whether the row holds a model pixel at all (a row without one can neither
      // propagate sideways nor vertically); s_m alternates between two rows so the next step's stores cannot overtake this step's reads
      int *s_mr = s_m + (step & 1) * 1024;
      if (colok) s_mr[c] = m;
      if (__syncthreads_or((colok && m >= 0) ? 1 : 0)) {
        // ---- horizontal propagation along the row (scan direction = traversal direction); skipped when no non-model pixel has a
        //      model pixel as its upstream neighbour (then nothing can be relabelled in this row)
        bool cand = false;
        if (colok && m < 0) cand = fwd ? (c >= 1 && s_mr[c - 1] >= 0) : (c + 1 < Wd && s_mr[c + 1] >= 0);
        if (__syncthreads_or(cand ? 1 : 0)) {
          const int s = nearest_flag(colok && m >= 0, fwd);                       // nearest source at or before c
          int L = -1;
          bool blocked = false;
          if (colok && m < 0) {
            if (s >= 0) { L = s_mr[s]; blocked = !refine_ok(coef(L), p_cur); }
            else blocked = true;
          }
          const int b = nearest_flag(colok && m < 0 && blocked, fwd);             // nearest blocked pixel at or before c
          // a pixel receives its source's label when no blocked pixel lies between the source and itself (the forward pass never
          // uses column W-1 as "current"; every receiver's chain starts at a source that IS a valid current)
          const bool recv = colok && m < 0 && s >= 0 && (b < 0 || (fwd ? (b < s) : (b > s)));
          if (recv) { m = L; parent[r * Wd + c] = croot[L]; mimg[r * Wd + c] = L; add_inlier(L); }
        }
        // ---- vertical check from the (updated) current row: only this thread ever relabels pixel (rn, c)
        const bool cur_ok = colok && m >= 0 && (fwd ? (c <= Wd - 2) : true);
        if (cur_ok && m_nxt < 0 && refine_ok(coef(m), p_nxt)) { m_nxt = m; parent[rn * Wd + c] = croot[m]; mimg[rn * Wd + c] = m; add_inlier(m); }
        if (!fwd) {
          // colIdx == 0: the "left" neighbour is the last pixel of the previous row (pointer arithmetic in PCL) = (rn, W-1), the pixel
          // the thread of column W-1 carries as its next label: column 0 posts its label, column W-1 applies the check itself
          const int seq = step + 1;
          if (c == 0) { s_post_m = m; __threadfence_block(); s_post_seq = seq; }
          if (c == Wd - 1) {
            while (s_post_seq != seq) { }
            __threadfence_block();
            const int m0 = s_post_m;
            if (m0 >= 0 && m_nxt < 0 && refine_ok(coef(m0), p_nxt)) { m_nxt = m0; parent[rn * Wd + c] = croot[m0]; mimg[rn * Wd + c] = m0; add_inlier(m0); }
          }
        }
      }
      m_cur = m_nxt; p_cur = p_nxt; m_nxt = m_n2; p_nxt = p_n2;
    }
  }
  flush_inliers();
  __syncthreads();
  for (int i = threadIdx.x; i < kRefineCoefSmem && i < ncand; i += blockDim.x) if (s_add[i]) atomicAdd(&csize[i], s_add[i]);
}

// ---------------------------------------------------------------------------------------------- P9: selection (PlaneExtractor.cpp:89-106,126-183)
__global__ void select_kernel(PlaneParamsDev prm, PlaneFramePtrs W, int n_frames) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  int K = W.cand_n[f];
  int best = -1, best_size = -1; double bv[4] = {0, 0, 0, 0};
  for (int k = 0; k < K; k++) {
    if (!W.cand_valid[f * kMaxRegions + k]) continue;
    int size = W.cand_size[f * kMaxRegions + k];
    if (size < prm.min_size) continue;
    float4 cf = W.cand_coef[f * kMaxRegions + k];
    if (cf.w < 0) { cf.x = -cf.x; cf.y = -cf.y; cf.z = -cf.z; cf.w = -cf.w; }
    double v[4] = {(double)cf.x, (double)cf.y, (double)cf.z, (double)cf.w};
    double cos_theta = v[0] * 0.0 + v[1] * 1.0 + v[2] * 0.0;
    cos_theta = cos_theta / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / 1.0;
    double theta = acos(cos_theta);
    if (theta > M_PI / 4 && theta < 3 * M_PI / 4) continue;
    if (size > best_size) { best_size = size; best = k; for (int j = 0; j < 4; j++) bv[j] = v[j]; }
  }
  if (best < 0) { W.out_found[f] = 0; W.out_inl[f] = 0; for (int j = 0; j < 4; j++) W.out_plane[4 * f + j] = 0; return; }
  double value = (bv[0] * 0.0 + bv[1] * 0.0 + bv[2] * 0.0 + bv[3]) / sqrt(bv[0] * bv[0] + bv[1] * bv[1] + bv[2] * bv[2]);
  if (value < 0) for (int j = 0; j < 4; j++) bv[j] = -bv[j];
  for (int j = 0; j < 4; j++) W.out_plane[4 * f + j] = bv[j];
  W.out_found[f] = 1; W.out_inl[f] = best_size;
}

// ---------------------------------------------------------------------------------------------- host driver
// debug_stage: 0 = normal run; 1 = stop after the distance map (out: float[NP]); 2 = after normals (float4[NP]);
// 3 = after connected components (int[NP] labels = min pixel index); 4 = after refinement (int[NP]).  First frame only.
int plane_extract_dbg(EslamHandle *h, cudaStream_t s, int n_frames, const uint16_t *depth_dev, size_t pitch, double *plane_cam_host, int *found_host,
                      int *n_inliers_host, int debug_stage, void *debug_out) {
  const int H = h->cam.rows, Wd = h->cam.cols;
  if (H > 1024 || Wd > 1024 || H < 24 || Wd < 24 || pitch < (size_t)Wd) { h->err = "plane_extract: image must be between 24 and 1024 pixels per side"; return ESLAM_ERR_BAD_ARG; }
  if (!h->pl) h->pl = new PlaneWorkspace();
  PlaneWorkspace &w = *h->pl;
  const size_t NP = (size_t)H * Wd, NI = (size_t)(H + 1) * ii_pitch(Wd) * 3;
  // the per-frame kernels (integral / distmap / refine: one CTA per frame, two resident per SM) are bound by their sequential row
  // chains, so a chunk is exactly two frames per SM: 296 on a B200
  const int chunk_max = 2 * (h->sm_count > 0 ? h->sm_count : 148);
  const int chunk = n_frames < chunk_max ? n_frames : chunk_max;
  ESLAM_CUDA_OK(h, w.pts.reserve(chunk * NP * 16)); ESLAM_CUDA_OK(h, w.nrm.reserve(chunk * NP * 16));
  ESLAM_CUDA_OK(h, w.iiX.reserve(chunk * NI * 8)); ESLAM_CUDA_OK(h, w.iiY.reserve(chunk * NI * 8));
  ESLAM_CUDA_OK(h, w.dm.reserve(chunk * NP * 4)); ESLAM_CUDA_OK(h, w.parent.reserve(chunk * NP * 4)); ESLAM_CUDA_OK(h, w.label.reserve(chunk * NP * 4));
  ESLAM_CUDA_OK(h, w.count.reserve(chunk * NP * 4)); ESLAM_CUDA_OK(h, w.maxidx.reserve(chunk * NP * 4));
  ESLAM_CUDA_OK(h, w.cmin.reserve(chunk * NP * 4)); ESLAM_CUDA_OK(h, w.cmax.reserve(chunk * NP * 4));
  ESLAM_CUDA_OK(h, w.cand_root.reserve((size_t)chunk * kMaxRegions * 4)); ESLAM_CUDA_OK(h, w.cand_n.reserve((size_t)chunk * 4));
  ESLAM_CUDA_OK(h, w.cand_coef.reserve((size_t)chunk * kMaxRegions * 16)); ESLAM_CUDA_OK(h, w.cand_valid.reserve((size_t)chunk * kMaxRegions * 4));
  ESLAM_CUDA_OK(h, w.cand_size.reserve((size_t)chunk * kMaxRegions * 4));
  ESLAM_CUDA_OK(h, w.crange.reserve((size_t)chunk * 6 * sizeof(unsigned)));
  ESLAM_CUDA_OK(h, w.cpart.reserve((size_t)chunk * ((NP + 255) / 256) * 6 * sizeof(unsigned)));
  ESLAM_CUDA_OK(h, w.out_plane.reserve((size_t)chunk * 32)); ESLAM_CUDA_OK(h, w.out_found.reserve((size_t)chunk * 4)); ESLAM_CUDA_OK(h, w.out_inl.reserve((size_t)chunk * 4));
  PlaneFramePtrs P;
  P.pts = w.pts.as<float4>(); P.nrm = w.nrm.as<float4>(); P.iiX = w.iiX.as<double>(); P.iiY = w.iiY.as<double>(); P.dm = w.dm.as<float>();
  P.parent = w.parent.as<int>(); P.label = w.label.as<int>(); P.count = w.count.as<int>(); P.maxidx = w.maxidx.as<int>(); P.cmin = w.cmin.as<int>(); P.cmax = w.cmax.as<int>();
  P.cand_root = w.cand_root.as<int>(); P.cand_n = w.cand_n.as<int>(); P.cand_coef = w.cand_coef.as<float4>(); P.cand_valid = w.cand_valid.as<int>();
  P.crange = w.crange.as<unsigned>(); P.cpart = w.cpart.as<unsigned>(); P.exact_mode = h->plane_exact ? 1 : 0;
  P.cand_size = w.cand_size.as<int>(); P.out_plane = w.out_plane.as<double>(); P.out_found = w.out_found.as<int>(); P.out_inl = w.out_inl.as<int>();
  PlaneParamsDev prm;
  prm.fx = h->cam.fx; prm.fy = h->cam.fy; prm.cx = h->cam.cx; prm.cy = h->cam.cy; prm.scale = h->cam.scale; prm.rows = H; prm.cols = Wd;
  const float AngTh = (float)h->prm.plane_angle_deg;
  prm.dis_th = (float)h->prm.plane_dist;
  prm.ang_cos = cosf((float)(0.017453 * AngTh));   // mps.setAngularThreshold(0.017453 * AngTh) -> PlaneCoefficientComparator: cosf
  prm.min_size = h->prm.plane_min_size;
  const int kIntSmemBytes = kIntWarps * 32 * kIntRingPitch * (int)sizeof(float);
  const int tpc = ((Wd + 31) / 32) * 32;
  const int strip_bytes = kStripRows * Wd * (int)sizeof(int);
  if (!h->attr_plane) {   // kernel attributes are per device: once per handle (the image size is fixed per handle)
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(integral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kIntSmemBytes));
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(integral_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // 2 CTAs (2 x 97 KB) per SM
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFitSmemBytes));
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(ccl_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, strip_bytes));
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(normals_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNxSmemBytes));
    h->attr_plane = true;
  }
  const size_t frame_stride = (size_t)H * pitch;
  // A chunk runs as two halves on two streams: most kernels of the stage are latency bound with one CTA per frame (distance map,
  // refinement, candidates, fit), so kernels of the two halves overlap on the SMs; the second stream is ordered after the caller's.
  const int nblocks = (int)((NP + 255) / 256);
  auto shifted = [&](const PlaneFramePtrs &B, int o) {
    PlaneFramePtrs Q = B;
    Q.pts += (size_t)o * NP; Q.nrm += (size_t)o * NP; Q.iiX += (size_t)o * NI; Q.iiY += (size_t)o * NI; Q.dm += (size_t)o * NP;
    Q.parent += (size_t)o * NP; Q.label += (size_t)o * NP; Q.count += (size_t)o * NP; Q.maxidx += (size_t)o * NP; Q.cmin += (size_t)o * NP; Q.cmax += (size_t)o * NP;
    Q.cand_root += (size_t)o * kMaxRegions; Q.cand_n += o; Q.cand_coef += (size_t)o * kMaxRegions; Q.cand_valid += (size_t)o * kMaxRegions; Q.cand_size += (size_t)o * kMaxRegions;
    Q.out_plane += (size_t)o * 4; Q.out_found += o; Q.out_inl += o; Q.crange += (size_t)o * 6; Q.cpart += (size_t)o * nblocks * 6;
    return Q;
  };
  if (!h->plane_stream2) {
    ESLAM_CUDA_OK(h, cudaStreamCreateWithFlags(&h->plane_stream2, cudaStreamNonBlocking));
    ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->plane_ev[0], cudaEventDisableTiming));
    ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->plane_ev[1], cudaEventDisableTiming));
  }
  for (int f0 = 0; f0 < n_frames; f0 += chunk) {
    const int Fc = (n_frames - f0) < chunk ? (n_frames - f0) : chunk;
    const int nsub = (debug_stage == 0 && Fc >= 32 && h->plane_streams > 1) ? 2 : 1;
    if (nsub == 2) { ESLAM_CUDA_OK(h, cudaEventRecord(h->plane_ev[0], s)); ESLAM_CUDA_OK(h, cudaStreamWaitEvent(h->plane_stream2, h->plane_ev[0], 0)); }
    for (int sub = 0; sub < nsub; sub++) {
    const int o = sub == 0 ? 0 : Fc / 2, F = nsub == 1 ? Fc : (sub == 0 ? Fc / 2 : Fc - Fc / 2);
    cudaStream_t s_outer = s;
    {
    cudaStream_t s = sub == 0 ? s_outer : h->plane_stream2;
    const PlaneFramePtrs P_outer = P;
    const PlaneFramePtrs P = shifted(P_outer, o);
    const int f0s = f0 + o;
    const uint16_t *d0 = depth_dev + (size_t)f0s * frame_stride;
    dim3 gpix((unsigned)((NP + 255) / 256), (unsigned)F);
    points_kernel<<<gpix, 256, 0, s>>>(d0, pitch, frame_stride, prm, P); ESLAM_LAUNCHED(h);
    crange_kernel<<<F, 192, 0, s>>>(P, (int)gpix.x); ESLAM_LAUNCHED(h);
    integral_kernel<<<F, 32 * kIntWarps, kIntSmemBytes, s>>>(prm, P); ESLAM_LAUNCHED(h);
    distmap_kernel<<<F, tpc, 0, s>>>(prm, P); ESLAM_LAUNCHED(h);
    if (debug_stage == 1) { ESLAM_CUDA_OK(h, cudaMemcpyAsync(debug_out, P.dm, NP * 4, cudaMemcpyDeviceToHost, s)); ESLAM_CUDA_OK(h, cudaStreamSynchronize(s)); return ESLAM_OK; }
    normals_kernel<<<dim3((gpix.x + 7) / 8, (unsigned)F), 256, 0, s>>>(prm, P); ESLAM_LAUNCHED(h);   // frames that need the recurrence (exit at once otherwise)
    normals_exact_kernel<<<dim3((unsigned)((Wd + kNxT - 1) / kNxT), (unsigned)((H + kNxT - 1) / kNxT), (unsigned)F), 512, kNxSmemBytes, s>>>(prm, P); ESLAM_LAUNCHED(h);
    if (debug_stage == 2) { ESLAM_CUDA_OK(h, cudaMemcpyAsync(debug_out, P.nrm, NP * 16, cudaMemcpyDeviceToHost, s)); ESLAM_CUDA_OK(h, cudaStreamSynchronize(s)); return ESLAM_OK; }
    const int nstrips = (H + kStripRows - 1) / kStripRows;
    ccl_strip_kernel<<<dim3((unsigned)nstrips, (unsigned)F), tpc, strip_bytes, s>>>(prm, P); ESLAM_LAUNCHED(h);
    if (nstrips > 1) { ccl_seam_kernel<<<dim3((unsigned)((Wd + 255) / 256), (unsigned)(nstrips - 1), (unsigned)F), 256, 0, s>>>(prm, P); ESLAM_LAUNCHED(h); }
    ccl_flatten_kernel<<<gpix, 256, 0, s>>>(prm, P); ESLAM_LAUNCHED(h);
    if (debug_stage == 3) { ESLAM_CUDA_OK(h, cudaMemcpyAsync(debug_out, P.label, NP * 4, cudaMemcpyDeviceToHost, s)); ESLAM_CUDA_OK(h, cudaStreamSynchronize(s)); return ESLAM_OK; }
    candidates_kernel<<<F, 1024, 0, s>>>(prm, P); ESLAM_LAUNCHED(h);
    dim3 gfit(8, (unsigned)F);    // 8 CTAs per frame stride over that frame's candidate regions
    fit_kernel<<<gfit, kFitThreads, kFitSmemBytes, s>>>(prm, P); ESLAM_LAUNCHED(h);
    model_image_kernel<<<gpix, 256, 0, s>>>(prm, P); ESLAM_LAUNCHED(h);
    refine_kernel<<<F, tpc, 0, s>>>(prm, P); ESLAM_LAUNCHED(h);
    if (debug_stage == 4) { ESLAM_CUDA_OK(h, cudaMemcpyAsync(debug_out, P.label, NP * 4, cudaMemcpyDeviceToHost, s)); ESLAM_CUDA_OK(h, cudaStreamSynchronize(s)); return ESLAM_OK; }
    select_kernel<<<(F + 63) / 64, 64, 0, s>>>(prm, P, F); ESLAM_LAUNCHED(h);
    ESLAM_CUDA_OK(h, cudaGetLastError());
    ESLAM_CUDA_OK(h, cudaMemcpyAsync(plane_cam_host + 4 * (size_t)f0s, P.out_plane, sizeof(double) * 4 * F, cudaMemcpyDeviceToHost, s));
    ESLAM_CUDA_OK(h, cudaMemcpyAsync(found_host + f0s, P.out_found, sizeof(int) * F, cudaMemcpyDeviceToHost, s));
    if (n_inliers_host) ESLAM_CUDA_OK(h, cudaMemcpyAsync(n_inliers_host + f0s, P.out_inl, sizeof(int) * F, cudaMemcpyDeviceToHost, s));
    }
    }
    if (nsub == 2) { ESLAM_CUDA_OK(h, cudaEventRecord(h->plane_ev[1], h->plane_stream2)); ESLAM_CUDA_OK(h, cudaStreamWaitEvent(s, h->plane_ev[1], 0)); }
    ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));   // the results of the chunk are on the host (the one host synchronisation of the entry)
  }
  return ESLAM_OK;
}

int plane_extract(EslamHandle *h, cudaStream_t s, int n_frames, const uint16_t *depth_dev, size_t pitch, double *plane_cam_host, int *found_host,
                  int *n_inliers_host) {
  return plane_extract_dbg(h, s, n_frames, depth_dev, pitch, plane_cam_host, found_host, n_inliers_host, 0, nullptr);
}

void plane_free(EslamHandle *h) {
  if (!h->pl) return;
  PlaneWorkspace &w = *h->pl;
  DevBuf *bufs[] = {&w.pts, &w.nrm, &w.iiX, &w.iiY, &w.dm, &w.parent, &w.label, &w.count, &w.maxidx, &w.cmin, &w.cmax, &w.cand_root, &w.cand_n, &w.cand_coef, &w.cand_valid, &w.cand_size,
                    &w.out_plane, &w.out_found, &w.out_inl, &w.crange, &w.cpart};
  for (DevBuf *b : bufs) b->release();
  w.pin.release();
  if (h->plane_stream2) { cudaStreamDestroy(h->plane_stream2); cudaEventDestroy(h->plane_ev[0]); cudaEventDestroy(h->plane_ev[1]); h->plane_stream2 = nullptr; }
  delete h->pl; h->pl = nullptr;
}

}  // namespace eslam

// extract.cu -- single-frame ellipsoid extraction, one CTA per (frame, bbox) object, batched over frames.
//
// Replaces EllipsoidExtractor::ExtractPointCloud + ProcessPCA + frame construction
//   src/pca/EllipsoidExtractor.cpp:86-168 (crop, 1 cm voxel grid, to world, supporting-plane filter, centre, clusters)
//   src/pca/EllipsoidExtractor.cpp:170-267, 308-347, 760-788 (PCA, chirality, gravity alignment, object frame)
//   src/symmetry/PointCloudFilter.cpp:21-56, 249-261 ; src/core/Geometry.cpp:44-66,160-166
// and the PCL calls behind them (VoxelGrid, EuclideanClusterExtraction, compute3DCentroid,
// computeCovarianceMatrixNormalized) with B200-native data-parallel algorithms that give the SAME integer results:
//   voxel grid        -> PCL voxel index, stable block radix sort, per-voxel float centroid in ascending point order
//   euclidean cluster -> uniform hash grid (cell ~ tolerance) + lock-free union-find; membership == connected components
//                        of {||p-q||^2 < r^2 in float}; order = size desc, then first index (PCL's discovery order)
// Data layout: per-object slices of flat HBM arrays (ExtractPtrs); a CTA streams its slice, which stays L2 resident.
#include "common.cuh"
#include "dmath.cuh"
#include "block_prims.cuh"
#include "extract_types.cuh"
#include "internal.h"

namespace eslam {

__device__ __forceinline__ int lower_bound_u32(const unsigned *a, int n, unsigned key) {
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}
// relaxed CTA-scope accesses for the lock-free union-find (one CTA per object): `volatile` would compile to system-scope strong
// loads/stores (LD.E.STRONG.SYS), which stall the dispatcher; the parents live in shared memory or in this CTA's HBM slice
__device__ __forceinline__ int vload(const int *p) { int v; asm volatile("ld.relaxed.cta.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void vstore(int *p, int v) { asm volatile("st.relaxed.cta.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// find with path halving: parent pointers only ever move towards the root (a smaller index), so the racy
// parent[i] = grandparent stores are benign (they always write an ancestor of i).
__device__ int uf_find(int *parent, int i) {
  int p = vload(parent + i);
  while (p != i) {
    int gp = vload(parent + p);
    if (gp != p) vstore(parent + i, gp);
    i = p; p = gp;
  }
  return i;
}
__device__ void uf_union(int *parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }   // hook the larger root under the smaller: final root == min index
    int old = atomicCAS(parent + a, a, b);
    if (old == a) return;
  }
}
// lower bound of `key` in the sorted keys k0[0..n): first through the shared-memory table of every 16th key, then four steps inside
// one 64-byte block of global memory (plain binary search when the table does not fit); uniform trip counts
__device__ __forceinline__ int lower_bound_2l(const unsigned *k0, const unsigned *coarse, int n, int ncoarse, int pow2, int pow2c, bool use_coarse, unsigned key) {
  int q = 0;
  if (use_coarse) {
    int cb = 0;   // number of 16-key blocks whose first key is < key
    for (int st = pow2c; st > 0; st >>= 1) { int t = cb + st; if (t <= ncoarse && coarse[t - 1] < key) cb = t; }
    if (cb > 0) {
      q = ((cb - 1) << 4) + 1;                     // k0[(cb-1)*16] < key; the bound lies in ((cb-1)*16, cb*16]
      const int hi = min(n, cb << 4);
#pragma unroll
      for (int st = 8; st > 0; st >>= 1) { int t = q + st; if (t <= hi && k0[t - 1] < key) q = t; }
    }
  } else {
    for (int st = pow2; st > 0; st >>= 1) { int t = q + st; if (t <= n && k0[t - 1] < key) q = t; }
  }
  return q;
}
__device__ __forceinline__ float l2_simple(float4 a, float4 b) {  // flann::L2_Simple<float>, no contraction
  float r = 0.0f, d;
  d = __fsub_rn(a.x, b.x); r = __fadd_rn(r, __fmul_rn(d, d));
  d = __fsub_rn(a.y, b.y); r = __fadd_rn(r, __fmul_rn(d, d));
  d = __fsub_rn(a.z, b.z); r = __fadd_rn(r, __fmul_rn(d, d));
  return r;
}

// PCL VoxelGrid indexing + stable sort.  On return k0[j]/v0[j] hold sorted voxel keys / original indices.
// `wide`: >= 40 KB of shared memory that is idle during the sort (8 x 2048 16-bit counters + 2048 digit offsets)
__device__ void voxel_sort(const float4 *pts, int n, float leaf, unsigned *k0, unsigned *v0, unsigned *k1, unsigned *v1, BlockScratch &S, RadixScratch &RS, int *wide, long long *dbg = nullptr) {
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float4 p = pts[i];
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  block_minmax3(mn, mx, S);
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();
  const float inv = __fdiv_rn(1.0f, leaf);
  long long dx = (long long)__fmul_rn(__fsub_rn(mx[0], mn[0]), inv) + 1;
  long long dy = (long long)__fmul_rn(__fsub_rn(mx[1], mn[1]), inv) + 1;
  long long dz = (long long)__fmul_rn(__fsub_rn(mx[2], mn[2]), inv) + 1;
  if (dx * dy * dz > 2147483647LL) {  // PCL returns the input unfiltered: every point is its own voxel
    for (int i = threadIdx.x; i < n; i += blockDim.x) { k0[i] = (unsigned)i; v0[i] = (unsigned)i; }
    __syncthreads();
    return;
  }
  int minb[3], divb[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    minb[a] = (int)floorf(__fmul_rn(mn[a], inv));
    int maxb = (int)floorf(__fmul_rn(mx[a], inv));
    divb[a] = maxb - minb[a] + 1;
  }
  const int mul1 = divb[0], mul2 = divb[0] * divb[1];
  const float fb0 = (float)minb[0], fb1 = (float)minb[1], fb2 = (float)minb[2];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float4 p = pts[i];
    int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), fb0);
    int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), fb1);
    int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), fb2);
    k0[i] = (unsigned)(i0 + i1 * mul1 + i2 * mul2);
    v0[i] = (unsigned)i;
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[2] = clock64();
  unsigned long long maxkey = (unsigned long long)divb[0] * (unsigned long long)divb[1] * (unsigned long long)divb[2];
  int bits = (maxkey >= 0xffffffffULL) ? 32 : bits_for((unsigned)maxkey);
  // wide passes (digits of up to 11 bits: two passes up to 22 key bits, three up to 32) instead of 8-bit ones; the counters live in the
  // idle union-find region.  The 1 cm grid over a crop that reaches the 6 m range limit has 27-bit keys: 3 passes instead of 4.
  if (bits > 8 && n <= 65535 && blockDim.x == 256)
    block_radix_sort_wide(k0, v0, k1, v1, n, bits, reinterpret_cast<unsigned short *>(wide), reinterpret_cast<unsigned *>(wide) + 8 * 2048 / 2, 2048, S);
  else block_radix_sort(k0, v0, k1, v1, n, bits, RS);
}

constexpr int kExtractThreads = 256;   // 3 CTAs/SM: different objects sit in different phases, hiding the many block barriers
constexpr int kSmemParents = 11264;    // 44 KB of dynamic shared memory per CTA: three CTAs (72.4 KB each with the static part) leave ~13 KB per SM for the small kernels that run beside them (bulk-copy pull, symmetry step)

// One tile (kExtractThreads sorted positions) of (voxel key, point) staged in shared memory -- every thread issues its own
// independent gather -- after which the first position of each voxel sums its members from shared memory.  Aliases RadixScratch
// (idle between sorts).
struct VoxelTile { unsigned key[kExtractThreads + 1]; float x[kExtractThreads], y[kExtractThreads], z[kExtractThreads]; };
static_assert(sizeof(VoxelTile) <= sizeof(RadixScratch), "VoxelTile aliases RadixScratch");
// Returns whether this thread's position is the head of a voxel; if so `c` is the voxel's float centroid (members added in
// ascending sorted position == ascending original index, PCL's centroid order).  Contains one __syncthreads(); the caller's next
// block-wide barrier (its compaction scan) separates this tile's reads from the next tile's staging.
__device__ __forceinline__ bool voxel_tile_centroid(const float4 *pts, const unsigned *k0, const unsigned *v0, int tile, int n, VoxelTile &T, float4 &c) {
  const int tid = threadIdx.x, j = tile + tid;
  unsigned key = 0; float4 p = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (j < n) { key = k0[j]; p = pts[v0[j]]; }
  T.key[tid + 1] = key; T.x[tid] = p.x; T.y[tid] = p.y; T.z[tid] = p.z;
  if (tid == 0) T.key[0] = (tile > 0) ? k0[tile - 1] : 0u;
  __syncthreads();
  const bool head = (j < n) && (j == 0 || T.key[tid] != key);
  if (head) {
    float sx = __fadd_rn(0.0f, p.x), sy = __fadd_rn(0.0f, p.y), sz = __fadd_rn(0.0f, p.z);
    const int lim = min(kExtractThreads, n - tile);
    int jj = tid + 1;
    while (jj < lim && T.key[jj + 1] == key) { sx = __fadd_rn(sx, T.x[jj]); sy = __fadd_rn(sy, T.y[jj]); sz = __fadd_rn(sz, T.z[jj]); jj++; }
    int members = jj - tid;
    if (jj == kExtractThreads) {   // the voxel continues past the tile: finish from global memory
      int g = tile + kExtractThreads;
      while (g < n && k0[g] == key) { const float4 q = pts[v0[g]]; sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y); sz = __fadd_rn(sz, q.z); g++; members++; }
    }
    const float cnt = (float)members;
    c = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), 0.0f);
  }
  return head;
}

template <int MINB, int SMEM_PARENTS, bool UQ, int PF>
__global__ void __launch_bounds__(kExtractThreads, MINB) extract_kernel(const uint16_t *__restrict__ depth, size_t pitch, size_t frame_stride, const ObjMeta *__restrict__ metas,
                                                      ObjState *__restrict__ states, ExtractParamsDev prm, ExtractPtrs W) {
  __shared__ BlockScratch S;
  __shared__ RadixScratch RS;
  __shared__ ScanScratch SS;
  __shared__ int s_sel;
  // UQ: per-warp queue of the (i, j) pairs that still need a union (packed 16 + 16 bits: at most 34 080 stride-3 samples per box)
  __shared__ unsigned s_uq[UQ ? kExtractThreads / 32 : 1][UQ ? 64 : 1];
  VoxelTile &VT = *reinterpret_cast<VoxelTile *>(&RS);
  extern __shared__ int s_parent[];   // union-find parents / cluster labels when the filtered cloud fits (kSmemParents), else HBM
  int phase = 0;
  const int tid = threadIdx.x;
  const ObjMeta m = metas[blockIdx.x];
  ObjState &st = states[blockIdx.x];
  if (tid == 0) {
    st.status = m.pre_status;
    st.n_crop = st.n_voxel = st.n_plane = st.n_clusters = st.n_cluster = st.n_sym = st.n_final = 0;
    st.run_sym = 0; st.symtype = m.symtype; st.best_seed = -1; st.done = 0; st.n_evals = 0; st.n_queries = 0; st.n_points = 0; st.n_hits = 0; st.n_ref_evals = 0; st.n_ref_points = 0; st.n_verify_fail = 0; for (int k = 0; k < 4; k++) st.sym_clk[k] = 0;
    for (int k = 0; k < 4; k++) st.bbox_i[k] = (int)m.bbox[k];
    for (int s = 0; s < 9; s++) { st.seed_prob[s] = st.seed_init[s] = st.seed_final[s] = st.seed_dual[s] = 0; for (int k = 0; k < 4; k++) st.seed_plane[s][k] = 0; }
  }
  if (m.pre_status != 0) return;
  if (tid == 0) st.phase_clk[0] = clock64();
  float4 *A = W.ptsA + m.off, *B = W.ptsB + m.off, *Cc = W.ptsC + m.off;
  unsigned *k0 = W.k0 + m.off, *k1 = W.k1 + m.off, *v0 = W.v0 + m.off, *v1 = W.v1 + m.off;
  int *parent = W.parent + m.off, *count = W.count + m.off;   // parent is re-pointed to shared memory once n3 is known
  const uint16_t *img = depth + (size_t)m.frame * frame_stride;
  const se3 Twc = se3_load(m.Twc);
  const hom34 Twc_h = se3_hom(Twc);

  // ---------------- S1: stride-3 crop + back-projection (PointCloudFilter.cpp:31-52) ----------------
  int n1 = 0;
  {
    const int total = m.nx * m.ny;
    for (int tile = 0; tile < total; tile += 4 * kExtractThreads) {
      unsigned short d[4]; int px[4], py[4]; bool keep[4]; int ex[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {   // four independent strided loads in flight per thread
        const int i = tile + k * kExtractThreads + tid;
        d[k] = 0; px[k] = 0; py[k] = 0;
        if (i < total) {
          const int iy = i / m.nx, ix = i - iy * m.nx;
          px[k] = m.x1 + 3 * ix; py[k] = m.y1 + 3 * iy;
          d[k] = img[(size_t)py[k] * pitch + px[k]];
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double Z = d[k] / prm.scale;
        keep[k] = (tile + k * kExtractThreads + tid < total) && !(Z <= 0.1 || Z > prm.depth_range);
      }
      int tot;
      block_flag_scan<4>(keep, ex, tot, SS, phase);
#pragma unroll
      for (int k = 0; k < 4; k++) if (keep[k]) {
        const double Z = d[k] / prm.scale;
        const double X = (px[k] - prm.cx) * Z / prm.fx;
        const double Y = (py[k] - prm.cy) * Z / prm.fy;
        A[n1 + ex[k]] = make_float4((float)X, (float)Y, (float)Z, 0.0f);   // QuadricPointCloudToPcl: double -> float
        if (W.d_crop) { double *q = W.d_crop + 3 * (size_t)(m.off + n1 + ex[k]); q[0] = X; q[1] = Y; q[2] = Z; }
      }
      n1 += tot;
    }
  }
  __syncthreads();
  if (tid == 0) st.phase_clk[1] = clock64();
  // ---------------- S2 + S3 + S4: 1 cm voxel grid -> world -> supporting-plane filter ----------------
  int n2 = 0, n3 = 0;
  if (n1 > 0) {
    voxel_sort(A, n1, prm.voxel_leaf, k0, v0, k1, v1, S, RS, s_parent);
    if (tid == 0) st.phase_clk[2] = clock64();
    for (int tile = 0; tile < n1; tile += blockDim.x) {
      bool head = false, keep = false;
      d3 pc = mk3(0, 0, 0), pw = mk3(0, 0, 0);
      float4 c;
      head = voxel_tile_centroid(A, k0, v0, tile, n1, VT, c);
      if (head) {
        pc = mk3((double)c.x, (double)c.y, (double)c.z);      // pclToQuadricPointCloud: float -> double
        pw = hom_apply(Twc_h, pc);                               // transformPointCloud
        keep = plane_signed_dist(prm.plane, pw) > prm.plane_margin;  // ApplySupportingPlaneFilter
      }
      int hr, kr, toth, totk;
      block_flag_scan2(head, keep, hr, kr, toth, totk, SS, phase);
      if (head && W.d_voxel) { double *q = W.d_voxel + 3 * (size_t)(m.off + n2 + hr); q[0] = pc.x; q[1] = pc.y; q[2] = pc.z; }
      if (keep) {
        B[n3 + kr] = make_float4((float)pw.x, (float)pw.y, (float)pw.z, 0.0f);   // QuadricPointCloudToPclXYZ
        if (W.d_world) { double *q = W.d_world + 3 * (size_t)(m.off + n3 + kr); q[0] = pw.x; q[1] = pw.y; q[2] = pw.z; }
      }
      n2 += toth; n3 += totk;
    }
  }
  __syncthreads();
  if (tid == 0) { st.n_crop = n1; st.n_voxel = n2; st.n_plane = n3; st.phase_clk[3] = clock64(); }
  if (n3 < 1) { if (tid == 0) st.status = ESLAM_OBJ_NO_POINTS_AFTER_PLANE; return; }
  if (n3 <= SMEM_PARENTS) parent = s_parent;   // block-uniform: pointer chasing at shared-memory instead of L2 latency

  // ---------------- S5: GetCenter (EllipsoidExtractor.cpp:583-643) ----------------
  if (tid == 0) {
    int x = (int)((m.bbox[0] + m.bbox[2]) / 2.0);
    int y = (int)((m.bbox[1] + m.bbox[3]) / 2.0);
    const int point_num = 10;
    int x_delta = (int)(fabs(m.bbox[0] - m.bbox[2]) / 4.0 / point_num);
    int y_delta = (int)(fabs(m.bbox[1] - m.bbox[3]) / 4.0 / point_num);
    double c0 = 0, c1 = 0, c2 = 0; int cnt = 0;
    for (int x_id = -point_num / 2; x_id < point_num / 2; x_id++)
      for (int y_id = -point_num / 2; y_id < point_num / 2; y_id++) {
        int x_ = x + x_id * x_delta, y_ = y + y_id * y_delta;
        if (x_ < 0 || y_ < 0 || x_ >= prm.cols || y_ >= prm.rows) continue;
        unsigned short d = img[(size_t)y_ * pitch + x_];
        float pz = (float)(d / prm.scale);
        if ((double)pz <= 0.1 || (double)pz > prm.depth_range) continue;
        float px = (float)((x_ - prm.cx) * (double)pz / prm.fx);
        float py = (float)((y_ - prm.cy) * (double)pz / prm.fy);
        c0 += (double)px; c1 += (double)py; c2 += (double)pz; cnt++;
      }
    int ok = cnt >= 2;
    if (ok) {
      double N = (double)cnt;
      d3 cw = hom_apply(Twc_h, mk3(c0 / N, c1 / N, c2 / N));
      S.bcast_d[0] = cw.x; S.bcast_d[1] = cw.y; S.bcast_d[2] = cw.z;
      st.center[0] = cw.x; st.center[1] = cw.y; st.center[2] = cw.z;
    }
    S.bcast_i[0] = ok;
  }
  __syncthreads();
  if (!S.bcast_i[0]) { if (tid == 0) st.status = ESLAM_OBJ_NO_CENTER; return; }
  const d3 center = mk3(S.bcast_d[0], S.bcast_d[1], S.bcast_d[2]);
  __syncthreads();

  // ---------------- S6: Euclidean clustering (EllipsoidExtractor.cpp:646-737) ----------------
  {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = tid; i < n3; i += blockDim.x) {
      float4 p = B[i];
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
    block_minmax3(mn, mx, S);
    float cs = prm.cluster_cell;
    long long gx, gy, gz;
    while (true) {
      gx = (long long)((mx[0] - mn[0]) / cs) + 1; gy = (long long)((mx[1] - mn[1]) / cs) + 1; gz = (long long)((mx[2] - mn[2]) / cs) + 1;
      if (gx * gy * gz < 2147483647LL) break;
      cs *= 2.0f;
    }
    const int nx = (int)gx, ny = (int)gy, nz = (int)gz;
    for (int i = tid; i < n3; i += blockDim.x) {
      float4 p = B[i];
      int cx = (int)((p.x - mn[0]) / cs), cy = (int)((p.y - mn[1]) / cs), cz = (int)((p.z - mn[2]) / cs);
      k0[i] = (unsigned)((cz * ny + cy) * nx + cx);
      v0[i] = (unsigned)i;
      parent[i] = i; count[i] = 0;
    }
    __syncthreads();
    {
      const int cbits = bits_for((unsigned)(gx * gy * gz));
      if (cbits > 8 && cbits <= 20 && n3 <= 65535)
        block_radix_sort_wide(k0, v0, k1, v1, n3, cbits, reinterpret_cast<unsigned short *>(RS.radix), RS.digit_base, 1024, S);
      else block_radix_sort(k0, v0, k1, v1, n3, cbits, RS);
    }
    if (tid == 0) st.phase_clk[4] = clock64();
    const float r2 = prm.cluster_r2;
    int pow2 = 1;
    while (pow2 < n3) pow2 <<= 1;
    // two-level lower bound: the first key of every 16-key block lives in shared memory (RadixScratch is idle now), so a search costs
    // ~log2(n/16) shared-memory steps plus 4 global steps inside one 64-byte block instead of ~log2(n) dependent global loads
    unsigned *coarse = RS.radix;
    const int ncoarse = (n3 + 15) >> 4;
    const bool use_coarse = ncoarse <= kRadixWarps * 256;
    int pow2c = 1;
    while (pow2c < ncoarse) pow2c <<= 1;
    if (use_coarse) for (int b = tid; b < ncoarse; b += blockDim.x) coarse[b] = k0[b << 4];
    // points in cell-sorted order with their original index in .w (the crop buffer A is free now): a neighbour candidate is one
    // sequential 16-byte load instead of the dependent pair v0[q] -> B[j]
    float4 *Bs = A;
    for (int pos = tid; pos < n3; pos += blockDim.x) { const int j = (int)v0[pos]; float4 pj = B[j]; pj.w = __int_as_float(j); Bs[pos] = pj; }
    __syncthreads();
    // Warp-converged neighbour search over the FORWARD half of the 27-cell neighbourhood (each unordered pair is tested exactly once,
    // from the point that comes first in (z, y, x) cell order; l2_simple is symmetric to the last bit): the rest of the own cell and
    // cell x+1 by walking on from the point's own sorted position, then rows (dz,dy) = (0,+1), (+1,-1), (+1,0), (+1,+1) over x-1..x+1.
    // The data-dependent scan of a row is re-converged with __any_sync each step, so divergent union-find retries never serialise
    // the rest of the warp.
    // UQ: a pair inside the radius whose points do not share a parent yet is rare per lane and iteration (the union-find code ran
    // with ~4 of 32 lanes active, 16 % of the kernel's issued instructions); the pairs go to a per-warp queue instead and are united
    // 32 at a time with every lane busy.  The final partition does not depend on the order of the unions.
    const unsigned lt_ = (1u << (tid & 31)) - 1u;
    unsigned *wq = s_uq[UQ ? (tid >> 5) : 0];
    int wcnt = 0;
    for (int base = 0; base < n3; base += blockDim.x) {
      const int pos = base + tid;
      const bool act = pos < n3;
      int i = 0, cx = 0, cy = 0, cz = 0;
      float4 p = make_float4(0, 0, 0, 0);
      if (act) {
        p = Bs[pos]; i = __float_as_int(p.w);
        cx = (int)((p.x - mn[0]) / cs); cy = (int)((p.y - mn[1]) / cs); cz = (int)((p.z - mn[2]) / cs);
      }
      const int xlo = max(cx - 1, 0), xhi = min(cx + 1, nx - 1);
#pragma unroll 1
      for (int row = 0; row < 5; row++) {
        const int zz = cz + (row >= 2 ? 1 : 0), yy = cy + (row == 0 ? 0 : (row == 1 ? 1 : row - 3));
        const bool valid = act && zz < nz && yy >= 0 && yy < ny;
        unsigned khi = 0; int q = n3;
        if (valid) {
          khi = (unsigned)((zz * ny + yy) * nx + xhi);
          q = (row == 0) ? pos + 1 : lower_bound_2l(k0, coarse, n3, ncoarse, pow2, pow2c, use_coarse, (unsigned)((zz * ny + yy) * nx + xlo));
        }
        // software pipelined walk: key and point of candidate q+1 are requested before candidate q is tested (a second lower bound for
        // the end of the row instead of the key loads was measured slower: 65 instructions per search against ~3 per candidate)
        // (PF = 2: two candidates ahead -- one iteration of the resident warps is shorter than an L2 round trip)
        unsigned kq = 0xffffffffu; float4 pq = make_float4(0, 0, 0, 0);
        unsigned kq1 = 0xffffffffu; float4 pq1 = make_float4(0, 0, 0, 0);
        if (valid && q < n3) { kq = k0[q]; pq = Bs[q]; }
        if (PF == 2 && valid && q + 1 < n3) { kq1 = k0[q + 1]; pq1 = Bs[q + 1]; }
        while (true) {
          const bool more = valid && q < n3 && kq <= khi;
          if (!__any_sync(0xffffffffu, more)) break;
          unsigned kn = 0xffffffffu; float4 pn = make_float4(0, 0, 0, 0);
          if (more && q + PF < n3) { kn = k0[q + PF]; pn = Bs[q + PF]; }
          bool want = false; int j = 0;
          if (more) {
            if (l2_simple(p, pq) < r2) {   // RadiusResultSet::addPoint: dist < radius^2
              j = __float_as_int(pq.w);
              want = vload(parent + i) != vload(parent + j);   // same parent => same set already: skip the finds
              if (!UQ && want) uf_union(parent, i, j);
            }
            q++;
            if (PF == 2) { kq = kq1; pq = pq1; kq1 = kn; pq1 = pn; } else { kq = kn; pq = pn; }
          }
          if (UQ) {
            const unsigned wm = __ballot_sync(0xffffffffu, want);
            if (wm) {
              if (want) wq[wcnt + __popc(wm & lt_)] = (unsigned)i | ((unsigned)j << 16);
              wcnt += __popc(wm);
              if (wcnt >= 32) {
                __syncwarp();
                const unsigned e = wq[tid & 31];
                uf_union(parent, (int)(e & 0xffffu), (int)(e >> 16));
                const int rest = wcnt - 32;
                unsigned t = 0;
                if ((tid & 31) < rest) t = wq[32 + (tid & 31)];
                __syncwarp();
                if ((tid & 31) < rest) wq[tid & 31] = t;
                wcnt = rest;
                __syncwarp();
              }
            }
          }
        }
      }
      // Pointer jumping over all parents every W.uf_jump-th batch of 256 points (shared-memory parents only): hooks go "larger root under smaller root", so
      // the trees grow chains between compressions -- the finds were 27 % of the kernel's issued instructions -- and with flat trees
      // the "same parent" test above recognises most pairs that are already in one set.  Concurrent readers only ever see ancestors.
      if (UQ && W.uf_jump > 0 && parent == s_parent && ((base / kExtractThreads) % W.uf_jump) == W.uf_jump - 1) {
        __syncthreads();
        for (int u = tid; u < n3; u += kExtractThreads) { const int pp = parent[u], gp = parent[pp]; if (gp != pp) parent[u] = gp; }
        __syncthreads();
      }
    }
    if (UQ) {   // drain
      __syncwarp();
      if ((tid & 31) < wcnt) { const unsigned e = wq[tid & 31]; uf_union(parent, (int)(e & 0xffffu), (int)(e >> 16)); }
    }
    __syncthreads();
    if (tid == 0) st.phase_clk[5] = clock64();
    // flatten; cluster sizes with one atomic per (warp, root)
    for (int base = 0; base < n3; base += blockDim.x) {
      const int i = base + tid;
      const bool act = i < n3;
      const int r = act ? uf_find(parent, i) : -1 - (tid & 31);
      const unsigned peers = __match_any_sync(0xffffffffu, r);
      if (act) {
        k1[i] = (unsigned)r;
        if ((__ffs(peers) - 1) == (tid & 31)) atomicAdd(count + r, __popc(peers));
      }
    }
    __syncthreads();
    for (int i = tid; i < n3; i += blockDim.x) parent[i] = (int)k1[i];   // flattened labels (root = min index of the component)
    __syncthreads();
    // accepted clusters (size >= MinClusterSize), ascending root == PCL discovery order
    int K = 0;
    for (int tile = 0; tile < n3; tile += blockDim.x) {
      int i = tile + tid;
      const bool acc[1] = {(i < n3) && (parent[i] == i) && (count[i] >= prm.min_cluster)};
      int tot, ex[1];
      block_flag_scan<1>(acc, ex, tot, SS, phase);
      if (acc[0]) v1[K + ex[0]] = (unsigned)i;
      K += tot;
    }
    __syncthreads();
    if (tid == 0) {  // stable sort by size descending (EuclideanClusterExtraction::extract sorts clusters by size)
      for (int a = 1; a < K; a++) {
        unsigned r = v1[a]; int c = count[r]; int b = a - 1;
        while (b >= 0 && count[v1[b]] < c) { v1[b + 1] = v1[b]; b--; }
        v1[b + 1] = r;
      }
      st.n_clusters = K; s_sel = -1;
    }
    __syncthreads();
    // first cluster whose minimum distance to the centre is < CenterDis (or the only cluster)  :696-714
    for (int c = 0; c < K; c++) {
      int root = (int)v1[c];
      double mind = 999999;
      for (int i = tid; i < n3; i += blockDim.x) {
        if (parent[i] != root) continue;
        float4 p = B[i];
        double dx = center.x - (double)p.x, dy = center.y - (double)p.y, dz = center.z - (double)p.z;
        double dis = sqrt(dx * dx + dy * dy + dz * dz);
        mind = fmin(mind, dis);
      }
      mind = block_min_double(mind, S);
      bool hit = mind < prm.center_dis;
      if (tid == 0) { if (K == 1) s_sel = c; if (hit) s_sel = c; }
      __syncthreads();
      if (hit) break;
    }
    __syncthreads();
    if (W.d_label) {
      for (int i = tid; i < n3; i += blockDim.x) {
        int lab = -1;
        if (count[parent[i]] >= prm.min_cluster) for (int c = 0; c < K; c++) if ((int)v1[c] == parent[i]) { lab = c; break; }
        W.d_label[m.off + i] = lab;
      }
    }
  }
  const int sel = s_sel;
  if (sel < 0) { if (tid == 0) st.status = ESLAM_OBJ_NO_CLUSTER; return; }
  int nc = 0;
  {
    const int root = (int)v1[sel];
    for (int tile = 0; tile < n3; tile += blockDim.x) {
      int i = tile + tid;
      const bool in[1] = {(i < n3) && (parent[i] == root)};
      int tot, ex[1];
      block_flag_scan<1>(in, ex, tot, SS, phase);
      if (in[0]) Cc[nc + ex[0]] = B[i];
      nc += tot;
    }
  }
  __syncthreads();

  // ---------------- S7 + S8: PCA, chirality, gravity alignment, object frame ----------------
  if (tid == 0) st.phase_clk[6] = clock64();
  {
    double s3[3] = {0, 0, 0};
    for (int i = tid; i < nc; i += blockDim.x) { float4 p = Cc[i]; s3[0] += (double)p.x; s3[1] += (double)p.y; s3[2] += (double)p.z; }
    block_sum_double<3>(s3, S);
    const double N = (double)nc;
    const double c0 = s3[0] / N, c1 = s3[1] / N, c2 = s3[2] / N;   // compute3DCentroid
    double cv[6] = {0, 0, 0, 0, 0, 0};  // (1,1) (1,2) (2,2) (0,0) (0,1) (0,2)  computeCovarianceMatrix dense branch
    for (int i = tid; i < nc; i += blockDim.x) {
      float4 p = Cc[i];
      double px = (double)p.x - c0, py = (double)p.y - c1, pz = (double)p.z - c2;
      cv[0] += py * py; cv[1] += py * pz; cv[2] += pz * pz;
      cv[3] += px * px; cv[4] += py * px; cv[5] += pz * px;
    }
    block_sum_double<6>(cv, S);
    if (tid == 0) {
      mat3 cov;
      cov.a[4] = cv[0] / N; cov.a[5] = cv[1] / N; cov.a[8] = cv[2] / N; cov.a[0] = cv[3] / N; cov.a[1] = cv[4] / N; cov.a[2] = cv[5] / N;
      cov.a[3] = cov.a[1]; cov.a[6] = cov.a[2]; cov.a[7] = cov.a[5];
      double ev[3]; mat3 rot;
      eig3_sym(cov, ev, rot);
      d3 cen = mk3((double)(float)c0, (double)(float)c1, (double)(float)c2);   // PointType c is float  (:184-191)
      m3_setcol(rot, 2, cross3(m3_col(rot, 0), m3_col(rot, 1)));               // AdjustChirality
      // AlignZAxisToGravity :207-267
      d3 nrm = mk3(prm.plane[0], prm.plane[1], prm.plane[2]);
      d3 zax = unit3(nrm);
      double max_abs = 0; bool max_pos = false; int max_id = -1;
      for (int i = 0; i < 3; i++) {
        double ct = dot3(m3_col(rot, i), zax);
        double a = fabs(ct);
        if (a > max_abs) { max_abs = a; max_pos = ct > 0; max_id = i; }
      }
      if (max_id < 0) max_id = 0;
      mat3 sw; double cs[3];
      d3 zv = m3_col(rot, max_id);
      if (!max_pos) zv = scl3(-1.0, zv);
      m3_setcol(sw, 2, zv); cs[2] = ev[max_id];
      int xi = (max_id + 1) % 3, yi = (max_id + 2) % 3;
      m3_setcol(sw, 0, m3_col(rot, xi)); cs[0] = ev[xi];
      m3_setcol(sw, 1, cross3(m3_col(sw, 2), m3_col(sw, 0))); cs[1] = ev[yi];
      // ApplyGravityPrior / calibRotMatAccordingToGroundPlane :766-788
      d3 zA = m3_col(sw, 2);
      d3 axis = unit3(cross3(zA, nrm));
      double cos_theta = dot3(nrm, zA) / norm3(nrm) / norm3(zA);
      double theta = acos(cos_theta);
      mat3 Rc = m3_mul(R_axis_angle(theta, axis), sw);
      // normalised frame :327-346
      d3 rx = unit3(m3_col(Rc, 0));
      d3 ry = cross3(zax, rx);
      mat3 Rwo; m3_setcol(Rwo, 0, rx); m3_setcol(Rwo, 1, ry); m3_setcol(Rwo, 2, zax);
      se3 Two; Two.r = q_from_R(Rwo); Two.t = cen;
      se3 Tow = se3_inv(Two);
      se3 Toc = se3_mul(Tow, Twc);
      st.pca_center[0] = cen.x; st.pca_center[1] = cen.y; st.pca_center[2] = cen.z;
      for (int i = 0; i < 9; i++) st.pca_rot[i] = Rc.a[i];
      st.pca_eval[0] = cs[0]; st.pca_eval[1] = cs[1]; st.pca_eval[2] = cs[2];
      se3_store(Two, st.T_wo); se3_store(Toc, st.T_oc);
      proj_matrix(se3_inv(Toc), prm.K, st.P);
      st.n_cluster = nc;
      hom34 h = se3_hom(Tow);
      for (int i = 0; i < 12; i++) S.bcast_d[i] = h.m[i];
    }
    __syncthreads();
  }
  hom34 Tow_h;
  for (int i = 0; i < 12; i++) Tow_h.m[i] = S.bcast_d[i];
  __syncthreads();

  // ---------------- symmetry grid (0.1 m) + transform to the object-normalised frame :322-347 ----------------
  if (tid == 0) st.phase_clk[7] = clock64();
  int M = 0;
  voxel_sort(Cc, nc, prm.sym_leaf, k0, v0, k1, v1, S, RS, s_parent);
  double *objn = W.objn + 3 * (size_t)m.off;
  float4 *objnf = W.objnf + m.off;
  // The 0.1 m voxels hold ~50 points each, and the centroid is a sequential float sum in ascending point order: (1) the first sorted
  // position of every voxel is compacted into v1, (2) one thread per voxel walks its members with eight independent gathers in
  // flight per trip (a head thread chasing v0[j] -> Cc[..] one member at a time across tile boundaries was 19 % of this kernel).
  for (int tile = 0; tile < nc; tile += blockDim.x) {
    const int j = tile + tid;
    const bool hd[1] = {j < nc && (j == 0 || k0[j - 1] != k0[j])};
    int tot, ex[1];
    block_flag_scan<1>(hd, ex, tot, SS, phase);
    if (hd[0]) v1[M + ex[0]] = (unsigned)j;
    M += tot;
  }
  __syncthreads();
  for (int vid = tid; vid < M; vid += blockDim.x) {
    const int j0 = (int)v1[vid], j1 = (vid + 1 < M) ? (int)v1[vid + 1] : nc;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (int jj = j0; jj < j1; jj += 8) {
      float4 q8[8];
#pragma unroll
      for (int u = 0; u < 8; u++) q8[u] = Cc[v0[min(jj + u, j1 - 1)]];
#pragma unroll
      for (int u = 0; u < 8; u++) if (jj + u < j1) { sx = __fadd_rn(sx, q8[u].x); sy = __fadd_rn(sy, q8[u].y); sz = __fadd_rn(sz, q8[u].z); }
    }
    const float cnt = (float)(j1 - j0);
    const float4 c = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), 0.0f);
    const d3 po = hom_apply(Tow_h, mk3((double)c.x, (double)c.y, (double)c.z));
    double *q = objn + 3 * (size_t)vid; q[0] = po.x; q[1] = po.y; q[2] = po.z;
    objnf[vid] = make_float4((float)po.x, (float)po.y, (float)po.z, 0.0f);   // initializeKDTree: double -> float
    if (W.d_sym) { double *qq = W.d_sym + 3 * (size_t)(m.off + vid); qq[0] = po.x; qq[1] = po.y; qq[2] = po.z; }
  }
  if (tid == 0) {
    st.phase_clk[8] = clock64();
    st.n_sym = M;
    st.run_sym = (prm.symmetry_open && m.symtype > 0 && M >= 1) ? 1 : 0;
  }
  // Everything but the object-normalised symmetry cloud (objn / objnf, M points) is dead now.  Its dirty lines would sit in L2 until
  // evicted and then be written back to HBM for nothing (the kernel wrote 1.04 MB per object to DRAM, more than it read): drop them.
  // A slice starts on a 128-byte boundary and holds a multiple of 128 bytes per array (cap is a multiple of 32 elements).
  if (W.l2_discard) {
    __syncthreads();
    const size_t r1 = (((size_t)n1 + 31) & ~(size_t)31), r3 = (((size_t)n3 + 31) & ~(size_t)31), rc = (((size_t)nc + 31) & ~(size_t)31);
    l2_discard_lines(A, r1 * 16, tid, kExtractThreads);
    l2_discard_lines(B, r3 * 16, tid, kExtractThreads);
    l2_discard_lines(Cc, rc * 16, tid, kExtractThreads);
    l2_discard_lines(k0, r1 * 4, tid, kExtractThreads); l2_discard_lines(k1, r1 * 4, tid, kExtractThreads);
    l2_discard_lines(v0, r1 * 4, tid, kExtractThreads); l2_discard_lines(v1, r1 * 4, tid, kExtractThreads);
    l2_discard_lines(W.count + m.off, r3 * 4, tid, kExtractThreads);
    if (n3 > SMEM_PARENTS) l2_discard_lines(W.parent + m.off, r3 * 4, tid, kExtractThreads);
  }
}

// Host entry, pinned (mapped) input: the SMs pull each object's box rectangle straight out of host memory over PCIe (16-byte
// loads, 4 in flight per thread) into the device-side frame buffer the extraction kernels read.  One launch per chunk replaces one
// DMA submission per frame (each costs ~5 us on the copy engine, more than moving a 200 KB rectangle), and only the pixels the path
// can read cross the bus.  Runs on the high-priority copy stream next to the compute kernels of the previous chunks.
// CTA shape: 128 threads x <= 32 registers = 4096 registers, exactly what three resident extract_kernel CTAs (3 x 20 480) leave free on an
// SM -- a 256-thread, 62-register version displaced one extraction CTA on every SM it ran on while the uploads were in flight.
constexpr int kPullThreads = 128;
__global__ void __launch_bounds__(kPullThreads, 16) pull_rects_kernel(const uint16_t *__restrict__ host_src, uint16_t *__restrict__ dst, size_t pitch, size_t frame_stride,
                                                                      const ObjMeta *__restrict__ metas, int n_obj, int rows, int cols) {
  for (int k = blockIdx.x; k < n_obj; k += gridDim.x) {
    const ObjMeta &m = metas[k];
    if (m.pre_status != 0 || m.nx == 0) continue;   // outside the image / empty (or inverted) box: no pixel of it is ever read
    const int x0 = m.x1 & ~7, x1 = min(cols, (m.x2 + 8) & ~7);
    const int y0 = m.y1, y1 = min(rows, m.y2 + 1);
    const int upr = (x1 - x0) >> 3;               // 16-byte units per row
    const int total = upr * (y1 - y0);
    const size_t base = (size_t)m.frame * frame_stride + (size_t)y0 * pitch + x0;
    const uint16_t *src = host_src + base;
    uint16_t *out = dst + base;
    const unsigned p32 = (unsigned)pitch;
    for (int u0 = threadIdx.x; u0 < total; u0 += 4 * kPullThreads) {
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int u = min(u0 + j * kPullThreads, total - 1);
        const int r = u / upr, c = u - r * upr;
        v[j] = __ldcs(reinterpret_cast<const uint4 *>(src + (r * p32 + c * 8u)));
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int u = u0 + j * kPullThreads;
        if (u < total) { const int r = u / upr, c = u - r * upr; *reinterpret_cast<uint4 *>(out + (r * p32 + c * 8u)) = v[j]; }
      }
    }
  }
}
// The same upload driven by the bulk-copy engine (TMA, cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS) instead of LDG/STG: one elected
// thread per one-warp CTA keeps kTmaStages row segments of its objects in flight -- bulk load host (mapped, pinned) -> shared memory
// with an mbarrier transaction count, bulk store shared -> device frame buffer -- so the pull costs no load/store issue slots and
// 352 bytes of registers next to the resident extraction / symmetry CTAs.  Row segments are 16-byte aligned multiples of 16 bytes.
constexpr int kTmaStages = 8;   // upper bound; the launch picks the depth (EslamHandle::pull_stages)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
struct PullCursor {   // (object, row) items of this CTA in a fixed order
  const ObjMeta *metas; int n_obj, rows, cols, k, r, nr, step;
  __device__ void settle() {   // move to the next object that has rows
    while (k < n_obj) {
      const ObjMeta &m = metas[k];
      if (m.pre_status == 0 && m.nx != 0) { nr = min(rows, m.y2 + 1) - m.y1; if (nr > 0) return; }
      k += step; r = 0;
    }
  }
  __device__ bool valid() const { return k < n_obj; }
  __device__ void next() { if (++r >= nr) { k += step; r = 0; settle(); } }
  __device__ void item(size_t pitch, size_t frame_stride, size_t &off, unsigned &bytes) const {
    const ObjMeta &m = metas[k];
    const int x0 = m.x1 & ~7, x1 = min(cols, (m.x2 + 8) & ~7);
    off = (size_t)m.frame * frame_stride + (size_t)(m.y1 + r) * pitch + x0;
    bytes = (unsigned)(x1 - x0) * 2u;
  }
};
__global__ void __launch_bounds__(32, 8) pull_rects_tma_kernel(const uint16_t *__restrict__ host_src, uint16_t *__restrict__ dst, size_t pitch, size_t frame_stride,
                                                               const ObjMeta *__restrict__ metas, int n_obj, int rows, int cols, int stage_bytes, int n_stages) {
  extern __shared__ __align__(128) unsigned char tma_buf[];   // [kTmaStages][stage_bytes]
  __shared__ __align__(8) unsigned long long bar[kTmaStages];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < n_stages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  PullCursor ld{metas, n_obj, rows, cols, (int)blockIdx.x, 0, 0, (int)gridDim.x}, st = ld;
  ld.settle(); st.settle();
  auto issue_load = [&](int stage) {
    size_t off; unsigned bytes;
    ld.item(pitch, frame_stride, off, bytes);
    const unsigned b = smem_u32(&bar[stage]), d = smem_u32(tma_buf + (size_t)stage * stage_bytes);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(host_src + off), "r"(bytes), "r"(b) : "memory");
    ld.next();
  };
  int issued = 0;
  for (; issued < n_stages && ld.valid(); issued++) issue_load(issued);
  for (int i = 0; st.valid(); i++) {
    const int stage = i % n_stages;
    const unsigned parity = (unsigned)(i / n_stages) & 1u, b = smem_u32(&bar[stage]);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(b), "r"(parity) : "memory");
    size_t off; unsigned bytes;
    st.item(pitch, frame_stride, off, bytes);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(smem_u32(tma_buf + (size_t)stage * stage_bytes)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    st.next();
    if (ld.valid()) {   // the stage is reused once the store has read it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue_load(stage);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int pull_rects_launch(EslamHandle *h, cudaStream_t s, const uint16_t *host_mapped, uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj) {
  ExtractWorkspace &w = *h->ex;
  if (n_obj <= 0) return ESLAM_OK;
  if (h->pull_tma) {   // default: bulk-copy engine instead of LDG/STG (+10 % end to end, profiles/r02_pull_tma_ab.txt); ESLAM_PULL_TMA=0 disables
    const int stage_bytes = (((h->cam.cols + 8) * 2) + 127) & ~127;
    const int want = h->pull_ctas_per_sm * h->sm_count, grid = n_obj < want ? n_obj : want;
    const int n_stages = h->pull_stages < 1 ? 1 : (h->pull_stages > kTmaStages ? kTmaStages : h->pull_stages);
    pull_rects_tma_kernel<<<grid, 32, (size_t)n_stages * stage_bytes, s>>>(host_mapped, depth_dev, pitch, frame_stride, w.meta.as<ObjMeta>() + obj0, n_obj, h->cam.rows,
                                                                              h->cam.cols, stage_bytes, n_stages);
    ESLAM_LAUNCHED(h);
    ESLAM_CUDA_OK(h, cudaGetLastError());
    return ESLAM_OK;
  }
  const int grid = n_obj < h->sm_count ? n_obj : h->sm_count;
  pull_rects_kernel<<<grid, kPullThreads, 0, s>>>(host_mapped, depth_dev, pitch, frame_stride, w.meta.as<ObjMeta>() + obj0, n_obj, h->cam.rows, h->cam.cols);
  ESLAM_LAUNCHED(h);
  ESLAM_CUDA_OK(h, cudaGetLastError());
  return ESLAM_OK;
}

int extract_launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj, const ExtractParamsDev &prm,
                   const ExtractPtrs &ptrs) {
  ExtractWorkspace &w = *h->ex;
  if (n_obj <= 0) return ESLAM_OK;
  // Shape: 3 CTAs/SM (80 registers, 48 KB of parents).  Measured alternative: 4 CTAs/SM with a 64-register cap and 32 KB of parents is
  // 9 % slower (profiles/r01_kernel_shape_ab.txt).
  void (*fn)(const uint16_t *, size_t, size_t, const ObjMeta *, ObjState *, ExtractParamsDev, ExtractPtrs) =
      !h->extract_uq ? extract_kernel<3, kSmemParents, false, 1> : (h->extract_pf == 2 ? extract_kernel<3, kSmemParents, true, 2> : extract_kernel<3, kSmemParents, true, 1>);
  if (!h->attr_extract) {
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemParents * (int)sizeof(int)));
    ESLAM_CUDA_OK(h, cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    h->attr_extract = true;
  }
  fn<<<n_obj, kExtractThreads, kSmemParents * sizeof(int), s>>>(depth_dev, pitch, frame_stride, w.meta.as<ObjMeta>() + obj0, w.state.as<ObjState>() + obj0, prm, ptrs);
  ESLAM_LAUNCHED(h);
  ESLAM_CUDA_OK(h, cudaGetLastError());
  return ESLAM_OK;
}

}  // namespace eslam

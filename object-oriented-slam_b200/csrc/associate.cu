// associate.cu -- N4: automatic data association of a BATCH of frames against one map snapshot, on the device.
//
// Replaces DataAssociationSolver::Solve / SolveCostMat (src/core/DataAssociation.cpp:16-135) for the frame-parallel case: every
// frame's single-frame ellipsoids (camera frame) are moved to the world frame (ellipsoid::transform_from(campose_wc) :53), the
// cost of (observation, map object) is the distance of their centres (:60-66), and the rows of a frame are resolved IN ORDER:
// the cheapest column if its cost is below 1 m (SETTING_DIS_THRESH :96), that column then set to 999 for the rows that follow
// (:120-124); otherwise the observation opens a new instance (:127-131).  Quirk kept: a match returns the COLUMN index of the
// instance-ordered map (:113-114), not the instance id.  The reference numbers new instances with a running counter over the
// frames it sees one by one; here the frames of a batch are independent (one warp each), so a second pass numbers the new
// instances in (frame, row) order from `first_new_instance` -- what the sequential call sequence would give on the same snapshot.
#include "common.cuh"
#include "dmath.cuh"
#include "internal.h"

namespace eslam {
namespace {

constexpr int kAssocMaxDet = 64;   // rows of one frame handled in shared memory (more: ESLAM_ERR_CAPACITY)

// one warp per frame; det_start[f .. f+1) = this frame's rows (frame order); out[row] = column index, -1 (no valid 3-D observation)
// or -2 (new instance, numbered by assoc_number_kernel); new_count[f] = number of -2 rows
__global__ void __launch_bounds__(128) assoc_frames_kernel(int n_frames, const int *__restrict__ det_start, const double *__restrict__ Twc7,
                                                           const double *__restrict__ det_pose7, const int *__restrict__ valid, int n_map,
                                                           const double *__restrict__ map_center3, int *__restrict__ out, int *__restrict__ new_count) {
  __shared__ double s_c[4][kAssocMaxDet][3];
  __shared__ unsigned char s_taken[4][1024];   // columns already associated in this frame (n_map <= 1024 in shared memory, else global scan below)
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * 4 + wid;
  if (f >= n_frames) return;
  const int r0 = det_start[f], r1 = det_start[f + 1];
  const se3 Twc = se3_load(Twc7 + 7 * f);
  for (int r = r0 + lane; r < r1; r += 32) {
    const se3 w = se3_mul(Twc, se3_load(det_pose7 + 7 * (size_t)r));   // transform_from: pose = Twc * pose
    s_c[wid][r - r0][0] = w.t.x; s_c[wid][r - r0][1] = w.t.y; s_c[wid][r - r0][2] = w.t.z;
  }
  for (int j = lane; j < n_map && j < 1024; j += 32) s_taken[wid][j] = 0;
  __syncwarp();
  int created = 0;
  for (int r = r0; r < r1; r++) {   // rows in order: an earlier row takes its column away from the later ones
    if (!valid[r]) { if (lane == 0) out[r] = -1; continue; }
    const double cx = s_c[wid][r - r0][0], cy = s_c[wid][r - r0][1], cz = s_c[wid][r - r0][2];
    double best = INFINITY; int bj = 0x7fffffff;
    for (int j = lane; j < n_map; j += 32) {
      const double dx = map_center3[3 * j] - cx, dy = map_center3[3 * j + 1] - cy, dz = map_center3[3 * j + 2] - cz;
      double c = sqrt(dx * dx + dy * dy + dz * dz);       // (pObj->pose.translation() - observed.translation()).norm()
      if (s_taken[wid][j]) c = 999.0;
      if (c < best) { best = c; bj = j; }                 // per lane: ascending j, strict < keeps the first minimum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                    // minCoeff: first index of the minimum
      const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    int res;
    if (n_map >= 1 && best < 1.0) { res = bj; if (lane == 0) s_taken[wid][bj] = 1; }
    else { res = -2; created++; }
    if (lane == 0) out[r] = res;
    __syncwarp();
  }
  if (lane == 0) new_count[f] = created;
}

// numbers the new instances in (frame, row) order: one CTA, exclusive scan of new_count over the frames
__global__ void __launch_bounds__(256) assoc_number_kernel(int n_frames, const int *__restrict__ det_start, const int *__restrict__ new_count, int first_new_instance,
                                                           int *__restrict__ out, int *__restrict__ total_new) {
  __shared__ int s_part[256];
  const int tid = threadIdx.x;
  const int per = (n_frames + 255) / 256, f0 = tid * per, f1 = min(n_frames, f0 + per);
  int sum = 0;
  for (int f = f0; f < f1; f++) sum += new_count[f];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) { int run = 0; for (int k = 0; k < 256; k++) { const int v = s_part[k]; s_part[k] = run; run += v; } *total_new = run; }
  __syncthreads();
  int next = first_new_instance + s_part[tid];
  for (int f = f0; f < f1; f++)
    for (int r = det_start[f]; r < det_start[f + 1]; r++)
      if (out[r] == -2) out[r] = next++;
}

}  // namespace

int associate_frames(EslamHandle *h, cudaStream_t s, int n_frames, const int *det_start_host, const double *Twc7, int n_det, const double *det_pose7, const int *valid,
                     int n_map, const double *map_center3, int first_new_instance, int *assoc_out, int *n_new_out) {
  if (n_frames <= 0 || n_det < 0 || n_map < 0 || n_map > 1024) return n_map > 1024 ? ESLAM_ERR_CAPACITY : ESLAM_ERR_BAD_ARG;
  for (int f = 0; f < n_frames; f++) {
    const int n = det_start_host[f + 1] - det_start_host[f];
    if (n < 0) return ESLAM_ERR_BAD_ARG;
    if (n > kAssocMaxDet) return ESLAM_ERR_CAPACITY;
  }
  if (det_start_host[0] != 0 || det_start_host[n_frames] != n_det) return ESLAM_ERR_BAD_ARG;
  DevBuf &buf = h->assoc_buf;
  const size_t o_start = 0, o_twc = o_start + (((size_t)(n_frames + 1) * 4 + 255) & ~(size_t)255), o_pose = o_twc + (size_t)n_frames * 56,
               o_map = o_pose + (size_t)(n_det > 0 ? n_det : 1) * 56, o_valid = o_map + (size_t)(n_map > 0 ? n_map : 1) * 24,
               o_out = o_valid + (((size_t)(n_det > 0 ? n_det : 1) * 4 + 7) & ~(size_t)7), o_cnt = o_out + (((size_t)(n_det > 0 ? n_det : 1) * 4 + 7) & ~(size_t)7),
               total = o_cnt + (size_t)(n_frames + 1) * 4;
  ESLAM_CUDA_OK(h, buf.reserve(total));
  char *b = buf.as<char>();
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(b + o_start, det_start_host, (size_t)(n_frames + 1) * 4, cudaMemcpyHostToDevice, s));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(b + o_twc, Twc7, (size_t)n_frames * 56, cudaMemcpyHostToDevice, s));
  if (n_det > 0) {
    ESLAM_CUDA_OK(h, cudaMemcpyAsync(b + o_pose, det_pose7, (size_t)n_det * 56, cudaMemcpyHostToDevice, s));
    ESLAM_CUDA_OK(h, cudaMemcpyAsync(b + o_valid, valid, (size_t)n_det * 4, cudaMemcpyHostToDevice, s));
  }
  if (n_map > 0) ESLAM_CUDA_OK(h, cudaMemcpyAsync(b + o_map, map_center3, (size_t)n_map * 24, cudaMemcpyHostToDevice, s));
  int *d_start = reinterpret_cast<int *>(b + o_start), *d_out = reinterpret_cast<int *>(b + o_out), *d_cnt = reinterpret_cast<int *>(b + o_cnt);
  assoc_frames_kernel<<<(n_frames + 3) / 4, 128, 0, s>>>(n_frames, d_start, reinterpret_cast<const double *>(b + o_twc), reinterpret_cast<const double *>(b + o_pose),
                                                         reinterpret_cast<const int *>(b + o_valid), n_map, reinterpret_cast<const double *>(b + o_map), d_out, d_cnt);
  ESLAM_LAUNCHED(h);
  assoc_number_kernel<<<1, 256, 0, s>>>(n_frames, d_start, d_cnt, first_new_instance, d_out, d_cnt + n_frames);
  ESLAM_LAUNCHED(h);
  ESLAM_CUDA_OK(h, cudaGetLastError());
  if (n_det > 0) ESLAM_CUDA_OK(h, cudaMemcpyAsync(assoc_out, d_out, (size_t)n_det * 4, cudaMemcpyDeviceToHost, s));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(n_new_out, d_cnt + n_frames, 4, cudaMemcpyDeviceToHost, s));
  ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));
  return ESLAM_OK;
}

}  // namespace eslam

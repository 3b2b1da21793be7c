// init_quadric.cu -- closed-form dual-quadric initialisation from bounding-box tangent planes, one warp per object.
//
// Replaces Initializer::initializeQuadric and helpers (src/core/Initializer.cpp:24-56, :58-91 getPlanesHomo,
// :107-145 fromDetectionsToLines, :147-164 getVectorFromPlanesHomo, :166-184 getQStarFromVectors (JacobiSVD),
// :186-248 getEllipsoidFromQStar).  The thin-V null vector of the (<=4F) x 10 system is found with a one-sided
// (Hestenes) Jacobi SVD: rows are spread over the 32 lanes, column dot products are warp-shuffle reductions.
#include "common.cuh"
#include "dmath.cuh"
#include "internal.h"

namespace eslam {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ bool inverse4_det(const double *A, double *inv, double *det_out) {  // Gauss-Jordan, partial pivoting
  double w[4][8];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { w[i][j] = A[i * 4 + j]; w[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  double det = 1.0;
  for (int c = 0; c < 4; c++) {
    int p = c; double best = fabs(w[c][c]);
    for (int r = c + 1; r < 4; r++) if (fabs(w[r][c]) > best) { best = fabs(w[r][c]); p = r; }
    if (best == 0.0) { *det_out = 0; return false; }
    if (p != c) { for (int j = 0; j < 8; j++) { double t = w[p][j]; w[p][j] = w[c][j]; w[c][j] = t; } det = -det; }
    double piv = w[c][c]; det *= piv;
    for (int j = 0; j < 8; j++) w[c][j] /= piv;
    for (int r = 0; r < 4; r++) if (r != c) { double f = w[r][c]; if (f != 0.0) for (int j = 0; j < 8; j++) w[r][j] -= f * w[c][j]; }
  }
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) inv[i * 4 + j] = w[i][4 + j];
  *det_out = det;
  return true;
}
__device__ void jacobi_eig4(double *A, double *ev) {  // eigenvalues only, ascending
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) off += A[i * 4 + j] * A[i * 4 + j];
    if (off < 1e-300) break;
    for (int p = 0; p < 4; p++) for (int q = p + 1; q < 4; q++) {
      double apq = A[p * 4 + q];
      if (apq == 0.0) continue;
      double tau = (A[q * 4 + q] - A[p * 4 + p]) / (2.0 * apq);
      double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
      for (int k = 0; k < 4; k++) { double akp = A[k * 4 + p], akq = A[k * 4 + q]; A[k * 4 + p] = c * akp - s * akq; A[k * 4 + q] = s * akp + c * akq; }
      for (int k = 0; k < 4; k++) { double apk = A[p * 4 + k], aqk = A[q * 4 + k]; A[p * 4 + k] = c * apk - s * aqk; A[q * 4 + k] = s * apk + c * aqk; }
    }
  }
  for (int i = 0; i < 4; i++) ev[i] = A[i * 4 + i];
  for (int i = 1; i < 4; i++) { double v = ev[i]; int j = i - 1; while (j >= 0 && ev[j] > v) { ev[j + 1] = ev[j]; j--; } ev[j + 1] = v; }
}

// one warp; A: [N x 10] row-major scratch in global memory (destroyed)
__global__ void init_quadric_kernel(int n_obs, const double *__restrict__ Twc7, const double *__restrict__ det5, const double *__restrict__ Kc, int rows, int cols,
                                    double *__restrict__ A, double *__restrict__ out10, int *__restrict__ ok_out) {
  const int lane = threadIdx.x & 31;
  __shared__ int s_n;
  __shared__ double s_V[100];
  double K[9];
  for (int i = 0; i < 9; i++) K[i] = Kc[i];
  if (lane == 0) {  // getPlanesHomo + getVectorFromPlanesHomo (sequential: keeps the reference's row order)
    int N = 0;
    for (int i = 0; i < n_obs; i++) {
      const double *det = det5 + 5 * i;
      if (det[0] < 1 && det[1] < 1 && det[2] < 1 && det[3] < 1) continue;
      se3 wc = se3_load(Twc7 + 7 * i);
      se3_normalize(wc);                       // SE3Quat(const Vector7d&) normalises  (se3quat.h:63-66)
      double P[12];
      proj_matrix(se3_inv(wc), K, P);
      double x1 = det[0], y1 = det[1], x2 = det[2], y2 = det[3];
      double lines[4][3] = {{1, 0, -x1}, {0, 1, -y1}, {1, 0, -x2}, {0, 1, -y2}};
      bool keep[4] = {x1 > 0 && x1 < cols - 1, y1 > 0 && y1 < rows - 1, x2 > 0 && x2 < cols - 1, y2 > 0 && y2 < rows - 1};
      for (int l = 0; l < 4; l++) if (keep[l]) {
        double p[4];
        for (int r = 0; r < 4; r++) { double s = P[r] * lines[l][0]; s += P[4 + r] * lines[l][1]; s += P[8 + r] * lines[l][2]; p[r] = s; }
        double *v = A + (size_t)N * 10;
        v[0] = p[0] * p[0]; v[1] = 2 * p[0] * p[1]; v[2] = 2 * p[0] * p[2]; v[3] = 2 * p[0] * p[3]; v[4] = p[1] * p[1];
        v[5] = 2 * p[1] * p[2]; v[6] = 2 * p[1] * p[3]; v[7] = p[2] * p[2]; v[8] = 2 * p[2] * p[3]; v[9] = p[3] * p[3];
        N++;
      }
    }
    s_n = N;
    for (int i = 0; i < 100; i++) s_V[i] = (i % 11 == 0) ? 1.0 : 0.0;
  }
  __syncwarp();
  const int N = s_n;
  if (N < 9) { if (lane == 0) { *ok_out = 0; for (int i = 0; i < 10; i++) out10[i] = 0; } return; }
  // one-sided Jacobi SVD
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (int p = 0; p < 10; p++) for (int q = p + 1; q < 10; q++) {
      double alpha = 0, beta = 0, gamma = 0;
      for (int k = lane; k < N; k += 32) { double ap = A[(size_t)k * 10 + p], aq = A[(size_t)k * 10 + q]; alpha += ap * ap; beta += aq * aq; gamma += ap * aq; }
      alpha = warp_sum(alpha); beta = warp_sum(beta); gamma = warp_sum(gamma);
      if (fabs(gamma) <= 1e-15 * sqrt(alpha * beta) || gamma == 0.0) continue;
      rotated = true;
      double zeta = (beta - alpha) / (2.0 * gamma);
      double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
      for (int k = lane; k < N; k += 32) { double ap = A[(size_t)k * 10 + p], aq = A[(size_t)k * 10 + q]; A[(size_t)k * 10 + p] = c * ap - s * aq; A[(size_t)k * 10 + q] = s * ap + c * aq; }
      if (lane < 10) { double vp = s_V[lane * 10 + p], vq = s_V[lane * 10 + q]; s_V[lane * 10 + p] = c * vp - s * vq; s_V[lane * 10 + q] = s * vp + c * vq; }
      __syncwarp();
    }
    if (!rotated) break;
  }
  int best = 0; double bestn = DBL_MAX;
  for (int j = 0; j < 10; j++) {
    double s2 = 0;
    for (int k = lane; k < N; k += 32) s2 += A[(size_t)k * 10 + j] * A[(size_t)k * 10 + j];
    s2 = warp_sum(s2);
    if (s2 < bestn) { bestn = s2; best = j; }
  }
  if (lane != 0) return;
  double q[10];
  for (int k = 0; k < 10; k++) q[k] = s_V[k * 10 + best];
  double QStar[16] = {q[0], q[1], q[2], q[3], q[1], q[4], q[5], q[6], q[2], q[5], q[7], q[8], q[3], q[6], q[8], q[9]};
  // getEllipsoidFromQStar
  double Qi[16], detQs;
  int ok = 1;
  for (int i = 0; i < 10; i++) out10[i] = 0;
  if (!inverse4_det(QStar, Qi, &detQs)) ok = 0;
  if (ok) {
    double cb = cbrt(detQs);
    double Q[16], Qs[16], ev[4];
    for (int i = 0; i < 16; i++) Q[i] = Qi[i] * cb;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Qs[i * 4 + j] = 0.5 * (Q[i * 4 + j] + Q[j * 4 + i]);
    jacobi_eig4(Qs, ev);
    int num_pos = int(ev[0] > 0) + int(ev[1] > 0) + int(ev[2] > 0) + int(ev[3] > 0);
    int num_neg = int(ev[0] < 0) + int(ev[1] < 0) + int(ev[2] < 0) + int(ev[3] < 0);
    if (!(num_pos == 3 && num_neg == 1) && !(num_pos == 1 && num_neg == 3)) ok = 0;
    if (ok) {
      if (ev[3] > 0) {
        for (int i = 0; i < 16; i++) Q[i] = -Q[i];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Qs[i * 4 + j] = 0.5 * (Q[i * 4 + j] + Q[j * 4 + i]);
        jacobi_eig4(Qs, ev);
      }
      double lambda[3] = {1.0 / ev[0], 1.0 / ev[1], 1.0 / ev[2]};
      mat3 Q33;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q33.a[i * 3 + j] = Q[i * 4 + j];
      double tmp[16], detQ;
      inverse4_det(Q, tmp, &detQ);
      double det33 = Q33.a[0] * (Q33.a[4] * Q33.a[8] - Q33.a[5] * Q33.a[7]) - Q33.a[1] * (Q33.a[3] * Q33.a[8] - Q33.a[5] * Q33.a[6]) +
                     Q33.a[2] * (Q33.a[3] * Q33.a[7] - Q33.a[4] * Q33.a[6]);
      double k = detQ / det33;
      double sc[3];
      for (int i = 0; i < 3; i++) sc[i] = sqrt(fabs(-k * lambda[i]));
      double t[3] = {QStar[3] / QStar[15], QStar[7] / QStar[15], QStar[11] / QStar[15]};
      double e3[3]; mat3 rot;
      eig3_sym(Q33, e3, rot);
      // rot_to_euler_zyx (matrix_utils.cpp:75-95)
      double roll, pitch, yaw;
      pitch = asin(-rot.a[6]);
      if (fabs(pitch - M_PI / 2.0) < 1.0e-3) { roll = 0.0; yaw = atan2(rot.a[5] - rot.a[1], rot.a[2] + rot.a[4]) + roll; }
      else if (fabs(pitch + M_PI / 2.0) < 1.0e-3) { roll = 0.0; yaw = atan2(rot.a[5] - rot.a[1], rot.a[2] + rot.a[4]) - roll; }
      else { roll = atan2(rot.a[7], rot.a[8]); yaw = atan2(rot.a[3], rot.a[0]); }
      // fromMinimalVector (Ellipsoid.cpp:16-22): zyx_euler_to_quat + SE3Quat(q,t) normalisation
      double sy = sin(yaw * 0.5), cy = cos(yaw * 0.5), sp = sin(pitch * 0.5), cp = cos(pitch * 0.5), sr = sin(roll * 0.5), cr = cos(roll * 0.5);
      se3 T;
      T.r.w = cr * cp * cy + sr * sp * sy; T.r.x = sr * cp * cy - cr * sp * sy; T.r.y = cr * sp * cy + sr * cp * sy; T.r.z = cr * cp * sy - sr * sp * cy;
      T.t = mk3(t[0], t[1], t[2]);
      se3_normalize(T);
      se3_store(T, out10);
      out10[7] = sc[0]; out10[8] = sc[1]; out10[9] = sc[2];
    }
  }
  *ok_out = ok;
}

int init_quadric_device(EslamHandle *h, cudaStream_t s, int n_obs, const double *Twc7, const double *det5, const double K[9], int rows, int cols,
                        double ellipsoid10[10], int *ok) {
  DevBuf in, scratch, out;
  size_t nin = (size_t)n_obs * 12 + 9;
  ESLAM_CUDA_OK(h, in.reserve(nin * 8));
  ESLAM_CUDA_OK(h, scratch.reserve((size_t)n_obs * 4 * 10 * 8 + 80));
  ESLAM_CUDA_OK(h, out.reserve(11 * 8));
  std::vector<double> host(nin);
  memcpy(host.data(), Twc7, sizeof(double) * 7 * n_obs);
  memcpy(host.data() + 7 * (size_t)n_obs, det5, sizeof(double) * 5 * n_obs);
  memcpy(host.data() + 12 * (size_t)n_obs, K, sizeof(double) * 9);
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(in.p, host.data(), nin * 8, cudaMemcpyHostToDevice, s));
  double *d = in.as<double>();
  init_quadric_kernel<<<1, 32, 0, s>>>(n_obs, d, d + 7 * (size_t)n_obs, d + 12 * (size_t)n_obs, rows, cols, scratch.as<double>(), out.as<double>(),
                                       reinterpret_cast<int *>(out.as<double>() + 10));
  ESLAM_LAUNCHED(h);
  ESLAM_CUDA_OK(h, cudaGetLastError());
  double res[11];
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(res, out.p, sizeof(res), cudaMemcpyDeviceToHost, s));
  ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));
  for (int i = 0; i < 10; i++) ellipsoid10[i] = res[i];
  *ok = *reinterpret_cast<int *>(&res[10]);
  in.release(); scratch.release(); out.release();
  return ESLAM_OK;
}

}  // namespace eslam

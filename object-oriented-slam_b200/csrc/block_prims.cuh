// block_prims.cuh -- CTA-wide primitives used by the per-object extraction kernels: deterministic reductions,
// ordered (stable) stream compaction, and a stable LSD radix sort of (key, value) pairs living in global memory.
// One CTA owns one object, so everything here synchronises with __syncthreads() only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eslam {

constexpr int kMaxWarps = 32;

struct BlockScratch {            // shared-memory scratch, one per CTA
  int warp_i[kMaxWarps + 1];
  float warp_f[kMaxWarps * 6];
  double warp_d[kMaxWarps * 9];
  int bcast_i[8];
  double bcast_d[16];
};

constexpr int kRadixWarps = 16;   // block_radix_sort supports CTAs of up to 512 threads
struct RadixScratch {
  unsigned int radix[kRadixWarps * 256];   // per-warp digit counters (8-bit digits), or 8 warps x 1024 16-bit counters (wide digits)
  unsigned int digit_base[1024];
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ int n_warps() { return (blockDim.x + 31) >> 5; }

// exclusive prefix sum of one int per thread (thread order); returns exclusive prefix, total through `total`
__device__ __forceinline__ int block_excl_scan(int v, BlockScratch &S, int &total) {
  int lane = lane_id(), w = warp_id();
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) S.warp_i[w] = inc;
  __syncthreads();
  if (w == 0) {
    int nw = n_warps();
    int x = (lane < nw) ? S.warp_i[lane] : 0;
    int xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
    if (lane < nw) S.warp_i[lane] = xi - x;
    if (lane == 31) S.warp_i[kMaxWarps] = xi;
  }
  __syncthreads();
  int res = S.warp_i[w] + inc - v;
  total = S.warp_i[kMaxWarps];
  __syncthreads();
  return res;
}

// ---- one-barrier ordered compaction scans for CTAs of exactly 256 threads (8 warps) ----
// Flags are ranked with ballots; the per-(item, warp) counts go through a double-buffered shared table, so one __syncthreads()
// per call suffices (call k+2 reuses the buffer of call k only after every warp has passed the barrier of call k+1).
// `phase` is a per-thread register that all threads of the CTA flip together.  Every thread of the CTA must call.
struct ScanScratch { int cnt[2][32]; };

// ITEMS flags per thread in item-major order: item k of thread t has rank position k*256 + t.  ITEMS <= 4.
template <int ITEMS>
__device__ __forceinline__ void block_flag_scan(const bool (&f)[ITEMS], int (&ex)[ITEMS], int &total, ScanScratch &S, int &phase) {
  static_assert(ITEMS >= 1 && ITEMS <= 4, "8 warps x ITEMS counters must fit one warp-wide scan");
  const int lane = lane_id(), w = warp_id();
  unsigned bal[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    bal[k] = __ballot_sync(0xffffffffu, f[k]);
    if (lane == k) S.cnt[phase][k * 8 + w] = __popc(bal[k]);
  }
  __syncthreads();
  const int c = (lane < ITEMS * 8) ? S.cnt[phase][lane] : 0;
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  total = __shfl_sync(0xffffffffu, inc, 31);
  const int excl = inc - c;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < ITEMS; k++) ex[k] = __shfl_sync(0xffffffffu, excl, k * 8 + w) + __popc(bal[k] & lt);
  phase ^= 1;
}
// two independent flag streams (a, b) ranked separately in thread order
__device__ __forceinline__ void block_flag_scan2(bool a, bool b, int &exa, int &exb, int &tota, int &totb, ScanScratch &S, int &phase) {
  const int lane = lane_id(), w = warp_id();
  const unsigned ba = __ballot_sync(0xffffffffu, a), bb = __ballot_sync(0xffffffffu, b);
  if (lane == 0) S.cnt[phase][w] = __popc(ba);
  if (lane == 1) S.cnt[phase][8 + w] = __popc(bb);
  __syncthreads();
  const int c = (lane < 16) ? S.cnt[phase][lane] : 0;
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  tota = __shfl_sync(0xffffffffu, inc, 7);
  totb = __shfl_sync(0xffffffffu, inc, 15) - tota;
  const int excl = inc - c;
  const unsigned lt = (1u << lane) - 1u;
  exa = __shfl_sync(0xffffffffu, excl, w) + __popc(ba & lt);
  exb = __shfl_sync(0xffffffffu, excl, 8 + w) - tota + __popc(bb & lt);
  phase ^= 1;
}

__device__ __forceinline__ int block_sum_int(int v, BlockScratch &S) { int tot; block_excl_scan(v, S, tot); return tot; }

// K-wide deterministic fp64 sum (fixed butterfly + fixed warp order): every thread gets the result in out[]
template <int K>
__device__ __forceinline__ void block_sum_double(double *v, BlockScratch &S) {
  int lane = lane_id(), w = warp_id(), nw = n_warps();
#pragma unroll
  for (int k = 0; k < K; k++) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) S.warp_d[w * K + k] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    double acc = 0;
    for (int i = 0; i < nw; i++) acc += S.warp_d[i * K + k];
    v[k] = acc;
  }
  __syncthreads();
}

// min/max of 3 floats (order independent => exact)
__device__ __forceinline__ void block_minmax3(float *mn, float *mx, BlockScratch &S) {
  int lane = lane_id(), w = warp_id(), nw = n_warps();
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float a = mn[k], b = mx[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o)); }
    if (lane == 0) { S.warp_f[w * 6 + k] = a; S.warp_f[w * 6 + 3 + k] = b; }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float a = S.warp_f[k], b = S.warp_f[3 + k];
    for (int i = 1; i < nw; i++) { a = fminf(a, S.warp_f[i * 6 + k]); b = fmaxf(b, S.warp_f[i * 6 + 3 + k]); }
    mn[k] = a; mx[k] = b;
  }
  __syncthreads();
}
__device__ __forceinline__ double block_min_double(double v, BlockScratch &S) {
  int lane = lane_id(), w = warp_id(), nw = n_warps();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) S.warp_d[w] = v;
  __syncthreads();
  double r = S.warp_d[0];
  for (int i = 1; i < nw; i++) r = fmin(r, S.warp_d[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ unsigned block_max_u32(unsigned v, BlockScratch &S) {
  int lane = lane_id(), w = warp_id(), nw = n_warps();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) S.warp_i[w] = (int)v;
  __syncthreads();
  unsigned r = (unsigned)S.warp_i[0];
  for (int i = 1; i < nw; i++) r = max(r, (unsigned)S.warp_i[i]);
  __syncthreads();
  return r;
}

// Two-pass variant for 256-thread CTAs: keys of up to 20 bits are sorted with two digits of ceil(bits/2) <= 10 bits instead of three
// 8-bit passes.  The per-warp counters are 16 bit (8 warps x 1024 digits in the same 16 KB), which bounds n at 65 535.  Same stable
// ranking as the 8-bit version (match_any inside a 32-key group, warp-major offsets across the CTA).
// cnt16: NW x dstride 16-bit counters, dbase: 2^ceil(bits/2) words -- both in shared memory (RadixScratch for <= 10-bit digits, the idle
// 48 KB union-find region for 11-bit digits).
static __device__ void block_radix_sort_wide(unsigned *k0, unsigned *v0, unsigned *k1, unsigned *v1, int n, int bits, unsigned short *cnt16, unsigned *dbase, int dstride,
                                             BlockScratch &BS) {
  const int lane = lane_id(), w = warp_id();
  constexpr int NW = 8;
  const int passes = (bits + 10) / 11 < 2 ? 2 : (bits + 10) / 11;   // two passes up to 22 bits, three up to 33
  const int dbits = (bits + passes - 1) / passes;
  const int D = 1 << dbits;
  const unsigned dmask = (unsigned)D - 1u;
  unsigned *cntw = reinterpret_cast<unsigned *>(cnt16);
  const int per = (D + 255) >> 8;              // digits per thread in the offset scan (<= 8)
  int chunk = (n + NW - 1) / NW;
  chunk = (chunk + 31) & ~31;
  const int beg = min(n, w * chunk), end = min(n, beg + chunk);
  unsigned *ki = k0, *vi = v0, *ko = k1, *vo = v1;
  for (int p = 0; p < passes; p++) {
    const int shift = p * dbits;
    for (int i = threadIdx.x; i < NW * dstride / 2; i += blockDim.x) cntw[i] = 0;
    __syncthreads();
    unsigned short *cnt = cnt16 + w * dstride;
    for (int i0 = beg; i0 < end; i0 += 128) {
      unsigned kk[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int j = i0 + u * 32 + lane; kk[u] = (j < end) ? ki[j] : 0u; }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = i0 + u * 32 + lane;
        if (i0 + u * 32 < end) {   // warp-uniform
          const bool valid = j < end;
          const unsigned d = valid ? ((kk[u] >> shift) & dmask) : (0x10000u + lane);
          const unsigned peers = __match_any_sync(0xffffffffu, d);
          if (valid && (__ffs(peers) - 1) == lane) cnt[d] = (unsigned short)(cnt[d] + __popc(peers));
          __syncwarp();
        }
      }
    }
    __syncthreads();
    // (digit major, warp minor) exclusive offsets: thread t owns digits t*per .. t*per+per-1
    unsigned tot8[8]; unsigned mine = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int d = threadIdx.x * per + q;
      unsigned run = 0;
      if (q < per && d < D) for (int ww = 0; ww < NW; ww++) { const unsigned c = cnt16[ww * dstride + d]; cnt16[ww * dstride + d] = (unsigned short)run; run += c; }
      tot8[q] = run; mine += run;
    }
    int total;
    unsigned ex = (unsigned)block_excl_scan((int)mine, BS, total);
#pragma unroll
    for (int q = 0; q < 8; q++) { const int d = threadIdx.x * per + q; if (q < per && d < D) dbase[d] = ex; ex += tot8[q]; }
    __syncthreads();
    for (int i0 = beg; i0 < end; i0 += 128) {
      unsigned kk[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int j = i0 + u * 32 + lane; const bool valid = j < end; kk[u] = valid ? ki[j] : 0u; vv[u] = valid ? vi[j] : 0u; }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = i0 + u * 32 + lane;
        if (i0 + u * 32 < end) {   // warp-uniform
          const bool valid = j < end;
          const unsigned key = kk[u], val = vv[u];
          const unsigned d = valid ? ((key >> shift) & dmask) : (0x10000u + lane);
          const unsigned peers = __match_any_sync(0xffffffffu, d);
          const unsigned base = valid ? ((unsigned)cnt[d] + dbase[d]) : 0u;
          const unsigned rank = __popc(peers & ((1u << lane) - 1u));
          __syncwarp();
          if (valid && (__ffs(peers) - 1) == lane) cnt[d] = (unsigned short)(cnt[d] + __popc(peers));
          __syncwarp();
          if (valid) { ko[base + rank] = key; vo[base + rank] = val; }
        }
      }
    }
    __syncthreads();
    unsigned *t;
    t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  if (ki != k0) {  // odd number of passes: copy back
    for (int i = threadIdx.x; i < n; i += blockDim.x) { k0[i] = ki[i]; v0[i] = vi[i]; }
    __syncthreads();
  }
}

// Stable LSD radix sort of n (key,val) pairs, 8-bit digits, keys < 2^bits.  Warp w owns the contiguous chunk
// [w*chunk, (w+1)*chunk); ranks inside a chunk come from __match_any_sync so equal keys keep their input order.
// Result ends in (k0,v0).  k1/v1 are scratch of the same size.
static __device__ void block_radix_sort(unsigned *k0, unsigned *v0, unsigned *k1, unsigned *v1, int n, int bits, RadixScratch &S) {
  const int lane = lane_id(), w = warp_id(), nw = n_warps();
  int chunk = (n + nw - 1) / nw;
  chunk = (chunk + 31) & ~31;
  const int beg = min(n, w * chunk), end = min(n, beg + chunk);
  unsigned *ki = k0, *vi = v0, *ko = k1, *vo = v1;
  int passes = (bits + 7) / 8;
  if (passes < 1) passes = 1;
  for (int p = 0; p < passes; p++) {
    const int shift = p * 8;
    for (int i = threadIdx.x; i < nw * 256; i += blockDim.x) S.radix[i] = 0;
    __syncthreads();
    unsigned *cnt = &S.radix[w * 256];
    // both loops run four 32-key groups per trip: the four (independent) global loads are issued first, so a warp pays one memory
    // latency per 128 keys instead of one per 32 (the loops were latency bound: 15 % of the extraction kernel)
    for (int i0 = beg; i0 < end; i0 += 128) {
      unsigned kk[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int j = i0 + u * 32 + lane; kk[u] = (j < end) ? ki[j] : 0u; }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = i0 + u * 32 + lane;
        if (i0 + u * 32 < end) {   // warp-uniform
          const bool valid = j < end;
          const unsigned d = valid ? ((kk[u] >> shift) & 255u) : (256u + lane);
          const unsigned peers = __match_any_sync(0xffffffffu, d);
          if (valid && (__ffs(peers) - 1) == lane) cnt[d] += __popc(peers);
          __syncwarp();
        }
      }
    }
    __syncthreads();
    // exclusive scan over (digit major, warp minor)
    if (threadIdx.x < 256) {
      unsigned run = 0;
      for (int ww = 0; ww < nw; ww++) { unsigned c = S.radix[ww * 256 + threadIdx.x]; S.radix[ww * 256 + threadIdx.x] = run; run += c; }
      S.digit_base[threadIdx.x] = run;
    }
    __syncthreads();
    if (w == 0) {  // exclusive scan of the 256 digit totals
      unsigned carry = 0;
      for (int b = 0; b < 256; b += 32) {
        unsigned x = S.digit_base[b + lane];
        unsigned inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        S.digit_base[b + lane] = carry + inc - x;
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    __syncthreads();
    for (int i0 = beg; i0 < end; i0 += 128) {
      unsigned kk[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int j = i0 + u * 32 + lane; const bool valid = j < end; kk[u] = valid ? ki[j] : 0u; vv[u] = valid ? vi[j] : 0u; }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = i0 + u * 32 + lane;
        if (i0 + u * 32 < end) {   // warp-uniform
          const bool valid = j < end;
          const unsigned key = kk[u], val = vv[u];
          const unsigned d = valid ? ((key >> shift) & 255u) : (256u + lane);
          const unsigned peers = __match_any_sync(0xffffffffu, d);
          const unsigned base = valid ? (cnt[d] + S.digit_base[d]) : 0u;
          const unsigned rank = __popc(peers & ((1u << lane) - 1u));
          __syncwarp();
          if (valid && (__ffs(peers) - 1) == lane) cnt[d] += __popc(peers);
          __syncwarp();
          if (valid) { ko[base + rank] = key; vo[base + rank] = val; }
        }
      }
    }
    __syncthreads();
    unsigned *t;
    t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  if (ki != k0) {  // odd number of passes: copy back
    for (int i = threadIdx.x; i < n; i += blockDim.x) { k0[i] = ki[i]; v0[i] = vi[i]; }
    __syncthreads();
  }
}

__device__ __forceinline__ int bits_for(unsigned maxval) { return maxval == 0 ? 1 : (32 - __clz(maxval)); }

}  // namespace eslam

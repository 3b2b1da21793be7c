// common.cuh -- handle, device-buffer helpers and launch accounting shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/eslam.h"

namespace eslam {

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  // grow-only device buffer
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct PinBuf {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; bytes = 0;
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

constexpr int kMaxChunks = 16;
constexpr int kMaxSmId = 256;   // upper bound of %smid + 1 the symmetry cache tables are sized for
constexpr int kSymCtas256 = 3, kSymCtas512 = 2;    // launch bounds (resident nine-warp CTAs per SM) of the two symmetry instantiations
constexpr int kSymSlots256 = 4, kSymSlots512 = 2;  // cache slots per SM: what shared memory admits at most (~50 KB / ~87 KB per CTA)
struct ExtractWorkspace;  // extract.cu
struct GraphWorkspace;    // edges.cu
struct SlamWorkspace;     // slam.cu
struct PlaneWorkspace;    // plane.cu
struct CopyPool;          // capi.cu: host threads that stage pageable input into pinned memory

}  // namespace eslam

struct EslamHandle {
  EslamParams prm;
  EslamCamera cam;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;  // owned stream used by the *_host entry points
  cudaStream_t copy_stream = nullptr;  // H2D stream of the pipelined *_host extraction
  cudaEvent_t chunk_ev[17] = {};       // [16]: metadata upload of the batch
  cudaStream_t chunk_stream[16] = {};   // per-chunk compute streams: chunks overlap each other and the remaining uploads
  cudaEvent_t chunk_done[16] = {};
  cudaStream_t chunk_aux[16] = {};      // round pipeline: the single-launch kernel of a chunk's large clouds (n_sym > 256) runs beside the rounds
  cudaEvent_t aux_fork[16] = {}, aux_join[16] = {};
  int h2d_mode = 4;      // upload: 4 = SMs pull each box rectangle from pinned host memory (falls back to 3 for pageable input);
                         // 3 = union rectangle of each frame's boxes, one batched DMA call per chunk; 2 = same, one call per frame;
                         // 1 = the boxes' rows (contiguous), 0 = whole frames (env ESLAM_H2D_MODE)
  int host_chunks = 8;   // pipeline depth of the host entry (env ESLAM_HOST_CHUNKS, 1..16)
  int dev_chunks = 8;    // object chunks / streams of the device-resident entry for batches of >= 512 objects (env ESLAM_DEV_CHUNKS)
  // symmetry kernel caches: one slot per RESIDENT seed warp (sm_count x 32 CTA slots), handed out through a per-SM bitmask, so the
  // same ~40 MB are reused by every batch and stay L2 resident
  eslam::DevBuf sym_pool_nn, sym_pool_px, sym_slotmask;   // round-1 kernel (ESLAM_SYM_IMPL=1) only
  eslam::DevBuf sym_pool256, sym_pool512, sym_slotmask2, sym_smid_map;
  cudaStream_t plane_stream2 = nullptr; cudaEvent_t plane_ev[2] = {nullptr, nullptr};   // second half of a ground-plane chunk
  int plane_streams = 1;    // env ESLAM_PLANE_STREAMS=2: a chunk as two halves on two streams (measured slower: 15.2 vs 12.3 ms per 296 frames)
  bool plane_exact = true;  // env ESLAM_PLANE_EXACT=0: every frame takes the integral-image recurrence (A/B, fallback test)
  int pull_stages = 4, pull_ctas_per_sm = 2;   // env ESLAM_PULL_STAGES (1..8 row segments in flight per pull CTA), ESLAM_PULL_CTAS (pull CTAs per SM)
  bool pull_tma = true;    // SM pull of the box rectangles through cp.async.bulk (TMA); env ESLAM_PULL_TMA=0: the LDG/STG pull kernel
  int sym_occ = 0;         // env ESLAM_SYM_OCC (tuning): 2 / 4 resident CTAs per SM instead of kSymCtas256
  int sym_impl = 3;        // env ESLAM_SYM_IMPL: 1 = round-1 kernel (A/B baseline), 2 = single launch (CTA per object, rounds inside), 3 = round pipeline (phase-pure launches)
  bool l2_discard = true;  // env ESLAM_L2_DISCARD=0: keep the dead scratch lines in L2 (they are written back to HBM when evicted)
  int step_threads = 0;    // env ESLAM_STEP_THREADS (tuning): 32 / 64 / 128 (default) threads per CTA of the symmetry step kernel
  int sym_window = 0;      // env ESLAM_SYM_WINDOW (tuning): chunks whose symmetry rounds may run at the same time (0 = all)
  bool extract_uq = true;  // cluster unions through a per-warp queue (env ESLAM_EXTRACT_UQ=0: inline unions, the A/B baseline)
  int uf_jump = 0;         // env ESLAM_UF_JUMP=n (tuning): pointer jumping over the cluster parents every n-th batch of 256 points; measured slower (0 = never)
  int extract_pf = 1;      // env ESLAM_EXTRACT_PF=2 (tuning): neighbour candidates requested two ahead instead of one
  bool sym_graph = true;   // round pipeline: a chunk's ~115 launches are captured once and replayed as a CUDA graph (env ESLAM_SYM_GRAPH=0: plain launches)
  bool sym_verify = false; // env ESLAM_SYM_VERIFY=1: every shortcut decision of the symmetry kernel is re-derived the long way and counted when it differs
  eslam::DevBuf assoc_buf;             // batched data association (associate.cu)
  eslam::DevBuf plane_depth;           // staging of eslam_extract_ground_plane_host (separate from the extraction slots)
  bool attr_extract = false, attr_symmetry = false, attr_symmetry_v1 = false, attr_plane = false;   // kernel attributes are per device: set once per handle, not once per process
  unsigned long long h2d_bytes = 0;    // depth bytes uploaded by the *_host extraction entry (running total)
  bool plane_set = false;
  double plane_world[4] = {0, 0, 1, 0};
  long long launches = 0;
  bool profiling = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float stage_ms[3] = {0, 0, 0};   // extract kernel, symmetry+finalize kernel, total of the last profiled batch
  std::string err;
  eslam::ExtractWorkspace *ex = nullptr;             // workspace of the batch being issued / last issued (one of ex_slot[])
  eslam::ExtractWorkspace *ex_slot[2] = {nullptr, nullptr};   // slot 0: synchronous entry points; both: pipelined submit / wait
  cudaStream_t slot_stream[2] = {nullptr, nullptr};
  cudaEvent_t slot_done[2] = {nullptr, nullptr};
  cudaEvent_t slot_start[2] = {nullptr, nullptr};   // device-resident pipelining: orders a slot's stream after the caller's stream
  int slot_objs[2] = {0, 0};                          // objects of the batch in flight on the slot (0 = free)
  int next_slot = 0;
  eslam::GraphWorkspace *gr = nullptr;
  eslam::SlamWorkspace *sl = nullptr;
  eslam::PlaneWorkspace *pl = nullptr;
  eslam::CopyPool *copy_pool = nullptr;
};

#define ESLAM_CUDA_OK(h, call)                                                                       \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      char _b[512];                                                                                  \
      snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      (h)->err = _b;                                                                                 \
      return ESLAM_ERR_CUDA;                                                                         \
    }                                                                                                \
  } while (0)

#define ESLAM_LAUNCHED(h) ((h)->launches++)

#ifdef __CUDACC__
namespace eslam {
// Drops the 128-byte lines of [p, p + bytes) from L2 WITHOUT write-back (PTX discard.global.L2, SASS CCTL.RML2): for scratch whose last
// reader is done.  p must be 128-byte aligned and the range must not share a line with live data; `tid` / `nthreads` spread the lines.
__device__ __forceinline__ void l2_discard_lines(const void *p, size_t bytes, int tid, int nthreads) {
  const char *a = reinterpret_cast<const char *>(p);
  for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)nthreads * 128) asm volatile("discard.global.L2 [%0], 128;" ::"l"(a + off) : "memory");
}
}  // namespace eslam
#endif

// capi.cu -- extern "C" boundary of libeslam_b200.so (include/eslam.h).  Host-side orchestration only: builds the
// tiny per-object metadata, moves bulk data H2D/D2H for the *_host entry points and launches the kernels.
// There is no CPU implementation of any hot-path stage in this library.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include "common.cuh"
#include "extract_types.cuh"
#include "internal.h"

using namespace eslam;

namespace eslam {
// Fork-join pool for the host-side staging copies of pageable input (box rows -> pinned memory).  Workers sleep between jobs.
struct CopyPool {
  std::vector<std::thread> th;
  std::mutex m; std::condition_variable cv_go, cv_done;
  std::function<void(int)> job; int n_items = 0; std::atomic<int> next{0}; int gen = 0, running = 0; bool quit = false;
  explicit CopyPool(int n) {
    for (int i = 0; i < n; i++) th.emplace_back([this] {
      int seen = 0;
      while (true) {
        { std::unique_lock<std::mutex> lk(m); cv_go.wait(lk, [&] { return quit || gen != seen; }); if (quit) return; seen = gen; }
        for (int it; (it = next.fetch_add(1)) < n_items;) job(it);
        { std::lock_guard<std::mutex> lk(m); if (--running == 0) cv_done.notify_one(); }
      }
    });
  }
  ~CopyPool() { { std::lock_guard<std::mutex> lk(m); quit = true; } cv_go.notify_all(); for (auto &t : th) t.join(); }
  void run(int n, std::function<void(int)> fn) {   // the calling thread works too
    if (n <= 0) return;
    { std::lock_guard<std::mutex> lk(m); job = std::move(fn); n_items = n; next = 0; running = (int)th.size(); gen++; }
    cv_go.notify_all();
    for (int it; (it = next.fetch_add(1)) < n_items;) job(it);
    std::unique_lock<std::mutex> lk(m); cv_done.wait(lk, [&] { return running == 0; });
  }
};
}  // namespace eslam

namespace {

int symmetry_type_of_label(int label) {  // EllipsoidExtractor::LoadSymmetryPrior, src/pca/EllipsoidExtractor.cpp:52-79
  switch (label) {
    case 58: return 0;
    case 59: case 62: case 57: case 66: case 63: case 64: case 41: return 1;
    case 28: return 2;
    default: return -1;
  }
}

// host fp64 helpers for the two trivial pose utilities of the ABI (Frame.cpp:13-14, Plane.cpp:117-122)
struct HQ { double x, y, z, w; };
void h_rot(const HQ &q, const double v[3], double o[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  for (int i = 0; i < 3; i++) uv[i] += uv[i];
  double c[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
  for (int i = 0; i < 3; i++) o[i] = v[i] + q.w * uv[i] + c[i];
}

// selects (and lazily creates) the extraction workspace the next batch uses
void use_slot(EslamHandle *h, int slot) {
  if (!h->ex_slot[slot]) h->ex_slot[slot] = new ExtractWorkspace();
  h->ex = h->ex_slot[slot];
}

ExtractParamsDev make_dev_params(const EslamHandle *h) {
  ExtractParamsDev d;
  const EslamParams &p = h->prm; const EslamCamera &c = h->cam;
  d.fx = c.fx; d.fy = c.fy; d.cx = c.cx; d.cy = c.cy; d.scale = c.scale; d.depth_range = p.depth_range;
  d.rows = c.rows; d.cols = c.cols;
  d.voxel_leaf = (float)p.voxel_leaf; d.sym_leaf = (float)p.symmetry_grid;
  for (int i = 0; i < 4; i++) d.plane[i] = h->plane_world[i];
  d.plane_margin = p.plane_margin;
  d.cluster_r2 = (float)(p.cluster_tolerance * p.cluster_tolerance);
  d.cluster_cell = (float)(p.cluster_tolerance * 1.01);
  d.min_cluster = p.min_cluster_size;
  d.center_dis = p.center_dis;
  d.sigma = p.symmetry_sigma;
  d.symmetry_open = p.symmetry_open; d.symmetry_iters = p.symmetry_iters;
  double K[9] = {c.fx, 0, c.cx, 0, c.fy, c.cy, 0, 0, 1};
  for (int i = 0; i < 9; i++) d.K[i] = K[i];
  return d;
}

// builds ObjMeta for a batch, (re)allocates the workspace, uploads the metadata
int prepare_batch(EslamHandle *h, cudaStream_t s, int n_frames, int n_obj, const int *frame_of_obj, const double *bbox4, const int *label, const double *Twc7,
                  bool want_dump, ExtractPtrs &P) {
  if (!h->ex) use_slot(h, 0);
  ExtractWorkspace &w = *h->ex;
  // next generation of the page-locked staging; wait (normally not at all) for the batch that used it two calls ago
  w.pin_cur ^= 1;
  if (w.pin_done[w.pin_cur]) ESLAM_CUDA_OK(h, cudaEventSynchronize(w.pin_done[w.pin_cur]));
  w.meta_host.resize(n_obj);
  long long off = 0;
  for (int i = 0; i < n_obj; i++) {
    ObjMeta &m = w.meta_host[i];
    memset(&m, 0, sizeof(m));
    m.frame = frame_of_obj[i]; m.label = label[i];
    m.symtype = h->prm.symmetry_open ? symmetry_type_of_label(label[i]) : -1;
    for (int k = 0; k < 4; k++) m.bbox[k] = bbox4[4 * i + k];
    // non-finite coordinates never reach the int casts (UB); an inverted box (x2 < x1 or y2 < y1) is an empty crop like in the
    // reference (status NO_POINTS_AFTER_PLANE through nx = ny = 0) and is skipped by every upload path
    const bool finite = isfinite(m.bbox[0]) && isfinite(m.bbox[1]) && isfinite(m.bbox[2]) && isfinite(m.bbox[3]) && fabs(m.bbox[0]) < 1e9 && fabs(m.bbox[1]) < 1e9 &&
                        fabs(m.bbox[2]) < 1e9 && fabs(m.bbox[3]) < 1e9;
    if (!finite) { m.x1 = m.y1 = m.x2 = m.y2 = -1; }
    else { m.x1 = (int)m.bbox[0]; m.y1 = (int)m.bbox[1]; m.x2 = (int)m.bbox[2]; m.y2 = (int)m.bbox[3]; }
    bool inside = finite && m.frame >= 0 && m.frame < n_frames && m.x1 >= 0 && m.y1 >= 0 && m.x2 <= h->cam.cols && m.y2 <= h->cam.rows;
    m.pre_status = inside ? 0 : ESLAM_OBJ_BBOX_OUTSIDE;
    m.nx = (inside && m.x2 > m.x1) ? (m.x2 - m.x1 + 2) / 3 : 0;
    m.ny = (inside && m.y2 > m.y1) ? (m.y2 - m.y1 + 2) / 3 : 0;
    if (m.nx == 0 || m.ny == 0) m.nx = m.ny = 0;   // empty crop: nothing of this box is ever read
    long long cap = (long long)m.nx * m.ny;
    cap = (cap + 31) & ~31LL;
    if (cap < 32) cap = 32;
    m.off = (int)off; m.cap = (int)cap;
    off += cap;
    if (off > 0x7fffffffLL / 4) return ESLAM_ERR_CAPACITY;
    if (inside) for (int k = 0; k < 7; k++) m.Twc[k] = Twc7[7 * m.frame + k];
  }
  w.cap_total = off; w.n_obj = n_obj;
  const size_t CAP = (size_t)off;
  ESLAM_CUDA_OK(h, w.meta.reserve(sizeof(ObjMeta) * (size_t)n_obj));
  ESLAM_CUDA_OK(h, w.state.reserve(sizeof(ObjState) * (size_t)n_obj));
  ESLAM_CUDA_OK(h, w.out.reserve(sizeof(EslamEllipsoidOut) * (size_t)n_obj));
  ESLAM_CUDA_OK(h, w.ptsA.reserve(CAP * 16)); ESLAM_CUDA_OK(h, w.ptsB.reserve(CAP * 16)); ESLAM_CUDA_OK(h, w.ptsC.reserve(CAP * 16));
  ESLAM_CUDA_OK(h, w.k0.reserve(CAP * 4)); ESLAM_CUDA_OK(h, w.k1.reserve(CAP * 4)); ESLAM_CUDA_OK(h, w.v0.reserve(CAP * 4)); ESLAM_CUDA_OK(h, w.v1.reserve(CAP * 4));
  ESLAM_CUDA_OK(h, w.parent.reserve(CAP * 4)); ESLAM_CUDA_OK(h, w.count.reserve(CAP * 4));
  ESLAM_CUDA_OK(h, w.objn.reserve(CAP * 24)); ESLAM_CUDA_OK(h, w.objnf.reserve(CAP * 16));
  P.ptsA = w.ptsA.as<float4>(); P.ptsB = w.ptsB.as<float4>(); P.ptsC = w.ptsC.as<float4>();
  P.k0 = w.k0.as<unsigned>(); P.k1 = w.k1.as<unsigned>(); P.v0 = w.v0.as<unsigned>(); P.v1 = w.v1.as<unsigned>();
  P.parent = w.parent.as<int>(); P.count = w.count.as<int>();
  P.objn = w.objn.as<double>(); P.objnf = w.objnf.as<float4>();
  P.nncache = nullptr; P.pixcache = nullptr; P.slotmask = nullptr;
  if (h->sym_impl == 1) {
    if (!h->sym_slotmask.p) {   // round-1 kernel: 2 * 256 entries per slot, 32 slots per SM id
      const size_t slots = (size_t)kMaxSmId * 32;
      ESLAM_CUDA_OK(h, h->sym_pool_nn.reserve(slots * 2 * 256 * sizeof(float4)));
      ESLAM_CUDA_OK(h, h->sym_pool_px.reserve(slots * 2 * 256 * sizeof(uint4)));
      ESLAM_CUDA_OK(h, h->sym_slotmask.reserve((size_t)kMaxSmId * sizeof(unsigned)));
      ESLAM_CUDA_OK(h, cudaMemset(h->sym_slotmask.p, 0, (size_t)kMaxSmId * sizeof(unsigned)));
    }
    P.nncache = h->sym_pool_nn.as<float4>(); P.pixcache = h->sym_pool_px.as<uint4>(); P.slotmask = h->sym_slotmask.as<unsigned>();
  }
  if (!h->sym_slotmask2.p) {   // once per handle (blocking): pools sized by the SM count, not by the batch
    const size_t sms = (size_t)h->sm_count;
    ESLAM_CUDA_OK(h, h->sym_pool256.reserve(sms * kSymSlots256 * symmetry_pool_bytes(256)));   // slot = sm * kSymSlots + bit, one per resident CTA
    ESLAM_CUDA_OK(h, h->sym_pool512.reserve(sms * kSymSlots512 * symmetry_pool_bytes(512)));
    ESLAM_CUDA_OK(h, h->sym_slotmask2.reserve(2 * sms * sizeof(unsigned)));
    ESLAM_CUDA_OK(h, h->sym_smid_map.reserve((kMaxSmId + 1) * sizeof(int)));
    std::vector<unsigned> mask(2 * sms);
    for (size_t i = 0; i < sms; i++) { mask[i] = ~0u << kSymSlots256; mask[sms + i] = ~0u << kSymSlots512; }
    std::vector<int> map(kMaxSmId + 1, -1);
    map[kMaxSmId] = 0;
    ESLAM_CUDA_OK(h, cudaMemcpy(h->sym_slotmask2.p, mask.data(), mask.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    ESLAM_CUDA_OK(h, cudaMemcpy(h->sym_smid_map.p, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (h->sym_impl == 3) {   // round pipeline: the seeds' state and caches live in HBM / L2 between the launches of a chunk
    ESLAM_CUDA_OK(h, w.symseed.reserve(symmetry_round_seed_bytes() * (size_t)n_obj));
    ESLAM_CUDA_OK(h, w.symcache.reserve(symmetry_round_cache_bytes() * (size_t)n_obj));
    ESLAM_CUDA_OK(h, w.symchunk.reserve(symmetry_round_desc_bytes() * (size_t)kMaxChunks));
    ESLAM_CUDA_OK(h, w.symlive.reserve(sizeof(int) * (size_t)kMaxChunks * (kSymMaxRounds + 2)));
    for (int g = 0; g < 2; g++) ESLAM_CUDA_OK(h, w.pin_symchunk2[g].reserve(symmetry_round_desc_bytes() * (size_t)kMaxChunks));
  }
  P.sympool256 = h->sym_pool256.as<unsigned char>(); P.sympool512 = h->sym_pool512.as<unsigned char>();
  P.slotmask2 = h->sym_slotmask2.as<unsigned>(); P.smid_map = h->sym_smid_map.as<int>(); P.sm_count = h->sm_count;
  P.l2_discard = (h->l2_discard && !want_dump) ? 1 : 0;
  P.uf_jump = h->uf_jump;
  P.d_crop = P.d_voxel = P.d_world = P.d_sym = P.d_final = nullptr; P.d_label = nullptr;
  if (want_dump) {
    ESLAM_CUDA_OK(h, w.d_crop.reserve(CAP * 24)); ESLAM_CUDA_OK(h, w.d_voxel.reserve(CAP * 24)); ESLAM_CUDA_OK(h, w.d_world.reserve(CAP * 24));
    ESLAM_CUDA_OK(h, w.d_sym.reserve(CAP * 24)); ESLAM_CUDA_OK(h, w.d_final.reserve(CAP * 96)); ESLAM_CUDA_OK(h, w.d_label.reserve(CAP * 4));
    P.d_crop = w.d_crop.as<double>(); P.d_voxel = w.d_voxel.as<double>(); P.d_world = w.d_world.as<double>();
    P.d_sym = w.d_sym.as<double>(); P.d_final = w.d_final.as<double>(); P.d_label = w.d_label.as<int>();
  }
  PinBuf &pm = w.pin_meta2[w.pin_cur];
  ESLAM_CUDA_OK(h, pm.reserve(sizeof(ObjMeta) * (size_t)n_obj));
  memcpy(pm.p, w.meta_host.data(), sizeof(ObjMeta) * (size_t)n_obj);
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(w.meta.p, pm.p, sizeof(ObjMeta) * (size_t)n_obj, cudaMemcpyHostToDevice, s));
  return ESLAM_OK;
}

// host_depth != nullptr: pipelined host entry -- the depth batch is uploaded in chunks on copy_stream while the kernels of the
// previous chunk run on `s` (objects must be ordered by frame, which is how a frame stream arrives; otherwise one chunk).
int run_extract_impl(EslamHandle *h, cudaStream_t s, int n_frames, const uint16_t *depth_dev, size_t pitch, int n_obj, const int *frame_of_obj, const double *bbox4,
                     const int *label, const double *Twc7, bool want_dump, ExtractPtrs &P, const uint16_t *host_depth);
int run_extract(EslamHandle *h, cudaStream_t s, int n_frames, const uint16_t *depth_dev, size_t pitch, int n_obj, const int *frame_of_obj, const double *bbox4,
                const int *label, const double *Twc7, bool want_dump, ExtractPtrs &P, const uint16_t *host_depth = nullptr) {
  const int rc = run_extract_impl(h, s, n_frames, depth_dev, pitch, n_obj, frame_of_obj, bbox4, label, Twc7, want_dump, P, host_depth);
  if (rc == ESLAM_OK && h->ex) {   // everything the batch reads from the staging generation is ordered before this point of s
    ExtractWorkspace &w = *h->ex;
    if (!w.pin_done[w.pin_cur]) ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&w.pin_done[w.pin_cur], cudaEventDisableTiming));
    ESLAM_CUDA_OK(h, cudaEventRecord(w.pin_done[w.pin_cur], s));
  }
  return rc;
}
int run_extract_impl(EslamHandle *h, cudaStream_t s, int n_frames, const uint16_t *depth_dev, size_t pitch, int n_obj, const int *frame_of_obj, const double *bbox4,
                     const int *label, const double *Twc7, bool want_dump, ExtractPtrs &P, const uint16_t *host_depth) {
  if (!h->plane_set) { h->err = "supporting plane not set (eslam_set_supporting_plane)"; return ESLAM_ERR_NO_PLANE; }
  if (n_obj <= 0 || n_frames <= 0 || !depth_dev || !frame_of_obj || !bbox4 || !label || !Twc7) return ESLAM_ERR_BAD_ARG;
  if (pitch < (size_t)h->cam.cols) return ESLAM_ERR_BAD_ARG;
  int rc = prepare_batch(h, s, n_frames, n_obj, frame_of_obj, bbox4, label, Twc7, want_dump, P);
  if (rc) return rc;
  ExtractParamsDev prm = make_dev_params(h);
  const size_t frame_stride = (size_t)h->cam.rows * pitch;
  if (h->profiling) ESLAM_CUDA_OK(h, cudaEventRecord(h->ev[0], s));
  if (!host_depth && (h->profiling || n_obj < 128)) {   // one stream, kernels back to back (per-kernel event times are meaningful)
    if ((rc = extract_launch(h, s, depth_dev, pitch, frame_stride, 0, n_obj, prm, P))) return rc;
    if (h->profiling) ESLAM_CUDA_OK(h, cudaEventRecord(h->ev[1], s));
    if ((rc = symmetry_finalize_launch(h, s, depth_dev, pitch, frame_stride, 0, n_obj, prm, P))) return rc;
    if (h->profiling) ESLAM_CUDA_OK(h, cudaEventRecord(h->ev[2], s));
    return ESLAM_OK;
  }
  bool sorted = true;
  for (int i = 1; i < n_obj; i++) if (frame_of_obj[i] < frame_of_obj[i - 1]) { sorted = false; break; }
  // Both kernels are latency / issue bound per CTA, so a chunk must not run alone: every chunk gets its own compute stream and the
  // extraction of one chunk overlaps the symmetry search of another (and, for the host entry, the uploads still in flight); the
  // caller's stream joins them at the end.  Device-resident input: chunks of objects.  Host input: chunks of frames (objects must be
  // ordered by frame, which is how a frame stream arrives; otherwise one chunk).
  int n_chunks = 1;
  if (!host_depth) n_chunks = n_obj >= 512 ? h->dev_chunks : (n_obj >= 256 ? 4 : 2);
  else if (sorted && n_frames >= 64) n_chunks = h->host_chunks;
  if (host_depth && !h->copy_stream) {
    int lo = 0, hi = 0;
    ESLAM_CUDA_OK(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    ESLAM_CUDA_OK(h, cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, hi));   // pull kernels go ahead of queued compute CTAs
  }
  // pinned + mapped host input (cudaHostAlloc / cudaHostRegister; torch pin_memory()): the SMs pull the rectangles themselves
  const uint16_t *host_mapped = nullptr;
  if (host_depth && h->h2d_mode == 4 && (pitch % 8) == 0 && (h->cam.cols % 8) == 0 && ((uintptr_t)depth_dev % 16) == 0) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host_depth) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer && ((uintptr_t)attr.devicePointer % 16) == 0)
      host_mapped = reinterpret_cast<const uint16_t *>(attr.devicePointer);
    else
      cudaGetLastError();
  }
  for (int c = 0; c < n_chunks; c++) {
    if (!h->chunk_ev[c]) ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->chunk_ev[c], cudaEventDisableTiming));
    if (!h->chunk_done[c]) ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->chunk_done[c], cudaEventDisableTiming));
    if (!h->chunk_stream[c]) ESLAM_CUDA_OK(h, cudaStreamCreateWithFlags(&h->chunk_stream[c], cudaStreamNonBlocking));
  }
  if (!h->chunk_ev[16]) ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->chunk_ev[16], cudaEventDisableTiming));
  ESLAM_CUDA_OK(h, cudaEventRecord(h->chunk_ev[16], s));   // metadata upload (prepare_batch) happened on s; earlier work on s is ordered too
  const ExtractWorkspace &w = *h->ex;
  const int rows = h->cam.rows, cols = h->cam.cols;
  const bool trace = getenv("ESLAM_TRACE") != nullptr;
  const auto t_issue0 = std::chrono::steady_clock::now();
  int o = 0;
  for (int c = 0; c < n_chunks; c++) {
    cudaStream_t cs = (n_chunks == 1) ? s : h->chunk_stream[c];
    int o1;
    if (host_depth) {
      const int f0 = (int)((long long)n_frames * c / n_chunks), f1 = (int)((long long)n_frames * (c + 1) / n_chunks);
      o1 = o;
      if (sorted) { while (o1 < n_obj && frame_of_obj[o1] < f1) o1++; } else o1 = n_obj;
      if (c == n_chunks - 1) o1 = n_obj;
      if (host_mapped) {
        ESLAM_CUDA_OK(h, cudaStreamWaitEvent(h->copy_stream, h->chunk_ev[16], 0));   // needs the ObjMeta table
        if ((rc = pull_rects_launch(h, h->copy_stream, host_mapped, const_cast<uint16_t *>(depth_dev), pitch, frame_stride, o, o1 - o))) return rc;
        for (int k = o; k < o1; k++) {
          const ObjMeta &m = w.meta_host[k];
          if (m.pre_status != 0 || m.nx == 0) continue;
          const int x0 = m.x1 & ~7, x1 = (m.x2 + 8) & ~7, y1 = m.y2 + 1;
          h->h2d_bytes += (size_t)((x1 < cols ? x1 : cols) - x0) * (size_t)((y1 < rows ? y1 : rows) - m.y1) * sizeof(uint16_t);
        }
      } else if (h->h2d_mode == 4 && (pitch % 8) == 0 && (h->cam.cols % 8) == 0 && ((uintptr_t)depth_dev % 16) == 0) {
        // Pageable host input (a plain cv::Mat): the rows of this chunk's box rectangles are copied by a few host threads into the
        // workspace's pinned staging buffer (same layout as the frame batch), then pulled by the bulk-copy kernel like mapped input.
        // The host copies chunk c + 1 while the GPU works on chunk c.
        ExtractWorkspace &ww = *h->ex;
        const size_t need = (size_t)n_frames * frame_stride * sizeof(uint16_t);
        if (ww.pin_depth.bytes < need) ESLAM_CUDA_OK(h, ww.pin_depth.reserve(need));
        uint16_t *stage = ww.pin_depth.as<uint16_t>();
        if (!h->copy_pool) { unsigned hw = std::thread::hardware_concurrency(); int nt = hw >= 16 ? 7 : (hw >= 4 ? (int)hw / 2 - 1 : 0); h->copy_pool = new eslam::CopyPool(nt); }
        std::vector<int> &rows_of = ww.stage_rows;   // prefix of row counts per object of the chunk
        rows_of.assign((size_t)(o1 - o) + 1, 0);
        for (int k = o; k < o1; k++) {
          const ObjMeta &m = w.meta_host[k];
          int nr = 0;
          if (m.pre_status == 0 && m.nx != 0) { const int y1 = m.y2 + 1 < rows ? m.y2 + 1 : rows; nr = y1 - m.y1; }
          rows_of[(size_t)(k - o) + 1] = rows_of[(size_t)(k - o)] + (nr > 0 ? nr : 0);
        }
        const int total_rows = rows_of.back(), per = 16, n_items = (total_rows + per - 1) / per;
        const int obase = o;
        h->copy_pool->run(n_items, [&, stage, obase](int it) {
          int r0 = it * per, r1 = r0 + per < total_rows ? r0 + per : total_rows;
          int k = (int)(std::upper_bound(rows_of.begin(), rows_of.end(), r0) - rows_of.begin()) - 1;
          for (int r = r0; r < r1; r++) {
            while (rows_of[(size_t)k + 1] <= r) k++;
            const ObjMeta &m = w.meta_host[obase + k];
            const int x0 = m.x1 & ~7, x1a = (m.x2 + 8) & ~7, x1 = x1a < cols ? x1a : cols;
            const size_t off = (size_t)m.frame * frame_stride + (size_t)(m.y1 + (r - rows_of[(size_t)k])) * pitch + x0;
            memcpy(stage + off, host_depth + off, (size_t)(x1 - x0) * sizeof(uint16_t));
          }
        });
        for (int k = o; k < o1; k++) {
          const ObjMeta &m = w.meta_host[k];
          if (m.pre_status != 0 || m.nx == 0) continue;
          const int x0 = m.x1 & ~7, x1 = (m.x2 + 8) & ~7, y1 = m.y2 + 1;
          h->h2d_bytes += (size_t)((x1 < cols ? x1 : cols) - x0) * (size_t)((y1 < rows ? y1 : rows) - m.y1) * sizeof(uint16_t);
        }
        ESLAM_CUDA_OK(h, cudaStreamWaitEvent(h->copy_stream, h->chunk_ev[16], 0));   // needs the ObjMeta table
        if ((rc = pull_rects_launch(h, h->copy_stream, stage, const_cast<uint16_t *>(depth_dev), pitch, frame_stride, o, o1 - o))) return rc;
      } else if (sorted && h->h2d_mode != 0) {
        // The kernels only ever read pixels inside the (int-truncated) boxes: crop and GetCenter sample inside the box, the lazy range
        // image is read where isInRange holds.  Upload the union rectangle of each frame's boxes (64-byte aligned columns) instead
        // of the whole 614 400-byte frame; frames without a valid box are not uploaded at all.
        int k = o;
        std::vector<cudaMemcpy3DBatchOp> &ops = h->ex->copy_ops;
        ops.clear();
        while (k < o1) {
          const int f = frame_of_obj[k];
          int x0 = cols, y0 = rows, x1 = -1, y1 = -1;
          for (; k < o1 && frame_of_obj[k] == f; k++) {
            const ObjMeta &m = w.meta_host[k];
            if (m.pre_status != 0 || m.nx == 0) continue;
            x0 = m.x1 < x0 ? m.x1 : x0; y0 = m.y1 < y0 ? m.y1 : y0; x1 = m.x2 > x1 ? m.x2 : x1; y1 = m.y2 > y1 ? m.y2 : y1;
          }
          if (x1 < 0) continue;
          x0 &= ~31; x1 = (x1 + 32) & ~31; if (x1 > cols) x1 = cols;
          y1 = y1 + 1 < rows ? y1 + 1 : rows;
          if (h->h2d_mode == 1) { x0 = 0; x1 = cols; }
          const size_t offp = (size_t)f * frame_stride + (size_t)y0 * pitch + x0;
          uint16_t *dst = const_cast<uint16_t *>(depth_dev) + offp;
          const uint16_t *src = host_depth + offp;
          const size_t wbytes = (size_t)(x1 - x0) * sizeof(uint16_t), nrows = (size_t)(y1 - y0);
          h->h2d_bytes += wbytes * nrows;
          if (h->h2d_mode >= 3) {   // one batched call per chunk: the per-frame rectangles as pointer-to-pointer 3-D ops (element = 1 byte)
            cudaMemcpy3DBatchOp op;
            memset(&op, 0, sizeof(op));
            op.src.type = cudaMemcpyOperandTypePointer; op.src.op.ptr.ptr = const_cast<uint16_t *>(src);
            op.src.op.ptr.rowLength = pitch * sizeof(uint16_t); op.src.op.ptr.layerHeight = nrows;
            op.dst.type = cudaMemcpyOperandTypePointer; op.dst.op.ptr.ptr = dst;
            op.dst.op.ptr.rowLength = pitch * sizeof(uint16_t); op.dst.op.ptr.layerHeight = nrows;
            op.extent.width = wbytes; op.extent.height = nrows; op.extent.depth = 1;
            op.srcAccessOrder = cudaMemcpySrcAccessOrderStream; op.flags = 0;
            ops.push_back(op);
          } else if (h->h2d_mode == 1 || (long long)(x1 - x0) * 4 >= (long long)pitch * 3) {   // (nearly) full rows: one contiguous copy
            ESLAM_CUDA_OK(h, cudaMemcpyAsync(dst - x0, src - x0, nrows * pitch * sizeof(uint16_t), cudaMemcpyHostToDevice, h->copy_stream));
          } else {
            ESLAM_CUDA_OK(h, cudaMemcpy2DAsync(dst, pitch * sizeof(uint16_t), src, pitch * sizeof(uint16_t), wbytes, nrows, cudaMemcpyHostToDevice, h->copy_stream));
          }
        }
        if (!ops.empty()) {
          size_t fail = 0;
          cudaError_t e = cudaMemcpy3DBatchAsync(ops.size(), ops.data(), &fail, 0, h->copy_stream);
          if (e != cudaSuccess) {   // driver without batched copies: issue them one by one, and stay in that mode
            cudaGetLastError();
            h->h2d_mode = 2;
            for (const cudaMemcpy3DBatchOp &op : ops)
              ESLAM_CUDA_OK(h, cudaMemcpy2DAsync(op.dst.op.ptr.ptr, op.dst.op.ptr.rowLength, op.src.op.ptr.ptr, op.src.op.ptr.rowLength, op.extent.width, op.extent.height,
                                                 cudaMemcpyHostToDevice, h->copy_stream));
          }
        }
      } else if (f1 > f0) {
        ESLAM_CUDA_OK(h, cudaMemcpyAsync(const_cast<uint16_t *>(depth_dev) + (size_t)f0 * frame_stride, host_depth + (size_t)f0 * frame_stride,
                                         (size_t)(f1 - f0) * frame_stride * sizeof(uint16_t), cudaMemcpyHostToDevice, h->copy_stream));
        h->h2d_bytes += (size_t)(f1 - f0) * frame_stride * sizeof(uint16_t);
      }
      ESLAM_CUDA_OK(h, cudaEventRecord(h->chunk_ev[c], h->copy_stream));
      ESLAM_CUDA_OK(h, cudaStreamWaitEvent(cs, h->chunk_ev[c], 0));
    } else {
      o1 = (int)((long long)n_obj * (c + 1) / n_chunks);
    }
    if (cs != s) ESLAM_CUDA_OK(h, cudaStreamWaitEvent(cs, h->chunk_ev[16], 0));
    if ((rc = extract_launch(h, cs, depth_dev, pitch, frame_stride, o, o1 - o, prm, P))) return rc;
    // ESLAM_SYM_WINDOW=k (tuning): the symmetry rounds of chunk c start when chunk c - k is done, so that at most k chunks' seed
    // caches (~123 KB per object) compete for L2; extraction of the later chunks is not held back
    if (h->sym_window > 0 && c >= h->sym_window && cs != s) ESLAM_CUDA_OK(h, cudaStreamWaitEvent(cs, h->chunk_done[c - h->sym_window], 0));
    if ((rc = symmetry_finalize_launch(h, cs, depth_dev, pitch, frame_stride, o, o1 - o, prm, P, c))) return rc;
    if (cs != s) { ESLAM_CUDA_OK(h, cudaEventRecord(h->chunk_done[c], cs)); ESLAM_CUDA_OK(h, cudaStreamWaitEvent(s, h->chunk_done[c], 0)); }
    o = o1;
  }
  if (h->profiling) { ESLAM_CUDA_OK(h, cudaEventRecord(h->ev[1], s)); ESLAM_CUDA_OK(h, cudaEventRecord(h->ev[2], s)); }
  if (trace) fprintf(stderr, "[eslam] run_extract: %d chunks, host issue %.3f ms (h2d_mode %d%s)\n", n_chunks,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_issue0).count(), h->h2d_mode, host_mapped ? ", SM pull" : "");
  return ESLAM_OK;
}

int fetch_dump(EslamHandle *h, cudaStream_t s, int n_obj, const ExtractPtrs &P, const EslamEllipsoidOut *out, EslamStageDump *dump) {
  ExtractWorkspace &w = *h->ex;
  std::vector<ObjState> states(n_obj);
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(states.data(), w.state.p, sizeof(ObjState) * (size_t)n_obj, cudaMemcpyDeviceToHost, s));
  ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));
  for (int i = 0; i < n_obj; i++) {
    const ObjMeta &m = w.meta_host[i];
    const ObjState &st = states[i];
    EslamStageDump &d = dump[i];
    const int cap = d.cap_points;
    auto cp = [&](double *dst, const double *src_dev, size_t off_pts, int n) -> int {
      if (!dst || n <= 0) return 0;
      int k = n < cap ? n : cap;
      cudaError_t e = cudaMemcpyAsync(dst, src_dev + 3 * off_pts, sizeof(double) * 3 * (size_t)k, cudaMemcpyDeviceToHost, s);
      return e == cudaSuccess ? 0 : 1;
    };
    int bad = 0;
    bad |= cp(d.crop_xyz, P.d_crop, m.off, out[i].n_crop);
    bad |= cp(d.voxel_xyz, P.d_voxel, m.off, out[i].n_voxel);
    bad |= cp(d.world_xyz, P.d_world, m.off, out[i].n_plane);
    bad |= cp(d.sym_xyz, P.d_sym, m.off, out[i].n_sym);
    bad |= cp(d.final_xyz, P.d_final, (size_t)4 * m.off, out[i].n_final);
    if (d.cluster_label && out[i].n_plane > 0 && out[i].status != ESLAM_OBJ_NO_POINTS_AFTER_PLANE && out[i].status != ESLAM_OBJ_NO_CENTER) {
      int k = out[i].n_plane < cap ? out[i].n_plane : cap;
      if (cudaMemcpyAsync(d.cluster_label, P.d_label + m.off, sizeof(int) * (size_t)k, cudaMemcpyDeviceToHost, s) != cudaSuccess) bad = 1;
    }
    if (bad) { h->err = "dump copy failed"; return ESLAM_ERR_CUDA; }
    for (int k = 0; k < 3; k++) { d.center_world[k] = st.center[k]; d.pca_center[k] = st.pca_center[k]; d.pca_eval[k] = st.pca_eval[k]; }
    for (int k = 0; k < 9; k++) d.pca_rot[k] = st.pca_rot[k];
    for (int k = 0; k < 7; k++) d.T_wo[k] = st.T_wo[k];
    for (int sd = 0; sd < 9; sd++) {
      d.seed_prob[sd] = st.seed_prob[sd]; d.seed_init_err[sd] = st.seed_init[sd]; d.seed_final_err[sd] = st.seed_final[sd]; d.seed_dual_dis[sd] = st.seed_dual[sd];
      for (int k = 0; k < 4; k++) d.seed_plane[sd][k] = st.seed_plane[sd][k];
    }
  }
  ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));
  return ESLAM_OK;
}

}  // namespace

extern "C" {

int eslam_abi_version(void) { return ESLAM_ABI_VERSION; }

void eslam_default_params(EslamParams *p) {
  p->depth_range = 6.0; p->voxel_leaf = 0.01; p->plane_margin = 0.05; p->cluster_tolerance = 0.02; p->min_cluster_size = 100; p->center_dis = 0.5;
  p->symmetry_open = 1; p->symmetry_grid = 0.1; p->symmetry_sigma = 0.1; p->symmetry_iters = 5;
  p->plane_min_size = 200; p->plane_angle_deg = 5.0; p->plane_dist = 0.1;
  p->edge3d_scale = 10000.0; p->gravity_prior_open = 1; p->gravity_prior_scale = 100.0; p->border_pixels = 10; p->length_limit_pixels = 0;
}
void eslam_default_camera(EslamCamera *c) { c->fx = 535.4; c->fy = 539.2; c->cx = 320.1; c->cy = 247.6; c->scale = 5000.0; c->rows = 480; c->cols = 640; }

int eslam_create(const EslamParams *p, const EslamCamera *c, int device, EslamHandle **out) {
  if (!p || !c || !out) return ESLAM_ERR_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return ESLAM_ERR_NO_DEVICE;  // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return ESLAM_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ESLAM_ERR_NO_DEVICE;
  EslamHandle *h = new (std::nothrow) EslamHandle();
  if (!h) return ESLAM_ERR_CUDA;
  h->prm = *p; h->cam = *c; h->device = device; h->sm_count = prop.multiProcessorCount;
  if (const char *e = getenv("ESLAM_H2D_MODE")) { int v = atoi(e); if (v >= 0 && v <= 4) h->h2d_mode = v; }
  if (const char *e = getenv("ESLAM_SYM_IMPL")) { int v = atoi(e); if (v >= 1 && v <= 3) h->sym_impl = v; }
  if (const char *e = getenv("ESLAM_PLANE_STREAMS")) h->plane_streams = atoi(e);
  if (const char *e = getenv("ESLAM_PLANE_EXACT")) h->plane_exact = atoi(e) != 0;
  if (const char *e = getenv("ESLAM_PULL_TMA")) h->pull_tma = atoi(e) != 0;
  if (const char *e = getenv("ESLAM_SYM_OCC")) h->sym_occ = atoi(e);
  if (const char *e = getenv("ESLAM_PULL_STAGES")) { int v = atoi(e); if (v >= 1 && v <= 8) h->pull_stages = v; }
  if (const char *e = getenv("ESLAM_PULL_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 8) h->pull_ctas_per_sm = v; }
  if (const char *e = getenv("ESLAM_L2_DISCARD")) h->l2_discard = atoi(e) != 0;
  if (const char *e = getenv("ESLAM_STEP_THREADS")) h->step_threads = atoi(e);
  if (const char *e = getenv("ESLAM_SYM_WINDOW")) h->sym_window = atoi(e);
  if (const char *e = getenv("ESLAM_EXTRACT_UQ")) h->extract_uq = atoi(e) != 0;
  if (const char *e = getenv("ESLAM_UF_JUMP")) { int v = atoi(e); if (v >= 0 && v <= 64) h->uf_jump = v; }
  if (const char *e = getenv("ESLAM_EXTRACT_PF")) h->extract_pf = atoi(e) == 2 ? 2 : 1;
  if (const char *e = getenv("ESLAM_SYM_GRAPH")) h->sym_graph = atoi(e) != 0;
  if (const char *e = getenv("ESLAM_SYM_VERIFY")) h->sym_verify = atoi(e) != 0;
  if (const char *e = getenv("ESLAM_DEV_CHUNKS")) { int v = atoi(e); if (v >= 1 && v <= kMaxChunks) h->dev_chunks = v; }
  if (const char *e = getenv("ESLAM_HOST_CHUNKS")) { int v = atoi(e); if (v >= 1 && v <= kMaxChunks) h->host_chunks = v; }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return ESLAM_ERR_CUDA; }
  *out = h;
  return ESLAM_OK;
}

void eslam_destroy(EslamHandle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (int sl = 0; sl < 2; sl++) {
    if (!h->ex_slot[sl]) continue;
    ExtractWorkspace &w = *h->ex_slot[sl];
    DevBuf *bufs[] = {&w.meta, &w.state, &w.out, &w.ptsA, &w.ptsB, &w.ptsC, &w.k0, &w.k1, &w.v0, &w.v1, &w.parent, &w.count, &w.objn, &w.objnf,
                      &w.d_crop, &w.d_voxel, &w.d_world, &w.d_sym, &w.d_final, &w.d_label, &w.depth};
    for (DevBuf *b : bufs) b->release();
    w.symseed.release(); w.symcache.release(); w.symchunk.release(); w.symlive.release(); w.pin_symchunk2[0].release(); w.pin_symchunk2[1].release();
    for (int g = 0; g < 2; g++) if (w.pin_done[g]) cudaEventDestroy(w.pin_done[g]);
    for (auto &g : w.symgraph) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    w.pin_meta2[0].release(); w.pin_meta2[1].release(); w.pin_out.release(); w.pin_depth.release();
    delete h->ex_slot[sl];
    if (h->slot_stream[sl]) cudaStreamDestroy(h->slot_stream[sl]);
    if (h->slot_done[sl]) cudaEventDestroy(h->slot_done[sl]);
    if (h->slot_start[sl]) cudaEventDestroy(h->slot_start[sl]);
  }
  h->assoc_buf.release();
  h->sym_pool_nn.release(); h->sym_pool_px.release(); h->sym_slotmask.release(); h->plane_depth.release();
  h->sym_pool256.release(); h->sym_pool512.release(); h->sym_slotmask2.release(); h->sym_smid_map.release();
  delete h->copy_pool; h->copy_pool = nullptr;
  slam_free(h);
  graph_free(h);
  plane_free(h);
  for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 17; i++) if (h->chunk_ev[i]) cudaEventDestroy(h->chunk_ev[i]);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < kMaxChunks; i++) { if (h->chunk_stream[i]) cudaStreamDestroy(h->chunk_stream[i]); if (h->chunk_done[i]) cudaEventDestroy(h->chunk_done[i]); }
  for (int i = 0; i < kMaxChunks; i++) { if (h->chunk_aux[i]) cudaStreamDestroy(h->chunk_aux[i]); if (h->aux_fork[i]) cudaEventDestroy(h->aux_fork[i]); if (h->aux_join[i]) cudaEventDestroy(h->aux_join[i]); }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// per-kernel device times of the most recent extraction batch (CUDA events on the launching stream)
int eslam_set_profiling(EslamHandle *h, int on) {
  if (!h) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  if (on && !h->ev[0]) for (int i = 0; i < 4; i++) ESLAM_CUDA_OK(h, cudaEventCreate(&h->ev[i]));
  h->profiling = on != 0;
  return ESLAM_OK;
}
int eslam_get_stage_ms(EslamHandle *h, float ms[3]) {
  if (!h || !h->profiling || !h->ev[0]) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  ESLAM_CUDA_OK(h, cudaEventSynchronize(h->ev[2]));
  ESLAM_CUDA_OK(h, cudaEventElapsedTime(&ms[0], h->ev[0], h->ev[1]));
  ESLAM_CUDA_OK(h, cudaEventElapsedTime(&ms[1], h->ev[1], h->ev[2]));
  ESLAM_CUDA_OK(h, cudaEventElapsedTime(&ms[2], h->ev[0], h->ev[2]));
  return ESLAM_OK;
}

const char *eslam_last_error(EslamHandle *h) { return h ? h->err.c_str() : "null handle"; }
long long eslam_launch_count(EslamHandle *h) { return h ? h->launches : 0; }
unsigned long long eslam_h2d_bytes(EslamHandle *h) { return h ? h->h2d_bytes : 0; }

size_t eslam_sizeof_params(void) { return sizeof(EslamParams); }
size_t eslam_sizeof_camera(void) { return sizeof(EslamCamera); }
size_t eslam_sizeof_ellipsoid_out(void) { return sizeof(EslamEllipsoidOut); }
size_t eslam_sizeof_stage_dump(void) { return sizeof(EslamStageDump); }
size_t eslam_sizeof_edge_batch(void) { return sizeof(EslamEdgeBatch); }
size_t eslam_sizeof_lm_stats(void) { return sizeof(EslamLmStats); }

// ---------------- ground plane ----------------
int eslam_extract_ground_plane_dev(EslamHandle *h, void *stream, int n_frames, const uint16_t *depth_dev, size_t pitch, double *plane_cam_host, int *found_host,
                                   int *n_inliers_host) {
  if (!h || n_frames <= 0 || !depth_dev || !plane_cam_host || !found_host) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  return plane_extract(h, s, n_frames, depth_dev, pitch, plane_cam_host, found_host, n_inliers_host);
}
int eslam_extract_ground_plane_host(EslamHandle *h, int n_frames, const uint16_t *depth, size_t pitch, double *plane_cam, int *found, int *n_inliers) {
  if (!h || n_frames <= 0 || !depth || !plane_cam || !found || pitch < (size_t)h->cam.cols) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  // own staging buffer: never aliases the extraction slots, so a batch submitted through _submit may still be in flight
  size_t bytes = (size_t)n_frames * h->cam.rows * pitch * sizeof(uint16_t);
  ESLAM_CUDA_OK(h, h->plane_depth.reserve(bytes));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(h->plane_depth.p, depth, bytes, cudaMemcpyHostToDevice, h->stream));
  return plane_extract(h, h->stream, n_frames, h->plane_depth.as<uint16_t>(), pitch, plane_cam, found, n_inliers);
}

// test/bench hook: symmetry work counters of object `obj` of the last batch: out = {cost evaluations (per half), NN scans, mirrored points, n_sym, NN cache hits}
int eslam_debug_counters(EslamHandle *h, int obj, unsigned long long out[5]) {
  if (!h || !h->ex || obj < 0 || obj >= h->ex->n_obj) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  ObjState st;
  ESLAM_CUDA_OK(h, cudaMemcpy(&st, h->ex->state.as<ObjState>() + obj, sizeof(ObjState), cudaMemcpyDeviceToHost));
  out[0] = st.n_evals; out[1] = st.n_queries; out[2] = st.n_points; out[3] = (unsigned long long)st.n_sym; out[4] = st.n_hits;
  return ESLAM_OK;
}

// test/bench hook: clock64() stamps of the extraction phases of object `obj` of the last batch
int eslam_debug_phase_clocks(EslamHandle *h, int obj, long long out[13]) {
  if (!h || !h->ex || obj < 0 || obj >= h->ex->n_obj) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  ObjState st;
  ESLAM_CUDA_OK(h, cudaMemcpy(&st, h->ex->state.as<ObjState>() + obj, sizeof(ObjState), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 9; i++) out[i] = st.phase_clk[i];
  for (int i = 0; i < 4; i++) out[9 + i] = st.sym_clk[i];
  return ESLAM_OK;
}

// test-only hook (not in include/eslam.h): intermediate buffers of the plane pipeline for stage-isolated parity tests
int eslam_debug_plane_stage(EslamHandle *h, const uint16_t *depth, size_t pitch, int stage, void *out) {
  if (!h || !depth || !out || stage < 1 || stage > 4 || pitch < (size_t)h->cam.cols) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  size_t bytes = (size_t)h->cam.rows * pitch * sizeof(uint16_t);
  ESLAM_CUDA_OK(h, h->plane_depth.reserve(bytes));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(h->plane_depth.p, depth, bytes, cudaMemcpyHostToDevice, h->stream));
  double pl[4]; int found = 0, inl = 0;
  return plane_extract_dbg(h, h->stream, 1, h->plane_depth.as<uint16_t>(), pitch, pl, &found, &inl, stage, out);
}

void eslam_plane_to_world(const double pc[4], const double Twc[7], double pw[4]) {
  // param = (Twc^T)^-1 * param with Twc = to_homogeneous_matrix() of the (possibly non-unit) quaternion pose:
  // generic 4x4 inverse like the reference (src/core/Plane.cpp:117-122), no rigid-body shortcut.
  const double x = Twc[3], y = Twc[4], z = Twc[5], w = Twc[6];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  double T[4][4] = {{1.0 - (tyy + tzz), txy - twz, txz + twy, Twc[0]},
                    {txy + twz, 1.0 - (txx + tzz), tyz - twx, Twc[1]},
                    {txz - twy, tyz + twx, 1.0 - (txx + tyy), Twc[2]},
                    {0, 0, 0, 1}};
  double a[4][8];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { a[i][j] = T[j][i]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }   // transpose | I
  for (int c = 0; c < 4; c++) {
    int p = c; double best = fabs(a[c][c]);
    for (int r = c + 1; r < 4; r++) if (fabs(a[r][c]) > best) { best = fabs(a[r][c]); p = r; }
    if (p != c) for (int j = 0; j < 8; j++) { double t = a[p][j]; a[p][j] = a[c][j]; a[c][j] = t; }
    double piv = a[c][c];
    for (int j = 0; j < 8; j++) a[c][j] /= piv;
    for (int r = 0; r < 4; r++) if (r != c) { double f = a[r][c]; if (f != 0.0) for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j]; }
  }
  for (int i = 0; i < 4; i++) { double s = a[i][4] * pc[0]; for (int k = 1; k < 4; k++) s += a[i][4 + k] * pc[k]; pw[i] = s; }
}

int eslam_set_supporting_plane(EslamHandle *h, const double pw[4]) {
  if (!h || !pw) return ESLAM_ERR_BAD_ARG;
  for (int i = 0; i < 4; i++) h->plane_world[i] = pw[i];
  h->plane_set = true;
  return ESLAM_OK;
}

int eslam_calibrate_measurement(const EslamParams *p, const EslamCamera *c, const double m[4]) {  // dataprocess_utils.cpp:150-190
  int x_length = (int)(m[2] - m[0]), y_length = (int)(m[3] - m[1]);
  if (x_length < p->length_limit_pixels || y_length < p->length_limit_pixels) return 1;
  int correct = 0; const int b = p->border_pixels;
  if (m[0] > b && m[0] < c->cols - 1 - b) correct++;
  if (m[2] > b && m[2] < c->cols - 1 - b) correct++;
  if (m[1] > b && m[1] < c->rows - 1 - b) correct++;
  if (m[3] > b && m[3] < c->rows - 1 - b) correct++;
  return correct != 4 ? 1 : 0;
}

// ---------------- ellipsoid extraction ----------------
int eslam_estimate_ellipsoids_dev(EslamHandle *h, void *stream, int n_frames, const uint16_t *depth_dev, size_t pitch, int n_obj, const int *frame_of_obj,
                                  const double *bbox4, const int *label, const double *Twc7) {
  if (!h) return ESLAM_ERR_BAD_ARG;
  if (h->slot_objs[0] || h->slot_objs[1]) { h->err = "a submitted batch is still in flight (eslam_estimate_ellipsoids_wait)"; return ESLAM_ERR_BAD_ARG; }
  cudaSetDevice(h->device);
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  use_slot(h, 0);
  ExtractPtrs P;
  return run_extract(h, s, n_frames, depth_dev, pitch, n_obj, frame_of_obj, bbox4, label, Twc7, false, P);
}
int eslam_fetch_ellipsoids(EslamHandle *h, void *stream, int n_obj, EslamEllipsoidOut *out) {
  if (!h || !h->ex || !out || n_obj <= 0 || n_obj > h->ex->n_obj) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(out, h->ex->out.p, sizeof(EslamEllipsoidOut) * (size_t)n_obj, cudaMemcpyDeviceToHost, s));
  ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));
  return ESLAM_OK;
}
int eslam_copy_ellipsoids_dev(EslamHandle *h, void *stream, int n_obj, void *records_dev) {
  if (!h || !h->ex || !records_dev || n_obj <= 0 || n_obj > h->ex->n_obj) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(records_dev, h->ex->out.p, sizeof(EslamEllipsoidOut) * (size_t)n_obj, cudaMemcpyDeviceToDevice, s));
  return ESLAM_OK;
}
int eslam_batch_counters(EslamHandle *h, unsigned long long out[12]) {
  if (!h || !h->ex || !out || h->ex->n_obj <= 0) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  std::vector<ObjState> st((size_t)h->ex->n_obj);
  ESLAM_CUDA_OK(h, cudaMemcpy(st.data(), h->ex->state.p, sizeof(ObjState) * st.size(), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 12; k++) out[k] = 0;
  for (const ObjState &o : st) {
    out[0] += o.n_ref_evals; out[1] += o.n_ref_points; out[2] += o.n_evals; out[3] += o.n_points; out[4] += o.n_queries; out[5] += o.n_hits;
    if ((unsigned long long)o.n_sym > out[6]) out[6] = (unsigned long long)o.n_sym;
    out[8] += o.n_verify_fail;
    for (int k = 0; k < 3; k++) out[9 + k] += (unsigned long long)o.sym_clk[k == 2 ? 3 : k];   // scalar phases, task phases, whole CTA
    if (o.run_sym) out[7]++;
  }
  return ESLAM_OK;
}
int eslam_estimate_ellipsoids_host(EslamHandle *h, int n_frames, const uint16_t *depth, size_t pitch, int n_obj, const int *frame_of_obj, const double *bbox4,
                                   const int *label, const double *Twc7, EslamEllipsoidOut *out, EslamStageDump *dump) {
  if (!h || !depth || !out || n_frames <= 0) return ESLAM_ERR_BAD_ARG;
  if (h->slot_objs[0] || h->slot_objs[1]) { h->err = "a submitted batch is still in flight (eslam_estimate_ellipsoids_wait)"; return ESLAM_ERR_BAD_ARG; }
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  use_slot(h, 0);
  size_t bytes = (size_t)n_frames * h->cam.rows * pitch * sizeof(uint16_t);
  ESLAM_CUDA_OK(h, h->ex->depth.reserve(bytes));
  ExtractPtrs P;
  int rc = run_extract(h, s, n_frames, h->ex->depth.as<uint16_t>(), pitch, n_obj, frame_of_obj, bbox4, label, Twc7, dump != nullptr, P, depth);
  if (rc) return rc;
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(out, h->ex->out.p, sizeof(EslamEllipsoidOut) * (size_t)n_obj, cudaMemcpyDeviceToHost, s));
  ESLAM_CUDA_OK(h, cudaStreamSynchronize(s));
  if (dump) return fetch_dump(h, s, n_obj, P, out, dump);
  return ESLAM_OK;
}

// Pipelined form of eslam_estimate_ellipsoids_host for frame streams: up to two batches in flight, each on its own workspace, so the
// upload and the ramp-up of batch k+1 run under the kernels of batch k.  `depth` must stay valid until the matching wait.
int eslam_estimate_ellipsoids_submit(EslamHandle *h, int n_frames, const uint16_t *depth, size_t pitch, int n_obj, const int *frame_of_obj, const double *bbox4,
                                     const int *label, const double *Twc7, int *ticket) {
  if (!h || !depth || !ticket || n_frames <= 0 || n_obj <= 0) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const int sl = h->next_slot;
  if (h->slot_objs[sl]) { h->err = "both pipeline slots are busy: wait for the oldest ticket first"; return ESLAM_ERR_BAD_ARG; }
  if (!h->slot_stream[sl]) ESLAM_CUDA_OK(h, cudaStreamCreateWithFlags(&h->slot_stream[sl], cudaStreamNonBlocking));
  if (!h->slot_done[sl]) ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->slot_done[sl], cudaEventDisableTiming));
  use_slot(h, sl);
  cudaStream_t s = h->slot_stream[sl];
  size_t bytes = (size_t)n_frames * h->cam.rows * pitch * sizeof(uint16_t);
  ESLAM_CUDA_OK(h, h->ex->depth.reserve(bytes));
  ExtractPtrs P;
  int rc = run_extract(h, s, n_frames, h->ex->depth.as<uint16_t>(), pitch, n_obj, frame_of_obj, bbox4, label, Twc7, false, P, depth);
  if (rc) return rc;
  ESLAM_CUDA_OK(h, h->ex->pin_out.reserve(sizeof(EslamEllipsoidOut) * (size_t)n_obj));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(h->ex->pin_out.p, h->ex->out.p, sizeof(EslamEllipsoidOut) * (size_t)n_obj, cudaMemcpyDeviceToHost, s));
  ESLAM_CUDA_OK(h, cudaEventRecord(h->slot_done[sl], s));
  h->slot_objs[sl] = n_obj;
  h->next_slot = sl ^ 1;
  *ticket = sl;
  return ESLAM_OK;
}
int eslam_estimate_ellipsoids_wait(EslamHandle *h, int ticket, int n_obj, EslamEllipsoidOut *out) {
  if (!h || !out || ticket < 0 || ticket > 1 || !h->slot_objs[ticket] || n_obj != h->slot_objs[ticket]) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  ESLAM_CUDA_OK(h, cudaEventSynchronize(h->slot_done[ticket]));
  memcpy(out, h->ex_slot[ticket]->pin_out.p, sizeof(EslamEllipsoidOut) * (size_t)n_obj);
  h->slot_objs[ticket] = 0;
  return ESLAM_OK;
}

int eslam_estimate_ellipsoids_dev_submit(EslamHandle *h, void *stream, int n_frames, const uint16_t *depth_dev, size_t pitch, int n_obj, const int *frame_of_obj,
                                         const double *bbox4, const int *label, const double *Twc7, int *ticket) {
  if (!h || !depth_dev || !ticket || n_frames <= 0 || n_obj <= 0) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  const int sl = h->next_slot;
  if (h->slot_objs[sl]) { h->err = "both pipeline slots are busy: collect the oldest ticket first"; return ESLAM_ERR_BAD_ARG; }
  if (!h->slot_stream[sl]) ESLAM_CUDA_OK(h, cudaStreamCreateWithFlags(&h->slot_stream[sl], cudaStreamNonBlocking));
  if (!h->slot_done[sl]) ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->slot_done[sl], cudaEventDisableTiming));
  if (!h->slot_start[sl]) ESLAM_CUDA_OK(h, cudaEventCreateWithFlags(&h->slot_start[sl], cudaEventDisableTiming));
  use_slot(h, sl);
  cudaStream_t s = h->slot_stream[sl], user = stream ? (cudaStream_t)stream : h->stream;
  ESLAM_CUDA_OK(h, cudaEventRecord(h->slot_start[sl], user));          // ordered after what the caller enqueued (e.g. the producer of the depth batch)
  ESLAM_CUDA_OK(h, cudaStreamWaitEvent(s, h->slot_start[sl], 0));
  ExtractPtrs P;
  int rc = run_extract(h, s, n_frames, depth_dev, pitch, n_obj, frame_of_obj, bbox4, label, Twc7, false, P);
  if (rc) return rc;
  ESLAM_CUDA_OK(h, cudaEventRecord(h->slot_done[sl], s));
  h->slot_objs[sl] = n_obj;
  h->next_slot = sl ^ 1;
  *ticket = sl;
  return ESLAM_OK;
}
int eslam_estimate_ellipsoids_dev_collect(EslamHandle *h, void *stream, int ticket, int n_obj, void *records_dev) {
  if (!h || !records_dev || ticket < 0 || ticket > 1 || !h->slot_objs[ticket] || n_obj != h->slot_objs[ticket]) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  cudaStream_t user = stream ? (cudaStream_t)stream : h->stream;
  ESLAM_CUDA_OK(h, cudaStreamWaitEvent(user, h->slot_done[ticket], 0));
  ESLAM_CUDA_OK(h, cudaMemcpyAsync(records_dev, h->ex_slot[ticket]->out.p, sizeof(EslamEllipsoidOut) * (size_t)n_obj, cudaMemcpyDeviceToDevice, user));
  // the slot's workspace may be reused by the next submit: its kernels run on the slot stream, which must not start before this copy
  ESLAM_CUDA_OK(h, cudaEventRecord(h->slot_start[ticket], user));
  ESLAM_CUDA_OK(h, cudaStreamWaitEvent(h->slot_stream[ticket], h->slot_start[ticket], 0));
  h->slot_objs[ticket] = 0;
  return ESLAM_OK;
}

// ---------------- initializer ----------------
int eslam_init_quadric(EslamHandle *h, int n_obs, const double *Twc7, const double *det5, const double K[9], int rows, int cols, double e10[10], int *ok) {
  if (!h || n_obs <= 0 || !Twc7 || !det5 || !K || !e10 || !ok) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return init_quadric_device(h, h->stream, n_obs, Twc7, det5, K, rows, cols, e10, ok);
}

// ---------------- graph ----------------
int eslam_edges_upload(EslamHandle *h, const EslamEdgeBatch *g, const double *ell) {
  if (!h || !g || !ell) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return graph_upload(h, g, ell, h->stream);
}
int eslam_edges_eval_dev(EslamHandle *h, void *stream) {
  if (!h) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return graph_eval(h, stream ? (cudaStream_t)stream : h->stream, true);
}
int eslam_edges_download(EslamHandle *h, double *res, double *jac, double *chi2) {
  if (!h) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return graph_download(h, res, jac, chi2, h->stream);
}
int eslam_edges_eval_host(EslamHandle *h, const EslamEdgeBatch *g, const double *ell, double *res, double *jac, double *chi2) {
  int rc = eslam_edges_upload(h, g, ell);
  if (rc) return rc;
  if ((rc = graph_eval(h, h->stream, true))) return rc;
  return graph_download(h, res, jac, chi2, h->stream);
}
int eslam_optimize_objects_host(EslamHandle *h, const EslamEdgeBatch *g, int iters, double *ell, EslamLmStats *stats) {
  int rc = eslam_edges_upload(h, g, ell);
  if (rc) return rc;
  return graph_optimize(h, iters, ell, stats, h->stream);
}

// test hook (not in eslam.h): out2 = { ordered sum as the LM kernels compute it (32 terms per step), the plain one-by-one loop }
int eslam_debug_ordered_sum(EslamHandle *h, int n, const double *v, double acc, double out2[2]) {
  if (!h || !out2) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return graph_debug_ordered_sum(h, h->stream, n, v, acc, out2);
}

int eslam_associate_frames_host(EslamHandle *h, int n_frames, const int *det_start, const double *Twc7, int n_det, const double *det_pose7_cam, const int *valid,
                                int n_map, const double *map_center3, int first_new_instance, int *assoc_out, int *n_new_out) {
  if (!h || !det_start || !Twc7 || !assoc_out || !n_new_out || (n_det > 0 && (!det_pose7_cam || !valid)) || (n_map > 0 && !map_center3)) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return associate_frames(h, h->stream, n_frames, det_start, Twc7, n_det, det_pose7_cam, valid, n_map, map_center3, first_new_instance, assoc_out, n_new_out);
}

void eslam_pose_inverse(const double T[7], double Ti[7]) {  // SE3Quat::inverse, se3quat.h:129-134
  HQ qc{-T[3], -T[4], -T[5], T[6]};
  double mt[3] = {T[0] * -1., T[1] * -1., T[2] * -1.}, o[3];
  h_rot(qc, mt, o);
  Ti[0] = o[0]; Ti[1] = o[1]; Ti[2] = o[2]; Ti[3] = qc.x; Ti[4] = qc.y; Ti[5] = qc.z; Ti[6] = qc.w;
}

void eslam_pose_mul(const double A[7], const double B[7], double out[7]) {  // SE3Quat::operator*, se3quat.h:110-116
  HQ a{A[3], A[4], A[5], A[6]}, b{B[3], B[4], B[5], B[6]};
  double o[3];
  h_rot(a, B, o);
  HQ r{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
       a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
  if (r.w < 0) { r.x *= -1; r.y *= -1; r.z *= -1; r.w *= -1; }   // normalizeRotation, se3quat.h:345-350
  double n = sqrt(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
  out[0] = A[0] + o[0]; out[1] = A[1] + o[1]; out[2] = A[2] + o[2];
  out[3] = r.x / n; out[4] = r.y / n; out[5] = r.z / n; out[6] = r.w / n;
}

int eslam_optimize_slam_host(EslamHandle *h, const EslamEdgeBatch *g, const double *odom_meas7, double odom_info, int iters, double *ell, double *Tcw_out7,
                             EslamLmStats *stats) {
  if (!h || !g || !ell) return ESLAM_ERR_BAD_ARG;
  cudaSetDevice(h->device);
  return slam_optimize(h, g, odom_meas7, odom_info, iters, ell, Tcw_out7, stats, h->stream);
}

}  // extern "C"

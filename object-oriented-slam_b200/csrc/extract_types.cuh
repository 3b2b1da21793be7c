// extract_types.cuh -- per-object metadata / state / workspace layout of the ellipsoid-extraction kernels.
#pragma once
#include "common.cuh"

namespace eslam {

struct ObjMeta {  // built on the host per batch (tiny), uploaded once
  int frame, label, symtype, pre_status;
  int x1, y1, x2, y2;   // int-truncated bbox (PointCloudFilter.cpp:26-29)
  int nx, ny;           // number of stride-3 samples per axis
  int off, cap;         // slice [off, off+cap) of every per-point workspace array
  double bbox[4];
  double Twc[7];
};

struct ObjState {  // device resident, one per object
  int status;
  int n_crop, n_voxel, n_plane, n_clusters, n_cluster, n_sym, n_final;
  int run_sym, symtype, best_seed, done;   // done: #seeds finished (last arriver finalises)
  int bbox_i[4];            // (int)bbox[k] as used by isInRange (SymmetrySolver.cpp:75)
  double center[3];
  double pca_center[3], pca_rot[9], pca_eval[3];
  double T_wo[7];           // object frame before completion
  double T_oc[7];           // camera pose in the object-normalised frame
  double P[12];             // K[I|0] hom(T_oc^-1)
  double seed_prob[9], seed_init[9], seed_final[9], seed_plane[9][4], seed_dual[9];
  long long phase_clk[9];   // clock64() stamps of the extraction phases (thread 0): start, crop, voxel sort, voxel+plane filter, cluster sort,
                            // neighbour search, flatten+select+gather, PCA, symmetry grid (tools/extract_phases.py)
  long long sym_clk[4];     // single-launch symmetry kernel, cycles of the CTA: scalar phases, task phases, scan phases, whole CTA (round-1 kernel: per-warp sums)
  double sym_R2;            // radius of the object-normalised cloud (bounds the finite-difference displacement), round pipeline
  unsigned long long n_evals, n_queries, n_points, n_hits;   // symmetry work counters (cost evaluations, NN scans, mirrored points, NN cache hits)
  unsigned long long n_ref_evals, n_ref_points;              // what the reference's call sequence performs (incl. the evaluations memoised here)
  unsigned long long n_verify_fail;                          // ESLAM_SYM_VERIFY=1: shortcut decisions that disagreed with the full path (must stay 0)
};

struct ExtractParamsDev {
  double fx, fy, cx, cy, scale, depth_range;
  int rows, cols;
  float voxel_leaf, sym_leaf;
  double plane[4];
  double plane_margin;
  float cluster_r2;          // (float)(tol*tol)
  float cluster_cell;        // tol * 1.01
  int min_cluster;
  double center_dis;
  double sigma;
  int symmetry_open, symmetry_iters;
  double K[9];
};

struct ExtractPtrs {  // device pointers of the per-point workspace (total capacity = sum of ObjMeta.cap)
  float4 *ptsA;     // S1 crop (float), later scratch
  float4 *ptsB;     // S3/S4 world points (float copy used by clustering)
  float4 *ptsC;     // S6 selected cluster (float)
  unsigned *k0, *k1, *v0, *v1;  // radix sort ping-pong
  int *parent, *count;
  double *objn;     // [cap*3] object-normalised cloud (double)
  float4 *objnf;    // float copy for the NN search
  float4 *nncache;  // [sm_count*32][2*kSymQueue] NN cache slot of a resident seed warp of the symmetry kernel (query float3 + neighbour index)
  uint4 *pixcache;  // same shape: per-point pixel cache (pixel key, range value, depth)
  unsigned *slotmask;  // [sm_count] bit b set = cache slot sm*32+b in use   (these three: round-1 kernel only, ESLAM_SYM_IMPL=1)
  // symmetry kernel caches: one slot per RESIDENT seed warp, handed out per SM through a bitmask (slot = dense_sm * 32 + bit)
  unsigned char *sympool256, *sympool512;   // slot = [2 NN banks][2*CAP] float4 + [2*CAP] uint4 pixel cache = CAP * 96 bytes
  unsigned *slotmask2;   // [2][sm_count]: CAP = 256 words, then CAP = 512 words; bits beyond the variant's slot count are pre-set
  int *smid_map;         // [kMaxSmId + 1]: dense SM index per %smid (first touch), last entry = next free index
  int sm_count;
  int uf_jump;           // clustering: pointer jumping over all parents every uf_jump-th batch of points (0 = never)
  int l2_discard;        // 1: dead per-object scratch is dropped from L2 without write-back when its last reader is done (discard.global.L2)
  // optional stage dump (device mirrors, NULL when no dump requested)
  double *d_crop, *d_voxel, *d_world, *d_sym, *d_final;  // d_final has 4x capacity
  int *d_label;
};

struct ExtractWorkspace {
  DevBuf meta, state, out, ptsA, ptsB, ptsC, k0, k1, v0, v1, parent, count, objn, objnf;
  DevBuf d_crop, d_voxel, d_world, d_sym, d_final, d_label;
  DevBuf symseed, symcache, symchunk, symlive;   // symmetry round pipeline: seed state, NN / pixel caches, per-chunk descriptors and live counters
  PinBuf pin_symchunk2[2];   // per-chunk descriptors, page-locked staging (two generations, see pin_cur)
  struct SymGraph { cudaGraphExec_t exec = nullptr; int n_obj = 0, rounds = 0, verify = 0; const void *desc = nullptr, *live = nullptr; } symgraph[16];   // replayable launch sequence per chunk
  DevBuf depth;   // staging for the *_host entry point
  PinBuf pin_meta2[2], pin_out, pin_depth;
  // The page-locked staging of the metadata / chunk descriptors is written by the CPU at issue time and read by an asynchronous copy
  // later: two generations, and a generation is rewritten only after the batch that used it has completed (pin_done) -- a caller
  // may issue the stream-based entry points back to back without synchronising in between.
  cudaEvent_t pin_done[2] = {nullptr, nullptr};
  int pin_cur = 0;
  std::vector<ObjMeta> meta_host;
  std::vector<cudaMemcpy3DBatchOp> copy_ops;   // per-chunk upload list of the host entry
  std::vector<int> stage_rows;                 // staged pull of pageable input: row-count prefix of the chunk's objects
  long long cap_total = 0;
  int n_obj = 0;
};

}  // namespace eslam

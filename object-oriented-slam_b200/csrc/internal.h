// internal.h -- C++ entry points between the translation units of libeslam_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include "../../include/eslam.h"

struct EslamHandle;
namespace eslam {
struct ExtractParamsDev;
struct ExtractPtrs;

// edges.cu
int graph_upload(EslamHandle *h, const EslamEdgeBatch *g, const double *ell, cudaStream_t s);
int graph_eval(EslamHandle *h, cudaStream_t s, bool jac);
int graph_download(EslamHandle *h, double *res, double *jac, double *chi2, cudaStream_t s);
int graph_optimize(EslamHandle *h, int iters, double *ell_inout, EslamLmStats *stats, cudaStream_t s);
void graph_free(EslamHandle *h);
int graph_debug_ordered_sum(EslamHandle *h, cudaStream_t s, int n, const double *v_host, double acc, double out2[2]);   // test hook

// slam.cu (SLAM mode: free poses + odometry edges)
int slam_optimize(EslamHandle *h, const EslamEdgeBatch *g, const double *odom_meas7, double odom_info, int iters, double *ell_inout, double *Tcw_out,
                  EslamLmStats *stats, cudaStream_t s);
void slam_free(EslamHandle *h);

// extract.cu / symmetry.cu
int extract_launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj, const ExtractParamsDev &prm,
                   const ExtractPtrs &ptrs);
int pull_rects_launch(EslamHandle *h, cudaStream_t s, const uint16_t *host_mapped, uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj);
int symmetry_finalize_launch(EslamHandle *h, cudaStream_t s, const uint16_t *depth_dev, size_t pitch, size_t frame_stride, int obj0, int n_obj,
                             const ExtractParamsDev &prm, const ExtractPtrs &ptrs, int chunk = 0);
size_t symmetry_round_seed_bytes();    // round pipeline (ESLAM_SYM_IMPL=3): persistent state per object (nine seeds)
size_t symmetry_round_cache_bytes();   // NN + pixel caches per object
size_t symmetry_round_desc_bytes();    // one chunk descriptor
constexpr int kSymMaxRounds = 126;     // live counters per chunk (rounds + 2)

size_t symmetry_pool_bytes(int cap);   // NN + pixel cache bytes of one resident symmetry CTA

// plane.cu
int plane_extract(EslamHandle *h, cudaStream_t s, int n_frames, const uint16_t *depth_dev, size_t pitch, double *plane_cam_host, int *found_host,
                  int *n_inliers_host);
int plane_extract_dbg(EslamHandle *h, cudaStream_t s, int n_frames, const uint16_t *depth_dev, size_t pitch, double *plane_cam_host, int *found_host,
                      int *n_inliers_host, int debug_stage, void *debug_out);
void plane_free(EslamHandle *h);

// associate.cu (N4: batched data association)
int associate_frames(EslamHandle *h, cudaStream_t s, int n_frames, const int *det_start_host, const double *Twc7, int n_det, const double *det_pose7, const int *valid,
                     int n_map, const double *map_center3, int first_new_instance, int *assoc_out, int *n_new_out);

// init_quadric.cu
int init_quadric_device(EslamHandle *h, cudaStream_t s, int n_obs, const double *Twc7, const double *det5, const double K[9], int rows, int cols,
                        double ellipsoid10[10], int *ok);
}  // namespace eslam

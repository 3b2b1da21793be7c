/*
 * eslam.h -- C ABI of the B200-native per-frame quadric-estimation hot path.
 *
 * This is the drop-in boundary.  The reference
 * (XunshanMan/Object-oriented-SLAM) has no FFI: Tracking.cpp calls C++ classes
 * directly.  Every entry point below names the reference interface it replaces
 * (file:line relative to the reference tree).  Signatures use only POD types,
 * plain pointers and sizes -- no Eigen / OpenCV / g2o / torch types.
 *
 * Conventions
 *   - poses are 7 doubles  x y z qx qy qz qw            (g2o::SE3Quat::toVector, se3quat.h:147)
 *   - ellipsoids are 10 doubles  pose7 + half-axes a b c  (Ellipsoid.cpp:152-157)
 *   - planes are 4 doubles  A B C D  with  Ax+By+Cz+D=0   (Plane.h:119)
 *   - depth images are row-major uint16 (CV_16UC1), `pitch` in ELEMENTS per row
 *   - every function returns 0 on success or a negative ESLAM_ERR_* code;
 *     per-object / per-frame status codes are the reference's own state codes.
 *   - *_host entry points take host pointers and do H2D/D2H themselves;
 *     *_dev entry points take device pointers for the bulk arrays (depth) and a
 *     cudaStream_t (passed as void*) and leave results in device buffers owned
 *     by the handle; small metadata (bboxes, poses, labels) is always host.
 *   - there is NO CPU fallback: if no CUDA device is usable, eslam_create fails
 *     with ESLAM_ERR_NO_DEVICE.
 */
#ifndef ESLAM_H_
#define ESLAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESLAM_ABI_VERSION 1

/* library-level error codes */
#define ESLAM_OK 0
#define ESLAM_ERR_NO_DEVICE (-1)
#define ESLAM_ERR_BAD_ARG (-2)
#define ESLAM_ERR_CUDA (-3)
#define ESLAM_ERR_CAPACITY (-4)
#define ESLAM_ERR_NO_PLANE (-5) /* supporting plane not set: EllipsoidExtractor.cpp:90 assert */

/* per-object status == EllipsoidExtractor::miSystemState (EllipsoidExtractor.h:101-102) */
#define ESLAM_OBJ_OK 0
#define ESLAM_OBJ_NO_CENTER 1    /* EllipsoidExtractor.cpp:134 */
#define ESLAM_OBJ_NO_CLUSTER 2   /* EllipsoidExtractor.cpp:146 */
#define ESLAM_OBJ_NO_POINTS 3    /* EllipsoidExtractor.cpp:127 */
#define ESLAM_OBJ_NO_POINTS_AFTER_PLANE 4 /* EllipsoidExtractor.cpp:118 */
#define ESLAM_OBJ_BBOX_OUTSIDE 5 /* not in the reference: bbox leaves the image (the reference reads out of bounds) */
#define ESLAM_OBJ_INTERNAL 6     /* not in the reference: the device pipeline did not finish the object (never expected; reported instead of a partial record) */

/* camera_intrinsic (include/core/Geometry.h:27-33) + image size */
typedef struct EslamCamera {
  double fx, fy, cx, cy;
  double scale; /* DepthMapFactor / Camera.scale, raw depth / scale = metres */
  int rows, cols;
} EslamCamera;

/* Every Config key the hot path reads, defaults = Example/param/TUM3.yaml */
typedef struct EslamParams {
  double depth_range;        /* EllipsoidExtractor_DEPTH_RANGE = 6        (Config.cpp:30) */
  double voxel_leaf;         /* 0.01 hard-coded                            (EllipsoidExtractor.cpp:98) */
  double plane_margin;       /* 0.05 hard-coded                            (EllipsoidExtractor.cpp:571) */
  double cluster_tolerance;  /* EllipsoidExtraction.Euclidean.ClusterTolerance = 0.02 */
  int min_cluster_size;      /* EllipsoidExtraction.Euclidean.MinClusterSize = 100 */
  double center_dis;         /* EllipsoidExtraction.Euclidean.CenterDis = 0.5 */
  int symmetry_open;         /* EllipsoidExtraction.Symmetry.Open = 1 */
  double symmetry_grid;      /* EllipsoidExtraction.Symmetry.GridSize = 0.1 */
  double symmetry_sigma;     /* SymmetrySolver.Sigma = 0.1 */
  int symmetry_iters;        /* 5 hard-coded                               (SymmetrySolver.cpp:263,319) */
  int plane_min_size;        /* Plane.MinSize = 200 */
  double plane_angle_deg;    /* Plane.AngleThreshold = 5   (read as float, PlaneExtractor.cpp:66) */
  double plane_dist;         /* Plane.DistanceThreshold = 0.1 (read as float, :67) */
  double edge3d_scale;       /* Optimizer.Edges.3DEllipsoid.Scale = 10000 */
  int gravity_prior_open;    /* Optimizer.Edges.GravityPrior.Open = 1 */
  double gravity_prior_scale;/* Optimizer.Edges.GravityPrior.Scale = 100 */
  int border_pixels;         /* Measurement.Border.Pixels = 10 */
  int length_limit_pixels;   /* Measurement.LengthLimit.Pixels = 0 */
} EslamParams;

/* Result of EllipsoidExtractor::EstimateLocalEllipsoid for one bbox
 * (EllipsoidExtractor.cpp:292-493) + GetResult() + GetSymmetryOutputData(). */
typedef struct EslamEllipsoidOut {
  double ellipsoid[10];   /* camera-frame ellipsoid e_local, toVector() */
  double prob;            /* e_local.prob (:482-489) */
  int status;             /* miSystemState, 0 == GetResult() true */
  int symmetry_type;      /* 0 none, 1 reflection, 2 dual reflection (:67-76) */
  int symmetry_result;    /* SymmetryOutputData.result */
  int best_seed;          /* index 0..8 of the winning init plane (Symmetry.cpp:122-126) */
  double sym_plane_world[4];  /* SymmetryOutputData.planeVec  (:387-389) */
  double sym_plane2_world[4]; /* SymmetryOutputData.planeVec2 (:391-396) */
  double sym_center_world[3]; /* SymmetryOutputData.center    (:383,424) */
  /* integer segment-mask statistics (bit-exact contract) */
  int n_crop;      /* valid stride-3 samples            (PointCloudFilter.cpp:31-52) */
  int n_voxel;     /* after 1 cm voxel grid             (EllipsoidExtractor.cpp:98) */
  int n_plane;     /* after supporting-plane filter     (:107) */
  int n_clusters;  /* clusters >= MinClusterSize        (:673) */
  int n_cluster;   /* points of the selected cluster    (:725) */
  int n_sym;       /* points after the symmetry grid    (:324) */
  int n_final;     /* points used for the extents       (:461) */
  int reserved;
} EslamEllipsoidOut;

/* Optional per-object stage dump used by the parity tests (caller-allocated,
 * capacities in points; arrays may be NULL to skip).  Same layout is filled by
 * the oracle, so each stage can be compared in isolation. */
typedef struct EslamStageDump {
  int cap_points;        /* capacity of every per-point array below */
  double *crop_xyz;      /* [n_crop*3]   S1 camera-frame points, scan order */
  double *voxel_xyz;     /* [n_voxel*3]  S2 centroids, ascending voxel key */
  double *world_xyz;     /* [n_plane*3]  S3+S4 world points kept by the plane filter */
  int *cluster_label;    /* [n_plane]    S6: -1 = no accepted cluster, else rank of the cluster (0 = largest) */
  double *sym_xyz;       /* [n_sym*3]    object-normalised cloud fed to the symmetry solver */
  double *final_xyz;     /* [n_final*3]  completed cloud in the final object frame */
  double center_world[3];/* S5 GetCenter */
  double pca_center[3];  /* S7 */
  double pca_rot[9];     /* S7+S8 rotMat after ApplyGravityPrior, row-major */
  double pca_eval[3];    /* S7 eigenvalues (after the AlignZ permutation) */
  double T_wo[7];        /* object frame before completion (:342-346) */
  double seed_prob[9];   /* S13 per-seed prob */
  double seed_init_err[9];
  double seed_final_err[9];
  double seed_plane[9][4];
  double seed_dual_dis[9];
} EslamStageDump;

/* ---- graph edges for the object BA (Optimizer.cpp:83-317, BasicEllipsoidEdges.cpp) ---- */
#define ESLAM_EDGE_PROJ 0     /* EdgeSE3EllipsoidProj        4-D  (BasicEllipsoidEdges.cpp:102-112) */
#define ESLAM_EDGE_9DOF 1     /* EdgeSE3Ellipsoid9DOF        9-D  (:68-77) */
#define ESLAM_EDGE_GRAVITY 2  /* EdgeEllipsoidGravityPlanePrior 1-D (:129-152) */
#define ESLAM_EDGE_MAXDIM 9

/* Structure-of-arrays edge batch.  Edge k connects camera pose `frame[k]`
 * (ignored for gravity edges) with ellipsoid vertex `object[k]`.  Order of the
 * edges is the g2o insertion order (internalId, optimizable_graph.cpp:273). */
typedef struct EslamEdgeBatch {
  int n_edges;
  const int *type;      /* [n_edges] ESLAM_EDGE_* */
  const int *frame;     /* [n_edges] index into Tcw */
  const int *object;    /* [n_edges] index into ellipsoids */
  const double *meas;   /* [n_edges*10] bbox4 | ellipsoid10 (camera frame) | plane4 ; unused tail ignored */
  const double *info;   /* [n_edges] scalar s: information = s * I  (Optimizer.cpp:189-193,219-223,269-274) */
  int n_frames;
  const double *Tcw;    /* [n_frames*7] world->camera (VertexSE3Expmap estimate), all fixed (mapping mode) */
  int n_objects;
  double K[9];          /* calibration, row-major */
} EslamEdgeBatch;

typedef struct EslamLmStats {
  int iterations;        /* solve() calls performed */
  int trials_total;      /* sum of LM trials (qmax) */
  int terminated;        /* 1 if a solve() returned Terminate */
  double chi2_initial;
  double chi2_final;
  double lambda_final;
} EslamLmStats;

typedef struct EslamHandle EslamHandle;

/* ------------------------------------------------------------------ lifecycle */
int eslam_abi_version(void);
void eslam_default_params(EslamParams *p);           /* TUM3.yaml values */
void eslam_default_camera(EslamCamera *c);           /* TUM3.yaml:62-70 */
/* replaces: new PlaneExtractor/EllipsoidExtractor/Optimizer/Initializer (Tracking.cpp:121,643-650,692-699) */
int eslam_create(const EslamParams *p, const EslamCamera *c, int device, EslamHandle **out);
void eslam_destroy(EslamHandle *h);
const char *eslam_last_error(EslamHandle *h);
/* number of kernel launches issued by this handle so far (bench.py "gpu_launches") */
long long eslam_launch_count(EslamHandle *h);
/* depth bytes the *_host extraction entry has uploaded so far (running total; bench.py "h2d_bytes_per_step"): only the union
 * rectangle of each frame's boxes travels -- the path never reads a pixel outside a box (PointCloudFilter.cpp:31-52,
 * EllipsoidExtractor.cpp:583-643, SymmetrySolver.cpp:75-84) */
unsigned long long eslam_h2d_bytes(EslamHandle *h);

/* profiling hooks (bench.py): device time of each kernel of the most recent extraction batch, measured with CUDA
 * events on the launching stream.  ms[0] = extract kernel, ms[1] = symmetry+finalize kernel, ms[2] = both. */
int eslam_set_profiling(EslamHandle *h, int on);
int eslam_get_stage_ms(EslamHandle *h, float ms[3]);

/* ------------------------------------------------------------------ ground plane */
/* replaces PlaneExtractor::extractGroundPlane (PlaneExtractor.cpp:126-183), batched over frames.
 * plane_cam: [n_frames*4] camera-frame plane (oriented so the camera is on the + side, :176-178)
 * found:     [n_frames]   1 = true, 0 = false (the bool result)
 * n_inliers: [n_frames]   inlier count of the chosen region (may be NULL) */
int eslam_extract_ground_plane_host(EslamHandle *h, int n_frames, const uint16_t *depth, size_t pitch,
                                    double *plane_cam, int *found, int *n_inliers);
int eslam_extract_ground_plane_dev(EslamHandle *h, void *stream, int n_frames, const uint16_t *depth_dev,
                                   size_t pitch, double *plane_cam_host, int *found_host, int *n_inliers_host);
/* g2o::plane::transform (Plane.cpp:117-122): camera-frame plane -> world with Twc */
void eslam_plane_to_world(const double plane_cam[4], const double Twc[7], double plane_world[4]);
/* replaces EllipsoidExtractor::SetSupportingPlane (EllipsoidExtractor.cpp:754-758) + Optimizer::SetGroundPlane (Optimizer.cpp:324-327) */
int eslam_set_supporting_plane(EslamHandle *h, const double plane_world[4]);

/* ------------------------------------------------------------------ ellipsoid extraction */
/* utils calibrateMeasurement (dataprocess_utils.cpp:150-190): returns 1 if the bbox is on the border / too small */
int eslam_calibrate_measurement(const EslamParams *p, const EslamCamera *c, const double bbox[4]);

/* replaces EllipsoidExtractor::EstimateLocalEllipsoid (EllipsoidExtractor.cpp:292-493), batched:
 *   n_frames depth images (frame f at depth + f*rows*pitch), n_obj boxes;
 *   frame_of_obj[i] says which frame box i belongs to; bbox4 [n_obj*4] x1 y1 x2 y2; label [n_obj];
 *   Twc7 [n_frames*7].  out [n_obj].  dump: NULL or [n_obj] stage dumps. */
int eslam_estimate_ellipsoids_host(EslamHandle *h, int n_frames, const uint16_t *depth, size_t pitch,
                                   int n_obj, const int *frame_of_obj, const double *bbox4, const int *label,
                                   const double *Twc7, EslamEllipsoidOut *out, EslamStageDump *dump);
/* same, depth already resident in HBM; results stay on the device until eslam_fetch_ellipsoids */
int eslam_estimate_ellipsoids_dev(EslamHandle *h, void *stream, int n_frames, const uint16_t *depth_dev,
                                  size_t pitch, int n_obj, const int *frame_of_obj, const double *bbox4,
                                  const int *label, const double *Twc7);
int eslam_fetch_ellipsoids(EslamHandle *h, void *stream, int n_obj, EslamEllipsoidOut *out);
/* device-to-device form of the fetch (multi-GPU "gather of map"): copies the n_obj records of the last
 * eslam_estimate_ellipsoids_dev batch into caller-owned DEVICE memory on `stream` (no synchronisation, no host round trip), so a
 * collective (ncclAllGather) can be enqueued on the same stream right behind it. */
int eslam_copy_ellipsoids_dev(EslamHandle *h, void *stream, int n_obj, void *records_dev);
/* work statistics of the last extraction batch, summed over its objects (bench.py roofline: the algorithmic bytes of the symmetry
 * kernel are counted, not assumed).  out[0] cost evaluations the reference performs (GetPointCloudProb calls, SymmetrySolver.cpp:49),
 * out[1] mirrored points those evaluations visit, out[2] evaluations this library actually ran, out[3] points it visited,
 * out[4] nearest-neighbour scans, out[5] NN cache hits, out[6] largest n_sym of the batch, out[7] objects with symmetry,
 * out[8] shortcut decisions that disagreed with the full path (only counted under ESLAM_SYM_VERIFY=1; must be 0),
 * out[9..11] clock64 cycles of the seed warps in single-plane evaluations / NN scans / scalar LM code. */
int eslam_batch_counters(EslamHandle *h, unsigned long long out[12]);
/* pipelined form of eslam_estimate_ellipsoids_host for frame streams (the reference calls EstimateLocalEllipsoid once per box and
 * blocks, Tracking.cpp:329; a batched caller keeps the device busy by submitting batch k+1 before collecting batch k).  At most two
 * tickets may be outstanding; `depth` and the metadata arrays of a submitted batch must stay valid until its wait returns; the
 * synchronous entry points refuse to run while a ticket is outstanding. */
int eslam_estimate_ellipsoids_submit(EslamHandle *h, int n_frames, const uint16_t *depth, size_t pitch,
                                     int n_obj, const int *frame_of_obj, const double *bbox4, const int *label,
                                     const double *Twc7, int *ticket);
int eslam_estimate_ellipsoids_wait(EslamHandle *h, int ticket, int n_obj, EslamEllipsoidOut *out);
/* the same pipelining for DEVICE-resident depth and records (no host copies, no host synchronisation): submit issues the batch on an
 * internal stream ordered after the work already enqueued on `stream` (the depth batch must stay valid until the matching collect);
 * collect makes `stream` wait for the batch and copies its n_obj records into caller-owned device memory on `stream`.  Tickets are
 * shared with the host form: at most two outstanding. */
int eslam_estimate_ellipsoids_dev_submit(EslamHandle *h, void *stream, int n_frames, const uint16_t *depth_dev, size_t pitch,
                                         int n_obj, const int *frame_of_obj, const double *bbox4, const int *label,
                                         const double *Twc7, int *ticket);
int eslam_estimate_ellipsoids_dev_collect(EslamHandle *h, void *stream, int ticket, int n_obj, void *records_dev);

/* ------------------------------------------------------------------ quadric initialisation */
/* replaces Initializer::initializeQuadric (Initializer.cpp:24-56): n_obs rows of Twc7 / det5 (x1 y1 x2 y2 rate)
 * ellipsoid10 out; *ok = getInitializeResult() */
int eslam_init_quadric(EslamHandle *h, int n_obs, const double *Twc7, const double *det5, const double K[9],
                       int rows, int cols, double ellipsoid10[10], int *ok);

/* ------------------------------------------------------------------ automatic data association (batched over frames) */
/* replaces DataAssociationSolver::Solve(pFrame, true) + SolveCostMat (src/core/DataAssociation.cpp:16-135) for n_frames frames
 * against ONE map snapshot: det_start [n_frames + 1] = row ranges of the frames (rows in frame order), Twc7 [n_frames*7],
 * det_pose7_cam [n_det*7] = pose of each single-frame ellipsoid in its camera frame (pFrame->mpLocalObjects[i]->pose),
 * valid [n_det] = 0 where mpLocalObjects[i] == NULL, map_center3 [n_map*3] = centres of the map objects in ascending instance
 * order (std::map order).  assoc_out [n_det]: -1 for an invalid row; the COLUMN index of the matched map object (the reference
 * returns the column, DataAssociation.cpp:113-114); or a new instance id, numbered from first_new_instance in (frame, row)
 * order (miInstanceNum++, :137-140).  *n_new_out = instances created.  At most 64 rows per frame and 1024 map objects
 * (ESLAM_ERR_CAPACITY beyond). */
int eslam_associate_frames_host(EslamHandle *h, int n_frames, const int *det_start, const double *Twc7, int n_det,
                                const double *det_pose7_cam, const int *valid, int n_map, const double *map_center3,
                                int first_new_instance, int *assoc_out, int *n_new_out);

/* ------------------------------------------------------------------ graph edges / BA */
/* One launch: residuals + g2o numeric Jacobians (delta 1e-9 central differences, base_binary_edge.hpp:131-205)
 * w.r.t. the 9-DoF ellipsoid vertex for all edges.
 *   ellipsoids [n_objects*10]; residuals [n_edges*9] (unused tail 0); jacobians [n_edges*9*9] row-major (dim x 9) */
int eslam_edges_eval_host(EslamHandle *h, const EslamEdgeBatch *g, const double *ellipsoids10,
                          double *residuals, double *jacobians, double *chi2);
/* device-resident variant for timing: upload once, evaluate `repeat` times, no D2H */
int eslam_edges_upload(EslamHandle *h, const EslamEdgeBatch *g, const double *ellipsoids10);
int eslam_edges_eval_dev(EslamHandle *h, void *stream);
int eslam_edges_download(EslamHandle *h, double *residuals, double *jacobians, double *chi2);

/* replaces Optimizer::GlobalObjectGraphOptimization's graph.optimize(iters) (Optimizer.cpp:290-291) with g2o's
 * Levenberg semantics (optimization_algorithm_levenberg.cpp:61-164); ellipsoids10 updated in place. */
int eslam_optimize_objects_host(EslamHandle *h, const EslamEdgeBatch *g, int iters, double *ellipsoids10,
                                EslamLmStats *stats);

/* SLAM mode of the same optimisation: Optimizer.cpp:126-159 with bSLAM_mode = true.  Every pose
 * g->Tcw[f], f >= 1, becomes a free VertexSE3Expmap (frame 0 stays fixed, :137), consecutive frames are tied by
 * EdgeSE3Expmap odometry edges (types_six_dof_expmap.h:83-99) inserted BEFORE the ellipsoid edges (they are created in
 * the pose loop), information = odom_info * I6 (:149-152, 1.0 in the reference).
 *   odom_meas7 [(n_frames-1)*7]: measurement of edge k between frames k and k+1 (Tcw_{k+1} * Tcw_k^-1, :141-143);
 *                                NULL = computed from g->Tcw like the reference does.
 *   ellipsoids10 updated in place; Tcw_out7 [n_frames*7] receives the optimised poses (may be NULL: the reference
 *   itself only writes the ellipsoids back, :293-306).
 * The linear system is solved exactly (block-tridiagonal pose chain + dense reduced object system) instead of the
 * reference's dense LDLT of the full matrix (linear_solver_dense.h:65-113). */
int eslam_optimize_slam_host(EslamHandle *h, const EslamEdgeBatch *g, const double *odom_meas7, double odom_info, int iters,
                             double *ellipsoids10, double *Tcw_out7, EslamLmStats *stats);

/* ------------------------------------------------------------------ Frame helper */
/* Frame::Frame pose handling (Frame.cpp:13-14): Twc7 -> Tcw7 = Twc.inverse() */
void eslam_pose_inverse(const double T7[7], double Tinv7[7]);
/* g2o::SE3Quat::operator* (se3quat.h:110-116, incl. normalizeRotation): out = A * B */
void eslam_pose_mul(const double A7[7], const double B7[7], double out7[7]);

#ifdef __cplusplus
}
#endif
#endif /* ESLAM_H_ */

// eslam_adapter.hpp -- C++ host-side mirror of the reference's operator classes over the C-ABI (include/eslam.h).
//
// The classes keep the reference's names and call shapes so that src/core/Tracking.cpp can be pointed at them:
//   EllipsoidSLAM::PlaneExtractor     src/plane/PlaneExtractor.h:41-70
//   EllipsoidSLAM::EllipsoidExtractor src/pca/EllipsoidExtractor.h:37-131
//   EllipsoidSLAM::Initializer        include/core/Initializer.h:36-80
//   EllipsoidSLAM::Optimizer          include/core/Optimizer.h:14-32
//   EllipsoidSLAM::Frame              include/core/Frame.h:12-41
// Two layers: (1) POD signatures (always available -- this is what is compiled and link-checked in this repository, which has no
// Eigen/OpenCV); (2) Eigen / cv::Mat overloads with the reference's exact signatures, enabled when those headers exist
// (compile-verified elsewhere).  No compute happens here: every method forwards to libeslam_b200.so (CUDA only, no CPU fallback).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/eslam.h"

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && __has_include(<opencv2/core.hpp>)
#define ESLAM_HAVE_EIGEN_OPENCV 1
#include <Eigen/Core>
#include <opencv2/core.hpp>
#endif
#endif

namespace EllipsoidSLAM {

using Pose7 = std::array<double, 7>;        // x y z qx qy qz qw
using Ellipsoid10 = std::array<double, 10>; // pose7 + a b c
using Plane4 = std::array<double, 4>;

struct camera_intrinsic { double fx, fy, cx, cy, scale; };   // include/core/Geometry.h:27-33

struct ellipsoid_pod {   // the fields of g2o::ellipsoid the path reads/writes (include/core/Ellipsoid.h:26-136)
  Ellipsoid10 vec{};
  double prob = 1.0;
  int miLabel = -1, miInstanceID = -1;
};

struct SymmetryOutputData {  // src/symmetry/Symmetry.h
  bool result = false; double prob = 0; int symmetryType = 0; Plane4 planeVec{}, planeVec2{}; std::array<double, 3> center{};
};

class Context {   // owns the EslamHandle shared by the extractor / optimizer objects (one per GPU)
 public:
  Context(const EslamParams &p, const EslamCamera &c, int device = 0) {
    if (eslam_create(&p, &c, device, &h_) != ESLAM_OK) throw std::runtime_error("eslam_create failed: a CUDA device is required (no CPU fallback)");
    cam_ = c; prm_ = p;
  }
  ~Context() { eslam_destroy(h_); }
  Context(const Context &) = delete;
  EslamHandle *get() const { return h_; }
  const EslamCamera &camera() const { return cam_; }
  const EslamParams &params() const { return prm_; }
  void check(int rc, const char *what) const { if (rc != ESLAM_OK) throw std::runtime_error(std::string(what) + ": " + eslam_last_error(h_)); }
 private:
  EslamHandle *h_ = nullptr; EslamCamera cam_{}; EslamParams prm_{};
};

class Frame {   // src/core/Frame.cpp:7-27
 public:
  static int total_frame;
  Frame(double timestamp_, const double *pose_tail7, const std::vector<std::array<double, 8>> &bboxMat, const uint16_t *depth, int rows, int cols)
      : timestamp(timestamp_), mmObservations(bboxMat), frame_img(depth, depth + (size_t)rows * cols) {
    for (int i = 0; i < 7; i++) cam_pose_Twc[i] = pose_tail7[i];
    eslam_pose_inverse(cam_pose_Twc.data(), cam_pose_Tcw.data());
    frame_seq_id = total_frame++;
    mvbOutlier.assign(bboxMat.size(), false);
  }
#ifdef ESLAM_HAVE_EIGEN_OPENCV
  // Frame(double timestamp, const Eigen::VectorXd &pose, const Eigen::MatrixXd &bboxMap, const cv::Mat &imDepth, const cv::Mat &imRGB)
  // (include/core/Frame.h:18-20; the last 7 entries of `pose` are x y z qx qy qz qw, Frame.cpp:13; the RGB image is only drawn on)
  Frame(double timestamp_, const Eigen::VectorXd &pose, const Eigen::MatrixXd &bboxMap, const cv::Mat &imDepth, const cv::Mat & /*imRGB*/ = cv::Mat())
      : timestamp(timestamp_) {
    for (int i = 0; i < 7; i++) cam_pose_Twc[i] = pose[pose.size() - 7 + i];
    eslam_pose_inverse(cam_pose_Twc.data(), cam_pose_Tcw.data());
    frame_seq_id = total_frame++;
    for (int r = 0; r < bboxMap.rows(); r++) { std::array<double, 8> row{}; for (int k = 0; k < 8 && k < bboxMap.cols(); k++) row[k] = bboxMap(r, k); mmObservations.push_back(row); }
    mvbOutlier.assign(mmObservations.size(), false);
    frame_img.resize((size_t)imDepth.rows * imDepth.cols);
    for (int r = 0; r < imDepth.rows; r++) std::copy(imDepth.ptr<uint16_t>(r), imDepth.ptr<uint16_t>(r) + imDepth.cols, frame_img.begin() + (size_t)r * imDepth.cols);
  }
  Frame(const Eigen::VectorXd &pose, const Eigen::MatrixXd &bboxMap, const cv::Mat &imDepth, const cv::Mat &imRGB = cv::Mat()) : Frame(0.0, pose, bboxMap, imDepth, imRGB) {}
#endif
  double timestamp; int frame_seq_id;
  Pose7 cam_pose_Twc{}, cam_pose_Tcw{};
  std::vector<std::array<double, 8>> mmObservations;   // id x1 y1 x2 y2 label rate instanceID
  std::vector<uint16_t> frame_img;
  std::vector<bool> mvbOutlier;
  std::vector<ellipsoid_pod *> mpLocalObjects;
};
inline int Frame::total_frame = 0;

class PlaneExtractor {
 public:
  explicit PlaneExtractor(Context &c) : ctx_(c) {}
  // bool extractGroundPlane(const cv::Mat &depth, g2o::plane &plane)   (PlaneExtractor.cpp:126-183)
  bool extractGroundPlane(const uint16_t *depth, size_t pitch, Plane4 &plane) {
    int found = 0, inl = 0;
    ctx_.check(eslam_extract_ground_plane_host(ctx_.get(), 1, depth, pitch, plane.data(), &found, &inl), "extractGroundPlane");
    n_inliers_ = inl;
    return found != 0;
  }
  int inliers() const { return n_inliers_; }
#ifdef ESLAM_HAVE_EIGEN_OPENCV
  bool extractGroundPlane(const cv::Mat &depth, Eigen::Vector4d &plane_param) {
    Plane4 p; bool ok = extractGroundPlane(depth.ptr<uint16_t>(0), depth.step / sizeof(uint16_t), p);
    for (int i = 0; i < 4; i++) plane_param[i] = p[i];
    return ok;
  }
#endif
 private:
  Context &ctx_; int n_inliers_ = 0;
};

class EllipsoidExtractor {
 public:
  explicit EllipsoidExtractor(Context &c) : ctx_(c) {}
  void OpenSymmetry() { if (!ctx_.params().symmetry_open) throw std::runtime_error("symmetry is a creation-time parameter (EslamParams.symmetry_open)"); }
  void SetSupportingPlane(const Plane4 &plane_world) { ctx_.check(eslam_set_supporting_plane(ctx_.get(), plane_world.data()), "SetSupportingPlane"); }
  // g2o::ellipsoid EstimateLocalEllipsoid(cv::Mat& depth, Vector4d& bbox, int label, VectorXd& pose, camera_intrinsic&)  (EllipsoidExtractor.cpp:292-493)
  ellipsoid_pod EstimateLocalEllipsoid(const uint16_t *depth, size_t pitch, const std::array<double, 4> &bbox, int label, const Pose7 &pose) {
    const int frame_of_obj = 0;
    EslamEllipsoidOut o;
    ctx_.check(eslam_estimate_ellipsoids_host(ctx_.get(), 1, depth, pitch, 1, &frame_of_obj, bbox.data(), &label, pose.data(), &o, nullptr), "EstimateLocalEllipsoid");
    miSystemState = o.status; mResult = (o.status == ESLAM_OBJ_OK);
    mSym.result = o.symmetry_result != 0; mSym.prob = o.prob; mSym.symmetryType = o.symmetry_type;
    for (int i = 0; i < 4; i++) { mSym.planeVec[i] = o.sym_plane_world[i]; mSym.planeVec2[i] = o.sym_plane2_world[i]; }
    for (int i = 0; i < 3; i++) mSym.center[i] = o.sym_center_world[i];
    ellipsoid_pod e; for (int i = 0; i < 10; i++) e.vec[i] = o.ellipsoid[i];
    e.prob = o.prob; e.miLabel = label;
    return e;
  }
  // batched form: every bbox of a batch of frames in one call
  std::vector<EslamEllipsoidOut> EstimateLocalEllipsoids(int n_frames, const uint16_t *depth, size_t pitch, const std::vector<int> &frame_of_obj,
                                                         const std::vector<double> &bbox4, const std::vector<int> &label, const std::vector<double> &Twc7) {
    std::vector<EslamEllipsoidOut> out(frame_of_obj.size());
    ctx_.check(eslam_estimate_ellipsoids_host(ctx_.get(), n_frames, depth, pitch, (int)frame_of_obj.size(), frame_of_obj.data(), bbox4.data(), label.data(), Twc7.data(),
                                              out.data(), nullptr), "EstimateLocalEllipsoids");
    return out;
  }
  // frame streams: submit batch k+1 before collecting batch k (at most two tickets outstanding; the input arrays of a submitted
  // batch must stay valid until its Collect returns) -- eslam_estimate_ellipsoids_submit / _wait
  int SubmitLocalEllipsoids(int n_frames, const uint16_t *depth, size_t pitch, const std::vector<int> &frame_of_obj, const std::vector<double> &bbox4,
                            const std::vector<int> &label, const std::vector<double> &Twc7) {
    int ticket = -1;
    ctx_.check(eslam_estimate_ellipsoids_submit(ctx_.get(), n_frames, depth, pitch, (int)frame_of_obj.size(), frame_of_obj.data(), bbox4.data(), label.data(), Twc7.data(),
                                                &ticket), "SubmitLocalEllipsoids");
    return ticket;
  }
  std::vector<EslamEllipsoidOut> CollectLocalEllipsoids(int ticket, size_t n_obj) {
    std::vector<EslamEllipsoidOut> out(n_obj);
    ctx_.check(eslam_estimate_ellipsoids_wait(ctx_.get(), ticket, (int)n_obj, out.data()), "CollectLocalEllipsoids");
    return out;
  }
#ifdef ESLAM_HAVE_EIGEN_OPENCV
  // the reference's exact call (src/pca/EllipsoidExtractor.h:57): depth CV_16UC1, bbox x1 y1 x2 y2, pose with x y z qx qy qz qw in
  // its last 7 entries; `camera` must be the intrinsics the Context was created with (checked).  to_reference<g2o::ellipsoid>() below
  // turns the result into the reference's own type.
  ellipsoid_pod EstimateLocalEllipsoid(cv::Mat &depth, Eigen::Vector4d &bbox, int label, Eigen::VectorXd &pose, camera_intrinsic &camera) {
    const EslamCamera &c = ctx_.camera();
    if (camera.fx != c.fx || camera.fy != c.fy || camera.cx != c.cx || camera.cy != c.cy || camera.scale != c.scale || depth.rows != c.rows || depth.cols != c.cols)
      throw std::runtime_error("EstimateLocalEllipsoid: camera / image size differ from the Context's (intrinsics are creation-time parameters)");
    Pose7 T; for (int i = 0; i < 7; i++) T[i] = pose[pose.size() - 7 + i];
    return EstimateLocalEllipsoid(depth.ptr<uint16_t>(0), depth.step / sizeof(uint16_t), {bbox[0], bbox[1], bbox[2], bbox[3]}, label, T);
  }
  // void SetSupportingPlane(g2o::plane *pPlane)  (EllipsoidExtractor.h:54): any type with a 4-vector member `param`
  template <class PlaneT> void SetSupportingPlane(PlaneT *pPlane) { SetSupportingPlane(Plane4{pPlane->param[0], pPlane->param[1], pPlane->param[2], pPlane->param[3]}); }
#endif
  bool GetResult() const { return mResult; }
  SymmetryOutputData GetSymmetryOutputData() const { return mSym; }
  int miSystemState = 0;
 private:
  Context &ctx_; bool mResult = false; SymmetryOutputData mSym;
};

struct Observation { int label; std::array<double, 4> bbox; double rate; Frame *pFrame; };   // include/core/Initializer.h:14-24
using Observations = std::vector<Observation *>;

class Initializer {
 public:
  Initializer(Context &c, int rows, int cols) : ctx_(c), rows_(rows), cols_(cols) {}
  // g2o::ellipsoid initializeQuadric(Observations&, Matrix3d& calib)   (Initializer.cpp:290-305)
  ellipsoid_pod initializeQuadric(const Observations &obs, const double K[9]) {
    std::vector<double> tw, det;
    for (const Observation *o : obs) { tw.insert(tw.end(), o->pFrame->cam_pose_Twc.begin(), o->pFrame->cam_pose_Twc.end()); det.insert(det.end(), o->bbox.begin(), o->bbox.end()); det.push_back(o->rate); }
    ellipsoid_pod e; int ok = 0;
    ctx_.check(eslam_init_quadric(ctx_.get(), (int)obs.size(), tw.data(), det.data(), K, rows_, cols_, e.vec.data(), &ok), "initializeQuadric");
    mbResult = ok != 0;
    if (mbResult && !obs.empty()) e.miLabel = obs[0]->label;
    return e;
  }
#ifdef ESLAM_HAVE_EIGEN_OPENCV
  // g2o::ellipsoid initializeQuadric(MatrixXd &pose_mat, MatrixXd &detection_mat, Matrix3d &calib)   (include/core/Initializer.h:46)
  // pose_mat rows: x y z qx qy qz qw (last 7 columns), detection_mat rows: x1 y1 x2 y2 (rate)
  ellipsoid_pod initializeQuadric(Eigen::MatrixXd &pose_mat, Eigen::MatrixXd &detection_mat, Eigen::Matrix3d &calib) {
    std::vector<double> tw, det; double K[9];
    for (int r = 0; r < pose_mat.rows(); r++) {
      for (int k = 0; k < 7; k++) tw.push_back(pose_mat(r, pose_mat.cols() - 7 + k));
      for (int k = 0; k < 4; k++) det.push_back(detection_mat(r, k));
      det.push_back(detection_mat.cols() > 4 ? detection_mat(r, 4) : 1.0);
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) K[3 * i + j] = calib(i, j);
    ellipsoid_pod e; int ok = 0;
    ctx_.check(eslam_init_quadric(ctx_.get(), (int)pose_mat.rows(), tw.data(), det.data(), K, rows_, cols_, e.vec.data(), &ok), "initializeQuadric");
    mbResult = ok != 0;
    return e;
  }
  // g2o::ellipsoid initializeQuadric(Observations &obs, Matrix3d &calib)   (include/core/Initializer.h:47) for the REFERENCE's own
  // Observation / Frame types: obs[i]->label, ->bbox (Vector4d), ->rate, ->pFrame->cam_pose_Twc (g2o::SE3Quat, toVector())
  template <class ObsVec> ellipsoid_pod initializeQuadric(ObsVec &obs, Eigen::Matrix3d &calib) {
    Eigen::MatrixXd pose_mat((int)obs.size(), 7), det_mat((int)obs.size(), 5);
    int r = 0;
    for (auto *o : obs) {
      const auto v = o->pFrame->cam_pose_Twc.toVector();
      for (int k = 0; k < 7; k++) pose_mat(r, k) = v[k];
      for (int k = 0; k < 4; k++) det_mat(r, k) = o->bbox[k];
      det_mat(r, 4) = o->rate; r++;
    }
    ellipsoid_pod e = initializeQuadric(pose_mat, det_mat, calib);
    if (mbResult && obs.size() > 0) e.miLabel = (*obs.begin())->label;
    return e;
  }
#endif
  bool getInitializeResult() const { return mbResult; }
 private:
  Context &ctx_; int rows_, cols_; bool mbResult = false;
};

class Optimizer {
 public:
  explicit Optimizer(Context &c) : ctx_(c) {}
  void SetGroundPlane(const Plane4 &normal) { mbGroundPlaneSet = true; mGroundPlaneNormal = normal; }
  // void GlobalObjectGraphOptimization(std::vector<Frame*>&, Map*, rows, cols, Matrix3d& K, std::map<int,Observations>&, ...)  (Optimizer.cpp:83-317)
  // `ellipsoids` plays the role of Map::GetAllEllipsoidsMap(): instance id -> ellipsoid, updated in place.
  EslamLmStats GlobalObjectGraphOptimization(const std::vector<Frame *> &pFrames, std::map<int, ellipsoid_pod *> &ellipsoids, const double K[9],
                                             const std::map<int, Observations> &objectObservations, int iters = 10) {
    const EslamParams &prm = ctx_.params();
    std::vector<int> inst; std::map<int, int> vid;
    for (const auto &kv : objectObservations) if (ellipsoids.count(kv.first)) { vid[kv.first] = (int)inst.size(); inst.push_back(kv.first); }
    std::vector<int> type, frame, object; std::vector<double> meas, info, Tcw, ell;
    auto push_meas = [&](const double *m, int n) { for (int i = 0; i < 10; i++) meas.push_back(i < n ? m[i] : 0.0); };
    for (int id : inst) {
      if (mbGroundPlaneSet && prm.gravity_prior_open == 1) {   // :183-196
        type.push_back(ESLAM_EDGE_GRAVITY); frame.push_back(0); object.push_back(vid[id]); push_meas(mGroundPlaneNormal.data(), 4);
        const double s = 1 * prm.gravity_prior_scale; info.push_back(s * s);
      }
      const Observations &obs = objectObservations.at(id);
      if (obs.size() > 2) for (const Observation *o : obs) {     // :201-245 (the NaN gate is applied below)
        type.push_back(ESLAM_EDGE_PROJ); frame.push_back(o->pFrame->frame_seq_id); object.push_back(vid[id]); push_meas(o->bbox.data(), 4); info.push_back((1.0 * 1.0) * o->rate);
      }
    }
    for (size_t f = 0; f < pFrames.size(); f++) for (const ellipsoid_pod *obj : pFrames[f]->mpLocalObjects) {   // :249-279
      if (!obj || !vid.count(obj->miInstanceID)) continue;
      type.push_back(ESLAM_EDGE_9DOF); frame.push_back((int)f); object.push_back(vid[obj->miInstanceID]); push_meas(obj->vec.data(), 10);
      const double s = 1.0 * std::sqrt(obj->prob); info.push_back((s * s) * prm.edge3d_scale);
    }
    for (const Frame *f : pFrames) Tcw.insert(Tcw.end(), f->cam_pose_Tcw.begin(), f->cam_pose_Tcw.end());
    for (int id : inst) ell.insert(ell.end(), ellipsoids[id]->vec.begin(), ellipsoids[id]->vec.end());
    EslamLmStats st{};
    if (inst.empty() || type.empty()) return st;
    auto make_batch = [&](EslamEdgeBatch &g) {
      g.n_edges = (int)type.size(); g.type = type.data(); g.frame = frame.data(); g.object = object.data(); g.meas = meas.data(); g.info = info.data();
      g.n_frames = (int)pFrames.size(); g.Tcw = Tcw.data(); g.n_objects = (int)inst.size(); for (int i = 0; i < 9; i++) g.K[i] = K[i];
    };
    EslamEdgeBatch g; make_batch(g);
    {  // NaN gate of the 2-D edges before insertion (:234-243)
      std::vector<double> res((size_t)g.n_edges * 9), chi((size_t)g.n_edges);
      ctx_.check(eslam_edges_eval_host(ctx_.get(), &g, ell.data(), res.data(), nullptr, chi.data()), "edges_eval");
      size_t w = 0;
      for (size_t k = 0; k < type.size(); k++) {
        if (type[k] == ESLAM_EDGE_PROJ && chi[k] != chi[k]) continue;
        type[w] = type[k]; frame[w] = frame[k]; object[w] = object[k]; info[w] = info[k]; for (int i = 0; i < 10; i++) meas[w * 10 + i] = meas[k * 10 + i]; w++;
      }
      type.resize(w); frame.resize(w); object.resize(w); info.resize(w); meas.resize(w * 10);
      make_batch(g);
    }
    ctx_.check(eslam_optimize_objects_host(ctx_.get(), &g, iters, ell.data(), &st), "optimize");
    for (size_t k = 0; k < inst.size(); k++) for (int i = 0; i < 10; i++) ellipsoids[inst[k]]->vec[i] = ell[k * 10 + i];
    return st;
  }
#ifdef ESLAM_HAVE_EIGEN_OPENCV
  void SetGroundPlane(Eigen::Vector4d &normal) { SetGroundPlane(Plane4{normal[0], normal[1], normal[2], normal[3]}); }   // Optimizer.h:23
  // void GlobalObjectGraphOptimization(std::vector<Frame *> &pFrames, Map *pMap, int rows, int cols, Matrix3d &mCalib,
  //                                    std::map<int, Observations> &objectObservations, bool save_graph, bool withAssociation, bool check_visibility)
  // (include/core/Optimizer.h:19-21) on the REFERENCE's own Frame / Map / Observation / g2o::ellipsoid types (a template, so this header
  // does not need them): pMap->GetAllEllipsoidsMap() -> std::map<int, g2o::ellipsoid *>, updated in place through fromVector();
  // frames give cam_pose_Tcw.toVector(), frame_seq_id and mpLocalObjects (toVector(), prob, miInstanceID).
  template <class FrameT, class MapT, class ObsMapT>
  void GlobalObjectGraphOptimization(std::vector<FrameT *> &pFrames, MapT *pMap, int /*rows*/, int /*cols*/, Eigen::Matrix3d &mCalib, ObsMapT &objectObservations,
                                     bool /*save_graph*/ = false, bool /*withAssociation*/ = false, bool /*check_visibility*/ = false) {
    auto emap = pMap->GetAllEllipsoidsMap();
    std::map<int, ellipsoid_pod> pods; std::map<int, ellipsoid_pod *> podp;
    for (auto &kv : emap) { ellipsoid_pod e; const auto v = kv.second->toVector(); for (int i = 0; i < 10; i++) e.vec[i] = v[i]; e.prob = kv.second->prob; e.miInstanceID = kv.first; pods[kv.first] = e; }
    for (auto &kv : pods) podp[kv.first] = &kv.second;
    std::vector<Frame> frames; frames.reserve(pFrames.size());
    std::vector<std::vector<ellipsoid_pod>> locals(pFrames.size());
    for (size_t f = 0; f < pFrames.size(); f++) {
      const auto tw = pFrames[f]->cam_pose_Twc.toVector();
      frames.emplace_back(pFrames[f]->timestamp, &tw[0], std::vector<std::array<double, 8>>(), nullptr, 0, 0);
      frames.back().frame_seq_id = pFrames[f]->frame_seq_id;
      for (auto *o : pFrames[f]->mpLocalObjects) {
        ellipsoid_pod e; e.miInstanceID = -1;
        if (o) { const auto v = o->toVector(); for (int i = 0; i < 10; i++) e.vec[i] = v[i]; e.prob = o->prob; e.miInstanceID = o->miInstanceID; e.miLabel = o->miLabel; }
        locals[f].push_back(e);
      }
    }
    std::vector<Frame *> fp;
    for (size_t f = 0; f < frames.size(); f++) { for (auto &e : locals[f]) frames[f].mpLocalObjects.push_back(e.miInstanceID >= 0 ? &e : nullptr); fp.push_back(&frames[f]); }
    std::vector<Observation> obs_store; std::map<int, Observations> obs;
    size_t total = 0; for (auto &kv : objectObservations) total += kv.second.size();
    obs_store.reserve(total);
    for (auto &kv : objectObservations) for (auto *o : kv.second) {
      Frame *pf = nullptr;
      for (size_t f = 0; f < pFrames.size(); f++) if (pFrames[f] == o->pFrame) pf = &frames[f];
      obs_store.push_back(Observation{o->label, {o->bbox[0], o->bbox[1], o->bbox[2], o->bbox[3]}, o->rate, pf});
      obs[kv.first].push_back(&obs_store.back());
    }
    double K[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) K[3 * i + j] = mCalib(i, j);
    GlobalObjectGraphOptimization(fp, podp, K, obs, 10);
    for (auto &kv : emap) { Eigen::Matrix<double, 10, 1> v; for (int i = 0; i < 10; i++) v[i] = pods[kv.first].vec[i]; kv.second->fromVector(v); }
  }
#endif
 private:
  Context &ctx_; bool mbGroundPlaneSet = false; Plane4 mGroundPlaneNormal{};
};

// DataAssociationSolver (include/core/DataAssociation.h, src/core/DataAssociation.cpp:16-135).  The reference resolves one frame against the
// Map it holds; here the map snapshot is passed as (instance id -> ellipsoid) and a whole list of frames can be resolved in ONE device
// launch (Solve(frames, map)): every frame against the same snapshot, new instances numbered in (frame, row) order.  A match returns
// the COLUMN index of the instance-ordered map like the reference (DataAssociation.cpp:113-114).
class DataAssociationSolver {
 public:
  explicit DataAssociationSolver(Context &c) : ctx_(c) {}
  std::vector<int> Solve(Frame *pFrame, const std::map<int, ellipsoid_pod> &map) { std::vector<Frame *> one{pFrame}; return Solve(one, map)[0]; }
  std::vector<std::vector<int>> Solve(const std::vector<Frame *> &frames, const std::map<int, ellipsoid_pod> &map) {
    std::vector<int> start{0}, valid; std::vector<double> twc, poses, centres;
    for (Frame *f : frames) {
      for (int k = 0; k < 7; k++) twc.push_back(f->cam_pose_Twc[k]);
      for (size_t i = 0; i < f->mmObservations.size(); i++) {
        const ellipsoid_pod *o = i < f->mpLocalObjects.size() ? f->mpLocalObjects[i] : nullptr;
        valid.push_back(o ? 1 : 0);
        for (int k = 0; k < 7; k++) poses.push_back(o ? o->vec[k] : 0.0);
      }
      start.push_back((int)valid.size());
    }
    for (auto &kv : map) for (int k = 0; k < 3; k++) centres.push_back(kv.second.vec[k]);   // std::map: ascending instance id
    std::vector<int> assoc(valid.size() ? valid.size() : 1); int n_new = 0;
    int rc = eslam_associate_frames_host(ctx_.get(), (int)frames.size(), start.data(), twc.data(), (int)valid.size(), poses.data(), valid.data(), (int)map.size(),
                                         centres.data(), miInstanceNum, assoc.data(), &n_new);
    ctx_.check(rc, "eslam_associate_frames_host");
    miInstanceNum += n_new;
    std::vector<std::vector<int>> out;
    for (size_t f = 0; f < frames.size(); f++) out.emplace_back(assoc.begin() + start[f], assoc.begin() + start[f + 1]);
    return out;
  }
  int CreateInstance() { return miInstanceNum++; }
 private:
  Context &ctx_; int miInstanceNum = 0;
};

#ifdef ESLAM_HAVE_EIGEN_OPENCV
// ellipsoid_pod -> the reference's g2o::ellipsoid (or any type with fromVector(Vector10d), prob, miLabel, miInstanceID)
template <class EllipsoidT> EllipsoidT to_reference(const ellipsoid_pod &p) {
  EllipsoidT e; Eigen::Matrix<double, 10, 1> v; for (int i = 0; i < 10; i++) v[i] = p.vec[i];
  e.fromVector(v); e.prob = p.prob; e.miLabel = p.miLabel; e.miInstanceID = p.miInstanceID;
  return e;
}
#endif

}  // namespace EllipsoidSLAM

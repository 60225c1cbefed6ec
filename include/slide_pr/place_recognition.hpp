// place_recognition.hpp -- header-only C++ adapter with the reference's own signatures.
//
// Drop-in for `class PlaceRecognition` -- the SlideMatch half and the SlideGraph entry
// (findInterLoopClosureWithClipper, semantic_clipper::run_semantic_clipper) --
// (backend/sloam/include/core/place_recognition.h:31-237) on top of the C-ABI in slide_pr.h.
// It is templated on the vector / matrix types so that it works with the reference's
//   std::vector<Eigen::Vector7d>, Eigen::Matrix3d, Eigen::Matrix4d, std::vector<Eigen::Vector4d>
// (Eigen is not a dependency of this header: any type with operator[] / operator()(row, col)
// and contiguous storage of 7 doubles per landmark works, e.g. std::array<double, 7> and the
// small Mat<R, C> below).  In sloam, `#include <slide_pr/place_recognition.hpp>` and
//   using PlaceRecognitionGpu = slide_pr::PlaceRecognition;
// replaces the member `PlaceRecognition inter_loopCloser_` (sloamNode.h:175); ParamInit's rosparam
// reads (place_recognition.cpp:24-75) fill `slide_pr_params` instead of the private members.
//
// Semantics kept from the reference: outputs are caller-owned references that are overwritten;
// failure is `return false` (findInterLoopClosure / findTransformation) or an early return with
// the outputs untouched (MatchMaps' sanity check, PR.cpp:169-175); no exceptions for "not found".
// Errors of the GPU path (no device, non-finite input) are reported through last_error() and
// make the bool functions return false, which the caller treats as "no closure this period"
// (sloamNode.cpp:647-649).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../slide_pr.h"

namespace slide_pr {

// minimal column-major fixed matrix (Eigen's default storage order) for builds without Eigen
template <int R, int C>
struct Mat {
  double v[R * C];
  double &operator()(int r, int c) { return v[c * R + r]; }
  double operator()(int r, int c) const { return v[c * R + r]; }
  static Mat Identity() {
    Mat m{};
    for (int i = 0; i < R * C; i++) m.v[i] = 0.0;
    for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0;
    return m;
  }
};
using Vector7d = std::array<double, 7>;
using Vector4d = std::array<double, 4>;
using Vector3d = std::array<double, 3>;

class PlaceRecognition {
 public:
  // public members of the reference class (place_recognition.h:34-43)
  bool visualize_matching_results = false;  // RViz markers are the caller's business; ignored here
  double min_loop_closure_overlap_percentage_ = 0.1;
  bool use_lsq = true;
  bool inter_loop_closure = true;

  explicit PlaceRecognition(const slide_pr_params &params) : p_(params) {
    use_lsq = params.use_lsq != 0;
    inter_loop_closure = params.inter_loop_closure != 0;
    const int rc = slide_pr_create(&p_, &h_);
    if (rc != SLIDE_PR_OK) throw std::runtime_error(std::string("slide_pr_create: ") + slide_pr_last_error(nullptr));
  }
  PlaceRecognition() : PlaceRecognition(defaults()) {}
  ~PlaceRecognition() { slide_pr_destroy(h_); }
  PlaceRecognition(const PlaceRecognition &) = delete;
  PlaceRecognition &operator=(const PlaceRecognition &) = delete;

  static slide_pr_params defaults() {
    slide_pr_params p;
    slide_pr_default_params(&p);
    return p;
  }
  const char *last_error() const { return slide_pr_last_error(h_); }
  const slide_pr_tf_result &last_result() const { return last_; }

  // members that findTransformation sets before MatchMaps (PR.cpp:786-787, 808-809)
  double match_x_half_range_ = 0.0, match_y_half_range_ = 0.0;

  // void MatchMaps(reference_objects, query_objects, R_t_out, best_num_inliers_out,
  //                map_objects_matched_out, detection_objects_matched_out)      PR.h:70-74
  template <class Vec7List, class Mat3, class Vec4List>
  void MatchMaps(const Vec7List &reference_objects, const Vec7List &query_objects, Mat3 &R_t_out,
                 int &best_num_inliers_out, Vec4List &map_objects_matched_out,
                 Vec4List &detection_objects_matched_out) {
    sync_members();
    const int32_t n_ref = (int32_t)reference_objects.size(), n_qry = (int32_t)query_objects.size();
    std::vector<int32_t> ri(n_qry > 0 ? n_qry : 1), qi(n_qry > 0 ? n_qry : 1);
    slide_pr_match_result res;
    const int rc = slide_pr_match_maps(h_, rows(reference_objects), n_ref, rows(query_objects), n_qry,
                                       match_x_half_range_, match_y_half_range_, ri.data(), qi.data(), &res);
    if (rc != SLIDE_PR_OK || res.status == SLIDE_PR_SANITY_RETURN) return;  // outputs untouched (PR.cpp:169-175)
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) R_t_out(r, c) = res.R_t[r * 3 + c];
    best_num_inliers_out = res.best_num_inliers;
    using V4 = typename Vec4List::value_type;
    map_objects_matched_out.clear();
    detection_objects_matched_out.clear();
    for (int k = 0; k < res.n_matched; k++) {  // PR.cpp:344-350: [label, x, y, z] of both rows
      V4 m, d;
      for (int c = 0; c < 4; c++) { m[c] = reference_objects[ri[k]][c]; d[c] = query_objects[qi[k]][c]; }
      map_objects_matched_out.push_back(m);
      detection_objects_matched_out.push_back(d);
    }
  }

  // bool findTransformation(reference_objects, query_objects, xyzYaw, transform_out)   PR.h:150-153
  template <class Vec7List, class Mat4>
  bool findTransformation(const Vec7List &reference_objects, const Vec7List &query_objects,
                          std::vector<double> &xyzYaw, Mat4 &transform_out) {
    sync_members();
    const int rc = slide_pr_find_transformation(h_, rows(reference_objects), (int32_t)reference_objects.size(),
                                                rows(query_objects), (int32_t)query_objects.size(), nullptr, nullptr, &last_);
    match_x_half_range_ = last_.half_x;
    match_y_half_range_ = last_.half_y;
    if (rc != SLIDE_PR_OK) return false;
    for (int i = 0; i < 4; i++) xyzYaw.push_back(last_.xyz_yaw[i]);  // getxyzYawfromTF pushes (PR.cpp:707-710)
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) transform_out(r, c) = last_.transform[r * 4 + c];
    return true;
  }

  // bool findInterLoopClosure(reference_objects, query_objects, tfFromQueryToCandidate)   PR.h:103-106
  template <class Vec7List, class Mat4>
  bool findInterLoopClosure(const Vec7List &reference_objects, const Vec7List &query_objects,
                            Mat4 &tfFromQueryToRef) {
    sync_members();
    double tf[16];
    const int rc = slide_pr_find_inter_loop_closure(h_, rows(reference_objects), (int32_t)reference_objects.size(),
                                                    rows(query_objects), (int32_t)query_objects.size(), tf, &last_);
    match_x_half_range_ = last_.half_x;  // the members findTransformation leaves behind (PR.cpp:786-787)
    match_y_half_range_ = last_.half_y;
    if (rc != SLIDE_PR_OK) return false;
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) tfFromQueryToRef(r, c) = tf[r * 4 + c];
    return true;
  }

  // rosparams sloam/place_recognition_slidegraph/* (PR.cpp:64-75); public so that ParamInit can fill them
  slide_pr_slidegraph_params slidegraph = slidegraph_defaults();
  static slide_pr_slidegraph_params slidegraph_defaults() {
    slide_pr_slidegraph_params p;
    slide_pr_slidegraph_default_params(&p);
    return p;
  }
  const slide_pr_sc_info &last_slidegraph() const { return last_sc_; }

  // bool findInterLoopClosureWithClipper(reference_objects, query_objects, tfFromQueryToRef)   PR.h:109-112
  // (SlideGraph: Delaunay triangles -> descriptor matching -> CLIPPER -> 2-D Kabsch, PR.cpp:541-630)
  template <class Vec7List, class Mat4>
  bool findInterLoopClosureWithClipper(const Vec7List &reference_objects, const Vec7List &query_objects,
                                       Mat4 &tfFromQueryToRef) {
    double tf[16];
    const int rc = slide_pr_find_inter_loop_closure_with_clipper(h_, rows(reference_objects), (int32_t)reference_objects.size(),
                                                                 rows(query_objects), (int32_t)query_objects.size(), &slidegraph,
                                                                 tf, &last_sc_);
    if (rc != SLIDE_PR_OK) return false;  // the reference inverts whatever the caller passed in (identity, sloamNode.cpp:620); left as is
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) tfFromQueryToRef(r, c) = tf[r * 4 + c];
    return true;
  }

  // bool findIntraLoopClosure(measurements, submap, query_pose, candidate_pose, tf)   PR.h:88-91
  // Pose4: the SE3's 4x4 matrix (query_pose.matrix()).
  template <class Vec7List, class Pose4, class Mat4>
  bool findIntraLoopClosure(const Vec7List &measurements, const Vec7List &submap, const Pose4 &query_pose,
                            const Pose4 &candidate_pose, Mat4 &tfFromQuery2Candidate) {
    sync_members();
    double qp[16], cp[16], tf[16];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) { qp[r * 4 + c] = query_pose(r, c); cp[r * 4 + c] = candidate_pose(r, c); }
    const int rc = slide_pr_find_intra_loop_closure(h_, rows(measurements), (int32_t)measurements.size(), rows(submap),
                                                    (int32_t)submap.size(), qp, cp, tf, &last_);
    match_x_half_range_ = last_.half_x;  // PR.cpp:808-809
    match_y_half_range_ = last_.half_y;
    if (rc != SLIDE_PR_OK) return false;
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) tfFromQuery2Candidate(r, c) = tf[r * 4 + c];
    return true;
  }

  // void solveLSQ(map_objects_matched, detection_objects_matched, xyzyaw_out, transform_out)   PR.h:127-130
  template <class Vec3List, class Mat4>
  void solveLSQ(const Vec3List &map_objects_matched_out, const Vec3List &detection_objects_matched_out,
                std::vector<double> &xyzyaw_out, Mat4 &transform_out) {
    const int k = (int)map_objects_matched_out.size();
    std::vector<double> tgt(3 * (size_t)(k > 0 ? k : 1)), src(3 * (size_t)(k > 0 ? k : 1));
    for (int i = 0; i < k; i++)
      for (int c = 0; c < 3; c++) { tgt[3 * i + c] = map_objects_matched_out[i][c]; src[3 * i + c] = detection_objects_matched_out[i][c]; }
    double xyz_yaw[4], tf[16];
    slide_pr_solve_lsq(tgt.data(), src.data(), k, xyz_yaw, tf);
    for (int i = 0; i < 4; i++) xyzyaw_out.push_back(xyz_yaw[i]);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) transform_out(r, c) = tf[r * 4 + c];
  }

  // void getxyzYawfromTF(tf, xyzYaw)   PR.h:138
  template <class Mat4>
  void getxyzYawfromTF(const Mat4 &tf, std::vector<double> &xyzYaw) {
    double t[16], out[4];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) t[r * 4 + c] = tf(r, c);
    slide_pr_get_xyz_yaw_from_tf(t, out);
    for (int i = 0; i < 4; i++) xyzYaw.push_back(out[i]);
  }

 private:
  template <class Vec7List>
  static const double *rows(const Vec7List &v) {
    static_assert(sizeof(typename Vec7List::value_type) == 7 * sizeof(double),
                  "landmark records must be 7 contiguous doubles [label,x,y,z,d1,d2,d3]");
    return v.empty() ? nullptr : reinterpret_cast<const double *>(v.data());
  }
  void sync_members() {  // the public members may be toggled between calls (place_recognition_test.cpp:202-207)
    if ((p_.use_lsq != 0) != use_lsq || (p_.inter_loop_closure != 0) != inter_loop_closure) {
      p_.use_lsq = use_lsq ? 1 : 0;
      p_.inter_loop_closure = inter_loop_closure ? 1 : 0;
      slide_pr_set_params(h_, &p_);
    }
  }
  slide_pr_params p_;
  slide_pr_handle *h_ = nullptr;
  slide_pr_tf_result last_{};
  slide_pr_sc_info last_sc_{};
  friend struct semantic_clipper_access;
};

// semantic_clipper::run_semantic_clipper(reference_map, query_map, tfFromQuery2Ref, sigma, epsilon,
// min_num_pairs, matching_threshold)   clipper_semantic_object/include/semantic_clipper.h:38
// Maps are std::vector<std::vector<double>> rows [label, x, y, z, d1, d2, d3] (semantic_clipper.cpp:144).
struct semantic_clipper_access {
  static slide_pr_handle *handle(PlaceRecognition &pr) { return pr.h_; }
};
namespace semantic_clipper {
template <class Mat4>
inline bool run_semantic_clipper(PlaceRecognition &gpu, const std::vector<std::vector<double>> &reference_map,
                                 const std::vector<std::vector<double>> &query_map, Mat4 &tfFromQuery2Ref, double sigma,
                                 double epsilon, int min_num_pairs, double matching_threshold) {
  std::vector<double> ref(7 * reference_map.size()), qry(7 * query_map.size());
  for (size_t i = 0; i < reference_map.size(); i++)
    for (int c = 0; c < 7; c++) ref[7 * i + c] = c < (int)reference_map[i].size() ? reference_map[i][c] : 0.0;
  for (size_t i = 0; i < query_map.size(); i++)
    for (int c = 0; c < 7; c++) qry[7 * i + c] = c < (int)query_map[i].size() ? query_map[i][c] : 0.0;
  slide_pr_slidegraph_params sp = gpu.slidegraph;
  sp.sigma = sigma; sp.epsilon = epsilon; sp.num_inliers_threshold = min_num_pairs; sp.matching_threshold = matching_threshold;
  double tf[16];
  slide_pr_sc_info info;
  const int rc = slide_pr_run_semantic_clipper(semantic_clipper_access::handle(gpu), ref.data(), (int32_t)reference_map.size(), qry.data(),
                                               (int32_t)query_map.size(), &sp, nullptr, 0, nullptr, 0, nullptr, 0, tf, &info);
  if (rc != SLIDE_PR_OK) return false;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) tfFromQuery2Ref(r, c) = tf[r * 4 + c];
  return true;
}
}  // namespace semantic_clipper

}  // namespace slide_pr

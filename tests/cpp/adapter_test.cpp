// adapter_test.cpp -- exercises include/slide_pr/place_recognition.hpp with the reference's call
// pattern (place_recognition_test.cpp:60-150): load two maps from text, findInterLoopClosure,
// print the result as one JSON line.  Built and run by tests/test_cpp_adapter.py.
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <slide_pr/place_recognition.hpp>

static std::vector<slide_pr::Vector7d> load(const char *path) {
  std::vector<slide_pr::Vector7d> out;
  std::ifstream f(path);
  std::string line;
  while (std::getline(f, line)) {
    std::istringstream iss(line);
    slide_pr::Vector7d o{};
    for (int i = 0; i < 7; i++) iss >> o[i];
    out.push_back(o);
  }
  return out;
}

int main(int argc, char **argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: adapter_test ref.txt qry.txt\n"); return 2; }
  slide_pr_params p = slide_pr::PlaceRecognition::defaults();
  // params/sloam-forest-parking-lot.yaml with ignore_dimension (fixture dims are zero)
  p.match_xy_step_size = 0.5;
  p.match_yaw_angle_step_size = slide_pr_deg2rad(5.0);
  p.match_threshold = 0.5;
  p.ignore_dimension = 1;
  p.min_num_inliers = 8;
  try {
    slide_pr::PlaceRecognition place_recoger(p);
    const auto reference_objects = load(argv[1]), query_objects = load(argv[2]);
    slide_pr::Mat<4, 4> tfFromQueryToRef = slide_pr::Mat<4, 4>::Identity();
    const bool closure_found = place_recoger.findInterLoopClosure(reference_objects, query_objects, tfFromQueryToRef);
    std::vector<double> xyzYawOut;
    place_recoger.getxyzYawfromTF(tfFromQueryToRef, xyzYawOut);
    // MatchMaps directly, with the half ranges findTransformation just computed
    slide_pr::Mat<3, 3> R_t = slide_pr::Mat<3, 3>::Identity();
    int best = 0;
    std::vector<slide_pr::Vector4d> map_matched, det_matched;
    const slide_pr_tf_result &last = place_recoger.last_result();
    std::vector<slide_pr::Vector7d> sref = reference_objects, sqry = query_objects;
    for (auto &o : sref) { o[1] -= last.centroid_ref[0]; o[2] -= last.centroid_ref[1]; }
    for (auto &o : sqry) { o[1] -= last.centroid_qry[0]; o[2] -= last.centroid_qry[1]; }
    place_recoger.MatchMaps(sref, sqry, R_t, best, map_matched, det_matched);
    // SlideGraph entry points (PR.h:109-112, semantic_clipper.h:38) with the reference's call pattern
    place_recoger.slidegraph.matching_threshold = 0.1;
    place_recoger.slidegraph.min_num_map_objects_to_start = 20;
    slide_pr::Mat<4, 4> tfClipper = slide_pr::Mat<4, 4>::Identity();
    const bool sg_found = place_recoger.findInterLoopClosureWithClipper(reference_objects, query_objects, tfClipper);
    const slide_pr_sc_info sg = place_recoger.last_slidegraph();
    std::vector<std::vector<double>> ref_rows, qry_rows;
    for (const auto &o : reference_objects) ref_rows.emplace_back(o.begin(), o.end());
    for (const auto &o : query_objects) qry_rows.emplace_back(o.begin(), o.end());
    slide_pr::Mat<4, 4> tfSc = slide_pr::Mat<4, 4>::Identity();
    const bool sc_found = slide_pr::semantic_clipper::run_semantic_clipper(place_recoger, ref_rows, qry_rows, tfSc, 0.1, 0.1, 10, 0.1);
    std::printf("{\"sg_found\": %s, \"sc_found\": %s, \"sg_triangles\": [%d, %d], \"sg_matches\": %lld, \"sg_inliers\": %d, "
                "\"sg_yaw\": %.17g, \"sc_yaw\": %.17g}\n", sg_found ? "true" : "false", sc_found ? "true" : "false",
                sg.n_triangles_model, sg.n_triangles_data, (long long)sg.n_triangle_matches, sg.n_inliers,
                std::atan2(tfClipper(1, 0), tfClipper(0, 0)), std::atan2(tfSc(1, 0), tfSc(0, 0)));
    std::printf("{\"found\": %s, \"xyz_yaw\": [%.17g, %.17g, %.17g, %.17g], \"best\": %d, \"n_matched\": %zu, "
                "\"R_t\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"hyp\": %lld}\n",
                closure_found ? "true" : "false", xyzYawOut[0], xyzYawOut[1], xyzYawOut[2], xyzYawOut[3], best,
                map_matched.size(), R_t(0, 0), R_t(0, 1), R_t(0, 2), R_t(1, 0), R_t(1, 1), R_t(1, 2),
                (long long)last.match.best_hyp_index);
  } catch (const std::exception &e) {
    std::printf("{\"error\": \"%s\"}\n", e.what());
    return 3;
  }
  return 0;
}

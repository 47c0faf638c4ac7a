// TEST INFRASTRUCTURE ONLY -- the back-end half of the drop-in on live reference objects.  This translation unit
// includes the UNMODIFIED reference headers (never copied) and the product's roborts_edu_slam_b200/csrc/backend_batcher.hpp,
// and feeds rsm_adapter::BackEndBatcher from a LIVE roborts_slam::SensorDataManager filled the way
// SlamProcessor::process fills it (slam/slam_processor.cpp: AddSensorData + AddMultiresolutionRangeData per map).
// tests/ compares what it returns with the reference's own classes run candidate by candidate (libref.so).
// Output: oracle/_ref/libbatcher.so, linked against roborts_edu_slam_b200/librsm.so.
#include <algorithm>
#include <cassert>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>

#include "scan_match/correlate_scan_matcher.h"
#include "slam/sensor_data_manager.h"
#include "backend_batcher.hpp"

using namespace roborts_slam;

namespace {
struct Harness {
  SensorDataManager sdm;
  rsm_adapter::BackEndConfig cfg;
  double sdm_res_[3] = {0.05, 0.1, 0.05};
  std::unique_ptr<rsm_adapter::BackEndBatcher> batcher;
};
}  // namespace

extern "C" {

// res = {fine, coarse, pub}; dev = {fine, coarse}; sizes = {fine, coarse}; passes = 3 x 8 doubles (size, sres, aoff, ares,
// threshold, use_point_size, use_center_penalty, type); opt = {on, iterate_max_times, cost_decrease_threshold,
// cost_min_threshold, max_update_distance, max_update_angle, optimize_failed_cost}
void* batcher_create(const double* res, const double* dev, const int* sizes, double blur_offset, float default_prob,
                     const double* passes, const double* opt) {
  auto* h = new Harness;
  rsm_adapter::BackEndConfig& c = h->cfg;
  c.fine_name = "fine_scan_match_map"; c.coarse_name = "coarse_scan_match_map"; c.pub_name = "pub_map";
  c.fine_resolution = res[0]; c.coarse_resolution = res[1];
  c.fine_deviation = dev[0]; c.coarse_deviation = dev[1]; c.gaussian_blur_offset = blur_offset; c.default_cell_prob = default_prob;
  c.fine_map_size = sizes[0]; c.coarse_map_size = sizes[1];
  for (int k = 0; k < 3; ++k) {
    const double* p = passes + 8 * k;
    rsm_pass_param& q = c.pass[k];
    q.search_space_size = p[0]; q.search_space_resolution = p[1]; q.search_angle_offset = p[2]; q.search_angle_resolution = p[3];
    q.response_threshold = p[4]; q.use_point_size = static_cast<int>(p[5]); q.use_center_penalty = p[6] != 0.0 ? 1 : 0;
    q.type = static_cast<int>(p[7]); q.reserved = 0;
  }
  c.use_optimize_scan_match = opt[0] != 0.0;
  c.optimize.iterate_max_times = static_cast<int>(opt[1]); c.optimize.cost_decrease_threshold = opt[2];
  c.optimize.cost_min_threshold = opt[3]; c.optimize.max_update_distance = opt[4]; c.optimize.max_update_angle = opt[5];
  c.optimize.reserved = 0;
  c.optimize_failed_cost = opt[6];
  h->sdm_res_[0] = res[0]; h->sdm_res_[1] = res[1]; h->sdm_res_[2] = res[2];
  try { h->batcher.reset(new rsm_adapter::BackEndBatcher(c, 0)); } catch (...) { delete h; return nullptr; }
  return h;
}
void batcher_destroy(void* hv) { delete static_cast<Harness*>(hv); }

// One accepted scan, as SlamProcessor::process stores it: the scan in metres with its sensor pose, and its copies in the
// cells of every map (RangeDataContainer::CreateFrom(scan, 1 / resolution), sensor_data_manager.h:99-115).
int batcher_add_scan(void* hv, int n, const double* xy_metres, const double* pose_world) {
  auto* h = static_cast<Harness*>(hv);
  auto rd = std::make_shared<RangeDataContainer2d>(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i) rd->AddDataPoint(Eigen::Vector2d(xy_metres[2 * i], xy_metres[2 * i + 1]));
  rd->set_sensor_origin(Eigen::Vector2d(0.0, 0.0));
  rd->set_sensor_pose(Eigen::Vector3d(pose_world[0], pose_world[1], pose_world[2]));
  h->sdm.AddSensorData(rd, OdometryData(Eigen::Vector3d(pose_world[0], pose_world[1], pose_world[2])));
  const std::string names[3] = {h->cfg.fine_name, h->cfg.coarse_name, h->cfg.pub_name};
  for (int k = 0; k < 3; ++k) {
    auto scaled = std::make_shared<RangeDataContainer2d>();
    scaled->CreateFrom(rd, 1 / h->sdm_res_[k]);
    h->sdm.AddMultiresolutionRangeData(names[k], scaled);
  }
  return h->sdm.current_data_index();
}

// what the fine-map copy of scan `id` holds (cells), so that the test feeds the reference exactly the same numbers
int batcher_get_fine_scan(void* hv, int id, double* xy_out, int cap) {
  auto* h = static_cast<Harness*>(hv);
  auto rd = h->sdm.GetMultiresolutionRangeData(h->cfg.fine_name, id);
  const int n = rd->GetSize();
  for (int i = 0; i < n && i < cap; ++i) { xy_out[2 * i] = rd->GetDataPoint(i)[0]; xy_out[2 * i + 1] = rd->GetDataPoint(i)[1]; }
  return n;
}

// SlamProcessor::UpdateRangeData (slam_processor.cpp:597-603): a corrected sensor pose
void batcher_set_pose(void* hv, int id, const double* pose_world) {
  auto* h = static_cast<Harness*>(hv);
  h->sdm.GetRangeData(id)->set_sensor_pose(Eigen::Vector3d(pose_world[0], pose_world[1], pose_world[2]));
}

// BackEndBatcher::TryCloseLoop on the live manager: returns the accepted chain index or -1
int batcher_try_close_loop(void* hv, int range_id, int n_chains, const int* chain_off, const int* chain_ids, const double* scan_pose,
                           const double* centre, const double* thresholds, double* best_pose, double* cov, double* stage1,
                           double* stage2) {
  auto* h = static_cast<Harness*>(hv);
  try {
    h->batcher->SyncScans(h->sdm);
    std::vector<std::vector<int>> chains(n_chains);
    for (int c = 0; c < n_chains; ++c) chains[c].assign(chain_ids + chain_off[c], chain_ids + chain_off[c + 1]);
    rsm_adapter::BackEndBatcher::LoopClosureThresholds th;
    th.min_response_coarse = thresholds[0]; th.max_variance_coarse = thresholds[1]; th.min_response_fine = thresholds[2];
    std::vector<double> s1, s2;
    const int hit = h->batcher->TryCloseLoop(range_id, chains, scan_pose, centre, th, best_pose, cov, &s1, &s2);
    for (int c = 0; c < n_chains; ++c) { stage1[c] = s1[c]; stage2[c] = s2[c]; }
    return hit;
  } catch (const std::exception& e) {
    LOG(WARNING) << "batcher_try_close_loop: " << e.what();
    return -2;
  }
}

}  // extern "C"

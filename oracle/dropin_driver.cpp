// TEST INFRASTRUCTURE ONLY -- proves the drop-in boundary.  This translation unit includes the
// UNMODIFIED reference headers (from /root/reference/src, never copied) and the product's C++
// adapter (roborts_edu_slam_b200/csrc/scan_matcher_adapter.hpp), and runs the adapter class on a
// LIVE reference ScanMatchMap / RangeDataContainer2d / CorrelationScanMatchParam -- the exact
// objects scan_matchers.h passes to BasedCorrelationScanMatch::ScanMatch.  tests/ compares the
// result with what the reference class returns for the same objects (libref.so).
// Output: oracle/_ref/libdropin.so, linked against roborts_edu_slam_b200/librsm.so.
#include <algorithm>
#include <cassert>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>

#include "scan_match/correlate_scan_matcher.h"
#include "scan_match/optimize_scan_matcher.h"
#include "ref_types.h"
#include "scan_matcher_adapter.hpp"

using namespace roborts_slam;

namespace {
std::shared_ptr<RangeDataContainer2d> MakeScan(int n, const double* xy, const double* pose_world = nullptr) {
  auto rd = std::make_shared<RangeDataContainer2d>(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i) rd->AddDataPoint(Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]));
  rd->set_sensor_origin(Eigen::Vector2d(0.0, 0.0));
  if (pose_world) rd->set_sensor_pose(Eigen::Vector3d(pose_world[0], pose_world[1], pose_world[2]));
  else rd->set_sensor_pose(Eigen::Vector3d(0.0, 0.0, 0.0));
  return rd;
}
std::shared_ptr<CorrelationScanMatchParam> MakeParam(const double* p) {
  auto q = std::make_shared<CorrelationScanMatchParam>();
  q->set_search_space_size(p[0]);
  q->set_search_space_resolution(p[1]);
  q->set_search_angle_offset(p[2]);
  q->set_search_angle_resolution(p[3]);
  q->set_response_threshold(p[4]);
  q->set_use_point_size(static_cast<int>(p[5]));
  q->set_use_center_penalty(p[6] != 0.0);
  q->set_correlation_scan_match_type(static_cast<CorrelationScanMatchType>(static_cast<int>(p[7])));
  q->set_max_depth(0);
  return q;
}
}  // namespace

extern "C" {

void* dropin_create(int device) {
  try { return new rsm_adapter::BasedCorrelationScanMatch(device); } catch (...) { return nullptr; }
}
void dropin_destroy(void* h) { delete static_cast<rsm_adapter::BasedCorrelationScanMatch*>(h); }

// the 3-pass chain of ScanMatchers::ScanMatch (scan_matchers.h:224-263, optimiser off) with the
// adapter in place of correlate_scan_matcher_; n_pass = 1 runs a single pass
double dropin_match_chain(void* h, void* ref_map, int n, const double* xy, const double* params, int n_pass,
                          double* pose_world, double* cov, double* resp_out, int* exact_used) {
  auto* m = static_cast<rsm_adapter::BasedCorrelationScanMatch*>(h);
  auto* rm = static_cast<RefMap*>(ref_map);
  auto rd = MakeScan(n, xy);
  Eigen::Vector3d pose(pose_world[0], pose_world[1], pose_world[2]);
  Eigen::Matrix3d c;
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c(r, k) = cov[3 * r + k];
  double score = 0.0;
  int used = 0;
  for (int i = 0; i < n_pass; ++i) {
    const double r = m->ScanMatch(rm->map, rd, MakeParam(params + 8 * i), pose, c);
    if (resp_out) resp_out[i] = r;
    score += r;
    used += m->last_detail().exact_sort_used;
  }
  score /= n_pass;
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) cov[3 * r + k] = c(r, k);
  pose_world[0] = pose[0]; pose_world[1] = pose[1]; pose_world[2] = pose[2];
  if (exact_used) *exact_used = used;
  return score;
}

// the adapter's Gauss-Newton class on a live reference map: op = {iterate_max_times, cost_decrease_threshold,
// cost_min_threshold, max_update_distance, max_update_angle}; returns the cost, pose_world in/out
// rsm_adapter::UpdateMapByRange in place of map->UpdateMapByRange (slam_processor.cpp:556,558): the reference's own
// update on the live host map + the same stamp on the device mirror.  1 = stamped, 0 = the map was extended instead.
int dropin_update_map(void* ref_map, int n, const double* xy, const double* pose_world, int use_blur, double deviation,
                      double occu_offset) {
  auto* rm = static_cast<RefMap*>(ref_map);
  return rsm_adapter::UpdateMapByRange(rm->map, MakeScan(n, xy, pose_world), use_blur != 0, deviation, occu_offset) ? 1 : 0;
}

// 1 = the device mirror of the map holds exactly the host map's cells, 0 = it differs, -1 = no mirror
int dropin_mirror_equals_host(void* ref_map) {
  auto* rm = static_cast<RefMap*>(ref_map);
  const std::vector<float> dev = rsm_adapter::DeviceMaps::Get(0).DownloadMirror(rm->map);
  const int n = rm->map->GetSizeX() * rm->map->GetSizeY();
  if (dev.empty() || static_cast<int>(dev.size()) != n) return -1;
  for (int i = 0; i < n; ++i) if (dev[i] != rm->map->GetCellValue(i)) return 0;
  return 1;
}

void dropin_sync_counts(long* full_uploads, long* incremental_updates) {
  *full_uploads = rsm_adapter::DeviceMaps::Get(0).full_uploads();
  *incremental_updates = rsm_adapter::DeviceMaps::Get(0).incremental_updates();
}

void* dropin_opt_create(int device) {
  try { return new rsm_adapter::BasedOptimizeScanMatch(device); } catch (...) { return nullptr; }
}
void dropin_opt_destroy(void* h) { delete static_cast<rsm_adapter::BasedOptimizeScanMatch*>(h); }
double dropin_optimize(void* h, void* ref_map, int n, const double* xy, const double* op, double* pose_world) {
  auto* m = static_cast<rsm_adapter::BasedOptimizeScanMatch*>(h);
  auto* rm = static_cast<RefMap*>(ref_map);
  auto rd = MakeScan(n, xy);
  auto q = std::make_shared<OptimizeScanMatchParam>();
  q->set_iterate_max_times(static_cast<int>(op[0]));
  q->set_cost_decrease_threshold(op[1]);
  q->set_cost_min_threshold(op[2]);
  q->set_max_update_distance(op[3]);
  q->set_max_update_angle_(op[4]);
  Eigen::Vector3d pose(pose_world[0], pose_world[1], pose_world[2]);
  const double cost = m->ScanMatch(rm->map, rd, q, pose);
  pose_world[0] = pose[0]; pose_world[1] = pose[1]; pose_world[2] = pose[2];
  return cost;
}

}  // extern "C"

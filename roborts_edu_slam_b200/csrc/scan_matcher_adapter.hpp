// scan_matcher_adapter.hpp -- drop-in replacement for the reference's correlative matcher class.
//
// A maintainer of RoboRTS-Edu-SLAM includes this header AFTER scan_match/correlate_scan_matcher.h
// (it uses the reference's own ScanMatchMap, RangeDataContainer2d, CorrelationScanMatchParam and
// Eigen types) and swaps the member type in scan_match/scan_matchers.h:400
//
//     std::unique_ptr<BasedCorrelationScanMatch> correlate_scan_matcher_;
// ->  std::unique_ptr<rsm_adapter::BasedCorrelationScanMatch> correlate_scan_matcher_;
//
// Nothing else changes: ScanMatchers::ScanMatch (scan_matchers.h:238-259), SlamProcessor::process
// (slam_processor.cpp:143) and ScanMatchInterface (:301) compile and behave as before.  The class
// keeps the signature, in/out semantics and error behaviour of
// BasedCorrelationScanMatch::ScanMatch (correlate_scan_matcher.h:784-875):
//   * !map->IsMapInit() or an empty scan -> returns 0.0, pose and covariance untouched (:792-795);
//   * the covariance is rewritten according to the pass type (:835-858);
//   * the pose is adopted only when response > response_threshold (:866-869);
//   * world <-> map conversions use the reference map's own GetMapCoordsPose / GetWorldCoordsPose,
//     so they are whatever Eigen computes in the host build.
// The lookup grid is mirrored on the device (DeviceMaps below, shared by both adapter classes): uploaded
// whole when the live map changed behind the mirror's back, stamped incrementally when the map update goes
// through rsm_adapter::UpdateMapByRange (two more edited lines, slam/slam_processor.cpp:556,558).
// INTEGRATION.md also shows how to keep the maps on the device only and how to batch loop-closure candidates.
//
// Link with -lrsm.  Not thread-safe per instance, like the reference (scan_match_mutex_).
#ifndef RSM_SCAN_MATCHER_ADAPTER_HPP_
#define RSM_SCAN_MATCHER_ADAPTER_HPP_

#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "rsm.h"

namespace rsm_adapter {

// One context per device and one device mirror per live reference map, shared by the correlative and the Gauss-Newton
// adapter (and by every instance of them): a map is uploaded once, not once per adapter class.  The mirror follows the
// host map in two ways:
//   * Sync()             whole-plane upload when the host map changed behind the mirror's back (another map_update_index()
//                        than the mirror saw, another size or scale) -- always correct, 33 ms for a 2400^2 fine map;
//   * UpdateMapByRange() the drop-in for the two calls at slam/slam_processor.cpp:556,558: runs the reference's own
//                        map->UpdateMapByRange(range_data, use_blur) and stamps the same scan on the mirror
//                        (rsm_grid_update_by_range_map, 0.1 ms), so the next match finds the mirror current.  The map's
//                        deviation and gaussian_blur_offset are private there, hence the two extra arguments (the call
//                        site has them in param_).  An update that resized the host map leaves the mirror stale and the
//                        next Sync() uploads the new plane.
// All entry points take the registry's mutex: the reference calls the matcher from two threads (scan_match_mutex_) and
// the map update from a third lock (map_mutex_), while a context is not re-entrant.
class DeviceMaps {
 public:
  static DeviceMaps& Get(int device = 0) {
    // (never destroyed: at process exit the CUDA runtime may already be gone)
    static std::mutex table_mu;
    static std::map<int, DeviceMaps*>* table = new std::map<int, DeviceMaps*>();
    std::lock_guard<std::mutex> lk(table_mu);
    DeviceMaps*& slot = (*table)[device];
    if (!slot) slot = new DeviceMaps(device);
    return *slot;
  }
  ~DeviceMaps() {
    for (auto& kv : mirrors_) if (kv.second.grid) rsm_grid_destroy(ctx_, kv.second.grid);
    if (ctx_) rsm_destroy(ctx_);
  }
  std::recursive_mutex& mutex() { return mu_; }
  rsm_ctx* context() const { return ctx_; }
  long full_uploads() const { return full_uploads_; }
  long incremental_updates() const { return incremental_updates_; }

  // the device grid of a live map, uploaded if the mirror is stale (call with mutex() held)
  rsm_grid* Sync(const std::shared_ptr<roborts_slam::ScanMatchMap>& map_ptr) {
    roborts_slam::ScanMatchMap& map = *map_ptr;
    Mirror& M = mirror_of(map_ptr);
    const int sx = map.GetSizeX(), sy = map.GetSizeY();
    if (M.grid && sx == M.sx && sy == M.sy && map.map_update_index() == M.update && map.get_scale_factor() == M.scale) return M.grid;
    if (!M.grid || sx != M.sx || sy != M.sy || map.get_scale_factor() != M.scale) {
      if (M.grid) { rsm_grid_destroy(ctx_, M.grid); M.grid = nullptr; }
      // the adapters do world<->map with the live map's own transform, so the offset handed to the library is irrelevant
      if (rsm_grid_create_from_scale(ctx_, sx, sy, map.get_scale_factor(), 0.0, 0.0, &M.grid) != RSM_OK)
        throw std::runtime_error(std::string("rsm_grid_create failed: ") + rsm_last_error(ctx_));
    }
    cells_.resize(static_cast<size_t>(sx) * sy);
    for (int i = 0; i < sx * sy; ++i) cells_[i] = map.GetCellValue(i);   // ProbabilityCell::prob_value_
    if (rsm_grid_upload_f32(ctx_, M.grid, cells_.data()) != RSM_OK)
      throw std::runtime_error(std::string("rsm_grid_upload_f32 failed: ") + rsm_last_error(ctx_));
    M.sx = sx; M.sy = sy; M.update = map.map_update_index(); M.scale = map.get_scale_factor();
    ++full_uploads_;
    return M.grid;
  }

  // OccuGridMap::UpdateMapByRange on the host map and, when the mirror was current, the same stamp on the device
  bool UpdateMapByRange(const std::shared_ptr<roborts_slam::ScanMatchMap>& map_ptr,
                        const std::shared_ptr<roborts_slam::RangeDataContainer2d>& range_data, bool use_blur,
                        double deviation, double gaussian_blur_offset) {
    std::lock_guard<std::recursive_mutex> lk(mu_);
    roborts_slam::ScanMatchMap& map = *map_ptr;
    Mirror& M = mirror_of(map_ptr);
    const bool current = M.grid && map.GetSizeX() == M.sx && map.GetSizeY() == M.sy && map.map_update_index() == M.update &&
                         map.get_scale_factor() == M.scale;
    // the pose in map cells BEFORE the host update: an update that extends the map moves its offset
    const Eigen::Vector3d pose_map = map.GetMapCoordsPose(range_data->sensor_pose());     // occu_grid_map.h:278
    const bool stamped = map.UpdateMapByRange(range_data, use_blur);
    if (!current || !stamped || map.GetSizeX() != M.sx || map.GetSizeY() != M.sy) return stamped;   // next Sync() uploads
    const int n = range_data->GetSize();
    pts_.resize(2 * static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) {
      const Eigen::Vector2d& p = range_data->GetDataPoint(i);
      pts_[2 * i] = p[0];
      pts_[2 * i + 1] = p[1];
    }
    const double pm[3] = {pose_map[0], pose_map[1], pose_map[2]};
    if (rsm_grid_update_by_range_map(ctx_, M.grid, deviation, gaussian_blur_offset, use_blur ? 1 : 0, pts_.data(), n, pm) == RSM_OK) {
      M.update = map.map_update_index();
      ++incremental_updates_;
    }
    return stamped;
  }

  // test aid: the mirror's cells (empty if the map has no mirror)
  std::vector<float> DownloadMirror(const std::shared_ptr<roborts_slam::ScanMatchMap>& map_ptr) {
    std::lock_guard<std::recursive_mutex> lk(mu_);
    Mirror& M = mirror_of(map_ptr);
    std::vector<float> out;
    if (!M.grid) return out;
    out.resize(static_cast<size_t>(M.sx) * M.sy);
    if (rsm_grid_download_f32(ctx_, M.grid, out.data()) != RSM_OK) out.clear();
    return out;
  }

 private:
  struct Mirror {
    rsm_grid* grid = nullptr;
    std::weak_ptr<roborts_slam::ScanMatchMap> map;
    int sx = 0, sy = 0, update = -2;
    double scale = 0.0;
  };
  explicit DeviceMaps(int device) {
    const int rc = rsm_create(device, &ctx_);
    if (rc != RSM_OK) throw std::runtime_error("rsm_create failed (status " + std::to_string(rc) + "): no CUDA device, and there is no CPU path");
  }
  // a weak_ptr tells a new map object at a recycled address from the one that was mirrored
  Mirror& mirror_of(const std::shared_ptr<roborts_slam::ScanMatchMap>& map_ptr) {
    for (auto it = mirrors_.begin(); it != mirrors_.end();) {
      if (it->second.map.expired()) { if (it->second.grid) rsm_grid_destroy(ctx_, it->second.grid); it = mirrors_.erase(it); }
      else ++it;
    }
    Mirror& M = mirrors_[map_ptr.get()];
    if (M.map.lock() != map_ptr) { if (M.grid) rsm_grid_destroy(ctx_, M.grid); M = Mirror(); M.map = map_ptr; }
    return M;
  }
  rsm_ctx* ctx_ = nullptr;
  std::recursive_mutex mu_;
  std::map<const void*, Mirror> mirrors_;
  std::vector<float> cells_;
  std::vector<double> pts_;
  long full_uploads_ = 0, incremental_updates_ = 0;
};

// drop-in for map->UpdateMapByRange(range_data, use_blur) on a scan-match map (slam/slam_processor.cpp:556,558)
inline bool UpdateMapByRange(const std::shared_ptr<roborts_slam::ScanMatchMap>& map,
                             const std::shared_ptr<roborts_slam::RangeDataContainer2d>& range_data, bool use_blur,
                             double deviation, double gaussian_blur_offset, int device = 0) {
  return DeviceMaps::Get(device).UpdateMapByRange(map, range_data, use_blur, deviation, gaussian_blur_offset);
}

class BasedCorrelationScanMatch {
 public:
  explicit BasedCorrelationScanMatch(int device = 0) : maps_(DeviceMaps::Get(device)), ctx_(maps_.context()) {}
  BasedCorrelationScanMatch(const BasedCorrelationScanMatch&) = delete;
  BasedCorrelationScanMatch& operator=(const BasedCorrelationScanMatch&) = delete;

  double ScanMatch(std::shared_ptr<roborts_slam::ScanMatchMap> map,
                   std::shared_ptr<roborts_slam::RangeDataContainer2d> range_data,
                   std::shared_ptr<roborts_slam::CorrelationScanMatchParam> scan_match_param,
                   Eigen::Vector3d& current_pose, Eigen::Matrix3d& cov_matrix) {
    if (!map->IsMapInit() || range_data->GetSize() == 0) {
      LOG(WARNING) << "Invalid scan match input !";
      return 0.0;
    }
    std::lock_guard<std::recursive_mutex> lk(maps_.mutex());
    rsm_grid* grid_ = maps_.Sync(map);
    const int n = range_data->GetSize();
    pts_.resize(2 * static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) {
      const Eigen::Vector2d& p = range_data->GetDataPoint(i);
      pts_[2 * i] = p[0];
      pts_[2 * i + 1] = p[1];
    }
    rsm_pass_param p;
    p.search_space_size = scan_match_param->search_space_size();
    p.search_space_resolution = scan_match_param->search_space_resolution();
    p.search_angle_offset = scan_match_param->search_angle_offset();
    p.search_angle_resolution = scan_match_param->search_angle_resolution();
    p.response_threshold = scan_match_param->response_threshold();
    p.use_point_size = scan_match_param->use_point_size();
    p.use_center_penalty = scan_match_param->use_center_penalty() ? 1 : 0;
    p.type = static_cast<int32_t>(scan_match_param->correlation_scan_match_type());
    p.reserved = 0;

    const Eigen::Vector3d center = map->GetMapCoordsPose(current_pose);   // :809-810
    const double center_map[3] = {center[0], center[1], center[2]};
    double cov[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[3 * r + c] = cov_matrix(r, c);
    double response = 0.0, best[3] = {0.0, 0.0, 0.0};
    const int rc = rsm_match_map(ctx_, grid_, pts_.data(), n, &p, center_map, cov, &response, best, &last_detail_);
    if (rc != RSM_OK) {
      // the reference has no error channel here; behave like its invalid-input branch
      LOG(WARNING) << "rsm_match_map failed: " << rsm_last_error(ctx_);
      return 0.0;
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov_matrix(r, c) = cov[3 * r + c];
    if (response > scan_match_param->response_threshold()) {
      current_pose = map->GetWorldCoordsPose(Eigen::Vector3d(best[0], best[1], best[2]));   // :866-869
    }
    return response;
  }

  const rsm_pass_detail& last_detail() const { return last_detail_; }
  rsm_ctx* context() const { return ctx_; }
  DeviceMaps& maps() { return maps_; }

 private:
  DeviceMaps& maps_;
  rsm_ctx* ctx_ = nullptr;
  std::vector<double> pts_;
  rsm_pass_detail last_detail_{};
};

#ifdef ROBORTS_SLAM_SCAN_MATCH_OPTIMIZE_SCAN_MATCHER_H
// Drop-in for the reference's Gauss-Newton matcher (scan_match/optimize_scan_matcher.h:60-237), defined when that
// header was included first (scan_matchers.h includes both).  Swap scan_matchers.h:401
//     std::unique_ptr<BasedOptimizeScanMatch> optimize_scan_matcher_;
// ->  std::unique_ptr<rsm_adapter::BasedOptimizeScanMatch> optimize_scan_matcher_;
// Same signature and behaviour: invalid input or a NaN step returns kMaxCost = 1000 with best_pose untouched
// (:73-76, :103-106); otherwise best_pose is the optimised pose (angle normalised, :125-127) and the return value
// the last cost.  World <-> map uses the live map's own transform.
class BasedOptimizeScanMatch {
 public:
  explicit BasedOptimizeScanMatch(int device = 0) : maps_(DeviceMaps::Get(device)), ctx_(maps_.context()) {}
  BasedOptimizeScanMatch(const BasedOptimizeScanMatch&) = delete;
  BasedOptimizeScanMatch& operator=(const BasedOptimizeScanMatch&) = delete;

  double ScanMatch(std::shared_ptr<roborts_slam::ScanMatchMap> map,
                   std::shared_ptr<roborts_slam::RangeDataContainer2d> range_data,
                   std::shared_ptr<roborts_slam::OptimizeScanMatchParam> optimize_scan_match_param,
                   Eigen::Vector3d& best_pose) {
    const double kMaxCost = 1.0 * 1000;
    if (!map->IsMapInit() || range_data->GetSize() == 0) {
      LOG(WARNING) << "Invalid scan match input !";
      return kMaxCost;
    }
    std::lock_guard<std::recursive_mutex> lk(maps_.mutex());
    rsm_grid* grid_ = maps_.Sync(map);
    const int n = range_data->GetSize();
    pts_.resize(2 * static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) {
      const Eigen::Vector2d& p = range_data->GetDataPoint(i);
      pts_[2 * i] = p[0];
      pts_[2 * i + 1] = p[1];
    }
    rsm_optimize_param q;
    q.iterate_max_times = optimize_scan_match_param->iterate_max_times();
    q.cost_decrease_threshold = optimize_scan_match_param->cost_decrease_threshold();
    q.cost_min_threshold = optimize_scan_match_param->cost_min_threshold();
    q.max_update_distance = optimize_scan_match_param->max_update_distance();
    q.max_update_angle = optimize_scan_match_param->max_update_angle();
    q.reserved = 0;
    const Eigen::Vector3d est = map->GetMapCoordsPose(best_pose);                 // :80-81
    double pose_map[3] = {est[0], est[1], est[2]};
    double cost = kMaxCost;
    const int rc = rsm_optimize_map(ctx_, grid_, pts_.data(), n, &q, pose_map, &cost, &last_iterations_);
    if (rc != RSM_OK) {
      LOG(WARNING) << "rsm_optimize_map failed: " << rsm_last_error(ctx_);
      return kMaxCost;
    }
    if (cost == kMaxCost && pose_map[0] == est[0] && pose_map[1] == est[1] && pose_map[2] == est[2]) return kMaxCost;   // NaN step
    best_pose = map->GetWorldCoordsPose(Eigen::Vector3d(pose_map[0], pose_map[1], pose_map[2]));   // :127
    return cost;
  }

  int last_iterations() const { return last_iterations_; }

 private:
  DeviceMaps& maps_;
  rsm_ctx* ctx_ = nullptr;
  std::vector<double> pts_;
  int32_t last_iterations_ = 0;
};
#endif  // ROBORTS_SLAM_SCAN_MATCH_OPTIMIZE_SCAN_MATCHER_H

}  // namespace rsm_adapter

#endif

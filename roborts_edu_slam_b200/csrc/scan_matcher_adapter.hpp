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
// The lookup grid is handed to the device whenever the live map changed (map_update_index(),
// size or address); INTEGRATION.md shows how to avoid that copy by building the grid on the
// device (rsm_grid_rasterize) and how to batch loop-closure candidates.
//
// Link with -lrsm.  Not thread-safe per instance, like the reference (scan_match_mutex_).
#ifndef RSM_SCAN_MATCHER_ADAPTER_HPP_
#define RSM_SCAN_MATCHER_ADAPTER_HPP_

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "rsm.h"

namespace rsm_adapter {

class BasedCorrelationScanMatch {
 public:
  explicit BasedCorrelationScanMatch(int device = 0) {
    const int rc = rsm_create(device, &ctx_);
    if (rc != RSM_OK) throw std::runtime_error("rsm_create failed (status " + std::to_string(rc) + "): no CUDA device, and there is no CPU path");
  }
  ~BasedCorrelationScanMatch() {
    if (grid_) rsm_grid_destroy(ctx_, grid_);
    if (ctx_) rsm_destroy(ctx_);
  }
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
    SyncGrid(map);
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

 private:
  // Hand the live map's prob_value_ plane to the device when it changed since the last call.
  // "Changed" = another map object (a weak_ptr tells a new object at a recycled address from the
  // one that was synced), another size / scale, or another update index: every write to the
  // cells goes through UpdateMapByRange, which ends in SetUpdated() (occu_grid_map.h:325).
  void SyncGrid(const std::shared_ptr<roborts_slam::ScanMatchMap>& map_ptr) {
    roborts_slam::ScanMatchMap& map = *map_ptr;
    const int sx = map.GetSizeX(), sy = map.GetSizeY();
    const bool same = grid_ && seen_map_.lock() == map_ptr && sx == seen_sx_ && sy == seen_sy_ &&
                      map.map_update_index() == seen_update_ && map.get_scale_factor() == seen_scale_;
    if (same) return;
    if (!grid_ || sx != seen_sx_ || sy != seen_sy_ || map.get_scale_factor() != seen_scale_) {
      if (grid_) { rsm_grid_destroy(ctx_, grid_); grid_ = nullptr; }
      // the adapter does world<->map itself, so the offset handed to the library is irrelevant
      if (rsm_grid_create_from_scale(ctx_, sx, sy, map.get_scale_factor(), 0.0, 0.0, &grid_) != RSM_OK)
        throw std::runtime_error(std::string("rsm_grid_create failed: ") + rsm_last_error(ctx_));
    }
    cells_.resize(static_cast<size_t>(sx) * sy);
    for (int i = 0; i < sx * sy; ++i) cells_[i] = map.GetCellValue(i);   // ProbabilityCell::prob_value_
    if (rsm_grid_upload_f32(ctx_, grid_, cells_.data()) != RSM_OK)
      throw std::runtime_error(std::string("rsm_grid_upload_f32 failed: ") + rsm_last_error(ctx_));
    seen_map_ = map_ptr; seen_sx_ = sx; seen_sy_ = sy;
    seen_update_ = map.map_update_index(); seen_scale_ = map.get_scale_factor();
  }

  rsm_ctx* ctx_ = nullptr;
  rsm_grid* grid_ = nullptr;
  std::weak_ptr<roborts_slam::ScanMatchMap> seen_map_;
  int seen_sx_ = 0, seen_sy_ = 0, seen_update_ = -2;
  double seen_scale_ = 0.0;
  std::vector<float> cells_;
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
  explicit BasedOptimizeScanMatch(int device = 0) {
    const int rc = rsm_create(device, &ctx_);
    if (rc != RSM_OK) throw std::runtime_error("rsm_create failed (status " + std::to_string(rc) + "): no CUDA device, and there is no CPU path");
  }
  ~BasedOptimizeScanMatch() {
    if (grid_) rsm_grid_destroy(ctx_, grid_);
    if (ctx_) rsm_destroy(ctx_);
  }
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
    SyncGrid(map);
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
  void SyncGrid(const std::shared_ptr<roborts_slam::ScanMatchMap>& map_ptr) {
    roborts_slam::ScanMatchMap& map = *map_ptr;
    const int sx = map.GetSizeX(), sy = map.GetSizeY();
    const bool same = grid_ && seen_map_.lock() == map_ptr && sx == seen_sx_ && sy == seen_sy_ &&
                      map.map_update_index() == seen_update_ && map.get_scale_factor() == seen_scale_;
    if (same) return;
    if (!grid_ || sx != seen_sx_ || sy != seen_sy_ || map.get_scale_factor() != seen_scale_) {
      if (grid_) { rsm_grid_destroy(ctx_, grid_); grid_ = nullptr; }
      if (rsm_grid_create_from_scale(ctx_, sx, sy, map.get_scale_factor(), 0.0, 0.0, &grid_) != RSM_OK)
        throw std::runtime_error(std::string("rsm_grid_create failed: ") + rsm_last_error(ctx_));
    }
    cells_.resize(static_cast<size_t>(sx) * sy);
    for (int i = 0; i < sx * sy; ++i) cells_[i] = map.GetCellValue(i);
    if (rsm_grid_upload_f32(ctx_, grid_, cells_.data()) != RSM_OK)
      throw std::runtime_error(std::string("rsm_grid_upload_f32 failed: ") + rsm_last_error(ctx_));
    seen_map_ = map_ptr; seen_sx_ = sx; seen_sy_ = sy;
    seen_update_ = map.map_update_index(); seen_scale_ = map.get_scale_factor();
  }

  rsm_ctx* ctx_ = nullptr;
  rsm_grid* grid_ = nullptr;
  std::weak_ptr<roborts_slam::ScanMatchMap> seen_map_;
  int seen_sx_ = 0, seen_sy_ = 0, seen_update_ = -2;
  double seen_scale_ = 0.0;
  std::vector<float> cells_;
  std::vector<double> pts_;
  int32_t last_iterations_ = 0;
};
#endif  // ROBORTS_SLAM_SCAN_MATCH_OPTIMIZE_SCAN_MATCHER_H

}  // namespace rsm_adapter

#endif

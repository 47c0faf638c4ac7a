// backend_batcher.hpp -- the back-end (loop-closure) side of the drop-in: many ScanMatchInterface calls as ONE batched
// call, fed from the reference's own live SensorDataManager.
//
// The reference matches loop-closure candidates one at a time through the ScanMatchFunc it hands to the pose graph
// (pose_graph/range_scan_pose_graph.h:30-35 -> SlamProcessor::ScanMatchInterface, slam/slam_processor.cpp:250-326):
// per candidate it resets the two back-end maps from the chain's scans, runs ScanMatchers::ScanMatch and scales the score
// by the map check.  Candidates that are evaluated against the SAME state -- every chain FindPossibleLoopClosure yields
// for one scan until one of them is accepted (range_scan_pose_graph.cpp:299-352), every near chain of LinkNearChains
// (:120-167) -- do not depend on each other, so they can go to the device together:
//
//   BackEndBatcher::SyncScans       mirrors the manager's multi-resolution scans into device scan stores (each scan is
//                                   uploaded once, poses are refreshed every call: SlamProcessor::UpdateRangeData)
//   BackEndBatcher::ScanMatchBatch  ScanMatchInterface for a list of (scan id, chain ids, seed pose) candidates:
//                                   rsm_scan_match_interface_batch[_opt]
//   BackEndBatcher::TryCloseLoop    RangeScanPoseGraph::TryCloseLoop's two-stage test over the chains of one scan: every
//                                   chain is matched from the scan's pose (stage 1, one batch), the chains that pass the
//                                   coarse thresholds are re-matched from the refined pose (stage 2, one batch), and the
//                                   FIRST chain, in the reference's order, whose fine response passes is the loop closure.
//                                   The reference stops evaluating against the old state at exactly that point (it corrects
//                                   the poses and re-enumerates), so the outcome is the same as its sequential loop.
//
// Include after slam/sensor_data_manager.h and scan_match/correlate_scan_matcher.h.  Link with -lrsm.
#ifndef RSM_BACKEND_BATCHER_HPP_
#define RSM_BACKEND_BATCHER_HPP_

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "rsm.h"

namespace rsm_adapter {

struct BackEndConfig {
  std::string fine_name, coarse_name, pub_name;   // keys of SensorDataManager::multiresolution_range_data_ (slam_processor.h)
  double fine_resolution = 0.05, coarse_resolution = 0.1;
  double fine_deviation = 0.15, coarse_deviation = 0.3, gaussian_blur_offset = 0.88;
  float default_cell_prob = 0.3f;                  // kMapUnknownCellProb
  int fine_map_size = 480, coarse_map_size = 240;  // cells per side of the back-end maps
  rsm_pass_param pass[3];                          // back_end_scan_match_param_: coarse, fine, super-fine
  bool use_optimize_scan_match = false;
  rsm_optimize_param optimize;
  double optimize_failed_cost = 20.0;
  bool use_map_check = false;
  rsm_map_check_param map_check;
};

struct BackEndCandidate {
  int range_id = -1;                 // the scan to match (id in the SensorDataManager)
  std::vector<int> chain_ids;        // the chain it is matched against
  double centre[2] = {0.0, 0.0};     // where the back-end maps are centred: SlamProcessor::current_sensor_pose_ (:451-455)
  double pose[3] = {0.0, 0.0, 0.0};  // in: seed; out: matched pose
  double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double response = 0.0;             // out: ScanMatchInterface's return value
};

class BackEndBatcher {
 public:
  explicit BackEndBatcher(const BackEndConfig& cfg, int device = 0) : cfg_(cfg) {
    if (rsm_create(device, &ctx_) != RSM_OK) throw std::runtime_error("rsm_create failed: no CUDA device, and there is no CPU path");
    if (rsm_scan_store_create(ctx_, &fine_) != RSM_OK || rsm_scan_store_create(ctx_, &coarse_) != RSM_OK ||
        rsm_scan_store_create(ctx_, &pub_) != RSM_OK)
      throw std::runtime_error("rsm_scan_store_create failed");
  }
  ~BackEndBatcher() {
    if (pub_grid_) rsm_grid_destroy(ctx_, pub_grid_);
    for (rsm_scan_store* s : {fine_, coarse_, pub_}) if (s) rsm_scan_store_destroy(ctx_, s);
    if (ctx_) rsm_destroy(ctx_);
  }
  BackEndBatcher(const BackEndBatcher&) = delete;
  BackEndBatcher& operator=(const BackEndBatcher&) = delete;

  // New scans of the manager go to the device stores once; the sensor poses of all scans are refreshed.
  void SyncScans(roborts_slam::SensorDataManager& sdm) {
    const int n = sdm.current_data_index() + 1;
    for (int id = rsm_scan_store_size(fine_); id < n; ++id) {
      const Eigen::Vector3d pose = sdm.GetRangeData(id)->sensor_pose();
      const double p[3] = {pose[0], pose[1], pose[2]};
      Add(fine_, sdm.GetMultiresolutionRangeData(cfg_.fine_name, id), p);
      Add(coarse_, sdm.GetMultiresolutionRangeData(cfg_.coarse_name, id), p);
      Add(pub_, sdm.GetMultiresolutionRangeData(cfg_.pub_name, id), p);
    }
    ids_.resize(n);
    poses_.resize(3 * static_cast<size_t>(n));
    for (int id = 0; id < n; ++id) {
      const Eigen::Vector3d pose = sdm.GetRangeData(id)->sensor_pose();
      ids_[id] = id;
      for (int k = 0; k < 3; ++k) poses_[3 * id + k] = pose[k];
    }
    if (n > 0 && rsm_scan_store_set_poses(ctx_, fine_, n, ids_.data(), poses_.data()) != RSM_OK)
      throw std::runtime_error(std::string("rsm_scan_store_set_poses failed: ") + rsm_last_error(ctx_));
  }

  // Occupancy of the live publishing map for the map check (MapCheckPenalize, slam_processor.cpp:573-595): a cell counts
  // when GetGridStates(cell) == GridStates_Occupied (occu_grid_map.h:447-471).  Call after the map changed.
  template <typename PubMapT>
  void SyncPubMap(PubMapT& pub_map, double resolution, double offset_x, double offset_y) {
    const int sx = pub_map.GetSizeX(), sy = pub_map.GetSizeY();
    if (!pub_grid_ || sx != pub_sx_ || sy != pub_sy_) {
      if (pub_grid_) { rsm_grid_destroy(ctx_, pub_grid_); pub_grid_ = nullptr; }
      if (rsm_grid_create(ctx_, sx, sy, resolution, offset_x, offset_y, &pub_grid_) != RSM_OK)
        throw std::runtime_error(std::string("rsm_grid_create failed: ") + rsm_last_error(ctx_));
      pub_sx_ = sx; pub_sy_ = sy;
    } else {
      rsm_grid_set_offset(ctx_, pub_grid_, offset_x, offset_y);
    }
    occ_.resize(static_cast<size_t>(sx) * sy);
    for (int i = 0; i < sx * sy; ++i) occ_[i] = pub_map.GetGridStates(i) == roborts_slam::GridStates_Occupied ? 1 : 0;
    if (rsm_grid_upload_occupancy(ctx_, pub_grid_, occ_.data()) != RSM_OK)
      throw std::runtime_error(std::string("rsm_grid_upload_occupancy failed: ") + rsm_last_error(ctx_));
  }

  // SlamProcessor::ScanMatchInterface for every candidate, in one batched call.
  void ScanMatchBatch(std::vector<BackEndCandidate>& cands, bool use_fine_scan_match = true) {
    const int n = static_cast<int>(cands.size());
    if (n == 0) return;
    off_.assign(1, 0);
    chain_.clear(); match_.resize(n); centres_.resize(2 * static_cast<size_t>(n)); seed_.resize(3 * static_cast<size_t>(n));
    covs_.resize(9 * static_cast<size_t>(n)); scores_.assign(n, 0.0);
    for (int i = 0; i < n; ++i) {
      const BackEndCandidate& c = cands[i];
      chain_.insert(chain_.end(), c.chain_ids.begin(), c.chain_ids.end());
      off_.push_back(static_cast<int64_t>(chain_.size()));
      match_[i] = c.range_id;
      centres_[2 * i] = c.centre[0]; centres_[2 * i + 1] = c.centre[1];
      for (int k = 0; k < 3; ++k) seed_[3 * i + k] = c.pose[k];
      for (int k = 0; k < 9; ++k) covs_[9 * i + k] = c.cov[k];
    }
    const bool check = cfg_.use_map_check && pub_grid_;
    int rc;
    if (cfg_.use_optimize_scan_match)
      rc = rsm_scan_match_interface_batch_opt(ctx_, fine_, coarse_, n, cfg_.fine_map_size, cfg_.fine_resolution, cfg_.fine_deviation,
                                              cfg_.coarse_map_size, cfg_.coarse_resolution, cfg_.coarse_deviation, cfg_.default_cell_prob,
                                              cfg_.gaussian_blur_offset, centres_.data(), off_.data(), chain_.data(), match_.data(),
                                              cfg_.pass, &cfg_.optimize, cfg_.optimize_failed_cost, use_fine_scan_match ? 1 : 0,
                                              seed_.data(), covs_.data(), scores_.data(), nullptr, check ? pub_grid_ : nullptr,
                                              check ? pub_ : nullptr, check ? &cfg_.map_check : nullptr);
    else
      rc = rsm_scan_match_interface_batch(ctx_, fine_, n, cfg_.fine_map_size, cfg_.fine_resolution, cfg_.default_cell_prob,
                                          cfg_.fine_deviation, cfg_.gaussian_blur_offset, centres_.data(), off_.data(), chain_.data(),
                                          match_.data(), cfg_.pass, use_fine_scan_match ? 1 : 0, seed_.data(), covs_.data(),
                                          scores_.data(), nullptr, check ? pub_grid_ : nullptr, check ? pub_ : nullptr,
                                          check ? &cfg_.map_check : nullptr);
    if (rc != RSM_OK) throw std::runtime_error(std::string("batched ScanMatchInterface failed: ") + rsm_last_error(ctx_));
    for (int i = 0; i < n; ++i) {
      BackEndCandidate& c = cands[i];
      for (int k = 0; k < 3; ++k) c.pose[k] = seed_[3 * i + k];
      for (int k = 0; k < 9; ++k) c.cov[k] = covs_[9 * i + k];
      c.response = scores_[i];
    }
  }

  struct LoopClosureThresholds {
    double min_response_coarse = 0.7, max_variance_coarse = 0.16, min_response_fine = 0.7;   // loop_match_* parameters
  };

  // RangeScanPoseGraph::TryCloseLoop's acceptance test (range_scan_pose_graph.cpp:305-352) over the candidate chains of one
  // scan, in the order FindPossibleLoopClosure yields them: returns the index of the first chain that closes the loop
  // (best_pose / cov = the result of its second match) or -1.  stage1 / stage2 (optional) receive every response.
  int TryCloseLoop(int range_id, const std::vector<std::vector<int>>& chains, const double scan_pose[3], const double centre[2],
                   const LoopClosureThresholds& th, double best_pose[3], double cov[9], std::vector<double>* stage1 = nullptr,
                   std::vector<double>* stage2 = nullptr) {
    std::vector<BackEndCandidate> first(chains.size());
    for (size_t i = 0; i < chains.size(); ++i) {
      first[i].range_id = range_id; first[i].chain_ids = chains[i];
      first[i].centre[0] = centre[0]; first[i].centre[1] = centre[1];
      for (int k = 0; k < 3; ++k) first[i].pose[k] = scan_pose[k];          // best_pose = range_data_ptr->sensor_pose()  (:308)
    }
    ScanMatchBatch(first, true);                                             // coarse_response  (:312-318)
    std::vector<BackEndCandidate> second;
    std::vector<int> which;
    for (size_t i = 0; i < first.size(); ++i) {
      const BackEndCandidate& c = first[i];
      if (c.response > th.min_response_coarse && c.cov[0] < th.max_variance_coarse && c.cov[4] < th.max_variance_coarse) {   // :322-324
        second.push_back(c);                                                 // re-match from the refined pose and covariance (:326-329)
        which.push_back(static_cast<int>(i));
      }
    }
    ScanMatchBatch(second, true);
    if (stage1) { stage1->clear(); for (const auto& c : first) stage1->push_back(c.response); }
    if (stage2) { stage2->assign(chains.size(), -1.0); for (size_t k = 0; k < second.size(); ++k) (*stage2)[which[k]] = second[k].response; }
    for (size_t k = 0; k < second.size(); ++k) {
      if (second[k].response < th.min_response_fine) continue;              // REJECTED (:333-334)
      for (int j = 0; j < 3; ++j) best_pose[j] = second[k].pose[j];
      for (int j = 0; j < 9; ++j) cov[j] = second[k].cov[j];
      return which[k];
    }
    return -1;
  }

  rsm_ctx* context() const { return ctx_; }

 private:
  void Add(rsm_scan_store* store, const std::shared_ptr<roborts_slam::RangeDataContainer2d>& rd, const double pose[3]) {
    if (!rd) throw std::runtime_error("SensorDataManager holds no scan of that resolution for a new id");
    const int n = rd->GetSize();
    pts_.resize(2 * static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) {
      const Eigen::Vector2d& p = rd->GetDataPoint(i);
      pts_[2 * i] = p[0];
      pts_[2 * i + 1] = p[1];
    }
    if (rsm_scan_store_add(ctx_, store, pts_.data(), n, pose, nullptr) != RSM_OK)
      throw std::runtime_error(std::string("rsm_scan_store_add failed: ") + rsm_last_error(ctx_));
  }

  BackEndConfig cfg_;
  rsm_ctx* ctx_ = nullptr;
  rsm_scan_store *fine_ = nullptr, *coarse_ = nullptr, *pub_ = nullptr;
  rsm_grid* pub_grid_ = nullptr;
  int pub_sx_ = 0, pub_sy_ = 0;
  std::vector<int32_t> ids_, chain_, match_;
  std::vector<int64_t> off_;
  std::vector<double> poses_, pts_, centres_, seed_, covs_, scores_;
  std::vector<uint8_t> occ_;
};

}  // namespace rsm_adapter

#endif

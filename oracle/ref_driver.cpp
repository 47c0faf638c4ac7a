// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the
// product path (roborts_edu_slam_b200/).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built from
// this file.
//
// This translation unit #includes the UNMODIFIED reference headers where they lie
// under /root/reference/src (never copied into this repo) against the stand-in
// Eigen/glog/boost/ros headers in oracle/standin/, and exposes a small C ABI so the
// Python tests can drive the reference's own
//   * OccuGridMap<ProbabilityCell>::InitMapWithRangeVec   (occu_grid_map.h:222-329)
//   * MultiResolutionCorrelateScanMatcher::ScanMatch      (correlate_scan_matcher.h:516-614)
//   * BasedCorrelationScanMatch::ScanMatch                (correlate_scan_matcher.h:784-875)
//   * BasedOptimizeScanMatch::ScanMatch                   (optimize_scan_matcher.h:68-131; its
//     H.ldlt().solve(b) resolves to the stand-in's restatement of Eigen 3.3's LDLT, the one piece
//     of that function that is therefore NOT the reference's own arithmetic)
// on arbitrary inputs.  scan_matchers.h itself is not compiled (sensor_data_manager / ROS
// plumbing); its chain (scan_matchers.h:179-289, optimiser off and on) is driven here by calling
// the reference's BasedOptimizeScanMatch / BasedCorrelationScanMatch exactly as those lines do.
//
// Output: oracle/_ref/libref.so (git-ignored).  Built by oracle/Makefile.
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>

#include <cfloat>
#include <cmath>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <condition_variable>

// The front-end map tests need GridMapBase::map_offset_ (no getter) after a resize: open the reference's
// classes for this translation unit only (test infrastructure; layout is unaffected, nothing is modified).
#define private public
#define protected public
#include "scan_match/correlate_scan_matcher.h"
#include "scan_match/optimize_scan_matcher.h"
#include "map/slam_map.h"

#undef private
#undef protected

#include "ref_types.h"

using namespace roborts_slam;

namespace {

std::shared_ptr<RangeDataContainer2d> MakeScan(int n, const double* xy, const double* pose_world) {
  auto rd = std::make_shared<RangeDataContainer2d>(n > 0 ? n : 1);
  for (int i = 0; i < n; ++i) rd->AddDataPoint(Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]));
  rd->set_sensor_origin(Eigen::Vector2d(0.0, 0.0));
  if (pose_world) rd->set_sensor_pose(Eigen::Vector3d(pose_world[0], pose_world[1], pose_world[2]));
  else rd->set_sensor_pose(Eigen::Vector3d(0.0, 0.0, 0.0));
  return rd;
}

std::shared_ptr<CorrelationScanMatchParam> MakeParam(const double* p) {
  // p = {size, sres, aoff, ares, threshold, use_point_size, use_center_penalty, type}
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

// A back-end style scan-match map (slam_processor.cpp:428-462): constructed, fully Reset()
// to default_prob, auto-resize off, just_update_occu on.
void* ref_map_create(double resolution, int size_x, int size_y, double off_x, double off_y,
                     double deviation, float default_prob, double occu_offset) {
  auto* h = new RefMap;
  h->map = std::make_shared<ScanMatchMap>(resolution, Eigen::Vector2i(size_x, size_y),
                                          Eigen::Vector2d(off_x, off_y), deviation, default_prob);
  h->map->set_cell_occu_prob_offset(occu_offset);
  h->map->set_use_auto_map_resize(false);
  h->map->set_just_update_occu(true);
  std::vector<std::shared_ptr<RangeDataContainer2d>> none;
  h->map->InitMapWithRangeVec(none, false, false);  // Reset(): every cell = default_prob
  return h;
}

void ref_map_destroy(void* m) { delete static_cast<RefMap*>(m); }

void ref_map_set_offset(void* m, double off_x, double off_y) {
  static_cast<RefMap*>(m)->map->set_map_offset(Eigen::Vector2d(off_x, off_y));
}

// InitMapWithRangeVec(vec, use_blur, use_reset_speedup=true), as ResetScanMatchMapWithRangeVec does.
// pts: concatenated (x,y) in CELL units, sensor frame; poses: n_scans x (x,y,theta) world metres.
int ref_map_build(void* m, int n_scans, const int* n_pts, const double* pts, const double* poses,
                  int use_blur) {
  auto* h = static_cast<RefMap*>(m);
  std::vector<std::shared_ptr<RangeDataContainer2d>> vec;
  size_t off = 0;
  for (int s = 0; s < n_scans; ++s) {
    vec.push_back(MakeScan(n_pts[s], pts + 2 * off, poses + 3 * s));
    off += n_pts[s];
  }
  h->map->InitMapWithRangeVec(vec, use_blur != 0, true);
  return h->map->IsMapInit() ? 0 : 1;
}

void ref_map_size(void* m, int* sx, int* sy) {
  auto* h = static_cast<RefMap*>(m);
  *sx = h->map->GetSizeX();
  *sy = h->map->GetSizeY();
}

double ref_map_cell_length(void* m) { return static_cast<RefMap*>(m)->map->GetCellLength(); }

void ref_map_read(void* m, float* out) {
  auto* h = static_cast<RefMap*>(m);
  const int n = h->map->GetGridCellNum();
  for (int i = 0; i < n; ++i) out[i] = h->map->GetCell(i).GetValue();
}

// Overwrite every cell (tests of externally supplied grids) and mark the map initialised.
void ref_map_write(void* m, const float* in) {
  auto* h = static_cast<RefMap*>(m);
  const int n = h->map->GetGridCellNum();
  for (int i = 0; i < n; ++i) h->map->GetCell(i).SetValue(in[i]);
  if (!h->map->IsMapInit()) h->map->SetUpdated();
}

void ref_world_to_map(void* m, const double* w, double* out) {
  Eigen::Vector3d r = static_cast<RefMap*>(m)->map->GetMapCoordsPose(Eigen::Vector3d(w[0], w[1], w[2]));
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

void ref_map_to_world(void* m, const double* p, double* out) {
  Eigen::Vector3d r = static_cast<RefMap*>(m)->map->GetWorldCoordsPose(Eigen::Vector3d(p[0], p[1], p[2]));
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

// Number of candidates a pass generates (n_ang * n_xy^2), computed the reference's way.
long ref_candidate_count(const double* p) {
  int n_ang = static_cast<int>(std::floor(p[2] * 2 / p[3]) + 1);
  int n_xy = static_cast<int>(util::Round(p[0] / p[1]) + 1);
  return static_cast<long>(n_ang) * n_xy * n_xy;
}

// The inner matcher only: returns best score, fills the SORTED candidate list
// (x, y, angle, angle_index, score) as the reference leaves it, plus the averaged best.
// center_map is the seed pose already in map coordinates.
double ref_match_candidates(void* m, int n, const double* xy, const double* param,
                            const double* center_map, long cap, double* cx, double* cy,
                            double* cang, int* cidx, double* cscore, long* n_out, double* best) {
  auto* h = static_cast<RefMap*>(m);
  auto rd = MakeScan(n, xy, nullptr);
  auto q = MakeParam(param);
  auto lut = std::make_shared<AngleSearchLookUpTable>();
  MultiResolutionCorrelateScanMatcher matcher;
  Eigen::Vector3d center(center_map[0], center_map[1], center_map[2]);
  double s = matcher.ScanMatch(h->map, rd, q, lut, center);
  std::vector<Candidate2D> c = matcher.candidate_pose();
  long k = std::min<long>(cap, static_cast<long>(c.size()));
  for (long i = 0; i < k; ++i) {
    if (cx) cx[i] = c[i].x();
    if (cy) cy[i] = c[i].y();
    if (cang) cang[i] = c[i].angle();
    if (cidx) cidx[i] = static_cast<int>(c[i].angle_index());
    if (cscore) cscore[i] = c[i].score();
  }
  if (n_out) *n_out = static_cast<long>(c.size());
  Candidate2D b = matcher.best_candidate_pose();
  if (best) { best[0] = b.x(); best[1] = b.y(); best[2] = b.angle(); best[3] = b.score(); }
  return s;
}

// One full pass: BasedCorrelationScanMatch::ScanMatch.  pose_world and cov (row-major 3x3)
// are in/out exactly as in the reference.  seconds (optional) = wall time of the call.
double ref_match(void* m, int n, const double* xy, const double* param, double* pose_world,
                 double* cov, double* seconds) {
  auto* h = static_cast<RefMap*>(m);
  auto rd = MakeScan(n, xy, nullptr);
  auto q = MakeParam(param);
  BasedCorrelationScanMatch matcher;
  Eigen::Vector3d pose(pose_world[0], pose_world[1], pose_world[2]);
  Eigen::Matrix3d c;
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c(r, k) = cov[3 * r + k];
  auto t0 = std::chrono::steady_clock::now();
  double resp = matcher.ScanMatch(h->map, rd, q, pose, c);
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) cov[3 * r + k] = c(r, k);
  pose_world[0] = pose[0]; pose_world[1] = pose[1]; pose_world[2] = pose[2];
  return resp;
}

// The coarse -> fine -> super-fine chain of ScanMatchers::ScanMatch (scan_matchers.h:224-263,
// use_optimize_scan_match = false), all three passes on the same (fine) map with the same
// matcher object, pose and covariance threaded through in place; returns the mean response.
// params = 3 x 8 doubles (coarse, fine, super).  resp_out (optional) = the 3 pass responses.
double ref_match_chain(void* m, int n, const double* xy, const double* params, int use_fine,
                       double* pose_world, double* cov, double* resp_out, double* seconds) {
  auto* h = static_cast<RefMap*>(m);
  auto rd = MakeScan(n, xy, nullptr);
  BasedCorrelationScanMatch matcher;
  Eigen::Vector3d process_pose(pose_world[0], pose_world[1], pose_world[2]);
  Eigen::Matrix3d c;
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c(r, k) = cov[3 * r + k];
  double score = 0.0;
  int times = 0;
  auto t0 = std::chrono::steady_clock::now();
  double r0 = matcher.ScanMatch(h->map, rd, MakeParam(params), process_pose, c);
  score += r0; times++;
  double r1 = 0.0, r2 = 0.0;
  if (use_fine) {
    r1 = matcher.ScanMatch(h->map, rd, MakeParam(params + 8), process_pose, c);
    score += r1; times++;
    r2 = matcher.ScanMatch(h->map, rd, MakeParam(params + 16), process_pose, c);
    score += r2; times++;
  }
  score /= times;
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (resp_out) { resp_out[0] = r0; resp_out[1] = r1; resp_out[2] = r2; }
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) cov[3 * r + k] = c(r, k);
  pose_world[0] = process_pose[0]; pose_world[1] = process_pose[1]; pose_world[2] = process_pose[2];
  return score;
}

// BasedOptimizeScanMatch::ScanMatch (optimize_scan_matcher.h:68-131) on one map.
// op = {iterate_max_times, cost_decrease_threshold, cost_min_threshold, max_update_distance, max_update_angle}
static std::shared_ptr<OptimizeScanMatchParam> MakeOptParam(const double* op) {
  auto q = std::make_shared<OptimizeScanMatchParam>();
  q->set_iterate_max_times(static_cast<int>(op[0]));
  q->set_cost_decrease_threshold(op[1]);
  q->set_cost_min_threshold(op[2]);
  q->set_max_update_distance(op[3]);
  q->set_max_update_angle_(op[4]);
  return q;
}

double ref_optimize(void* m, int n, const double* xy, const double* op, double* pose_world) {
  auto* h = static_cast<RefMap*>(m);
  auto rd = MakeScan(n, xy, nullptr);
  BasedOptimizeScanMatch opt;
  Eigen::Vector3d pose(pose_world[0], pose_world[1], pose_world[2]);
  double cost = opt.ScanMatch(h->map, rd, MakeOptParam(op), pose);
  pose_world[0] = pose[0]; pose_world[1] = pose[1]; pose_world[2] = pose[2];
  return cost;
}

// ScanMatchers::ScanMatch with use_optimize_scan_match = true (scan_matchers.h:179-289; MapSizeCheck left to
// the caller): the optimiser runs on the coarse map with the coarse-resolution scan, the correlative passes on
// the fine map.  resp_out (optional) = {optimiser cost, coarse, fine, super responses} (0 where a step is skipped).
double ref_match_chain_opt(void* m_coarse, int n_c, const double* xy_c, void* m_fine, int n_f, const double* xy_f,
                           const double* params, const double* op, double optimize_failed_cost, int use_fine,
                           double* pose_world, double* cov, double* resp_out) {
  auto* hc = static_cast<RefMap*>(m_coarse);
  auto* hf = static_cast<RefMap*>(m_fine);
  auto rd_c = MakeScan(n_c, xy_c, nullptr);
  auto rd_f = MakeScan(n_f, xy_f, nullptr);
  BasedOptimizeScanMatch optimizer;
  BasedCorrelationScanMatch matcher;
  Eigen::Vector3d best_pose(pose_world[0], pose_world[1], pose_world[2]);
  Eigen::Matrix3d c;
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c(r, k) = cov[3 * r + k];
  double scan_match_score = 0.0;
  int scan_match_times = 0;
  Eigen::Vector3d process_pose(best_pose);
  double optimize_cost = optimizer.ScanMatch(hc->map, rd_c, MakeOptParam(op), process_pose);
  scan_match_score = optimize_failed_cost / (optimize_cost + optimize_failed_cost);
  scan_match_times++;
  double r0 = 0.0, r1 = 0.0, r2 = 0.0;
  if (!use_fine || optimize_cost > optimize_failed_cost) {
    scan_match_score = 0.0;
    scan_match_times--;
    process_pose = best_pose;
    r0 = matcher.ScanMatch(hf->map, rd_f, MakeParam(params), process_pose, c);
    scan_match_score += r0;
    scan_match_times++;
  }
  best_pose = process_pose;
  if (use_fine) {
    r1 = matcher.ScanMatch(hf->map, rd_f, MakeParam(params + 8), process_pose, c);
    scan_match_score += r1; scan_match_times++;
    r2 = matcher.ScanMatch(hf->map, rd_f, MakeParam(params + 16), process_pose, c);
    scan_match_score += r2; scan_match_times++;
  }
  best_pose = process_pose;
  scan_match_score /= scan_match_times;
  if (resp_out) { resp_out[0] = optimize_cost; resp_out[1] = r0; resp_out[2] = r1; resp_out[3] = r2; }
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) cov[3 * r + k] = c(r, k);
  pose_world[0] = best_pose[0]; pose_world[1] = best_pose[1]; pose_world[2] = best_pose[2];
  return scan_match_score;
}

// ---- front-end scan-match map: incremental UpdateMapByRange with auto-resize (slam_processor.cpp:466-512, 529-571) ----
// Constructed like CreateAllMap does (no Reset: cell 0 = default_prob, every other cell kDefaultCellProb = 0.5,
// grid_map_base.h:150-163), auto-resize on, just_update_occu on.
void* ref_frontend_map_create(double resolution, int size_x, int size_y, double off_x, double off_y, double deviation,
                              float default_prob, double occu_offset, double extend_factor) {
  auto* h = new RefMap;
  h->map = std::make_shared<ScanMatchMap>(resolution, Eigen::Vector2i(size_x, size_y), Eigen::Vector2d(off_x, off_y),
                                          deviation, default_prob);
  h->map->set_extend_factor(extend_factor);
  h->map->set_cell_occu_prob_offset(occu_offset);
  h->map->set_use_auto_map_resize(true);
  h->map->set_just_update_occu(true);
  return h;
}

// UpdateMapByRange(range_data, use_blur): 1 = the scan was stamped, 0 = the map was extended instead (the scan is
// dropped, occu_grid_map.h:296-300).  geom_out = {size_x, size_y, map_offset_x, map_offset_y}.
int ref_frontend_map_update(void* m, int n, const double* xy, const double* pose_world, int use_blur, double* geom_out) {
  auto* h = static_cast<RefMap*>(m);
  auto rd = MakeScan(n, xy, pose_world);
  const bool ok = h->map->UpdateMapByRange(rd, use_blur != 0);
  if (geom_out) {
    geom_out[0] = h->map->GetSizeX(); geom_out[1] = h->map->GetSizeY();
    geom_out[2] = h->map->map_offset_[0]; geom_out[3] = h->map->map_offset_[1];
  }
  return ok ? 1 : 0;
}

// MapSizeCheck's bound update (scan_matchers.h:365-390): 1 = inside, 0 = the map was extended.
int ref_frontend_map_size_check(void* m, const double* pose_world, double range_max, double offset, double* geom_out) {
  auto* h = static_cast<RefMap*>(m);
  Eigen::Vector3d center_pose = h->map->GetMapCoordsPose(Eigen::Vector3d(pose_world[0], pose_world[1], pose_world[2]));
  const double max_size = (range_max + offset) / h->map->GetCellLength();
  BoundBox2d box(Eigen::Vector2d((center_pose.x() - max_size), (center_pose.y() - max_size)),
                 Eigen::Vector2d((center_pose.x() + max_size), (center_pose.y() + max_size)));
  const bool ok = h->map->UpdateBound(box);
  if (geom_out) {
    geom_out[0] = h->map->GetSizeX(); geom_out[1] = h->map->GetSizeY();
    geom_out[2] = h->map->map_offset_[0]; geom_out[3] = h->map->map_offset_[1];
  }
  return ok ? 1 : 0;
}

// GaussianBlur kernel as the reference builds it (occu_grid_map.h:83-105): returns half size,
// fills k[(2h+1)^2] if non-null; -1 if the blur parameters are rejected.
int ref_blur_kernel(double sigma, double resolution, double* k, int cap) {
  GaussianBlur g(sigma, resolution);
  if (!g.GetBlurStates()) return -1;
  int n = g.GetKernelGridSize();
  if (k && cap >= n * n) std::memcpy(k, g.GetKernelValuePointer(), sizeof(double) * n * n);
  return g.GetKernelGridHalfSize();
}


// ---- publishing map + MapFeedbackResponsePenalty (occu_grid_map.h:331-392, 447-471) --------------
// A front-end style PubMap (CountCell): no blur, free space ray-traced, auto-resize off.
void* ref_pubmap_create(double resolution, int size_x, int size_y, double off_x, double off_y, float default_prob) {
  auto* m = new std::shared_ptr<PubMap>(std::make_shared<PubMap>(resolution, Eigen::Vector2i(size_x, size_y),
                                                                 Eigen::Vector2d(off_x, off_y), 0.0, default_prob));
  (*m)->set_use_auto_map_resize(false);
  (*m)->set_just_update_occu(false);
  std::vector<std::shared_ptr<RangeDataContainer2d>> none;
  (*m)->InitMapWithRangeVec(none, false, false);
  return m;
}

void ref_pubmap_destroy(void* m) { delete static_cast<std::shared_ptr<PubMap>*>(m); }

// UpdateMapByRange(scan, use_blur=false): pts in CELL units of this map, sensor frame; pose world metres.
int ref_pubmap_update(void* m, int n_pts, const double* pts, const double* pose_world) {
  auto& map = *static_cast<std::shared_ptr<PubMap>*>(m);
  return map->UpdateMapByRange(MakeScan(n_pts, pts, pose_world), false) ? 0 : 1;
}

// Per cell: prob_value_, pass_count_ and the occupancy the no-blur check uses
// (CountCellFunctions::GetGridStates == GridStates_Occupied, grid_map_cell.h:125-136).
void ref_pubmap_read(void* m, float* value, float* pass_count, unsigned char* occupied) {
  auto& map = *static_cast<std::shared_ptr<PubMap>*>(m);
  CountCellFunctions f;
  const int n = map->GetGridCellNum();
  for (int i = 0; i < n; ++i) {
    auto& c = map->GetCell(i);
    if (value) value[i] = c.GetValue();
    if (pass_count) pass_count[i] = c.pass_count();
    if (occupied) occupied[i] = f.GetGridStates(c) == GridStates_Occupied ? 1 : 0;
  }
}

// The publishing map as the front end keeps it (slam_processor.cpp:477-483, 538-553): constructed, never Reset,
// auto-resize on, ray-traced free space, CountCellFunctions knobs set before every update.
void* ref_pubmap_create_frontend(double resolution, int size_x, int size_y, double off_x, double off_y, double extend_factor) {
  auto* m = new std::shared_ptr<PubMap>(std::make_shared<PubMap>(resolution, Eigen::Vector2i(size_x, size_y),
                                                                 Eigen::Vector2d(off_x, off_y)));
  (*m)->set_extend_factor(extend_factor);
  (*m)->set_use_auto_map_resize(true);
  return m;
}

void ref_pubmap_set_factors(void* m, float free_factor, float occu_factor, float occu_threshold, float min_pass_through) {
  auto& map = *static_cast<std::shared_ptr<PubMap>*>(m);
  map->SetOccuThreshold(occu_threshold);
  map->SetMinPassThrough(min_pass_through);
  map->SetUpdateFreeFactor(free_factor);
  map->SetUpdateOccupiedFactor(occu_factor);
}

// 1 = updated, 0 = extended instead (scan dropped).  geom_out = {size_x, size_y, map_offset_x, map_offset_y}.
int ref_pubmap_update_geom(void* m, int n_pts, const double* pts, const double* pose_world, double* geom_out) {
  auto& map = *static_cast<std::shared_ptr<PubMap>*>(m);
  const bool ok = map->UpdateMapByRange(MakeScan(n_pts, pts, pose_world), false);
  if (geom_out) {
    geom_out[0] = map->GetSizeX(); geom_out[1] = map->GetSizeY();
    geom_out[2] = map->map_offset_[0]; geom_out[3] = map->map_offset_[1];
  }
  return ok ? 1 : 0;
}

// Per cell: value, pass count, hit count and the map's own GetGridStates == Occupied (with the knobs set above).
void ref_pubmap_read_all(void* m, float* value, float* pass_count, float* hit_count, unsigned char* occupied) {
  auto& map = *static_cast<std::shared_ptr<PubMap>*>(m);
  const int n = map->GetGridCellNum();
  for (int i = 0; i < n; ++i) {
    auto& c = map->GetCell(i);
    if (value) value[i] = c.GetValue();
    if (pass_count) pass_count[i] = c.pass_count_;
    if (hit_count) hit_count[i] = c.hit_count_;
    if (occupied) occupied[i] = map->GetGridStates(i) == GridStates_Occupied ? 1 : 0;
  }
}

// origin: RangeDataContainer::sensor_origin() (NULL = (0,0)).  use_logistic re-states the one line of
// SlamProcessor::MapCheckPenalize (slam/slam_processor.cpp:589-591; that file needs ROS and cannot be
// compiled here) on top of the reference's own MapFeedbackResponsePenalty.
double ref_pubmap_penalty(void* m, int n_pts, const double* pts, const double* pose_world, const double* origin,
                          int check_point_num, double bound_tolerance, double penalty_gain, int use_blur, int use_logistic) {
  auto& map = *static_cast<std::shared_ptr<PubMap>*>(m);
  auto scan = MakeScan(n_pts, pts, nullptr);
  if (origin) scan->set_sensor_origin(Eigen::Vector2d(origin[0], origin[1]));
  double penalty = map->MapFeedbackResponsePenalty(scan, Eigen::Vector3d(pose_world[0], pose_world[1], pose_world[2]),
                                                   check_point_num, bound_tolerance, penalty_gain, use_blur != 0);
  if (use_logistic) penalty = (1 / (1 + exp(-10 * (penalty - 0.4))));
  return penalty;
}

}  // extern "C"

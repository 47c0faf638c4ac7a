/* rsm.h -- C ABI of the B200-native correlative scan matcher (librsm.so).
 *
 * Drop-in boundary for ONE hot path of KevinLADLee/RoboRTS-Edu-SLAM: the brute-force
 * correlative scan matcher in src/scan_match/correlate_scan_matcher.h plus the lookup-grid
 * construction it reads from (src/map/occu_grid_map.h).  The reference has no FFI layer of its
 * own; each entry point below cites the reference interface it replaces
 * (file:line under the reference's src/).  INTEGRATION.md shows the C++ adapter a maintainer
 * would drop into the reference so that scan_matchers.h / slam_processor.cpp compile unchanged.
 *
 * Conventions
 *   - plain pointers and sizes only; every call returns an rsm_status (0 = OK) and never throws;
 *   - all host arrays are caller-owned and only borrowed for the duration of the call;
 *   - scan points are (x,y) pairs in CELL units in the sensor frame, exactly what
 *     RangeDataContainer::CreateFrom(scan, 1/resolution) holds (slam/sensor_data_manager.h:99-115);
 *   - poses are (x, y, theta) in world metres / radians; covariances are row-major 3x3;
 *   - a context is not re-entrant (the reference serialises callers with scan_match_mutex_,
 *     scan_matchers.h:298-300); use one context per GPU / per caller thread;
 *   - there is NO CPU fallback: without a CUDA device rsm_create fails.
 *
 * Results are identical to the reference's CPU matcher on the same inputs: bit-exact cell
 * indices, candidate scores, response and best pose; covariance within 1e-6 relative (bit-equal
 * unless equal scores tie inside a covariance prefix; RSM_OPT_STRICT_TIES routes those to the exact path).
 * "The reference" = its own headers compiled against a stand-in for Eigen (oracle/standin): the
 * Affine2d products of the map transforms and the 3 x 3 LDLT of the Gauss-Newton step follow Eigen 3.3's
 * evaluation order as restated there, unverified against real Eigen (none in the build image).
 */
#ifndef RSM_H_
#define RSM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsm_ctx rsm_ctx;   /* opaque: device buffers, streams, scratch */
typedef struct rsm_grid rsm_grid; /* opaque: one device-resident lookup grid (ScanMatchMap) */

typedef enum rsm_status {
  RSM_OK = 0,
  RSM_ERR_NO_DEVICE = 1,       /* no CUDA device / driver: there is no CPU path */
  RSM_ERR_CUDA = 2,            /* a CUDA runtime call failed; see rsm_last_error */
  RSM_ERR_INVALID = 3,         /* bad argument */
  RSM_ERR_WINDOW = 4,          /* search window + scan extent leaves the grid (the reference would
                                  read out of bounds here; its callers prevent it with MapSizeCheck,
                                  scan_matchers.h:365-390) */
  RSM_ERR_UNSUPPORTED = 5,     /* e.g. FAST (branch-and-bound) pass type */
  RSM_ERR_NOT_INIT = 6,        /* grid has no content yet (reference: !IsMapInit()) */
  RSM_NEED_EXACT = 7           /* rsm_match_finish only: a consumed candidate set holds exact score ties whose order
                                  the reference's unstable std::sort decides (correlate_scan_matcher.h:607-608);
                                  nothing was written -- gather every rank's slice (rsm_match_slice_scores) and call
                                  rsm_match_finish_exact, which runs that same sort on the whole array */
} rsm_status;

/* CorrelationScanMatchType, correlate_scan_matcher.h:34-39 */
enum { RSM_COARSE = 0, RSM_FINE = 1, RSM_SUPER = 2, RSM_FAST = 3 };

/* CorrelationScanMatchParam, correlate_scan_matcher.h:41-86 (max_depth only matters to FAST) */
typedef struct rsm_pass_param {
  double search_space_size;       /* full width of the square translation window, metres */
  double search_space_resolution; /* metres */
  double search_angle_offset;     /* half range, radians */
  double search_angle_resolution; /* radians */
  double response_threshold;
  int32_t use_point_size;
  int32_t use_center_penalty;     /* bool */
  int32_t type;                   /* RSM_COARSE / RSM_FINE / RSM_SUPER */
  int32_t reserved;
} rsm_pass_param;

/* Optional per-pass detail (diagnostics and parity tests). */
typedef struct rsm_pass_detail {
  double best_score;      /* top candidate score, not clamped */
  double best_pose_map[3];/* (averaged) best candidate in map cells / rad */
  int64_t n_candidates;
  int32_t n_avg;          /* size of the averaging set of FindBestCandidate (:670-710) */
  int32_t exact_sort_used;/* 1 if exact ties forced the host std::sort path (same result, slower) */
  int32_t pose_updated;   /* response > threshold */
  int32_t n_ang, n_xy, visited, divisor;
  int32_t reserved;
} rsm_pass_detail;

/* Counters since the context was created or last reset. */
typedef struct rsm_stats {
  int64_t kernel_launches;      /* kernels of this library launched */
  int64_t score_launches;       /* launches of the scoring kernel */
  int64_t evals;                /* candidate-beam evaluations scored (n_ang*n_xy^2*visited) */
  int64_t passes;               /* single passes completed */
  int64_t exact_sort_passes;    /* passes that needed the exact-tie host sort */
  int64_t h2d_bytes, d2h_bytes; /* bytes copied by this library */
  double score_kernel_ms;       /* CUDA-event time of scoring kernels (only while profiling is on) */
  double raster_kernel_ms;      /* CUDA-event time of rasterisation kernels (profiling on) */
  double select_kernel_ms;      /* CUDA-event time of selection kernels (profiling on) */
  /* host wall-clock per phase of a pass, summed (always on): 0 geometry + angle tables + upload,
   * 1 wait for scoring + selection, 2 host stage 1 (best pose, positional covariance),
   * 3 same-(x,y) gather round trip, 4 host stage 2 (angular covariance), 5 exact-sort path,
   * 6 set-up of a batched back-end step (grid slots, chain lookup), 7 its pipelined chain as a whole (wall clock; with
   * several lanes the phases 0-5 of the lanes overlap and add up to more than this) */
  double phase_ms[8];
} rsm_stats;

/* ---- context ------------------------------------------------------------------------- */
int rsm_create(int device, rsm_ctx** out);
void rsm_destroy(rsm_ctx* ctx);
const char* rsm_last_error(const rsm_ctx* ctx);
const char* rsm_version(void);
int rsm_set_profiling(rsm_ctx* ctx, int on); /* record CUDA events around every kernel class */
int rsm_get_stats(rsm_ctx* ctx, rsm_stats* out);
int rsm_reset_stats(rsm_ctx* ctx);
int rsm_synchronize(rsm_ctx* ctx);
/* Context options.  RSM_OPT_STRICT_TIES (default 0): the fast path already takes the exact std::sort path whenever
 * the ORDER of exactly tied scores could change which candidates are consumed (ties inside the averaging set, at
 * the 20/21 covariance cuts).  Ties strictly inside a 20-element covariance prefix keep the same set but may add
 * the terms in another order than the reference's sort permutation: the covariance then agrees to rounding
 * (<= 1e-6 relative, the contract) instead of bit for bit.  With the option on those passes take the exact path
 * too (bit-equal covariance under every tie pattern; a 1.2 M-candidate pass then costs a 50 ms host sort).
 * RSM_OPT_LANES (default 0 = automatic): sub-batches a batched chain call is cut into, pipelined over as many
 * streams so that one sub-batch's host finalisation overlaps another's kernels (1 = no pipelining). */
enum { RSM_OPT_STRICT_TIES = 1, RSM_OPT_LANES = 2 };
int rsm_set_option(rsm_ctx* ctx, int option, int value);
/* CUDA-event stopwatch on the context's own stream (what bench.py times with). */
int rsm_timer_start(rsm_ctx* ctx);
int rsm_timer_stop(rsm_ctx* ctx, double* elapsed_ms);
/* Overwrite a scratch buffer larger than L2 (126 MB) so the next call starts cold. */
int rsm_flush_l2(rsm_ctx* ctx);

/* ---- lookup grid ------------------------------------------------------------------------
 * Replaces the live ScanMatchMap the reference matcher reads through GetGridProbValue
 * (map/occu_grid_map.h:395-397) and its world<->map transform (map/grid_map_base.h:68-93).
 * offset is GridMapBase::map_offset_ (metres). */
int rsm_grid_create(rsm_ctx* ctx, int size_x, int size_y, double resolution, double offset_x,
                    double offset_y, rsm_grid** out);
/* Same, from GridMapBase::get_scale_factor() (= 1/resolution as the map computed it) instead of
 * the resolution, so that an adapter holding a live reference map reproduces its cell length bit
 * for bit (map/grid_map_base.h:297-309). */
int rsm_grid_create_from_scale(rsm_ctx* ctx, int size_x, int size_y, double scale_factor,
                               double offset_x, double offset_y, rsm_grid** out);
void rsm_grid_destroy(rsm_ctx* ctx, rsm_grid* grid);
int rsm_grid_set_offset(rsm_ctx* ctx, rsm_grid* grid, double offset_x, double offset_y);
/* Hand over an existing grid: prob[y*size_x + x] = ProbabilityCell::prob_value_ (map/grid_map_cell.h:301-328). */
int rsm_grid_upload_f32(rsm_ctx* ctx, rsm_grid* grid, const float* prob);
/* Occupancy of a publishing map (PubMap = OccuGridMap<CountCell>, map/slam_map.h:35) for the map
 * check below: occupied[y*size_x + x] != 0 where CheckOccuLineVisitorCallback would count the cell
 * (map/occu_grid_map.h:447-471): CountCellFunctions::GetGridStates(cell) == GridStates_Occupied
 * (pass_count >= 2 and value >= 0.5, map/grid_map_cell.h:125-136) without blur, value >
 * cell_occu_prob_offset with blur.  The grid object only lends its size and world<->map transform. */
int rsm_grid_upload_occupancy(rsm_ctx* ctx, rsm_grid* grid, const uint8_t* occupied);
/* OccuGridMap::MapFeedbackResponsePenalty (map/occu_grid_map.h:331-392) through
 * SlamProcessor::MapCheckPenalize (slam/slam_processor.cpp:573-595), for n candidate poses in one
 * launch: coeff_out[i] = max(1 + 2*gain - gain * (rays of scan i blocked by an occupied cell farther
 * than bound_tolerance from the ray's end), 0.1), 0.0 for a pose outside the map, and through
 * 1/(1+exp(-10*(coeff-0.4))) when use_logistic (the loop-closure call, slam_processor.cpp:313-317).
 * Scan i = pts_xy[2*pts_begin[i] .. 2*(pts_begin[i]+pts_count[i])) in cells of the publishing map,
 * sensor frame (ranges may overlap or coincide); sensor_origin_xy = RangeDataContainer::sensor_origin()
 * (NULL = (0,0)).  check_point_num / bound_tolerance / penalty_gain = map_check_* parameters. */
int rsm_map_check_penalize(rsm_ctx* ctx, const rsm_grid* pub_map, int n, const double* poses_world,
                           const double* pts_xy, const int64_t* pts_begin, const int32_t* pts_count,
                           const double* sensor_origin_xy, int check_point_num, double bound_tolerance,
                           double penalty_gain, int use_logistic, double* coeff_out);
/* Device-side construction from base scans; replaces OccuGridMap::InitMapWithRangeVec in the
 * back-end configuration (just_update_occu, no auto-resize; map/occu_grid_map.h:222-329,
 * 474-497, 531-576; slam/slam_processor.cpp:448-462).  sigma = map deviation, occu_offset =
 * gaussian_blur_offset.  pts_xy = concatenated scan points (cells, sensor frame), n_pts[i] points
 * for scan i, poses_world = n_scans x (x,y,theta) sensor poses. */
int rsm_grid_rasterize(rsm_ctx* ctx, rsm_grid* grid, float default_prob, double sigma,
                       double occu_offset, int use_blur, int n_scans, const int32_t* n_pts,
                       const double* pts_xy, const double* poses_world);
/* ---- front-end map maintenance ----------------------------------------
 * The front-end scan-match maps are constructed once and then only stamped (just_update_occu,
 * slam/slam_processor.cpp:466-512) by one UpdateMapByRange per accepted scan (:529-571); keeping them
 * on the device replaces the whole-map upload an adapter would otherwise make after every scan.
 * The resize POLICY (bound box bookkeeping, GridMapBase::UpdateBound / ExtendSize,
 * map/grid_map_base.h:188-274) stays with the caller -- the reference's own map object or a restatement
 * -- which tells the device map what happened:
 *   rsm_grid_fill             a constructed map: new CellType[n]{default_prob} gives cell 0 the map's
 *                             default_prob and every other cell kDefaultCellProb = 0.5
 *                             (map/grid_map_base.h:150-163, map/grid_map_cell.h:30); not yet IsMapInit().
 *   rsm_grid_update_by_range  UpdateMapByRange(range_data, use_blur) returned true: stamp that scan
 *                             (map/occu_grid_map.h:258-329, 531-576); no reset.
 *   rsm_grid_extend           UpdateMapByRange / MapSizeCheck extended the map (and dropped the scan,
 *                             :296-300): new size, where the old cell (0,0) went
 *                             (pre_grid_offset = -GetFloorMin()), the new map_offset_.  New cells get
 *                             fill_prob, cell 0 first_cell_prob, then the old rows are copied in
 *                             (grid_map_base.h:222-238).  Drops an uploaded occupancy mask. */
/* The resize policy itself, restated on the host (no device, no context needed): GridMapBase's bound-box
 * bookkeeping -- UpdateBound / ExtendSize(EXTEND_PARTLY) (map/grid_map_base.h:188-274, util/boundbox.h) -- as
 * UpdateMapByRange (map/occu_grid_map.h:278-300) and MapSizeCheck (scan_match/scan_matchers.h:365-390) drive
 * it.  One object per map, fed every scan / check in the reference's order.  *fits = 1: the scan is to be
 * stamped (rsm_grid_update_by_range / rsm_pubmap_update_by_range); *fits = 0: the map was extended instead
 * and the scan is dropped, as in the reference -- `geometry` then holds what rsm_grid_extend /
 * rsm_pubmap_extend need.  half_kernel = rsm_blur_half_size(deviation, resolution) of the map (0 for the
 * publishing map), use_blur as passed to UpdateMapByRange.  With this a caller needs no reference map object. */
typedef struct rsm_map_bounds rsm_map_bounds;
typedef struct rsm_map_geometry {
  int32_t size_x, size_y;                         /* map size after the call */
  int32_t pre_grid_offset_x, pre_grid_offset_y;   /* where the old cell (0,0) went in the last extension */
  double offset_x, offset_y;                      /* GridMapBase::map_offset_ after the call */
} rsm_map_geometry;
int rsm_blur_half_size(double sigma, double resolution);   /* GaussianBlur half size, -1 if rejected (occu_grid_map.h:40-59) */
int rsm_map_bounds_create(int size_x, int size_y, double scale_factor, double offset_x, double offset_y,
                          double extend_factor, rsm_map_bounds** out);
void rsm_map_bounds_destroy(rsm_map_bounds* bounds);
int rsm_map_bounds_update_scan(rsm_map_bounds* bounds, const double* pts_xy, int n_pts, const double pose_world[3],
                               int half_kernel, int use_blur, int* fits, rsm_map_geometry* geometry /* nullable */);
int rsm_map_bounds_size_check(rsm_map_bounds* bounds, const double pose_world[3], double range_max, double offset,
                              int* fits, rsm_map_geometry* geometry /* nullable */);
int rsm_grid_fill(rsm_ctx* ctx, rsm_grid* grid, float fill_prob, float first_cell_prob);
int rsm_grid_update_by_range(rsm_ctx* ctx, rsm_grid* grid, double sigma, double occu_offset, int use_blur,
                             const double* pts_xy, int n_pts, const double pose_world[3]);
/* The same for a caller that owns the world<->map transform (the C++ adapter mirrors a live reference map, whose
 * map_offset_ is private: it passes GetMapCoordsPose(sensor_pose), map/grid_map_base.h:83-87). */
int rsm_grid_update_by_range_map(rsm_ctx* ctx, rsm_grid* grid, double sigma, double occu_offset, int use_blur,
                                 const double* pts_xy, int n_pts, const double pose_map[3]);
int rsm_grid_extend(rsm_ctx* ctx, rsm_grid* grid, int new_size_x, int new_size_y, int pre_grid_offset_x,
                    int pre_grid_offset_y, double new_offset_x, double new_offset_y, float fill_prob,
                    float first_cell_prob);
int rsm_grid_geometry(const rsm_grid* grid, int* size_x, int* size_y, double* offset_x, double* offset_y);
int rsm_grid_download_f32(rsm_ctx* ctx, rsm_grid* grid, float* prob_out);

/* Publishing map on the device: PubMap = OccuGridMap<CountCell> (map/slam_map.h:35) as hit / pass / value
 * planes.  rsm_pubmap_update_by_range is UpdateMapByRange(range_data) with ray-traced free space
 * (map/occu_grid_map.h:258-329, 125-187, 474-530; map/grid_map_cell.h:92-108) for a scan the reference's
 * resize policy accepted; update_free_factor / update_occu_factor are the CountCellFunctions knobs
 * SlamProcessor::UpdateMap sets before every update (slam/slam_processor.cpp:538-551).  rsm_pubmap_extend
 * mirrors ExtendSize like rsm_grid_extend.  rsm_pubmap_refresh_occupancy evaluates
 * GetGridStates == Occupied (pass_count >= min_pass_through and value >= occu_threshold,
 * map/grid_map_cell.h:125-136) into the mask rsm_map_check_penalize reads: pass
 * rsm_pubmap_check_grid(pm) as its pub_map -- no occupancy upload. */
typedef struct rsm_pubmap rsm_pubmap;
int rsm_pubmap_create(rsm_ctx* ctx, int size_x, int size_y, double resolution, double offset_x, double offset_y,
                      float default_prob, rsm_pubmap** out);
void rsm_pubmap_destroy(rsm_ctx* ctx, rsm_pubmap* pm);
int rsm_pubmap_update_by_range(rsm_ctx* ctx, rsm_pubmap* pm, const double* pts_xy, int n_pts,
                               const double pose_world[3], float update_free_factor, float update_occu_factor);
int rsm_pubmap_extend(rsm_ctx* ctx, rsm_pubmap* pm, int new_size_x, int new_size_y, int pre_grid_offset_x,
                      int pre_grid_offset_y, double new_offset_x, double new_offset_y);
int rsm_pubmap_refresh_occupancy(rsm_ctx* ctx, rsm_pubmap* pm, float occu_threshold, float min_pass_through);
const rsm_grid* rsm_pubmap_check_grid(const rsm_pubmap* pm);
int rsm_pubmap_download(rsm_ctx* ctx, const rsm_pubmap* pm, float* prob_out /* nullable */,
                        float* pass_out /* nullable */, float* hit_out /* nullable */,
                        uint8_t* occupied_out /* nullable */);
/* 1 if the grid is held as exact 2^-25 fixed point (integer gather path), 0 if as float32. */
int rsm_grid_is_fixed_point(const rsm_grid* grid);
int rsm_world_to_map(const rsm_grid* grid, const double pose_world[3], double pose_map[3]);
int rsm_map_to_world(const rsm_grid* grid, const double pose_map[3], double pose_world[3]);

/* ---- scans ------------------------------------------------------------------------------
 * A scan kept resident on the device.  The reference matches the same RangeDataContainer2d
 * several times (three passes per chain, one chain per loop-closure candidate); uploading it
 * once replaces the shared_ptr<RangeDataContainer2d> it passes around
 * (slam/sensor_data_manager.h:88-343).  pts_xy as everywhere: cells, sensor frame. */
typedef struct rsm_scan rsm_scan;
int rsm_scan_create(rsm_ctx* ctx, const double* pts_xy, int n_pts, rsm_scan** out);
void rsm_scan_destroy(rsm_ctx* ctx, rsm_scan* scan);

/* ---- matching ---------------------------------------------------------------------------
 * rsm_match: one pass.  Replaces BasedCorrelationScanMatch::ScanMatch(map, range_data, param,
 * current_pose&, cov_matrix&) -> response (scan_match/correlate_scan_matcher.h:784-875).
 * pose_world and cov are in/out exactly as there: cov is rewritten according to the pass type,
 * pose_world only if response > response_threshold.  Invalid input (grid not initialised or
 * n_pts == 0) returns RSM_OK with *response = 0 and the outputs untouched, like :792-795. */
int rsm_match(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts,
              const rsm_pass_param* param, double pose_world[3], double cov[9], double* response,
              rsm_pass_detail* detail /* nullable */);

/* rsm_match_map: the same pass for a caller that owns the world<->map transform (the C++ adapter
 * calls the reference map's own GetMapCoordsPose / GetWorldCoordsPose, map/grid_map_base.h:83-93).
 * center_map = seed pose in map cells / rad; cov in/out as in rsm_match; best_map_out = (averaged)
 * best candidate in map coordinates, which the caller converts and adopts iff *response exceeds
 * its threshold (:866-869). */
int rsm_match_map(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts,
                  const rsm_pass_param* param, const double center_map[3], double cov[9],
                  double* response, double best_map_out[3], rsm_pass_detail* detail /* nullable */);

/* Same pass on a device-resident scan (no host->device copy of the points). */
int rsm_match_resident(rsm_ctx* ctx, const rsm_grid* grid, const rsm_scan* scan,
                       const rsm_pass_param* param, double pose_world[3], double cov[9],
                       double* response, rsm_pass_detail* detail /* nullable */);

/* rsm_match_chain: coarse -> fine -> super-fine on one grid.  Replaces ScanMatchers::ScanMatch
 * with the optimiser off (scan_match/scan_matchers.h:179-289); params[0..2] = coarse, fine,
 * super.  *score = mean of the pass responses; responses (nullable) = the individual ones. */
int rsm_match_chain(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts,
                    const rsm_pass_param params[3], int use_fine, double pose_world[3],
                    double cov[9], double* score, double responses[3] /* nullable */);

/* rsm_match_batch: n independent chains in batched launches (no reference equivalent; the
 * reference runs loop-closure candidates one at a time, pose_graph/range_scan_pose_graph.cpp:
 * 153,312,329).  Pair i uses grids[i], points pts_xy[pts_offset[i] .. pts_offset[i+1]) (offsets
 * in points), params[3*i .. 3*i+2] (or params[0..2] for every pair when shared_params != 0),
 * poses_world[3*i..], covs[9*i..]; scores[i] = mean response, responses (nullable) 3 per pair. */
int rsm_match_batch(rsm_ctx* ctx, int n, const rsm_grid* const* grids, const double* pts_xy,
                    const int64_t* pts_offset, const rsm_pass_param* params, int shared_params,
                    int use_fine, double* poses_world, double* covs, double* scores,
                    double* responses /* nullable */);

/* rsm_loop_closure_batch: the whole back-end ScanMatchInterface step for n (scan, chain) pairs
 * (slam/slam_processor.cpp:250-326 without the pub-map penalty): for every pair rasterise a
 * size x size grid centred on centres_world[2*i..] from its base scans, then run the chain.
 * Base scans of pair i are scans [scan_offset[i], scan_offset[i+1]) of the base arrays. */
int rsm_loop_closure_batch(rsm_ctx* ctx, int n, int grid_size, double resolution, float default_prob,
                           double sigma, double occu_offset, const double* centres_world,
                           const int64_t* scan_offset, const int32_t* base_n_pts,
                           const double* base_pts_xy, const double* base_poses_world,
                           const double* pts_xy, const int64_t* pts_offset,
                           const rsm_pass_param params[3], int use_fine, double* poses_world,
                           double* covs, double* scores, double* responses /* nullable */);

/* ---- scan store and the batched back-end step -------------------------
 * A scan store is the device-resident counterpart of one resolution of
 * SensorDataManager::multiresolution_range_data_[name] (slam/sensor_data_manager.h:514-525): every
 * accepted scan is added once (points in cells of that resolution, sensor frame -- what
 * RangeDataContainer::CreateFrom(scan, 1/resolution) holds, :99-115 -- plus its sensor pose) and is
 * addressed by the id the reference uses for it.  Poses change when the pose graph is optimised
 * (SlamProcessor::UpdateRangeData, slam/slam_processor.cpp:597-603): rsm_scan_store_set_poses. */
typedef struct rsm_scan_store rsm_scan_store;
int rsm_scan_store_create(rsm_ctx* ctx, rsm_scan_store** out);
void rsm_scan_store_destroy(rsm_ctx* ctx, rsm_scan_store* store);
int rsm_scan_store_add(rsm_ctx* ctx, rsm_scan_store* store, const double* pts_xy, int n_pts,
                       const double pose_world[3], int32_t* id_out /* nullable */);
int rsm_scan_store_set_poses(rsm_ctx* ctx, rsm_scan_store* store, int n, const int32_t* ids,
                             const double* poses_world);
int rsm_scan_store_get_pose(const rsm_scan_store* store, int32_t id, double pose_world[3]);
int rsm_scan_store_size(const rsm_scan_store* store);

/* map_check_* parameters of SlamProcessor::MapCheckPenalize (slam/slam_processor.cpp:573-595) */
typedef struct rsm_map_check_param {
  double bound_tolerance;
  double penalty_gain;
  int32_t check_point_num;
  int32_t use_logistic;
} rsm_map_check_param;

/* rsm_scan_match_interface_batch: SlamProcessor::ScanMatchInterface (slam/slam_processor.cpp:250-326)
 * for n loop-closure candidates in batched launches, with scans named by id the way the pose graph
 * names them (ScanMatchFunc(range_data, closest_id, range_id, pose&, cov&, ...),
 * pose_graph/range_scan_pose_graph.h:30-35; called once per candidate chain at
 * pose_graph/range_scan_pose_graph.cpp:153,312,329).  Candidate i matches scan match_ids[i] against
 * the chain chain_ids[chain_offset[i] .. chain_offset[i+1]): a grid_size^2 grid centred on
 * centres_world[2i..] is reset from the chain's scans at their stored poses (:448-462), then the
 * coarse/fine/super chain runs from poses_world[3i..] (in/out; covs in/out as in rsm_match_batch).
 * Only ids travel host->device: the points are already resident.
 * With pub_map != NULL the step ends like the reference's: scores[i] *= MapCheckPenalize(scan
 * match_ids[i] of pub_store -- the same scan in cells of the publishing map --, matched pose,
 * use_logistic), clamped to 1 (:313-317).  responses (nullable) = the three pass responses. */
int rsm_scan_match_interface_batch(rsm_ctx* ctx, const rsm_scan_store* store, int n, int grid_size,
                                   double resolution, float default_prob, double sigma, double occu_offset,
                                   const double* centres_world, const int64_t* chain_offset,
                                   const int32_t* chain_ids, const int32_t* match_ids,
                                   const rsm_pass_param params[3], int use_fine, double* poses_world,
                                   double* covs, double* scores, double* responses /* nullable */,
                                   const rsm_grid* pub_map /* nullable */,
                                   const rsm_scan_store* pub_store /* nullable */,
                                   const rsm_map_check_param* check /* nullable */);

/* ---- Gauss-Newton matcher ----------------------------------------------
 * OptimizeScanMatchParam, scan_match/optimize_scan_matcher.h:33-58 */
typedef struct rsm_optimize_param {
  double cost_decrease_threshold;
  double cost_min_threshold;
  double max_update_distance;   /* metres */
  double max_update_angle;      /* radians */
  int32_t iterate_max_times;    /* >= 1 */
  int32_t reserved;
} rsm_optimize_param;

/* rsm_optimize: replaces BasedOptimizeScanMatch::ScanMatch(map, range_data, param, best_pose&) -> cost
 * (scan_match/optimize_scan_matcher.h:68-131): Gauss-Newton on the bilinearly interpolated lookup grid.
 * pose_world in/out; *cost = the reference's return value (1000 = kMaxCost for an uninitialised grid, an
 * empty scan or a NaN step, with pose_world untouched); iterations (nullable) = cost evaluations made.
 * Per-point terms are evaluated on the device and added in point order, so H, b and the cost carry
 * the reference's bits; the 3x3 solve follows Eigen 3.3's LDLT on the host. */
int rsm_optimize(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts,
                 const rsm_optimize_param* param, double pose_world[3], double* cost,
                 int32_t* iterations /* nullable */);
/* The same for a caller that owns the world<->map transform (the C++ adapter uses the live reference map's
 * GetMapCoordsPose / GetWorldCoordsPose): pose_map = the estimate in map cells / rad in, the optimised
 * estimate with its angle normalised (:125) out -- untouched when *cost comes back as kMaxCost. */
int rsm_optimize_map(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts,
                     const rsm_optimize_param* param, double pose_map[3], double* cost,
                     int32_t* iterations /* nullable */);
/* n independent problems, one launch per iteration over those still iterating.  Problem i uses
 * grids[i] and points pts_xy[pts_offset[i] .. pts_offset[i+1]). */
int rsm_optimize_batch(rsm_ctx* ctx, int n, const rsm_grid* const* grids, const double* pts_xy,
                       const int64_t* pts_offset, const rsm_optimize_param* param, double* poses_world,
                       double* costs, int32_t* iterations /* nullable */);
/* rsm_match_chain_opt: ScanMatchers::ScanMatch with use_optimize_scan_match on
 * (scan_match/scan_matchers.h:179-289): the optimiser on the coarse map with the scan in coarse-map
 * cells; if it fails (cost > optimize_failed_cost) or use_fine is 0, its result is dropped and the
 * coarse correlative pass runs from the seed; then fine and super-fine on the fine map.
 * responses (nullable) = {optimiser cost, coarse, fine, super responses}, 0 for a step not run. */
int rsm_match_chain_opt(rsm_ctx* ctx, const rsm_grid* coarse_grid, const double* pts_coarse, int n_coarse,
                        const rsm_grid* fine_grid, const double* pts_fine, int n_fine,
                        const rsm_pass_param params[3], const rsm_optimize_param* opt,
                        double optimize_failed_cost, int use_fine, double pose_world[3], double cov[9],
                        double* score, double responses[4] /* nullable */);

/* rsm_scan_match_interface_batch with the Gauss-Newton pre-step (use_optimize_scan_match, the reference's in-code default,
 * scan_match/scan_matchers.h:205-232): ScanMatchInterface resets BOTH back-end maps from the chain
 * (slam/slam_processor.cpp:282-285) -- the coarse one from coarse_store, whose scan i is scan i of fine_store in
 * coarse-map cells -- runs the optimiser on the coarse map, keeps its pose when cost <= optimize_failed_cost and
 * the fine passes follow, else runs the coarse correlative pass on the fine map from the seed; then fine and
 * super-fine.  scores[i] = the reference's average (optimize_failed_cost / (cost + optimize_failed_cost) stands in
 * for the dropped coarse response), map check as above.  responses (nullable): 4 per candidate -- optimiser cost,
 * coarse, fine, super-fine responses (0 for a step not run). */
int rsm_scan_match_interface_batch_opt(rsm_ctx* ctx, const rsm_scan_store* fine_store, const rsm_scan_store* coarse_store,
                                       int n, int grid_size, double resolution, double sigma, int coarse_grid_size,
                                       double coarse_resolution, double coarse_sigma, float default_prob,
                                       double occu_offset, const double* centres_world, const int64_t* chain_offset,
                                       const int32_t* chain_ids, const int32_t* match_ids,
                                       const rsm_pass_param params[3], const rsm_optimize_param* optimize,
                                       double optimize_failed_cost, int use_fine, double* poses_world, double* covs,
                                       double* scores, double* responses /* nullable, 4 per candidate */,
                                       const rsm_grid* pub_map /* nullable */, const rsm_scan_store* pub_store /* nullable */,
                                       const rsm_map_check_param* check /* nullable */);

/* ---- map rebuilds from the scan store ---------------------------------------------------------
 * SlamProcessor::CorrectPoseAndMap (slam/slam_processor.cpp:329-371) rebuilds all three maps from every
 * stored scan at its corrected pose after a loop closure: InitMapWithRangeVec = Reset + one
 * UpdateMapByRange per scan (map/occu_grid_map.h:222-255).  With the scans in a store these are one call
 * each, ids in the reference's order (the publishing map's list ends with map_min_passthrough extra copies
 * of id 0, :351-354; its float counters accumulate in list order).  Map extents stay the caller's policy. */
int rsm_grid_rebuild(rsm_ctx* ctx, rsm_grid* grid, const rsm_scan_store* store, int n, const int32_t* ids,
                     float default_prob, double sigma, double occu_offset, int use_blur);
int rsm_pubmap_rebuild(rsm_ctx* ctx, rsm_pubmap* pm, const rsm_scan_store* store, int n, const int32_t* ids,
                       float update_free_factor, float update_occu_factor);

/* ---- parity / multi-GPU building blocks --------------------------------------------------
 * rsm_pass_scores: penalised score of every candidate of one pass in candidate order
 * k = (angle_index*n_xy + x_index)*n_xy + y_index (the order of correlate_scan_matcher.h:552-584),
 * restricted to angle indices [angle_begin, angle_end) (pass 0, -1 for all).  scores_out holds
 * (angle_end-angle_begin)*n_xy^2 doubles.  center given in world coordinates. */
int rsm_pass_scores(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts,
                    const rsm_pass_param* param, const double pose_world[3], int angle_begin,
                    int angle_end, double* scores_out, int64_t capacity, int64_t* n_written);

/* Angle-sliced single window: one large search window cut along the angle index
 * over several GPUs (one context per GPU / rank).  The reference's winner is a tolerance-set
 * average plus two top-20 prefixes, not a single argmax, so the exchange is three calls around
 * two small all-gathers (e.g. torch.distributed.all_gather over NCCL):
 *
 *   rsm_match_partial   every rank scores angle indices [angle_begin, angle_end) and packs its
 *                       local maximum, its candidates within 1e-2 of it and its local top-21 into
 *                       `partial` (RSM_PARTIAL_BYTES); the slice's scores stay on the device.
 *        -- all-gather the partials --
 *   rsm_match_merge     every rank merges all partials (global maximum, averaging set, best
 *                       pose, global top-21) and writes its slice's scores of the same-(x,y)
 *                       columns into `columns` (RSM_COLUMNS_BYTES).
 *        -- all-gather the columns --
 *   rsm_match_finish    every rank finalises exactly like rsm_match: pose_world / cov in/out,
 *                       response, detail.  All ranks obtain identical results.
 *
 * If the consumed sets contain exact ties (whose order the reference's unstable sort decides),
 * rsm_match_finish returns RSM_NEED_EXACT on every rank (the decision only reads gathered data) and
 * writes nothing.  The ranks then exchange their slices' scores and every rank runs the reference's
 * own sort on the whole array:
 *
 *   rsm_match_slice_scores   copies this rank's slice ((angle_end - angle_begin) * n_xy^2 doubles,
 *                            candidate order) to host memory; capacity 0 only reports the size.
 *        -- all-gather the slices (rank order = angle order) --
 *   rsm_match_finish_exact   slices[r] / counts[r] = rank r's scores; together they must cover every
 *                            candidate of the window exactly once, in angle order.  Same outputs as
 *                            rsm_match_finish; detail->exact_sort_used = 1. */
#define RSM_PARTIAL_BYTES 65536
#define RSM_COLUMNS_BYTES 131072
int rsm_match_partial(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts,
                      const rsm_pass_param* param, const double pose_world[3], int angle_begin,
                      int angle_end, void* partial);
int rsm_match_merge(rsm_ctx* ctx, const void* const* partials, int n_partials, void* columns);
int rsm_match_finish(rsm_ctx* ctx, const void* const* partials, int n_partials,
                     const void* const* columns, double pose_world[3], double cov[9],
                     double* response, rsm_pass_detail* detail /* nullable */);
int rsm_match_slice_scores(rsm_ctx* ctx, double* scores_out, int64_t capacity, int64_t* n_slice);
int rsm_match_finish_exact(rsm_ctx* ctx, const double* const* slices, const int64_t* counts, int n_slices,
                           double pose_world[3], double cov[9], double* response,
                           rsm_pass_detail* detail /* nullable */);

/* The same exchange inside the library, over NCCL on the context's own stream (one process per GPU; libnccl.so.2 is
 * bound at run time, the copy the process already holds if there is one).  Rank 0 draws an id with rsm_comm_unique_id
 * and hands it to the others by any means (torch.distributed broadcast, MPI, a file); every rank then calls
 * rsm_comm_init on its context -- collectively, like ncclCommInitRank.  rsm_match_sliced is BasedCorrelationScanMatch::
 * ScanMatch (scan_match/correlate_scan_matcher.h:784-875) for ONE window cut along the angle index over the ranks
 * (contiguous slices, earlier ranks one angle more): every rank passes the same grid content, scan, parameters and
 * seed and receives the same, reference-identical result.  Per match: the slice is scored and selected, its partial is
 * packed on the device, ncclAllGather (16 KB per rank), host merge, the same-(x,y) columns go straight from the score
 * array into the exchange buffer, ncclAllGather (<= 9 columns x slice length), host finalisation -- two stream
 * synchronisations in all.  Exact ties in a consumed set: the slices' scores are all-gathered device to device and
 * every rank runs the reference's sort on the whole array (detail->exact_sort_used = 1).  world_size 1 needs no NCCL. */
#define RSM_COMM_ID_BYTES 128
int rsm_comm_unique_id(void* id_out /* RSM_COMM_ID_BYTES */);
int rsm_comm_init(rsm_ctx* ctx, int rank, int world_size, const void* id /* RSM_COMM_ID_BYTES; may be NULL for world_size 1 */);
int rsm_comm_destroy(rsm_ctx* ctx);
int rsm_match_sliced(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_pass_param* param,
                     double pose_world[3], double cov[9], double* response, rsm_pass_detail* detail /* nullable */);

/* ---- measurement support ------------------------------------------------------------------
 * Gather-bandwidth micro-benchmark used as the roofline denominator of the scoring kernel:
 * mode 0 = shared-memory row segments (32 consecutive 4-byte words per warp load, the access
 * shape of the scoring kernel), 1 = shared-memory random words, 2 = global-memory row segments
 * over an L1/L2-resident footprint, 3 = global-memory random words.  footprint_bytes: tile size
 * (modes 0/1, <= 200 KB) or buffer size (modes 2/3), rounded down to a power of two.
 * *gbps = bytes gathered / CUDA-event time. */
int rsm_microbench_gather(rsm_ctx* ctx, int mode, int64_t footprint_bytes, int iters, double* gbps);

/* The stream plan of the staged scoring kernel, for inspection and tests (host only; no device needed).
 * runs[3 i .. 3 i + 2] = beams visited, window width n_xy and search angles of job i; variant 0 / 1 / 2 = the
 * 81 x 96 paired-row, 64 x 64 and 96 x 96 tile mappings (negative: the one the library picks for the widest window).
 * The jobs' (angle, tile) items form one sequence of beams cut into at most max_ctas contiguous shares.
 * out receives 12 ints per share: item0, beam0, item1, beam1 (last item inclusive, its end beam exclusive), then for
 * the first and for the last item {ticket or -1, first partial slot, this share's part, parts} when other shares
 * visit the same item.  Returns the number of shares (or -needed when cap is too small); *n_items, *n_tickets,
 * *n_slots describe the launch. */
int rsm_stream_plan(int n_runs, const int* runs, int variant, int max_ctas, int* out, int cap, int64_t* n_items,
                    int* n_tickets, int* n_slots);

#ifdef __cplusplus
}
#endif
#endif /* RSM_H_ */

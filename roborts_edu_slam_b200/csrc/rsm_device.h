// rsm_device.h -- structures shared by the host orchestration (rsm_api.cu) and the kernels
// (rsm_kernels.cu).  Plain-old-data only; everything a kernel needs about one job is in one
// struct so that a batched launch is "an array of jobs + a CTA prefix array".
#ifndef RSM_DEVICE_H_
#define RSM_DEVICE_H_

#include <stdint.h>

namespace rsm {

// Lookup-grid cells are held either as exact fixed point (value * 2^25 in an int32; every
// probability the reference writes is a multiple of 2^-25) or as the raw
// float32 the reference stores (fallback when some value is not representable).
constexpr int kFixShift = 25;
constexpr double kFixScale = 1.0 / 33554432.0;  // 2^-25
constexpr int kFixOne = 1 << kFixShift;

// Device grids are padded to one of two row pitches (cells) so that the scoring kernel reaches the
// rows of a tile with immediate offsets (no address arithmetic per gather); wider grids keep
// pitch = size_x rounded up to 32 and use the run-time stride variant.  129 and 17 cache lines
// per row: consecutive rows fall into different L1 sets.
constexpr int kPitchSmall = 544;
constexpr int kPitchLarge = 4128;
inline int grid_pitch_for(int size_x) {
  return size_x <= kPitchSmall ? kPitchSmall : size_x <= kPitchLarge ? kPitchLarge : (size_x + 31) / 32 * 32;
}

constexpr int kTopK = 21;          // top-20 covariance prefix + 1 to detect a tie at the cut
constexpr int kChunk = 32;         // beams per shared-memory table chunk of the scoring kernel
constexpr int kMaxDevices = 64;   // per-device launch configuration caches
constexpr int kSelectBuf = 2048;
constexpr int kSelectThreads = 512;   // select_kernel CTA size
constexpr int kSelectSlice = 8192;    // candidates per select CTA (at least); a multiple of 8 * kSelectThreads
static_assert(kSelectSlice % (8 * kSelectThreads) == 0, "block_top_k caches whole iterations");   // per-CTA candidate buffer of the selection kernel
constexpr int kMaxCols = 9;        // same-(x,y) columns gathered for the angular covariance

enum ErrBits {
  kErrWindow = 1,       // a candidate's endpoint cell fell outside the grid (index was clamped)
  kErrPoolFull = 2,     // averaging-set pool overflowed its capacity
  kErrSelectFull = 4    // selection buffer overflowed (massive ties / flat score field)
};

struct Entry {  // one candidate handed back to the host
  double score;
  long long index;
};

// One scoring job = one pass of one match restricted to an angle slice.
struct ScoreJob {
  const void* grid;      // int32 fixed-point or float32 cells, row-major, `pitch` cells per row
  const double* pts;     // P x 2 scan points, cells, sensor frame
  const double* trig;    // n_ang x 3 : cos(angle_i), sin(angle_i), angle_i   (host libm values)
  double* score;         // out: ang_count * n_xy * n_xy penalised scores, candidate order
  unsigned long long* best_key;  // out: atomicMax of the order-preserving key of the scores
  int* err;              // out: OR of ErrBits
  int pitch, size_x, size_y;
  int P, step, V;        // points, beam stride, beams visited = ceil(P/step)
  int n_xy, ang_begin, ang_count;
  int tiles_x, tiles_y;  // candidate tiles per angle
  int use_penalty;
  int f_int, stepoff;    // affine variant: search step in cells (exact integer) and f_int * pitch
  int n_split;           // staged variant: CTAs (one cluster) sharing the beams of one (angle, tile);
                         // tiled / flat variants: CTAs sharing them through `acc` (small fixed-point launches)
  int pad0;
  unsigned long long* acc;  // n_split > 1 (tiled / flat): integer partial sums per candidate, zeroed before the launch
  int* tickets;             // n_split > 1 (tiled / flat): arrival counter per group of CTAs sharing candidates
  const void* tmap;      // staged variant: two CUtensorMaps over the grid (wide box, tall box), 128 B each
  double divisor;        // use_point_size after the reference's adjustment (:561-566)
  double sx, sy, f;      // search_space_start_x/y and space_step_factor (:546-548), in cells
  double cx, cy, ca;     // centre pose in map coordinates
  double m2;             // cell_len * cell_len
  double half_size;      // search_space_size / 2
  double gain;           // distance penalty gain (:759-761)
};

// Stream plan of the staged scoring kernel (rsm_score.cu, score_stream_kernel): the (job, angle, tile) items of a
// launch form one sequence of beams; every persistent CTA takes a contiguous range of it.  An item cut between CTAs
// is finished by whichever of them arrives last (ticket), from the integer partial sums the others left in a slot.
struct StreamCta {
  int item0, beam0;      // first item of the range and the first beam of it this CTA visits
  int item1, beam1;      // last item (inclusive) and the end of its beams (exclusive)
  int ticket0, slot0, part0, parts0;   // item0 shared with other CTAs: its ticket, its first partial slot, this CTA's part,
                                       // the number of parts; ticket0 < 0: item0 is this CTA's alone
  int ticket1, slot1, part1, parts1;   // the same for item1 when item1 != item0
};

// Averaging-set candidates of all jobs of a launch land in one pool (appended with an atomic
// counter, so the order is arbitrary; the host buckets by job and sorts).
struct PoolEntry {
  double score;
  int index;   // candidate index within the job's score array
  int job;
};

struct SelectJob {
  const double* score;
  long long n;                          // candidates in this job (angle slice x n_xy^2)
  const unsigned long long* best_key;   // written by the scoring kernel
  Entry* top_list; int* top_count;      // per-CTA top-kTopK lists: top_list[cta*kTopK + r], top_count[cta]
  Entry* final_top; int* final_count;   // merged by the job's last CTA: the job's top-kTopK, descending
  int* spec_cols;                       // [0] = number of speculative columns, [1..9] = ix*n_xy+iy
  double* spec_out;                     // spec_out[c * n_ang + ia]  (nullptr: no speculative gather)
  int* done;                            // arrival counter of the job's CTAs (zeroed before launch)
  int* err;
  int job_id;
  int n_cta;                            // CTAs working on this job
  int n_xy, n_ang;                      // window width and number of (local) angles
  long long slice;                      // candidates per CTA
};

struct GatherJob {
  const double* score;
  double* out;           // out[c * n_ang + ia]
  int n_xy, n_ang, n_cols;
  int cols[kMaxCols];    // ix * n_xy + iy
};

// Angle-sliced single window (include/rsm.h, rsm_match_partial ...): what one rank tells the others about its slice.
// A partial is this header followed by n_pool averaging-set candidates and n_top top-list candidates, 16 bytes
// each: {double score, int64 global candidate index}.
constexpr uint32_t kPartialMagic = 0x52534d50u;   // 'RSMP'
struct PartialHeader {
  uint32_t magic; int32_t a0, a1, n_ang, n_xy, n_pool, n_top, flags;   // flags: 1 = exact path needed, 2 = window left the grid
  unsigned long long best_key;
  double reserved[3];
};
static_assert(sizeof(PartialHeader) == 64, "PartialHeader layout");
struct ColumnsHeader { uint32_t magic; int32_t a0, a1, n_cols; int32_t cols[kMaxCols]; int32_t pad[3]; };
static_assert(sizeof(ColumnsHeader) == 64, "ColumnsHeader layout");
constexpr int kPartialDevBytes = 16384;           // partial as packed on the device for the in-library exchange

// device-side packing of a slice's partial (one job) for rsm_match_sliced
struct PackJob {
  const unsigned long long* best_key; const int* err; const int* pool_count; const PoolEntry* pool;
  const Entry* ftop; const int* fcnt;
  long long base;        // global index of the slice's first candidate
  int pool_cap, a0, a1, n_ang, n_xy, pad0;
  char* out;             // kPartialDevBytes
};

// One base scan to stamp into one grid.
struct RasterScan {
  void* grid;
  const double* pts;     // this scan's points (cells, sensor frame)
  int n_pts;
  int pitch, size_x, size_y;
  int start_x, start_y;  // beam start cell (sensor origin), points landing there are skipped
  double c, s, tx, ty;   // Translation(tx,ty) * Rotation(theta) with host libm cos/sin
};

struct FillJob {
  void* grid;
  long long n_cells;
  int value;             // raw 32-bit pattern (fixed-point int or float bits)
};

// One pose of a map-check batch (MapFeedbackResponsePenalty): the rays of one scan through the
// publishing map's occupancy mask.
struct PenaltyJob {
  const double* pts;     // scan points, cells of the publishing map, sensor frame
  int n_pts, step;       // points and check stride (:361-368)
  int sx0, sy0;          // beam start cell (host: pose transform of the sensor origin, :357-359)
  double c, s, tx, ty;   // cos / sin (host libm) and translation of the pose in map cells
  int* blocked;          // out: rays that met an occupied cell beyond the tolerance
  int pad0, pad1;
};

// One Gauss-Newton evaluation (BasedOptimizeScanMatch::UpdateCost) of one (grid, scan, pose estimate).
constexpr int kOptSums = 10;   // H00 H10 H20 H11 H21 H22 b0 b1 b2 cost
struct OptimizeJob {
  const void* grid;      // lookup grid cells (fixed-point int32 or float32)
  const double* pts;     // scan points, cells of this grid, sensor frame
  int n_pts;
  int size_x, size_y, pitch;
  int fixed, pad0;
  double c, s, tx, ty;   // cos / sin of the estimate's heading (host libm) and its translation in map cells
  double* out;           // kOptSums sums in point order + the number of points inside the map
};

// One UpdateMapByRange of a publishing map (OccuGridMap<CountCell>): ray-traced free space + occupied end cells.
struct PubScan {
  float* hit; float* pass; float* prob;   // CountCell::hit_count_, pass_count_, prob_value_ planes (dense, size_x per row)
  int* mark;                              // CountCell::update_index_
  const double* pts;                      // scan points, cells of this map, sensor frame
  int n_pts;
  int size_x, size_y;
  int start_x, start_y;                   // beam start cell
  int free_tag, occ_tag;                  // cur_mark_free_index / cur_mark_occu_index of this update
  int bx0, by0, bx1, by1;                 // cells the update can touch (inclusive): written by the mark pass, read by the apply pass
  float add_pass, add_hit;                // 1.0f + update_free_factor_, 1.0f + update_occu_factor_
  double c, s, tx, ty;                    // pose in map cells: cos / sin from the host libm
};

}  // namespace rsm

#endif

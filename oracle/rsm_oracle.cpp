// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's correlative scan matcher
// hot path, written from scratch on flat arrays.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load liboracle.so; nothing under
// roborts_edu_slam_b200/ may import, link or execute it.
//
// Pinning: the reference holds no golden vectors for this path.  The
// restatement is pinned against the reference's own headers compiled in place
// (oracle/_ref/libref.so, ref_driver.cpp) by tests/test_oracle.py, and against
// the fixtures that build wrote to tests/golden/ (used where /root/reference is absent).
//
// C++ rather than C for one reason: the reference's winner is defined by libstdc++'s unstable
// std::sort (correlate_scan_matcher.h:607-608); the same std::sort is used here on
// (score, index) pairs -- the permutation introsort produces depends only on the comparison
// outcomes, not on the element type.
//
// All citations are file:line under /root/reference/src.  Compiled with -ffp-contract=off:
// every a*b+c below is two rounded operations, as on the reference's baseline x86-64 build.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

const double kMaxVariance = 500.0;      // util/slam_util.h:57
const double kDoubleTolerance = 1e-06;  // util/slam_util.h:59
const int kMaxVarianceUse = 20;         // scan_match/correlate_scan_matcher.h:1033

// util/slam_util.h:70-73
inline bool DoubleEqual(double a, double b, double tol = kDoubleTolerance) {
  const double delta = a - b;
  return delta < 0.0 ? delta >= -std::fabs(tol) : delta <= std::fabs(tol);
}
// util/slam_util.h:75-77
inline double Round(double v) { return v >= 0.0 ? std::floor(v + 0.5) : std::ceil(v - 0.5); }

struct Param {
  double size, sres, aoff, ares, threshold;
  int use_point_size;
  bool use_center_penalty;
  int type;  // 0 coarse, 1 fine, 2 super, 3 fast
};

Param ReadParam(const double* p) {
  Param q;
  q.size = p[0]; q.sres = p[1]; q.aoff = p[2]; q.ares = p[3]; q.threshold = p[4];
  q.use_point_size = static_cast<int>(p[5]);
  q.use_center_penalty = p[6] != 0.0;
  q.type = static_cast<int>(p[7]);
  return q;
}

// map/grid_map_base.h:68-93 with Eigen's evaluation order (see oracle/standin/Eigen/Geometry):
// world_to_map = Scaling(s,s) * Translation(off): linear diag(s,s), translation (s*offx, s*offy);
// Affine * v = translation + (l00*x + l01*y).
struct MapTf {
  double s, tx, ty;          // world -> map
  double i00, i01, i10, i11, itx, ity;  // map -> world
};

MapTf MakeTf(double scale, double off_x, double off_y) {
  MapTf t;
  t.s = scale;
  t.tx = scale * off_x;
  t.ty = scale * off_y;
  const double l00 = scale, l01 = 0.0, l10 = 0.0, l11 = scale;
  const double det = l00 * l11 - l10 * l01;
  const double invdet = 1.0 / det;
  t.i00 = l11 * invdet;
  t.i10 = -l10 * invdet;
  t.i01 = -l01 * invdet;
  t.i11 = l00 * invdet;
  t.itx = (-t.i00) * t.tx + (-t.i01) * t.ty;
  t.ity = (-t.i10) * t.tx + (-t.i11) * t.ty;
  return t;
}
inline void WorldToMap(const MapTf& t, const double* w, double* m) {
  m[0] = t.tx + (t.s * w[0] + 0.0 * w[1]);
  m[1] = t.ty + (0.0 * w[0] + t.s * w[1]);
  m[2] = w[2];
}
inline void MapToWorld(const MapTf& t, const double* m, double* w) {
  w[0] = t.itx + (t.i00 * m[0] + t.i01 * m[1]);
  w[1] = t.ity + (t.i10 * m[0] + t.i11 * m[1]);
  w[2] = m[2];
}

struct Scored {
  double score;
  int index;
};

struct PassGeometry {
  int n_ang, n_xy, step, divisor;
  double start_x, start_y, factor, start_angle;
};

// correlate_scan_matcher.h:154,160-164 (angles), :538,546-548 (translations), :560-566 (beam stride)
PassGeometry Geometry(const Param& q, int P, double cell_len, const double* center) {
  PassGeometry g;
  g.n_ang = static_cast<int>(std::floor(q.aoff * 2 / q.ares) + 1);
  g.start_angle = center[2] - q.aoff;
  g.n_xy = static_cast<int>(Round(q.size / q.sres) + 1);
  g.start_x = center[0] - (q.size / cell_len) * 0.5;
  g.start_y = center[1] - (q.size / cell_len) * 0.5;
  g.factor = q.sres / cell_len;
  int use = q.use_point_size;
  if (P < 2 * use) { use = P; g.step = 1; } else { g.step = P / (use - 1); }
  g.divisor = use;
  return g;
}

// One brute-force pass: scores in candidate order k = (ia*n_xy + ix)*n_xy + iy, penalised.
// correlate_scan_matcher.h:552-603, 637-662, 718-745.  serach_angle_size_/2 (== aoff*2/2) is
// what the angle table receives (:526,536); aoff*2/2 == aoff exactly in binary floating point.
// Angles [a_begin, a_end) only (a_end < 0: all); score[] is indexed relative to a_begin.  Every angle is independent of
// the others, so slices computed on several threads concatenate to the bits of the full pass.
void ScorePass(const float* grid, int size_x, double cell_len, int P, const double* pts,
               const Param& q, const double* center, const PassGeometry& g, double* score_out,
               int32_t* dump_gx, int32_t* dump_gy, int a_begin = 0, int a_end = -1) {
  std::vector<double> lx(P), ly(P);
  const double half_range = (q.aoff * 2) / 2;
  const double start_angle = center[2] - half_range;
  if (a_end < 0) a_end = g.n_ang;
  double* score = score_out - static_cast<size_t>(a_begin) * g.n_xy * g.n_xy;
  for (int ia = a_begin; ia < a_end; ++ia) {
    const double angle = start_angle + ia * q.ares;
    const double c = std::cos(angle), s = std::sin(angle);
    for (int p = 0; p < P; ++p) {
      lx[p] = c * pts[2 * p] - s * pts[2 * p + 1];
      ly[p] = s * pts[2 * p] + c * pts[2 * p + 1];
    }
    for (int ix = 0; ix < g.n_xy; ++ix) {
      const double x = g.start_x + ix * g.factor;
      for (int iy = 0; iy < g.n_xy; ++iy) {
        const double y = g.start_y + iy * g.factor;
        const size_t k = (static_cast<size_t>(ia) * g.n_xy + ix) * g.n_xy + iy;
        double sum = 0.0;
        int v = 0;
        for (int p = 0; p < P; p += g.step, ++v) {
          const int gx = static_cast<int>(lx[p] + x + 0.5);
          const int gy = static_cast<int>(ly[p] + y + 0.5);
          if (dump_gx) {  // only for the index-parity test on tiny cases
            const size_t V = (P + g.step - 1) / g.step;
            dump_gx[k * V + v] = gx;
            dump_gy[k * V + v] = gy;
          }
          sum += static_cast<double>(grid[static_cast<size_t>(gy) * size_x + gx]);
        }
        score[k] = sum / g.divisor;
      }
    }
  }
  if (!q.use_center_penalty) return;
  const double gain = (q.type == 0) ? 0.4 : 0.2;  // :588-602, :760-761
  for (int ia = a_begin; ia < a_end; ++ia) {
    const double angle = start_angle + ia * q.ares;
    for (int ix = 0; ix < g.n_xy; ++ix) {
      const double x = g.start_x + ix * g.factor;
      for (int iy = 0; iy < g.n_xy; ++iy) {
        const double y = g.start_y + iy * g.factor;
        const size_t k = (static_cast<size_t>(ia) * g.n_xy + ix) * g.n_xy + iy;
        if (DoubleEqual(score[k], 0.0)) continue;
        const double dx = x - center[0], dy = y - center[1];
        double d2 = dx * dx + dy * dy;
        d2 *= (cell_len * cell_len);
        double dp = 1.0 - (gain * d2 / (q.size / 2));
        dp = std::max(dp, 0.5);
        const double da = angle - center[2];
        const double a2 = da * da;  // pow(x, 2)
        double ap = 1.0 - (0.25 * a2 / 0.349);
        ap = std::max(ap, 0.9);
        score[k] = score[k] * (dp * ap);
      }
    }
  }
}

struct Best {
  double x, y, angle, score;
};

struct PassOut {
  Best best;
  double response;
  long n_avg;
};

// Everything after the scores exist: sort (:607), FindBestCandidate (:670-710), covariance
// (:835-858, :887-956, :965-1019), response clamp and pose write-back (:861-869).
PassOut Finish(std::vector<Scored>& cand, const Param& q, const PassGeometry& g, double cell_len,
               const double* center, double* cov /*row-major 3x3, in/out*/) {
  const double start_angle = center[2] - (q.aoff * 2) / 2;
  auto pose_of = [&](int k, double* x, double* y, double* a) {
    const int iy = k % g.n_xy;
    const int ix = (k / g.n_xy) % g.n_xy;
    const int ia = k / (g.n_xy * g.n_xy);
    *x = g.start_x + ix * g.factor;
    *y = g.start_y + iy * g.factor;
    *a = start_angle + ia * q.ares;
  };
  std::sort(cand.begin(), cand.end(), [](const Scored& a, const Scored& b) { return a.score > b.score; });

  PassOut out;
  Best best;
  pose_of(cand[0].index, &best.x, &best.y, &best.angle);
  best.score = cand[0].score;
  double ax = 0.0, ay = 0.0, tx = 0.0, ty = 0.0, ssum = 0.0;
  long count = 0;
  for (const Scored& c : cand) {
    if (!DoubleEqual(c.score, best.score, 1e-2)) break;
    double x, y, a;
    pose_of(c.index, &x, &y, &a);
    ax += x * c.score;
    ay += y * c.score;
    tx += std::cos(a) * c.score;
    ty += std::sin(a) * c.score;
    ssum += c.score;
    ++count;
  }
  if (count > 1) {
    ax /= ssum; ay /= ssum; tx /= ssum; ty /= ssum;
    best.x = ax; best.y = ay; best.angle = std::atan2(ty, tx);
  }
  out.best = best;
  out.n_avg = count;

  const double max_ang_var = 4 * (q.ares * q.ares);  // :801
  auto C = [&](int r, int c) -> double& { return cov[3 * r + c]; };
  auto positional = [&]() {  // :887-956
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C(r, c) = (r == c) ? 1.0 : 0.0;
    if (best.score < kDoubleTolerance) {
      C(0, 0) = kMaxVariance; C(1, 1) = kMaxVariance; C(2, 2) = max_ang_var;
      return;
    }
    double vxx = 0.0, vxy = 0.0, vyy = 0.0, norm = 0.0;
    const double bound = std::min(best.score - 0.1, 0.5);
    int n = 0;
    for (const Scored& c : cand) {
      if (c.score > bound && n < kMaxVarianceUse) {
        double x, y, a;
        pose_of(c.index, &x, &y, &a);
        norm += c.score;
        vxx += ((x - best.x) * (x - best.x)) * c.score;
        vxy += ((x - best.x) * (y - best.y) * c.score);
        vyy += ((y - best.y) * (y - best.y)) * c.score;
        ++n;
      } else {
        break;
      }
    }
    if (norm > kDoubleTolerance) {
      double xx = vxx / norm, xy = vxy / norm, yy = vyy / norm;
      const double r = q.sres / cell_len;
      const double min_var = 0.1 * (r * r);
      xx = std::max(xx, min_var);
      yy = std::max(yy, min_var);
      const double m2 = cell_len * cell_len;
      C(0, 0) = (xx * m2) / best.score;
      C(0, 1) = (xy * m2) / best.score;
      C(1, 0) = (xy * m2) / best.score;
      C(1, 1) = (yy * m2) / best.score;
      C(2, 2) = max_ang_var;
    }
    if (DoubleEqual(C(0, 0), 0.0)) C(0, 0) = kMaxVariance;
    if (DoubleEqual(C(1, 1), 0.0)) C(1, 1) = kMaxVariance;
  };
  auto angular = [&]() {  // :965-1019
    if (best.score < kDoubleTolerance) { C(2, 2) = max_ang_var; return; }
    const double tol = q.sres / cell_len;
    const double bound = std::min(best.score - 0.1, 0.5);
    double norm = 0.0, acc = 0.0;
    int n = 0;
    for (const Scored& c : cand) {
      if (c.score >= bound && n < kMaxVarianceUse) {
        double x, y, a;
        pose_of(c.index, &x, &y, &a);
        if (DoubleEqual(x, best.x, tol) && DoubleEqual(y, best.y, tol)) {
          norm += c.score;
          acc += ((a - best.angle) * (a - best.angle)) * c.score;
          ++n;
        }
      }
    }
    double var = max_ang_var;
    if (norm > kDoubleTolerance) {
      var = acc / norm;  // the /4 branch at :1008-1010 is overwritten at :1012
    } else {
      var = 200 * max_ang_var;
    }
    C(2, 2) = var;
  };
  switch (q.type) {
    case 3:
    case 0: positional(); angular(); break;
    case 1: positional(); break;
    case 2: angular(); break;
    default: break;
  }
  out.response = best.score > 1.0 ? 1.0 : best.score;
  return out;
}

double OnePass(const float* grid, int size_x, double scale, double off_x, double off_y, int P,
               const double* pts, const Param& q, double* pose_world, double* cov,
               double* best_map_out, long* n_avg_out) {
  if (P == 0 || grid == nullptr) return 0.0;  // :792-795
  const MapTf tf = MakeTf(scale, off_x, off_y);
  const double cell_len = 1 / scale;  // grid_map_base.h:307-309
  double center[3];
  WorldToMap(tf, pose_world, center);
  const PassGeometry g = Geometry(q, P, cell_len, center);
  const size_t n = static_cast<size_t>(g.n_ang) * g.n_xy * g.n_xy;
  std::vector<double> score(n);
  ScorePass(grid, size_x, cell_len, P, pts, q, center, g, score.data(), nullptr, nullptr);
  std::vector<Scored> cand(n);
  for (size_t k = 0; k < n; ++k) { cand[k].score = score[k]; cand[k].index = static_cast<int>(k); }
  PassOut o = Finish(cand, q, g, cell_len, center, cov);
  if (best_map_out) { best_map_out[0] = o.best.x; best_map_out[1] = o.best.y; best_map_out[2] = o.best.angle; best_map_out[3] = o.best.score; }
  if (n_avg_out) *n_avg_out = o.n_avg;
  if (o.response > q.threshold) {
    const double b[3] = {o.best.x, o.best.y, o.best.angle};
    MapToWorld(tf, b, pose_world);
  }
  return o.response;
}

}  // namespace

extern "C" {

// occu_grid_map.h:40-59, 83-105
int orc_blur_kernel(double sigma, double resolution, double* k, int cap) {
  const double lo = 0.5 * resolution, hi = 10 * resolution;
  if (!(sigma > lo && sigma < hi && resolution > 0)) return -1;
  const int half = static_cast<int>((sigma / resolution) * std::sqrt(std::log(2)));
  const int n = 2 * half + 1;
  if (k && cap >= n * n) {
    for (int i = -half; i <= half; ++i)
      for (int j = -half; j <= half; ++j) {
        const double d = std::hypot(i * resolution, j * resolution);
        const double q = d / sigma;
        k[(i + half) + n * (j + half)] = std::exp(-0.5 * (q * q));
      }
  }
  return half;
}

// Lookup-grid build in the back-end configuration (just_update_occu, no auto-resize):
// Reset to default, then per base scan transform / truncate / bounds-skip / stamp.
// occu_grid_map.h:222-329, 474-497, 531-576; grid_map_cell.h:361-365; grid_map_base.h:339-346.
// use_blur == 0 (or blur parameters the reference rejects): the SET_CELL_OCCUPIED path, SetCellOccu (:499-516).
// reset != 0: InitMapWithRangeVec (Reset to default_prob first); reset == 0: UpdateMapByRange on the map as it is
// (the front-end scan-match maps, just_update_occu, slam_processor.cpp:529-571)
int orc_grid_stamp(float* grid, int size_x, int size_y, float default_prob, double sigma,
                   double resolution, double occu_offset, double off_x, double off_y, int n_scans,
                   const int* n_pts, const double* pts, const double* poses, int use_blur, int reset) {
  std::vector<double> kernel(21 * 21);
  int half = orc_blur_kernel(sigma, resolution, kernel.data(), 21 * 21);
  // GaussianBlur rejected the parameters: half_kernel_size_ = 0, and UpdateMapByRange drops use_blur
  // (map/occu_grid_map.h:47-59, 265-268)
  if (half < 0) { half = 0; use_blur = 0; }
  const int ks = 2 * half + 1;
  const size_t ncell = static_cast<size_t>(size_x) * size_y;
  if (reset) for (size_t i = 0; i < ncell; ++i) grid[i] = default_prob;
  const double scale = 1.0 / resolution;
  const MapTf tf = MakeTf(scale, off_x, off_y);
  auto set_prob = [&](int x, int y, float prob) {
    float& c = grid[static_cast<size_t>(y) * size_x + x];
    if (c < prob && prob <= 1.0f) c = prob;
  };
  // SET_CELL_OCCUPIED (use_blur false): SetCellOccu (:499-516) acts once per cell and update --
  // cell.update_index_ < cur_mark_occu_index -- and adds ProbabilityCellFunctions' update_occu_factor_
  // (0.5f, map/grid_map_cell.h:333-336, 338-342), clamped to 1.  No ray is traced on these maps
  // (just_update_occu), so update_index_ never equals cur_mark_free_index; the indices of earlier
  // updates are always smaller (cur_update_index only grows), hence a per-call plane is equivalent.
  std::vector<int> update_index;
  if (!use_blur) update_index.assign(ncell, -1);
  int cur_update_index = 0;
  const float update_occu_factor = 0.5f;
  size_t off = 0;
  const double tol = half + 1;
  for (int s = 0; s < n_scans; ++s) {
    const int cur_mark_occu_index = cur_update_index + 2;   // :272-273
    double pm[3];
    WorldToMap(tf, poses + 3 * s, pm);
    const double c = std::cos(pm[2]), sn = std::sin(pm[2]);
    // beam start = transform * origin(0,0) -> translation + (c*0 + (-s)*0)
    const int sx0 = static_cast<int>((pm[0] + (c * 0.0 + (-sn) * 0.0)) + 0.5);
    const int sy0 = static_cast<int>((pm[1] + (sn * 0.0 + c * 0.0)) + 0.5);
    for (int i = 0; i < n_pts[s]; ++i) {
      const double px = pts[2 * (off + i)], py = pts[2 * (off + i) + 1];
      const double mx = pm[0] + (c * px + (-sn) * py);
      const double my = pm[1] + (sn * px + c * py);
      const int ex = static_cast<int>(mx + 0.5), ey = static_cast<int>(my + 0.5);
      if (ex == sx0 && ey == sy0) continue;
      if (!(ex > tol && ex < size_x - tol && ey > tol && ey < size_y - tol)) continue;
      if (!use_blur) {
        const size_t at = static_cast<size_t>(ey) * size_x + ex;
        if (update_index[at] < cur_mark_occu_index) {
          float v = grid[at] + update_occu_factor;
          if (v > 1.0f) v = 1.0f;
          grid[at] = v;
          update_index[at] = cur_mark_occu_index;
        }
        continue;
      }
      set_prob(ex, ey, 1.0f);
      for (int j = -half; j <= half; ++j)
        for (int ii = -half; ii <= half; ++ii)
          set_prob(ex + ii, ey + j, static_cast<float>(kernel[(ii + half) + ks * (j + half)] * occu_offset));
    }
    off += n_pts[s];
    cur_update_index += 3;   // :326
  }
  return 0;
}

int orc_grid_build(float* grid, int size_x, int size_y, float default_prob, double sigma,
                   double resolution, double occu_offset, double off_x, double off_y, int n_scans,
                   const int* n_pts, const double* pts, const double* poses, int use_blur) {
  return orc_grid_stamp(grid, size_x, size_y, default_prob, sigma, resolution, occu_offset, off_x, off_y, n_scans, n_pts, pts,
                        poses, use_blur, 1);
}

// GridMapBase::ExtendSize's cell copy (map/grid_map_base.h:222-238): a new array whose cell 0 holds `first` and every
// other cell `fill` (new CellType[n]{default}), the old rows copied in at (pre_x, pre_y).
void orc_grid_extend(const float* old_grid, int old_sx, int old_sy, float* new_grid, int new_sx, int new_sy, int pre_x,
                     int pre_y, float fill, float first) {
  const size_t n = static_cast<size_t>(new_sx) * new_sy;
  for (size_t i = 0; i < n; ++i) new_grid[i] = fill;
  new_grid[0] = first;
  for (int r = 0; r < old_sy; ++r)
    std::memcpy(new_grid + static_cast<size_t>(pre_y + r) * new_sx + pre_x, old_grid + static_cast<size_t>(r) * old_sx,
                sizeof(float) * old_sx);
}

// OccuGridMap<CountCell>::UpdateMapByRange without blur (the publishing map; map/occu_grid_map.h:258-329, 125-187,
// 474-530; map/grid_map_cell.h:92-108), in the reference's own order: per beam the Bresenham walk marks cells free
// (once per update), then the end cell is un-freed and marked occupied (once per update).
// hit / pass / value / index: the CountCell planes; cur_update_index in/out (advanced by 3).
void orc_pubmap_update(float* hit, float* pass, float* value, int* index, int size_x, int size_y, double scale, double off_x,
                       double off_y, int n, const double* pts, const double* pose_world, float free_factor, float occu_factor,
                       int* cur_update_index) {
  const MapTf tf = MakeTf(scale, off_x, off_y);
  double pm[3];
  WorldToMap(tf, pose_world, pm);
  const double c = std::cos(pm[2]), sn = std::sin(pm[2]);
  const int mark_free = *cur_update_index + 1, mark_occu = *cur_update_index + 2;   // :272-273
  const int sx0 = static_cast<int>((pm[0] + (c * 0.0 + (-sn) * 0.0)) + 0.5);
  const int sy0 = static_cast<int>((pm[1] + (sn * 0.0 + c * 0.0)) + 0.5);
  auto in_map = [&](int x, int y) { return x > 1 && x < size_x - 1 && y > 1 && y < size_y - 1; };   // PointInMap(x, y, 0 + 1), :476
  auto set_free = [&](int x, int y) {                                                // :499-509, grid_map_cell.h:100-103
    if (!in_map(x, y)) return;
    const size_t k = static_cast<size_t>(y) * size_x + x;
    if (index[k] < mark_free) {
      pass[k] += (1.0f + free_factor);
      value[k] = hit[k] / pass[k];
      index[k] = mark_free;
    }
  };
  auto set_occu = [&](int x, int y) {                                                // :511-528, grid_map_cell.h:92-108
    if (!in_map(x, y)) return;
    const size_t k = static_cast<size_t>(y) * size_x + x;
    if (index[k] < mark_occu) {
      if (index[k] == mark_free) { pass[k] -= (1.0f + free_factor); value[k] = hit[k] / pass[k]; }
      hit[k] += (1.0f + occu_factor);
      pass[k] += (1.0f + free_factor);
      value[k] = hit[k] / pass[k];
      if (value[k] > 1.0f) value[k] = 1.0f;
      index[k] = mark_occu;
    }
  };
  for (int i = 0; i < n; ++i) {
    const double px = pts[2 * i], py = pts[2 * i + 1];
    const int ex = static_cast<int>((pm[0] + (c * px + (-sn) * py)) + 0.5);
    const int ey = static_cast<int>((pm[1] + (sn * px + c * py)) + 0.5);
    if (ex == sx0 && ey == sy0) continue;                                            // :312
    int x0 = sx0, y0 = sy0, x1 = ex, y1 = ey;                                        // ErgodLineBresenhami, :125-187
    const bool steep = std::abs(y1 - y0) > std::abs(x1 - x0);
    if (steep) { std::swap(x0, y0); std::swap(x1, y1); }
    if (x0 > x1) { std::swap(x0, x1); std::swap(y0, y1); }
    const int delta_x = x1 - x0, delta_y = std::abs(y1 - y0), y_step = y0 < y1 ? 1 : -1;
    int error = 0, y = y0;
    for (int x = x0; x <= x1; ++x) {
      const int qx = steep ? y : x, qy = steep ? x : y;
      error += delta_y;
      if (2 * error >= delta_x) { y += y_step; error -= delta_x; }
      set_free(qx, qy);
    }
    set_occu(ex, ey);
  }
  *cur_update_index += 3;                                                            // :326
}

void orc_world_to_map(double scale, double off_x, double off_y, const double* w, double* out) {
  WorldToMap(MakeTf(scale, off_x, off_y), w, out);
}
void orc_map_to_world(double scale, double off_x, double off_y, const double* m, double* out) {
  MapToWorld(MakeTf(scale, off_x, off_y), m, out);
}

// geometry of a pass: out = {n_ang, n_xy, step, divisor, visited}
void orc_pass_geometry(const double* param, int P, double cell_len, const double* center_map,
                       long* out, double* dout /*start_x,start_y,factor,start_angle*/) {
  const Param q = ReadParam(param);
  const PassGeometry g = Geometry(q, P, cell_len, center_map);
  out[0] = g.n_ang; out[1] = g.n_xy; out[2] = g.step; out[3] = g.divisor;
  out[4] = (P + g.step - 1) / g.step;
  if (dout) { dout[0] = g.start_x; dout[1] = g.start_y; dout[2] = g.factor; dout[3] = g.start_angle; }
}

// Penalised scores of every candidate in candidate order; optional (gx,gy) dump
// [n_cand][visited] for the cell-index parity test.
int orc_scores(const float* grid, int size_x, double cell_len, int P, const double* pts,
               const double* param, const double* center_map, double* score, int32_t* gx,
               int32_t* gy) {
  const Param q = ReadParam(param);
  const PassGeometry g = Geometry(q, P, cell_len, center_map);
  ScorePass(grid, size_x, cell_len, P, pts, q, center_map, g, score, gx, gy);
  return 0;
}

// BasedCorrelationScanMatch::ScanMatch restated (correlate_scan_matcher.h:784-875).
// One angle slice of a pass (the full-size config-5 golden is scored on all host threads, one slice each).
int orc_scores_slice(const float* grid, int size_x, double cell_len, int P, const double* pts, const double* param,
                     const double* center_map, int a_begin, int a_end, double* score_out) {
  const Param q = ReadParam(param);
  const PassGeometry g = Geometry(q, P, cell_len, center_map);
  if (a_begin < 0 || a_end > g.n_ang || a_end < a_begin) return 1;
  ScorePass(grid, size_x, cell_len, P, pts, q, center_map, g, score_out, nullptr, nullptr, a_begin, a_end);
  return 0;
}

// Everything after the scores exist (sort, FindBestCandidate, covariance, response, pose write-back) on a score
// array computed elsewhere (orc_scores / orc_scores_slice): what orc_match does after ScorePass.
double orc_finish_scores(const double* score, long n_scores, int P, double scale, double off_x, double off_y,
                         const double* param, double* pose_world, double* cov, double* best_map_out, long* n_avg_out) {
  const Param q = ReadParam(param);
  const MapTf tf = MakeTf(scale, off_x, off_y);
  const double cell_len = 1 / scale;
  double center[3];
  WorldToMap(tf, pose_world, center);
  const PassGeometry g = Geometry(q, P, cell_len, center);
  const size_t n = static_cast<size_t>(g.n_ang) * g.n_xy * g.n_xy;
  if (static_cast<size_t>(n_scores) != n) return -1.0;
  std::vector<Scored> cand(n);
  for (size_t k = 0; k < n; ++k) { cand[k].score = score[k]; cand[k].index = static_cast<int>(k); }
  PassOut o = Finish(cand, q, g, cell_len, center, cov);
  if (best_map_out) { best_map_out[0] = o.best.x; best_map_out[1] = o.best.y; best_map_out[2] = o.best.angle; best_map_out[3] = o.best.score; }
  if (n_avg_out) *n_avg_out = o.n_avg;
  if (o.response > q.threshold) {
    const double b[3] = {o.best.x, o.best.y, o.best.angle};
    MapToWorld(tf, b, pose_world);
  }
  return o.response;
}

double orc_match(const float* grid, int size_x, int size_y, double scale, double off_x,
                 double off_y, int P, const double* pts, const double* param, double* pose_world,
                 double* cov, double* best_map_out, long* n_avg_out, double* seconds) {
  (void)size_y;
  auto t0 = std::chrono::steady_clock::now();
  double r = OnePass(grid, size_x, scale, off_x, off_y, P, pts, ReadParam(param), pose_world, cov,
                     best_map_out, n_avg_out);
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return r;
}

// ScanMatchers::ScanMatch with the optimiser off (scan_matchers.h:224-263, 281).
double orc_match_chain(const float* grid, int size_x, int size_y, double scale, double off_x,
                       double off_y, int P, const double* pts, const double* params, int use_fine,
                       double* pose_world, double* cov, double* resp_out, double* seconds) {
  (void)size_y;
  auto t0 = std::chrono::steady_clock::now();
  double score = 0.0, r[3] = {0.0, 0.0, 0.0};
  int times = 0;
  r[0] = OnePass(grid, size_x, scale, off_x, off_y, P, pts, ReadParam(params), pose_world, cov, nullptr, nullptr);
  score += r[0]; ++times;
  if (use_fine) {
    r[1] = OnePass(grid, size_x, scale, off_x, off_y, P, pts, ReadParam(params + 8), pose_world, cov, nullptr, nullptr);
    score += r[1]; ++times;
    r[2] = OnePass(grid, size_x, scale, off_x, off_y, P, pts, ReadParam(params + 16), pose_world, cov, nullptr, nullptr);
    score += r[2]; ++times;
  }
  score /= times;
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (resp_out) { resp_out[0] = r[0]; resp_out[1] = r[1]; resp_out[2] = r[2]; }
  return score;
}

}  // extern "C"

namespace {

// ---- BasedOptimizeScanMatch (scan_match/optimize_scan_matcher.h:60-237) ---------------------------
// Gauss-Newton on the bilinearly interpolated lookup grid.  Pinning: every line below is checked
// bit for bit against the reference's own header compiled in place (oracle/_ref), EXCEPT the 3x3
// solve: the reference calls Eigen's H.ldlt().solve(b); real Eigen is not in this image, so both
// this file and the stand-in header restate Eigen 3.3's LDLT (Cholesky/LDLT.h) -- that one step
// is "parity unpinned" (it is association-free for 3x3 apart from d22 - (l20*t0 + l21*t1) and the
// two-term sums of the triangular solves).

// Eigen 3.3 LDLT<Matrix3d>::compute (ldlt_inplace<Lower>::unblocked) + _solve_impl, column-major 3x3.
void Ldlt3Solve(const double* H /* column-major */, const double* b, double* x) {
  double m[9];
  for (int i = 0; i < 9; ++i) m[i] = H[i];
  auto M = [&](int r, int c) -> double& { return m[c * 3 + r]; };
  int tr[3];
  double temp[3];
  const int n = 3;
  for (int k = 0; k < n; ++k) {
    int big = k;
    double best = std::fabs(M(k, k));
    for (int i = k + 1; i < n; ++i) if (std::fabs(M(i, i)) > best) { best = std::fabs(M(i, i)); big = i; }
    tr[k] = big;
    if (k != big) {
      const int sz = n - big - 1;
      for (int j = 0; j < k; ++j) std::swap(M(k, j), M(big, j));
      for (int i = 0; i < sz; ++i) std::swap(M(big + 1 + i, k), M(big + 1 + i, big));
      std::swap(M(k, k), M(big, big));
      for (int i = k + 1; i < big; ++i) { const double t = M(i, k); M(i, k) = M(big, i); M(big, i) = t; }
    }
    const int rs = n - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; ++j) temp[j] = M(j, j) * M(k, j);
      double dot = M(k, 0) * temp[0];
      for (int j = 1; j < k; ++j) dot = dot + M(k, j) * temp[j];
      M(k, k) -= dot;
      for (int i = 0; i < rs; ++i)
        for (int j = 0; j < k; ++j) M(k + 1 + i, k) = M(k + 1 + i, k) + M(k + 1 + i, j) * (-1.0 * temp[j]);
    }
    const double akk = M(k, k);
    const bool pivot_ok = std::fabs(akk) > 0.0;
    if (k == 0 && !pivot_ok) { for (int j = 0; j < n; ++j) tr[j] = j; break; }
    if (rs > 0 && pivot_ok) for (int i = 0; i < rs; ++i) M(k + 1 + i, k) /= akk;
  }
  for (int i = 0; i < n; ++i) x[i] = b[i];
  for (int k = 0; k < n; ++k) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
  for (int i = 1; i < n; ++i) {
    double sum = M(i, 0) * x[0];
    for (int j = 1; j < i; ++j) sum = sum + M(i, j) * x[j];
    x[i] -= sum;
  }
  for (int i = 0; i < n; ++i) { if (std::fabs(M(i, i)) > DBL_MIN) x[i] /= M(i, i); else x[i] = 0.0; }
  for (int i = n - 2; i >= 0; --i) {
    double sum = M(i + 1, i) * x[i + 1];
    for (int j = i + 2; j < n; ++j) sum = sum + M(j, i) * x[j];
    x[i] -= sum;
  }
  for (int k = n - 1; k >= 0; --k) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
}

// util/slam_util.h:79-87
inline double MaxAbxLimit(double value, double limit) {
  if (value > std::fabs(limit)) value = std::fabs(limit);
  else if (value < -std::fabs(limit)) value = -std::fabs(limit);
  return value;
}
// util/slam_util.h:103-111
inline double NormalizeAngle(double angle) {
  double a = std::fmod(std::fmod(angle, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI);
  if (a > M_PI) a -= 2.0 * M_PI;
  return a;
}

// UpdateCost (optimize_scan_matcher.h:154-220): H (column-major, symmetric by construction), b, cost
// for one pose estimate in map cells.  A cell index past the end of the cell array (x1 == size_x in the
// last row: the reference reads out of bounds there) reads 0.
void OptimizeCost(const float* grid, int size_x, int size_y, int P, const double* pts, const double* est,
                  double* H, double* b, double* cost_out, long* oob_reads = nullptr) {
  const double c = std::cos(est[2]), sn = std::sin(est[2]);        // :95-97, :197-198
  for (int i = 0; i < 9; ++i) H[i] = 0.0;
  b[0] = b[1] = b[2] = 0.0;
  double cost = 0.0;
  int valid_point = 1;                                             // :160
  const long n_cells = static_cast<long>(size_x) * size_y;
  auto cell = [&](int x, int y) -> double {
    const long idx = static_cast<long>(y) * size_x + x;            // grid_map_base.h:352-354
    if (idx >= 0 && idx < n_cells) return static_cast<double>(grid[idx]);
    if (oob_reads) ++*oob_reads;                                   // undefined behaviour in the reference
    return 0.0;
  };
  for (int p = 0; p < P; ++p) {
    const double lx = pts[2 * p], ly = pts[2 * p + 1];
    const double x = (c * lx + (-sn) * ly) + est[0];               // rotation * local_point + translation (:167)
    const double y = (sn * lx + c * ly) + est[1];
    if (!(x > 0 && x < size_x && y > 0 && y < size_y)) continue;   // PointInMap, grid_map_base.h:330-337
    const double x0 = std::floor(x), y0 = std::floor(y), x1 = std::ceil(x), y1 = std::ceil(y);
    const double m00 = cell(static_cast<int>(x0), static_cast<int>(y0));
    const double m01 = cell(static_cast<int>(x0), static_cast<int>(y1));
    const double m10 = cell(static_cast<int>(x1), static_cast<int>(y0));
    const double m11 = cell(static_cast<int>(x1), static_cast<int>(y1));
    double r = ((y - y0) * (m11 * (x - x0) + m01 * (x1 - x)) + (y1 - y) * (m10 * (x - x0) + m00 * (x1 - x)));   // :186-187
    r = (r >= 0) ? ((r <= 1) ? r : 1) : 0;                         // :190-191
    const double error = 1 - r;
    cost += (error * error);
    const double ds0 = (-sn * lx - c * ly);                        // :197
    const double ds1 = (c * lx - sn * ly);                         // :198
    const double dm0 = ((y - y0) * (m11 - m01) + (y1 - y) * (m10 - m00));   // :200
    const double dm1 = ((x - x0) * (m11 - m10) + (x1 - x) * (m01 - m00));   // :201
    double J[3];                                                   // J = -de_m * de_s (:204), lazy 1x2 * 2x3 product
    J[0] = (-dm0) * 1.0 + (-dm1) * 0.0;
    J[1] = (-dm0) * 0.0 + (-dm1) * 1.0;
    J[2] = (-dm0) * ds0 + (-dm1) * ds1;
    for (int cc = 0; cc < 3; ++cc)
      for (int rr = 0; rr < 3; ++rr) H[cc * 3 + rr] = H[cc * 3 + rr] + J[rr] * J[cc];   // :206
    for (int rr = 0; rr < 3; ++rr) b[rr] = b[rr] + (-J[rr]) * error;                     // :207
    ++valid_point;
  }
  cost *= (1000.0 / valid_point);                                  // :218, kCostPointSize = 1000
  *cost_out = cost;
}

// BasedOptimizeScanMatch::ScanMatch (:68-131).  op = {iterate_max_times, cost_decrease_threshold,
// cost_min_threshold, max_update_distance, max_update_angle}.  iterate_max_times < 1 would return the
// previous call's cost in the reference (a stale member); here it returns NaN.
double Optimize(const float* grid, int size_x, int size_y, double scale, double off_x, double off_y, int P,
                const double* pts, const double* op, double* pose_world, int* iters_out, long* oob_reads = nullptr) {
  const double kMaxCost = 1.0 * 1000;
  if (iters_out) *iters_out = 0;
  if (P == 0) return kMaxCost;                                     // :73-76 (map initialised by the caller)
  const MapTf tf = MakeTf(scale, off_x, off_y);
  double est[3];
  WorldToMap(tf, pose_world, est);
  const double map_resolution = 1 / scale;                         // GetCellLength(), grid_map_base.h:307-309
  const int iterate_max_times = static_cast<int>(op[0]);
  double cost = std::nan(""), last_cost = 0.0;
  int iter = 0;
  for (; iter < iterate_max_times; ++iter) {
    last_cost = cost;
    double H[9], b[3], det[3];
    OptimizeCost(grid, size_x, size_y, P, pts, est, H, b, &cost, oob_reads);
    Ldlt3Solve(H, b, det);                                         // :135-141
    if (std::isnan(det[0]) || std::isnan(det[1]) || std::isnan(det[2])) { if (iters_out) *iters_out = iter; return kMaxCost; }   // :103-106
    if (iter > 0 && (last_cost - cost < op[1] || cost < op[2])) break;   // :112-118
    est[0] += MaxAbxLimit(det[0], op[3] / map_resolution);         // :143-152
    est[1] += MaxAbxLimit(det[1], op[3] / map_resolution);
    est[2] += MaxAbxLimit(det[2], op[4]);
  }
  if (iters_out) *iters_out = iter;
  est[2] = NormalizeAngle(est[2]);                                 // :125
  MapToWorld(tf, est, pose_world);                                 // :127
  return cost;
}

}  // namespace

extern "C" {

// oob_out (optional): reads past the end of the cell array (a point in the last row's strip size_y-1 < y < size_y:
// undefined behaviour in the reference, 0.0 here and in the CUDA path)
double orc_optimize(const float* grid, int size_x, int size_y, double scale, double off_x, double off_y, int P,
                    const double* pts, const double* op, double* pose_world, int* iters_out, long* oob_out) {
  long oob = 0;
  const double c = Optimize(grid, size_x, size_y, scale, off_x, off_y, P, pts, op, pose_world, iters_out, &oob);
  if (oob_out) *oob_out = oob;
  return c;
}

// H (row-major 3x3 = column-major, it is symmetric), b, cost of one UpdateCost evaluation at a map pose
void orc_optimize_cost(const float* grid, int size_x, int size_y, int P, const double* pts, const double* pose_map,
                       double* H, double* b, double* cost) {
  OptimizeCost(grid, size_x, size_y, P, pts, pose_map, H, b, cost);
}

void orc_ldlt3_solve(const double* H_colmajor, const double* b, double* x) { Ldlt3Solve(H_colmajor, b, x); }

// ScanMatchers::ScanMatch with use_optimize_scan_match = true (scan_matchers.h:179-289): optimiser on the
// coarse map / coarse-resolution scan, correlative passes on the fine map.
// resp_out (optional) = {optimiser cost, coarse, fine, super responses}.
double orc_match_chain_opt(const float* grid_c, int size_xc, int size_yc, double scale_c, double off_xc, double off_yc,
                           int Pc, const double* pts_c, const float* grid_f, int size_xf, int size_yf, double scale_f,
                           double off_xf, double off_yf, int Pf, const double* pts_f, const double* params,
                           const double* op, double optimize_failed_cost, int use_fine, double* pose_world,
                           double* cov, double* resp_out) {
  (void)size_yf;
  double best_pose[3] = {pose_world[0], pose_world[1], pose_world[2]};
  double process_pose[3] = {best_pose[0], best_pose[1], best_pose[2]};
  double scan_match_score = 0.0;
  int scan_match_times = 0;
  const double optimize_cost = Optimize(grid_c, size_xc, size_yc, scale_c, off_xc, off_yc, Pc, pts_c, op, process_pose, nullptr);   // :207
  scan_match_score = optimize_failed_cost / (optimize_cost + optimize_failed_cost);          // :211
  scan_match_times++;
  double r[3] = {0.0, 0.0, 0.0};
  if (!use_fine || optimize_cost > optimize_failed_cost) {                                   // :224-226
    scan_match_score = 0.0;
    scan_match_times--;
    for (int k = 0; k < 3; ++k) process_pose[k] = best_pose[k];                              // :232
    r[0] = OnePass(grid_f, size_xf, scale_f, off_xf, off_yf, Pf, pts_f, ReadParam(params), process_pose, cov, nullptr, nullptr);
    scan_match_score += r[0];
    scan_match_times++;
  }
  if (use_fine) {                                                                            // :247-263
    r[1] = OnePass(grid_f, size_xf, scale_f, off_xf, off_yf, Pf, pts_f, ReadParam(params + 8), process_pose, cov, nullptr, nullptr);
    scan_match_score += r[1]; scan_match_times++;
    r[2] = OnePass(grid_f, size_xf, scale_f, off_xf, off_yf, Pf, pts_f, ReadParam(params + 16), process_pose, cov, nullptr, nullptr);
    scan_match_score += r[2]; scan_match_times++;
  }
  scan_match_score /= scan_match_times;                                                      // :281
  if (resp_out) { resp_out[0] = optimize_cost; resp_out[1] = r[0]; resp_out[2] = r[1]; resp_out[3] = r[2]; }
  for (int k = 0; k < 3; ++k) pose_world[k] = process_pose[k];
  return scan_match_score;
}


// OccuGridMap::MapFeedbackResponsePenalty (map/occu_grid_map.h:331-392): rays from the sensor to
// every check point, walked with LineVisitor::ErgodLineBresenhami (:125-187); the callback
// (CheckOccuLineVisitorCallback, :447-471) adds 1 to the ray's result, while that result is still
// below 1, for a cell that counts as occupied and lies farther than bound_tolerance from the ray's
// end point.  `occupied[y*size_x+x]` is that per-cell test (GetGridStates == Occupied without
// blur, value > cell_occu_prob_offset with blur), evaluated by the caller.
// pts: raw scan points in cells of this map, sensor frame; origin: RangeDataContainer::sensor_origin().
double orc_map_feedback_penalty(const unsigned char* occupied, int size_x, int size_y, double scale, double off_x,
                                double off_y, int n_pts, const double* pts, const double* pose_world,
                                const double* origin, int check_point_num, double bound_tolerance, double penalty_gain) {
  if (bound_tolerance < 0 || check_point_num <= 0 || penalty_gain <= 0.0 || penalty_gain >= 1.0) return 1.0;   // :337-341
  const MapTf tf = MakeTf(scale, off_x, off_y);
  double pm[3];
  WorldToMap(tf, pose_world, pm);
  if (!(pm[0] > 0.0 && pm[0] < size_x && pm[1] > 0.0 && pm[1] < size_y)) return 0.0;                            // :351-353
  const double c = std::cos(pm[2]), sn = std::sin(pm[2]);
  const int sx0 = static_cast<int>((pm[0] + (c * origin[0] + (-sn) * origin[1])) + 0.5);                        // :357-359
  const int sy0 = static_cast<int>((pm[1] + (sn * origin[0] + c * origin[1])) + 0.5);
  if (check_point_num == 1 && n_pts >= 2) return std::nan("");   // the reference divides by zero here (:367)
  int step = 1;
  if (n_pts < 2 * check_point_num) { check_point_num = n_pts; step = 1; }                                       // :361-368
  else step = n_pts / (check_point_num - 1);
  double penalty = 0;
  for (int p = 0; p < n_pts; p += step) {
    const double px = pts[2 * p], py = pts[2 * p + 1];
    const int ex = static_cast<int>((pm[0] + (c * px + (-sn) * py)) + 0.5);
    const int ey = static_cast<int>((pm[1] + (sn * px + c * py)) + 0.5);
    if ((sx0 == ex && sy0 == ey) || !(double(ex) > 0.0 && double(ex) < size_x && double(ey) > 0.0 && double(ey) < size_y)) continue;
    // ErgodLineBresenhami(start, end)
    double res = 0.0;
    int x0 = sx0, y0 = sy0, x1 = ex, y1 = ey;
    const bool steep = std::abs(y1 - y0) > std::abs(x1 - x0);
    if (steep) { std::swap(x0, y0); std::swap(x1, y1); }
    if (x0 > x1) { std::swap(x0, x1); std::swap(y0, y1); }
    const int delta_x = x1 - x0, delta_y = std::abs(y1 - y0);
    int error = 0, y = y0;
    const int y_step = y0 < y1 ? 1 : -1;
    for (int x = x0; x <= x1; ++x) {
      const int qx = steep ? y : x, qy = steep ? x : y;
      error += delta_y;
      if (2 * error >= delta_x) { y += y_step; error -= delta_x; }
      double penalty_sum = 0.0;
      // (a sensor cell outside the map makes the reference read out of bounds; such cells count as free here)
      if (qx >= 0 && qx < size_x && qy >= 0 && qy < size_y && occupied[static_cast<size_t>(qy) * size_x + qx]) {
        const double ddx = double(ex) - double(qx), ddy = double(ey) - double(qy);
        if (std::sqrt(ddx * ddx + ddy * ddy) > bound_tolerance) penalty_sum += 1.0;                              // slam_util.h:94-96
      }
      if (res < 1.0) res += penalty_sum;
    }
    penalty += res;
  }
  penalty *= penalty_gain;
  return std::max((1.0 + 2 * penalty_gain - penalty), 0.1);                                                      // :389-390
}

// SlamProcessor::MapCheckPenalize (slam/slam_processor.cpp:573-595), use_map_check_feedback on.
double orc_map_check_penalize(const unsigned char* occupied, int size_x, int size_y, double scale, double off_x,
                              double off_y, int n_pts, const double* pts, const double* pose_world, const double* origin,
                              int check_point_num, double bound_tolerance, double penalty_gain, int use_logistic) {
  double penalty = orc_map_feedback_penalty(occupied, size_x, size_y, scale, off_x, off_y, n_pts, pts, pose_world, origin,
                                            check_point_num, bound_tolerance, penalty_gain);
  if (use_logistic) penalty = (1 / (1 + exp(-10 * (penalty - 0.4))));
  return penalty;
}

}  // extern "C"

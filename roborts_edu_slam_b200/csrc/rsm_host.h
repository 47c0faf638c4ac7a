// rsm_host.h -- host-side exact arithmetic of the matcher: map transforms, pass geometry and the
// finalisation the reference performs on its sorted candidate vector (FindBestCandidate and the
// two covariance routines).  Header-only, plain C++; compiled for baseline x86-64 with
// -ffp-contract=off so that every expression is evaluated as the reference's build evaluates it.
// All citations are file:line under the reference's src/.
#ifndef RSM_HOST_H_
#define RSM_HOST_H_

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/rsm.h"

namespace rsm {

constexpr double kMaxVariance = 500.0;          // util/slam_util.h:57
constexpr double kDoubleTolerance = 1e-06;      // util/slam_util.h:59
constexpr int kMaxVarianceUsePointSize = 20;    // scan_match/correlate_scan_matcher.h:1033
constexpr double kResponseFilterTolerance = 1e-2;  // :763

// util/slam_util.h:70-73
inline bool DoubleEqual(double a, double b, double tolerance = kDoubleTolerance) {
  const double delta = a - b;
  return delta < 0.0 ? delta >= -std::fabs(tolerance) : delta <= std::fabs(tolerance);
}
// util/slam_util.h:75-77
inline double Round(double value) { return value >= 0.0 ? std::floor(value + 0.5) : std::ceil(value - 0.5); }

// world <-> map of GridMapBase (map/grid_map_base.h:68-93).  The reference builds
// world_to_map = AlignedScaling2d(s,s) * Translation2d(offset) and inverts it with Eigen; the
// member values and the products below follow Eigen 3.3's evaluation order: linear part diag(s,s),
// translation (s*ox, s*oy); inverse via invdet = 1/(s*s - 0*0); Affine*v = t + (l00*x + l01*y).
struct MapTransform {
  double s, tx, ty;
  double i00, i01, i10, i11, itx, ity;
  void set(double scale, double off_x, double off_y) {
    s = scale;
    tx = scale * off_x;
    ty = scale * off_y;
    const double l00 = scale, l01 = 0.0, l10 = 0.0, l11 = scale;
    const double det = l00 * l11 - l10 * l01;
    const double invdet = 1.0 / det;
    i00 = l11 * invdet;
    i10 = -l10 * invdet;
    i01 = -l01 * invdet;
    i11 = l00 * invdet;
    itx = (-i00) * tx + (-i01) * ty;
    ity = (-i10) * tx + (-i11) * ty;
  }
  void world_to_map(const double* w, double* m) const {
    m[0] = tx + (s * w[0] + 0.0 * w[1]);
    m[1] = ty + (0.0 * w[0] + s * w[1]);
    m[2] = w[2];
  }
  void map_to_world(const double* m, double* w) const {
    w[0] = itx + (i00 * m[0] + i01 * m[1]);
    w[1] = ity + (i10 * m[0] + i11 * m[1]);
    w[2] = m[2];
  }
};

// Everything about one pass that follows from (param, P, cell_len, centre).
struct PassGeo {
  int n_ang, n_xy, step, divisor, visited;
  double start_x, start_y, factor, start_angle, ares;
  double cell_len, center[3];
  int64_t n_cand() const { return int64_t(n_ang) * n_xy * n_xy; }
  double x_of(int ix) const { return start_x + ix * factor; }      // :569
  double y_of(int iy) const { return start_y + iy * factor; }      // :572
  double angle_of(int ia) const { return start_angle + ia * ares; }  // :164
  void decode(int64_t k, int* ia, int* ix, int* iy) const {
    if (k >= 0 && k <= 0x7fffffff) {            // every window up to 2^31 candidates: 32-bit divisions
      const uint32_t u = uint32_t(k), n = uint32_t(n_xy), q = u / n;
      *iy = int(u - q * n);
      const uint32_t a = q / n;
      *ix = int(q - a * n);
      *ia = int(a);
      return;
    }
    *iy = int(k % n_xy);
    *ix = int((k / n_xy) % n_xy);
    *ia = int(k / (int64_t(n_xy) * n_xy));
  }
};

// cos / sin of one search angle as the reference's std::cos / std::sin give them (correlate_scan_matcher.h:171-172).
// glibc's sincos shares its reduction and kernels with sin and cos and returns their bits (checked on 4e7 arguments with
// glibc 2.39; tests/test_host_logic.py keeps checking) at ~0.65 of the cost of the two calls.
inline void angle_trig(double angle, double* c, double* s) {
#if defined(__GLIBC__)
  ::sincos(angle, s, c);
#else
  *c = std::cos(angle);
  *s = std::sin(angle);
#endif
}

inline PassGeo make_geo(const rsm_pass_param& q, int P, double cell_len, const double* center) {
  PassGeo g;
  const double angle_offset = (q.search_angle_offset * 2) / 2;   // serach_angle_size_ / 2 (:526, :536)
  g.n_ang = static_cast<int>(std::floor(angle_offset * 2 / q.search_angle_resolution) + 1);  // :154
  g.start_angle = center[2] - angle_offset;                       // :161
  g.ares = q.search_angle_resolution;
  g.n_xy = static_cast<int>(Round(q.search_space_size / q.search_space_resolution) + 1);     // :538
  g.start_x = center[0] - (q.search_space_size / cell_len) * 0.5;  // :546
  g.start_y = center[1] - (q.search_space_size / cell_len) * 0.5;  // :547
  g.factor = q.search_space_resolution / cell_len;                 // :548
  int use = q.use_point_size;                                      // :561-566
  if (P < 2 * use) { use = P; g.step = 1; } else { g.step = P / (use - 1); }
  g.divisor = use;
  g.visited = (P + g.step - 1) / g.step;
  g.cell_len = cell_len;
  g.center[0] = center[0]; g.center[1] = center[1]; g.center[2] = center[2];
  return g;
}

struct Cand {
  double score;
  int64_t index;
};

struct BestPose {
  double x, y, angle, score;
  int n_avg;
};

inline bool by_score_desc(const Cand& a, const Cand& b) { return a.score > b.score; }

// FindBestCandidate (:670-710).  `a` = the candidates with DoubleEqual(score, top, 1e-2), sorted
// by descending score in the reference's order; a[0] is the top candidate.
// trig (optional): the pass's angle table, cos / sin / angle per search-angle index -- the values std::cos / std::sin give for
// angle_of(ia) (angle_trig), so that an averaging set of dozens of candidates costs no libm call.
inline BestPose find_best(const PassGeo& g, const Cand* a, size_t n, const double* trig = nullptr) {
  BestPose b;
  int ia, ix, iy;
  g.decode(a[0].index, &ia, &ix, &iy);
  b.x = g.x_of(ix); b.y = g.y_of(iy); b.angle = g.angle_of(ia); b.score = a[0].score;
  if (n == 1) { b.n_avg = 1; return b; }       // one candidate: the sums below are computed but not used (count > 1, :699)
  double ax = 0.0, ay = 0.0, tx = 0.0, ty = 0.0, ssum = 0.0;
  int count = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!DoubleEqual(a[i].score, b.score, kResponseFilterTolerance)) break;
    g.decode(a[i].index, &ia, &ix, &iy);
    const double s = a[i].score, ang = g.angle_of(ia);
    ax += g.x_of(ix) * s;
    ay += g.y_of(iy) * s;
    tx += (trig ? trig[3 * ia] : std::cos(ang)) * s;
    ty += (trig ? trig[3 * ia + 1] : std::sin(ang)) * s;
    ssum += s;
    ++count;
  }
  if (count > 1) {
    ax /= ssum; ay /= ssum; tx /= ssum; ty /= ssum;
    b.x = ax; b.y = ay; b.angle = std::atan2(ty, tx);
  }
  b.n_avg = count;
  return b;
}

inline double max_angular_variance(const rsm_pass_param& q) {
  return 4 * (q.search_angle_resolution * q.search_angle_resolution);  // :801
}

// ComputePositionalCovariance (:887-956).  `top` = the head of the sorted candidate vector
// (at least min(n_cand, 20) entries).  cov row-major 3x3.
inline void positional_cov(const PassGeo& g, const rsm_pass_param& q, const BestPose& best,
                           const Cand* top, size_t n_top, double* cov) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[3 * r + c] = (r == c) ? 1.0 : 0.0;
  const double mav = max_angular_variance(q);
  if (best.score < kDoubleTolerance) {
    cov[0] = kMaxVariance; cov[4] = kMaxVariance; cov[8] = mav;
    return;
  }
  double vxx = 0.0, vxy = 0.0, vyy = 0.0, norm = 0.0;
  const double bound = std::min(best.score - 0.1, 0.5);
  int used = 0;
  for (size_t i = 0; i < n_top; ++i) {
    const double s = top[i].score;
    if (s > bound && used < kMaxVarianceUsePointSize) {
      int ia, ix, iy;
      g.decode(top[i].index, &ia, &ix, &iy);
      const double x = g.x_of(ix), y = g.y_of(iy);
      norm += s;
      vxx += ((x - best.x) * (x - best.x)) * s;
      vxy += ((x - best.x) * (y - best.y) * s);
      vyy += ((y - best.y) * (y - best.y)) * s;
      ++used;
    } else {
      break;
    }
  }
  if (norm > kDoubleTolerance) {
    double xx = vxx / norm, xy = vxy / norm, yy = vyy / norm;
    const double r = q.search_space_resolution / g.cell_len;
    const double min_variance = 0.1 * (r * r);
    xx = std::max<double>(xx, min_variance);
    yy = std::max<double>(yy, min_variance);
    const double m2 = g.cell_len * g.cell_len;
    cov[0] = (xx * m2) / best.score;
    cov[1] = (xy * m2) / best.score;
    cov[3] = (xy * m2) / best.score;
    cov[4] = (yy * m2) / best.score;
    cov[8] = mav;
  }
  if (DoubleEqual(cov[0], 0.0)) cov[0] = kMaxVariance;
  if (DoubleEqual(cov[4], 0.0)) cov[4] = kMaxVariance;
}

inline double cov_score_bound(const BestPose& best) { return std::min(best.score - 0.1, 0.5); }

inline bool same_xy(const PassGeo& g, const BestPose& best, int ix, int iy) {
  const double tol = g.factor;  // linear_tolerance = search resolution / map resolution (:840)
  return DoubleEqual(g.x_of(ix), best.x, tol) && DoubleEqual(g.y_of(iy), best.y, tol);
}

// ComputeAngularCovariance (:965-1019).  `xy` = the candidates with the "same" (x,y) as the best
// pose and score >= bound, in sorted order (at least the first 20 of them).
inline void angular_cov(const PassGeo& g, const rsm_pass_param& q, const BestPose& best,
                        const Cand* xy, size_t n_xy_list, double* cov) {
  const double mav = max_angular_variance(q);
  if (best.score < kDoubleTolerance) { cov[8] = mav; return; }
  const double bound = cov_score_bound(best);
  double norm = 0.0, acc = 0.0;
  int used = 0;
  for (size_t i = 0; i < n_xy_list; ++i) {
    const double s = xy[i].score;
    if (s >= bound && used < kMaxVarianceUsePointSize) {
      int ia, ix, iy;
      g.decode(xy[i].index, &ia, &ix, &iy);
      if (same_xy(g, best, ix, iy)) {
        const double a = g.angle_of(ia);
        norm += s;
        acc += ((a - best.angle) * (a - best.angle)) * s;
        ++used;
      }
    }
  }
  double var;
  if (norm > kDoubleTolerance) var = acc / norm;   // (:1008-1012: the /4 value is overwritten)
  else var = 200 * mav;
  cov[8] = var;
}

// ---- host side of the Gauss-Newton matcher (scan_match/optimize_scan_matcher.h) --------------------
// util/slam_util.h:79-87
inline double max_abs_limit(double value, double limit) {
  const double lim = std::fabs(limit);
  if (value > lim) return lim;
  if (value < -lim) return -lim;
  return value;
}
// util/slam_util.h:103-111
inline double normalize_angle(double angle) {
  const double two_pi = 2.0 * M_PI;
  double a = std::fmod(std::fmod(angle, two_pi) + two_pi, two_pi);
  if (a > M_PI) a -= two_pi;
  return a;
}

// det = H.ldlt().solve(b) for a symmetric 3x3 H (optimize_scan_matcher.h:135-141).  Follows Eigen 3.3's
// LDLT (Cholesky/LDLT.h): in-place lower factorisation with the largest remaining diagonal entry as pivot,
// then x = P^T L^-T D^+ L^-1 P b with pivots of magnitude <= DBL_MIN treated as zero.  a[r][c] is the
// working copy; only its lower triangle is read after the swaps.
inline void ldlt3_solve(const double H[3][3], const double b[3], double x[3]) {
  double a[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[r][c] = H[r][c];
  int perm[3] = {0, 1, 2};
  double w[3];
  for (int k = 0; k < 3; ++k) {
    int piv = k;
    for (int i = k + 1; i < 3; ++i) if (std::fabs(a[i][i]) > std::fabs(a[piv][piv])) piv = i;
    perm[k] = piv;
    if (piv != k) {
      for (int c = 0; c < k; ++c) std::swap(a[k][c], a[piv][c]);               // row heads
      for (int r = piv + 1; r < 3; ++r) std::swap(a[r][k], a[r][piv]);          // column tails
      std::swap(a[k][k], a[piv][piv]);
      for (int i = k + 1; i < piv; ++i) std::swap(a[i][k], a[piv][i]);          // the strip between them
    }
    if (k > 0) {
      for (int c = 0; c < k; ++c) w[c] = a[c][c] * a[k][c];
      double dot = a[k][0] * w[0];
      for (int c = 1; c < k; ++c) dot = dot + a[k][c] * w[c];
      a[k][k] -= dot;
      for (int r = k + 1; r < 3; ++r)
        for (int c = 0; c < k; ++c) a[r][k] = a[r][k] + a[r][c] * (-1.0 * w[c]);
    }
    const double d = a[k][k];
    if (k == 0 && !(std::fabs(d) > 0.0)) { perm[0] = 0; perm[1] = 1; perm[2] = 2; break; }   // zero diagonal: nothing to eliminate
    if (std::fabs(d) > 0.0) for (int r = k + 1; r < 3; ++r) a[r][k] /= d;
  }
  for (int i = 0; i < 3; ++i) x[i] = b[i];
  for (int k = 0; k < 3; ++k) if (perm[k] != k) std::swap(x[k], x[perm[k]]);
  x[1] -= a[1][0] * x[0];
  x[2] -= (a[2][0] * x[0] + a[2][1] * x[1]);
  for (int i = 0; i < 3; ++i) { if (std::fabs(a[i][i]) > DBL_MIN) x[i] /= a[i][i]; else x[i] = 0.0; }
  x[1] -= a[2][1] * x[2];
  x[0] -= (a[1][0] * x[1] + a[2][0] * x[2]);
  for (int k = 2; k >= 0; --k) if (perm[k] != k) std::swap(x[k], x[perm[k]]);
}

// ---- map resize policy (GridMapBase::UpdateBound / ExtendSize, map/grid_map_base.h:188-274; util/boundbox.h) ----------
// Host-only bookkeeping: which scans fit the map, and when and how the map grows.  The cells live elsewhere (on the
// device); this object only reproduces the reference's decisions and geometry, bit for bit.
struct BoundBox2 {
  // BoundBox<double>: an empty box is min = (FLT_MAX, FLT_MAX), max = (FLT_MIN, FLT_MIN) -- FLT_MIN being the smallest
  // positive float, not the most negative one (util/boundbox.h:38-42)
  double min_x = double(FLT_MAX), min_y = double(FLT_MAX), max_x = double(FLT_MIN), max_y = double(FLT_MIN);
  void add_point(double x, double y) {                       // :103-109
    if (x < min_x) min_x = x;
    if (y < min_y) min_y = y;
    if (x > max_x) max_x = x;
    if (y > max_y) max_y = y;
  }
  void add_box(const BoundBox2& b) { add_point(b.min_x, b.min_y); add_point(b.max_x, b.max_y); }   // :111-115
  bool in_bounds(double x, double y) const { return x > min_x && x < max_x && y > min_y && y < max_y; }   // :122-126
  int size_x() const { return int(std::ceil(max_x) - std::floor(min_x)); }   // GetBoxSize, :85-90
  int size_y() const { return int(std::ceil(max_y) - std::floor(min_y)); }
};

struct MapBounds {
  int size_x = 0, size_y = 0;
  double scale = 1.0, off_x = 0.0, off_y = 0.0, extend_factor = 1.0;
  BoundBox2 box;                                             // GridMapBase::bound_box_
  MapTransform tf;
  // what the last extension did (valid after update() returned false)
  int pre_x = 0, pre_y = 0;

  void init(int sx, int sy, double scale_factor, double ox, double oy, double extend) {
    size_x = sx; size_y = sy; scale = scale_factor; off_x = ox; off_y = oy;
    if (extend > 0) extend_factor = extend;                   // set_extend_factor, :181-185
    box = BoundBox2();
    tf.set(scale, off_x, off_y);
  }
  bool point_in_map(double x, double y) const { return x > 0 && x < size_x && y > 0 && y < size_y; }   // :330-337

  // ExtendSize(EXTEND_PARTLY), :188-254
  void extend_partly() {
    BoundBox2 tmp;
    tmp.add_box(box);
    tmp.add_point(0.0, 0.0);
    tmp.add_point(double(size_x), double(size_y));
    double mnx = tmp.min_x, mny = tmp.min_y, mxx = tmp.max_x, mxy = tmp.max_y;
    const double bsx = double(tmp.size_x()) * extend_factor, bsy = double(tmp.size_y()) * extend_factor;
    if (box.min_x <= 0.0) mnx -= bsx;
    if (box.min_y <= 0.0) mny -= bsy;
    if (box.max_x >= double(size_x)) mxx += bsx;
    if (box.max_y >= double(size_y)) mxy += bsy;
    tmp.add_point(mnx, mny);
    tmp.add_point(mxx, mxy);
    const double fx = std::floor(tmp.min_x), fy = std::floor(tmp.min_y);
    off_x -= fx / scale;                                      // map_offset_ -= GetFloorMin() / scale_factor_
    off_y -= fy / scale;
    pre_x = int(-fx); pre_y = int(-fy);                       // pre_grid_offset = (-GetFloorMin()).cast<int>()
    size_x = tmp.size_x(); size_y = tmp.size_y();
    const BoundBox2 old = box;
    box.min_x = old.min_x - tmp.min_x; box.min_y = old.min_y - tmp.min_y;     // ResetWithBound(min - tmp.min, max - tmp.min)
    box.max_x = old.max_x - tmp.min_x; box.max_y = old.max_y - tmp.min_y;
    tf.set(scale, off_x, off_y);                              // SetMapTransform
  }

  // UpdateBound(bound_box), :257-274: true = the box is inside the map (or was already covered); false = the map was
  // extended (pre_x / pre_y / size / offset updated)
  bool update(const BoundBox2& b) {
    if (box.in_bounds(b.min_x, b.min_y) && box.in_bounds(b.max_x, b.max_y)) return true;
    box.add_box(b);
    if (!point_in_map(box.min_x, box.min_y) || !point_in_map(box.max_x, box.max_y)) { extend_partly(); return false; }
    return true;
  }

  // the bound box UpdateMapByRange hands to UpdateBound (occu_grid_map.h:278-294): hull of the scan's points in map
  // cells (pose_transform * p = t + (c x - s y, s x + c y), host libm), grown by the blur kernel's half size
  BoundBox2 scan_box(const double* pts_xy, int n, const double* pose_world, int half_kernel, bool use_blur) const {
    double pm[3];
    tf.world_to_map(pose_world, pm);
    const double c = std::cos(pm[2]), s = std::sin(pm[2]);
    BoundBox2 b;
    for (int i = 0; i < n; ++i) {
      const double px = pts_xy[2 * i], py = pts_xy[2 * i + 1];
      b.add_point(pm[0] + (c * px + (-s) * py), pm[1] + (s * px + c * py));
    }
    if (use_blur) { b.min_x -= half_kernel; b.min_y -= half_kernel; b.max_x += half_kernel; b.max_y += half_kernel; }   // :288-290
    return b;
  }
  // MapSizeCheck's box (scan_match/scan_matchers.h:365-380)
  BoundBox2 range_box(const double* pose_world, double range_max, double offset) const {
    double pm[3];
    tf.world_to_map(pose_world, pm);
    const double max_size = (range_max + offset) / (1 / scale);
    BoundBox2 b;
    b.min_x = pm[0] - max_size; b.min_y = pm[1] - max_size; b.max_x = pm[0] + max_size; b.max_y = pm[1] + max_size;
    return b;
  }
};

}  // namespace rsm
#endif

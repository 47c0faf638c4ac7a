// TEST SHIM (CPU): exposes the product's host-side arithmetic (roborts_edu_slam_b200/csrc/rsm_host.h -- the code that
// runs on the host between kernels) so that tests/test_host_logic.py can check it against the oracle without a GPU.
// hs_finalize glues the header's functions together the way run_pass's exact path (finish_exact, rsm_api.cu) does.
#include <algorithm>
#include <vector>

#include "rsm_host.h"

using namespace rsm;

extern "C" {

void hs_ldlt3(const double* H_rowmajor, const double* b, double* x) {
  double H[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[r][c] = H_rowmajor[3 * r + c];
  ldlt3_solve(H, b, x);
}
double hs_normalize_angle(double a) { return normalize_angle(a); }
// the product's angle-table entry for n angles: out[2i] = cos, out[2i+1] = sin
void hs_angle_trig(const double* angles, long n, double* out) {
  for (long i = 0; i < n; ++i) angle_trig(angles[i], &out[2 * i], &out[2 * i + 1]);
}
double hs_max_abs_limit(double v, double lim) { return max_abs_limit(v, lim); }

void hs_world_to_map(double scale, double off_x, double off_y, const double* w, double* m) {
  MapTransform t; t.set(scale, off_x, off_y); t.world_to_map(w, m);
}
void hs_map_to_world(double scale, double off_x, double off_y, const double* m, double* w) {
  MapTransform t; t.set(scale, off_x, off_y); t.map_to_world(m, w);
}

// out_i = {n_ang, n_xy, step, divisor, visited}; out_d = {start_x, start_y, factor, start_angle}
void hs_geometry(const rsm_pass_param* q, int P, double cell_len, const double* center, long* out_i, double* out_d) {
  const PassGeo g = make_geo(*q, P, cell_len, center);
  out_i[0] = g.n_ang; out_i[1] = g.n_xy; out_i[2] = g.step; out_i[3] = g.divisor; out_i[4] = g.visited;
  out_d[0] = g.start_x; out_d[1] = g.start_y; out_d[2] = g.factor; out_d[3] = g.start_angle;
}

// Everything after the scores exist, on the full score array (candidate order): returns the response;
// best = {x, y, angle, score} in map coordinates; cov row-major 3x3 in/out; n_avg = size of the averaging set.
double hs_finalize(const rsm_pass_param* q, int P, double cell_len, const double* center, const double* score, long n,
                   double* best_out, double* cov, int* n_avg) {
  const PassGeo g = make_geo(*q, P, cell_len, center);
  std::vector<Cand> c(n);
  for (long k = 0; k < n; ++k) { c[k].score = score[k]; c[k].index = k; }
  std::sort(c.begin(), c.end(), by_score_desc);
  size_t na = 0;
  while (na < c.size() && DoubleEqual(c[na].score, c[0].score, kResponseFilterTolerance)) ++na;
  const BestPose best = find_best(g, c.data(), na);
  const int type = q->type;
  if (type == RSM_COARSE || type == RSM_FINE)
    positional_cov(g, *q, best, c.data(), std::min<size_t>(c.size(), 21), cov);
  if (type == RSM_COARSE || type == RSM_SUPER) {
    std::vector<Cand> xy;
    const double bound = cov_score_bound(best);
    for (const Cand& e : c) {
      if (!(e.score >= bound)) break;
      int ia, ix, iy;
      g.decode(e.index, &ia, &ix, &iy);
      if (same_xy(g, best, ix, iy)) { xy.push_back(e); if (xy.size() >= size_t(kMaxVarianceUsePointSize)) break; }
    }
    angular_cov(g, *q, best, xy.data(), xy.size(), cov);
  }
  best_out[0] = best.x; best_out[1] = best.y; best_out[2] = best.angle; best_out[3] = best.score;
  *n_avg = best.n_avg;
  return std::min(best.score, 1.0);
}

// map resize policy: one object, updated scan by scan; geom = {size_x, size_y, off_x, off_y, pre_x, pre_y}
void* hs_bounds_create(int sx, int sy, double scale, double ox, double oy, double extend) {
  MapBounds* b = new MapBounds;
  b->init(sx, sy, scale, ox, oy, extend);
  return b;
}
void hs_bounds_destroy(void* h) { delete static_cast<MapBounds*>(h); }
static void bounds_geom(const MapBounds* b, double* geom) {
  geom[0] = b->size_x; geom[1] = b->size_y; geom[2] = b->off_x; geom[3] = b->off_y; geom[4] = b->pre_x; geom[5] = b->pre_y;
}
int hs_bounds_update_scan(void* h, const double* pts, int n, const double* pose, int half_kernel, int use_blur, double* geom) {
  MapBounds* b = static_cast<MapBounds*>(h);
  const bool ok = b->update(b->scan_box(pts, n, pose, half_kernel, use_blur != 0));
  bounds_geom(b, geom);
  return ok ? 1 : 0;
}
int hs_bounds_size_check(void* h, const double* pose, double range_max, double offset, double* geom) {
  MapBounds* b = static_cast<MapBounds*>(h);
  const bool ok = b->update(b->range_box(pose, range_max, offset));
  bounds_geom(b, geom);
  return ok ? 1 : 0;
}

}  // extern "C"

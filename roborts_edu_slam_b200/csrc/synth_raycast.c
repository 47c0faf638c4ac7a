/* Synthetic workload generation (not on the hot path, not the oracle): ray-casts laser
 * scans out of an occupancy bitmap the way SURVEY.md section 8(d) defines the benchmark
 * inputs.  Pixel (px,py), py = 0 at the bottom, covers [px*res,(px+1)*res) x [py*res,...).
 * A sample outside the bitmap counts as occupied.  For beam i the sensor-frame bearing is
 * a_i = -fov/2 + fov*i/(beams-1); the ray is sampled at r = r0 + k*dr until it hits or
 * reaches r_max; a hit is kept only when r < keep_frac*r_max.  Kept points are written as
 * (cos a_i * r, sin a_i * r) in METRES, sensor frame.  Returns the number of points kept.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

int synth_raycast(const uint8_t* occ, int width, int height, double res,
                  double x, double y, double heading,
                  int beams, double fov, double r_max, double r0, double dr, double keep_frac,
                  double* out_xy) {
  int n = 0;
  const double keep = keep_frac * r_max;
  for (int i = 0; i < beams; ++i) {
    const double a = (beams > 1) ? (-0.5 * fov + fov * (double)i / (double)(beams - 1)) : 0.0;
    const double c = cos(heading + a), s = sin(heading + a);
    for (int k = 0;; ++k) {
      const double r = r0 + (double)k * dr;
      if (r >= r_max) break;
      const double wx = x + c * r, wy = y + s * r;
      const int px = (int)floor(wx / res), py = (int)floor(wy / res);
      int hit = (px < 0 || py < 0 || px >= width || py >= height);
      if (!hit) hit = occ[(size_t)py * width + px] != 0;
      if (hit) {
        if (r < keep) {
          out_xy[2 * n] = cos(a) * r;
          out_xy[2 * n + 1] = sin(a) * r;
          ++n;
        }
        break;
      }
    }
  }
  return n;
}

/* Distance (in pixels, Chebyshev) from (px,py) to the nearest occupied pixel or map edge is
 * at least `clear` -> 1, else 0.  Used to draw free-space poses for the batched workload. */
int synth_is_clear(const uint8_t* occ, int width, int height, int px, int py, int clear) {
  if (px - clear < 0 || py - clear < 0 || px + clear >= width || py + clear >= height) return 0;
  for (int j = -clear; j <= clear; ++j)
    for (int i = -clear; i <= clear; ++i)
      if (occ[(size_t)(py + j) * width + (px + i)]) return 0;
  return 1;
}

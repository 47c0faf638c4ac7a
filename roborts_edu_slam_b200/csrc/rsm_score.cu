// rsm_score.cu -- the scoring kernel: every (angle, x, y) candidate x every visited beam.
//
// Reference: MultiResolutionCorrelateScanMatcher::ScanMatch's triple loop + GetResponse +
// PenalizeResponse (scan_match/correlate_scan_matcher.h:552-603, 637-662, 718-745).
//
// Mapping (DESIGN.md section 4 has the long version).  One CTA of 128 threads owns one search
// angle and one tile of LX x rows translations; lanes walk consecutive x translations, so a warp
// load touches ONE grid row segment (coalesced, 1-2 cache lines) instead of 32 scattered cells;
// each thread owns RY consecutive y translations and keeps their sums in registers: no cross-lane
// reduction, and the float32 fallback adds in the reference's beam order.  The cell index of a
// (beam, candidate) pair separates into gx(beam, x) and gy(beam, y), each produced in FP64 with
// the reference's exact operation order (no FMA) -- per (beam, x) and (beam, y), never per
// candidate -- and staged through double-buffered shared-memory tables.
//
//   general variant   tables hold gx[beam][x] and gy*pitch[beam][y]; inner loop per beam:
//                     1 table load + RY x (row-offset load, address add, 4-byte gather, add)
//   affine variant    (search step an exact integer number F of cells: the coarse passes).  With
//                     t0 = (lut + first candidate of the tile) + 0.5, every other candidate of the
//                     tile differs from t0 by j*F up to < 1e-9 of accumulated rounding (three
//                     roundings per coordinate, |coordinates| < 2^20), so whenever frac(t0) lies in
//                     (1e-6, 1 - 1e-6) the truncated index of candidate j is PROVABLY
//                     trunc(t0) + j*F.  The table then holds one int per (beam, y-slot):
//                     gy0*pitch + gx0; lanes add tx*F, rows are reached by constant strides.  The
//                     ~4e-6 of (beam, slot) pairs that fail the test (or touch the grid border)
//                     recompute their indices exactly per thread.  No per-candidate FP64 at all.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
#include <cuda_runtime.h>
#include <stdint.h>

#include "rsm_device.h"
#include "rsm_kernels.h"

namespace rsm {

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ unsigned long long score_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

__device__ __forceinline__ int find_job(const int* __restrict__ cta_begin, int n_jobs, int b) {
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(cta_begin + mid) <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <int RYP> struct RowVec;
template <> struct RowVec<1> { __device__ static void load(const int* p, int* r) { r[0] = p[0]; } };
template <> struct RowVec<2> { __device__ static void load(const int* p, int* r) { int2 v = *reinterpret_cast<const int2*>(p); r[0] = v.x; r[1] = v.y; } };
template <> struct RowVec<4> { __device__ static void load(const int* p, int* r) { int4 v = *reinterpret_cast<const int4*>(p); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; } };
template <> struct RowVec<8> { __device__ static void load(const int* p, int* r) { RowVec<4>::load(p, r); RowVec<4>::load(p + 4, r + 4); } };

// threads per CTA: 4 y-slots of 32 lanes, or 8 y-slots of 16 / 8 / 4 lanes
__host__ __device__ constexpr int threads_of(int lx) { return lx >= 16 ? 128 : lx * 8; }
__host__ __device__ constexpr int ryp_of(int ry) { return (ry <= 1) ? 1 : (ry <= 2) ? 2 : (ry <= 4) ? 4 : 8; }
// resident-thread target per SM: 1536 (<= 40 registers) for RY <= 4, 1280 (<= 48) above
__host__ __device__ constexpr int min_blocks_of(int ry, int nt) {
  return ((ry <= 4 ? 1536 : 1280) / nt) > 32 ? 32 : ((ry <= 4 ? 1536 : 1280) / nt);
}

// (int)(a + b + 0.5): static_cast<int> truncation of the reference (:647-648)
__device__ __forceinline__ int cell_index(double lut, double cand) {
  return __double2int_rz(dadd(dadd(lut, cand), 0.5));
}

// PITCH > 0: AFFINE with a unit search step on a grid of that compile-time row pitch -- the RY rows
// of a thread are then reached with immediate offsets.
template <bool FIXED, bool AFFINE, int LX, int RY, int PITCH>
__global__ void __launch_bounds__(threads_of(LX), min_blocks_of(RY, threads_of(LX)))
score_kernel(const ScoreJob* __restrict__ jobs, const int* __restrict__ cta_begin, int n_jobs) {
  constexpr int RYP = ryp_of(RY);
  constexpr int PC = kChunk;
  constexpr int NT = threads_of(LX);
  constexpr int SLOTS = NT / LX;      // y slots per CTA; each owns RY consecutive y translations
  constexpr int ROWS = SLOTS * RY;
  constexpr int GXW = AFFINE ? 1 : LX;                // ints per beam in the x table (unused if affine)
  constexpr int GYW = AFFINE ? SLOTS : SLOTS * RYP;   // ints per beam in the y table
  const int tid = threadIdx.x;
  const int tx = tid % LX;       // lane position along x
  const int ts = tid / LX;       // y slot

  __shared__ ScoreJob J;
  __shared__ int s_job;
  __shared__ unsigned long long s_wmax[NT / 32];
  __shared__ double sLut[2][PC][2];     // rotated endpoints (x, y) of the chunk's beams
  __shared__ double sX[LX];             // candidate x of this tile
  __shared__ double sY[ROWS];           // candidate y of this tile
  __shared__ __align__(16) int sGX[2][PC][GXW];  // general: cell x per (beam, lane)
  __shared__ __align__(16) int sGY[2][PC][GYW];  // general: cell y * pitch per (beam, row);
                                                 // affine: gy0*pitch + gx0 per (beam, slot), -1 = recompute
  __shared__ int sAff[3][SLOTS];        // affine: 1 while every beam of chunk c passed the test for the slot
                                        // (indexed c % 3: reset one iteration before it is rebuilt)

  if (tid == 0) s_job = find_job(cta_begin, n_jobs, blockIdx.x);
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(jobs + s_job);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < int(sizeof(ScoreJob) / 4); i += NT) dst[i] = __ldg(src + i);
  }
  const int first_cta = __ldg(cta_begin + s_job);
  __syncthreads();

  const int local = blockIdx.x - first_cta;
  const int tiles = J.tiles_x * J.tiles_y;
  const int ia_local = local / tiles;
  const int tile = local - ia_local * tiles;
  const int tx0 = (tile % J.tiles_x) * LX;
  const int ty0 = (tile / J.tiles_x) * ROWS;
  const int ia = J.ang_begin + ia_local;
  const double cs = __ldg(J.trig + 3 * ia), sn = __ldg(J.trig + 3 * ia + 1), ang = __ldg(J.trig + 3 * ia + 2);
  const int V = J.V, n_xy = J.n_xy, pitch = J.pitch, size_x = J.size_x, size_y = J.size_y;
  const int stepoff = PITCH > 0 ? PITCH : J.stepoff;   // affine: (search step in cells) * pitch
  const int nchunks = (V + PC - 1) / PC;

  if (AFFINE && tid < SLOTS) { sAff[0][tid] = 1; sAff[1][tid] = 1; sAff[2][tid] = 1; }
  // candidate coordinates of this tile: x = start_x + x_index * factor   (:569, :572)
  for (int i = tid; i < LX + ROWS; i += NT) {
    if (i < LX) sX[i] = dadd(J.sx, dmul((double)(tx0 + i), J.f));
    else sY[i - LX] = dadd(J.sy, dmul((double)(ty0 + i - LX), J.f));
  }

  // rotated endpoint of visited beam v: (cos*px - sin*py, sin*px + cos*py)   (:179-180)
  auto lut_chunk = [&](int c) {
    if (tid < PC) {
      const int v = c * PC + tid;
      if (v < V) {
        const int p = v * J.step;
        const double px = __ldg(J.pts + 2 * p), py = __ldg(J.pts + 2 * p + 1);
        sLut[c & 1][tid][0] = dsub(dmul(cs, px), dmul(sn, py));
        sLut[c & 1][tid][1] = dadd(dmul(sn, px), dmul(cs, py));
      }
    }
  };
  // cell index tables of chunk c; returns kErrWindow if a real candidate's cell left the grid
  auto build_chunk = [&](int c) -> int {
    int e = 0;
    const int b = c & 1;
    const int npc = min(PC, V - c * PC);
    if (AFFINE) {
      const int F = J.f_int;
      for (int q = tid; q < npc * SLOTS; q += NT) {
        const int pc = q / SLOTS, sl = q % SLOTS;
        const double tx_ = dadd(dadd(sLut[b][pc][0], sX[0]), 0.5);
        const double ty_ = dadd(dadd(sLut[b][pc][1], sY[sl * RY]), 0.5);
        const int gx0 = __double2int_rz(tx_), gy0 = __double2int_rz(ty_);
        const double fx = tx_ - (double)gx0, fy = ty_ - (double)gy0;
        const bool ok = fx > 1e-6 && fx < 1.0 - 1e-6 && fy > 1e-6 && fy < 1.0 - 1e-6 &&
                        gx0 >= 0 && gx0 + (LX - 1) * F < size_x && gy0 >= 0 && gy0 + (RY - 1) * F < size_y;
        sGY[b][pc][sl] = ok ? gy0 * pitch + gx0 : -1;
        if (!ok) sAff[c % 3][sl] = 0;
      }
    } else {
      for (int q = tid; q < npc * LX; q += NT) {
        const int pc = q / LX, j = q % LX;
        int g = cell_index(sLut[b][pc][0], sX[j]);
        if (g < 0 || g >= size_x) {
          if (tx0 + j < n_xy) e = kErrWindow;
          g = max(0, min(g, size_x - 1));
        }
        sGX[b][pc][j] = g;
      }
      for (int q = tx; q < npc * RY; q += LX) {
        const int pc = q / RY, rr = q % RY;
        int g = cell_index(sLut[b][pc][1], sY[ts * RY + rr]);
        if (g < 0 || g >= size_y) {
          if (ty0 + ts * RY + rr < n_xy) e = kErrWindow;
          g = max(0, min(g, size_y - 1));
        }
        sGY[b][pc][ts * RYP + rr] = g * pitch;
      }
    }
    return e;
  };

  unsigned int a32[RY];
  unsigned long long a64[RY];
  double ad[RY];
#pragma unroll
  for (int r = 0; r < RY; ++r) { a32[r] = 0u; a64[r] = 0ull; ad[r] = 0.0; }

  const int* gridI = reinterpret_cast<const int*>(J.grid);
  const float* gridF = reinterpret_cast<const float*>(J.grid);
  int err = 0;

  lut_chunk(0);
  __syncthreads();
  err |= build_chunk(0);
  if (nchunks > 1) lut_chunk(1);
  __syncthreads();

  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) err |= build_chunk(c + 1);
    if (c + 2 < nchunks) lut_chunk(c + 2);

    const int b = c & 1;
    const int npc = min(PC, V - c * PC);
    if (AFFINE) {
      const int lane_off = tx * J.f_int;
      if (tid < SLOTS) sAff[(c + 2) % 3][tid] = 1;   // chunk c+2 is built next iteration
      if (sAff[c % 3][ts]) {
        // fast path: every beam of the chunk reaches its RY rows by constant strides
#pragma unroll 2
        for (int pc = 0; pc < npc; ++pc) {
          const int base = sGY[b][pc][ts] + lane_off;
#pragma unroll
          for (int r = 0; r < RY; ++r) {
            if (FIXED) a32[r] += (unsigned int)__ldg(gridI + (base + r * stepoff));
            else ad[r] = dadd(ad[r], (double)__ldg(gridF + (base + r * stepoff)));
          }
        }
      } else {
        for (int pc = 0; pc < npc; ++pc) {
          const int b0 = sGY[b][pc][ts];
          if (b0 >= 0) {
            const int base = b0 + lane_off;
#pragma unroll
            for (int r = 0; r < RY; ++r) {
              if (FIXED) a32[r] += (unsigned int)__ldg(gridI + (base + r * stepoff));
              else ad[r] = dadd(ad[r], (double)__ldg(gridF + (base + r * stepoff)));
            }
          } else {
            // exact recomputation for this (beam, slot): frac(t0) too close to a cell boundary,
            // or the tile touches the grid border
            int gx = cell_index(sLut[b][pc][0], sX[tx]);
            if (gx < 0 || gx >= size_x) {
              if (tx0 + tx < n_xy) err |= kErrWindow;
              gx = max(0, min(gx, size_x - 1));
            }
            const double ly = sLut[b][pc][1];
#pragma unroll
            for (int r = 0; r < RY; ++r) {
              int g = cell_index(ly, sY[ts * RY + r]);
              if (g < 0 || g >= size_y) {
                if (ty0 + ts * RY + r < n_xy) err |= kErrWindow;
                g = max(0, min(g, size_y - 1));
              }
              if (FIXED) a32[r] += (unsigned int)__ldg(gridI + (g * pitch + gx));
              else ad[r] = dadd(ad[r], (double)__ldg(gridF + (g * pitch + gx)));
            }
          }
        }
      }
    } else {
#pragma unroll 2
      for (int pc = 0; pc < npc; ++pc) {
        const int gx = sGX[b][pc][tx];
        int ro[RYP];
        RowVec<RYP>::load(&sGY[b][pc][ts * RYP], ro);
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          if (FIXED) a32[r] += (unsigned int)__ldg(gridI + (ro[r] + gx));
          else ad[r] = dadd(ad[r], (double)__ldg(gridF + (ro[r] + gx)));
        }
      }
    }
    if (FIXED) {
      // <= 32 cells of <= 2^25 each fit a uint32; spill into the 64-bit sum once per chunk
#pragma unroll
      for (int r = 0; r < RY; ++r) { a64[r] += a32[r]; a32[r] = 0u; }
    }
    __syncthreads();
  }

  // epilogue: response = sum / divisor (:659), centre penalty (:727-743), store, block maximum
  const int ix = tx0 + tx;
  unsigned long long kmax = 0ull;
  if (ix < n_xy) {
    const double x = sX[tx];
    const double dx = dsub(x, J.cx);
    const double dx2 = dmul(dx, dx);
    const double da = dsub(ang, J.ca);
    const double a2 = dmul(da, da);
    const double ap = fmax(dsub(1.0, ddiv(dmul(0.25, a2), 0.349)), 0.9);
    double* out = J.score + ((long long)ia_local * n_xy + ix) * n_xy;
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int iy = ty0 + ts * RY + r;
      if (iy < n_xy) {
        double sum = FIXED ? dmul((double)a64[r], kFixScale) : ad[r];
        double sc = ddiv(sum, J.divisor);
        if (J.use_penalty) {
          // DoubleEqual(score, 0.0) with the default 1e-6 tolerance skips the penalty (:728)
          const bool zero = sc < 0.0 ? (sc >= -1e-06) : (sc <= 1e-06);
          if (!zero) {
            const double dy = dsub(sY[ts * RY + r], J.cy);
            double d2 = dadd(dx2, dmul(dy, dy));
            d2 = dmul(d2, J.m2);
            const double dp = fmax(dsub(1.0, ddiv(dmul(J.gain, d2), J.half_size)), 0.5);
            sc = dmul(sc, dmul(dp, ap));
          }
        }
        out[iy] = sc;
        const unsigned long long k = score_key(sc);
        kmax = k > kmax ? k : kmax;
      }
    }
  }
  kmax = warp_max_u64(kmax);
  if ((tid & 31) == 0) s_wmax[tid >> 5] = kmax;
  if (err) atomicOr(J.err, err);
  __syncthreads();
  if (tid == 0) {
    unsigned long long m = 0ull;
    for (int w = 0; w < NT / 32; ++w) m = s_wmax[w] > m ? s_wmax[w] : m;
    atomicMax(J.best_key, m);
  }
}

// ---- variant table ----------------------------------------------------------------------------
typedef void (*ScoreFn)(const ScoreJob*, const int*, int);

template <bool FIXED, bool AFFINE, int LX, int PITCH>
static ScoreFn pick_ry(int ry) {
  switch (ry) {
    case 1: return score_kernel<FIXED, AFFINE, LX, 1, PITCH>;
    case 2: return score_kernel<FIXED, AFFINE, LX, 2, PITCH>;
    case 3: return score_kernel<FIXED, AFFINE, LX, 3, PITCH>;
    case 4: return score_kernel<FIXED, AFFINE, LX, 4, PITCH>;
    case 5: return score_kernel<FIXED, AFFINE, LX, 5, PITCH>;
    case 6: return score_kernel<FIXED, AFFINE, LX, 6, PITCH>;
    case 7: return score_kernel<FIXED, AFFINE, LX, 7, PITCH>;
    case 8: return score_kernel<FIXED, AFFINE, LX, 8, PITCH>;
    default: return nullptr;
  }
}

template <bool FIXED>
static ScoreFn pick_lx(bool affine, int lx, int ry, int pitch) {
  if (affine) {
    if (FIXED && pitch == kPitchSmall) {
      switch (lx) {
        case 16: return pick_ry<FIXED, true, 16, kPitchSmall>(ry);
        case 32: return pick_ry<FIXED, true, 32, kPitchSmall>(ry);
        default: break;
      }
    }
    if (FIXED && pitch == kPitchLarge) {
      switch (lx) {
        case 16: return pick_ry<FIXED, true, 16, kPitchLarge>(ry);
        case 32: return pick_ry<FIXED, true, 32, kPitchLarge>(ry);
        default: break;
      }
    }
    switch (lx) {
      case 4: return pick_ry<FIXED, true, 4, 0>(ry);
      case 8: return pick_ry<FIXED, true, 8, 0>(ry);
      case 16: return pick_ry<FIXED, true, 16, 0>(ry);
      case 32: return pick_ry<FIXED, true, 32, 0>(ry);
      default: return nullptr;
    }
  }
  switch (lx) {
    case 4: return pick_ry<FIXED, false, 4, 0>(ry);
    case 8: return pick_ry<FIXED, false, 8, 0>(ry);
    case 16: return pick_ry<FIXED, false, 16, 0>(ry);
    case 32: return pick_ry<FIXED, false, 32, 0>(ry);
    default: return nullptr;
  }
}

// const_pitch: kPitchSmall / kPitchLarge when every job of the launch has that row pitch AND a unit
// search step (immediate-offset variant), else 0
static ScoreFn score_fn(bool fixed, bool affine, int lx, int ry, int const_pitch) {
  return fixed ? pick_lx<true>(affine, lx, ry, const_pitch) : pick_lx<false>(affine, lx, ry, 0);
}

int score_threads(int lx) { return threads_of(lx); }

// resident CTAs per SM of a variant (0 if the variant does not exist)
int score_occupancy(bool fixed, bool affine, int lx, int ry, int const_pitch) {
  ScoreFn fn = score_fn(fixed, affine, lx, ry, const_pitch);
  if (!fn) return 0;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads_of(lx), 0) != cudaSuccess) { cudaGetLastError(); return 0; }
  return nb;
}

cudaError_t launch_score(bool fixed, bool affine, int lx, int ry, int const_pitch, int n_cta, cudaStream_t st,
                         const ScoreJob* jobs, const int* cta_begin, int n_jobs) {
  ScoreFn fn = score_fn(fixed, affine, lx, ry, const_pitch);
  if (!fn) return cudaErrorInvalidValue;
  fn<<<n_cta, threads_of(lx), 0, st>>>(jobs, cta_begin, n_jobs);
  return cudaGetLastError();
}

}  // namespace rsm

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
// Kernels in this file, by the windows they take (rsm_api.cu, pass_begin, picks):
//   score_stream_kernel<Map>   >= 48 translations per axis, unit step, fixed point, from half a wave of (angle, tile) items on:
//                              persistent CTAs over one beam sequence, TMA boxes in shared memory (the headline kernel)
//   score_staged_kernel<RX,RY> the same windows in small launches: thread-block clusters share an item's beams
//   score_patch_kernel<NR>     batches of unit-step windows of <= 16 translations (the coarse pass of a back-end chain)
//   score_flat_kernel<F,K>     3 .. 12 translations at a non-integer step (fine / super-fine passes)
//   score_kernel<...>          everything else (the tiled L1 kernel described above)
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <type_traits>

#include "rsm_device.h"
#include "rsm_kernels.h"
#include "rsm_select.cuh"

namespace rsm {

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
  __shared__ double sLut[3][PC][2];     // rotated endpoints (x, y) of the chunk's beams, indexed chunk % 3: while chunk c
                                        // is gathered (its exact-recompute path reads them), chunk c+1's tables are built
                                        // from theirs and chunk c+2's are written
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

  // small launches: S CTAs share the beams of one (angle, tile) -- integer partial sums, so the order is free
  const int S = (FIXED && J.n_split > 1) ? J.n_split : 1;
  const int local = (blockIdx.x - first_cta) / S;
  const int split = (blockIdx.x - first_cta) - local * S;
  const int tiles = J.tiles_x * J.tiles_y;
  const int ia_local = local / tiles;
  const int tile = local - ia_local * tiles;
  const int tx0 = (tile % J.tiles_x) * LX;
  const int ty0 = (tile / J.tiles_x) * ROWS;
  const int ia = J.ang_begin + ia_local;
  const double cs = __ldg(J.trig + 3 * ia), sn = __ldg(J.trig + 3 * ia + 1), ang = __ldg(J.trig + 3 * ia + 2);
  const int V = J.V, n_xy = J.n_xy, pitch = J.pitch, size_x = J.size_x, size_y = J.size_y;
  const int stepoff = PITCH > 0 ? PITCH : J.stepoff;   // affine: (search step in cells) * pitch
  const int nchunks_all = (V + PC - 1) / PC;
  const int per_split = (nchunks_all + S - 1) / S;
  const int c_begin = min(split * per_split, nchunks_all);
  const int nchunks = min(nchunks_all, c_begin + per_split);      // this CTA gathers chunks [c_begin, nchunks)

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
        sLut[c % 3][tid][0] = dsub(dmul(cs, px), dmul(sn, py));
        sLut[c % 3][tid][1] = dadd(dmul(sn, px), dmul(cs, py));
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
        const double tx_ = dadd(dadd(sLut[c % 3][pc][0], sX[0]), 0.5);
        const double ty_ = dadd(dadd(sLut[c % 3][pc][1], sY[sl * RY]), 0.5);
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
        int g = cell_index(sLut[c % 3][pc][0], sX[j]);
        if (g < 0 || g >= size_x) {
          if (tx0 + j < n_xy) e = kErrWindow;
          g = max(0, min(g, size_x - 1));
        }
        sGX[b][pc][j] = g;
      }
      for (int q = tx; q < npc * RY; q += LX) {
        const int pc = q / RY, rr = q % RY;
        int g = cell_index(sLut[c % 3][pc][1], sY[ts * RY + rr]);
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

  if (c_begin < nchunks) lut_chunk(c_begin);
  __syncthreads();
  if (c_begin < nchunks) err |= build_chunk(c_begin);
  if (c_begin + 1 < nchunks) lut_chunk(c_begin + 1);
  __syncthreads();

  for (int c = c_begin; c < nchunks; ++c) {
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
            int gx = cell_index(sLut[c % 3][pc][0], sX[tx]);
            if (gx < 0 || gx >= size_x) {
              if (tx0 + tx < n_xy) err |= kErrWindow;
              gx = max(0, min(gx, size_x - 1));
            }
            const double ly = sLut[c % 3][pc][1];
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

  const int ix = tx0 + tx;
  if (FIXED && S > 1) {
    // add this CTA's partial sums; the last CTA of the (angle, tile) to arrive reads the totals and finishes
    __shared__ int s_ticket;
    if (ix < n_xy) {
      unsigned long long* acc = J.acc + ((long long)ia_local * n_xy + ix) * n_xy;
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        const int iy = ty0 + ts * RY + r;
        if (iy < n_xy && a64[r]) atomicAdd(acc + iy, a64[r]);
      }
    }
    if (err) atomicOr(J.err, err);
    err = 0;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(J.tickets + local, 1);
    __syncthreads();
    if (s_ticket != S - 1) return;
    __threadfence();
    if (ix < n_xy) {
      const unsigned long long* acc = J.acc + ((long long)ia_local * n_xy + ix) * n_xy;
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        const int iy = ty0 + ts * RY + r;
        if (iy < n_xy) a64[r] = __ldcg(acc + iy);
      }
    }
  }
  // epilogue: response = sum / divisor (:659), centre penalty (:727-743), store, block maximum
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

// =================================================================================================
// Patch kernel: small unit-step windows (<= 16 translations per axis) of a BATCH of matches, fixed point.
// =================================================================================================
// The back-end chain's coarse pass is 13 x 13 translations: in the tiled kernel a warp load covers two
// 16-cell row segments (13 useful) that straddle L1 lines, ~9 evaluations per L1 cycle.  Here a CTA owns
// kAngles consecutive search angles of one match, one warp per angle.  Per chunk of 32 beams lane b of warp a
// computes the patch origin of (angle a, beam b) in FP64 (reference association, same provable index test as
// above); the CTA takes the hull of the 256 origins, and if hull + patch fits the 112 x 96-cell tile the CTA copies
// that rectangle from the grid into shared memory (16-byte loads), each cell then serving up to 256 patches.
// A warp gathers its angle's window with NR shared-memory loads per beam: lanes 0-15 read row r, lanes 16-31
// row r+1; the tile pitch of 112 = 16 (mod 32) puts the two half-warps into disjoint banks.  The origin of beam j
// reaches the warp by shuffle.  Chunks whose hull does not fit (far beams, where eight angles fan out) gather
// straight from global memory with the same origins; beams that fail the index test take exact per-thread indices.
namespace patch {
#ifdef RSM_STAGED_DEBUG
__device__ unsigned long long g_pdbg[8];
#define PDBG_ADD(i, v) atomicAdd(&g_pdbg[i], (unsigned long long)(v))
#else
#define PDBG_ADD(i, v)
#endif
constexpr int kAngles = 8;
constexpr int kThreads = kAngles * 32;
constexpr int kBoxW = 112, kBoxH = 96;   // capacity; only the hull of a chunk is copied
constexpr int kPC = 32;

template <int NR>
__global__ void __launch_bounds__(kThreads, 4)
score_patch_kernel(const ScoreJob* __restrict__ jobs, const int* __restrict__ cta_begin, int n_jobs) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ ScoreJob J;
  __shared__ int s_job;
  __shared__ __align__(16) int tile[kBoxW * kBoxH];
  __shared__ __align__(16) int sBase[kAngles][kPC];   // origins of the chunk's beams that passed the index test, compacted
  __shared__ int s_box[kAngles][4];
  __shared__ unsigned long long s_wmax[kAngles];

  if (tid == 0) s_job = find_job(cta_begin, n_jobs, blockIdx.x);
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(jobs + s_job);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < int(sizeof(ScoreJob) / 4); i += kThreads) dst[i] = __ldg(src + i);
  }
  const int first_cta = __ldg(cta_begin + s_job);
  __syncthreads();

  const int ia_local = (blockIdx.x - first_cta) * kAngles + warp;
  const bool active = ia_local < J.ang_count;
  const int ia = J.ang_begin + (active ? ia_local : J.ang_count - 1);
  const double cs = __ldg(J.trig + 3 * ia), sn = __ldg(J.trig + 3 * ia + 1), ang = __ldg(J.trig + 3 * ia + 2);
  const int V = J.V, n_xy = J.n_xy, pitch = J.pitch, size_x = J.size_x, size_y = J.size_y;
  const int ixl = lane & 15, hp = lane >> 4;
  const double x0 = dadd(J.sx, dmul(0.0, J.f)), y0 = dadd(J.sy, dmul(0.0, J.f));   // candidate (0, 0)   (:569, :572)
  const double xc = dadd(J.sx, dmul((double)ixl, J.f));
  const int* __restrict__ grid = reinterpret_cast<const int*>(J.grid);
  const int nchunks = (V + kPC - 1) / kPC;
  const int* const q_tile = tile + hp * kBoxW + ixl;              // this lane's cell of a patch whose origin is tile[0]
  const int* const q_grid = grid + hp * pitch + ixl;
  const int* const bases = sBase[warp];

  unsigned int a32[NR];
  unsigned long long a64[NR];
#pragma unroll
  for (int k = 0; k < NR; ++k) { a32[k] = 0u; a64[k] = 0ull; }
  int err = 0;

  for (int c = 0; c < nchunks; ++c) {
    const int v = c * kPC + lane;
    const bool vb = v < V;
    double lx = 0.0, ly = 0.0;
    int gx0 = 0, gy0 = 0;
    bool ok = false;
    if (vb) {
      const int p = v * J.step;
      const double px = __ldg(J.pts + 2 * p), py = __ldg(J.pts + 2 * p + 1);
      lx = dsub(dmul(cs, px), dmul(sn, py));                      // :179-180
      ly = dadd(dmul(sn, px), dmul(cs, py));
      const double tx_ = dadd(dadd(lx, x0), 0.5), ty_ = dadd(dadd(ly, y0), 0.5);
      gx0 = __double2int_rz(tx_); gy0 = __double2int_rz(ty_);
      const double fx = tx_ - (double)gx0, fy = ty_ - (double)gy0;
      ok = active && fx > 1e-6 && fx < 1.0 - 1e-6 && fy > 1e-6 && fy < 1.0 - 1e-6 &&
           gx0 >= 0 && gx0 + 15 < size_x && gy0 >= 0 && gy0 + 2 * NR - 1 < size_y;
    }
    // hull of the origins of the CTA's 8 angles x 32 beams; copy it to shared memory if hull + patch fit the tile
    {
      const int xmin = __reduce_min_sync(0xffffffffu, ok ? gx0 : 0x7fffffff), xmax = __reduce_max_sync(0xffffffffu, ok ? gx0 : -1);
      const int ymin = __reduce_min_sync(0xffffffffu, ok ? gy0 : 0x7fffffff), ymax = __reduce_max_sync(0xffffffffu, ok ? gy0 : -1);
      if (lane == 0) { s_box[warp][0] = xmin; s_box[warp][1] = xmax; s_box[warp][2] = ymin; s_box[warp][3] = ymax; }
    }
    __syncthreads();
    int bx0 = 0x7fffffff, bx1 = -1, by0 = 0x7fffffff, by1 = -1;
#pragma unroll
    for (int w = 0; w < kAngles; ++w) {
      bx0 = min(bx0, s_box[w][0]); bx1 = max(bx1, s_box[w][1]); by0 = min(by0, s_box[w][2]); by1 = max(by1, s_box[w][3]);
    }
    const int xl = bx0 & ~3;                                       // 16-byte aligned rows
    const bool fits = bx1 >= 0 && (bx1 - xl + 16 <= kBoxW) && (by1 - by0 + 2 * NR <= kBoxH);
    if (tid == 0) { PDBG_ADD(0, 1); PDBG_ADD(1, fits ? 1 : 0); PDBG_ADD(2, bx1 >= 0 ? bx1 - xl + 16 : 0); PDBG_ADD(3, bx1 >= 0 ? by1 - by0 + 2 * NR : 0); }
    // the warp's safe beams, compacted: the gather loop below has no per-beam branch
    const unsigned int safe_mask = __ballot_sync(0xffffffffu, ok);
    const unsigned int slow_mask = __ballot_sync(0xffffffffu, vb && active && !ok);
    const int n_safe = __popc(safe_mask);
    if (ok) sBase[warp][__popc(safe_mask & ((1u << lane) - 1u))] = fits ? (gy0 - by0) * kBoxW + (gx0 - xl) : gy0 * pitch + gx0;
    if (fits) {
      const int w4 = (bx1 - xl + 16 + 3) >> 2, hh = by1 - by0 + 2 * NR;       // the hull, in int4 columns x rows
      for (int q = tid; q < w4 * hh; q += kThreads) {
        const int r = q / w4, c4 = q - r * w4;
        const int gy = by0 + r, gx = xl + 4 * c4;
        int4 val = make_int4(0, 0, 0, 0);
        if (gy < size_y && gx + 3 < pitch) val = __ldg(reinterpret_cast<const int4*>(grid + (size_t)gy * pitch + gx));
        *reinterpret_cast<int4*>(tile + r * kBoxW + 4 * c4) = val;
      }
      __syncthreads();
      int j = 0;
      for (; j + 2 <= n_safe; j += 2) {
        const int2 b2 = *reinterpret_cast<const int2*>(bases + j);
        const int* qa = q_tile + b2.x;
        const int* qb = q_tile + b2.y;
        int va[NR], vbv[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) { va[k] = qa[k * 2 * kBoxW]; vbv[k] = qb[k * 2 * kBoxW]; }
#pragma unroll
        for (int k = 0; k < NR; ++k) a32[k] += (unsigned int)va[k] + (unsigned int)vbv[k];
      }
      if (j < n_safe) {
        const int* qa = q_tile + bases[j];
#pragma unroll
        for (int k = 0; k < NR; ++k) a32[k] += (unsigned int)qa[k * 2 * kBoxW];
      }
    } else {
      __syncwarp();
      for (int j = 0; j < n_safe; ++j) {
        const int* qa = q_grid + bases[j];
#pragma unroll
        for (int k = 0; k < NR; ++k) a32[k] += (unsigned int)__ldg(qa + k * 2 * pitch);
      }
    }
    // exact indices for the beams whose frac(t0) is too close to a cell boundary, or whose patch touches the grid border
    for (unsigned int um = slow_mask; um; um &= um - 1u) {
      const int j = __ffs(um) - 1;
      const double lxj = __shfl_sync(0xffffffffu, lx, j), lyj = __shfl_sync(0xffffffffu, ly, j);
      int gx = cell_index(lxj, xc);
      if (gx < 0 || gx >= size_x) { if (ixl < n_xy) err |= kErrWindow; gx = max(0, min(gx, size_x - 1)); }
#pragma unroll
      for (int k = 0; k < NR; ++k) {
        const int iy = 2 * k + hp;
        int gy = cell_index(lyj, dadd(J.sy, dmul((double)iy, J.f)));
        if (gy < 0 || gy >= size_y) { if (iy < n_xy) err |= kErrWindow; gy = max(0, min(gy, size_y - 1)); }
        a32[k] += (unsigned int)__ldg(grid + (size_t)gy * pitch + gx);
      }
    }
    // two beams are added before the spill: <= 32 cells of <= 2^25 per chunk still fit 32 bits (2^30)
#pragma unroll
    for (int k = 0; k < NR; ++k) { a64[k] += a32[k]; a32[k] = 0u; }
    __syncthreads();     // the tile, s_box and sBase are reused by the next chunk
  }

  // epilogue: response = sum / divisor (:659), centre penalty (:727-743), store, block maximum
  unsigned long long kmax = 0ull;
  if (active && ixl < n_xy) {
    const double dx = dsub(xc, J.cx);
    const double dx2 = dmul(dx, dx);
    const double da = dsub(ang, J.ca);
    const double a2 = dmul(da, da);
    const double ap = fmax(dsub(1.0, ddiv(dmul(0.25, a2), 0.349)), 0.9);
    double* out = J.score + ((long long)ia_local * n_xy + ixl) * n_xy;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int iy = 2 * k + hp;
      if (iy < n_xy) {
        double sc = ddiv(dmul((double)a64[k], kFixScale), J.divisor);
        if (J.use_penalty) {
          const bool zero = sc < 0.0 ? (sc >= -1e-06) : (sc <= 1e-06);     // :728
          if (!zero) {
            const double dy = dsub(dadd(J.sy, dmul((double)iy, J.f)), J.cy);
            double d2 = dadd(dx2, dmul(dy, dy));
            d2 = dmul(d2, J.m2);
            const double dp = fmax(dsub(1.0, ddiv(dmul(J.gain, d2), J.half_size)), 0.5);
            sc = dmul(sc, dmul(dp, ap));
          }
        }
        out[iy] = sc;
        const unsigned long long key = score_key(sc);
        kmax = key > kmax ? key : kmax;
      }
    }
  }
  kmax = warp_max_u64(kmax);
  if (lane == 0) s_wmax[warp] = kmax;
  if (err) atomicOr(J.err, err);
  __syncthreads();
  if (tid == 0) {
    unsigned long long m = 0ull;
    for (int w = 0; w < kAngles; ++w) m = s_wmax[w] > m ? s_wmax[w] : m;
    atomicMax(J.best_key, m);
  }
}
}  // namespace patch

int score_patch_angles() { return patch::kAngles; }
#ifdef RSM_STAGED_DEBUG
extern "C" void rsm_debug_patch(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, patch::g_pdbg, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(patch::g_pdbg, z, sizeof z); }
}
#endif

cudaError_t launch_score_patch(int n_xy, int n_cta, cudaStream_t st, const ScoreJob* jobs, const int* cta_begin, int n_jobs) {
  if (n_cta <= 0) return cudaSuccess;
  if (n_xy <= 8) patch::score_patch_kernel<4><<<n_cta, patch::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
  else if (n_xy <= 14) patch::score_patch_kernel<7><<<n_cta, patch::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
  else patch::score_patch_kernel<8><<<n_cta, patch::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
  return cudaGetLastError();
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


// =================================================================================================
// flat variant: small windows with a non-integer search step (the fine / super-fine passes)
// =================================================================================================
// A fine pass is 11 x 11 translations x 11 angles, a super-fine pass 3 x 3 x 21: tiles of such
// windows leave most lanes of the tiled kernel idle, and their index tables are as large as the
// work itself.  B200 issues only ~17 FP64 operations per clock per SM (measured), so computing
// (int)(lut + x + 0.5) in FP64 per evaluation is out of the question too.  Instead:
//   * a CTA of 256 threads takes 256 consecutive candidates of the job in (angle, y, x) order
//     (spanning up to 30 angles); every thread owns ONE candidate;
//   * coordinates are carried as 32-bit fixed point with 16 fractional bits: lutq = rn(lut * 2^16)
//     per (angle, beam) (built per 32-beam chunk from the FP64 endpoint), xq = rn(x * 2^16) + 2^15
//     per thread.  t = lutq + xq is an integer add; it differs from the reference's FP64 value
//     T = (lut + x) + 0.5 by at most 2^-16 + 1e-11 cells, so whenever frac(t) lies in
//     [2, 65533] / 65536 and t >= 0, trunc(T) is PROVABLY t >> 16;
//   * the ~1.2e-4 of evaluations that fail the test (or leave the grid) recompute the index with
//     the exact FP64 expression.
namespace flat {

constexpr int kThreads = 256;
constexpr int kPC = 32;          // beams per chunk
constexpr int kMaxAngles = 32;   // angles a CTA may span
constexpr double kQ = 65536.0;

__device__ __forceinline__ int to_q(double v) {   // rn(v * 2^16), saturated so that sums cannot overflow
  return __double2int_rn(fmin(fmax(dmul(v, kQ), -1073741824.0), 1073741824.0));
}

// K candidates per thread: a CTA takes K * 256 consecutive candidates, thread t the candidates k0 + t + j * 256.
// Larger windows (the 11 x 11 x 11 fine pass) amortise the chunk tables over more gathers and keep K independent
// loads in flight per beam.
//
// Per chunk of 32 beams:
//   1. table: lutq = rn(rotated endpoint * 2^16) per (angle of the CTA, beam), a SAFE bit per entry -- set when, for
//      every translation of the window, frac(lutq + xq) stays 2 / 65536 away from a cell boundary (so trunc is exact for
//      all candidates of that angle and beam: no per-candidate test in the gather loop) -- and the hull of the cells
//      the safe entries can touch;
//   2. if the hull fits the shared-memory tile (104 x 96 cells), it is copied there with 16-byte loads; every copied
//      cell then serves up to K * 256 gathers, and a warp's 32 gathers (2-3 grid rows x 5-6 columns) cost one
//      shared-memory wavefront instead of 2-3 L1 lines;
//   3. gather: 2 adds, 2 shifts, 1 multiply-add and 1 load per evaluation; entries without the SAFE bit recompute the
//      index with the exact FP64 expression (~1e-3 of the beams), chunks whose hull does not fit gather from global memory.
constexpr int kTileW = 104, kTileH = 96;     // pitch = 8 (mod 32): the 2-3 rows a warp's gathers touch fall into disjoint banks

template <bool FIXED, int K>
__global__ void __launch_bounds__(kThreads, K <= 2 ? 4 : 3)
score_flat_kernel(const ScoreJob* __restrict__ jobs, const int* __restrict__ cta_begin, int n_jobs) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ ScoreJob J;
  __shared__ int s_job;
  __shared__ unsigned long long s_wmax[kThreads / 32];
  __shared__ int2 sLutQ[kMaxAngles][kPC];
  __shared__ unsigned int sSafe[kMaxAngles];          // bit b: beam b of the chunk is safe for every candidate of the angle
  __shared__ int sXq[16], sYq[16];                    // rn(x_i * 2^16) + 2^15 per translation index
  __shared__ int sHull[kThreads / 32][4];
  __shared__ __align__(16) int tile[kTileW * kTileH];

  if (tid == 0) s_job = find_job(cta_begin, n_jobs, blockIdx.x);
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(jobs + s_job);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < int(sizeof(ScoreJob) / 4); i += kThreads) dst[i] = __ldg(src + i);
  }
  const int first_cta = __ldg(cta_begin + s_job);
  __syncthreads();

  const int n_xy = J.n_xy, plane = n_xy * n_xy;
  const int n_local = J.ang_count * plane;
  const int S = (FIXED && J.n_split > 1) ? J.n_split : 1;      // CTAs sharing the beams of one block of candidates
  const int group = (blockIdx.x - first_cta) / S, split = (blockIdx.x - first_cta) - group * S;
  const int k0 = group * (kThreads * K);                       // first candidate of this CTA, (angle, y, x) order
  const int a_first = k0 / plane;
  const int a_last = min(k0 + kThreads * K - 1, n_local - 1) / plane;
  const int n_a = a_last - a_first + 1;
  if (tid < n_xy) {
    sXq[tid] = to_q(dadd(J.sx, dmul((double)tid, J.f))) + 32768;     // :569
    sYq[tid] = to_q(dadd(J.sy, dmul((double)tid, J.f))) + 32768;     // :572
  }
  __syncthreads();
  int kk[K], a_rel[K], xq[K], yq[K], cix[K], ciy[K];
  bool live[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    kk[j] = k0 + tid + j * kThreads;
    live[j] = kk[j] < n_local;
    const int kc = live[j] ? kk[j] : n_local - 1;
    const int ia_l = kc / plane, rem = kc - ia_l * plane;
    ciy[j] = rem / n_xy; cix[j] = rem - ciy[j] * n_xy;
    a_rel[j] = ia_l - a_first;
    xq[j] = sXq[cix[j]];
    yq[j] = sYq[ciy[j]];
  }
  const int xq_lo = sXq[0], xq_hi = sXq[n_xy - 1], yq_lo = sYq[0], yq_hi = sYq[n_xy - 1];
  const int V = J.V, pitch = J.pitch;
  const int size_x = J.size_x, size_y = J.size_y;
  const int nchunks_all = (V + kPC - 1) / kPC;
  const int per_split = (nchunks_all + S - 1) / S;
  const int c_begin = min(split * per_split, nchunks_all);
  const int nchunks = min(nchunks_all, c_begin + per_split);   // this CTA gathers chunks [c_begin, nchunks)
  const int* gridI = reinterpret_cast<const int*>(J.grid);
  const float* gridF = reinterpret_cast<const float*>(J.grid);

  // exact FP64 index of (candidate j, beam v)
  auto exact_index = [&](int j, int v, int* err) -> int {
    const int ia = J.ang_begin + a_first + a_rel[j];
    const double cs = __ldg(J.trig + 3 * ia), sn = __ldg(J.trig + 3 * ia + 1);
    const int p = v * J.step;
    const double px = __ldg(J.pts + 2 * p), py = __ldg(J.pts + 2 * p + 1);
    int gx = cell_index(dsub(dmul(cs, px), dmul(sn, py)), dadd(J.sx, dmul((double)cix[j], J.f)));      // :647-648
    int gy = cell_index(dadd(dmul(sn, px), dmul(cs, py)), dadd(J.sy, dmul((double)ciy[j], J.f)));
    if ((unsigned int)gx >= (unsigned int)size_x || (unsigned int)gy >= (unsigned int)size_y) {
      if (live[j]) *err = kErrWindow;
      gx = max(0, min(gx, size_x - 1));
      gy = max(0, min(gy, size_y - 1));
    }
    return gy * pitch + gx;
  };

  unsigned int a32[K];
  unsigned long long a64[K];
  double ad[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { a32[j] = 0u; a64[j] = 0ull; ad[j] = 0.0; }
  int err = 0;
  for (int c = c_begin; c < nchunks; ++c) {
    const int npc = min(kPC, V - c * kPC);
    // ---- 1. table, SAFE bits, hull: warp w takes the angles w, w + 8, ...; lane = beam of the chunk ----
    int hx0 = 0x7fffffff, hx1 = -1, hy0 = 0x7fffffff, hy1 = -1;
    for (int a = warp; a < n_a; a += kThreads / 32) {
      bool safe = false;
      if (lane < npc) {
        const int ja = J.ang_begin + a_first + a;
        const double cs = __ldg(J.trig + 3 * ja), sn = __ldg(J.trig + 3 * ja + 1);
        const int p = (c * kPC + lane) * J.step;
        const double px = __ldg(J.pts + 2 * p), py = __ldg(J.pts + 2 * p + 1);
        const int qx = to_q(dsub(dmul(cs, px), dmul(sn, py))), qy = to_q(dadd(dmul(sn, px), dmul(cs, py)));   // :179-180
        sLutQ[a][lane] = make_int2(qx, qy);
        safe = true;
        for (int i = 0; i < n_xy; ++i) {
          const unsigned int fx = (unsigned int)(qx + sXq[i] - 2) & 0xffffu, fy = (unsigned int)(qy + sYq[i] - 2) & 0xffffu;
          if (fx > 65531u || fy > 65531u) safe = false;
        }
        // cells the window's candidates touch for this beam (xq ascending with the translation index)
        const int gx_lo = (qx + xq_lo) >> 16, gx_hi = (qx + xq_hi) >> 16, gy_lo = (qy + yq_lo) >> 16, gy_hi = (qy + yq_hi) >> 16;
        if (gx_lo < 0 || gx_hi >= size_x || gy_lo < 0 || gy_hi >= size_y) safe = false;     // off the grid: exact path reports it
        if (safe) { hx0 = min(hx0, gx_lo); hx1 = max(hx1, gx_hi); hy0 = min(hy0, gy_lo); hy1 = max(hy1, gy_hi); }
      }
      const unsigned int m = __ballot_sync(0xffffffffu, safe);
      if (lane == 0) sSafe[a] = m;
    }
    hx0 = __reduce_min_sync(0xffffffffu, hx0); hx1 = __reduce_max_sync(0xffffffffu, hx1);
    hy0 = __reduce_min_sync(0xffffffffu, hy0); hy1 = __reduce_max_sync(0xffffffffu, hy1);
    if (lane == 0) { sHull[warp][0] = hx0; sHull[warp][1] = hx1; sHull[warp][2] = hy0; sHull[warp][3] = hy1; }
    __syncthreads();
    int bx0 = 0x7fffffff, bx1 = -1, by0 = 0x7fffffff, by1 = -1;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      bx0 = min(bx0, sHull[w][0]); bx1 = max(bx1, sHull[w][1]); by0 = min(by0, sHull[w][2]); by1 = max(by1, sHull[w][3]);
    }
    const int xl = bx0 & ~3;                                        // 16-byte aligned rows
    const bool fits = bx1 >= 0 && (bx1 - xl + 1 <= kTileW) && (by1 - by0 + 1 <= kTileH);
    // ---- 2. the hull into shared memory ----
    if (fits) {
      const int w4 = (bx1 - xl + 1 + 3) >> 2, hh = by1 - by0 + 1;
      for (int q = tid; q < w4 * hh; q += kThreads) {
        const int r = q / w4, c4 = q - r * w4;
        const int gx = xl + 4 * c4;
        int4 val = make_int4(0, 0, 0, 0);
        if (gx + 3 < pitch) val = __ldg(reinterpret_cast<const int4*>(gridI + (size_t)(by0 + r) * pitch + gx));
        *reinterpret_cast<int4*>(tile + r * kTileW + 4 * c4) = val;
      }
      __syncthreads();
    }
    // ---- 3. gather ----
    // Four specialised loops: cells from the tile or from global memory; with or without the per-entry SAFE test.
    // A warp takes the test-free loop unless one of its lanes' angles has an unsafe beam in this chunk (~10 % of the
    // warp-chunks), so the common loop is 2 adds, 2 shifts, a multiply-add and a load per evaluation.
    const unsigned int valid = npc == kPC ? 0xffffffffu : ((1u << npc) - 1u);
    const int2* l[K];
    unsigned int slow[K];
    bool any_slow = false;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      l[j] = sLutQ[a_rel[j]];
      slow[j] = ~sSafe[a_rel[j]] & valid;
      any_slow = any_slow || slow[j] != 0u;
    }
    any_slow = __any_sync(0xffffffffu, any_slow);
    auto gather = [&](auto in_tile, auto checked) {
      constexpr bool TILE = decltype(in_tile)::value, CHECK = decltype(checked)::value;
      const int org = TILE ? by0 * kTileW + xl : 0;                 // tile index = gy * kTileW + gx - org
      const int row = TILE ? kTileW : pitch;
#pragma unroll 4
      for (int pc = 0; pc < npc; ++pc) {
        int vv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int2 e = l[j][pc];
          const int at = ((e.y + yq[j]) >> 16) * row + ((e.x + xq[j]) >> 16) - org;
          if (CHECK && ((slow[j] >> pc) & 1u)) vv[j] = __ldg(gridI + exact_index(j, c * kPC + pc, &err));
          else vv[j] = TILE ? tile[at] : __ldg(gridI + at);
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
          if (FIXED) a32[j] += (unsigned int)vv[j];
          else ad[j] = dadd(ad[j], (double)__int_as_float(vv[j]));
        }
      }
    };
    if (fits) {
      if (any_slow) gather(std::true_type{}, std::true_type{}); else gather(std::true_type{}, std::false_type{});
    } else {
      if (any_slow) gather(std::false_type{}, std::true_type{}); else gather(std::false_type{}, std::false_type{});
    }
    if (FIXED) {
#pragma unroll
      for (int j = 0; j < K; ++j) { a64[j] += a32[j]; a32[j] = 0u; }
    }
    __syncthreads();      // the table, the SAFE bits and the tile are rewritten by the next chunk
  }
  (void)gridF;

  if (FIXED && S > 1) {
    __shared__ int s_ticket;
#pragma unroll
    for (int j = 0; j < K; ++j) if (live[j] && a64[j]) atomicAdd(J.acc + kk[j], a64[j]);
    if (err) atomicOr(J.err, err);
    err = 0;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(J.tickets + group, 1);
    __syncthreads();
    if (s_ticket != S - 1) return;
    __threadfence();
#pragma unroll
    for (int j = 0; j < K; ++j) if (live[j]) a64[j] = __ldcg(J.acc + kk[j]);
  }
  unsigned long long key = 0ull;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (!live[j]) continue;
    const int ia_l = a_first + a_rel[j], ia = J.ang_begin + ia_l;
    const double x = dadd(J.sx, dmul((double)cix[j], J.f)), y = dadd(J.sy, dmul((double)ciy[j], J.f));
    double sc = ddiv(FIXED ? dmul((double)a64[j], kFixScale) : ad[j], J.divisor);    // :659
    if (J.use_penalty) {
      const bool zero = sc < 0.0 ? (sc >= -1e-06) : (sc <= 1e-06);             // :728
      if (!zero) {
        const double ang = __ldg(J.trig + 3 * ia + 2);
        const double dx = dsub(x, J.cx), dy = dsub(y, J.cy);
        double d2 = dadd(dmul(dx, dx), dmul(dy, dy));
        d2 = dmul(d2, J.m2);
        const double dp = fmax(dsub(1.0, ddiv(dmul(J.gain, d2), J.half_size)), 0.5);
        const double da = dsub(ang, J.ca);
        const double ap = fmax(dsub(1.0, ddiv(dmul(0.25, dmul(da, da)), 0.349)), 0.9);
        sc = dmul(sc, dmul(dp, ap));
      }
    }
    J.score[((long long)ia_l * n_xy + cix[j]) * n_xy + ciy[j]] = sc;
    const unsigned long long kj = score_key(sc);
    key = kj > key ? kj : key;
  }
  key = warp_max_u64(key);
  if ((tid & 31) == 0) s_wmax[tid >> 5] = key;
  if (err) atomicOr(J.err, err);
  __syncthreads();
  if (tid == 0) {
    unsigned long long m = 0ull;
    for (int w = 0; w < kThreads / 32; ++w) m = s_wmax[w] > m ? s_wmax[w] : m;
    atomicMax(J.best_key, m);
  }
}

}  // namespace flat

// widest window of the flat variant, and candidates per thread for a window width: a CTA of k * 256 consecutive
// candidates must not span more than kMaxAngles angles, and small jobs should not be mostly padding
int score_flat_max_nxy() { return 12; }
int score_flat_k(int n_xy) {
  const int plane = n_xy * n_xy;
  if (plane < 36) return 1;      // 3 x 3 .. 5 x 5: up to 30 angles per 256 candidates already
  return plane < 64 ? 2 : 3;     // 121 x 11 = 1331 candidates -> two CTAs of 768
}
int score_flat_ctas(int n_local, int k) { return (n_local + flat::kThreads * k - 1) / (flat::kThreads * k); }

cudaError_t launch_score_flat(bool fixed, int k, int n_cta, cudaStream_t st, const ScoreJob* jobs, const int* cta_begin, int n_jobs) {
  if (n_cta <= 0) return cudaSuccess;
  if (fixed) {
    if (k == 1) flat::score_flat_kernel<true, 1><<<n_cta, flat::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
    else if (k == 2) flat::score_flat_kernel<true, 2><<<n_cta, flat::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
    else flat::score_flat_kernel<true, 3><<<n_cta, flat::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
  } else {
    if (k == 1) flat::score_flat_kernel<false, 1><<<n_cta, flat::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
    else if (k == 2) flat::score_flat_kernel<false, 2><<<n_cta, flat::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
    else flat::score_flat_kernel<false, 3><<<n_cta, flat::kThreads, 0, st>>>(jobs, cta_begin, n_jobs);
  }
  return cudaGetLastError();
}

// =================================================================================================
// staged variant: the grid window of a beam round lives in shared memory, filled by TMA bulk copies
// =================================================================================================
// For windows of >= ~48 translations per axis on a fixed-point grid with a unit search step.
// The L1 path above is bound by L1 load issue: a warp's 32 consecutive cells start at an arbitrary
// cell, so every gather touches two 128-byte lines (~1.8 cycles per warp load).  Shared memory has
// no such penalty (32 consecutive words are conflict-free at any offset: 1 cycle).  Staging only
// pays if each staged cell is gathered several times, which needs BIG candidate tiles:
//   * a CTA owns one angle, one 96 x 96 tile of translations (18 candidates per thread in
//     registers) and one contiguous slice of the beams; 16 compute warps + 1 producer warp;
//   * the producer walks the slice in ROUNDS of up to 32 consecutive beams whose tile-origin
//     boxes fit a 20 480-cell buffer (prefix min/max with warp shuffles), publishes the round's
//     tile-relative bases, arms the buffer's FULL mbarrier with the byte count and issues one
//     cp.async.bulk (TMA, SASS UBLKCP) per tile row; it runs one round ahead (two buffers) and
//     re-uses a buffer when the 16 compute warps have arrived on its EMPTY mbarrier;
//   * compute warps wait on FULL, gather with per-row strides and immediate column offsets
//     (1 LDS + 0.5 IADD3 per evaluation), arrive on EMPTY -- no block-wide barrier in the loop;
//   * beams are split over CTAs to fill the machine when a job has few (angle, tile) pairs;
//     partial sums are integers, so they are combined with 64-bit atomic adds and the last CTA of
//     an (angle, tile) (atomic ticket) finalises the scores.
// Per-beam tile origins use the same provable affine index test as the L1 path (tile-wide);
// beams that fail it or whose tile leaves the grid are added afterwards with exact per-thread
// indices read from global memory -- integer sums are order-independent.
namespace staged {

constexpr int kComputeWarps = 16;
constexpr int kThreads = (kComputeWarps + 1) * 32;   // + 1 producer warp
constexpr int kBufCells = 20480;            // 80 KB per buffer
constexpr int kBoxW0 = 160, kBoxH0 = 128;   // the two TMA box shapes (both kBufCells cells): wide and tall
constexpr int kBoxW1 = 128, kBoxH1 = 160;
constexpr int kRound = 32;                  // beams per round at most
#ifdef RSM_STAGED_DEBUG
__device__ unsigned long long g_dbg[32];
#define DBG_T() clock64()
#define DBG_ADD(i, v) atomicAdd(&g_dbg[i], (unsigned long long)(v))
#else
#define DBG_T() 0ll
#define DBG_ADD(i, v) do {} while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// Polling wait with a sleep between polls: a waiting warp must not take issue slots from the compute
// warps of its scheduler (measured: a producer spinning on try_wait slowed its four neighbours by 15%).
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity, unsigned int sleep_ns) {
  const uint32_t a = smem_u32(b);
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(sleep_ns);
  }
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned long long ld_dsmem_u64(const void* local, int cta_rank) {
  uint32_t remote;
  unsigned long long v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(cta_rank));
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(remote) : "memory");
  return v;
}
// One box of the grid (a CUtensorMap built by the host, held in global memory) into shared memory.
__device__ __forceinline__ void tma_box(void* dst, const void* tmap, int x, int y, uint64_t* b) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(b))
      : "memory");
}
__device__ __forceinline__ void tmap_acquire(const void* tmap) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}

// One round of one compute warp: every beam of the round adds its NX x RY cells to the accumulators.
// FULL: all RY rows of the warp are inside the window (no row predicate).
template <int RX, int RY, int NX, bool FULL>
__device__ __forceinline__ void accumulate_round(unsigned int (&lo)[RX * RY], const int* tile_buf, const int* bases, int n,
                                                 int sw, int rows) {
#pragma unroll 2
  for (int j = 0; j < n; ++j) {
    const int* q = tile_buf + bases[j];
    int v[RY][NX];
#pragma unroll
    for (int ry = 0; ry < RY; ++ry) {
      const int* qr = q + ry * sw;
#pragma unroll
      for (int rx = 0; rx < NX; ++rx) v[ry][rx] = (FULL || ry < rows) ? qr[rx * 32] : 0;
    }
#pragma unroll
    for (int ry = 0; ry < RY; ++ry)
#pragma unroll
      for (int rx = 0; rx < NX; ++rx) lo[ry * RX + rx] += (unsigned int)v[ry][rx];
  }
}

template <int RX, int RY>
__global__ void __launch_bounds__(kThreads, 1)
score_staged_kernel(const ScoreJob* __restrict__ jobs, const int* __restrict__ cta_begin, int n_jobs) {
  constexpr int TILE_X = 32 * RX, TILE_Y = kComputeWarps * RY, NC = RX * RY;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp == kComputeWarps;
  const long long k0 = DBG_T(); (void)k0;

  __shared__ ScoreJob J;
  __shared__ int s_job;
  __shared__ unsigned long long s_wmax[kComputeWarps];
  __shared__ double sX[TILE_X], sY[TILE_Y];
  __shared__ int sBase[2][kRound];          // tile-relative cell offset of each beam of the round, -1 = skip
  __shared__ int sMeta[2][4];               // beams in the round, row pitch, last-round flag
  __shared__ __align__(8) uint64_t sFull[2], sEmpty[2];
  __shared__ int sUnsafeCount;
  __shared__ int sUnsafe[64];               // beams that need exact per-thread indices
  extern __shared__ __align__(128) unsigned char dyn[];
  int* buf0 = reinterpret_cast<int*>(dyn);
  int* buf1 = buf0 + kBufCells;
  int2* sBeam = reinterpret_cast<int2*>(buf1 + kBufCells);   // tile-origin cell of every beam of the slice

  if (tid == 0) s_job = find_job(cta_begin, n_jobs, blockIdx.x);
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(jobs + s_job);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < int(sizeof(ScoreJob) / 4); i += kThreads) dst[i] = __ldg(src + i);
  }
  const int first_cta = __ldg(cta_begin + s_job);
  if (tid == 0) {
    mbar_init(&sFull[0], 1); mbar_init(&sFull[1], 1);
    mbar_init(&sEmpty[0], kComputeWarps); mbar_init(&sEmpty[1], kComputeWarps);
    sUnsafeCount = 0;
  }
  __syncthreads();

  const int S = J.n_split;
  const int local = blockIdx.x - first_cta;
  const int tiles = J.tiles_x * J.tiles_y;
  const int per_angle = tiles * S;
  const int ia_local = local / per_angle;
  const int rem = local - ia_local * per_angle;
  const int tile = rem / S, split = rem - tile * S;
  const int tx0 = (tile % J.tiles_x) * TILE_X;
  const int ty0 = (tile / J.tiles_x) * TILE_Y;
  const int ia = J.ang_begin + ia_local;
  const double cs = __ldg(J.trig + 3 * ia), sn = __ldg(J.trig + 3 * ia + 1), ang = __ldg(J.trig + 3 * ia + 2);
  const int V = J.V, n_xy = J.n_xy, pitch = J.pitch, size_x = J.size_x, size_y = J.size_y;
  const int vb = (V + S - 1) / S;
  const int v_begin = min(V, split * vb), v_end = min(V, v_begin + vb);
  const int nv = v_end - v_begin;
  const int* __restrict__ grid = reinterpret_cast<const int*>(J.grid);
  // the part of the tile that is inside the window: whole 32-lane groups in x, single rows in y
  const int nx_act = min(RX, (n_xy - tx0 + 31) >> 5);
  // footprint of a beam in the box: the window's cells only.  Lanes beyond the window in the last 32-lane group read
  // whatever follows (the next box row, at worst the first cells after the buffer: still this CTA's shared memory);
  // their sums are never stored.  81 instead of 96 cells leaves room for ~20 % more beams per box.
  const int ext_x = min(32 * nx_act, n_xy - tx0), ext_y = min(TILE_Y, n_xy - ty0);
  const int rows = max(0, min(RY, n_xy - (ty0 + warp * RY)));      // rows of this warp inside the window

  for (int i = tid; i < TILE_X + TILE_Y; i += kThreads) {
    if (i < TILE_X) sX[i] = dadd(J.sx, dmul((double)(tx0 + i), J.f));
    else sY[i - TILE_X] = dadd(J.sy, dmul((double)(ty0 + i - TILE_X), J.f));
  }
  __syncthreads();
  // tile-origin cell of every beam of the slice + the affine safety test for the whole tile
  for (int j = tid; j < nv; j += kThreads) {
    const int p = (v_begin + j) * J.step;
    const double px = __ldg(J.pts + 2 * p), py = __ldg(J.pts + 2 * p + 1);
    const double lx = dsub(dmul(cs, px), dmul(sn, py));
    const double ly = dadd(dmul(sn, px), dmul(cs, py));
    const double tx_ = dadd(dadd(lx, sX[0]), 0.5), ty_ = dadd(dadd(ly, sY[0]), 0.5);
    const int gx0 = __double2int_rz(tx_), gy0 = __double2int_rz(ty_);
    const double fx = tx_ - (double)gx0, fy = ty_ - (double)gy0;
    const bool ok = fx > 1e-6 && fx < 1.0 - 1e-6 && fy > 1e-6 && fy < 1.0 - 1e-6 &&
                    gx0 >= 0 && gx0 + ext_x - 1 < size_x && gy0 >= 0 && gy0 + ext_y - 1 < size_y;
    sBeam[j] = ok ? make_int2(gx0, gy0) : make_int2(-1, -1);
    if (!ok) { const int pos = atomicAdd(&sUnsafeCount, 1); if (pos < 64) sUnsafe[pos] = j; }
  }
  __syncthreads();

  // Accumulators: 32-bit sums kept below 2^31 between rounds (a round adds at most 32 * 2^25 = 2^30),
  // the overflow counted in units of 2^31, four 8-bit counters per register (total < 2^36 for
  // V <= 2048, so a counter stays below 32).  64-bit accumulators here would cost 36 registers and
  // leave ptxas no room to keep the 18 shared-memory loads of a beam in flight.
  unsigned int lo[NC];
  unsigned int hi[(NC + 3) / 4];
#pragma unroll
  for (int c = 0; c < NC; ++c) lo[c] = 0u;
#pragma unroll
  for (int c = 0; c < (NC + 3) / 4; ++c) hi[c] = 0u;

  const long long k1 = DBG_T(); (void)k1;
  if (nv > 0) {
    if (producer) {
      // ---- producer warp: plan rounds, publish bases, arm FULL, issue one TMA box copy per round ----
      const char* tmaps = reinterpret_cast<const char*>(J.tmap);
      if (lane == 0) { tmap_acquire(tmaps); tmap_acquire(tmaps + 128); }
      __syncwarp();
      int b = 0;
      for (int r = 0; b < nv; ++r) {
        const int which = r & 1;
        const long long t0 = DBG_T();
        const long long t1 = t0;
        const int j = b + lane;
        int2 e = make_int2(-1, -1);
        if (j < nv) e = sBeam[j];
        const bool live = j < nv, safe = live && e.x >= 0;
        int xmin = safe ? e.x : 0x7fffffff, xmax = safe ? e.x : -1, ymin = safe ? e.y : 0x7fffffff, ymax = safe ? e.y : -1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {      // inclusive prefix min / max over lanes 0..lane
          const int a0 = __shfl_up_sync(0xffffffffu, xmin, o), a1 = __shfl_up_sync(0xffffffffu, xmax, o);
          const int a2 = __shfl_up_sync(0xffffffffu, ymin, o), a3 = __shfl_up_sync(0xffffffffu, ymax, o);
          if (lane >= o) { xmin = min(xmin, a0); xmax = max(xmax, a1); ymin = min(ymin, a2); ymax = max(ymax, a3); }
        }
        // the longest run of beams whose tile footprints share one box, wide or tall
        const bool any = xmax >= 0;
        const int xl = xmin & ~3;   // the inner box coordinate must be a multiple of 16 bytes
        const int dx = xmax - xl + ext_x, dy = ymax - ymin + ext_y;
        const bool fit0 = live && (!any || (dx <= kBoxW0 && dy <= kBoxH0));
        const bool fit1 = live && (!any || (dx <= kBoxW1 && dy <= kBoxH1));
        const unsigned int vote0 = __ballot_sync(0xffffffffu, fit0), vote1 = __ballot_sync(0xffffffffu, fit1);
        const int n0 = (vote0 == 0xffffffffu) ? 32 : (__ffs(~vote0) - 1);   // leading run of ones (>= 1: one beam always fits)
        const int n1 = (vote1 == 0xffffffffu) ? 32 : (__ffs(~vote1) - 1);
        const bool tall = n1 > n0;
        const int n = tall ? n1 : n0, rw = tall ? kBoxW1 : kBoxW0;
        const int srcl = n - 1;
        const bool rany = __shfl_sync(0xffffffffu, (int)any, srcl) != 0;
        int rxl = __shfl_sync(0xffffffffu, xl, srcl), ryl = __shfl_sync(0xffffffffu, ymin, srcl);
        if (!rany) { rxl = 0; ryl = 0; }
        // the round is planned in registers while the compute warps still read this buffer's bases; only
        // publishing them and the copy itself wait for the buffer
        if (r >= 2) mbar_wait(&sEmpty[which], (uint32_t)(((r >> 1) - 1) & 1), 100);
        if (lane < n) sBase[which][lane] = safe ? (e.y - ryl) * rw + (e.x - rxl) : -1;
        const unsigned int unsafe = __ballot_sync(0xffffffffu, lane < n && !safe);
        if (lane == 0) { sMeta[which][0] = n; sMeta[which][1] = rw; sMeta[which][2] = (b + n >= nv) ? 1 : 0; sMeta[which][3] = unsafe == 0u; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        const long long t2 = DBG_T();
        if (lane == 0) {
          mbar_expect_tx(&sFull[which], (uint32_t)(kBufCells * 4));
          tma_box(which ? buf1 : buf0, tmaps + (tall ? 128 : 0), rxl, ryl, &sFull[which]);
        }
        b += n;
#ifdef RSM_STAGED_DEBUG
        const long long t3 = DBG_T();
        if (lane == 0) { DBG_ADD(0, 1); DBG_ADD(1, t1 - t0); DBG_ADD(2, t2 - t1); DBG_ADD(3, t3 - t2); DBG_ADD(5, n); DBG_ADD(6, tall); }
#endif
      }
    } else {
      // ---- compute warps -----------------------------------------------------------------------------
      for (int r = 0;; ++r) {
        const int which = r & 1;
        const long long c0 = DBG_T();
        mbar_wait(&sFull[which], (uint32_t)((r >> 1) & 1), 40);
        const long long c1 = DBG_T();
        const int n = sMeta[which][0], sw = sMeta[which][1], last = sMeta[which][2], all_safe = sMeta[which][3];
        const int* tile_buf = (which ? buf1 : buf0) + (warp * RY) * sw + lane;
        const int* bases = sBase[which];
        if (rows > 0) {
          if (all_safe && rows == RY) {
            if (nx_act == RX) accumulate_round<RX, RY, RX, true>(lo, tile_buf, bases, n, sw, rows);
            else if (RX > 2 && nx_act == 2) accumulate_round<RX, RY, (RX > 2 ? 2 : 1), true>(lo, tile_buf, bases, n, sw, rows);
            else accumulate_round<RX, RY, 1, true>(lo, tile_buf, bases, n, sw, rows);
          } else if (all_safe) {
            // the warp that straddles the window edge in y: same loop, loads predicated per row
            if (nx_act == RX) accumulate_round<RX, RY, RX, false>(lo, tile_buf, bases, n, sw, rows);
            else if (RX > 2 && nx_act == 2) accumulate_round<RX, RY, (RX > 2 ? 2 : 1), false>(lo, tile_buf, bases, n, sw, rows);
            else accumulate_round<RX, RY, 1, false>(lo, tile_buf, bases, n, sw, rows);
          } else {
            // a round with a beam that failed the tile-wide index test (rare)
            for (int j = 0; j < n; ++j) {
              const int base = bases[j];
              if (base < 0) continue;
              const int* q = tile_buf + base;
#pragma unroll
              for (int ry = 0; ry < RY; ++ry) {
                if (ry >= rows) continue;
                const int* qr = q + ry * sw;
#pragma unroll
                for (int rx = 0; rx < RX; ++rx)
                  if (rx < nx_act) lo[ry * RX + rx] += (unsigned int)qr[rx * 32];
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) { hi[c >> 2] += (lo[c] >> 31) << (8 * (c & 3)); lo[c] &= 0x7fffffffu; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sEmpty[which]);
#ifdef RSM_STAGED_DEBUG
        if (lane == 0 && warp == 0) { DBG_ADD(8, c1 - c0); DBG_ADD(9, DBG_T() - c1); }
        if (lane == 0 && warp == 15) { DBG_ADD(10, c1 - c0); DBG_ADD(11, DBG_T() - c1); }
#endif
        if (last) break;
      }
    }
  }
  __syncthreads();
  const long long k2 = DBG_T(); (void)k2;

  unsigned long long a64[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) a64[c] = ((unsigned long long)((hi[c >> 2] >> (8 * (c & 3))) & 0xffu) << 31) + lo[c];

  int err = 0;
  if (!producer) {
    // beams that failed the tile-wide test: exact indices, straight from global memory
    const int n_unsafe = sUnsafeCount;
    const int n_slow = n_unsafe <= 64 ? n_unsafe : nv;      // more than the list holds: scan every beam
    for (int u = 0; u < n_slow; ++u) {
      const int j = n_unsafe <= 64 ? sUnsafe[u] : u;
      if (sBeam[j].x >= 0) continue;
      const int p = (v_begin + j) * J.step;
      const double px = __ldg(J.pts + 2 * p), py = __ldg(J.pts + 2 * p + 1);
      const double lx = dsub(dmul(cs, px), dmul(sn, py));
      const double ly = dadd(dmul(sn, px), dmul(cs, py));
      int gy[RY];
#pragma unroll
      for (int ry = 0; ry < RY; ++ry) {
        int g = cell_index(ly, sY[warp * RY + ry]);
        if (g < 0 || g >= size_y) { if (ty0 + warp * RY + ry < n_xy) err |= kErrWindow; g = max(0, min(g, size_y - 1)); }
        gy[ry] = g * pitch;
      }
#pragma unroll
      for (int rx = 0; rx < RX; ++rx) {
        int g = cell_index(lx, sX[lane + 32 * rx]);
        if (g < 0 || g >= size_x) { if (tx0 + lane + 32 * rx < n_xy) err |= kErrWindow; g = max(0, min(g, size_x - 1)); }
#pragma unroll
        for (int ry = 0; ry < RY; ++ry) a64[ry * RX + rx] += (unsigned int)__ldg(grid + (gy[ry] + g));
      }
    }
  }

  // ---- beam splits: the S CTAs of an (angle, tile) form a thread-block cluster.  Every CTA parks its
  // integer partial sums in its own shared memory; after a cluster barrier CTA k sums accumulator
  // slots c = k, k + S, ... over all peers through distributed shared memory and finishes those.
  const long long k_angle = (long long)ia_local * n_xy * n_xy;
  if (S > 1) {
    unsigned long long* part = reinterpret_cast<unsigned long long*>(dyn);   // [NC][compute threads], stage buffers are free now
    if (!producer) {
#pragma unroll
      for (int c = 0; c < NC; ++c) part[c * (kComputeWarps * 32) + tid] = a64[c];
    }
    cluster_sync();
    if (!producer) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c % S != split) continue;
        unsigned long long sum = 0ull;
        for (int peer = 0; peer < S; ++peer) sum += ld_dsmem_u64(&part[c * (kComputeWarps * 32) + tid], peer);
        a64[c] = sum;
      }
    }
  }
  const long long k3 = DBG_T(); (void)k3;
  // epilogue: response = sum / divisor (:659), centre penalty (:727-743), store, block maximum
  unsigned long long kmax = 0ull;
  if (!producer) {
    const double da = dsub(ang, J.ca);
    const double ap = fmax(dsub(1.0, ddiv(dmul(0.25, dmul(da, da)), 0.349)), 0.9);
#pragma unroll
    for (int rx = 0; rx < RX; ++rx) {
      const int ix = tx0 + lane + 32 * rx;
      if (ix >= n_xy) continue;
      const double dx = dsub(sX[lane + 32 * rx], J.cx);
      const double dx2 = dmul(dx, dx);
      double* out = J.score + k_angle + (long long)ix * n_xy;
#pragma unroll
      for (int ry = 0; ry < RY; ++ry) {
        const int iy = ty0 + warp * RY + ry;
        if (iy >= n_xy || (S > 1 && (ry * RX + rx) % S != split)) continue;
        double sc = ddiv(dmul((double)a64[ry * RX + rx], kFixScale), J.divisor);
        if (J.use_penalty) {
          const bool zero = sc < 0.0 ? (sc >= -1e-06) : (sc <= 1e-06);
          if (!zero) {
            const double dy = dsub(sY[warp * RY + ry], J.cy);
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
    kmax = warp_max_u64(kmax);
    if (lane == 0) s_wmax[warp] = kmax;
  }
  if (err) atomicOr(J.err, err);
  __syncthreads();
  const long long k4 = DBG_T(); (void)k4;
  if (tid == 0) {
    unsigned long long m = 0ull;
    for (int w = 0; w < kComputeWarps; ++w) m = s_wmax[w] > m ? s_wmax[w] : m;
    atomicMax(J.best_key, m);
#ifdef RSM_STAGED_DEBUG
    { const int o = S > 1 ? 24 : 12; DBG_ADD(o, 1); DBG_ADD(o + 1, k1 - k0); DBG_ADD(o + 2, k2 - k1); DBG_ADD(o + 3, k3 - k2); DBG_ADD(o + 4, k4 - k3); DBG_ADD(o + 5, DBG_T() - k4); }
#endif
  }
  if (S > 1) cluster_sync();   // peers may still be reading this CTA's partial sums
}


// =================================================================================================
// stream plan: persistent CTAs over ONE sequence of (angle, tile, beam) units
// =================================================================================================
// The cluster plan above quantises: config 2 is 181 (angle, tile) items on 148 SMs -- one full wave and a
// fifth of a second one.  Here the host lays every item of a launch end to end, weighs it (beams x the
// tile's shared-memory wavefronts per beam + fixed costs) and gives each of the resident CTAs an equal,
// contiguous share (StreamCta).  A CTA walks its items with one running TMA pipeline; an item cut between
// CTAs is finished by whichever arrives last: everyone leaves integer partial sums in a slot of global
// scratch, takes a ticket, and the last one adds the other slots to its registers and runs the epilogue.
// Nobody waits for anybody, so the kernel cannot deadlock when fewer CTAs are resident than launched.
//
// Two candidate-to-thread mappings:
//   MapLanes<RX, RY>   as the cluster kernel: a warp owns RY rows, lanes walk RX x 32 columns.
//   MapPaired          for windows of 32 m + 17 .. 32 m + 32 columns (81 = 2 * 32 + 16 + 1): tiles of 81 columns.
//                      Columns 0..63 as above (two loads per row); columns 64..79 of rows r and r + 4 share ONE
//                      load (lanes 0-15 / 16-31) and column 80 of the warp's eight rows one more (lanes 0-7).
//                      The box pitch is 4 (mod 32) words, so rows r and r + 4 are 16 banks apart and eight rows
//                      spread over eight banks: both loads are conflict-free.  21 loads for 8 x 81 cells (20.25 at
//                      best) where three lane groups per row take 24.
template <int RX, int RY>
struct MapLanes {
  static constexpr int kWarps = 16, kRows = RY, kTileX = 32 * RX, kTileY = kWarps * RY, kNC = RX * RY;
  static constexpr int kW0 = 160, kH0 = 128, kW1 = 128, kH1 = 160;
  static constexpr int kBuffers = 2, kCells = kBufCells;          // two 80 KB boxes in flight
  __device__ static __forceinline__ bool slot(int c, int lane, int& dx, int& dy) {
    dx = lane + 32 * (c % RX); dy = c / RX;
    return true;
  }
  template <int SW, int NX, bool FULL>
  __device__ static __forceinline__ void beam(unsigned int (&lo)[kNC], const int* q, int rows) {
    int v[RY][NX];
#pragma unroll
    for (int ry = 0; ry < RY; ++ry)
#pragma unroll
      for (int rx = 0; rx < NX; ++rx) v[ry][rx] = (FULL || ry < rows) ? q[ry * SW + rx * 32] : 0;
#pragma unroll
    for (int ry = 0; ry < RY; ++ry)
#pragma unroll
      for (int rx = 0; rx < NX; ++rx) lo[ry * RX + rx] += (unsigned int)v[ry][rx];
  }
  // the offsets of four beams come with one 16-byte load (a broadcast: one wavefront instead of four)
  template <int SW, int NX, bool FULL>
  __device__ static __forceinline__ void run(unsigned int (&lo)[kNC], const int* q0, const int* bases, int n, int rows) {
    int j = 0;
    for (; j + 4 <= n; j += 4) {
      const int4 b = *reinterpret_cast<const int4*>(bases + j);
      beam<SW, NX, FULL>(lo, q0 + b.x, rows); beam<SW, NX, FULL>(lo, q0 + b.y, rows);
      beam<SW, NX, FULL>(lo, q0 + b.z, rows); beam<SW, NX, FULL>(lo, q0 + b.w, rows);
    }
    for (; j < n; ++j) beam<SW, NX, FULL>(lo, q0 + bases[j], rows);
  }
  template <int SW>
  __device__ static __forceinline__ void round(unsigned int (&lo)[kNC], const int* buf, const int* bases, int n, int warp, int lane,
                                               int rows, int ext_x) {
    const int* q0 = buf + warp * RY * SW + lane;
    const int nx = (ext_x + 31) >> 5;
    if (rows == RY) {
      if (nx == RX) run<SW, RX, true>(lo, q0, bases, n, rows);
      else if (RX > 2 && nx == 2) run<SW, (RX > 2 ? 2 : 1), true>(lo, q0, bases, n, rows);
      else run<SW, 1, true>(lo, q0, bases, n, rows);
    } else {
      if (nx == RX) run<SW, RX, false>(lo, q0, bases, n, rows);
      else if (RX > 2 && nx == 2) run<SW, (RX > 2 ? 2 : 1), false>(lo, q0, bases, n, rows);
      else run<SW, 1, false>(lo, q0, bases, n, rows);
    }
  }
};

struct MapPaired {
  static constexpr int kWarps = 12, kRows = 8, kTileX = 81, kTileY = kWarps * 8, kNC = 21;
  static constexpr int kW0 = 132, kH0 = 124, kW1 = 100, kH1 = 163;     // pitches = 4 (mod 32)
  // THREE 64 KB boxes in flight: a box lands through the same shared-memory port the gathers saturate, so with two
  // buffers it arrived ~540 cycles after the warps wanted it (8 % of a round); a third gives it two rounds to trickle in.
  // Same fill per beam as two 80 KB boxes (simulated on the config-2 / config-5 scans: 27.0 / 28.4 wavefronts per beam).
  static constexpr int kBuffers = 3, kCells = 16384;
  __device__ static __forceinline__ bool slot(int c, int lane, int& dx, int& dy) {
    if (c < 16) { dx = lane + 32 * (c & 1); dy = c >> 1; return true; }
    if (c < 20) { dx = 64 + (lane & 15); dy = (c - 16) + 4 * (lane >> 4); return true; }
    dx = 80; dy = lane & 7;
    return lane < 8;
  }
  template <int SW, int NF, bool PAIR, bool COL, bool FULL>
  __device__ static __forceinline__ void beam(unsigned int (&lo)[kNC], const int* qa, const int* qp, const int* qc, int rows, int rows_p,
                                              bool col_ok) {
    int va[8][NF], vp[4], vc = 0;
#pragma unroll
    for (int ry = 0; ry < 8; ++ry)
#pragma unroll
      for (int rx = 0; rx < NF; ++rx) va[ry][rx] = (FULL || ry < rows) ? qa[ry * SW + rx * 32] : 0;
    if (PAIR) {
#pragma unroll
      for (int p = 0; p < 4; ++p) vp[p] = (FULL || p < rows_p) ? qp[p * SW] : 0;
    }
    if (COL) vc = (FULL || col_ok) ? qc[0] : 0;
#pragma unroll
    for (int ry = 0; ry < 8; ++ry)
#pragma unroll
      for (int rx = 0; rx < NF; ++rx) lo[ry * 2 + rx] += (unsigned int)va[ry][rx];
    if (PAIR) {
#pragma unroll
      for (int p = 0; p < 4; ++p) lo[16 + p] += (unsigned int)vp[p];
    }
    if (COL) lo[20] += (unsigned int)vc;
  }
  // the offsets of four beams come with one 16-byte load (a broadcast: one wavefront instead of four)
  template <int SW, int NF, bool PAIR, bool COL, bool FULL>
  __device__ static __forceinline__ void run(unsigned int (&lo)[kNC], const int* qa0, const int* qp0, const int* qc0, const int* bases,
                                             int n, int rows, int rows_p, bool col_ok) {
    int j = 0;
    for (; j + 4 <= n; j += 4) {
      const int4 b = *reinterpret_cast<const int4*>(bases + j);
      beam<SW, NF, PAIR, COL, FULL>(lo, qa0 + b.x, qp0 + b.x, qc0 + b.x, rows, rows_p, col_ok);
      beam<SW, NF, PAIR, COL, FULL>(lo, qa0 + b.y, qp0 + b.y, qc0 + b.y, rows, rows_p, col_ok);
      beam<SW, NF, PAIR, COL, FULL>(lo, qa0 + b.z, qp0 + b.z, qc0 + b.z, rows, rows_p, col_ok);
      beam<SW, NF, PAIR, COL, FULL>(lo, qa0 + b.w, qp0 + b.w, qc0 + b.w, rows, rows_p, col_ok);
    }
    for (; j < n; ++j) {
      const int b = bases[j];
      beam<SW, NF, PAIR, COL, FULL>(lo, qa0 + b, qp0 + b, qc0 + b, rows, rows_p, col_ok);
    }
  }
  template <int SW>
  __device__ static __forceinline__ void round(unsigned int (&lo)[kNC], const int* buf, const int* bases, int n, int warp, int lane,
                                               int rows, int ext_x) {
    const int* qa = buf + warp * 8 * SW + lane;
    const int* qp = buf + (warp * 8 + 4 * (lane >> 4)) * SW + 64 + (lane & 15);
    const int* qc = buf + (warp * 8 + (lane & 7)) * SW + 80;
    const int rows_p = rows - 4 * (lane >> 4);
    const bool col_ok = (lane & 7) < rows;
    if (rows == 8) {
      if (ext_x > 80) run<SW, 2, true, true, true>(lo, qa, qp, qc, bases, n, rows, rows_p, col_ok);
      else if (ext_x > 64) run<SW, 2, true, false, true>(lo, qa, qp, qc, bases, n, rows, rows_p, col_ok);
      else if (ext_x > 32) run<SW, 2, false, false, true>(lo, qa, qp, qc, bases, n, rows, rows_p, col_ok);
      else run<SW, 1, false, false, true>(lo, qa, qp, qc, bases, n, rows, rows_p, col_ok);
    } else {
      if (ext_x > 80) run<SW, 2, true, true, false>(lo, qa, qp, qc, bases, n, rows, rows_p, col_ok);
      else if (ext_x > 64) run<SW, 2, true, false, false>(lo, qa, qp, qc, bases, n, rows, rows_p, col_ok);
      else if (ext_x > 32) run<SW, 2, false, false, false>(lo, qa, qp, qc, bases, n, rows, rows_p, col_ok);
      else run<SW, 1, false, false, false>(lo, qa, qp, qc, bases, n, rows, rows_p, col_ok);
    }
  }
};

template <class Map>
__global__ void __launch_bounds__((Map::kWarps + 1) * 32, 1)
score_stream_kernel(const ScoreJob* __restrict__ jobs, const int* __restrict__ item_begin, int n_jobs,
                    const StreamCta* __restrict__ plan, unsigned long long* __restrict__ partials, int* __restrict__ tickets) {
  constexpr int kW = Map::kWarps, kT = (kW + 1) * 32, kCT = kW * 32, NC = Map::kNC, RY = Map::kRows;
  constexpr int TILE_X = Map::kTileX, TILE_Y = Map::kTileY;
  constexpr int NB = Map::kBuffers, kCells = Map::kCells;
  static_assert(Map::kW0 * Map::kH0 <= kCells && Map::kW1 * Map::kH1 <= kCells, "box larger than its buffer");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp == kW;

  __shared__ ScoreJob J;
  __shared__ int s_job, s_ticket;
  __shared__ unsigned long long s_wmax[kW];
  __shared__ double sX[TILE_X], sY[TILE_Y], sDX2[TILE_X], sDY2[TILE_Y];
  __shared__ __align__(16) int sBase[NB][kRound];
  __shared__ int sMeta[NB][4];
  __shared__ __align__(8) uint64_t sFull[NB], sEmpty[NB];
  __shared__ int sUnsafeCount;
  __shared__ int sUnsafe[64];
  extern __shared__ __align__(128) unsigned char dyn[];
  int* bufs = reinterpret_cast<int*>(dyn);                         // NB box buffers of kCells cells
  int2* sBeam = reinterpret_cast<int2*>(bufs + NB * kCells + 128);   // (+ 512 bytes: lanes beyond the window read past a box)

  const StreamCta P = plan[blockIdx.x];
  if (tid == 0) {
    for (int b = 0; b < NB; ++b) { mbar_init(&sFull[b], 1); mbar_init(&sEmpty[b], kW); }
  }
  int r = 0;                       // rounds of this CTA so far: the pipeline runs on across items

  for (int item = P.item0; item <= P.item1; ++item) {
    __syncthreads();               // the previous item is finished: J, the tables and the beam list are free
    const long long d0 = DBG_T(); (void)d0;
    if (tid == 0) { s_job = find_job(item_begin, n_jobs, item); sUnsafeCount = 0; }
    __syncthreads();
    {
      const int* src = reinterpret_cast<const int*>(jobs + s_job);
      int* dst = reinterpret_cast<int*>(&J);
      for (int i = tid; i < int(sizeof(ScoreJob) / 4); i += kT) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int local = item - __ldg(item_begin + s_job);
    const int tiles = J.tiles_x * J.tiles_y;
    const int ia_local = local / tiles;
    const int tile = local - ia_local * tiles;
    const int tx0 = (tile % J.tiles_x) * TILE_X;
    const int ty0 = (tile / J.tiles_x) * TILE_Y;
    const int ia = J.ang_begin + ia_local;
    const double cs = __ldg(J.trig + 3 * ia), sn = __ldg(J.trig + 3 * ia + 1), ang = __ldg(J.trig + 3 * ia + 2);
    const int V = J.V, n_xy = J.n_xy, pitch = J.pitch, size_x = J.size_x, size_y = J.size_y;
    const int v_begin = item == P.item0 ? P.beam0 : 0;
    const int v_end = item == P.item1 ? P.beam1 : V;
    const int nv = v_end - v_begin;
    const bool first_shared = item == P.item0 && P.ticket0 >= 0;
    const bool last_shared = item != P.item0 && item == P.item1 && P.ticket1 >= 0;
    const bool shared = first_shared || last_shared;
    const int* __restrict__ grid = reinterpret_cast<const int*>(J.grid);
    const int ext_x = min(TILE_X, n_xy - tx0), ext_y = min(TILE_Y, n_xy - ty0);
    const int rows = max(0, min(RY, n_xy - (ty0 + warp * RY)));

    for (int i = tid; i < TILE_X + TILE_Y; i += kT) {
      if (i < TILE_X) {
        const double x = dadd(J.sx, dmul((double)(tx0 + i), J.f));
        const double d = dsub(x, J.cx);
        sX[i] = x; sDX2[i] = dmul(d, d);
      } else {
        const double y = dadd(J.sy, dmul((double)(ty0 + i - TILE_X), J.f));
        const double d = dsub(y, J.cy);
        sY[i - TILE_X] = y; sDY2[i - TILE_X] = dmul(d, d);
      }
    }
    __syncthreads();
    for (int j = tid; j < nv; j += kT) {
      const int p = (v_begin + j) * J.step;
      const double px = __ldg(J.pts + 2 * p), py = __ldg(J.pts + 2 * p + 1);
      const double lx = dsub(dmul(cs, px), dmul(sn, py));
      const double ly = dadd(dmul(sn, px), dmul(cs, py));
      const double tx_ = dadd(dadd(lx, sX[0]), 0.5), ty_ = dadd(dadd(ly, sY[0]), 0.5);
      const int gx0 = __double2int_rz(tx_), gy0 = __double2int_rz(ty_);
      const double fx = tx_ - (double)gx0, fy = ty_ - (double)gy0;
      const bool ok = fx > 1e-6 && fx < 1.0 - 1e-6 && fy > 1e-6 && fy < 1.0 - 1e-6 &&
                      gx0 >= 0 && gx0 + ext_x - 1 < size_x && gy0 >= 0 && gy0 + ext_y - 1 < size_y;
      sBeam[j] = ok ? make_int2(gx0, gy0) : make_int2(-1, -1);
      if (!ok) { const int pos = atomicAdd(&sUnsafeCount, 1); if (pos < 64) sUnsafe[pos] = j; }
    }
    __syncthreads();

    const long long d1 = DBG_T(); (void)d1;
    unsigned int lo[NC];
    unsigned int hi[(NC + 3) / 4];
#pragma unroll
    for (int c = 0; c < NC; ++c) lo[c] = 0u;
#pragma unroll
    for (int c = 0; c < (NC + 3) / 4; ++c) hi[c] = 0u;

    if (nv <= 0) {
      // (the host never plans an empty share)
    } else if (producer) {
      const char* tmaps = reinterpret_cast<const char*>(J.tmap);
      if (lane == 0) { tmap_acquire(tmaps); tmap_acquire(tmaps + 128); }
      __syncwarp();
      int b = 0;
      for (; b < nv; ++r) {
        const int which = r % NB, lap = r / NB;
        const long long p0 = DBG_T(); (void)p0;
        const int j = b + lane;
        int2 e = make_int2(-1, -1);
        if (j < nv) e = sBeam[j];
        const bool live = j < nv, safe = live && e.x >= 0;
        int xmin = safe ? e.x : 0x7fffffff, xmax = safe ? e.x : -1, ymin = safe ? e.y : 0x7fffffff, ymax = safe ? e.y : -1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int a0 = __shfl_up_sync(0xffffffffu, xmin, o), a1 = __shfl_up_sync(0xffffffffu, xmax, o);
          const int a2 = __shfl_up_sync(0xffffffffu, ymin, o), a3 = __shfl_up_sync(0xffffffffu, ymax, o);
          if (lane >= o) { xmin = min(xmin, a0); xmax = max(xmax, a1); ymin = min(ymin, a2); ymax = max(ymax, a3); }
        }
        const bool any = xmax >= 0;
        const int xl = xmin & ~3;
        const int dx = xmax - xl + ext_x, dy = ymax - ymin + ext_y;
        const bool fit0 = live && (!any || (dx <= Map::kW0 && dy <= Map::kH0));
        const bool fit1 = live && (!any || (dx <= Map::kW1 && dy <= Map::kH1));
        const unsigned int vote0 = __ballot_sync(0xffffffffu, fit0), vote1 = __ballot_sync(0xffffffffu, fit1);
        const int n0 = (vote0 == 0xffffffffu) ? 32 : (__ffs(~vote0) - 1);
        const int n1 = (vote1 == 0xffffffffu) ? 32 : (__ffs(~vote1) - 1);
        const bool tall = n1 > n0;
        // (a beam whose own footprint fits neither box cannot occur: the tile is smaller than both)
        const int n = max(1, tall ? n1 : n0), rw = tall ? Map::kW1 : Map::kW0;
        const int srcl = n - 1;
        const bool rany = __shfl_sync(0xffffffffu, (int)any, srcl) != 0;
        int rxl = __shfl_sync(0xffffffffu, xl, srcl), ryl = __shfl_sync(0xffffffffu, ymin, srcl);
        if (!rany) { rxl = 0; ryl = 0; }
        const long long p1 = DBG_T(); (void)p1;
        if (r >= NB) mbar_wait(&sEmpty[which], (uint32_t)((lap - 1) & 1), 100);
        const long long p2 = DBG_T(); (void)p2;
        if (lane < n) sBase[which][lane] = safe ? (e.y - ryl) * rw + (e.x - rxl) : -1;
        const unsigned int unsafe = __ballot_sync(0xffffffffu, lane < n && !safe);
        if (lane == 0) { sMeta[which][0] = n; sMeta[which][1] = tall ? 1 : 0; sMeta[which][2] = (b + n >= nv) ? 1 : 0; sMeta[which][3] = unsafe == 0u; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_expect_tx(&sFull[which], (uint32_t)((tall ? Map::kW1 * Map::kH1 : Map::kW0 * Map::kH0) * 4));
          tma_box(bufs + which * kCells, tmaps + (tall ? 128 : 0), rxl, ryl, &sFull[which]);
        }
        b += n;
#ifdef RSM_STAGED_DEBUG
        if (lane == 0) { DBG_ADD(5, p1 - p0); DBG_ADD(6, p2 - p1); DBG_ADD(7, 1); DBG_ADD(13, DBG_T() - p2); }
#endif
      }
    } else {
      int pend = 0;                // beams added since the sums were last folded into the overflow counters
      for (;; ++r) {
        const int which = r % NB;
        const long long c0 = DBG_T(); (void)c0;
        mbar_wait(&sFull[which], (uint32_t)((r / NB) & 1), 40);
        const long long c1 = DBG_T(); (void)c1;
        const int n = sMeta[which][0], tall = sMeta[which][1], last = sMeta[which][2], all_safe = sMeta[which][3];
        const int* buf = bufs + which * kCells;
        const int* bases = sBase[which];
        // a beam adds at most 2^25 per sum: after a flush (sums < 2^31) 64 beams fit before the next one is due
        if (pend + n > 64) {
#pragma unroll
          for (int c = 0; c < NC; ++c) { hi[c >> 2] += (lo[c] >> 31) << (8 * (c & 3)); lo[c] &= 0x7fffffffu; }
          pend = 0;
        }
        pend += n;
        if (rows > 0) {
          if (all_safe) {
            if (tall) Map::template round<Map::kW1>(lo, buf, bases, n, warp, lane, rows, ext_x);
            else Map::template round<Map::kW0>(lo, buf, bases, n, warp, lane, rows, ext_x);
          } else {
            // a round with a beam that failed the tile-wide index test (rare): its cells come from global memory later
            const int sw = tall ? Map::kW1 : Map::kW0;
            for (int j = 0; j < n; ++j) {
              const int base = bases[j];
              if (base < 0) continue;
#pragma unroll
              for (int c = 0; c < NC; ++c) {
                int dx, dy;
                const bool used = Map::slot(c, lane, dx, dy);
                if (used && dx < ext_x && dy < rows) lo[c] += (unsigned int)buf[base + (warp * RY + dy) * sw + dx];
              }
            }
          }
        }
        const long long c2 = DBG_T(); (void)c2;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sEmpty[which]);
#ifdef RSM_STAGED_DEBUG
        if (lane == 0 && (warp == 0 || warp == 5)) { const int o = warp == 0 ? 8 : 16; DBG_ADD(o, c1 - c0); DBG_ADD(o + 1, c2 - c1); DBG_ADD(o + 2, DBG_T() - c2); DBG_ADD(o + 3, 1); DBG_ADD(o + 4, n); }
#endif
        if (last) { ++r; break; }
      }
    }
    __syncthreads();
    const long long d2 = DBG_T(); (void)d2;

    unsigned long long a64[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) a64[c] = ((unsigned long long)((hi[c >> 2] >> (8 * (c & 3))) & 0xffu) << 31) + lo[c];

    int err = 0;
    if (!producer && rows > 0) {
      const int n_unsafe = sUnsafeCount;
      const int n_slow = n_unsafe <= 64 ? n_unsafe : nv;
      for (int u = 0; u < n_slow; ++u) {
        const int j = n_unsafe <= 64 ? sUnsafe[u] : u;
        if (sBeam[j].x >= 0) continue;
        const int p = (v_begin + j) * J.step;
        const double px = __ldg(J.pts + 2 * p), py = __ldg(J.pts + 2 * p + 1);
        const double lx = dsub(dmul(cs, px), dmul(sn, py));
        const double ly = dadd(dmul(sn, px), dmul(cs, py));
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          int dx, dy;
          const bool used = Map::slot(c, lane, dx, dy);
          if (!used || dx >= ext_x || dy >= rows) continue;
          int gx = cell_index(lx, sX[dx]), gy = cell_index(ly, sY[warp * RY + dy]);
          if (gx < 0 || gx >= size_x || gy < 0 || gy >= size_y) {
            err |= kErrWindow;
            gx = max(0, min(gx, size_x - 1)); gy = max(0, min(gy, size_y - 1));
          }
          a64[c] += (unsigned int)__ldg(grid + (gy * pitch + gx));
        }
      }
    }

    if (shared) {
      const int ticket = first_shared ? P.ticket0 : P.ticket1, slot0 = first_shared ? P.slot0 : P.slot1;
      const int part = first_shared ? P.part0 : P.part1, parts = first_shared ? P.parts0 : P.parts1;
      if (!producer) {
        unsigned long long* mine = partials + (size_t)(slot0 + part) * (NC * kCT) + tid;
#pragma unroll
        for (int c = 0; c < NC; ++c) __stcg(mine + c * kCT, a64[c]);
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) s_ticket = atomicAdd(tickets + ticket, 1);
      __syncthreads();
      if (s_ticket != parts - 1) {          // somebody else finishes this item
        if (err) atomicOr(J.err, err);
#ifdef RSM_STAGED_DEBUG
        if (tid == 0) { DBG_ADD(0, 1); DBG_ADD(1, d1 - d0); DBG_ADD(2, d2 - d1); DBG_ADD(3, DBG_T() - d2); }
#endif
        continue;
      }
      __threadfence();
      if (!producer) {
        for (int q = 0; q < parts; ++q) {
          if (q == part) continue;
          const unsigned long long* theirs = partials + (size_t)(slot0 + q) * (NC * kCT) + tid;
#pragma unroll
          for (int c = 0; c < NC; ++c) a64[c] += __ldcg(theirs + c * kCT);
        }
      }
    }

    const long long d3 = DBG_T(); (void)d3;
    // epilogue: response = sum / divisor (:659), centre penalty (:727-743), store, block maximum.  A rolled loop over the
    // thread's slots with the sums staged through shared memory (the box buffers are idle): unrolled, the 2 x NC
    // double-precision divisions alone are ~90 KB of straight-line code that runs once per item -- from a cold
    // instruction cache every time (measured: 19.6 k -> 15.4 k cycles per item).
    const long long k_angle = (long long)ia_local * n_xy * n_xy;
    unsigned long long* stage = reinterpret_cast<unsigned long long*>(dyn) + tid;    // [NC][kCT], thread-private slots
    unsigned long long kmax = 0ull;
    if (!producer) {
#pragma unroll
      for (int c = 0; c < NC; ++c) stage[c * kCT] = a64[c];
      const double da = dsub(ang, J.ca);
      const double ap = fmax(dsub(1.0, ddiv(dmul(0.25, dmul(da, da)), 0.349)), 0.9);
      const double divisor = J.divisor, m2 = J.m2, gain = J.gain, half_size = J.half_size;
      const int use_penalty = J.use_penalty;
      double* out = J.score + k_angle + (long long)tx0 * n_xy + (ty0 + warp * RY);
#pragma unroll 1
      for (int c = 0; c < NC; ++c) {
        int dx, dy;
        const bool used = Map::slot(c, lane, dx, dy);
        if (used && dx < ext_x && dy < rows) {
          double sc = ddiv(dmul((double)stage[c * kCT], kFixScale), divisor);
          if (use_penalty) {
            const bool zero = sc < 0.0 ? (sc >= -1e-06) : (sc <= 1e-06);
            if (!zero) {
              double d2 = dadd(sDX2[dx], sDY2[warp * RY + dy]);
              d2 = dmul(d2, m2);
              const double dp = fmax(dsub(1.0, ddiv(dmul(gain, d2), half_size)), 0.5);
              sc = dmul(sc, dmul(dp, ap));
            }
          }
          out[dx * n_xy + dy] = sc;
          const unsigned long long key = score_key(sc);
          kmax = key > kmax ? key : kmax;
        }
      }
      const unsigned long long wmax = warp_max_u64(kmax);
      if (lane == 0) s_wmax[warp] = wmax;
    }
    if (err) atomicOr(J.err, err);
    __syncthreads();
    unsigned long long item_key = 0ull;
    for (int w = 0; w < kW; ++w) item_key = s_wmax[w] > item_key ? s_wmax[w] : item_key;
    if (tid == 0) atomicMax(J.best_key, item_key);
    const long long d4 = DBG_T(); (void)d4;
    if (tid == 0) {
#ifdef RSM_STAGED_DEBUG
      DBG_ADD(0, 1); DBG_ADD(1, d1 - d0); DBG_ADD(2, d2 - d1); DBG_ADD(3, d3 - d2); DBG_ADD(4, d4 - d3); DBG_ADD(24, 1);
#endif
    }
  }
}

}  // namespace staged

// Tile shapes of the staged variant: 0 = 96 x 96 candidates per CTA, 1 = 64 x 64.
int score_staged_variant(int n_xy) { return n_xy > 64 ? 0 : 1; }

void score_staged_tile(int variant, int* tile_x, int* tile_y) {
  *tile_x = variant == 0 ? 96 : 64;
  *tile_y = staged::kComputeWarps * (variant == 0 ? 6 : 4);
}

void score_staged_boxes(int box_w[2], int box_h[2]) {
  box_w[0] = staged::kBoxW0; box_h[0] = staged::kBoxH0;
  box_w[1] = staged::kBoxW1; box_h[1] = staged::kBoxH1;
}

// (+ 128 bytes: lanes beyond the window in the last lane group may read up to 15 cells past a box buffer)
size_t score_staged_smem(int beams_per_split) { return size_t(2 * staged::kBufCells) * 4 + size_t(beams_per_split) * 8 + 128; }

static void (*staged_fn(int variant))(const ScoreJob*, const int*, int) {
  return variant == 0 ? staged::score_staged_kernel<3, 6> : staged::score_staged_kernel<2, 4>;
}

// guards the per-device launch-configuration caches below: contexts of different caller threads share them
static std::mutex g_config_mutex;

static cudaError_t staged_configure(int variant, size_t smem) {
  static size_t configured[kMaxDevices][2] = {{0, 0}};   // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(g_config_mutex);
  if (smem > configured[dev][variant]) {
    cudaError_t e = cudaFuncSetAttribute(staged_fn(variant), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev][variant] = smem;
  }
  return cudaSuccess;
}

// CTAs of the staged variant that can be resident at once when launched as clusters of n_split
// (a cluster lives inside one GPC, so sizes that do not divide the GPC's SM count leave SMs idle).
int score_staged_resident_ctas(int variant, int n_split, int max_beams_per_split) {
  static std::atomic<int> cache[2][9];
  if (n_split < 1 || n_split > 8) return 0;
  if (const int hit = cache[variant][n_split].load(std::memory_order_acquire)) return hit;
  const size_t smem = score_staged_smem(max_beams_per_split > 2048 ? max_beams_per_split : 2048);
  if (staged_configure(variant, smem) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_split * 64); cfg.blockDim = dim3(staged::kThreads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = n_split; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, staged_fn(variant), &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  cache[variant][n_split].store(n * n_split, std::memory_order_release);
  return n * n_split;
}

cudaError_t launch_score_staged(int variant, int n_split, int n_cta, int max_beams_per_split, cudaStream_t st,
                                const ScoreJob* jobs, const int* cta_begin, int n_jobs) {
  const size_t smem = score_staged_smem(max_beams_per_split);
  cudaError_t e = staged_configure(variant, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_cta); cfg.blockDim = dim3(staged::kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = n_split; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, staged_fn(variant), jobs, cta_begin, n_jobs);
}

// ---- stream plan: host side ------------------------------------------------------------------------
// variant 0: MapPaired (81 x 96 tiles), 1: MapLanes<2, 4> (64 x 64), 2: MapLanes<3, 6> (96 x 96, the cluster kernel's mapping)
int score_stream_variant(int n_xy) { return n_xy > 64 ? 0 : 1; }

void score_stream_tile(int variant, int* tile_x, int* tile_y) {
  if (variant == 0) { *tile_x = staged::MapPaired::kTileX; *tile_y = staged::MapPaired::kTileY; }
  else if (variant == 1) { *tile_x = 64; *tile_y = 64; }
  else { *tile_x = 96; *tile_y = 96; }
}

void score_stream_boxes(int variant, int box_w[2], int box_h[2]) {
  if (variant == 0) {
    box_w[0] = staged::MapPaired::kW0; box_h[0] = staged::MapPaired::kH0;
    box_w[1] = staged::MapPaired::kW1; box_h[1] = staged::MapPaired::kH1;
  } else {
    box_w[0] = staged::kBoxW0; box_h[0] = staged::kBoxH0;
    box_w[1] = staged::kBoxW1; box_h[1] = staged::kBoxH1;
  }
}

// shared-memory load instructions one beam costs a CTA on a tile whose window part is ext_x x ext_y
int score_stream_weight(int variant, int ext_x, int ext_y) {
  int w = 0;
  if (variant == 0) {
    const int per_row = ext_x > 32 ? 2 : 1;
    for (int y = 0; y < ext_y; y += 8) {
      const int rows = ext_y - y < 8 ? ext_y - y : 8;
      w += rows * per_row + (ext_x > 64 ? (rows < 4 ? rows : 4) : 0) + (ext_x > 80 ? 1 : 0) + 1;
    }
  } else {
    const int ry = variant == 1 ? 4 : 6, nx = (ext_x + 31) / 32;
    for (int y = 0; y < ext_y; y += ry) w += (ext_y - y < ry ? ext_y - y : ry) * nx + 1;
  }
  return w;
}

size_t score_stream_partial_words(int variant) {
  return variant == 0 ? size_t(staged::MapPaired::kNC) * staged::MapPaired::kWarps * 32
       : variant == 1 ? size_t(8) * 16 * 32 : size_t(18) * 16 * 32;
}

size_t score_stream_smem(int variant, int max_beams) {
  const size_t boxes = variant == 0 ? size_t(staged::MapPaired::kBuffers) * staged::MapPaired::kCells : size_t(2) * staged::kBufCells;
  return boxes * 4 + 512 + size_t(max_beams) * 8;
}

typedef void (*StreamFn)(const ScoreJob*, const int*, int, const StreamCta*, unsigned long long*, int*);
static StreamFn stream_fn(int variant) {
  return variant == 0 ? staged::score_stream_kernel<staged::MapPaired>
       : variant == 1 ? staged::score_stream_kernel<staged::MapLanes<2, 4>> : staged::score_stream_kernel<staged::MapLanes<3, 6>>;
}

cudaError_t launch_score_stream(int variant, int n_cta, int max_beams, cudaStream_t st, const ScoreJob* jobs, const int* item_begin,
                                int n_jobs, const StreamCta* plan, unsigned long long* partials, int* tickets) {
  static size_t configured[kMaxDevices][3] = {{0, 0, 0}};
  if (variant < 0 || variant > 2) return cudaErrorInvalidValue;
  const size_t smem = score_stream_smem(variant, max_beams);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  {
    std::lock_guard<std::mutex> lock(g_config_mutex);
    if (smem > configured[dev][variant]) {
      cudaError_t e = cudaFuncSetAttribute(stream_fn(variant), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured[dev][variant] = smem;
    }
  }
  const int threads = variant == 0 ? (staged::MapPaired::kWarps + 1) * 32 : staged::kThreads;
  stream_fn(variant)<<<n_cta, threads, smem, st>>>(jobs, item_begin, n_jobs, plan, partials, tickets);
  return cudaGetLastError();
}

#ifdef RSM_STAGED_DEBUG
extern "C" int rsm_debug_staged(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, staged::g_dbg, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(staged::g_dbg, z, sizeof z); }
  return 0;
}
#endif

}  // namespace rsm

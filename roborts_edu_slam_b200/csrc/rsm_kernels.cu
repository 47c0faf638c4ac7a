// rsm_kernels.cu -- hand-written sm_100a kernels of the correlative scan matcher.
//
//   score_kernel    the hot loop: every (angle, x, y) candidate x every visited beam
//                   (reference: MultiResolutionCorrelateScanMatcher::ScanMatch loop + GetResponse +
//                   PenalizeResponse, scan_match/correlate_scan_matcher.h:552-603, 637-662, 718-745)
//   select_kernel   what the reference gets from std::sort: the maximum's neighbourhood
//                   (averaging set of FindBestCandidate, :670-710) and the global top-21
//                   (covariance prefix, :887-956)
//   gather_kernel   scores of the <= 9 same-(x,y) columns for the angular covariance (:965-1019)
//   raster_kernel   lookup-grid construction (OccuGridMap::UpdateMapByRange / SetCellOccuBlur,
//                   map/occu_grid_map.h:258-329, 531-576) as atomicMax compositing
//
// Design (DESIGN.md has the long version): this is a gather-reduce, not a contraction, so no
// tensor cores.  Lanes of a warp walk consecutive x translations, so one warp load touches one
// grid row segment (1-2 cache lines) instead of 32 scattered cells; each thread owns RY
// consecutive y translations and keeps their sums in registers, so there is no cross-lane
// reduction and the float32 fallback adds in the reference's beam order.  Cell indices are
// produced per (beam, x) and (beam, y) -- never per candidate -- in FP64 with the reference's
// exact operation order (fused multiply-add is off), staged through double-buffered shared
// memory tables, so the inner loop is: 1 table load + RY x (address add, 4-byte gather, add).
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
#include <cuda_runtime.h>
#include <stdint.h>

#include "rsm_device.h"
#include "rsm_kernels.h"

namespace rsm {

// ---- exact FP64 helpers: every reference operation is one IEEE op, never contracted ---------
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// order-preserving map double -> uint64 (so atomicMax works on scores of either sign)
__device__ __forceinline__ unsigned long long score_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_score(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// largest j with cta_begin[j] <= b
__device__ __forceinline__ int find_job(const int* __restrict__ cta_begin, int n_jobs, int b) {
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(cta_begin + mid) <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// =================================================================================================
// scoring
// =================================================================================================
template <int RYP> struct RowVec;
template <> struct RowVec<1> { __device__ static void load(const int* p, int* r) { r[0] = p[0]; } };
template <> struct RowVec<2> { __device__ static void load(const int* p, int* r) { int2 v = *reinterpret_cast<const int2*>(p); r[0] = v.x; r[1] = v.y; } };
template <> struct RowVec<4> { __device__ static void load(const int* p, int* r) { int4 v = *reinterpret_cast<const int4*>(p); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; } };
template <> struct RowVec<8> { __device__ static void load(const int* p, int* r) { RowVec<4>::load(p, r); RowVec<4>::load(p + 4, r + 4); } };

template <bool FIXED, int LX, int RY>
__global__ void __launch_bounds__(256)
score_kernel(const ScoreJob* __restrict__ jobs, const int* __restrict__ cta_begin, int n_jobs) {
  constexpr int RYP = (RY <= 1) ? 1 : (RY <= 2) ? 2 : (RY <= 4) ? 4 : 8;
  constexpr int PC = kChunk;
  const int NT = blockDim.x;
  const int tid = threadIdx.x;
  const int slots = NT / LX;     // y slots per CTA; each owns RY consecutive y translations
  const int rows = slots * RY;
  const int tx = tid % LX;       // lane position along x
  const int ts = tid / LX;       // y slot

  __shared__ ScoreJob J;
  __shared__ int s_job;
  __shared__ unsigned long long s_wmax[8];
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  double* sLut = reinterpret_cast<double*>(dyn_smem);   // [2][PC][2] rotated endpoints (x,y)
  double* sX = sLut + 2 * PC * 2;                        // [LX]   candidate x of this tile
  double* sY = sX + LX;                                  // [rows] candidate y of this tile
  int* sGX = reinterpret_cast<int*>(sY + rows);          // [2][PC][LX]          cell x
  int* sGY = sGX + 2 * PC * LX;                          // [2][PC][slots][RYP]  cell y * pitch

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
  const int ty0 = (tile / J.tiles_x) * rows;
  const int ia = J.ang_begin + ia_local;
  const double cs = __ldg(J.trig + 3 * ia), sn = __ldg(J.trig + 3 * ia + 1), ang = __ldg(J.trig + 3 * ia + 2);
  const int V = J.V, n_xy = J.n_xy;
  const int nchunks = (V + PC - 1) / PC;

  // candidate coordinates of this tile: x = start_x + x_index * factor   (:569, :572)
  for (int i = tid; i < LX + rows; i += NT) {
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
        double* d = sLut + ((c & 1) * PC + tid) * 2;
        d[0] = dsub(dmul(cs, px), dmul(sn, py));
        d[1] = dadd(dmul(sn, px), dmul(cs, py));
      }
    }
  };
  int err = 0;
  // cell index tables of chunk c: (int)(lut + candidate + 0.5), truncation toward zero   (:647-648)
  auto build_chunk = [&](int c) {
    const int npc = min(PC, V - c * PC);
    const double* lut = sLut + (c & 1) * PC * 2;
    int* gxt = sGX + (c & 1) * PC * LX;
    int* gyt = sGY + (c & 1) * PC * slots * RYP;
    for (int q = tid; q < npc * LX; q += NT) {
      const int pc = q / LX, j = q % LX;
      int g = __double2int_rz(dadd(dadd(lut[pc * 2], sX[j]), 0.5));
      if (g < 0 || g >= J.size_x) {
        if (tx0 + j < n_xy) err |= kErrWindow;
        g = max(0, min(g, J.size_x - 1));
      }
      gxt[pc * LX + j] = g;
    }
    for (int q = tx; q < npc * RY; q += LX) {
      const int pc = q / RY, rr = q % RY;
      int g = __double2int_rz(dadd(dadd(lut[pc * 2 + 1], sY[ts * RY + rr]), 0.5));
      if (g < 0 || g >= J.size_y) {
        if (ty0 + ts * RY + rr < n_xy) err |= kErrWindow;
        g = max(0, min(g, J.size_y - 1));
      }
      gyt[(pc * slots + ts) * RYP + rr] = g * J.pitch;
    }
  };

  unsigned int a32[RY];
  unsigned long long a64[RY];
  double ad[RY];
#pragma unroll
  for (int r = 0; r < RY; ++r) { a32[r] = 0u; a64[r] = 0ull; ad[r] = 0.0; }

  const int* gridI = reinterpret_cast<const int*>(J.grid);
  const float* gridF = reinterpret_cast<const float*>(J.grid);

  lut_chunk(0);
  __syncthreads();
  build_chunk(0);
  if (nchunks > 1) lut_chunk(1);
  __syncthreads();

  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) build_chunk(c + 1);
    if (c + 2 < nchunks) lut_chunk(c + 2);

    const int npc = min(PC, V - c * PC);
    const int* gxp = sGX + (c & 1) * PC * LX + tx;
    const int* gyp = sGY + ((c & 1) * PC * slots + ts) * RYP;
    const int gy_stride = slots * RYP;
    if (npc == PC) {
#pragma unroll 8
      for (int pc = 0; pc < PC; ++pc) {
        const int gx = gxp[pc * LX];
        int ro[RYP];
        RowVec<RYP>::load(gyp + pc * gy_stride, ro);
#pragma unroll
        for (int r = 0; r < RY; ++r) {
          if (FIXED) a32[r] += (unsigned int)__ldg(gridI + (ro[r] + gx));
          else ad[r] = dadd(ad[r], (double)__ldg(gridF + (ro[r] + gx)));
        }
      }
    } else {
      for (int pc = 0; pc < npc; ++pc) {
        const int gx = gxp[pc * LX];
        int ro[RYP];
        RowVec<RYP>::load(gyp + pc * gy_stride, ro);
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
    for (int w = 0; w < (NT >> 5); ++w) m = s_wmax[w] > m ? s_wmax[w] : m;
    atomicMax(J.best_key, m);
  }
}

size_t score_smem_bytes(int lx, int ry, int nt) {
  const int ryp = (ry <= 1) ? 1 : (ry <= 2) ? 2 : (ry <= 4) ? 4 : 8;
  const int slots = nt / lx;
  const int rows = slots * ry;
  return size_t(2 * kChunk * 2 + lx + rows) * 8 + size_t(2 * kChunk * lx + 2 * kChunk * slots * ryp) * 4;
}

template <bool FIXED, int LX>
static cudaError_t launch_score_ry(int ry, int n_cta, int nt, size_t smem, cudaStream_t st,
                                   const ScoreJob* jobs, const int* cta_begin, int n_jobs) {
#define RSM_CASE(R)                                                                            \
  case R:                                                                                      \
    score_kernel<FIXED, LX, R><<<n_cta, nt, smem, st>>>(jobs, cta_begin, n_jobs);              \
    break;
  switch (ry) {
    RSM_CASE(1) RSM_CASE(2) RSM_CASE(3) RSM_CASE(4) RSM_CASE(5) RSM_CASE(6) RSM_CASE(7) RSM_CASE(8)
    default: return cudaErrorInvalidValue;
  }
#undef RSM_CASE
  return cudaGetLastError();
}

cudaError_t launch_score(bool fixed, int lx, int ry, int nt, int n_cta, cudaStream_t st,
                         const ScoreJob* jobs, const int* cta_begin, int n_jobs) {
  const size_t smem = score_smem_bytes(lx, ry, nt);
#define RSM_LX(L)                                                                               \
  case L:                                                                                       \
    return fixed ? launch_score_ry<true, L>(ry, n_cta, nt, smem, st, jobs, cta_begin, n_jobs)   \
                 : launch_score_ry<false, L>(ry, n_cta, nt, smem, st, jobs, cta_begin, n_jobs);
  switch (lx) {
    RSM_LX(4) RSM_LX(8) RSM_LX(16) RSM_LX(32)
    default: return cudaErrorInvalidValue;
  }
#undef RSM_LX
}

// =================================================================================================
// selection
// =================================================================================================
// Block-wide "pop the maximum" over one (key, tag) pair per thread.  Returns the winning key and
// tag to every thread; among equal keys the smallest tag wins.  tag must be unique per live item.
struct KeyTag { unsigned long long key; unsigned int tag; };

__device__ __forceinline__ KeyTag block_argmax(unsigned long long key, unsigned int tag,
                                               unsigned long long* s_k, unsigned int* s_t) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
    unsigned int t2 = __shfl_xor_sync(0xffffffffu, tag, o);
    if (k2 > key || (k2 == key && t2 < tag)) { key = k2; tag = t2; }
  }
  __syncthreads();  // protect s_k/s_t from the previous round's readers
  if (lane == 0) { s_k[warp] = key; s_t[warp] = tag; }
  __syncthreads();
  KeyTag r;
  r.key = s_k[0]; r.tag = s_t[0];
  for (int w = 1; w < nw; ++w) {
    const unsigned long long k2 = s_k[w];
    const unsigned int t2 = s_t[w];
    if (k2 > r.key || (k2 == r.key && t2 < r.tag)) { r.key = k2; r.tag = t2; }
  }
  return r;
}

__global__ void __launch_bounds__(256)
select_kernel(const SelectJob* __restrict__ jobs, const int* __restrict__ cta_begin, int n_jobs,
              PoolEntry* __restrict__ pool, int pool_cap, int* __restrict__ pool_count) {
  __shared__ SelectJob J;
  __shared__ int s_job;
  __shared__ unsigned long long s_k[8];
  __shared__ unsigned int s_t[8];
  __shared__ unsigned long long s_bkey[kSelectBuf];
  __shared__ unsigned int s_bidx[kSelectBuf];   // index relative to the slice start
  __shared__ int s_count;
  const int tid = threadIdx.x, NT = blockDim.x;

  if (tid == 0) { s_job = find_job(cta_begin, n_jobs, blockIdx.x); s_count = 0; }
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(jobs + s_job);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < int(sizeof(SelectJob) / 4); i += NT) dst[i] = __ldg(src + i);
  }
  const int cta = blockIdx.x - __ldg(cta_begin + s_job);
  __syncthreads();

  const long long lo = (long long)cta * J.slice;
  const long long hi = min(J.n, lo + J.slice);
  const double best = key_score(*J.best_key);

  // pass 1: averaging-set candidates (DoubleEqual(score, best, 1e-2), :685) and thread maxima
  unsigned long long tmax = 0ull;
  for (long long k = lo + tid; k < hi; k += NT) {
    const double s = J.score[k];
    const double delta = dsub(s, best);
    const bool near_best = delta < 0.0 ? (delta >= -1e-2) : (delta <= 1e-2);
    if (near_best) {
      const int pos = atomicAdd(pool_count, 1);
      if (pos < pool_cap) { PoolEntry e; e.score = s; e.index = (int)k; e.job = J.job_id; pool[pos] = e; }
      else atomicOr(J.err, kErrPoolFull);
    }
    const unsigned long long key = score_key(s);
    tmax = key > tmax ? key : tmax;
  }
  // threshold = the kTopK-th largest thread maximum: a lower bound of the slice's kTopK-th largest
  // element, so everything in the slice's top-kTopK is >= threshold.
  unsigned long long thr = 0ull;
  {
    unsigned long long mine = tmax;
    for (int r = 0; r < kTopK; ++r) {
      KeyTag w = block_argmax(mine, (unsigned int)tid, s_k, s_t);
      thr = w.key;
      if (w.key == 0ull) break;  // fewer than kTopK non-empty threads: keep everything
      if (w.tag == (unsigned int)tid) mine = 0ull;
    }
  }
  __syncthreads();
  // pass 2: everything >= threshold goes to the shared buffer
  for (long long k = lo + tid; k < hi; k += NT) {
    const unsigned long long key = score_key(J.score[k]);
    if (key >= thr) {
      const int pos = atomicAdd(&s_count, 1);
      if (pos < kSelectBuf) { s_bkey[pos] = key; s_bidx[pos] = (unsigned int)(k - lo); }
    }
  }
  __syncthreads();
  int count = s_count;
  if (count > kSelectBuf) {
    if (tid == 0) atomicOr(J.err, kErrSelectFull);
    count = kSelectBuf;
  }
  // pop the kTopK largest of the buffer
  const int emit = min(count, kTopK);
  for (int r = 0; r < emit; ++r) {
    unsigned long long key = 0ull;
    unsigned int tag = 0xffffffffu;
    for (int i = tid; i < count; i += NT) {
      const unsigned long long k2 = s_bkey[i];
      if (k2 > key || (k2 == key && (unsigned int)i < tag)) { key = k2; tag = (unsigned int)i; }
    }
    KeyTag w = block_argmax(key, tag, s_k, s_t);
    if (tid == 0) {
      Entry e;
      e.score = key_score(w.key);
      e.index = lo + s_bidx[w.tag];
      J.top_list[cta * kTopK + r] = e;
      s_bkey[w.tag] = 0ull;
    }
    __syncthreads();
  }
  if (tid == 0) J.top_count[cta] = emit;
}

cudaError_t launch_select(int n_cta, cudaStream_t st, const SelectJob* jobs, const int* cta_begin,
                          int n_jobs, PoolEntry* pool, int pool_cap, int* pool_count) {
  select_kernel<<<n_cta, 256, 0, st>>>(jobs, cta_begin, n_jobs, pool, pool_cap, pool_count);
  return cudaGetLastError();
}

// =================================================================================================
// gather of same-(x,y) columns
// =================================================================================================
__global__ void __launch_bounds__(128) gather_kernel(const GatherJob* __restrict__ jobs) {
  const GatherJob& J = jobs[blockIdx.x];
  const long long plane = (long long)J.n_xy * J.n_xy;
  const int total = J.n_cols * J.n_ang;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = i / J.n_ang, ia = i - c * J.n_ang;
    J.out[i] = J.score[(long long)ia * plane + J.cols[c]];
  }
}

cudaError_t launch_gather(int n_jobs, cudaStream_t st, const GatherJob* jobs) {
  gather_kernel<<<n_jobs, 128, 0, st>>>(jobs);
  return cudaGetLastError();
}

// =================================================================================================
// lookup-grid construction
// =================================================================================================
__global__ void __launch_bounds__(256) fill_kernel(const FillJob* __restrict__ jobs, int ctas_per_job) {
  const FillJob J = jobs[blockIdx.x / ctas_per_job];
  const int part = blockIdx.x % ctas_per_job;
  int4* g4 = reinterpret_cast<int4*>(J.grid);
  const long long n4 = J.n_cells >> 2;
  const int4 v4 = make_int4(J.value, J.value, J.value, J.value);
  for (long long i = (long long)part * blockDim.x + threadIdx.x; i < n4; i += (long long)ctas_per_job * blockDim.x)
    g4[i] = v4;
  if (part == 0) {
    int* g = reinterpret_cast<int*>(J.grid);
    for (long long i = (n4 << 2) + threadIdx.x; i < J.n_cells; i += blockDim.x) g[i] = J.value;
  }
}

cudaError_t launch_fill(int n_jobs, int ctas_per_job, cudaStream_t st, const FillJob* jobs) {
  fill_kernel<<<n_jobs * ctas_per_job, 256, 0, st>>>(jobs, ctas_per_job);
  return cudaGetLastError();
}

// One CTA per base scan.  Cells hold non-negative values whose 32-bit pattern orders like the
// value (fixed-point ints, or IEEE floats >= 0), so "cell = max(cell, v)" is an integer atomicMax
// and the result does not depend on the order the reference visits points and scans in.
// stamp[(2h+1)^2]: pattern of float(kernel[i,j] * occu_offset), 0 where that exceeds 1.0f
// (SetGridProbability ignores prob > 1, map/grid_map_cell.h:361-365); one = pattern of 1.0f.
__global__ void __launch_bounds__(128)
raster_kernel(const RasterScan* __restrict__ scans, const int* __restrict__ stamp, int half, int one) {
  const RasterScan S = scans[blockIdx.x];
  int* grid = reinterpret_cast<int*>(S.grid);
  const int ks = 2 * half + 1;
  const int tol = half + 1;
  for (int i = threadIdx.x; i < S.n_pts; i += blockDim.x) {
    const double px = S.pts[2 * i], py = S.pts[2 * i + 1];
    // Affine * v = translation + (l00*x + l01*y), l = (c, -s; s, c)   (occu_grid_map.h:279-283)
    const double mx = dadd(S.tx, dadd(dmul(S.c, px), dmul(-S.s, py)));
    const double my = dadd(S.ty, dadd(dmul(S.s, px), dmul(S.c, py)));
    const int ex = __double2int_rz(dadd(mx, 0.5));
    const int ey = __double2int_rz(dadd(my, 0.5));
    if (ex == S.start_x && ey == S.start_y) continue;                    // :312
    if (!(ex > tol && ex < S.size_x - tol && ey > tol && ey < S.size_y - tol)) continue;  // :476
    atomicMax(grid + (long long)ey * S.pitch + ex, one);                  // :544
    for (int j = -half; j <= half; ++j)
      for (int ii = -half; ii <= half; ++ii) {
        const int v = __ldg(stamp + (ii + half) + ks * (j + half));
        if (v) atomicMax(grid + (long long)(ey + j) * S.pitch + (ex + ii), v);   // :560-573
      }
  }
}

cudaError_t launch_raster(int n_scans, cudaStream_t st, const RasterScan* scans, const int* stamp, int half, int one) {
  if (n_scans == 0) return cudaSuccess;
  raster_kernel<<<n_scans, 128, 0, st>>>(scans, stamp, half, one);
  return cudaGetLastError();
}

// L2 flush: stream a buffer larger than L2 through it
__global__ void __launch_bounds__(256) flush_kernel(int4* buf, long long n4, int v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    int4 x = buf[i];
    x.x += v;
    buf[i] = x;
  }
}

cudaError_t launch_flush(cudaStream_t st, void* buf, long long bytes, int v) {
  flush_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<int4*>(buf), bytes / 16, v);
  return cudaGetLastError();
}

}  // namespace rsm

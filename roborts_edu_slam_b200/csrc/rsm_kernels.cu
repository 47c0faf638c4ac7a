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
#include <limits.h>

#include <atomic>
#include <stdint.h>

#include "rsm_device.h"
#include "rsm_kernels.h"
#include "rsm_select.cuh"

namespace rsm {

// One CTA per slice of a job's score array; the last CTA of a job to finish merges the per-CTA
// lists into the job's final top-kTopK and gathers, speculatively, the scores of the 3x3
// translation neighbourhood of the best candidate over all angles: those are the same-(x,y)
// columns the angular covariance needs whenever the averaged best pose stays in that
// neighbourhood (always when the averaging set has one member).
__global__ void __launch_bounds__(kSelectThreads)
select_kernel(const SelectJob* __restrict__ jobs, const int* __restrict__ cta_begin, int n_jobs,
              PoolEntry* __restrict__ pool, int pool_cap, int* __restrict__ pool_count) {
  __shared__ SelectJob J;
  __shared__ int s_job;
  __shared__ unsigned long long s_k[kSelectThreads / 32];
  __shared__ unsigned long long s_bkey[kSelectBuf];
  __shared__ unsigned int s_bidx[kSelectBuf];
  __shared__ int s_count[2];
  __shared__ int s_last;
  extern __shared__ unsigned long long s_cache[];   // kSelectSlice keys of the first pass, for the second
  const int tid = threadIdx.x, NT = blockDim.x;

  if (tid == 0) s_job = find_job(cta_begin, n_jobs, blockIdx.x);
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(jobs + s_job);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = tid; i < int(sizeof(SelectJob) / 4); i += NT) dst[i] = __ldg(src + i);
  }
  const int cta = blockIdx.x - __ldg(cta_begin + s_job);
  __syncthreads();
  SDBG_MIN(0); SDBG_MAX(1);

  const long long lo = (long long)cta * J.slice;
  const long long hi = min(J.n, lo + J.slice);
  const unsigned int n_slice = hi > lo ? (unsigned int)(hi - lo) : 0u;
  const double best = key_score(*J.best_key);
  // keys below this cannot be within 1e-2 of the best score: skips the FP64 test for almost all
  const unsigned long long near_floor = score_key(dsub(best, 0.0101));
  const double* __restrict__ sc = J.score + lo;

  bool overflow = false;
  Entry* my_list = J.top_list + (size_t)cta * kTopK;
  const int emitted = block_top_k(
      n_slice, [&](unsigned int i) { return score_key(sc[i]); },
      [&](unsigned int i, unsigned long long key) {
        // averaging-set candidates: DoubleEqual(score, best, 1e-2)   (:685)
        if (key < near_floor) return;
        const double s = key_score(key);
        const double delta = dsub(s, best);
        const bool near_best = delta < 0.0 ? (delta >= -1e-2) : (delta <= 1e-2);
        if (near_best) {
          const int pos = atomicAdd(pool_count, 1);
          if (pos < pool_cap) { PoolEntry e; e.score = s; e.index = (int)(lo + i); e.job = J.job_id; pool[pos] = e; }
          else atomicOr(J.err, kErrPoolFull);
        }
      },
      [&](int r, unsigned long long key, unsigned int i) { Entry e; e.score = key_score(key); e.index = lo + i; my_list[r] = e; },
      s_k, s_bkey, s_bidx, s_count, &overflow, s_cache, kSelectSlice);
  // unused slots get a score whose key is 0 so that the merge needs no per-list count
  for (int r = emitted + tid; r < kTopK; r += NT) { Entry e; e.score = __longlong_as_double(-1ll); e.index = -1; my_list[r] = e; }
  __threadfence();     // every thread's list entries are visible device-wide before the ticket
  __syncthreads();
  SDBG_MIN(2); SDBG_MAX(3);
  if (tid == 0) {
    J.top_count[cta] = emitted;
    if (overflow) atomicOr(J.err, kErrSelectFull);
    __threadfence();
    s_last = (atomicAdd(J.done, 1) == J.n_cta - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  SDBG(4);

  // ---- last CTA of the job: merge the per-CTA lists, gather the best candidate's columns -----------
  select_job_tail(J, J.n_cta, emitted, s_k, s_bkey, s_bidx, s_count, s_cache, kSelectSlice);
  SDBG(6);
}
#ifdef RSM_SELECT_DEBUG
extern "C" int rsm_debug_select(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_sel_dbg, sizeof(unsigned long long) * 16);
  unsigned long long z[16]; for (int i = 0; i < 16; ++i) z[i] = (i == 0 || i == 2) ? ~0ull : 0ull;
  cudaMemcpyToSymbol(g_sel_dbg, z, sizeof z);
  return 0;
}
#endif

cudaError_t launch_select(int n_cta, cudaStream_t st, const SelectJob* jobs, const int* cta_begin,
                          int n_jobs, PoolEntry* pool, int pool_cap, int* pool_count) {
  const size_t smem = size_t(kSelectSlice) * 8;
  // function attributes are per device; contexts of different caller threads may get here at the same time
  static std::atomic<bool> configured[kMaxDevices];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  if (!configured[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // idempotent
    if (e != cudaSuccess) return e;
    configured[dev].store(true, std::memory_order_release);
  }
  select_kernel<<<n_cta, kSelectThreads, smem, st>>>(jobs, cta_begin, n_jobs, pool, pool_cap, pool_count);
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

// Partial of one slice, packed where it lies (rsm_match_sliced): header, the pool's candidates, the slice's top list.
__global__ void __launch_bounds__(256) pack_partial_kernel(PackJob J) {
  PartialHeader* H = reinterpret_cast<PartialHeader*>(J.out);
  Entry* out = reinterpret_cast<Entry*>(J.out + sizeof(PartialHeader));
  const int cap = int((kPartialDevBytes - sizeof(PartialHeader)) / sizeof(Entry)) - kTopK;
  const int e = *J.err;
  const int count = *J.pool_count;
  const bool exact = (e & (kErrPoolFull | kErrSelectFull)) || count > J.pool_cap || count > cap;
  const int n_pool = exact ? 0 : count;
  const int n_top = *J.fcnt;
  for (int i = threadIdx.x; i < n_pool; i += blockDim.x) { Entry c; c.score = J.pool[i].score; c.index = J.base + J.pool[i].index; out[i] = c; }
  for (int i = threadIdx.x; i < n_top; i += blockDim.x) { Entry c = J.ftop[i]; c.index += J.base; out[n_pool + i] = c; }
  if (threadIdx.x == 0) {
    PartialHeader h;
    h.magic = kPartialMagic; h.a0 = J.a0; h.a1 = J.a1; h.n_ang = J.n_ang; h.n_xy = J.n_xy; h.n_pool = n_pool; h.n_top = n_top;
    h.flags = (exact ? 1 : 0) | ((e & kErrWindow) ? 2 : 0);
    h.best_key = *J.best_key;
    h.reserved[0] = h.reserved[1] = h.reserved[2] = 0.0;
    *H = h;
  }
}

cudaError_t launch_pack_partial(cudaStream_t st, const PackJob& job) {
  pack_partial_kernel<<<1, 256, 0, st>>>(job);
  return cudaGetLastError();
}

// same-(x,y) columns of a slice straight into an exchange buffer: out[c * stride + ia]
__global__ void __launch_bounds__(128) gather_columns_kernel(GatherJob J, int stride) {
  const long long plane = (long long)J.n_xy * J.n_xy;
  const int total = J.n_cols * J.n_ang;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i / J.n_ang, ia = i - c * J.n_ang;
    J.out[(long long)c * stride + ia] = J.score[(long long)ia * plane + J.cols[c]];
  }
}

cudaError_t launch_gather_columns(cudaStream_t st, const GatherJob& job, int stride) {
  gather_columns_kernel<<<4, 128, 0, st>>>(job, stride);
  return cudaGetLastError();
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
// Two phases per chunk of points: (1) one thread per point computes the end cell in FP64 (the reference's
// association) and parks it in shared memory; (2) one thread per (point, stamp row) composites 2h+1 consecutive
// cells.  A cell is read first (L2) and the atomic is only issued when it would raise the value: the maximum is
// monotone, so a stale read can only cause a redundant atomic, never a missed one.  Neighbouring scan points
// and the scans of one chain overlap heavily, which removes most of the atomics.
constexpr int kRasterChunk = 1024;
__global__ void __launch_bounds__(256)
raster_kernel(const RasterScan* __restrict__ scans, const int* __restrict__ stamp, int half, int one) {
  __shared__ int2 s_cell[kRasterChunk];
  const RasterScan S = scans[blockIdx.x];
  int* grid = reinterpret_cast<int*>(S.grid);
  const int ks = 2 * half + 1;
  const int tol = half + 1;
  for (int base = 0; base < S.n_pts; base += kRasterChunk) {
    const int cnt = min(kRasterChunk, S.n_pts - base);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const double px = S.pts[2 * (base + i)], py = S.pts[2 * (base + i) + 1];
      // Affine * v = translation + (l00*x + l01*y), l = (c, -s; s, c)   (occu_grid_map.h:279-283)
      const double mx = dadd(S.tx, dadd(dmul(S.c, px), dmul(-S.s, py)));
      const double my = dadd(S.ty, dadd(dmul(S.s, px), dmul(S.c, py)));
      int ex = __double2int_rz(dadd(mx, 0.5));
      const int ey = __double2int_rz(dadd(my, 0.5));
      if (ex == S.start_x && ey == S.start_y) ex = INT_MIN;                                   // :312
      else if (!(ex > tol && ex < S.size_x - tol && ey > tol && ey < S.size_y - tol)) ex = INT_MIN;  // :476
      s_cell[i] = make_int2(ex, ey);
    }
    __syncthreads();
    for (int w = threadIdx.x; w < cnt * ks; w += blockDim.x) {
      const int i = w / ks, j = w - i * ks - half;
      const int2 c = s_cell[i];
      if (c.x == INT_MIN) continue;
      int* row = grid + (long long)(c.y + j) * S.pitch + c.x;
      for (int ii = -half; ii <= half; ++ii) {
        int v = __ldg(stamp + (ii + half) + ks * (j + half));                                   // :560-573
        if (j == 0 && ii == 0) v = max(v, one);                                                  // :544
        if (v && __ldcg(row + ii) < v) atomicMax(row + ii, v);
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_raster(int n_scans, cudaStream_t st, const RasterScan* scans, const int* stamp, int half, int one) {
  if (n_scans == 0) return cudaSuccess;
  raster_kernel<<<n_scans, 256, 0, st>>>(scans, stamp, half, one);
  return cudaGetLastError();
}

// SET_CELL_OCCUPIED (UpdateMapByRange with use_blur false, or blur parameters the reference rejects): SetCellOccu
// (map/occu_grid_map.h:499-516) adds ProbabilityCellFunctions' update_occu_factor_ (0.5f, map/grid_map_cell.h:333-342),
// clamped to 1, ONCE per cell and update -- the reference guards with the cell's update_index_.  That is not a maximum,
// so the scans of one grid are applied one after the other: one CTA per grid walks its scans in order; within a scan
// the first thread to reach a cell claims it by setting the (otherwise unused) sign bit together with the new value in
// one compare-and-swap, later points of the same scan see the bit and pass; a second sweep over the scan's points
// clears the bits.  Cells are non-negative fixed-point ints or float bit patterns, so bit 31 is free in both.
__global__ void __launch_bounds__(256)
raster_occu_kernel(const RasterScan* __restrict__ scans, const int* __restrict__ group_begin, int half, int fixed) {
  const int s_begin = group_begin[blockIdx.x], s_end = group_begin[blockIdx.x + 1];
  const int tol = half + 1;
  const unsigned int kTag = 0x80000000u;
  for (int s = s_begin; s < s_end; ++s) {
    const RasterScan S = scans[s];
    unsigned int* grid = reinterpret_cast<unsigned int*>(S.grid);
    for (int sweep = 0; sweep < 2; ++sweep) {
      for (int i = threadIdx.x; i < S.n_pts; i += blockDim.x) {
        const double px = S.pts[2 * i], py = S.pts[2 * i + 1];
        const double mx = dadd(S.tx, dadd(dmul(S.c, px), dmul(-S.s, py)));      // occu_grid_map.h:279-283
        const double my = dadd(S.ty, dadd(dmul(S.s, px), dmul(S.c, py)));
        const int ex = __double2int_rz(dadd(mx, 0.5)), ey = __double2int_rz(dadd(my, 0.5));
        if (ex == S.start_x && ey == S.start_y) continue;                                          // :312
        if (!(ex > tol && ex < S.size_x - tol && ey > tol && ey < S.size_y - tol)) continue;       // :476
        unsigned int* cell = grid + (long long)ey * S.pitch + ex;
        if (sweep == 1) { atomicAnd(cell, ~kTag); continue; }
        unsigned int old = *reinterpret_cast<volatile unsigned int*>(cell);
        while (!(old & kTag)) {
          float v = fixed ? __fmul_rn(__int2float_rn((int)old), 2.98023223876953125e-08f) : __uint_as_float(old);   // 2^-25
          v = __fadd_rn(v, 0.5f);                                                                // UpdateSetOccupied
          if (v > 1.0f) v = 1.0f;
          const unsigned int nv = (fixed ? (unsigned int)__float2int_rn(__fmul_rn(v, 33554432.0f)) : __float_as_uint(v)) | kTag;
          const unsigned int seen = atomicCAS(cell, old, nv);
          if (seen == old) break;
          old = seen;
        }
      }
      __syncthreads();     // every claim of this scan is in place before the bits are cleared; every bit is cleared before the next scan
    }
  }
}

cudaError_t launch_raster_occu(int n_groups, cudaStream_t st, const RasterScan* scans, const int* group_begin, int half, int fixed) {
  if (n_groups == 0) return cudaSuccess;
  raster_occu_kernel<<<n_groups, 256, 0, st>>>(scans, group_begin, half, fixed);
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

// ---- gather-bandwidth micro-benchmarks (roofline denominators) -------------------
// Every warp reads `iters` row segments of 32 consecutive 4-byte words (the access shape of
// score_kernel) or 32 independent random words, from a shared-memory tile or from a global
// footprint that lives in L1/L2.
__device__ __forceinline__ unsigned int mix(unsigned int x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(1024)
microbench_kernel(const int* __restrict__ g, unsigned int words, int iters, int mode, unsigned long long* sink) {
  extern __shared__ int tile[];
  const bool use_smem = mode < 2;
  const bool random = (mode & 1) != 0;
  if (use_smem) {
    for (unsigned int i = threadIdx.x; i < words + 32u; i += blockDim.x) tile[i] = (int)(i * 2654435761u);
    __syncthreads();
  }
  // `words` is a power of two.  One LCG step yields a base; 8 loads then read 8 row segments (or
  // 8 scattered words per lane) at fixed strides from it, so address generation costs ~0.4
  // instructions per load and the load pipe is the limit.  Bases are 4-byte aligned, i.e. row
  // segments are misaligned with respect to 128-byte lines like the scoring kernel's.
  const unsigned int lane = threadIdx.x & 31;
  const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr unsigned int kStride = 1061u;                 // words between the 8 loads of a step (odd)
  const unsigned int mask = words / 2 - 1u;               // bases in the lower half; 8 strides stay inside
  unsigned int st = random ? mix(gtid * 9781u + 7u) : mix((gtid >> 5) * 9781u + 1u);
  const unsigned int add = random ? 0u : lane;
  unsigned int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  if (use_smem) {
#pragma unroll 1
    for (int it = 0; it < iters; it += 8) {
      st = st * 1664525u + 1013904223u;
      const int* p = tile + ((st >> 7) & mask) + add;
      acc0 += p[0] + p[kStride];
      acc1 += p[2 * kStride] + p[3 * kStride];
      acc2 += p[4 * kStride] + p[5 * kStride];
      acc3 += p[6 * kStride] + p[7 * kStride];
    }
  } else {
#pragma unroll 1
    for (int it = 0; it < iters; it += 8) {
      st = st * 1664525u + 1013904223u;
      const int* p = g + ((st >> 7) & mask) + add;
      acc0 += __ldg(p) + __ldg(p + kStride);
      acc1 += __ldg(p + 2 * kStride) + __ldg(p + 3 * kStride);
      acc2 += __ldg(p + 4 * kStride) + __ldg(p + 5 * kStride);
      acc3 += __ldg(p + 6 * kStride) + __ldg(p + 7 * kStride);
    }
  }
  const unsigned int acc = acc0 + acc1 + acc2 + acc3;
  if (acc == 0x12345u) atomicAdd(sink, 1ull);   // keeps the loads alive
}

cudaError_t launch_microbench(cudaStream_t st, int mode, const int* g, unsigned int words, int iters,
                              int n_cta, unsigned long long* sink) {
  const size_t smem = mode < 2 ? size_t(words + 32) * 4 : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  microbench_kernel<<<n_cta, 1024, smem, st>>>(g, words, iters, mode, sink);
  return cudaGetLastError();
}

cudaError_t launch_flush(cudaStream_t st, void* buf, long long bytes, int v) {
  flush_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<int4*>(buf), bytes / 16, v);
  return cudaGetLastError();
}


// =================================================================================================
// map check: rays of a scan through the publishing map (MapFeedbackResponsePenalty)
// =================================================================================================
// One CTA per pose, one thread per checked ray.  A ray is LineVisitor::ErgodLineBresenhami from the
// sensor cell to the beam's end cell (map/occu_grid_map.h:125-187); it counts once if any visited
// cell is occupied and farther than the tolerance from the end cell (:447-471 -- the callback only
// adds while the ray's result is below 1, so the walk stops at the first such cell).
__global__ void __launch_bounds__(128)
penalty_kernel(const PenaltyJob* __restrict__ jobs, const unsigned char* __restrict__ occ, int size_x, int size_y,
               double bound_tolerance) {
  const PenaltyJob J = jobs[blockIdx.x];
  int blocked = 0;
  for (int r = threadIdx.x; r * (long long)J.step < J.n_pts; r += blockDim.x) {
    const int p = r * J.step;
    const double px = J.pts[2 * p], py = J.pts[2 * p + 1];
    // pose_transform * point, + 0.5, truncate (:372-375); Affine * v = translation + (l00 x + l01 y)
    const int ex = __double2int_rz(dadd(dadd(J.tx, dadd(dmul(J.c, px), dmul(-J.s, py))), 0.5));
    const int ey = __double2int_rz(dadd(dadd(J.ty, dadd(dmul(J.s, px), dmul(J.c, py))), 0.5));
    if ((ex == J.sx0 && ey == J.sy0) || !(ex > 0 && ex < size_x && ey > 0 && ey < size_y)) continue;   // :377 (PointInMap is strict)
    int x0 = J.sx0, y0 = J.sy0, x1 = ex, y1 = ey;
    const bool steep = abs(y1 - y0) > abs(x1 - x0);
    if (steep) { int t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
    if (x0 > x1) { int t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
    const int delta_x = x1 - x0, delta_y = abs(y1 - y0), y_step = y0 < y1 ? 1 : -1;
    int error = 0, y = y0;
    for (int x = x0; x <= x1; ++x) {
      const int qx = steep ? y : x, qy = steep ? x : y;
      error += delta_y;
      if (2 * error >= delta_x) { y += y_step; error -= delta_x; }
      // (only a sensor cell outside the map can put a visited cell outside it; the reference reads
      //  out of bounds there, here such cells count as free)
      if (qx >= 0 && qx < size_x && qy >= 0 && qy < size_y && occ[(size_t)qy * size_x + qx]) {
        const int ddx = ex - qx, ddy = ey - qy;
        if (__dsqrt_rn((double)((long long)ddx * ddx + (long long)ddy * ddy)) > bound_tolerance) { ++blocked; break; }       // slam_util.h:94-96
      }
    }
  }
  blocked = __reduce_add_sync(0xffffffffu, blocked);
  __shared__ int s_sum[4];
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = blocked;
  __syncthreads();
  if (threadIdx.x == 0) *J.blocked = s_sum[0] + s_sum[1] + s_sum[2] + s_sum[3];
}

cudaError_t launch_penalty(int n_jobs, cudaStream_t st, const PenaltyJob* jobs, const unsigned char* occ, int size_x,
                           int size_y, double bound_tolerance) {
  penalty_kernel<<<n_jobs, 128, 0, st>>>(jobs, occ, size_x, size_y, bound_tolerance);
  return cudaGetLastError();
}

// =================================================================================================
// Gauss-Newton cost evaluation (BasedOptimizeScanMatch::UpdateCost, scan_match/optimize_scan_matcher.h:154-220)
// =================================================================================================
// One CTA per (grid, scan, pose estimate).  The reference adds the per-point terms of H = sum J^T J,
// b = sum -J^T e and cost = sum e^2 in point order, in FP64; a tree reduction would round differently.
// So the work is split by what can run in parallel without changing a bit: all threads compute the
// terms of a chunk of points (bilinear read of four cells, Jacobian, products -- every operation the
// reference's, none fused) into shared memory, then ten threads -- one per distinct sum; H is symmetric
// term by term because J_r*J_c is commutative -- add the chunk's terms in point order.  A point
// outside the map contributes +0.0 terms, which leaves every sum unchanged (the sums start at +0 and can
// never become -0).  The chain of ~P dependent adds is the latency floor of the kernel; the ten chains
// and all jobs of a batch run side by side.
constexpr int kOptChunk = 256;
__device__ __forceinline__ double opt_cell(const OptimizeJob& J, int x, int y) {
  // flat cell index y*size_x + x as the reference computes it (grid_map_base.h:352-354): x == size_x wraps into
  // the next row; past the last cell the reference reads out of bounds, here that reads 0
  if (x >= J.size_x) { x -= J.size_x; y += 1; }
  if (y >= J.size_y) return 0.0;
  const size_t at = (size_t)y * J.pitch + x;
  if (J.fixed) return dmul((double)(reinterpret_cast<const int*>(J.grid)[at]), 1.0 / 33554432.0);
  return (double)(reinterpret_cast<const float*>(J.grid)[at]);
}

__global__ void __launch_bounds__(128)
optimize_kernel(const OptimizeJob* __restrict__ jobs) {
  __shared__ double s_term[kOptSums][kOptChunk + 1];
  __shared__ int s_valid[4];
  const OptimizeJob J = jobs[blockIdx.x];
  double acc = 0.0;
  int valid = 0;
  for (int base = 0; base < J.n_pts; base += kOptChunk) {
    const int cnt = min(kOptChunk, J.n_pts - base);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const double lx = J.pts[2 * (base + i)], ly = J.pts[2 * (base + i) + 1];
      const double x = dadd(dadd(dmul(J.c, lx), dmul(-J.s, ly)), J.tx);     // rotation * local_point + translation (:167)
      const double y = dadd(dadd(dmul(J.s, lx), dmul(J.c, ly)), J.ty);
      double t[kOptSums];
#pragma unroll
      for (int k = 0; k < kOptSums; ++k) t[k] = 0.0;
      if (x > 0 && x < J.size_x && y > 0 && y < J.size_y) {                  // PointInMap (grid_map_base.h:330-337)
        ++valid;
        const double x0 = floor(x), y0 = floor(y), x1 = ceil(x), y1 = ceil(y);
        const double m00 = opt_cell(J, (int)x0, (int)y0), m01 = opt_cell(J, (int)x0, (int)y1);
        const double m10 = opt_cell(J, (int)x1, (int)y0), m11 = opt_cell(J, (int)x1, (int)y1);
        const double fx = dsub(x, x0), gx = dsub(x1, x), fy = dsub(y, y0), gy = dsub(y1, y);
        double r = dadd(dmul(fy, dadd(dmul(m11, fx), dmul(m01, gx))), dmul(gy, dadd(dmul(m10, fx), dmul(m00, gx))));   // :186-187
        r = (r >= 0) ? ((r <= 1) ? r : 1) : 0;                               // :190-191
        const double e = dsub(1.0, r);
        const double ds0 = dsub(dmul(-J.s, lx), dmul(J.c, ly));              // :197
        const double ds1 = dsub(dmul(J.c, lx), dmul(J.s, ly));               // :198
        const double dm0 = dadd(dmul(fy, dsub(m11, m01)), dmul(gy, dsub(m10, m00)));   // :200
        const double dm1 = dadd(dmul(fx, dsub(m11, m10)), dmul(gx, dsub(m01, m00)));   // :201
        const double j0 = dadd(dmul(-dm0, 1.0), dmul(-dm1, 0.0));            // J = -de_m * de_s (:204)
        const double j1 = dadd(dmul(-dm0, 0.0), dmul(-dm1, 1.0));
        const double j2 = dadd(dmul(-dm0, ds0), dmul(-dm1, ds1));
        t[0] = dmul(j0, j0); t[1] = dmul(j1, j0); t[2] = dmul(j2, j0);       // :206
        t[3] = dmul(j1, j1); t[4] = dmul(j2, j1); t[5] = dmul(j2, j2);
        t[6] = dmul(-j0, e); t[7] = dmul(-j1, e); t[8] = dmul(-j2, e);       // :207
        t[9] = dmul(e, e);                                                   // :193
      }
#pragma unroll
      for (int k = 0; k < kOptSums; ++k) s_term[k][i] = t[k];
    }
    __syncthreads();
    if (threadIdx.x < kOptSums) {
      const double* row = s_term[threadIdx.x];
#pragma unroll 8
      for (int i = 0; i < cnt; ++i) acc = dadd(acc, row[i]);
    }
    __syncthreads();
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0) s_valid[threadIdx.x >> 5] = valid;
  __syncthreads();
  if (threadIdx.x < kOptSums) J.out[threadIdx.x] = acc;
  if (threadIdx.x == 0) J.out[kOptSums] = (double)(s_valid[0] + s_valid[1] + s_valid[2] + s_valid[3]);
}

cudaError_t launch_optimize(int n_jobs, cudaStream_t st, const OptimizeJob* jobs) {
  if (n_jobs == 0) return cudaSuccess;
  optimize_kernel<<<n_jobs, 128, 0, st>>>(jobs);
  return cudaGetLastError();
}

// =================================================================================================
// publishing map (OccuGridMap<CountCell>::UpdateMapByRange, map/occu_grid_map.h:258-329, 474-530)
// =================================================================================================
// Within one update every cell changes at most once per kind: SetCellFree acts only while update_index_ <
// cur_mark_free_index, SetCellOccu only while update_index_ < cur_mark_occu_index, and an occupied cell's own ray
// has marked it free first (the Bresenham walk ends on the end cell).  So a cell's fate depends only on whether
// any ray passed it (free) and whether any ray ended on it (occupied), not on the beam order: pass 1 records
// that with atomicMax on the update index exactly as the reference stores it; pass 2 applies the float updates
// the reference would have made, in its order: free; or free, un-free, occupied.
__global__ void __launch_bounds__(128) pub_mark_kernel(PubScan* __restrict__ scan) {
  const PubScan S = *scan;
  int bx0 = S.start_x, bx1 = S.start_x, by0 = S.start_y, by1 = S.start_y;   // hull of the start cell and the end cells
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n_pts; i += gridDim.x * blockDim.x) {
  const double px = S.pts[2 * i], py = S.pts[2 * i + 1];
  const int ex = __double2int_rz(dadd(dadd(S.tx, dadd(dmul(S.c, px), dmul(-S.s, py))), 0.5));   // :301-310
  const int ey = __double2int_rz(dadd(dadd(S.ty, dadd(dmul(S.s, px), dmul(S.c, py))), 0.5));
  if (ex == S.start_x && ey == S.start_y) continue;                                                // :312
  bx0 = min(bx0, ex); bx1 = max(bx1, ex); by0 = min(by0, ey); by1 = max(by1, ey);
  // LineVisitor::ErgodLineBresenhami (:125-187)
  int x0 = S.start_x, y0 = S.start_y, x1 = ex, y1 = ey;
  const bool steep = abs(y1 - y0) > abs(x1 - x0);
  if (steep) { int t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
  if (x0 > x1) { int t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
  const int delta_x = x1 - x0, delta_y = abs(y1 - y0), y_step = y0 < y1 ? 1 : -1;
  int error = 0, y = y0;
  for (int x = x0; x <= x1; ++x) {
    const int qx = steep ? y : x, qy = steep ? x : y;
    error += delta_y;
    if (2 * error >= delta_x) { y += y_step; error -= delta_x; }
    // CellUpdate: PointInMap(x, y, half_kernel_size_ + 1) with half_kernel_size_ = 0 (:476, grid_map_base.h:339-346)
    if (qx > 1 && qx < S.size_x - 1 && qy > 1 && qy < S.size_y - 1) atomicMax(S.mark + (size_t)qy * S.size_x + qx, S.free_tag);
  }
  if (ex > 1 && ex < S.size_x - 1 && ey > 1 && ey < S.size_y - 1) atomicMax(S.mark + (size_t)ey * S.size_x + ex, S.occ_tag);
  }
  // the cells this update can have touched, for the apply pass (every ray stays inside the hull of its two ends)
  bx0 = __reduce_min_sync(0xffffffffu, bx0); bx1 = __reduce_max_sync(0xffffffffu, bx1);
  by0 = __reduce_min_sync(0xffffffffu, by0); by1 = __reduce_max_sync(0xffffffffu, by1);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&scan->bx0, max(bx0, 0)); atomicMax(&scan->bx1, min(bx1, S.size_x - 1));
    atomicMin(&scan->by0, max(by0, 0)); atomicMax(&scan->by1, min(by1, S.size_y - 1));
  }
}

__global__ void __launch_bounds__(256) pub_apply_kernel(const PubScan* __restrict__ scan) {
  const PubScan S = *scan;
  const int w = S.bx1 - S.bx0 + 1, h = S.by1 - S.by0 + 1;
  if (w <= 0 || h <= 0) return;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < (long long)w * h; k += (long long)gridDim.x * blockDim.x) {
    const int x = S.bx0 + (int)(k % w), y = S.by0 + (int)(k / w);
    const size_t c = (size_t)y * S.size_x + x;
    const int m = S.mark[c];
    if (m == S.free_tag) {                              // CountCellFunctions::UpdateSetFree (grid_map_cell.h:100-103)
      const float pass = __fadd_rn(S.pass[c], S.add_pass);
      S.pass[c] = pass;
      S.prob[c] = __fdiv_rn(S.hit[c], pass);
    } else if (m == S.occ_tag) {                        // UpdateSetFree, UpdateUnsetFree, UpdateSetOccupied (:92-108)
      float pass = __fadd_rn(S.pass[c], S.add_pass);
      pass = __fsub_rn(pass, S.add_pass);
      const float hit = __fadd_rn(S.hit[c], S.add_hit);
      pass = __fadd_rn(pass, S.add_pass);
      float prob = __fdiv_rn(hit, pass);
      if (prob > 1.0f) prob = 1.0f;
      S.hit[c] = hit; S.pass[c] = pass; S.prob[c] = prob;
    }
  }
}

cudaError_t launch_pub_update(cudaStream_t st, PubScan* scan) {
  // `scan` is the device copy, its box initialised empty (bx0 = by0 = INT_MAX, bx1 = by1 = -1); the launch shapes
  // are fixed upper bounds, the kernels stride over n_pts / the box themselves
  pub_mark_kernel<<<16, 128, 0, st>>>(scan);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  pub_apply_kernel<<<296, 256, 0, st>>>(scan);
  return cudaGetLastError();
}

// CountCellFunctions::GetGridStates == GridStates_Occupied (grid_map_cell.h:125-136): what the map check reads
__global__ void __launch_bounds__(256)
pub_occupancy_kernel(const float* __restrict__ pass, const float* __restrict__ prob, long long n, float thr, float min_pass,
                     unsigned char* __restrict__ occ) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    occ[i] = (pass[i] >= min_pass && !(prob[i] < thr)) ? 1 : 0;
}

cudaError_t launch_pub_occupancy(cudaStream_t st, const float* pass, const float* prob, long long n_cells, float occu_threshold,
                                 float min_pass_through, unsigned char* occ) {
  pub_occupancy_kernel<<<296, 256, 0, st>>>(pass, prob, n_cells, occu_threshold, min_pass_through, occ);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) fill_f32_kernel(float* p, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
cudaError_t launch_fill_f32(cudaStream_t st, float* p, long long n, float v) {
  fill_f32_kernel<<<296, 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}

}  // namespace rsm

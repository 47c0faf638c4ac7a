// rsm_select.cuh -- device code shared by the kernels of rsm_kernels.cu and rsm_score.cu: exact FP64 helpers, the
// order-preserving score key, the block-wide top-kTopK of the selection kernel and its per-job tail (merge of the per-CTA
// tops, speculative gather of the best candidate's columns).
#ifndef RSM_SELECT_CUH_
#define RSM_SELECT_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "rsm_device.h"

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
// selection
// =================================================================================================
#ifdef RSM_SELECT_DEBUG
static __device__ unsigned long long g_sel_dbg[16];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define SDBG(i) do { if (threadIdx.x == 0) g_sel_dbg[i] = gtime(); } while (0)
#define SDBG_MIN(i) do { if (threadIdx.x == 0) atomicMin(&g_sel_dbg[i], gtime()); } while (0)
#define SDBG_MAX(i) do { if (threadIdx.x == 0) atomicMax(&g_sel_dbg[i], gtime()); } while (0)
#define SDBG_DUR(i, t0) do { if (threadIdx.x == 0) atomicMax(&g_sel_dbg[i], gtime() - (t0)); } while (0)
#define SDBG_T() gtime()
#else
#define SDBG(i) do {} while (0)
#define SDBG_MIN(i) do {} while (0)
#define SDBG_MAX(i) do {} while (0)
#define SDBG_DUR(i, t0) do {} while (0)
#define SDBG_T() 0ull
#endif
// Second half of the block-wide top-kTopK: s_count[0] keys above the threshold wait in s_bkey / s_bidx[0 ..), up to kTopK keys
// tied with it in the buffer's tail (s_count[1]); sorts them and calls emit(r, key, index) for the first kTopK.  The caller
// has synchronised the block after the last buffer write.
template <typename Emit>
__device__ int top_k_finish(Emit emit, unsigned long long* s_bkey, unsigned int* s_bidx, int* s_count, bool* overflow) {
  const int tid = threadIdx.x, NT = blockDim.x;
  int count = s_count[0];
  const int n_eq = min(s_count[1], kTopK);
  if (count > kSelectBuf - kTopK) { *overflow = true; count = kSelectBuf - kTopK; }
  __syncthreads();
  // compact: move the tied keys right behind the others
  unsigned long long kk = 0ull;
  unsigned int ii = 0u;
  if (tid < n_eq) { kk = s_bkey[kSelectBuf - kTopK + tid]; ii = s_bidx[kSelectBuf - kTopK + tid]; }
  __syncthreads();
  if (tid < n_eq) { s_bkey[count + tid] = kk; s_bidx[count + tid] = ii; }
  __syncthreads();
  count += n_eq;
  int pow2 = 32;
  while (pow2 < count) pow2 <<= 1;
  if (pow2 <= NT) {      // (every lane of the network needs its partner: a CTA of 416 threads sorts up to 256 pairs this way)
    // bitonic sort of the buffered (key, index) pairs, one per thread in registers: descending key,
    // ascending index among equal keys; the first kTopK are the answer.  Exchanges at distances
    // below 32 are warp shuffles, the others go through shared memory.
    const int N = pow2;
    unsigned long long k = tid < count ? s_bkey[tid] : 0ull;
    unsigned int idx = tid < count ? s_bidx[tid] : 0xffffffffu;
    __syncthreads();
    for (int kk = 2; kk <= N; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        unsigned long long ok;
        unsigned int oi;
        if (j >= 32) {
          s_bkey[tid] = k; s_bidx[tid] = idx;
          __syncthreads();
          ok = s_bkey[tid ^ j]; oi = s_bidx[tid ^ j];
          __syncthreads();
        } else {
          ok = __shfl_xor_sync(0xffffffffu, k, j);
          oi = __shfl_xor_sync(0xffffffffu, idx, j);
        }
        const bool other_first = ok > k || (ok == k && oi < idx);   // the partner's pair sorts before this one
        const bool lower = (tid & j) == 0, descending_run = (tid & kk) == 0;
        if ((lower == descending_run) ? other_first : !other_first) { k = ok; idx = oi; }
      }
    }
    if (tid < min(count, kTopK)) emit(tid, k, idx);
    __syncthreads();
      return min(count, kTopK);
  }
  // (massive ties only) rank of every buffered key among the buffered keys; ranks < kTopK are the answer
  for (int i = tid; i < count; i += NT) {
    const unsigned long long ki = s_bkey[i];
    const unsigned int ei = s_bidx[i];
    int rank = 0;
    for (int j = 0; j < count; j += 8) {      // 8 independent shared-memory loads in flight
      unsigned long long kj[8];
      unsigned int ej[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool in = j + u < count;
        kj[u] = in ? s_bkey[j + u] : 0ull;
        ej[u] = in ? s_bidx[j + u] : 0xffffffffu;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) rank += (kj[u] > ki || (kj[u] == ki && ej[u] < ei)) ? 1 : 0;
      if (rank >= kTopK) break;               // already out of the top list
    }
    if (rank < kTopK) emit(rank, ki, ei);
  }
  __syncthreads();
#ifdef RSM_SELECT_DEBUG
  if (threadIdx.x == 0) atomicMax(&g_sel_dbg[11], (unsigned long long)count);
#endif
  return min(count, kTopK);
}

// Block-wide top-kTopK of n keys given by key_at(i), i in [0, n): emit(r, key, i) is called for
// r = 0.. in descending key order (smallest i first among equal keys), each rank by one thread.
// Two streaming passes over the keys (visit(i, key) is called once per element during the
// first), then a rank computation over the few keys that reached the threshold.  Returns the
// number emitted; *overflow is set when more than kSelectBuf keys reach the threshold (massive ties).
template <typename KeyAt, typename Visit, typename Emit>
__device__ int block_top_k(unsigned int n, KeyAt key_at, Visit visit, Emit emit, unsigned long long* s_k,
                           unsigned long long* s_bkey, unsigned int* s_bidx, int* s_count, bool* overflow,
                           unsigned long long* s_cache, unsigned int cache_n) {
  const int tid = threadIdx.x, NT = blockDim.x;
  constexpr int U = 8;   // independent loads in flight per thread: these passes are pure L2 latency otherwise
  unsigned long long tmax = 0ull;
  const unsigned long long d0 = SDBG_T();
  for (unsigned int i0 = tid; i0 < n; i0 += U * NT) {
    unsigned long long k[U];
#pragma unroll
    for (int u = 0; u < U; ++u) k[u] = (i0 + u * NT < n) ? key_at(i0 + u * NT) : 0ull;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * NT < n) {
        if (i0 < cache_n) s_cache[i0 + u * NT] = k[u];   // read back by this same thread in the second pass
        visit(i0 + u * NT, k[u]);
        tmax = k[u] > tmax ? k[u] : tmax;
      }
    }
  }
  // threshold: every warp ranks the high words of its lanes' maxima with shuffles and reports the
  // kTopK-th largest; the largest report is the threshold.  The lanes' maxima are distinct
  // elements, so at least kTopK keys are at or above any warp's report (low words cleared): a
  // lower bound of the kTopK-th largest key, tight enough that only a few dozen keys pass.  Warps
  // with fewer than kTopK non-empty lanes report 0; if all do, everything is kept (n is small).
  unsigned long long thr = 0ull;
  __syncthreads();
  SDBG_DUR(8, d0);
  {
    const unsigned int lane = tid & 31;
    const unsigned int h = (unsigned int)(tmax >> 32);
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const unsigned int hj = __shfl_sync(0xffffffffu, h, j);
      rank += (hj > h || (hj == h && j < (int)lane)) ? 1 : 0;
    }
    __syncthreads();   // previous users of s_k / s_count / the buffer are done
    if (rank == kTopK - 1) s_k[tid >> 5] = (unsigned long long)h << 32;
    if (tid == 0) { s_count[0] = 0; s_count[1] = 0; }
    __syncthreads();
    for (int w = 0; w < (NT >> 5); ++w) thr = s_k[w] > thr ? s_k[w] : thr;
  }
  for (unsigned int i0 = tid; i0 < n; i0 += U * NT) {
    unsigned long long k[U];
    if (i0 < cache_n) {      // cache_n is a multiple of U * NT: an iteration is cached as a whole or not at all
#pragma unroll
      for (int u = 0; u < U; ++u) k[u] = (i0 + u * NT < n) ? s_cache[i0 + u * NT] : 0ull;
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) k[u] = (i0 + u * NT < n) ? key_at(i0 + u * NT) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (k[u] > thr) {
        const int pos = atomicAdd(s_count, 1);
        if (pos < kSelectBuf - kTopK) { s_bkey[pos] = k[u]; s_bidx[pos] = i0 + u * NT; }
      } else if (k[u] == thr && k[u] != 0ull) {
        // keys tied with the threshold: kTopK of them are enough (score fields on maps with few
        // distinct cell values tie massively), kept in the tail of the buffer
        const int pos = atomicAdd(s_count + 1, 1);
        if (pos < kTopK) { s_bkey[kSelectBuf - kTopK + pos] = k[u]; s_bidx[kSelectBuf - kTopK + pos] = i0 + u * NT; }
      }
    }
  }
  __syncthreads();
  SDBG_DUR(9, d0);
  return top_k_finish(emit, s_bkey, s_bidx, s_count, overflow);
}

// Speculative gather of the best candidate's 3x3 translation neighbourhood over all angles: the same-(x,y) columns the
// angular covariance needs whenever the averaged best pose stays in that neighbourhood (always when the averaging set has
// one member).  J.final_top[0] is the job's best candidate (n_final > 0), written before a block barrier.
__device__ __forceinline__ void spec_gather(const SelectJob& J, int n_final) {
  const int tid = threadIdx.x, NT = blockDim.x;
  // ---- speculative gather of the best candidate's 3x3 translation neighbourhood -------------------
  if (n_final > 0 && J.spec_out != nullptr) {
    const long long kbest = J.final_top[0].index;       // written by thread 0 above, visible after the barrier
    const int n_xy = J.n_xy;
    const long long plane = (long long)n_xy * n_xy;
    const int rem = (int)(kbest % plane);
    const int bx = rem / n_xy, by = rem % n_xy;
    const int x0 = max(bx - 1, 0), x1 = min(bx + 1, n_xy - 1), y0 = max(by - 1, 0), y1 = min(by + 1, n_xy - 1);
    const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, ncol = nx * ny;
    if (tid == 0) J.spec_cols[0] = ncol;
    if (tid < ncol) J.spec_cols[1 + tid] = (x0 + tid / ny) * n_xy + (y0 + tid % ny);
    const int nang = J.n_ang;
    for (int i0 = tid; i0 < ncol * nang; i0 += 4 * NT) {      // 4 independent loads in flight
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * NT;
        if (i < ncol * nang) {
          const int c = i / nang, ia = i - c * nang;
          const int col = (x0 + c / ny) * n_xy + (y0 + c % ny);
          v[u] = J.score[(long long)ia * plane + col];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * NT < ncol * nang) J.spec_out[i0 + u * NT] = v[u];
    }
  } else if (tid == 0 && J.spec_cols != nullptr) {
    J.spec_cols[0] = 0;
  }
  __syncthreads();
  SDBG(6);
}

// The per-job tail, run by the last CTA to finish one of the job's lists: merge the n_lists per-list tops
// (J.top_list, kTopK entries each, unused ones with key 0) into the job's final top-kTopK, then gather, speculatively,
// the scores of the 3x3 translation neighbourhood of the best candidate over all angles: those are the same-(x,y)
// columns the angular covariance needs whenever the averaged best pose stays in that neighbourhood.  own_count: entries
// of the caller's own list when it is the only one.  s_cache: kSelectSlice keys or nullptr (cache_n 0).
__device__ __forceinline__ void select_job_tail(const SelectJob& J, int n_lists, int own_count, unsigned long long* s_k,
                                                unsigned long long* s_bkey, unsigned int* s_bidx, int* s_count,
                                                unsigned long long* s_cache, unsigned int cache_n) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const unsigned int n_ent = (unsigned int)n_lists * kTopK;
  const Entry* ent = J.top_list;
  bool overflow = false;
  int n_final;
  if (n_lists == 1) {
    // a single list (every small pass): it already is the job's top-kTopK
    n_final = own_count;
    for (int r = tid; r < own_count; r += NT) J.final_top[r] = ent[r];
  } else {
    n_final = block_top_k(
        n_ent,
        [&](unsigned int i) -> unsigned long long { return score_key(ent[i].score); },
        [](unsigned int, unsigned long long) {},
        [&](int r, unsigned long long, unsigned int i) { J.final_top[r] = ent[i]; },
        s_k, s_bkey, s_bidx, s_count, &overflow, s_cache, cache_n);
  }
  if (tid == 0) {
    *J.final_count = n_final;
    if (overflow) atomicOr(J.err, kErrSelectFull);
  }
  __syncthreads();
  SDBG(5);
  spec_gather(J, n_final);
}

}  // namespace rsm
#endif

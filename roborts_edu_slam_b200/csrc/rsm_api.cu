// rsm_api.cu -- the C ABI of include/rsm.h: context and grid management, the host side of a pass
// (geometry, libm angle tables, launch, exact finalisation) and the chain / batch drivers.
//
// Host/device split of one pass (reference: BasedCorrelationScanMatch::ScanMatch,
// scan_match/correlate_scan_matcher.h:784-875):
//   host    world->map, n_ang / n_xy / stride, cos/sin of every search angle from the HOST libm
//           (CUDA's sin/cos are not bit-identical to glibc's; n_ang is at most ~10^3 per pass)
//   device  score_kernel  -> penalised score of every candidate + running maximum
//           select_kernel -> averaging-set candidates + global top-21
//           gather_kernel -> scores of the same-(x,y) columns (angular covariance only)
//   host    FindBestCandidate / covariance on the few dozen candidates that matter, in the
//           reference's expression order; atan2 from the host libm.
// If the candidates that are consumed contain exact score ties whose order libstdc++'s unstable
// sort would decide, the pass falls back to the *exact* path: the score array is copied back and
// sorted with the same std::sort call the reference makes, which reproduces its order.
// There is no CPU scoring path anywhere in this file.
#include <cuda.h>           // CUtensorMap types only; the encoder is fetched through the runtime
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>        // environ

#include <algorithm>
#include <map>
#include <memory>
#include <tuple>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rsm.h"
#include "rsm_device.h"
#include "rsm_host.h"
#include <climits>

#include "rsm_kernels.h"

using namespace rsm;

// ---------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------
struct rsm_grid {
  int size_x = 0, size_y = 0, pitch = 0;
  double resolution = 0, scale = 0, off_x = 0, off_y = 0;
  MapTransform tf;
  bool fixed = true;   // cells are int32 value*2^25 (else float32)
  bool init = false;   // has content (reference: IsMapInit())
  bool owned = true;   // d_cells allocated by this grid (false: a slot of a batch pool)
  void* d_cells = nullptr;
  unsigned char* d_occ = nullptr;   // publishing-map occupancy mask (rsm_grid_upload_occupancy), size_x bytes per row
  double cell_len() const { return 1 / scale; }   // map/grid_map_base.h:307-309
};

struct rsm_scan {
  double* d_pts = nullptr;
  int n = 0;
};

// Device-resident counterpart of SensorDataManager::multiresolution_range_data_[name]
// (slam/sensor_data_manager.h:514-525): scans addressed by id, each with its current sensor pose.
struct rsm_scan_store {
  struct Entry { double* d_pts; int n; double pose[3]; };
  std::vector<Entry> scans;
  std::vector<void*> chunks;          // device allocations; scans never move once stored
  size_t chunk_used = 0, chunk_cap = 0;
};

// GridMapBase's bound-box bookkeeping (host only): see MapBounds in rsm_host.h
struct rsm_map_bounds { MapBounds b; };

// Device-resident publishing map: OccuGridMap<CountCell> (map/slam_map.h:35) as three float planes + the update index.
struct rsm_pubmap {
  rsm_grid g;                      // geometry + the occupancy mask the map check reads (g.d_occ); no lookup cells
  float* d_hit = nullptr;
  float* d_pass = nullptr;
  float* d_prob = nullptr;
  int* d_mark = nullptr;
  float default_prob = 0.5f;
  int cur_update_index = 0;        // occu_grid_map.h:207, advanced by 3 per update
};

namespace {

struct Buf {
  char* p = nullptr;
  size_t cap = 0;
};

struct Layout {
  size_t off = 0;
  size_t take(size_t bytes, size_t align = 256) {
    off = (off + align - 1) / align * align;
    size_t o = off;
    off += bytes;
    return o;
  }
};

enum KernelClass { KC_SCORE = 0, KC_SELECT = 1, KC_RASTER = 2, KC_OTHER = 3, KC_N = 4 };

}  // namespace

namespace {
// Small persistent worker pool for the per-item host work of large batches (angle tables,
// FindBestCandidate, covariances): items are independent, so the range is cut into chunks that
// the workers and the calling thread pull from an atomic counter.
// set while a thread runs one lane of a pipelined chain: per-item loops then run inline (the lanes are the parallelism)
thread_local bool t_lane_worker = false;

class HostPool {
 public:
  explicit HostPool(int fixed_threads = 0) : fixed_(fixed_threads) {}
  ~HostPool() { stop(); }
  void set_threads(int n) { if (n != fixed_ && workers_.empty()) fixed_ = n; }
  void run(int n, int grain, const std::function<void(int, int)>& fn) {
    if (n <= 0) return;
    const int chunks = (n + grain - 1) / grain;
    if (chunks <= 1 || threads_wanted() <= 1 || t_lane_worker) { fn(0, n); return; }
    start();
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn; n_ = n; grain_ = grain; next_.store(0); pending_ = int(workers_.size()); ++gen_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }
 private:
  int threads_wanted() const { return fixed_ > 0 ? fixed_ : env_threads(); }
 public:
  // RSM_HOST_THREADS caps the pool (one process per GPU shares the host cores with its peers)
  static int env_threads() {
    static const int n = [] {
      unsigned hc = std::thread::hardware_concurrency();
      int t = int(std::min<unsigned>(hc ? hc : 1, 16));
      if (const char* e = std::getenv("RSM_HOST_THREADS")) { const int v = std::atoi(e); if (v >= 1) t = std::min(v, 64); }
      return t;
    }();
    return n;
  }
 private:
  void work() {
    for (;;) {
      const int c = next_.fetch_add(1);
      const int b = c * grain_;
      if (b >= n_) break;
      (*fn_)(b, std::min(n_, b + grain_));
    }
  }
  void start() {
    if (!workers_.empty()) return;
    const int t = threads_wanted() - 1;
    for (int i = 0; i < t; ++i)
      workers_.emplace_back([this] {
        unsigned seen = 0;
        for (;;) {
          {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
            if (quit_) return;
            seen = gen_;
          }
          work();
          {
            std::lock_guard<std::mutex> lk(m_);
            if (--pending_ == 0) done_cv_.notify_one();
          }
        }
      });
  }
  void stop() {
    { std::lock_guard<std::mutex> lk(m_); quit_ = true; }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
    workers_.clear();
  }
  int fixed_ = 0;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int, int)>* fn_ = nullptr;
  int n_ = 0, grain_ = 1, pending_ = 0;
  unsigned gen_ = 0;
  bool quit_ = false;
  std::atomic<int> next_{0};
};
}  // namespace

struct SliceState {   // what rsm_match_partial leaves behind for merge / finish
  bool valid = false, merged = false, exact_needed = false;
  rsm_pass_param param;
  PassGeo geo;
  const rsm_grid* grid = nullptr;
  int a0 = 0, a1 = 0;
  double* d_score = nullptr;     // slice scores, resident in the work arena
  char* d_gjob = nullptr; double* d_gout = nullptr;
  BestPose best;
  std::vector<Cand> top;
  int n_cols = 0;
  int cols[kMaxCols];
};

// One in-flight pass: a stream and the buffers its launches read and write.  Lane 0 is the context's own
// stream; a batched chain call cuts its items into sub-batches that run on lanes 0..L-1, so that the host
// finalisation of one sub-batch overlaps the kernels of the others (one context per GPU is enough).
struct Lane {
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;                 // second staged scoring launch, concurrent with the first
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t done = nullptr;                     // blocking-sync event: a waiting host thread sleeps instead of spinning
  Buf d_work, h_up, h_down;                       // device arena, pinned staging (up / down)
  Buf d_aux, h_aux;                               // raster descriptors of a loop-closure sub-batch (live while a pass is prepared)
  std::vector<cudaEvent_t> ev_pool;
  struct Span { cudaEvent_t a, b; int kc; };
  std::vector<Span> spans;
  size_t ev_used = 0;
  rsm_stats stats = {};                           // what this lane's passes counted since the last merge into the context's
  // CUDA graphs of whole passes (upload, zeroing, score launches, select, read-back) enqueued on this lane, keyed by
  // everything that shapes the launch sequence; the per-call data travels in the pinned buffer.  Per lane: lane threads
  // capture and replay side by side.
  struct PassGraph { int seen = 0; cudaGraphExec_t exec = nullptr; };
  std::map<std::vector<long long>, PassGraph> graphs;
};

// stream plan of the staged scoring kernel (plan_stream below)
struct StreamRun { int V, n_xy, ang_count, tiles_x, tiles_y; };
struct StreamPlan { std::vector<StreamCta> ctas; std::vector<int> item_begin; int n_tickets = 0, n_slots = 0; long long n_items = 0; };

struct rsm_ctx {
  int device = 0;
  HostPool pool;
  HostPool lane_pool{1};                          // one thread per lane of a pipelined chain (resized before first use)
  std::mutex mu;                                  // err string, tensor-map cache: touched by lane threads
  SliceState slice;
  Lane L0;
  std::vector<Lane*> extra_lanes;                 // lanes 1.. (created on first use)
  cudaStream_t& stream = L0.stream;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  std::string err;
  rsm_stats stats;
  bool profiling = false;
  bool strict_ties = false;                       // RSM_OPT_STRICT_TIES
  int lanes_wanted = 0;                           // RSM_OPT_LANES (0 = automatic)
  Buf& d_work = L0.d_work;
  Buf& h_up = L0.h_up;
  Buf& h_down = L0.h_down;
  Buf d_pts, d_flush, d_pool_grids, d_pool_grids2;
  // in-library exchange of the angle-sliced match (rsm_comm_init / rsm_match_sliced): an NCCL communicator of this
  // context's own, exchange buffers on the device and their pinned host mirror
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  Buf d_xchg, h_xchg;
  // staged scoring variant: tensor maps of the grids seen so far, keyed by (cells, size, pitch)
  struct alignas(64) TmapPair { CUtensorMap box[2]; };
  std::map<std::tuple<const void*, int, int, int, int>, TmapPair> tmaps;   // + box set
  void* encode_tiled = nullptr;   // cuTensorMapEncodeTiled
  int sm_count = 0;               // multiprocessors of the device (stream plan: persistent CTAs)
  std::map<std::vector<int>, std::shared_ptr<const StreamPlan>> stream_plans;   // by (beams, window, angles) per job: a front end repeats its shape
};

namespace {

struct PhaseTimer {   // host wall clock, accumulated into rsm_stats::phase_ms
  rsm_stats* st;
  std::chrono::steady_clock::time_point t;
  explicit PhaseTimer(rsm_stats* s) : st(s), t(std::chrono::steady_clock::now()) {}
  void lap(int phase) {
    auto n = std::chrono::steady_clock::now();
    st->phase_ms[phase] += std::chrono::duration<double, std::milli>(n - t).count();
    t = n;
  }
};

// lane counters -> context counters (by the calling thread, when no lane thread is running)
void merge_lane_stats(rsm_ctx* ctx, Lane* L) {
  rsm_stats& a = ctx->stats;
  rsm_stats& b = L->stats;
  a.kernel_launches += b.kernel_launches; a.score_launches += b.score_launches; a.evals += b.evals; a.passes += b.passes;
  a.exact_sort_passes += b.exact_sort_passes; a.h2d_bytes += b.h2d_bytes; a.d2h_bytes += b.d2h_bytes;
  a.score_kernel_ms += b.score_kernel_ms; a.raster_kernel_ms += b.raster_kernel_ms; a.select_kernel_ms += b.select_kernel_ms;
  for (int i = 0; i < 8; ++i) a.phase_ms[i] += b.phase_ms[i];
  std::memset(&b, 0, sizeof b);
}

int fail(rsm_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err = buf;
  }
  return code;
}

// Every entry point makes the context's device current for the calling thread (a new host thread
// starts on device 0) and puts the caller's device back on return.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(const rsm_ctx* ctx) {
    if (!ctx) return;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != ctx->device) cudaSetDevice(ctx->device); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Tensor maps (wide box, tall box) over a fixed-point grid, for the TMA box copies of the staged
// scoring variant.  Out-of-grid parts of a box are filled with zeros.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The pair is copied out under the context's mutex: lane threads share the cache.
// box_set: -1 = the cluster kernel's boxes, >= 0 = those of that stream-plan variant.
int grid_tmaps(rsm_ctx* ctx, const rsm_grid* g, int box_set, rsm_ctx::TmapPair* out) {
  const auto key = std::make_tuple((const void*)g->d_cells, g->size_x, g->size_y, g->pitch, box_set);
  const char* what = nullptr;
  int code = 0;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->tmaps.find(key);
    if (it != ctx->tmaps.end()) { *out = it->second; return RSM_OK; }
    if (!ctx->encode_tiled) {
      cudaDriverEntryPointQueryResult q;
      void* fn = nullptr;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
          q != cudaDriverEntryPointSuccess)
        what = "cuTensorMapEncodeTiled is not available from this driver";
      else
        ctx->encode_tiled = fn;
    }
    if (!what) {
      if (ctx->tmaps.size() > 8192) ctx->tmaps.clear();
      rsm_ctx::TmapPair pair;
      int bw[2], bh[2];
      if (box_set < 0) score_staged_boxes(bw, bh); else score_stream_boxes(box_set, bw, bh);
      for (int b = 0; b < 2 && !what; ++b) {
        const cuuint64_t dims[2] = {cuuint64_t(g->size_x), cuuint64_t(g->size_y)};
        const cuuint64_t strides[1] = {cuuint64_t(g->pitch) * 4};
        const cuuint32_t box[2] = {cuuint32_t(bw[b]), cuuint32_t(bh[b])};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
            &pair.box[b], CU_TENSOR_MAP_DATA_TYPE_INT32, 2, g->d_cells, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { what = "cuTensorMapEncodeTiled failed"; code = int(r); }
      }
      if (!what) { ctx->tmaps.emplace(key, pair); *out = pair; return RSM_OK; }
    }
  }
  return fail(ctx, RSM_ERR_CUDA, "%s (%d)", what, code);
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(ctx, RSM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int ensure_dev(rsm_ctx* ctx, Buf& b, size_t bytes, cudaStream_t st = nullptr) {
  if (bytes <= b.cap) return RSM_OK;
  CU(cudaStreamSynchronize(st ? st : ctx->stream));
  if (b.p) CU(cudaFree(b.p));
  b.p = nullptr; b.cap = 0;
  size_t cap = bytes + bytes / 4 + 4096;
  CU(cudaMalloc(reinterpret_cast<void**>(&b.p), cap));
  b.cap = cap;
  return RSM_OK;
}

int ensure_pinned(rsm_ctx* ctx, Buf& b, size_t bytes, cudaStream_t st = nullptr) {
  if (bytes <= b.cap) return RSM_OK;
  CU(cudaStreamSynchronize(st ? st : ctx->stream));
  if (b.p) CU(cudaFreeHost(b.p));
  b.p = nullptr; b.cap = 0;
  size_t cap = bytes + bytes / 4 + 4096;
  CU(cudaMallocHost(reinterpret_cast<void**>(&b.p), cap));
  b.cap = cap;
  return RSM_OK;
}

// Lanes beyond the context's own stream, created on first use.
int get_lane(rsm_ctx* ctx, int k, Lane** out) {
  if (k == 0) { *out = &ctx->L0; return RSM_OK; }
  while (int(ctx->extra_lanes.size()) < k) {
    Lane* L = new Lane;
    // Lanes get distinct stream priorities (lane 1 the highest, lane 0 -- the context's own stream -- the lowest).
    // Kernels of equal priority share the SMs evenly, so the sub-batches of a chain would all finish a pass at the
    // same moment, finalise on the host at the same moment and leave the GPU idle meanwhile; with priorities the
    // lanes finish one after the other and one lane's host stage overlaps the others' kernels.
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    const int prio = std::min(least, greatest + int(ctx->extra_lanes.size()));
    const bool no_prio = std::getenv("RSM_NO_LANE_PRIORITY") != nullptr;
    if ((no_prio ? cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking)
                 : cudaStreamCreateWithPriority(&L->stream, cudaStreamNonBlocking, prio)) != cudaSuccess ||
        (no_prio ? cudaStreamCreateWithFlags(&L->stream2, cudaStreamNonBlocking)
                 : cudaStreamCreateWithPriority(&L->stream2, cudaStreamNonBlocking, prio)) != cudaSuccess ||
        cudaEventCreateWithFlags(&L->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&L->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&L->done, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) {
      delete L;
      return fail(ctx, RSM_ERR_CUDA, "creating a pipeline lane failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    ctx->extra_lanes.push_back(L);
  }
  *out = ctx->extra_lanes[k - 1];
  return RSM_OK;
}

void destroy_lane(Lane& L) {
  for (auto& g : L.graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  L.graphs.clear();
  if (L.d_work.p) cudaFree(L.d_work.p);
  if (L.d_aux.p) cudaFree(L.d_aux.p);
  if (L.h_up.p) cudaFreeHost(L.h_up.p);
  if (L.h_down.p) cudaFreeHost(L.h_down.p);
  if (L.h_aux.p) cudaFreeHost(L.h_aux.p);
  for (cudaEvent_t e : L.ev_pool) cudaEventDestroy(e);
  if (L.ev_fork) cudaEventDestroy(L.ev_fork);
  if (L.ev_join) cudaEventDestroy(L.ev_join);
  if (L.done) cudaEventDestroy(L.done);
  if (L.stream2) cudaStreamDestroy(L.stream2);
  if (L.stream) cudaStreamDestroy(L.stream);
}

// CUDA-event bracket around a kernel class while profiling is on
struct Prof {
  rsm_ctx* ctx; Lane* L; int kc; cudaEvent_t a = nullptr, b = nullptr;
  Prof(rsm_ctx* c, int k, Lane* lane = nullptr) : ctx(c), L(lane ? lane : &c->L0), kc(k) {
    if (!ctx->profiling) return;
    while (L->ev_pool.size() < L->ev_used + 2) {
      cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; L->ev_pool.push_back(e);
    }
    a = L->ev_pool[L->ev_used++]; b = L->ev_pool[L->ev_used++];
    cudaEventRecord(a, L->stream);
  }
  ~Prof() {
    if (!a) return;
    cudaEventRecord(b, L->stream);
    L->spans.push_back({a, b, kc});
  }
};

// call after the lane's stream has been synchronised
void harvest_profile(rsm_ctx* ctx, Lane* L = nullptr) {
  if (!L) L = &ctx->L0;
  for (auto& s : L->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      if (s.kc == KC_SCORE) L->stats.score_kernel_ms += ms;
      else if (s.kc == KC_SELECT) L->stats.select_kernel_ms += ms;
      else if (s.kc == KC_RASTER) L->stats.raster_kernel_ms += ms;
    }
  }
  L->spans.clear();
  L->ev_used = 0;
}

int sync_stream(rsm_ctx* ctx, Lane* L = nullptr) {
  if (!L) L = &ctx->L0;
  CU(cudaStreamSynchronize(L->stream));
  harvest_profile(ctx, L);
  return RSM_OK;
}

// Wait for everything enqueued on the lane so far.  Batches sleep on a blocking-sync event (a rank of a
// multi-GPU job has few host cores, and a spinning wait takes one from the worker pool); single matches
// spin in cudaStreamSynchronize, whose wake-up is some 20 us faster.
int wait_lane(rsm_ctx* ctx, Lane* L, bool blocking) {
  if (blocking && L->done) {
    CU(cudaEventRecord(L->done, L->stream));
    // Poll and yield first: a blocking wait wakes ~50 us after the event (three times per chain and lane), a polling
    // thread that yields between queries costs the other lanes and the worker pool next to nothing.  After 2 ms sleep.
    static const bool yield_wait = [] { const char* e = std::getenv("RSM_WAIT"); return !(e && std::strcmp(e, "block") == 0); }();
    if (yield_wait) {
      const auto t0 = std::chrono::steady_clock::now();
      for (;;) {
        const cudaError_t q = cudaEventQuery(L->done);
        if (q == cudaSuccess) { harvest_profile(ctx, L); return RSM_OK; }
        if (q != cudaErrorNotReady) return fail(ctx, RSM_ERR_CUDA, "cudaEventQuery failed: %s", cudaGetErrorString(q));
        if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
        std::this_thread::yield();
      }
    }
    CU(cudaEventSynchronize(L->done));
    harvest_profile(ctx, L);
    return RSM_OK;
  }
  return sync_stream(ctx, L);
}

inline double key_to_score(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d;
  std::memcpy(&d, &b, 8);
  return d;
}

// value representable as an int32 multiple of 2^-25 in [0, 1] ?
inline bool fix_ok(float v, int* out) {
  if (!(v >= 0.0f && v <= 1.0f)) return false;
  const double s = std::ldexp(static_cast<double>(v), kFixShift);
  const double r = std::floor(s);
  if (r != s) return false;
  *out = static_cast<int>(r);
  return true;
}

struct TileCfg { int lx, ry, rows; bool affine; };

// Tile shape of the scoring kernel for a window of n_xy x n_xy translations: lanes along x (lx),
// ry consecutive y per thread.  Minimises padded work, per-beam overhead and table-build overhead.
TileCfg pick_tile(int n_xy, bool affine_ok) {
  int lx = 4;
  while (lx < n_xy && lx < 32) lx <<= 1;
  const bool affine = affine_ok;
  TileCfg best{lx, 1, score_rows(lx, 1), affine};
  double best_cost = 1e300;
  const double alpha = affine ? 1.0 : 3.0;   // per-beam cost relative to one row's gather
  for (int ry = 1; ry <= 8; ++ry) {
    const int rows = score_rows(lx, ry);
    const int tx = (n_xy + lx - 1) / lx, ty = (n_xy + rows - 1) / rows;
    const double padded = double(tx) * lx * double(ty) * rows;
    const double cost = padded * (1.0 + alpha / ry) * (1.0 + 4.0 * double(lx + rows) / (double(lx) * rows));
    if (cost < best_cost) { best_cost = cost; best = TileCfg{lx, ry, rows, affine}; }
  }
  if (const char* e = std::getenv("RSM_RY")) {   // tuning aid: force the rows-per-thread choice
    const int ry = std::atoi(e);
    if (ry >= 1 && ry <= 8) best = TileCfg{lx, ry, score_rows(lx, ry), affine};
  }
  return best;
}

// ---------------------------------------------------------------------------------------------
// one pass over a batch of items
// ---------------------------------------------------------------------------------------------
struct PassItem {
  const rsm_grid* grid = nullptr;
  const double* d_pts = nullptr;   // device points of this scan
  const double* h_pts = nullptr;   // or host points: they travel in the pass's own upload block (no copy of their own)
  int P = 0;
  rsm_pass_param param;
  double* pose_world = nullptr;    // in/out
  double* cov = nullptr;           // in/out
  double response = 0.0;
  rsm_pass_detail detail;
  int ang_begin = 0, ang_end = -1; // angle slice (-1: all)
  const double* center_map = nullptr;  // non-null: seed given in map coordinates, pose_world untouched
  // internal
  bool active = false;
  PassGeo geo;
  int a0 = 0, a1 = 0;              // resolved slice
  int64_t n_local = 0;             // candidates scored here
  size_t score_off = 0, trig_off = 0, spec_off = 0, acc_off = 0, ticket_off = 0;
  int sel_cta0 = 0, sel_ncta = 0;
  bool exact = false;
  BestPose best;
  std::vector<Cand> a_list, top, xy;
  int n_cols = 0;
  int cols[kMaxCols];
  size_t gather_off = 0;
};

enum PassMode { MODE_MATCH = 0, MODE_SCORES = 1, MODE_PARTIAL = 2 };


// exact path: reproduce the reference's sort on a full score array (candidate k of the array = global index base + k)
void finish_exact_core(const PassGeo& g, const rsm_pass_param& param, std::vector<Cand>& c, BestPose& best, double* cov) {
  std::sort(c.begin(), c.end(), by_score_desc);   // correlate_scan_matcher.h:607-608
  size_t na = 0;
  while (na < c.size() && DoubleEqual(c[na].score, c[0].score, kResponseFilterTolerance)) ++na;
  best = find_best(g, c.data(), na);
  const int type = param.type;
  if (type == RSM_COARSE || type == RSM_FINE || type == RSM_FAST)
    positional_cov(g, param, best, c.data(), std::min<size_t>(c.size(), kTopK), cov);
  if (type == RSM_COARSE || type == RSM_SUPER || type == RSM_FAST) {
    std::vector<Cand> xy;
    const double bound = cov_score_bound(best);
    for (const Cand& e : c) {
      if (!(e.score >= bound)) break;   // sorted: nothing further can qualify
      int ia, ix, iy;
      g.decode(e.index, &ia, &ix, &iy);
      if (same_xy(g, best, ix, iy)) { xy.push_back(e); if (xy.size() >= size_t(kMaxVarianceUsePointSize)) break; }
    }
    angular_cov(g, param, best, xy.data(), xy.size(), cov);
  }
}

void finish_exact(PassItem& it, const double* score) {
  const PassGeo& g = it.geo;
  const int64_t n = it.n_local;
  const int64_t base = int64_t(it.a0) * g.n_xy * g.n_xy;
  std::vector<Cand> c(n);
  for (int64_t k = 0; k < n; ++k) { c[k].score = score[k]; c[k].index = base + k; }
  finish_exact_core(g, it.param, c, it.best, it.cov);
}

bool adjacent_tie(const std::vector<Cand>& v, size_t n) {
  for (size_t i = 1; i < n && i < v.size(); ++i) if (v[i].score == v[i - 1].score) return true;
  return false;
}

// entries of a sorted list that ComputePositionalCovariance consumes (:904-921)
size_t positional_prefix(const std::vector<Cand>& top, double bound) {
  size_t used = 0;
  while (used < top.size() && used < size_t(kMaxVarianceUsePointSize) && top[used].score > bound) ++used;
  return used;
}

// Everything one pass keeps between its launch (pass_begin) and its host finalisation (pass_end).
struct PassRun {
  Lane* lane = nullptr;
  std::vector<PassItem>* items = nullptr;
  std::vector<int> act;
  PassMode mode = MODE_MATCH;
  bool pending = false, blocking = false;
  double* scores_out = nullptr;
  size_t o_best = 0, o_err = 0, o_poolcnt = 0, o_fcnt = 0, o_ftop = 0, o_speccols = 0, o_spec = 0, o_pool = 0,
         o_gjobs = 0, o_gout = 0, o_score = 0, o_trig = 0;
  size_t head_bytes = 0, gather_doubles = 0;
  int pool_first = 0, pool_cap = 0;
};

// ---- geometry of every item; returns the active ones ------------------------------------------------
int pass_geometry(rsm_ctx* ctx, std::vector<PassItem>& items, std::vector<int>& act) {
  const int n_items = int(items.size());
  act.clear();
  for (int i = 0; i < n_items; ++i) {
    PassItem& it = items[i];
    it.active = false; it.exact = false; it.response = 0.0;
    std::memset(&it.detail, 0, sizeof it.detail);
    if (!it.grid || !it.grid->init || it.P <= 0) continue;   // :792-795 -> response 0, outputs untouched
    if (it.param.type == RSM_FAST) return fail(ctx, RSM_ERR_UNSUPPORTED, "FAST (branch-and-bound) pass type is not provided");
    if (!(it.param.search_space_resolution > 0) || !(it.param.search_angle_resolution > 0) ||
        !(it.param.search_space_size >= 0) || !(it.param.search_angle_offset >= 0) || it.param.use_point_size < 2)
      return fail(ctx, RSM_ERR_INVALID, "bad pass parameters");
    double center[3];
    if (it.center_map) { center[0] = it.center_map[0]; center[1] = it.center_map[1]; center[2] = it.center_map[2]; }
    else it.grid->tf.world_to_map(it.pose_world, center);
    it.geo = make_geo(it.param, it.P, it.grid->cell_len(), center);
    if (it.geo.n_ang < 1 || it.geo.n_xy < 1 || it.geo.n_cand() > (int64_t(1) << 31) - 1)
      return fail(ctx, RSM_ERR_INVALID, "search window has %lld candidates", (long long)it.geo.n_cand());
    it.a0 = std::max(0, it.ang_begin);
    it.a1 = it.ang_end < 0 ? it.geo.n_ang : std::min(it.ang_end, it.geo.n_ang);
    if (it.a1 <= it.a0) continue;
    it.n_local = int64_t(it.a1 - it.a0) * it.geo.n_xy * it.geo.n_xy;
    it.active = true;
    act.push_back(i);
  }
  return RSM_OK;
}

// Items of one launch share the kernel plan (tile shape, variant), which is derived from the window width and
// the search step: a batch that mixes windows is cut into groups of equal (n_xy, step, cell type).
bool same_plan(const PassItem& a, const PassItem& b) {
  return a.geo.n_xy == b.geo.n_xy && a.geo.factor == b.geo.factor && a.grid->fixed == b.grid->fixed;
}

// The RSM_* environment switches as they are right now, from ONE pass over environ: a pass consults a dozen of them, and a
// getenv each is a few microseconds of a 14 us launch path.  (Read per pass, not cached: tests flip them at run time.)
struct EnvRsm {
  const char* entry[16];
  int n = 0;
  EnvRsm() {
    for (char** e = ::environ; e && *e; ++e)
      if ((*e)[0] == 'R' && (*e)[1] == 'S' && (*e)[2] == 'M' && (*e)[3] == '_' && n < 16) entry[n++] = *e + 4;
  }
  // value of RSM_<name> or nullptr
  const char* get(const char* name) const {
    const size_t len = std::strlen(name);
    for (int i = 0; i < n; ++i)
      if (std::strncmp(entry[i], name, len) == 0 && entry[i][len] == '=') return entry[i] + len + 1;
    return nullptr;
  }
};

// ---- stream plan of the staged scoring kernel (rsm_score.cu: score_stream_kernel) ----------------------
// The (job, angle, tile) items of a launch laid end to end; an item costs, in beams of a full tile: its beams x the
// tile's share of a full tile's shared-memory loads + 19 fixed (job fetch, beam table, pipeline fill) + 55 x that share
// for the epilogue.  n_cta equal shares; a cut closer than 24 beams to an item's end moves there.

void plan_stream(const std::vector<StreamRun>& runs, int variant, int max_ctas, StreamPlan& out) {
  constexpr double kFixed = 19.0, kEpilogue = 55.0;
  int tx, ty;
  score_stream_tile(variant, &tx, &ty);
  const double w_full = double(score_stream_weight(variant, tx, ty));
  // per run: the tiles of one angle (the pattern repeats for every angle)
  struct Info { double start, per_angle; long long item0; int tiles; size_t t0; };
  std::vector<Info> info(runs.size());
  std::vector<double> len, rho;
  double total = 0.0;
  long long n_items = 0;
  out.item_begin.assign(runs.size() + 1, 0);
  for (size_t r = 0; r < runs.size(); ++r) {
    const StreamRun& R = runs[r];
    Info& I = info[r];
    I.start = total; I.item0 = n_items; I.tiles = R.tiles_x * R.tiles_y; I.t0 = len.size(); I.per_angle = 0.0;
    for (int t = 0; t < I.tiles; ++t) {
      const int ex = std::min(tx, R.n_xy - (t % R.tiles_x) * tx), ey = std::min(ty, R.n_xy - (t / R.tiles_x) * ty);
      const double q = double(score_stream_weight(variant, ex, ey)) / w_full;
      rho.push_back(q);
      len.push_back(R.V * q + kFixed + kEpilogue * q);
      I.per_angle += len.back();
    }
    total += I.per_angle * R.ang_count;
    out.item_begin[r] = int(n_items);
    n_items += (long long)R.ang_count * I.tiles;
  }
  out.item_begin[runs.size()] = int(n_items);
  out.n_items = n_items;
  const int G = int(std::max<long long>(1, std::min<long long>(max_ctas, (long long)(total / 96.0))));
  // cuts: (item, beam, run of the item), strictly increasing; the k-th lies at k / G of the total length
  struct Cut { long long item; int beam; int run; };
  std::vector<Cut> cuts;
  cuts.reserve(size_t(G) + 1);
  cuts.push_back({0, 0, 0});
  size_t r = 0;
  for (int k = 1; k < G; ++k) {
    const double target = total * k / G;
    while (r + 1 < runs.size() && target >= info[r + 1].start) ++r;
    const StreamRun& R = runs[r];
    const Info& I = info[r];
    double off = target - I.start;
    const int ia = int(std::min<double>(R.ang_count - 1, std::floor(off / I.per_angle)));
    off -= ia * I.per_angle;
    int t = 0;
    while (t + 1 < I.tiles && off >= len[I.t0 + t]) { off -= len[I.t0 + t]; ++t; }
    const long long item = I.item0 + (long long)ia * I.tiles + t;
    const int guard = std::min(24, R.V / 2);
    const int beam = int(std::lround((off - kFixed) / rho[I.t0 + t]));
    Cut c;
    if (beam < guard || beam <= 0) c = {item, 0, int(r)};
    else if (beam > R.V - guard || beam >= R.V) {
      c = {item + 1, 0, int(r)};
      if (c.item >= I.item0 + (long long)R.ang_count * I.tiles) c.run = int(r) + 1;
    } else c = {item, beam, int(r)};
    const Cut& l = cuts.back();
    if (c.item < n_items && (c.item > l.item || (c.item == l.item && c.beam > l.beam))) cuts.push_back(c);
  }
  cuts.push_back({n_items, 0, int(runs.size())});
  // shared items: those with a cut inside; parts = cuts inside + 1, in beam order
  const size_t nc = cuts.size();
  std::vector<int> cut_ticket(nc, -1), cut_ord(nc, 0), t_slot0, t_parts;
  for (size_t c = 1; c + 1 < nc; ++c) {
    if (cuts[c].beam == 0) continue;
    if (cut_ticket[c - 1] >= 0 && cuts[c - 1].item == cuts[c].item && cuts[c - 1].beam > 0) {
      cut_ticket[c] = cut_ticket[c - 1]; cut_ord[c] = cut_ord[c - 1] + 1;
    } else {
      cut_ticket[c] = int(t_parts.size()); cut_ord[c] = 0;
      t_parts.push_back(1);
    }
    t_parts[cut_ticket[c]]++;
  }
  t_slot0.resize(t_parts.size());
  int slots = 0;
  for (size_t t = 0; t < t_parts.size(); ++t) { t_slot0[t] = slots; slots += t_parts[t]; }
  out.n_tickets = int(t_parts.size()); out.n_slots = slots;
  out.ctas.resize(nc - 1);
  for (size_t c = 0; c + 1 < nc; ++c) {
    StreamCta& A = out.ctas[c];
    std::memset(&A, 0, sizeof A);
    A.ticket0 = A.ticket1 = -1;
    const Cut& b = cuts[c];
    const Cut& e = cuts[c + 1];
    A.item0 = int(b.item); A.beam0 = b.beam;
    if (e.beam == 0) {
      A.item1 = int(e.item - 1);
      const int run = (e.run < int(runs.size()) && e.item > info[e.run].item0) ? e.run : e.run - 1;
      A.beam1 = runs[run].V;
    } else { A.item1 = int(e.item); A.beam1 = e.beam; }
    if (b.beam > 0) { A.ticket0 = cut_ticket[c]; A.part0 = cut_ord[c] + 1; }
    else if (e.beam > 0 && e.item == b.item) { A.ticket0 = cut_ticket[c + 1]; A.part0 = 0; }
    if (A.ticket0 >= 0) { A.slot0 = t_slot0[A.ticket0]; A.parts0 = t_parts[A.ticket0]; }
    if (e.beam > 0 && A.item1 != A.item0) {
      A.ticket1 = cut_ticket[c + 1]; A.part1 = 0; A.slot1 = t_slot0[A.ticket1]; A.parts1 = t_parts[A.ticket1];
    }
  }
}

// ---- launch: host preparation + everything enqueued on the lane's stream ------------------------------
int pass_begin(rsm_ctx* ctx, Lane* lane, std::vector<PassItem>& items, const std::vector<int>& act_in, PassMode mode,
               double* scores_out, int64_t scores_cap, int64_t* scores_written, PassRun& R) {
  rsm_stats& ST = lane->stats;
  PhaseTimer pt(&ST);
  const EnvRsm env;      // the debugging switches, read once per pass
  R = PassRun();
  R.lane = lane; R.items = &items; R.act = act_in; R.mode = mode; R.scores_out = scores_out;
  const std::vector<int>& act = R.act;
  if (scores_written) *scores_written = 0;
  if (act.empty()) return RSM_OK;
  const int na = int(act.size());
  cudaStream_t st = lane->stream;

  // ---- layouts -------------------------------------------------------------------------------
  // the affine variant needs the search step to be the same exact integer number of cells for every job
  bool affine_ok = true;
  {
    const double f0 = items[act[0]].geo.factor;
    if (!(f0 >= 1.0 && f0 <= 64.0 && f0 == std::floor(f0))) affine_ok = false;
    for (int a = 0; a < na && affine_ok; ++a) {
      // same step for every job, and window coordinates small enough for the rounding bound of
      // the affine index test (rsm_score.cu) to hold
      const PassGeo& g = items[act[a]].geo;
      const double lim = 1048576.0;
      if (g.factor != f0 || !(std::fabs(g.start_x) < lim && std::fabs(g.start_y) < lim &&
                              std::fabs(g.x_of(g.n_xy + 64)) < lim && std::fabs(g.y_of(g.n_xy + 64)) < lim))
        affine_ok = false;
    }
  }
  TileCfg cfg = pick_tile(items[act[0]].geo.n_xy, affine_ok);
  // staged variant (TMA boxes of the grid in shared memory, rsm_score.cu): windows of 48 and more
  // translations per axis at unit step on fixed-point grids.  Measured on B200 against the L1 path:
  // 0.224 vs 0.265 ms on config 2, 11.3 vs 13.6 ms on the wide window.  RSM_NO_STAGED=1 turns it off.
  bool use_staged = affine_ok && items[act[0]].geo.factor == 1.0 && items[act[0]].geo.n_xy >= 48 &&
                    env.get("NO_STAGED") == nullptr;
  int max_V = 0;
  for (int a = 0; a < na && use_staged; ++a) {
    const PassItem& it = items[act[a]];
    if (!it.grid->fixed || it.geo.visited > 2048 || (it.grid->pitch & 3)) use_staged = false;
    max_V = std::max(max_V, it.geo.visited);
  }
  int n_split = 1, staged_variant = 0, split_b = 0;   // split_b: split of the second launch, 0 = one launch
  bool use_stream = false;
  int stream_variant = 0;
  static const StreamPlan kNoPlan;
  std::shared_ptr<const StreamPlan> splan_hold;      // keeps a cached plan alive while this pass uses it
  const StreamPlan* splan = &kNoPlan;
  long long items_a = 0;                             // (angle, tile) items of the first launch
  if (use_staged) {
    int tx, ty;
    int max_nxy = 0;
    for (int a = 0; a < na; ++a) max_nxy = std::max(max_nxy, items[act[a]].geo.n_xy);
    staged_variant = score_staged_variant(max_nxy);
    score_staged_tile(staged_variant, &tx, &ty);
    cfg.lx = tx; cfg.rows = ty; cfg.ry = 0; cfg.affine = true;
    // Beams are split over the CTAs of a cluster (n_split) so that waves of CTAs are full.  A CTA
    // costs, in beams: its share of the beams, a fixed part (job fetch, beam table, pipeline fill)
    // worth about 19, its share of the epilogue (55 / split) and the cluster exchange (10).  A wave
    // is the CTAs resident at once for that cluster size.  With few waves one split size leaves
    // the last wave mostly empty, so the angles may be cut in two launches: whole waves at split
    // A, the remaining angles at a larger split B that fills one more, shorter wave.
    long long work_items = 0;
    int min_V = 1 << 30;
    for (int a = 0; a < na; ++a) {
      const PassItem& it = items[act[a]];
      const int t1 = (it.geo.n_xy + tx - 1) / tx, t2 = (it.geo.n_xy + ty - 1) / ty;
      work_items += (long long)(it.a1 - it.a0) * t1 * t2;
      min_V = std::min(min_V, it.geo.visited);
    }
    const int s_max = std::max(1, std::min(8, min_V / 64));
    int resident[9] = {0};
    double unit[9] = {0};
    for (int sp = 1; sp <= s_max; ++sp) {
      resident[sp] = score_staged_resident_ctas(staged_variant, sp, max_V);
      unit[sp] = double(max_V) / sp + 19.0 + 55.0 / sp + (sp > 1 ? 10.0 : 0.0);
    }
    const bool two_phase_ok = env.get("ONE_PHASE") == nullptr;
    double best_cost = 0.0;
    for (int sa = 1; sa <= s_max; ++sa) {
      if (resident[sa] <= 0) continue;
      const long long per_wave = resident[sa] / sa;                       // work items per full wave at split sa
      const long long full = (work_items + per_wave - 1) / per_wave;      // waves if everything runs at split sa
      const double one = double(full) * unit[sa];
      if (best_cost == 0.0 || one < best_cost) { best_cost = one; n_split = sa; split_b = 0; items_a = work_items; }
      if (!two_phase_ok || full < 2 || full > 16) continue;
      const long long in_a = (full - 1) * per_wave;                        // whole waves at split sa
      for (int sb = sa + 1; sb <= s_max; ++sb) {
        if (resident[sb] <= 0) continue;
        const long long rest = work_items - in_a;
        const long long waves_b = (rest * sb + resident[sb] - 1) / resident[sb];
        const double two = double(full - 1) * unit[sa] + double(waves_b) * unit[sb] + 12.0;   // + a launch boundary
        if (two < best_cost) { best_cost = two; n_split = sa; split_b = sb; items_a = in_a; }
      }
    }
    if (env.get("DEBUG_SPLIT"))
      std::fprintf(stderr, "[rsm] staged plan: %lld items, split %d for the first %lld, split %d for the rest, cost %.0f\n",
                   work_items, n_split, items_a, split_b, best_cost);
    // Stream plan (persistent CTAs over one sequence of beams) from half a wave of items on: no wave quantisation
    // and the paired-row mapping for 65+ columns.  Below that the cluster plan splits the epilogue too.
    // RSM_STAGED_PLAN=cluster|stream forces one; RSM_STREAM_CTAS / RSM_STREAM_VARIANT override the shape (tests).
    if (ctx->sm_count <= 0) cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
    const char* plan_env = env.get("STAGED_PLAN");
    use_stream = plan_env ? std::strcmp(plan_env, "stream") == 0 : 2 * work_items >= ctx->sm_count;
    if (use_stream) {
      stream_variant = score_stream_variant(max_nxy);
      if (const char* e = env.get("STREAM_VARIANT")) stream_variant = std::max(0, std::min(2, std::atoi(e)));
      score_stream_tile(stream_variant, &tx, &ty);
      cfg.lx = tx; cfg.rows = ty;
      std::vector<StreamRun> runs(na);
      for (int a = 0; a < na; ++a) {
        const PassItem& it = items[act[a]];
        runs[a] = {it.geo.visited, it.geo.n_xy, it.a1 - it.a0, (it.geo.n_xy + tx - 1) / tx, (it.geo.n_xy + ty - 1) / ty};
      }
      int max_ctas = std::max(1, ctx->sm_count);
      if (const char* e = env.get("STREAM_CTAS")) max_ctas = std::max(1, std::atoi(e));
      {
        std::vector<int> key = {stream_variant, max_ctas};
        for (const StreamRun& r : runs) key.insert(key.end(), {r.V, r.n_xy, r.ang_count});
        std::shared_ptr<const StreamPlan> hit;
        {
          std::lock_guard<std::mutex> lk(ctx->mu);
          auto it = ctx->stream_plans.find(key);
          if (it != ctx->stream_plans.end()) hit = it->second;
        }
        if (!hit) {
          auto made = std::make_shared<StreamPlan>();
          plan_stream(runs, stream_variant, max_ctas, *made);
          hit = made;
          std::lock_guard<std::mutex> lk(ctx->mu);
          if (ctx->stream_plans.size() >= 32) ctx->stream_plans.clear();
          ctx->stream_plans[key] = hit;
        }
        splan_hold = hit;
        splan = hit.get();
      }
      if (splan->n_items > INT_MAX / 2) use_stream = false;
      if (env.get("DEBUG_SPLIT"))
        std::fprintf(stderr, "[rsm] stream plan: variant %d, %lld items over %zu CTAs, %d shared (%d partial slots)\n", stream_variant,
                     splan->n_items, splan->ctas.size(), splan->n_tickets, splan->n_slots);
    }
  }
  // flat variant: small windows whose step is not an integer number of cells (fine / super-fine passes)
  bool use_flat = !use_staged && !cfg.affine && env.get("NO_FLAT") == nullptr;
  int flat_k = 1;                                  // candidates per thread of the flat kernel
  for (int a = 0; a < na && use_flat; ++a) {
    const PassGeo& g = items[act[a]].geo;
    // 16.16 fixed-point coordinates: window coordinates must stay well inside +-2^14 cells
    if (g.n_xy < 3 || g.n_xy > score_flat_max_nxy() || !(std::fabs(g.start_x) < 8192.0 && std::fabs(g.start_y) < 8192.0 &&
                                       std::fabs(g.x_of(g.n_xy)) < 8192.0 && std::fabs(g.y_of(g.n_xy)) < 8192.0))
      use_flat = false;
  }
  if (use_flat) flat_k = score_flat_k(items[act[0]].geo.n_xy);
  // patch variant: batches of small unit-step windows on fixed-point grids (the back-end chain's coarse pass).
  // Its CTAs are 8 angles x the whole window x every beam, four resident per SM: below one full wave of them
  // (592) the tiled kernel's many small CTAs balance better (measured: 64-pair sub-batches lose 4 %, 256-pair
  // batches gain 18 % of the coarse pass).  RSM_NO_PATCH=1 / RSM_FORCE_PATCH=1 override.
  bool use_patch = !use_staged && !use_flat && cfg.affine && items[act[0]].geo.factor == 1.0 && env.get("NO_PATCH") == nullptr;
  int patch_nxy = 0, patch_ctas = 0;
  for (int a = 0; a < na && use_patch; ++a) {
    const PassItem& it = items[act[a]];
    if (!it.grid->fixed || (it.grid->pitch & 3) || it.geo.n_xy < 3 || it.geo.n_xy > 16) use_patch = false;
    patch_nxy = std::max(patch_nxy, it.geo.n_xy);
    patch_ctas += (it.a1 - it.a0 + score_patch_angles() - 1) / score_patch_angles();
  }
  if (use_patch && patch_ctas < 256 && env.get("FORCE_PATCH") == nullptr) use_patch = false;
  // immediate-offset variant: unit search step and the same padded pitch for every job
  int const_pitch = 0;
  if (!use_staged && !use_patch && cfg.affine && items[act[0]].geo.factor == 1.0 && cfg.lx >= 16) {
    const_pitch = items[act[0]].grid->pitch;
    if (const_pitch != kPitchSmall && const_pitch != kPitchLarge) const_pitch = 0;
    for (int a = 1; a < na && const_pitch; ++a) if (items[act[a]].grid->pitch != const_pitch) const_pitch = 0;
  }
  const int rows = cfg.rows;
  // launches of the tiled / flat kernels on fixed-point grids that do not fill the machine: split the beams of every
  // (angle, tile) over `beam_split` CTAs that add integer partial sums into per-candidate accumulators; the last one
  // to arrive finishes.  A single front-end match is 1 .. 81 CTAs that would each walk ~1000 beams alone; the
  // super-fine pass of a 512-pair batch is 512 CTAs, less than one wave.
  int beam_split = 1;
  if (!use_staged && !use_patch && env.get("NO_BEAM_SPLIT") == nullptr) {
    bool all_fixed = true;
    long long base_ctas = 0, cands = 0;
    int min_chunks = 1 << 30;
    for (int a = 0; a < na; ++a) {
      const PassItem& it = items[act[a]];
      const PassGeo& g = it.geo;
      if (!it.grid->fixed) all_fixed = false;
      base_ctas += use_flat ? score_flat_ctas(int(it.n_local), flat_k)
                            : (long long)(it.a1 - it.a0) * ((g.n_xy + cfg.lx - 1) / cfg.lx) * ((g.n_xy + rows - 1) / rows);
      cands += it.n_local;
      min_chunks = std::min(min_chunks, (g.visited + 31) / 32);
    }
    long long target = 6 * 148;     // CTAs wanted: these are 128 .. 256-thread CTAs, several resident per SM
    if (const char* e = env.get("SPLIT_TARGET")) target = std::max(1, std::atoi(e));
    if (all_fixed && base_ctas > 0 && cands <= (1 << 20)) {
      if (2 * base_ctas <= target)
        beam_split = int(std::max<long long>(1, std::min<long long>({(long long)min_chunks, 32, (target + base_ctas - 1) / base_ctas})));
      else if (use_flat && base_ctas < 4 * target)   // a batch's super-fine pass: 3-4 waves of shorter CTAs balance better
        beam_split = int(std::max<long long>(1, std::min<long long>({(long long)min_chunks / 4, 8, (4 * target + base_ctas - 1) / base_ctas})));
    }
  }
  std::vector<ScoreJob> sjobs(na);
  std::vector<int> s_cta(na + 1, 0);
  std::vector<SelectJob> ljobs(na);
  std::vector<int> l_cta(na + 1, 0);
  Layout dl;   // device work arena
  const size_t o_sjobs = dl.take(sizeof(ScoreJob) * (use_staged ? 2 * size_t(na) : size_t(na)));   // staged: up to two launches
  const size_t o_scta = dl.take(sizeof(int) * (use_staged ? 2 * (size_t(na) + 1) : size_t(na) + 1));
  const size_t o_ljobs = dl.take(sizeof(SelectJob) * na);
  const size_t o_lcta = dl.take(sizeof(int) * (na + 1));
  size_t trig_doubles = 0;
  for (int a = 0; a < na; ++a) { items[act[a]].trig_off = trig_doubles; trig_doubles += size_t(items[act[a]].geo.n_ang) * 3; }
  const size_t o_trig = dl.take(trig_doubles * 8);
  size_t hpts_doubles = 0;
  for (int a = 0; a < na; ++a) if (items[act[a]].h_pts) hpts_doubles += size_t(items[act[a]].P) * 2;
  // (rounded up to 2 KB: the size of the upload is part of a pass graph's key, and scans of one sensor differ by a few points)
  const size_t o_hpts = dl.take((hpts_doubles * 8 + 2047) / 2048 * 2048, 16);
  const size_t o_tmaps = dl.take(use_staged ? size_t(na) * 256 : 0, 128);
  const size_t o_splan = dl.take(use_stream ? splan->ctas.size() * sizeof(StreamCta) : 0, 16);
  const size_t up_bytes = dl.off;          // everything above is uploaded in one copy
  // zero-initialised block: [beam-split accumulators and tickets (not read back),] best keys, err flags, pool counter
  size_t acc_total = 0, ticket_total = 0;
  if (beam_split > 1)
    for (int a = 0; a < na; ++a) {
      PassItem& it = items[act[a]];
      const PassGeo& g = it.geo;
      it.acc_off = acc_total; it.ticket_off = ticket_total;
      acc_total += size_t(it.n_local);
      ticket_total += use_flat ? size_t(score_flat_ctas(int(it.n_local), flat_k))
                               : size_t(it.a1 - it.a0) * ((g.n_xy + cfg.lx - 1) / cfg.lx) * ((g.n_xy + rows - 1) / rows);
    }
  const size_t o_acc = dl.take(acc_total * 8, 256);
  const size_t o_tickets = dl.take(ticket_total * 4, 4);
  const size_t o_stickets = dl.take(use_stream ? size_t(splan->n_tickets) * 4 : 0, 4);   // stream plan: tickets of the shared items
  const size_t zero_begin = beam_split > 1 ? o_acc : use_stream ? o_stickets : dl.off;
  const bool zero_wide = beam_split > 1 || use_stream;      // the zeroed block starts before the best keys
  const size_t o_best = dl.take(size_t(na) * 8, beam_split > 1 ? 8 : 256);
  const size_t o_err = dl.take(size_t(na) * 4, 4);
  const size_t o_poolcnt = dl.take(4, 4);
  const size_t o_done = dl.take(size_t(na) * 4, 4);
  const size_t zero_end = dl.off;
  int total_sel_cta = 0;
  int64_t total_cand = 0;
  for (int a = 0; a < na; ++a) {
    PassItem& it = items[act[a]];
    int ncta = int(std::min<int64_t>(592, (it.n_local + kSelectSlice - 1) / kSelectSlice));
    if (ncta < 1) ncta = 1;
    it.sel_cta0 = total_sel_cta; it.sel_ncta = ncta;
    total_sel_cta += ncta;
    total_cand += it.n_local;
  }
  const size_t o_fcnt = dl.take(size_t(na) * 4, 4);
  const size_t o_speccols = dl.take(size_t(na) * (kMaxCols + 1) * 4, 4);
  const size_t o_ftop = dl.take(size_t(na) * kTopK * sizeof(Entry), 16);
  size_t spec_doubles = 0;
  for (int a = 0; a < na; ++a) {
    PassItem& it = items[act[a]];
    it.spec_off = spec_doubles;
    if (it.param.type == RSM_COARSE || it.param.type == RSM_SUPER) spec_doubles += size_t(kMaxCols) * (it.a1 - it.a0);
  }
  const size_t o_spec = dl.take(spec_doubles * 8, 16);
  const int pool_cap = int(std::min<int64_t>(std::max<int64_t>(int64_t(na) * 64, 65536), total_cand));
  const size_t o_pool = dl.take(size_t(pool_cap) * sizeof(PoolEntry), 16);
  const size_t down_end = dl.off;          // [o_best, down_end) is what the host reads back
  const size_t o_topcnt = dl.take(size_t(total_sel_cta) * 4, 4);   // per-CTA lists stay on the device
  const size_t o_top = dl.take(size_t(total_sel_cta) * kTopK * sizeof(Entry), 16);
  const size_t o_gjobs = dl.take(sizeof(GatherJob) * na);
  size_t gather_doubles = 0;
  for (int a = 0; a < na; ++a) { items[act[a]].gather_off = gather_doubles; gather_doubles += size_t(kMaxCols) * items[act[a]].geo.n_ang; }
  const size_t o_gout = dl.take(gather_doubles * 8);
  size_t score_doubles = 0;
  for (int a = 0; a < na; ++a) { items[act[a]].score_off = score_doubles; score_doubles += size_t(items[act[a]].n_local); }
  const size_t o_score = dl.take(score_doubles * 8);
  const size_t o_spart = dl.take(use_stream ? size_t(splan->n_slots) * score_stream_partial_words(stream_variant) * 8 : 0, 256);
  int rc = ensure_dev(ctx, lane->d_work, dl.off, st);
  if (rc) return rc;
  char* dw = lane->d_work.p;
  rc = ensure_pinned(ctx, lane->h_up, std::max(up_bytes, sizeof(GatherJob) * size_t(na)), st);
  if (rc) return rc;
  rc = ensure_pinned(ctx, lane->h_down, std::max(down_end - o_best, gather_doubles * 8), st);
  if (rc) return rc;

  // ---- fill jobs, angle tables ---------------------------------------------------------------
  char* up = lane->h_up.p;
  double* h_trig = reinterpret_cast<double*>(up + o_trig);
  int cta = 0;
  size_t hpts_used = 0;
  bool any_fixed = false, any_float = false;
  ctx->pool.run(na, 32, [&](int a0, int a1) {
    for (int a = a0; a < a1; ++a) {
      const PassItem& it = items[act[a]];
      const PassGeo& g = it.geo;
      double* t = h_trig + it.trig_off;
      for (int ia = 0; ia < g.n_ang; ++ia) {
        const double angle = g.angle_of(ia);
        angle_trig(angle, &t[3 * ia], &t[3 * ia + 1]);      // correlate_scan_matcher.h:171-172
        t[3 * ia + 2] = angle;
      }
    }
  });
  for (int a = 0; a < na; ++a) {
    PassItem& it = items[act[a]];
    const PassGeo& g = it.geo;
    ScoreJob& J = sjobs[a];
    std::memset(&J, 0, sizeof J);
    J.grid = it.grid->d_cells;
    J.pts = it.d_pts;
    if (it.h_pts) {       // host scan: behind the angle tables in the upload block
      std::memcpy(up + o_hpts + hpts_used * 8, it.h_pts, size_t(it.P) * 16);
      J.pts = reinterpret_cast<const double*>(dw + o_hpts) + hpts_used;
      hpts_used += size_t(it.P) * 2;
    }
    J.trig = reinterpret_cast<const double*>(dw + o_trig) + it.trig_off;
    J.score = reinterpret_cast<double*>(dw + o_score) + it.score_off;
    J.best_key = reinterpret_cast<unsigned long long*>(dw + o_best) + a;
    J.err = reinterpret_cast<int*>(dw + o_err) + a;
    J.pitch = it.grid->pitch; J.size_x = it.grid->size_x; J.size_y = it.grid->size_y;
    J.P = it.P; J.step = g.step; J.V = g.visited;
    J.n_xy = g.n_xy; J.ang_begin = it.a0; J.ang_count = it.a1 - it.a0;
    J.tiles_x = (g.n_xy + cfg.lx - 1) / cfg.lx;
    J.tiles_y = (g.n_xy + rows - 1) / rows;
    J.use_penalty = it.param.use_center_penalty ? 1 : 0;
    J.f_int = cfg.affine ? int(g.factor) : 0;
    J.stepoff = J.f_int * it.grid->pitch;
    J.n_split = use_stream ? 1 : use_staged ? n_split : beam_split;
    if (beam_split > 1) {
      J.acc = reinterpret_cast<unsigned long long*>(dw + o_acc) + it.acc_off;
      J.tickets = reinterpret_cast<int*>(dw + o_tickets) + it.ticket_off;
    }
    if (use_staged) {
      rsm_ctx::TmapPair tp;
      rc = grid_tmaps(ctx, it.grid, use_stream ? stream_variant : -1, &tp);
      if (rc) return rc;
      std::memcpy(up + o_tmaps + size_t(a) * 256, &tp, 256);
      J.tmap = dw + o_tmaps + size_t(a) * 256;
    }
    J.divisor = double(g.divisor);
    J.sx = g.start_x; J.sy = g.start_y; J.f = g.factor;
    J.cx = g.center[0]; J.cy = g.center[1]; J.ca = g.center[2];
    J.m2 = g.cell_len * g.cell_len;                                   // :733
    J.half_size = it.param.search_space_size / 2;                     // :734
    J.gain = (it.param.type == RSM_COARSE) ? 0.4 : 0.2;               // :588-602, :759-761
    s_cta[a] = cta;
    cta += use_flat ? score_flat_ctas(int(it.n_local), flat_k) * beam_split
                    : use_patch ? (J.ang_count + score_patch_angles() - 1) / score_patch_angles()
                                : J.ang_count * J.tiles_x * J.tiles_y * (use_staged ? n_split : beam_split);
    (it.grid->fixed ? any_fixed : any_float) = true;
    SelectJob& L = ljobs[a];
    std::memset(&L, 0, sizeof L);
    L.score = J.score; L.n = it.n_local; L.best_key = J.best_key;
    L.top_list = reinterpret_cast<Entry*>(dw + o_top) + size_t(it.sel_cta0) * kTopK;
    L.top_count = reinterpret_cast<int*>(dw + o_topcnt) + it.sel_cta0;
    L.final_top = reinterpret_cast<Entry*>(dw + o_ftop) + size_t(a) * kTopK;
    L.final_count = reinterpret_cast<int*>(dw + o_fcnt) + a;
    const bool wants_angular = it.param.type == RSM_COARSE || it.param.type == RSM_SUPER;
    L.spec_cols = reinterpret_cast<int*>(dw + o_speccols) + size_t(a) * (kMaxCols + 1);
    L.spec_out = wants_angular ? reinterpret_cast<double*>(dw + o_spec) + it.spec_off : nullptr;
    L.done = reinterpret_cast<int*>(dw + o_done) + a;
    L.n_xy = g.n_xy; L.n_ang = it.a1 - it.a0;
    L.err = J.err; L.job_id = a; L.n_cta = it.sel_ncta;
    L.slice = (it.n_local + it.sel_ncta - 1) / it.sel_ncta;
    l_cta[a] = it.sel_cta0;
    ST.evals += it.n_local * g.visited;
  }
  s_cta[na] = cta; l_cta[na] = total_sel_cta;
  if (any_fixed && any_float) return fail(ctx, RSM_ERR_UNSUPPORTED, "a batch must not mix fixed-point and float32 grids");
  // staged variant: cut the (job, angle) sequence into the launches of the plan
  struct StagedLaunch { int split = 1, n_jobs = 0, n_cta = 0, beams = 0; size_t jobs_off = 0, cta_off = 0; };
  StagedLaunch launches[2];
  int n_launches = 0;
  if (use_stream) {
    std::memcpy(up + o_sjobs, sjobs.data(), sizeof(ScoreJob) * na);
    std::memcpy(up + o_scta, splan->item_begin.data(), sizeof(int) * (na + 1));
    std::memcpy(up + o_splan, splan->ctas.data(), sizeof(StreamCta) * splan->ctas.size());
    n_launches = 1;
  } else if (use_staged) {
    std::vector<ScoreJob> pj[2];
    std::vector<int> pc[2];
    long long left_a = split_b ? items_a : (long long)1 << 60;
    for (int a = 0; a < na; ++a) {
      const ScoreJob& J = sjobs[a];
      const long long tiles = (long long)J.tiles_x * J.tiles_y;
      const int ang_a = int(std::min<long long>(J.ang_count, left_a / tiles));
      left_a -= (long long)ang_a * tiles;
      if (ang_a < J.ang_count) left_a = 0;             // the cut is one point of the sequence
      for (int ph = 0; ph < 2; ++ph) {
        const int first = ph == 0 ? 0 : ang_a, count = ph == 0 ? ang_a : J.ang_count - ang_a;
        if (count <= 0) continue;
        ScoreJob K = J;
        K.ang_begin = J.ang_begin + first; K.ang_count = count;
        K.score = J.score + (long long)first * J.n_xy * J.n_xy;
        K.n_split = ph == 0 ? n_split : split_b;
        pc[ph].push_back(launches[ph].n_cta);
        launches[ph].n_cta += int(count * tiles) * K.n_split;
        pj[ph].push_back(K);
      }
    }
    size_t jobs_off = o_sjobs, cta_off = o_scta;
    for (int ph = 0; ph < 2; ++ph) {
      if (pj[ph].empty()) continue;
      StagedLaunch& L = launches[n_launches++];
      L = launches[ph];
      L.split = ph == 0 ? n_split : split_b;
      L.n_jobs = int(pj[ph].size());
      L.beams = (max_V + L.split - 1) / L.split;
      pc[ph].push_back(L.n_cta);
      L.jobs_off = jobs_off; L.cta_off = cta_off;
      std::memcpy(up + jobs_off, pj[ph].data(), sizeof(ScoreJob) * pj[ph].size());
      std::memcpy(up + cta_off, pc[ph].data(), sizeof(int) * pc[ph].size());
      jobs_off += sizeof(ScoreJob) * pj[ph].size();
      cta_off += sizeof(int) * pc[ph].size();
    }
  } else {
    std::memcpy(up + o_sjobs, sjobs.data(), sizeof(ScoreJob) * na);
    std::memcpy(up + o_scta, s_cta.data(), sizeof(int) * (na + 1));
  }
  std::memcpy(up + o_ljobs, ljobs.data(), sizeof(SelectJob) * na);
  std::memcpy(up + o_lcta, l_cta.data(), sizeof(int) * (na + 1));

  // ---- launch --------------------------------------------------------------------------------
  char* dn = lane->h_down.p;
  const size_t head_bytes = o_pool - o_best;
  const int pool_first = std::min(pool_cap, std::max(4096, na * 8));
  const bool fork = use_staged && !use_stream && n_launches == 2 && env.get("NO_FORK") == nullptr;
  auto enqueue_score = [&]() -> int {
    CU(cudaMemcpyAsync(dw, up, up_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dw + (zero_wide ? zero_begin : o_best), 0, zero_end - (zero_wide ? zero_begin : o_best), st));
    {
      Prof p(ctx, KC_SCORE, lane);
      if (use_flat)
        CU(launch_score_flat(any_fixed, flat_k, cta, st, reinterpret_cast<const ScoreJob*>(dw + o_sjobs),
                               reinterpret_cast<const int*>(dw + o_scta), na));
      else if (use_patch)
        CU(launch_score_patch(patch_nxy, cta, st, reinterpret_cast<const ScoreJob*>(dw + o_sjobs),
                              reinterpret_cast<const int*>(dw + o_scta), na));
      else if (use_stream)
        CU(launch_score_stream(stream_variant, int(splan->ctas.size()), max_V, st, reinterpret_cast<const ScoreJob*>(dw + o_sjobs),
                               reinterpret_cast<const int*>(dw + o_scta), na, reinterpret_cast<const StreamCta*>(dw + o_splan),
                               reinterpret_cast<unsigned long long*>(dw + o_spart), reinterpret_cast<int*>(dw + o_stickets)));
      else if (use_staged) {
        // the launches cover disjoint angles: the second one goes to a side stream so that its
        // clusters take SMs as soon as CTAs of the first retire (no kernel-boundary drain between them)
        if (fork) {
          CU(cudaEventRecord(lane->ev_fork, st));
          CU(cudaStreamWaitEvent(lane->stream2, lane->ev_fork, 0));
        }
        for (int l = 0; l < n_launches; ++l) {
          const StagedLaunch& L = launches[l];
          CU(launch_score_staged(staged_variant, L.split, L.n_cta, L.beams, (fork && l == 1) ? lane->stream2 : st,
                                 reinterpret_cast<const ScoreJob*>(dw + L.jobs_off), reinterpret_cast<const int*>(dw + L.cta_off), L.n_jobs));
        }
        if (fork) {
          CU(cudaEventRecord(lane->ev_join, lane->stream2));
          CU(cudaStreamWaitEvent(st, lane->ev_join, 0));
        }
      } else
        CU(launch_score(any_fixed, cfg.affine, cfg.lx, cfg.ry, any_fixed ? const_pitch : 0, cta, st,
                        reinterpret_cast<const ScoreJob*>(dw + o_sjobs), reinterpret_cast<const int*>(dw + o_scta), na));
    }
    return RSM_OK;
  };
  auto enqueue_tail = [&]() -> int {
    {
      Prof p(ctx, KC_SELECT, lane);
      CU(launch_select(total_sel_cta, st, reinterpret_cast<const SelectJob*>(dw + o_ljobs),
                       reinterpret_cast<const int*>(dw + o_lcta), na, reinterpret_cast<PoolEntry*>(dw + o_pool),
                       pool_cap, reinterpret_cast<int*>(dw + o_poolcnt)));
    }
    // read back everything up to the pool, plus a first slice of the pool
    CU(cudaMemcpyAsync(dn, dw + o_best, head_bytes + size_t(pool_first) * sizeof(PoolEntry), cudaMemcpyDeviceToHost, st));
    return RSM_OK;
  };
  ST.h2d_bytes += up_bytes;
  ST.kernel_launches += use_staged ? n_launches : 1;
  ST.score_launches++;

  if (mode == MODE_SCORES) {
    // parity/debug: hand the raw score array of item 0 back
    rc = enqueue_score();
    if (rc) return rc;
    PassItem& it = items[act[0]];
    if (it.n_local > scores_cap) return fail(ctx, RSM_ERR_INVALID, "scores_out too small: need %lld", (long long)it.n_local);
    CU(cudaMemcpyAsync(scores_out, dw + o_score + it.score_off * 8, size_t(it.n_local) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(lane->h_down.p, dw + o_err, 4, cudaMemcpyDeviceToHost, st));
    rc = sync_stream(ctx, lane);
    if (rc) return rc;
    ST.d2h_bytes += size_t(it.n_local) * 8;
    if (scores_written) *scores_written = it.n_local;
    int e0; std::memcpy(&e0, lane->h_down.p, 4);
    if (e0 & kErrWindow) return fail(ctx, RSM_ERR_WINDOW, "search window + scan extent leaves the grid");
    return RSM_OK;
  }

  // A whole pass replayed as one CUDA graph once its shape has been seen twice: one launch call
  // instead of a copy, a memset, up to three kernels, two event pairs and a read-back.
  bool launched = false;
  // (batches too: a sub-batch's pass is five or six driver calls, ~25 us of host time on the lane's critical path three
  //  times per chain -- and host time is what bounds a rank with few cores)
  if (!ctx->profiling && env.get("NO_GRAPH") == nullptr) {
    std::vector<long long> key = {(long long)(intptr_t)dw, (long long)(intptr_t)up, (long long)(intptr_t)dn, (long long)up_bytes,
                                  (long long)o_best, (long long)zero_end, use_flat * (1 + 8 * flat_k) + 2 * (use_patch ? patch_nxy : 0) + 64 * beam_split, use_staged, any_fixed, cfg.affine, cfg.lx, cfg.ry,
                                  const_pitch, staged_variant, cta, na, (long long)o_sjobs, (long long)o_scta, n_launches, fork,
                                  total_sel_cta, (long long)o_ljobs, (long long)o_lcta, (long long)o_pool, pool_cap,
                                  (long long)o_poolcnt, (long long)head_bytes, pool_first, use_stream, stream_variant,
                                  (long long)splan->ctas.size(), (long long)o_splan, (long long)o_spart, (long long)o_stickets, (long long)zero_begin, max_V};
    for (int l = 0; l < n_launches; ++l) {
      const StagedLaunch& L = launches[l];
      key.insert(key.end(), {L.split, L.n_cta, L.beams, (long long)L.jobs_off, (long long)L.cta_off, L.n_jobs});
    }
    if (lane->graphs.size() > 64) {
      for (auto& g : lane->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
      lane->graphs.clear();
    }
    Lane::PassGraph& G = lane->graphs[key];
    // (one capture at a time in the process: lanes see a new shape together, and nothing is gained by capturing side by side)
    static std::mutex capture_mu;
    std::unique_lock<std::mutex> capture_lock(capture_mu, std::defer_lock);
    if (!G.exec && G.seen + 1 == 2) capture_lock.lock();
    if (!G.exec && ++G.seen == 2 && cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int r = enqueue_score();
      if (r == RSM_OK) r = enqueue_tail();
      cudaGraph_t graph = nullptr;
      const cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (r != RSM_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&G.exec, graph, 0) != cudaSuccess) {
        G.exec = nullptr;
        cudaGetLastError();
      }
      if (graph) cudaGraphDestroy(graph);
    }
    if (capture_lock.owns_lock()) capture_lock.unlock();
    if (G.exec) {
      CU(cudaGraphLaunch(G.exec, st));
      launched = true;
    }
  }
  if (!launched) {
    rc = enqueue_score();
    if (rc) return rc;
    rc = enqueue_tail();
    if (rc) return rc;
  }
  ST.kernel_launches++;
  R.pending = true;
  // batches sleep on the lane's event unless the host has cores to spare for every lane thread (a spinning wait
  // wakes ~50 us sooner: +10 % on a 512-pair batch; RSM_SPIN_WAIT=0 / 1 overrides)
  static const int spin_env = [] { const char* e = std::getenv("RSM_SPIN_WAIT"); return e ? std::atoi(e) : -1; }();
  const bool spare_cores = HostPool::env_threads() >= 2 * std::max(1, int(ctx->extra_lanes.size()) + 1);
  R.blocking = na > 8 && (spin_env < 0 ? !spare_cores : spin_env == 0);
  R.o_trig = o_trig;
  R.o_best = o_best; R.o_err = o_err; R.o_poolcnt = o_poolcnt; R.o_fcnt = o_fcnt; R.o_ftop = o_ftop;
  R.o_speccols = o_speccols; R.o_spec = o_spec; R.o_pool = o_pool; R.o_gjobs = o_gjobs; R.o_gout = o_gout; R.o_score = o_score;
  R.head_bytes = head_bytes; R.gather_doubles = gather_doubles; R.pool_first = pool_first; R.pool_cap = pool_cap;
  pt.lap(0);
  return RSM_OK;
}

// ---- wait for the lane, then the host side of the pass -----------------------------------------------
int pass_end(rsm_ctx* ctx, PassRun& R) {
  if (!R.pending) return RSM_OK;
  R.pending = false;
  Lane* lane = R.lane;
  rsm_stats& ST = lane->stats;
  PhaseTimer pt(&ST);
  std::vector<PassItem>& items = *R.items;
  const std::vector<int>& act = R.act;
  const int na = int(act.size());
  const PassMode mode = R.mode;
  cudaStream_t st = lane->stream;
  char* dw = lane->d_work.p;
  char* dn = lane->h_down.p;
  const size_t o_best = R.o_best, o_err = R.o_err, o_poolcnt = R.o_poolcnt, o_fcnt = R.o_fcnt, o_ftop = R.o_ftop,
               o_speccols = R.o_speccols, o_spec = R.o_spec, o_pool = R.o_pool, o_gjobs = R.o_gjobs, o_gout = R.o_gout,
               o_score = R.o_score, head_bytes = R.head_bytes, gather_doubles = R.gather_doubles;
  const int pool_first = R.pool_first, pool_cap = R.pool_cap;
  double* scores_out = R.scores_out;
  int rc = wait_lane(ctx, lane, R.blocking);
  if (rc) return rc;
  ST.d2h_bytes += head_bytes + size_t(pool_first) * sizeof(PoolEntry);
  const unsigned long long* h_best = reinterpret_cast<const unsigned long long*>(dn);
  const int* h_err = reinterpret_cast<const int*>(dn + (o_err - o_best));
  int pool_count = *reinterpret_cast<const int*>(dn + (o_poolcnt - o_best));
  const int* h_fcnt = reinterpret_cast<const int*>(dn + (o_fcnt - o_best));
  const Entry* h_ftop = reinterpret_cast<const Entry*>(dn + (o_ftop - o_best));
  const int* h_speccols = reinterpret_cast<const int*>(dn + (o_speccols - o_best));
  const double* h_spec = reinterpret_cast<const double*>(dn + (o_spec - o_best));
  const PoolEntry* h_pool = reinterpret_cast<const PoolEntry*>(dn + (o_pool - o_best));
  bool pool_overflow = pool_count > pool_cap;
  if (pool_overflow) pool_count = pool_cap;
  if (pool_count > pool_first) {
    CU(cudaMemcpyAsync(dn + head_bytes + size_t(pool_first) * sizeof(PoolEntry),
                       dw + o_pool + size_t(pool_first) * sizeof(PoolEntry),
                       size_t(pool_count - pool_first) * sizeof(PoolEntry), cudaMemcpyDeviceToHost, st));
    rc = sync_stream(ctx, lane);
    if (rc) return rc;
    ST.d2h_bytes += size_t(pool_count - pool_first) * sizeof(PoolEntry);
  }

  pt.lap(1);
  if (mode == MODE_PARTIAL) {
    // pack item 0's local result; the slice's scores stay where they are (ctx->slice)
    PassItem& it = items[act[0]];
    if (h_err[0] & kErrWindow) return fail(ctx, RSM_ERR_WINDOW, "search window + scan extent leaves the grid");
    const int64_t base = int64_t(it.a0) * it.geo.n_xy * it.geo.n_xy;
    PartialHeader H;
    std::memset(&H, 0, sizeof H);
    H.magic = kPartialMagic; H.a0 = it.a0; H.a1 = it.a1; H.n_ang = it.geo.n_ang; H.n_xy = it.geo.n_xy;
    H.best_key = h_best[0];
    const int cap = int((RSM_PARTIAL_BYTES - sizeof(PartialHeader)) / sizeof(Cand)) - kTopK;
    if ((h_err[0] & (kErrPoolFull | kErrSelectFull)) || pool_overflow || pool_count > cap) H.flags |= 1;
    Cand* out = reinterpret_cast<Cand*>(reinterpret_cast<char*>(scores_out) + sizeof(PartialHeader));
    if (!(H.flags & 1)) {
      for (int e = 0; e < pool_count; ++e) out[H.n_pool++] = Cand{h_pool[e].score, base + h_pool[e].index};
    }
    for (int r = 0; r < h_fcnt[0]; ++r) out[H.n_pool + H.n_top++] = Cand{h_ftop[r].score, base + h_ftop[r].index};
    std::memcpy(scores_out, &H, sizeof H);
    SliceState& S = ctx->slice;
    S.valid = true; S.merged = false; S.exact_needed = false;
    S.param = it.param; S.geo = it.geo; S.grid = it.grid; S.a0 = it.a0; S.a1 = it.a1;
    S.d_score = reinterpret_cast<double*>(dw + o_score) + it.score_off;
    S.d_gjob = dw + o_gjobs; S.d_gout = reinterpret_cast<double*>(dw + o_gout) + it.gather_off;
    return RSM_OK;
  }
  // ---- stage 1: best pose, positional covariance ---------------------------------------------
  for (int a = 0; a < na; ++a) {
    PassItem& it = items[act[a]];
    if (h_err[a] & kErrWindow) return fail(ctx, RSM_ERR_WINDOW, "search window + scan extent leaves the grid (item %d)", act[a]);
    it.a_list.clear(); it.top.clear(); it.xy.clear(); it.n_cols = 0;
    if ((h_err[a] & (kErrPoolFull | kErrSelectFull)) || pool_overflow) it.exact = true;
  }
  for (int e = 0; e < pool_count; ++e) {
    const PoolEntry& pe = h_pool[e];
    PassItem& it = items[act[pe.job]];
    if (it.exact) continue;
    it.a_list.push_back(Cand{pe.score, int64_t(it.a0) * it.geo.n_xy * it.geo.n_xy + pe.index});
  }
  const bool strict = ctx->strict_ties;
  // same-(x,y) entries of the gathered columns -> sorted list -> angular covariance; false = exact path needed
  auto angular_from_columns = [strict](PassItem& it, const double* vals, const int* where) -> bool {
    const PassGeo& g = it.geo;
    const double bound = cov_score_bound(it.best);
    const int nang = it.a1 - it.a0;
    it.xy.clear();
    for (int c = 0; c < it.n_cols; ++c)
      for (int ia = 0; ia < nang; ++ia) {
        const double s = vals[size_t(where ? where[c] : c) * nang + ia];
        if (s >= bound) it.xy.push_back(Cand{s, (int64_t(it.a0 + ia) * g.n_xy * g.n_xy) + it.cols[c]});
      }
    if (it.xy.size() > size_t(kTopK)) {
      std::nth_element(it.xy.begin(), it.xy.begin() + kTopK, it.xy.end(), by_score_desc);
      it.xy.resize(kTopK);
    }
    std::sort(it.xy.begin(), it.xy.end(), by_score_desc);
    if (it.xy.size() > size_t(kMaxVarianceUsePointSize) &&
        it.xy[kMaxVarianceUsePointSize].score == it.xy[kMaxVarianceUsePointSize - 1].score) return false;
    if (strict && adjacent_tie(it.xy, kMaxVarianceUsePointSize)) return false;
    angular_cov(g, it.param, it.best, it.xy.data(), it.xy.size(), it.cov);
    return true;
  };
  ctx->pool.run(na, 16, [&](int a_begin, int a_end) {
    for (int a = a_begin; a < a_end; ++a) {
      PassItem& it = items[act[a]];
      if (it.exact) continue;
      const PassGeo& g = it.geo;
      const double top_score = key_to_score(h_best[a]);
      std::sort(it.a_list.begin(), it.a_list.end(), by_score_desc);
      if (it.a_list.empty() || it.a_list[0].score != top_score || adjacent_tie(it.a_list, it.a_list.size())) { it.exact = true; continue; }
      // (the angle table of this pass is still in the lane's upload buffer: cos / sin of every search angle)
      it.best = find_best(g, it.a_list.data(), it.a_list.size(), reinterpret_cast<const double*>(lane->h_up.p + R.o_trig) + it.trig_off);
      if (size_t(it.best.n_avg) != it.a_list.size()) { it.exact = true; continue; }   // cannot happen; be safe
      const int64_t base = int64_t(it.a0) * g.n_xy * g.n_xy;
      // the job's top-kTopK as merged on the device (descending); their VALUES are unique whatever
      // the order of equal scores, and equal scores at the cut are detected below
      const Entry* ft = h_ftop + size_t(a) * kTopK;
      for (int r = 0; r < h_fcnt[a]; ++r) it.top.push_back(Cand{ft[r].score, base + ft[r].index});
      const int type = it.param.type;
      const double bound = cov_score_bound(it.best);
      if (type == RSM_COARSE || type == RSM_FINE) {
        // the 20-element prefix is unambiguous unless the 20th and 21st scores tie above the bound
        if (it.top.size() == size_t(kTopK) && it.top[kTopK - 1].score > bound &&
            it.top[kTopK - 1].score == it.top[kTopK - 2].score) { it.exact = true; continue; }
        // ties inside the prefix keep the set but not the reference's summation order (RSM_OPT_STRICT_TIES)
        if (strict && adjacent_tie(it.top, positional_prefix(it.top, bound))) { it.exact = true; continue; }
        positional_cov(g, it.param, it.best, it.top.data(), it.top.size(), it.cov);
      }
      if (type == RSM_COARSE || type == RSM_SUPER) {
        if (it.best.score < kDoubleTolerance) {
          angular_cov(g, it.param, it.best, nullptr, 0, it.cov);
        } else {
          int xs[8], ys[8], nx = 0, ny = 0;
          const double tol = g.factor;
          for (int ix = 0; ix < g.n_xy && nx < 8; ++ix) if (DoubleEqual(g.x_of(ix), it.best.x, tol)) xs[nx++] = ix;
          for (int iy = 0; iy < g.n_xy && ny < 8; ++iy) if (DoubleEqual(g.y_of(iy), it.best.y, tol)) ys[ny++] = iy;
          if (nx * ny > kMaxCols) { it.exact = true; continue; }
          it.n_cols = 0;
          for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) it.cols[it.n_cols++] = xs[i] * g.n_xy + ys[j];
          if (it.n_cols == 0) { angular_cov(g, it.param, it.best, nullptr, 0, it.cov); continue; }
          // the device already gathered the 3x3 neighbourhood of the top candidate: use it when it
          // covers every same-(x,y) column, else leave n_cols set for the gather round trip
          const int* sc_cols = h_speccols + size_t(a) * (kMaxCols + 1);
          const int nspec = sc_cols[0];
          int where[kMaxCols];
          bool covered = nspec > 0;
          for (int c = 0; c < it.n_cols && covered; ++c) {
            where[c] = -1;
            for (int q = 0; q < nspec; ++q) if (sc_cols[1 + q] == it.cols[c]) where[c] = q;
            if (where[c] < 0) covered = false;
          }
          if (!covered) continue;
          if (!angular_from_columns(it, h_spec + it.spec_off, where)) { it.n_cols = 0; it.exact = true; continue; }
          it.n_cols = 0;   // no round trip needed
        }
      }
    }
  });
  std::vector<GatherJob> gjobs;
  std::vector<int> gitem;
  for (int a = 0; a < na; ++a) {
    PassItem& it = items[act[a]];
    if (it.exact || it.n_cols == 0) continue;
    const PassGeo& g = it.geo;
    GatherJob G;
    std::memset(&G, 0, sizeof G);
    G.score = reinterpret_cast<const double*>(dw + o_score) + it.score_off;
    G.out = reinterpret_cast<double*>(dw + o_gout) + it.gather_off;
    G.n_xy = g.n_xy; G.n_ang = it.a1 - it.a0; G.n_cols = it.n_cols;
    for (int c = 0; c < it.n_cols; ++c) G.cols[c] = it.cols[c];
    gjobs.push_back(G); gitem.push_back(act[a]);
  }
  const bool need_gather = !gjobs.empty();
  pt.lap(2);
  // ---- stage 2: angular covariance from the same-(x,y) columns ---------------------------------
  if (need_gather) {
    const int ng = int(gjobs.size());
    std::memcpy(lane->h_up.p, gjobs.data(), sizeof(GatherJob) * ng);
    CU(cudaMemcpyAsync(dw + o_gjobs, lane->h_up.p, sizeof(GatherJob) * ng, cudaMemcpyHostToDevice, st));
    CU(launch_gather(ng, st, reinterpret_cast<const GatherJob*>(dw + o_gjobs)));
    ST.kernel_launches++;
    CU(cudaMemcpyAsync(lane->h_down.p, dw + o_gout, gather_doubles * 8, cudaMemcpyDeviceToHost, st));
    rc = sync_stream(ctx, lane);
    if (rc) return rc;
    ST.h2d_bytes += sizeof(GatherJob) * ng;
    ST.d2h_bytes += gather_doubles * 8;
    pt.lap(3);
    const double* h_g = reinterpret_cast<const double*>(lane->h_down.p);
    ctx->pool.run(ng, 16, [&](int g_begin, int g_end) {
      for (int gi = g_begin; gi < g_end; ++gi) {
        PassItem& it = items[gitem[gi]];
        if (!angular_from_columns(it, h_g + it.gather_off, nullptr)) it.exact = true;
      }
    });
  }
  pt.lap(4);
  // ---- exact path for the items that need the reference's own sort order ----------------------
  {
    std::vector<int> ex;
    std::vector<size_t> ex_off;
    size_t ex_doubles = 0;
    for (int a = 0; a < na; ++a)
      if (items[act[a]].exact) { ex.push_back(act[a]); ex_off.push_back(ex_doubles); ex_doubles += size_t(items[act[a]].n_local); }
    if (!ex.empty()) {
      rc = ensure_pinned(ctx, lane->h_down, ex_doubles * 8, st);
      if (rc) return rc;
      double* h_sc = reinterpret_cast<double*>(lane->h_down.p);
      for (size_t i = 0; i < ex.size(); ++i) {
        const PassItem& it = items[ex[i]];
        CU(cudaMemcpyAsync(h_sc + ex_off[i], dw + o_score + it.score_off * 8, size_t(it.n_local) * 8,
                           cudaMemcpyDeviceToHost, st));
      }
      rc = sync_stream(ctx, lane);
      if (rc) return rc;
      ST.d2h_bytes += ex_doubles * 8;
      ctx->pool.run(int(ex.size()), 1, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i) finish_exact(items[ex[i]], h_sc + ex_off[i]);
      });
      ST.exact_sort_passes += int64_t(ex.size());
    }
  }
  pt.lap(5);
  // ---- response, pose write-back (:861-869) ----------------------------------------------------
  for (int a = 0; a < na; ++a) {
    PassItem& it = items[act[a]];
    const double bs = it.best.score;
    it.response = bs > 1.0 ? 1.0 : bs;
    it.detail.best_score = bs;
    it.detail.best_pose_map[0] = it.best.x; it.detail.best_pose_map[1] = it.best.y; it.detail.best_pose_map[2] = it.best.angle;
    it.detail.n_candidates = it.geo.n_cand();
    it.detail.n_avg = it.best.n_avg;
    it.detail.exact_sort_used = it.exact ? 1 : 0;
    it.detail.n_ang = it.geo.n_ang; it.detail.n_xy = it.geo.n_xy;
    it.detail.visited = it.geo.visited; it.detail.divisor = it.geo.divisor;
    it.detail.pose_updated = 0;
    if (it.response > it.param.response_threshold) {
      const double b[3] = {it.best.x, it.best.y, it.best.angle};
      if (!it.center_map) it.grid->tf.map_to_world(b, it.pose_world);
      it.detail.pose_updated = 1;
    }
    ST.passes++;
  }
  return RSM_OK;
}

// groups of the active items that share a kernel plan, in first-appearance order
std::vector<std::vector<int>> plan_groups(const std::vector<PassItem>& items, const std::vector<int>& act) {
  std::vector<std::vector<int>> groups;
  for (int i : act) {
    bool placed = false;
    for (auto& g : groups) if (same_plan(items[g[0]], items[i])) { g.push_back(i); placed = true; break; }
    if (!placed) groups.push_back(std::vector<int>{i});
  }
  return groups;
}

// One pass over a batch of items on the context's own lane.
int run_pass(rsm_ctx* ctx, std::vector<PassItem>& items, PassMode mode, double* scores_out,
             int64_t scores_cap, int64_t* scores_written) {
  std::vector<int> act;
  int rc = pass_geometry(ctx, items, act);
  if (rc) return rc;
  if (scores_written) *scores_written = 0;
  if (act.empty()) return RSM_OK;
  if (mode != MODE_MATCH) {       // single-item modes
    PassRun R;
    rc = pass_begin(ctx, &ctx->L0, items, std::vector<int>{act[0]}, mode, scores_out, scores_cap, scores_written, R);
    if (rc) return rc;
    return pass_end(ctx, R);
  }
  for (const auto& g : plan_groups(items, act)) {
    PassRun R;
    rc = pass_begin(ctx, &ctx->L0, items, g, mode, nullptr, 0, nullptr, R);
    if (rc) return rc;
    rc = pass_end(ctx, R);
    if (rc) return rc;
  }
  return RSM_OK;
}

int upload_points(rsm_ctx* ctx, const double* pts_xy, size_t n_points, double** d_out) {
  const size_t bytes = n_points * 16;
  int rc = ensure_dev(ctx, ctx->d_pts, bytes);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ctx->d_pts.p, pts_xy, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += bytes;
  *d_out = reinterpret_cast<double*>(ctx->d_pts.p);
  return RSM_OK;
}

// How many sub-batches (lanes) a batched chain of n items is cut into.  Each sub-batch must still fill the GPU
// by itself for most of a kernel (a coarse pass of 128 chains is ~4 waves of the patch kernel's CTAs).
int chain_lanes(const rsm_ctx* ctx, int n) {
  int lanes = ctx->lanes_wanted;
  if (const char* e = std::getenv("RSM_LANES")) lanes = std::atoi(e);
  if (lanes <= 0) {
    lanes = n >= 384 ? 4 : n >= 192 ? 3 : n >= 64 ? 2 : 1;
    // eight sub-batches of 64 hide more of the host stages between the passes than four of 128 lose in kernel
    // efficiency (512 pairs on one B200: 2.91 -> 2.63 ms per step); with four host cores per rank (8 ranks on a 32-core
    // box) six do best (3.56 -> 3.44 ms: the step is bound by the host there)
    if (n >= 512) lanes = HostPool::env_threads() >= 6 ? 8 : 6;
  }
  return std::max(1, std::min({lanes, 8, n}));
}

// The coarse -> fine -> super-fine chain over a batch (scan_matchers.h:224-263, 281), software-pipelined: the
// items are cut into contiguous sub-batches, one lane (stream + buffers) each; while the host finalises pass k of
// one sub-batch and prepares its pass k+1, the kernels of the other sub-batches run.  `pre(lane, first, count)`
// (optional) enqueues work a sub-batch's first pass depends on -- the grid rasterisation of the batched back-end
// step -- on that lane's stream.
int run_chain(rsm_ctx* ctx, int n, const rsm_grid* const* grids, double* const* d_pts, const int* n_pts,
              const rsm_pass_param* params, bool shared_params, bool use_fine, double* poses, double* covs,
              double* scores, double* responses, const std::function<int(Lane*, int, int)>* pre = nullptr,
              const unsigned char* skip_first = nullptr) {   // skip_first[i]: item i does not run pass 0 (its response is 0)
  std::vector<PassItem> items(n);
  std::vector<double> sum(n, 0.0);
  const int n_pass = use_fine ? 3 : 1;
  auto set_pass = [&](int i, int pass) {
    PassItem& it = items[i];
    it.grid = (skip_first && skip_first[i] && pass == 0) ? nullptr : grids[i]; it.d_pts = d_pts[i]; it.P = n_pts[i];
    it.param = params[(shared_params ? 0 : 3 * i) + pass];
    it.pose_world = poses + 3 * i; it.cov = covs + 9 * i;
    it.ang_begin = 0; it.ang_end = -1;
  };
  auto account = [&](int pass) {
    for (int i = 0; i < n; ++i) {
      sum[i] += items[i].response;
      if (responses) responses[3 * i + pass] = items[i].response;
    }
  };
  int n_lanes = chain_lanes(ctx, n);
  // the lanes' launches assume one kernel plan per sub-batch: same parameters and the same kind of grid for every item
  const rsm_grid* g0 = nullptr;
  for (int i = 0; i < n && n_lanes > 1; ++i) {
    if (!grids[i]) continue;
    if (!g0) g0 = grids[i];
    else if (grids[i]->scale != g0->scale || grids[i]->fixed != g0->fixed) n_lanes = 1;
  }
  if (n_lanes <= 1 || !shared_params) {
    // one lane (small batches), or windows that may differ from item to item: pass after pass, grouped by plan
    if (pre) { int rc = (*pre)(&ctx->L0, 0, n); if (rc) return rc; }
    for (int pass = 0; pass < n_pass; ++pass) {
      for (int i = 0; i < n; ++i) set_pass(i, pass);
      int rc = run_pass(ctx, items, MODE_MATCH, nullptr, 0, nullptr);
      if (rc) return rc;
      account(pass);
    }
  } else {
    // One host thread per lane runs that sub-batch's whole chain: launch pass k, sleep on the lane's event, finalise,
    // launch pass k + 1.  The lanes' kernels share the GPU stream by stream, so one lane's host stage overlaps the
    // others' kernels without any lane waiting for another.
    struct Sub { Lane* lane = nullptr; int first = 0, count = 0, rc = RSM_OK; std::vector<PassItem> items; PassRun run; };
    std::vector<Sub> subs(n_lanes);
    int rc_all = RSM_OK;
    for (int l = 0; l < n_lanes && rc_all == RSM_OK; ++l) {
      Sub& S = subs[l];
      S.first = int((long long)n * l / n_lanes);
      S.count = int((long long)n * (l + 1) / n_lanes) - S.first;
      S.items.resize(S.count);
      rc_all = get_lane(ctx, l, &S.lane);
      // lanes 1.. must see what the caller enqueued on the context's stream (point uploads, grid writes)
      if (rc_all == RSM_OK && l > 0) {
        if (cudaEventRecord(ctx->L0.ev_fork, ctx->stream) != cudaSuccess || cudaStreamWaitEvent(S.lane->stream, ctx->L0.ev_fork, 0) != cudaSuccess)
          rc_all = fail(ctx, RSM_ERR_CUDA, "lane fork failed");
      }
    }
    if (rc_all != RSM_OK) return rc_all;
    const int device = ctx->device;
    auto lane_chain = [&](Sub& S) {
      cudaSetDevice(device);
      if (pre) { S.rc = (*pre)(S.lane, S.first, S.count); if (S.rc) return; }
      for (int pass = 0; pass < n_pass; ++pass) {
        for (int k = 0; k < S.count; ++k) {
          PassItem& it = S.items[k];
          const int i = S.first + k;
          it.grid = (skip_first && skip_first[i] && pass == 0) ? nullptr : grids[i]; it.d_pts = d_pts[i]; it.P = n_pts[i];
          it.param = params[pass];                      // shared parameters on this path
          it.pose_world = poses + 3 * i; it.cov = covs + 9 * i;
          it.ang_begin = 0; it.ang_end = -1;
        }
        std::vector<int> act;
        S.rc = pass_geometry(ctx, S.items, act);
        if (S.rc == RSM_OK) S.rc = pass_begin(ctx, S.lane, S.items, act, MODE_MATCH, nullptr, 0, nullptr, S.run);
        if (S.rc == RSM_OK) S.rc = pass_end(ctx, S.run);
        if (S.rc) return;
        for (int k = 0; k < S.count; ++k) {
          sum[S.first + k] += S.items[k].response;
          if (responses) responses[3 * (S.first + k) + pass] = S.items[k].response;
        }
      }
    };
    ctx->lane_pool.set_threads(8);
    ctx->lane_pool.run(n_lanes, 1, [&](int l0, int l1) {
      const bool was = t_lane_worker;
      t_lane_worker = true;
      for (int l = l0; l < l1; ++l) lane_chain(subs[l]);
      t_lane_worker = was;
    });
    for (int l = 0; l < n_lanes; ++l) if (subs[l].rc != RSM_OK && rc_all == RSM_OK) rc_all = subs[l].rc;
    if (rc_all != RSM_OK) {
      // leave no launch in flight that reads this call's buffers
      for (int l = 0; l < n_lanes; ++l) cudaStreamSynchronize(subs[l].lane->stream);
      return rc_all;
    }
    // the context's stream continues after every lane (later calls on it may rewrite the grids / points)
    for (int l = 1; l < n_lanes; ++l) {
      if (cudaEventRecord(subs[l].lane->ev_join, subs[l].lane->stream) != cudaSuccess ||
          cudaStreamWaitEvent(ctx->stream, subs[l].lane->ev_join, 0) != cudaSuccess)
        return fail(ctx, RSM_ERR_CUDA, "lane join failed");
    }
  }
  for (int i = 0; i < n; ++i) {
    if (responses && !use_fine) { responses[3 * i + 1] = 0.0; responses[3 * i + 2] = 0.0; }
    scores[i] = sum[i] / n_pass;
  }
  return RSM_OK;
}

// MapFeedbackResponsePenalty (map/occu_grid_map.h:331-392) + MapCheckPenalize (slam/slam_processor.cpp:573-595) for n
// poses in one launch.  Points of pose i: host range pts_xy[2*pts_begin[i]..] (uploaded here, [lo, hi) = the hull of the
// ranges) or, when dev_pts is given, the device-resident scan dev_pts[i].
int map_check_core(rsm_ctx* ctx, const rsm_grid* pub_map, int n, const double* poses_world, const double* pts_xy,
                   int64_t lo, int64_t hi, const int64_t* pts_begin, const double* const* dev_pts, const int32_t* pts_count,
                   const double* sensor_origin_xy, int check_point_num, double bound_tolerance, double penalty_gain,
                   int use_logistic, double* coeff_out) {
  // occu_grid_map.h:337-341: out-of-range knobs switch the check off
  if (bound_tolerance < 0 || check_point_num <= 0 || penalty_gain <= 0.0 || penalty_gain >= 1.0) {
    for (int i = 0; i < n; ++i) coeff_out[i] = use_logistic ? (1 / (1 + std::exp(-10 * (1.0 - 0.4)))) : 1.0;
    return RSM_OK;
  }
  for (int i = 0; i < n; ++i)
    if (check_point_num == 1 && pts_count[i] >= 2)   // the reference divides by zero (occu_grid_map.h:367)
      return fail(ctx, RSM_ERR_INVALID, "rsm_map_check_penalize: check_point_num = 1 with 2 or more points");
  const double ox = sensor_origin_xy ? sensor_origin_xy[0] : 0.0, oy = sensor_origin_xy ? sensor_origin_xy[1] : 0.0;
  const size_t pts_bytes = dev_pts ? 0 : size_t(hi - lo) * 16;
  Layout dl;
  const size_t o_jobs = dl.take(sizeof(PenaltyJob) * size_t(n));
  const size_t o_pts = dl.take(pts_bytes, 16);
  const size_t up_bytes = dl.off;
  const size_t o_out = dl.take(size_t(n) * 4, 16);
  int rc = ensure_dev(ctx, ctx->d_work, dl.off);
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_up, up_bytes);
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_down, size_t(n) * 4);
  if (rc) return rc;
  char* dw = ctx->d_work.p;
  char* up = ctx->h_up.p;
  if (pts_bytes) std::memcpy(up + o_pts, pts_xy + 2 * lo, pts_bytes);
  PenaltyJob* jobs = reinterpret_cast<PenaltyJob*>(up + o_jobs);
  std::vector<char> outside(n, 0);
  for (int i = 0; i < n; ++i) {
    PenaltyJob& J = jobs[i];
    std::memset(&J, 0, sizeof J);
    double pm[3];
    pub_map->tf.world_to_map(poses_world + 3 * i, pm);                                   // :349
    // PointInMap(x, y) is strict on both sides (grid_map_base.h:339-346); outside -> 0.0 (:351-353)
    outside[i] = !(pm[0] > 0.0 && pm[0] < pub_map->size_x && pm[1] > 0.0 && pm[1] < pub_map->size_y);
    const double c = std::cos(pm[2]), sn = std::sin(pm[2]);                              // Rotation2Dd, host libm
    J.c = c; J.s = sn; J.tx = pm[0]; J.ty = pm[1];
    J.sx0 = static_cast<int>((pm[0] + (c * ox + (-sn) * oy)) + 0.5);                      // :357-359
    J.sy0 = static_cast<int>((pm[1] + (sn * ox + c * oy)) + 0.5);
    const int all = pts_count[i];
    J.n_pts = outside[i] ? 0 : all;
    J.step = (all < 2 * check_point_num) ? 1 : all / (check_point_num - 1);                // :361-368
    J.pts = dev_pts ? dev_pts[i] : reinterpret_cast<const double*>(dw + o_pts) + 2 * (pts_begin[i] - lo);
    J.blocked = reinterpret_cast<int*>(dw + o_out) + i;
  }
  CU(cudaMemcpyAsync(dw, up, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_penalty(n, ctx->stream, reinterpret_cast<const PenaltyJob*>(dw + o_jobs), pub_map->d_occ, pub_map->size_x,
                    pub_map->size_y, bound_tolerance));
  CU(cudaMemcpyAsync(ctx->h_down.p, dw + o_out, size_t(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_stream(ctx);
  if (rc) return rc;
  ctx->stats.h2d_bytes += up_bytes; ctx->stats.d2h_bytes += size_t(n) * 4; ctx->stats.kernel_launches++;
  const int* blocked = reinterpret_cast<const int*>(ctx->h_down.p);
  for (int i = 0; i < n; ++i) {
    double coeff;
    if (outside[i]) coeff = 0.0;
    else {
      double penalty = double(blocked[i]);            // the reference adds 1.0 per blocked ray
      penalty *= penalty_gain;
      coeff = std::max((1.0 + 2 * penalty_gain - penalty), 0.1);                           // :389-390
    }
    if (use_logistic) coeff = (1 / (1 + std::exp(-10 * (coeff - 0.4))));                   // slam_processor.cpp:589-591
    coeff_out[i] = coeff;
  }
  return RSM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* rsm_version(void) { return "rsm 0.1 (sm_100a)"; }

int rsm_create(int device, rsm_ctx** out) {
  if (!out) return RSM_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return RSM_ERR_NO_DEVICE;
  }
  rsm_ctx* ctx = new rsm_ctx;
  std::memset(&ctx->stats, 0, sizeof ctx->stats);
  ctx->device = device;
  Lane& L = ctx->L0;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&L.stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&L.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&L.ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&L.done, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&ctx->t0) != cudaSuccess || cudaEventCreate(&ctx->t1) != cudaSuccess) {
    destroy_lane(L);
    delete ctx;
    return RSM_ERR_CUDA;
  }
  *out = ctx;
  return RSM_OK;
}

void rsm_destroy(rsm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (Lane* L : ctx->extra_lanes) { cudaStreamSynchronize(L->stream); destroy_lane(*L); delete L; }
  rsm_comm_destroy(ctx);
  Buf* dev[] = {&ctx->d_pts, &ctx->d_flush, &ctx->d_pool_grids, &ctx->d_pool_grids2, &ctx->d_xchg};
  for (Buf* b : dev) if (b->p) cudaFree(b->p);
  if (ctx->h_xchg.p) cudaFreeHost(ctx->h_xchg.p);
  cudaEventDestroy(ctx->t0); cudaEventDestroy(ctx->t1);
  destroy_lane(ctx->L0);
  delete ctx;
}

const char* rsm_last_error(const rsm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int rsm_set_profiling(rsm_ctx* ctx, int on) { if (!ctx) return RSM_ERR_INVALID; ctx->profiling = on != 0; return RSM_OK; }
int rsm_set_option(rsm_ctx* ctx, int option, int value) {
  if (!ctx) return RSM_ERR_INVALID;
  switch (option) {
    case RSM_OPT_STRICT_TIES: ctx->strict_ties = value != 0; return RSM_OK;
    case RSM_OPT_LANES: if (value < 0 || value > 8) break; ctx->lanes_wanted = value; return RSM_OK;
    default: break;
  }
  return fail(ctx, RSM_ERR_INVALID, "rsm_set_option: unknown option %d or value %d out of range", option, value);
}
int rsm_get_stats(rsm_ctx* ctx, rsm_stats* out) {
  if (!ctx || !out) return RSM_ERR_INVALID;
  merge_lane_stats(ctx, &ctx->L0);
  for (Lane* L : ctx->extra_lanes) merge_lane_stats(ctx, L);
  *out = ctx->stats;
  return RSM_OK;
}
int rsm_reset_stats(rsm_ctx* ctx) {
  if (!ctx) return RSM_ERR_INVALID;
  std::memset(&ctx->stats, 0, sizeof ctx->stats);
  std::memset(&ctx->L0.stats, 0, sizeof ctx->L0.stats);
  for (Lane* L : ctx->extra_lanes) std::memset(&L->stats, 0, sizeof L->stats);
  return RSM_OK;
}
int rsm_synchronize(rsm_ctx* ctx) { if (!ctx) return RSM_ERR_INVALID; DeviceGuard device_guard(ctx); return sync_stream(ctx); }

int rsm_timer_start(rsm_ctx* ctx) {
  DeviceGuard device_guard(ctx);
  if (!ctx) return RSM_ERR_INVALID;
  CU(cudaEventRecord(ctx->t0, ctx->stream));
  return RSM_OK;
}
int rsm_timer_stop(rsm_ctx* ctx, double* elapsed_ms) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !elapsed_ms) return RSM_ERR_INVALID;
  CU(cudaEventRecord(ctx->t1, ctx->stream));
  CU(cudaEventSynchronize(ctx->t1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ctx->t0, ctx->t1));
  *elapsed_ms = ms;
  return RSM_OK;
}

int rsm_flush_l2(rsm_ctx* ctx) {
  DeviceGuard device_guard(ctx);
  if (!ctx) return RSM_ERR_INVALID;
  const size_t bytes = size_t(256) << 20;
  int rc = ensure_dev(ctx, ctx->d_flush, bytes);
  if (rc) return rc;
  CU(launch_flush(ctx->stream, ctx->d_flush.p, (long long)bytes, 1));
  return sync_stream(ctx);
}

// ---- grids -----------------------------------------------------------------------------------
int rsm_grid_create(rsm_ctx* ctx, int size_x, int size_y, double resolution, double offset_x, double offset_y, rsm_grid** out) {
  if (!(resolution > 0)) return fail(ctx, RSM_ERR_INVALID, "rsm_grid_create: bad resolution");
  int rc = rsm_grid_create_from_scale(ctx, size_x, size_y, 1.0 / resolution, offset_x, offset_y, out);   // map/grid_map_base.h:50
  if (rc == RSM_OK) (*out)->resolution = resolution;
  return rc;
}

int rsm_grid_create_from_scale(rsm_ctx* ctx, int size_x, int size_y, double scale_factor, double offset_x, double offset_y, rsm_grid** out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !out || size_x <= 0 || size_y <= 0 || size_x > 32768 || size_y > 32768 || !(scale_factor > 0))
    return fail(ctx, RSM_ERR_INVALID, "rsm_grid_create: bad arguments");
  rsm_grid* g = new rsm_grid;
  g->size_x = size_x; g->size_y = size_y; g->pitch = grid_pitch_for(size_x);
  g->resolution = 1 / scale_factor;
  g->scale = scale_factor;
  g->off_x = offset_x; g->off_y = offset_y;
  g->tf.set(g->scale, offset_x, offset_y);
  const size_t bytes = size_t(g->pitch) * size_y * 4;   // pitch is a multiple of 32 cells
  cudaError_t e = cudaMalloc(&g->d_cells, bytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(g->d_cells, 0, bytes, ctx->stream);   // padding columns are never read, keep them defined
  if (e != cudaSuccess) { delete g; return fail(ctx, RSM_ERR_CUDA, "cudaMalloc(grid) failed: %s", cudaGetErrorString(e)); }
  *out = g;
  return RSM_OK;
}

void rsm_grid_destroy(rsm_ctx* ctx, rsm_grid* grid) {
  DeviceGuard device_guard(ctx);
  if (!grid) return;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  if (grid->owned && grid->d_cells) cudaFree(grid->d_cells);
  if (grid->d_occ) cudaFree(grid->d_occ);
  delete grid;
}

int rsm_grid_set_offset(rsm_ctx* ctx, rsm_grid* grid, double offset_x, double offset_y) {
  if (!grid) return fail(ctx, RSM_ERR_INVALID, "null grid");
  grid->off_x = offset_x; grid->off_y = offset_y;
  grid->tf.set(grid->scale, offset_x, offset_y);   // set_map_offset -> SetMapTransform (grid_map_base.h:285-289)
  return RSM_OK;
}

int rsm_grid_upload_f32(rsm_ctx* ctx, rsm_grid* grid, const float* prob) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !prob) return fail(ctx, RSM_ERR_INVALID, "rsm_grid_upload_f32: null argument");
  const size_t n = size_t(grid->size_x) * grid->size_y;
  int rc = ensure_pinned(ctx, ctx->h_up, n * 4);
  if (rc) return rc;
  int* fx = reinterpret_cast<int*>(ctx->h_up.p);
  bool ok = true;
  for (size_t i = 0; i < n; ++i) if (!fix_ok(prob[i], &fx[i])) { ok = false; break; }
  if (!ok) std::memcpy(ctx->h_up.p, prob, n * 4);
  grid->fixed = ok;
  CU(cudaMemcpy2DAsync(grid->d_cells, size_t(grid->pitch) * 4, ctx->h_up.p, size_t(grid->size_x) * 4,
                       size_t(grid->size_x) * 4, size_t(grid->size_y), cudaMemcpyHostToDevice, ctx->stream));
  rc = sync_stream(ctx);
  if (rc) return rc;
  ctx->stats.h2d_bytes += n * 4;
  grid->init = true;
  return RSM_OK;
}

int rsm_grid_download_f32(rsm_ctx* ctx, rsm_grid* grid, float* prob_out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !prob_out) return fail(ctx, RSM_ERR_INVALID, "rsm_grid_download_f32: null argument");
  const size_t n = size_t(grid->size_x) * grid->size_y;
  int rc = ensure_pinned(ctx, ctx->h_down, n * 4);
  if (rc) return rc;
  CU(cudaMemcpy2DAsync(ctx->h_down.p, size_t(grid->size_x) * 4, grid->d_cells, size_t(grid->pitch) * 4,
                       size_t(grid->size_x) * 4, size_t(grid->size_y), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_stream(ctx);
  if (rc) return rc;
  ctx->stats.d2h_bytes += n * 4;
  if (grid->fixed) {
    const int* fx = reinterpret_cast<const int*>(ctx->h_down.p);
    for (size_t i = 0; i < n; ++i) prob_out[i] = static_cast<float>(std::ldexp(static_cast<double>(fx[i]), -kFixShift));
  } else {
    std::memcpy(prob_out, ctx->h_down.p, n * 4);
  }
  return RSM_OK;
}

int rsm_grid_upload_occupancy(rsm_ctx* ctx, rsm_grid* grid, const uint8_t* occupied) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !occupied) return fail(ctx, RSM_ERR_INVALID, "rsm_grid_upload_occupancy: null argument");
  const size_t n = size_t(grid->size_x) * grid->size_y;
  if (!grid->d_occ) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&grid->d_occ), n);
    if (e != cudaSuccess) return fail(ctx, RSM_ERR_CUDA, "cudaMalloc(occupancy) failed: %s", cudaGetErrorString(e));
  }
  int rc = ensure_pinned(ctx, ctx->h_up, n);
  if (rc) return rc;
  std::memcpy(ctx->h_up.p, occupied, n);
  CU(cudaMemcpyAsync(grid->d_occ, ctx->h_up.p, n, cudaMemcpyHostToDevice, ctx->stream));
  rc = sync_stream(ctx);           // the staging buffer is reused by the next call
  if (rc) return rc;
  ctx->stats.h2d_bytes += n;
  return RSM_OK;
}

// MapFeedbackResponsePenalty + MapCheckPenalize for n poses in one launch.
int rsm_map_check_penalize(rsm_ctx* ctx, const rsm_grid* pub_map, int n, const double* poses_world, const double* pts_xy,
                           const int64_t* pts_begin, const int32_t* pts_count, const double* sensor_origin_xy,
                           int check_point_num, double bound_tolerance, double penalty_gain, int use_logistic,
                           double* coeff_out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !pub_map || n < 0 || (n > 0 && (!poses_world || !pts_begin || !pts_count || !coeff_out)))
    return fail(ctx, RSM_ERR_INVALID, "rsm_map_check_penalize: bad arguments");
  if (n == 0) return RSM_OK;
  if (!pub_map->d_occ) return fail(ctx, RSM_ERR_NOT_INIT, "rsm_map_check_penalize: no occupancy uploaded for this map");
  int64_t lo = INT64_MAX, hi = 0;
  for (int i = 0; i < n; ++i) {
    if (pts_count[i] < 0 || pts_begin[i] < 0) return fail(ctx, RSM_ERR_INVALID, "rsm_map_check_penalize: bad point range");
    if (pts_count[i] == 0) continue;
    lo = std::min(lo, pts_begin[i]);
    hi = std::max(hi, pts_begin[i] + pts_count[i]);
  }
  if (hi <= lo) { lo = 0; hi = 0; }
  if (hi > lo && !pts_xy) return fail(ctx, RSM_ERR_INVALID, "rsm_map_check_penalize: pts_xy is null");
  return map_check_core(ctx, pub_map, n, poses_world, pts_xy, lo, hi, pts_begin, nullptr, pts_count, sensor_origin_xy,
                        check_point_num, bound_tolerance, penalty_gain, use_logistic, coeff_out);
}

int rsm_grid_is_fixed_point(const rsm_grid* grid) { return grid && grid->fixed ? 1 : 0; }

int rsm_world_to_map(const rsm_grid* grid, const double pose_world[3], double pose_map[3]) {
  if (!grid) return RSM_ERR_INVALID;
  grid->tf.world_to_map(pose_world, pose_map);
  return RSM_OK;
}
int rsm_map_to_world(const rsm_grid* grid, const double pose_map[3], double pose_world[3]) {
  if (!grid) return RSM_ERR_INVALID;
  grid->tf.map_to_world(pose_map, pose_world);
  return RSM_OK;
}

}  // extern "C"

namespace {

// GaussianBlur (map/occu_grid_map.h:40-59, 83-105): -1 if the parameters are rejected
int blur_kernel(double sigma, double resolution, std::vector<double>& k) {
  const double lo = 0.5 * resolution, hi = 10 * resolution;
  if (!(sigma > lo && sigma < hi && resolution > 0)) return -1;
  const int half = static_cast<int>((sigma / resolution) * std::sqrt(std::log(2)));
  const int n = 2 * half + 1;
  k.assign(size_t(n) * n, 0.0);
  for (int i = -half; i <= half; ++i)
    for (int j = -half; j <= half; ++j) {
      const double d = std::hypot(i * resolution, j * resolution);
      const double q = d / sigma;
      k[(i + half) + n * (j + half)] = std::exp(-0.5 * (q * q));
    }
  return half;
}

struct RasterPlan {
  bool fixed = true;
  bool blur = true;      // false: SET_CELL_OCCUPIED (use_blur off, or blur parameters the reference rejects)
  int half = 0, one = 0, fill = 0;
  std::vector<int> stamp;
};

// Decide the cell representation for a rasterised grid and build the stamp patterns.
int plan_raster(rsm_ctx* ctx, float default_prob, double sigma, double resolution, double occu_offset, int use_blur, RasterPlan& pl) {
  std::vector<double> k;
  int half = blur_kernel(sigma, resolution, k);
  if (!(default_prob >= 0.0f)) return fail(ctx, RSM_ERR_INVALID, "default_prob must be >= 0");
  // GaussianBlur rejected the parameters: half_kernel_size_ = 0 and UpdateMapByRange drops use_blur (occu_grid_map.h:47-59, 265-268)
  if (half < 0) { half = 0; use_blur = 0; }
  pl.half = half;
  pl.blur = use_blur != 0;
  int tmp;
  const float onef = 1.0f;
  if (!pl.blur) {
    // every value SetCellOccu can write is fl(v + 0.5f) clamped to 1: a 2^-25 multiple whenever v is one
    pl.fixed = fix_ok(default_prob, &tmp);
    pl.stamp.assign(1, 0);
    if (pl.fixed) { pl.one = kFixOne; fix_ok(default_prob, &pl.fill); }
    else { std::memcpy(&pl.one, &onef, 4); std::memcpy(&pl.fill, &default_prob, 4); }
    return RSM_OK;
  }
  const int ks = 2 * half + 1;
  std::vector<float> probs(size_t(ks) * ks);
  for (size_t i = 0; i < probs.size(); ++i) probs[i] = static_cast<float>(k[i] * occu_offset);   // :567
  bool fixed = fix_ok(default_prob, &tmp);
  for (float p : probs) if (p <= 1.0f && !(p >= 0.0f && fix_ok(p, &tmp))) fixed = false;
  pl.fixed = fixed;
  pl.stamp.assign(probs.size(), 0);
  for (size_t i = 0; i < probs.size(); ++i) {
    const float p = probs[i];
    if (!(p <= 1.0f) || !(p > 0.0f)) continue;   // prob > 1 is ignored by SetGridProbability; <= 0 never raises a cell
    if (fixed) fix_ok(p, &pl.stamp[i]); else std::memcpy(&pl.stamp[i], &p, 4);
  }
  if (fixed) { pl.one = kFixOne; fix_ok(default_prob, &pl.fill); }
  else { std::memcpy(&pl.one, &onef, 4); std::memcpy(&pl.fill, &default_prob, 4); }
  return RSM_OK;
}

// Stamp n_scans scans (device descriptors at d_scans) on stream st: blur = max-compositing, one CTA per scan in any
// order; non-blur = one CTA per grid (scans [group_begin[g], group_begin[g+1]) in order; d_groups on the device).
cudaError_t enqueue_raster(const RasterPlan& pl, cudaStream_t st, int n_scans, const RasterScan* d_scans, const int* d_stamp,
                           int n_groups, const int* d_groups) {
  if (pl.blur) return launch_raster(n_scans, st, d_scans, d_stamp, pl.half, pl.one);
  return launch_raster_occu(n_groups, st, d_scans, d_groups, pl.half, pl.fixed ? 1 : 0);
}

// Fill the RasterScan of one base scan (host libm cos/sin = Eigen::Rotation2Dd, occu_grid_map.h:278-303)
void make_raster_scan(const rsm_grid* g, const double* pose_world, const double* d_pts, int n_pts, RasterScan& S,
                      bool pose_in_map = false) {
  double pm[3];
  if (pose_in_map) { pm[0] = pose_world[0]; pm[1] = pose_world[1]; pm[2] = pose_world[2]; }
  else g->tf.world_to_map(pose_world, pm);
  const double c = std::cos(pm[2]), s = std::sin(pm[2]);
  S.grid = g->d_cells; S.pts = d_pts; S.n_pts = n_pts;
  S.pitch = g->pitch; S.size_x = g->size_x; S.size_y = g->size_y;
  S.c = c; S.s = s; S.tx = pm[0]; S.ty = pm[1];
  const double ox = pm[0] + (c * 0.0 + (-s) * 0.0), oy = pm[1] + (s * 0.0 + c * 0.0);
  S.start_x = static_cast<int>(ox + 0.5);
  S.start_y = static_cast<int>(oy + 0.5);
}

}  // namespace

extern "C" {

int rsm_grid_rasterize(rsm_ctx* ctx, rsm_grid* grid, float default_prob, double sigma, double occu_offset, int use_blur,
                       int n_scans, const int32_t* n_pts, const double* pts_xy, const double* poses_world) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || n_scans < 0 || (n_scans > 0 && (!n_pts || !pts_xy || !poses_world)))
    return fail(ctx, RSM_ERR_INVALID, "rsm_grid_rasterize: bad arguments");
  RasterPlan pl;
  int rc = plan_raster(ctx, default_prob, sigma, grid->resolution, occu_offset, use_blur, pl);
  if (rc) return rc;
  size_t total = 0;
  for (int s = 0; s < n_scans; ++s) total += size_t(n_pts[s]);
  Layout dl;
  const size_t o_fill = dl.take(sizeof(FillJob));
  const size_t o_scans = dl.take(sizeof(RasterScan) * std::max(1, n_scans));
  const size_t o_stamp = dl.take(pl.stamp.size() * 4);
  const size_t o_groups = dl.take(2 * sizeof(int), 4);
  const size_t up_bytes = dl.off;
  rc = ensure_dev(ctx, ctx->d_work, dl.off);
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_up, up_bytes);
  if (rc) return rc;
  double* d_pts = nullptr;
  if (total) { rc = upload_points(ctx, pts_xy, total, &d_pts); if (rc) return rc; }
  char* up = ctx->h_up.p;
  char* dw = ctx->d_work.p;
  FillJob F; F.grid = grid->d_cells; F.n_cells = (long long)grid->pitch * grid->size_y; F.value = pl.fill;
  std::memcpy(up + o_fill, &F, sizeof F);
  RasterScan* hs = reinterpret_cast<RasterScan*>(up + o_scans);
  size_t off = 0;
  for (int s = 0; s < n_scans; ++s) {
    make_raster_scan(grid, poses_world + 3 * s, d_pts + 2 * off, n_pts[s], hs[s]);
    off += n_pts[s];
  }
  std::memcpy(up + o_stamp, pl.stamp.data(), pl.stamp.size() * 4);
  const int groups[2] = {0, n_scans};
  std::memcpy(up + o_groups, groups, sizeof groups);
  CU(cudaMemcpyAsync(dw, up, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += up_bytes;
  {
    Prof p(ctx, KC_RASTER);
    CU(launch_fill(1, 148, ctx->stream, reinterpret_cast<const FillJob*>(dw + o_fill)));
    if (n_scans > 0)
      CU(enqueue_raster(pl, ctx->stream, n_scans, reinterpret_cast<const RasterScan*>(dw + o_scans),
                        reinterpret_cast<const int*>(dw + o_stamp), 1, reinterpret_cast<const int*>(dw + o_groups)));
  }
  ctx->stats.kernel_launches += (n_scans > 0) ? 2 : 1;
  rc = sync_stream(ctx);
  if (rc) return rc;
  grid->fixed = pl.fixed;
  if (n_scans > 0) grid->init = true;   // SetUpdated() per UpdateMapByRange (occu_grid_map.h:325)
  return RSM_OK;
}

// ---- front-end map maintenance: the scan-match maps stay on the device between scans ----
int rsm_grid_fill(rsm_ctx* ctx, rsm_grid* grid, float fill_prob, float first_cell_prob) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !(fill_prob >= 0.0f) || !(first_cell_prob >= 0.0f)) return fail(ctx, RSM_ERR_INVALID, "rsm_grid_fill: bad arguments");
  int vf = 0, v0 = 0;
  const bool fixed = fix_ok(fill_prob, &vf) && fix_ok(first_cell_prob, &v0);
  if (!fixed) { std::memcpy(&vf, &fill_prob, 4); std::memcpy(&v0, &first_cell_prob, 4); }
  int rc = ensure_dev(ctx, ctx->d_work, sizeof(FillJob) * 2);
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_up, sizeof(FillJob) * 2);
  if (rc) return rc;
  FillJob* F = reinterpret_cast<FillJob*>(ctx->h_up.p);
  F[0].grid = grid->d_cells; F[0].n_cells = (long long)grid->pitch * grid->size_y; F[0].value = vf;
  F[1].grid = grid->d_cells; F[1].n_cells = 1; F[1].value = v0;      // new CellType[n]{default}: only cell 0 gets it
  CU(cudaMemcpyAsync(ctx->d_work.p, F, sizeof(FillJob) * 2, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_fill(1, 148, ctx->stream, reinterpret_cast<const FillJob*>(ctx->d_work.p)));
  CU(launch_fill(1, 1, ctx->stream, reinterpret_cast<const FillJob*>(ctx->d_work.p) + 1));
  ctx->stats.kernel_launches += 2; ctx->stats.h2d_bytes += sizeof(FillJob) * 2;
  rc = sync_stream(ctx);
  if (rc) return rc;
  grid->fixed = fixed;
  grid->init = false;       // a constructed map is not IsMapInit() until its first update (grid_map_base.h:311-317)
  return RSM_OK;
}

}  // extern "C"
namespace {
int grid_update_core(rsm_ctx* ctx, rsm_grid* grid, double sigma, double occu_offset, int use_blur, const double* pts_xy, int n_pts,
                     const double pose_world[3], bool pose_in_map);
}
extern "C" {
int rsm_grid_update_by_range(rsm_ctx* ctx, rsm_grid* grid, double sigma, double occu_offset, int use_blur,
                             const double* pts_xy, int n_pts, const double pose_world[3]) {
  return grid_update_core(ctx, grid, sigma, occu_offset, use_blur, pts_xy, n_pts, pose_world, false);
}
int rsm_grid_update_by_range_map(rsm_ctx* ctx, rsm_grid* grid, double sigma, double occu_offset, int use_blur,
                                 const double* pts_xy, int n_pts, const double pose_map[3]) {
  return grid_update_core(ctx, grid, sigma, occu_offset, use_blur, pts_xy, n_pts, pose_map, true);
}
}  // extern "C"
namespace {
int grid_update_core(rsm_ctx* ctx, rsm_grid* grid, double sigma, double occu_offset, int use_blur, const double* pts_xy, int n_pts,
                     const double pose_world[3], bool pose_in_map) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || n_pts < 0 || (n_pts > 0 && !pts_xy) || !pose_world)
    return fail(ctx, RSM_ERR_INVALID, "rsm_grid_update_by_range: bad arguments");
  RasterPlan pl;
  // the stamp patterns depend on the map's cell representation, which the fill / upload decided
  int rc = plan_raster(ctx, 0.5f, sigma, grid->resolution, occu_offset, use_blur, pl);
  if (rc) return rc;
  if (!pl.blur) pl.fixed = grid->fixed;     // SetCellOccu works on either cell representation
  if (grid->fixed && !pl.fixed) return fail(ctx, RSM_ERR_UNSUPPORTED, "rsm_grid_update_by_range: the blur levels are not 2^-25 multiples but the map is held in fixed point");
  if (pl.blur && !grid->fixed && pl.fixed) {     // float map: the same stamps as float patterns
    std::vector<double> k;
    blur_kernel(sigma, grid->resolution, k);
    for (size_t i = 0; i < pl.stamp.size(); ++i) {
      const float p = static_cast<float>(k[i] * occu_offset);
      pl.stamp[i] = 0;
      if (p <= 1.0f && p > 0.0f) std::memcpy(&pl.stamp[i], &p, 4);
    }
    const float onef = 1.0f;
    std::memcpy(&pl.one, &onef, 4);
    pl.fixed = false;
  }
  Layout dl;
  const size_t o_scan = dl.take(sizeof(RasterScan));
  const size_t o_stamp = dl.take(pl.stamp.size() * 4);
  const size_t o_groups = dl.take(2 * sizeof(int), 4);
  const size_t up_bytes = dl.off;
  rc = ensure_dev(ctx, ctx->d_work, dl.off);
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_up, up_bytes);
  if (rc) return rc;
  double* d_pts = nullptr;
  if (n_pts) { rc = upload_points(ctx, pts_xy, size_t(n_pts), &d_pts); if (rc) return rc; }
  char* up = ctx->h_up.p;
  char* dw = ctx->d_work.p;
  make_raster_scan(grid, pose_world, d_pts, n_pts, *reinterpret_cast<RasterScan*>(up + o_scan), pose_in_map);
  std::memcpy(up + o_stamp, pl.stamp.data(), pl.stamp.size() * 4);
  const int groups[2] = {0, 1};
  std::memcpy(up + o_groups, groups, sizeof groups);
  CU(cudaMemcpyAsync(dw, up, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += up_bytes;
  {
    Prof p(ctx, KC_RASTER);
    CU(enqueue_raster(pl, ctx->stream, 1, reinterpret_cast<const RasterScan*>(dw + o_scan), reinterpret_cast<const int*>(dw + o_stamp),
                      1, reinterpret_cast<const int*>(dw + o_groups)));
  }
  ctx->stats.kernel_launches++;
  rc = sync_stream(ctx);
  if (rc) return rc;
  grid->init = true;      // SetUpdated() (occu_grid_map.h:325)
  return RSM_OK;
}
}  // namespace
extern "C" {

int rsm_grid_extend(rsm_ctx* ctx, rsm_grid* grid, int new_size_x, int new_size_y, int pre_grid_offset_x, int pre_grid_offset_y,
                    double new_offset_x, double new_offset_y, float fill_prob, float first_cell_prob) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !grid->owned || new_size_x <= 0 || new_size_y <= 0 || new_size_x > 32768 || new_size_y > 32768 ||
      pre_grid_offset_x < 0 || pre_grid_offset_y < 0 || pre_grid_offset_x + grid->size_x > new_size_x ||
      pre_grid_offset_y + grid->size_y > new_size_y)
    return fail(ctx, RSM_ERR_INVALID, "rsm_grid_extend: the old map must fit inside the new one");
  int vf = 0, v0 = 0;
  const bool fill_fixed = fix_ok(fill_prob, &vf) && fix_ok(first_cell_prob, &v0);
  if (grid->fixed && !fill_fixed) return fail(ctx, RSM_ERR_UNSUPPORTED, "rsm_grid_extend: fill value is not a 2^-25 multiple but the map is held in fixed point");
  if (!grid->fixed) { std::memcpy(&vf, &fill_prob, 4); std::memcpy(&v0, &first_cell_prob, 4); }
  const int new_pitch = grid_pitch_for(new_size_x);
  void* d_new = nullptr;
  cudaError_t e = cudaMalloc(&d_new, size_t(new_pitch) * new_size_y * 4);
  if (e != cudaSuccess) return fail(ctx, RSM_ERR_CUDA, "cudaMalloc(extended grid) failed: %s", cudaGetErrorString(e));
  int rc = ensure_dev(ctx, ctx->d_work, sizeof(FillJob) * 2);
  if (rc == RSM_OK) rc = ensure_pinned(ctx, ctx->h_up, sizeof(FillJob) * 2);
  if (rc) { cudaFree(d_new); return rc; }
  FillJob* F = reinterpret_cast<FillJob*>(ctx->h_up.p);
  F[0].grid = d_new; F[0].n_cells = (long long)new_pitch * new_size_y; F[0].value = vf;
  F[1].grid = d_new; F[1].n_cells = 1; F[1].value = v0;
  // grid_map_base.h:226-238: new CellType[n]{default_cell_prob_}, then the old rows are copied in at pre_grid_offset
  cudaError_t err = cudaMemcpyAsync(ctx->d_work.p, F, sizeof(FillJob) * 2, cudaMemcpyHostToDevice, ctx->stream);
  if (err == cudaSuccess) err = launch_fill(1, 148, ctx->stream, reinterpret_cast<const FillJob*>(ctx->d_work.p));
  if (err == cudaSuccess) err = launch_fill(1, 1, ctx->stream, reinterpret_cast<const FillJob*>(ctx->d_work.p) + 1);
  if (err == cudaSuccess)
    err = cudaMemcpy2DAsync(static_cast<char*>(d_new) + (size_t(pre_grid_offset_y) * new_pitch + pre_grid_offset_x) * 4, size_t(new_pitch) * 4,
                            grid->d_cells, size_t(grid->pitch) * 4, size_t(grid->size_x) * 4, size_t(grid->size_y),
                            cudaMemcpyDeviceToDevice, ctx->stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
  if (err != cudaSuccess) { cudaFree(d_new); return fail(ctx, RSM_ERR_CUDA, "rsm_grid_extend failed: %s", cudaGetErrorString(err)); }
  ctx->stats.kernel_launches += 2;
  cudaFree(grid->d_cells);
  if (grid->d_occ) { cudaFree(grid->d_occ); grid->d_occ = nullptr; }
  grid->d_cells = d_new;
  grid->size_x = new_size_x; grid->size_y = new_size_y; grid->pitch = new_pitch;
  grid->off_x = new_offset_x; grid->off_y = new_offset_y;
  grid->tf.set(grid->scale, new_offset_x, new_offset_y);      // SetMapTransform (grid_map_base.h:253)
  return RSM_OK;
}

int rsm_grid_geometry(const rsm_grid* grid, int* size_x, int* size_y, double* offset_x, double* offset_y) {
  if (!grid) return RSM_ERR_INVALID;
  if (size_x) *size_x = grid->size_x;
  if (size_y) *size_y = grid->size_y;
  if (offset_x) *offset_x = grid->off_x;
  if (offset_y) *offset_y = grid->off_y;
  return RSM_OK;
}

// ---- publishing map on the device (second half) -------------------------------------
}  // extern "C"
namespace {
int pubmap_alloc(rsm_ctx* ctx, int sx, int sy, float default_prob, float** hit, float** pass, float** prob, int** mark) {
  const size_t n = size_t(sx) * sy;
  *hit = *pass = *prob = nullptr; *mark = nullptr;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(hit), n * 4);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(pass), n * 4);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(prob), n * 4);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(mark), n * 4);
  // CountCell::ResetGridCell: prob = val, pass = hit = 0, update_index = -1 (grid_map_cell.h:64-69)
  if (e == cudaSuccess) e = cudaMemsetAsync(*hit, 0, n * 4, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(*pass, 0, n * 4, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(*mark, 0xff, n * 4, ctx->stream);
  if (e == cudaSuccess) e = launch_fill_f32(ctx->stream, *prob, (long long)n, default_prob);
  if (e != cudaSuccess) {
    cudaFree(*hit); cudaFree(*pass); cudaFree(*prob); cudaFree(*mark);
    return fail(ctx, RSM_ERR_CUDA, "publishing map allocation failed: %s", cudaGetErrorString(e));
  }
  return RSM_OK;
}
}  // namespace
extern "C" {

int rsm_pubmap_create(rsm_ctx* ctx, int size_x, int size_y, double resolution, double offset_x, double offset_y,
                      float default_prob, rsm_pubmap** out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !out || size_x <= 0 || size_y <= 0 || size_x > 32768 || size_y > 32768 || !(resolution > 0))
    return fail(ctx, RSM_ERR_INVALID, "rsm_pubmap_create: bad arguments");
  rsm_pubmap* pm = new rsm_pubmap;
  pm->default_prob = default_prob;
  rsm_grid& g = pm->g;
  g.size_x = size_x; g.size_y = size_y; g.pitch = size_x;
  g.resolution = resolution; g.scale = 1.0 / resolution;      // map/grid_map_base.h:50
  g.off_x = offset_x; g.off_y = offset_y;
  g.tf.set(g.scale, offset_x, offset_y);
  g.owned = false; g.d_cells = nullptr; g.init = false;
  int rc = pubmap_alloc(ctx, size_x, size_y, default_prob, &pm->d_hit, &pm->d_pass, &pm->d_prob, &pm->d_mark);
  if (rc == RSM_OK && cudaMalloc(reinterpret_cast<void**>(&g.d_occ), size_t(size_x) * size_y) != cudaSuccess)
    rc = fail(ctx, RSM_ERR_CUDA, "cudaMalloc(occupancy) failed");
  if (rc == RSM_OK && cudaMemsetAsync(g.d_occ, 0, size_t(size_x) * size_y, ctx->stream) != cudaSuccess) rc = RSM_ERR_CUDA;
  if (rc == RSM_OK) rc = sync_stream(ctx);
  if (rc) { delete pm; return rc; }
  *out = pm;
  return RSM_OK;
}

void rsm_pubmap_destroy(rsm_ctx* ctx, rsm_pubmap* pm) {
  DeviceGuard device_guard(ctx);
  if (!pm) return;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  cudaFree(pm->d_hit); cudaFree(pm->d_pass); cudaFree(pm->d_prob); cudaFree(pm->d_mark);
  if (pm->g.d_occ) cudaFree(pm->g.d_occ);
  delete pm;
}

const rsm_grid* rsm_pubmap_check_grid(const rsm_pubmap* pm) { return pm ? &pm->g : nullptr; }

}  // extern "C"
namespace {
// descriptor of one UpdateMapByRange of a publishing map; advances the map's update index
void make_pub_scan(rsm_pubmap* pm, const double* d_pts, int n_pts, const double* pose_world, float update_free_factor,
                   float update_occu_factor, PubScan& S) {
  rsm_grid& g = pm->g;
  std::memset(&S, 0, sizeof S);
  double pmap[3];
  g.tf.world_to_map(pose_world, pmap);                                       // occu_grid_map.h:278
  const double c = std::cos(pmap[2]), s = std::sin(pmap[2]);                 // Rotation2Dd, host libm
  S.c = c; S.s = s; S.tx = pmap[0]; S.ty = pmap[1];
  S.start_x = static_cast<int>((pmap[0] + (c * 0.0 + (-s) * 0.0)) + 0.5);    // :303-305, sensor_origin = (0, 0)
  S.start_y = static_cast<int>((pmap[1] + (s * 0.0 + c * 0.0)) + 0.5);
  S.hit = pm->d_hit; S.pass = pm->d_pass; S.prob = pm->d_prob; S.mark = pm->d_mark;
  S.pts = d_pts; S.n_pts = n_pts; S.size_x = g.size_x; S.size_y = g.size_y;
  S.free_tag = pm->cur_update_index + 1; S.occ_tag = pm->cur_update_index + 2;   // :272-273
  S.add_pass = 1.0f + update_free_factor; S.add_hit = 1.0f + update_occu_factor; // grid_map_cell.h:93-94
  S.bx0 = INT_MAX; S.by0 = INT_MAX; S.bx1 = -1; S.by1 = -1;                  // filled by the mark pass
  pm->cur_update_index += 3;                                                 // :326
  g.init = true;                                                             // SetUpdated()
}
}  // namespace
extern "C" {

int rsm_pubmap_update_by_range(rsm_ctx* ctx, rsm_pubmap* pm, const double* pts_xy, int n_pts, const double pose_world[3],
                               float update_free_factor, float update_occu_factor) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !pm || n_pts < 0 || (n_pts > 0 && !pts_xy) || !pose_world)
    return fail(ctx, RSM_ERR_INVALID, "rsm_pubmap_update_by_range: bad arguments");
  double* d_pts = nullptr;
  int rc = RSM_OK;
  if (n_pts > 0) { rc = upload_points(ctx, pts_xy, size_t(n_pts), &d_pts); if (rc) return rc; }
  PubScan S;
  make_pub_scan(pm, d_pts, n_pts, pose_world, update_free_factor, update_occu_factor, S);
  if (n_pts == 0) return RSM_OK;
  rc = ensure_dev(ctx, ctx->d_work, sizeof S);
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_up, sizeof S);
  if (rc) return rc;
  std::memcpy(ctx->h_up.p, &S, sizeof S);
  CU(cudaMemcpyAsync(ctx->d_work.p, ctx->h_up.p, sizeof S, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_pub_update(ctx->stream, reinterpret_cast<PubScan*>(ctx->d_work.p)));
  ctx->stats.kernel_launches += 2; ctx->stats.h2d_bytes += sizeof S;
  return sync_stream(ctx);
}

// ---- rebuilds from the scan store (SlamProcessor::CorrectPoseAndMap, slam/slam_processor.cpp:329-371) -------------
int rsm_pubmap_rebuild(rsm_ctx* ctx, rsm_pubmap* pm, const rsm_scan_store* store, int n, const int32_t* ids,
                       float update_free_factor, float update_occu_factor) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !pm || !store || n < 0 || (n > 0 && !ids)) return fail(ctx, RSM_ERR_INVALID, "rsm_pubmap_rebuild: bad arguments");
  for (int i = 0; i < n; ++i)
    if (ids[i] < 0 || size_t(ids[i]) >= store->scans.size()) return fail(ctx, RSM_ERR_INVALID, "rsm_pubmap_rebuild: unknown scan id %d", ids[i]);
  rsm_grid& g = pm->g;
  const size_t cells = size_t(g.size_x) * g.size_y;
  // InitMapWithRangeVec: Reset() (every cell ResetGridCell(default)), indices restart (occu_grid_map.h:222-237)
  CU(cudaMemsetAsync(pm->d_hit, 0, cells * 4, ctx->stream));
  CU(cudaMemsetAsync(pm->d_pass, 0, cells * 4, ctx->stream));
  CU(cudaMemsetAsync(pm->d_mark, 0xff, cells * 4, ctx->stream));
  CU(launch_fill_f32(ctx->stream, pm->d_prob, (long long)cells, pm->default_prob));
  pm->cur_update_index = 0;
  g.init = false;
  if (n == 0) return sync_stream(ctx);
  int rc = ensure_dev(ctx, ctx->d_work, sizeof(PubScan) * size_t(n));
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_up, sizeof(PubScan) * size_t(n));
  if (rc) return rc;
  PubScan* hs = reinterpret_cast<PubScan*>(ctx->h_up.p);
  for (int i = 0; i < n; ++i) {
    const rsm_scan_store::Entry& E = store->scans[ids[i]];
    make_pub_scan(pm, E.d_pts, E.n, E.pose, update_free_factor, update_occu_factor, hs[i]);
  }
  CU(cudaMemcpyAsync(ctx->d_work.p, hs, sizeof(PubScan) * size_t(n), cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += sizeof(PubScan) * size_t(n);
  // one update after the other on the stream: the float counters of a cell accumulate in scan order
  for (int i = 0; i < n; ++i) {
    if (hs[i].n_pts == 0) continue;
    CU(launch_pub_update(ctx->stream, reinterpret_cast<PubScan*>(ctx->d_work.p) + i));
    ctx->stats.kernel_launches += 2;
  }
  return sync_stream(ctx);
}

int rsm_grid_rebuild(rsm_ctx* ctx, rsm_grid* grid, const rsm_scan_store* store, int n, const int32_t* ids, float default_prob,
                     double sigma, double occu_offset, int use_blur) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !store || n < 0 || (n > 0 && !ids)) return fail(ctx, RSM_ERR_INVALID, "rsm_grid_rebuild: bad arguments");
  for (int i = 0; i < n; ++i)
    if (ids[i] < 0 || size_t(ids[i]) >= store->scans.size()) return fail(ctx, RSM_ERR_INVALID, "rsm_grid_rebuild: unknown scan id %d", ids[i]);
  RasterPlan pl;
  int rc = plan_raster(ctx, default_prob, sigma, grid->resolution, occu_offset, use_blur, pl);
  if (rc) return rc;
  Layout dl;
  const size_t o_fill = dl.take(sizeof(FillJob));
  const size_t o_scans = dl.take(sizeof(RasterScan) * std::max(1, n));
  const size_t o_stamp = dl.take(pl.stamp.size() * 4);
  const size_t o_groups = dl.take(2 * sizeof(int), 4);
  const size_t up_bytes = dl.off;
  rc = ensure_dev(ctx, ctx->d_work, dl.off);
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_up, up_bytes);
  if (rc) return rc;
  char* up = ctx->h_up.p;
  char* dw = ctx->d_work.p;
  FillJob F; F.grid = grid->d_cells; F.n_cells = (long long)grid->pitch * grid->size_y; F.value = pl.fill;
  std::memcpy(up + o_fill, &F, sizeof F);
  RasterScan* hs = reinterpret_cast<RasterScan*>(up + o_scans);
  for (int i = 0; i < n; ++i) {
    const rsm_scan_store::Entry& E = store->scans[ids[i]];
    make_raster_scan(grid, E.pose, E.d_pts, E.n, hs[i]);
  }
  std::memcpy(up + o_stamp, pl.stamp.data(), pl.stamp.size() * 4);
  const int groups[2] = {0, n};
  std::memcpy(up + o_groups, groups, sizeof groups);
  CU(cudaMemcpyAsync(dw, up, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += up_bytes;
  {
    Prof p(ctx, KC_RASTER);
    CU(launch_fill(1, 148, ctx->stream, reinterpret_cast<const FillJob*>(dw + o_fill)));
    if (n > 0)
      CU(enqueue_raster(pl, ctx->stream, n, reinterpret_cast<const RasterScan*>(dw + o_scans), reinterpret_cast<const int*>(dw + o_stamp),
                        1, reinterpret_cast<const int*>(dw + o_groups)));
  }
  ctx->stats.kernel_launches += (n > 0) ? 2 : 1;
  rc = sync_stream(ctx);
  if (rc) return rc;
  grid->fixed = pl.fixed;
  if (n > 0) grid->init = true;
  return RSM_OK;
}

int rsm_pubmap_extend(rsm_ctx* ctx, rsm_pubmap* pm, int new_size_x, int new_size_y, int pre_grid_offset_x, int pre_grid_offset_y,
                      double new_offset_x, double new_offset_y) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !pm) return fail(ctx, RSM_ERR_INVALID, "rsm_pubmap_extend: null argument");
  rsm_grid& g = pm->g;
  if (new_size_x <= 0 || new_size_y <= 0 || new_size_x > 32768 || new_size_y > 32768 || pre_grid_offset_x < 0 || pre_grid_offset_y < 0 ||
      pre_grid_offset_x + g.size_x > new_size_x || pre_grid_offset_y + g.size_y > new_size_y)
    return fail(ctx, RSM_ERR_INVALID, "rsm_pubmap_extend: the old map must fit inside the new one");
  float *hit, *pass, *prob; int* mark;
  int rc = pubmap_alloc(ctx, new_size_x, new_size_y, pm->default_prob, &hit, &pass, &prob, &mark);   // grid_map_base.h:226
  if (rc) return rc;
  unsigned char* occ = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&occ), size_t(new_size_x) * new_size_y) != cudaSuccess) {
    cudaFree(hit); cudaFree(pass); cudaFree(prob); cudaFree(mark);
    return fail(ctx, RSM_ERR_CUDA, "cudaMalloc(occupancy) failed");
  }
  const size_t at = size_t(pre_grid_offset_y) * new_size_x + pre_grid_offset_x;
  const size_t np = size_t(new_size_x) * 4, op = size_t(g.size_x) * 4;
  cudaError_t e = cudaMemcpy2DAsync(hit + at, np, pm->d_hit, op, op, g.size_y, cudaMemcpyDeviceToDevice, ctx->stream);   // :228-232
  if (e == cudaSuccess) e = cudaMemcpy2DAsync(pass + at, np, pm->d_pass, op, op, g.size_y, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync(prob + at, np, pm->d_prob, op, op, g.size_y, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync(mark + at, np, pm->d_mark, op, op, g.size_y, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(occ, 0, size_t(new_size_x) * new_size_y, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cudaFree(hit); cudaFree(pass); cudaFree(prob); cudaFree(mark); cudaFree(occ);
    return fail(ctx, RSM_ERR_CUDA, "rsm_pubmap_extend failed: %s", cudaGetErrorString(e));
  }
  cudaFree(pm->d_hit); cudaFree(pm->d_pass); cudaFree(pm->d_prob); cudaFree(pm->d_mark); cudaFree(g.d_occ);
  pm->d_hit = hit; pm->d_pass = pass; pm->d_prob = prob; pm->d_mark = mark; g.d_occ = occ;
  g.size_x = new_size_x; g.size_y = new_size_y; g.pitch = new_size_x;
  g.off_x = new_offset_x; g.off_y = new_offset_y;
  g.tf.set(g.scale, new_offset_x, new_offset_y);
  pm->cur_update_index += 3;                                                 // occu_grid_map.h:296-299
  return RSM_OK;
}

int rsm_pubmap_refresh_occupancy(rsm_ctx* ctx, rsm_pubmap* pm, float occu_threshold, float min_pass_through) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !pm) return fail(ctx, RSM_ERR_INVALID, "rsm_pubmap_refresh_occupancy: null argument");
  CU(launch_pub_occupancy(ctx->stream, pm->d_pass, pm->d_prob, (long long)pm->g.size_x * pm->g.size_y, occu_threshold,
                          min_pass_through, pm->g.d_occ));
  ctx->stats.kernel_launches++;
  return sync_stream(ctx);
}

int rsm_pubmap_download(rsm_ctx* ctx, const rsm_pubmap* pm, float* prob_out, float* pass_out, float* hit_out, uint8_t* occupied_out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !pm) return fail(ctx, RSM_ERR_INVALID, "rsm_pubmap_download: null argument");
  const size_t n = size_t(pm->g.size_x) * pm->g.size_y;
  if (prob_out) CU(cudaMemcpyAsync(prob_out, pm->d_prob, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (pass_out) CU(cudaMemcpyAsync(pass_out, pm->d_pass, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (hit_out) CU(cudaMemcpyAsync(hit_out, pm->d_hit, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (occupied_out) CU(cudaMemcpyAsync(occupied_out, pm->g.d_occ, n, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_stream(ctx);
}

// ---- map resize policy (host only; no device, no context) -------------------------------------------------------
int rsm_blur_half_size(double sigma, double resolution) {
  std::vector<double> k;
  return blur_kernel(sigma, resolution, k);     // -1: the reference rejects these blur parameters (occu_grid_map.h:40-59)
}

int rsm_map_bounds_create(int size_x, int size_y, double scale_factor, double offset_x, double offset_y, double extend_factor,
                          rsm_map_bounds** out) {
  if (!out || size_x <= 0 || size_y <= 0 || !(scale_factor > 0)) return RSM_ERR_INVALID;
  rsm_map_bounds* h = new rsm_map_bounds;
  h->b.init(size_x, size_y, scale_factor, offset_x, offset_y, extend_factor);
  *out = h;
  return RSM_OK;
}

void rsm_map_bounds_destroy(rsm_map_bounds* bounds) { delete bounds; }

}  // extern "C"
namespace {
void bounds_geometry(const MapBounds& b, rsm_map_geometry* g) {
  if (!g) return;
  g->size_x = b.size_x; g->size_y = b.size_y; g->pre_grid_offset_x = b.pre_x; g->pre_grid_offset_y = b.pre_y;
  g->offset_x = b.off_x; g->offset_y = b.off_y;
}
}  // namespace
extern "C" {

int rsm_map_bounds_update_scan(rsm_map_bounds* bounds, const double* pts_xy, int n_pts, const double pose_world[3], int half_kernel,
                               int use_blur, int* fits, rsm_map_geometry* geometry) {
  // n_pts == 0 would hand an empty (FLT_MAX .. FLT_MIN) box to UpdateBound and send the reference's map size to ~3e38
  if (!bounds || !pts_xy || n_pts <= 0 || !pose_world || !fits || half_kernel < 0) return RSM_ERR_INVALID;
  *fits = bounds->b.update(bounds->b.scan_box(pts_xy, n_pts, pose_world, half_kernel, use_blur != 0)) ? 1 : 0;
  bounds_geometry(bounds->b, geometry);
  return RSM_OK;
}

int rsm_map_bounds_size_check(rsm_map_bounds* bounds, const double pose_world[3], double range_max, double offset, int* fits,
                              rsm_map_geometry* geometry) {
  if (!bounds || !pose_world || !fits) return RSM_ERR_INVALID;
  *fits = bounds->b.update(bounds->b.range_box(pose_world, range_max, offset)) ? 1 : 0;
  bounds_geometry(bounds->b, geometry);
  return RSM_OK;
}

// ---- matching ----------------------------------------------------------------------------------
int rsm_match(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_pass_param* param,
              double pose_world[3], double cov[9], double* response, rsm_pass_detail* detail) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !param || !pose_world || !cov || !response || n_pts < 0 || (n_pts > 0 && !pts_xy))
    return fail(ctx, RSM_ERR_INVALID, "rsm_match: bad arguments");
  *response = 0.0;
  if (detail) std::memset(detail, 0, sizeof *detail);
  if (!grid->init || n_pts == 0) return RSM_OK;   // correlate_scan_matcher.h:792-795
  std::vector<PassItem> items(1);
  items[0].grid = grid; items[0].h_pts = pts_xy; items[0].P = n_pts; items[0].param = *param;   // the scan rides in the pass's upload
  items[0].pose_world = pose_world; items[0].cov = cov;
  int rc = run_pass(ctx, items, MODE_MATCH, nullptr, 0, nullptr);
  if (rc) return rc;
  *response = items[0].response;
  if (detail) *detail = items[0].detail;
  return RSM_OK;
}

int rsm_match_map(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_pass_param* param,
                  const double center_map[3], double cov[9], double* response, double best_map_out[3],
                  rsm_pass_detail* detail) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !param || !center_map || !cov || !response || !best_map_out || n_pts < 0 || (n_pts > 0 && !pts_xy))
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_map: bad arguments");
  *response = 0.0;
  if (detail) std::memset(detail, 0, sizeof *detail);
  if (!grid->init || n_pts == 0) return RSM_OK;   // correlate_scan_matcher.h:792-795
  double unused_pose[3] = {0.0, 0.0, 0.0};
  std::vector<PassItem> items(1);
  items[0].grid = grid; items[0].h_pts = pts_xy; items[0].P = n_pts; items[0].param = *param;
  items[0].pose_world = unused_pose; items[0].cov = cov; items[0].center_map = center_map;
  int rc = run_pass(ctx, items, MODE_MATCH, nullptr, 0, nullptr);
  if (rc) return rc;
  *response = items[0].response;
  for (int i = 0; i < 3; ++i) best_map_out[i] = items[0].detail.best_pose_map[i];
  if (detail) *detail = items[0].detail;
  return RSM_OK;
}

int rsm_scan_create(rsm_ctx* ctx, const double* pts_xy, int n_pts, rsm_scan** out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !out || n_pts < 0 || (n_pts > 0 && !pts_xy)) return fail(ctx, RSM_ERR_INVALID, "rsm_scan_create: bad arguments");
  rsm_scan* s = new rsm_scan;
  s->n = n_pts;
  if (n_pts > 0) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s->d_pts), size_t(n_pts) * 16);
    if (e != cudaSuccess) { delete s; return fail(ctx, RSM_ERR_CUDA, "cudaMalloc(scan) failed: %s", cudaGetErrorString(e)); }
    e = cudaMemcpyAsync(s->d_pts, pts_xy, size_t(n_pts) * 16, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(s->d_pts); delete s; return fail(ctx, RSM_ERR_CUDA, "scan upload failed: %s", cudaGetErrorString(e)); }
    ctx->stats.h2d_bytes += size_t(n_pts) * 16;
  }
  *out = s;
  return RSM_OK;
}

void rsm_scan_destroy(rsm_ctx* ctx, rsm_scan* scan) {
  DeviceGuard device_guard(ctx);
  if (!scan) return;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  if (scan->d_pts) cudaFree(scan->d_pts);
  delete scan;
}

int rsm_match_resident(rsm_ctx* ctx, const rsm_grid* grid, const rsm_scan* scan, const rsm_pass_param* param,
                       double pose_world[3], double cov[9], double* response, rsm_pass_detail* detail) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !scan || !param || !pose_world || !cov || !response)
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_resident: bad arguments");
  *response = 0.0;
  if (detail) std::memset(detail, 0, sizeof *detail);
  if (!grid->init || scan->n == 0) return RSM_OK;
  std::vector<PassItem> items(1);
  items[0].grid = grid; items[0].d_pts = scan->d_pts; items[0].P = scan->n; items[0].param = *param;
  items[0].pose_world = pose_world; items[0].cov = cov;
  int rc = run_pass(ctx, items, MODE_MATCH, nullptr, 0, nullptr);
  if (rc) return rc;
  *response = items[0].response;
  if (detail) *detail = items[0].detail;
  return RSM_OK;
}

int rsm_microbench_gather(rsm_ctx* ctx, int mode, int64_t footprint_bytes, int iters, double* gbps) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !gbps || mode < 0 || mode > 3 || footprint_bytes < 1024 || iters < 4)
    return fail(ctx, RSM_ERR_INVALID, "rsm_microbench_gather: bad arguments");
  if (mode < 2 && footprint_bytes > 200 * 1024) return fail(ctx, RSM_ERR_INVALID, "shared-memory tile must be <= 200 KB");
  unsigned int words = 32768;   // largest power of two not above the requested footprint (>= 128 KB)
  while (size_t(words) * 2 * 4 <= size_t(footprint_bytes)) words *= 2;
  int rc = ensure_dev(ctx, ctx->d_flush, std::max<size_t>(size_t(words) * 4 + 256, size_t(256) << 20));
  if (rc) return rc;
  unsigned long long* sink = reinterpret_cast<unsigned long long*>(ctx->d_flush.p);
  const int* g = reinterpret_cast<const int*>(ctx->d_flush.p + 64);
  const int n_cta = mode < 2 ? 148 : 148 * 2;
  iters = std::max(8, iters / 8 * 8);
  CU(launch_microbench(ctx->stream, mode, g, words, iters, n_cta, sink));   // warm-up (fills L2 / I-cache)
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
  CU(cudaEventRecord(a, ctx->stream));
  CU(launch_microbench(ctx->stream, mode, g, words, iters, n_cta, sink));
  CU(cudaEventRecord(b, ctx->stream));
  CU(cudaEventSynchronize(b));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a); cudaEventDestroy(b);
  ctx->stats.kernel_launches += 2;
  const double bytes = double(n_cta) * 1024.0 * double(iters) * 4.0;
  *gbps = bytes / (double(ms) * 1e-3) / 1e9;
  return RSM_OK;
}

int rsm_stream_plan(int n_runs, const int* runs, int variant, int max_ctas, int* out, int cap, int64_t* n_items,
                    int* n_tickets, int* n_slots) {
  if (n_runs < 1 || !runs || max_ctas < 1 || variant > 2) return 0;
  int widest = 0;
  for (int i = 0; i < n_runs; ++i) widest = std::max(widest, runs[3 * i + 1]);
  if (variant < 0) variant = score_stream_variant(widest);
  int tx, ty;
  score_stream_tile(variant, &tx, &ty);
  std::vector<StreamRun> rr(n_runs);
  for (int i = 0; i < n_runs; ++i) {
    if (runs[3 * i] < 1 || runs[3 * i + 1] < 1 || runs[3 * i + 2] < 1) return 0;
    rr[i] = {runs[3 * i], runs[3 * i + 1], runs[3 * i + 2], (runs[3 * i + 1] + tx - 1) / tx, (runs[3 * i + 1] + ty - 1) / ty};
  }
  StreamPlan P;
  plan_stream(rr, variant, max_ctas, P);
  if (n_items) *n_items = P.n_items;
  if (n_tickets) *n_tickets = P.n_tickets;
  if (n_slots) *n_slots = P.n_slots;
  if (int(P.ctas.size()) > cap || !out) return -int(P.ctas.size());
  static_assert(sizeof(StreamCta) == 12 * sizeof(int), "StreamCta layout");
  std::memcpy(out, P.ctas.data(), P.ctas.size() * sizeof(StreamCta));
  return int(P.ctas.size());
}

int rsm_match_chain(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_pass_param params[3],
                    int use_fine, double pose_world[3], double cov[9], double* score, double responses[3]) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !params || !pose_world || !cov || !score || n_pts < 0 || (n_pts > 0 && !pts_xy))
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_chain: bad arguments");
  double* d_pts = nullptr;
  if (n_pts > 0) { int rc = upload_points(ctx, pts_xy, size_t(n_pts), &d_pts); if (rc) return rc; }
  const rsm_grid* grids[1] = {grid};
  double* dp[1] = {d_pts};
  int np[1] = {n_pts};
  return run_chain(ctx, 1, grids, dp, np, params, true, use_fine != 0, pose_world, cov, score, responses);
}

int rsm_match_batch(rsm_ctx* ctx, int n, const rsm_grid* const* grids, const double* pts_xy, const int64_t* pts_offset,
                    const rsm_pass_param* params, int shared_params, int use_fine, double* poses_world, double* covs,
                    double* scores, double* responses) {
  DeviceGuard device_guard(ctx);
  if (!ctx || n < 0 || (n > 0 && (!grids || !pts_offset || !params || !poses_world || !covs || !scores)))
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_batch: bad arguments");
  if (n == 0) return RSM_OK;
  const size_t total = size_t(pts_offset[n]);
  double* d_pts = nullptr;
  if (total) { int rc = upload_points(ctx, pts_xy, total, &d_pts); if (rc) return rc; }
  std::vector<double*> dp(n);
  std::vector<int> np(n);
  for (int i = 0; i < n; ++i) { dp[i] = d_pts + 2 * pts_offset[i]; np[i] = int(pts_offset[i + 1] - pts_offset[i]); }
  return run_chain(ctx, n, grids, dp.data(), np.data(), params, shared_params != 0, use_fine != 0, poses_world, covs, scores, responses);
}

}  // extern "C"

namespace {

struct BaseRef { const double* d_pts; int n; const double* pose_world; };   // one base scan of a chain

// ScanMatchInterface for n (scan, chain) pairs (slam/slam_processor.cpp:250-326): per pair reset a
// grid_size^2 grid centred on centres_world[2i..] from the pair's base scans (:448-462), run the chain, and --
// when a publishing map is given -- scale the score by the map check and clamp it to 1 (:313-317).
// base[scan_offset[i] .. scan_offset[i+1]) = base scans of pair i, all points already on the device.
// The Gauss-Newton pre-step of the back-end chain (ScanMatchers::ScanMatch with use_optimize_scan_match_, the in-code
// default; scan_match/scan_matchers.h:205-232): everything the batched step needs about the coarse map and its scans.
struct BackendOpt {
  const std::vector<BaseRef>* base = nullptr;   // the chains' scans in coarse-map cells (same order as the fine ones)
  double* const* dp = nullptr;                  // the match scans in coarse-map cells
  const int* np = nullptr;
  int grid_size = 0;
  double resolution = 0, sigma = 0;
  const rsm_optimize_param* op = nullptr;
  double failed_cost = 0;
};

int optimize_core(rsm_ctx* ctx, int n, const rsm_grid* const* grids, const double* const* d_pts, const int* n_pts,
                  const rsm_optimize_param* op, double* poses_world, double* costs, int32_t* iterations, bool map_coords);

// Reset + stamp one pool of back-end grids: slot i is a grid_size^2 grid centred on centres_world[2i..]
// (ResetScanMatchMapWithRangeVec, slam/slam_processor.cpp:448-462).
struct GridPool {
  RasterPlan pl;
  std::vector<rsm_grid> gs;
  std::vector<const rsm_grid*> gp;
  size_t cells = 0;
  const int64_t* scan_offset = nullptr;
  const std::vector<BaseRef>* base = nullptr;
};

int make_grid_pool(rsm_ctx* ctx, Buf& pool_buf, int n, int grid_size, double resolution, float default_prob, double sigma,
                   double occu_offset, const double* centres_world, const int64_t* scan_offset, const std::vector<BaseRef>& base,
                   GridPool& P) {
  // use_blur = true as in both shipped configurations; blur parameters the reference's GaussianBlur rejects select the
  // SET_CELL_OCCUPIED update, as there (map/occu_grid_map.h:265-268)
  int rc = plan_raster(ctx, default_prob, sigma, resolution, occu_offset, 1, P.pl);
  if (rc) return rc;
  const int pool_pitch = grid_pitch_for(grid_size);
  P.cells = size_t(pool_pitch) * grid_size;
  const size_t slot = (P.cells * 4 + 255) / 256 * 256;
  rc = ensure_dev(ctx, pool_buf, slot * n);
  if (rc) return rc;
  P.gs.assign(n, rsm_grid());
  P.gp.resize(n);
  P.scan_offset = scan_offset; P.base = &base;
  const double scale = 1.0 / resolution;
  const double cell_len = 1 / scale;
  for (int i = 0; i < n; ++i) {
    rsm_grid& g = P.gs[i];
    g.size_x = g.size_y = grid_size; g.pitch = pool_pitch;
    g.resolution = resolution; g.scale = scale;
    // ResetScanMatchMapWithRangeVec: offset = -(pose - 0.5 * size * cell_len)   (slam_processor.cpp:451-455)
    g.off_x = -(centres_world[2 * i] - 0.5 * grid_size * cell_len);
    g.off_y = -(centres_world[2 * i + 1] - 0.5 * grid_size * cell_len);
    g.tf.set(scale, g.off_x, g.off_y);
    g.fixed = P.pl.fixed; g.owned = false;
    g.d_cells = pool_buf.p + slot * i;
    g.init = scan_offset[i + 1] > scan_offset[i];
    P.gp[i] = &g;
  }
  return RSM_OK;
}

// Reset + stamp the grids of pairs [first, first + count) of a pool on a lane's stream (the descriptors live in the lane's
// own aux buffers at aux_off: a pass preparation reuses the work arena right away)
int raster_pool(rsm_ctx* ctx, const GridPool& P, Lane* L, int first, int count) {
  const int64_t* scan_offset = P.scan_offset;
  const std::vector<BaseRef>& base = *P.base;
  const int64_t s0 = scan_offset[first], s1 = scan_offset[first + count];
  Layout dl;
  const size_t o_fill = dl.take(sizeof(FillJob) * size_t(count));
  const size_t o_scans = dl.take(sizeof(RasterScan) * size_t(std::max<int64_t>(1, s1 - s0)));
  const size_t o_stamp = dl.take(P.pl.stamp.size() * 4);
  const size_t o_groups = dl.take(sizeof(int) * size_t(count + 1), 4);
  const size_t up_bytes = dl.off;
  int rc2 = ensure_dev(ctx, L->d_aux, dl.off, L->stream);
  if (rc2) return rc2;
  rc2 = ensure_pinned(ctx, L->h_aux, up_bytes, L->stream);
  if (rc2) return rc2;
  char* up = L->h_aux.p;
  char* dw = L->d_aux.p;
  FillJob* hf = reinterpret_cast<FillJob*>(up + o_fill);
  RasterScan* hs = reinterpret_cast<RasterScan*>(up + o_scans);
  int* hg = reinterpret_cast<int*>(up + o_groups);
  for (int k = 0; k < count; ++k) {
    const int i = first + k;
    hf[k].grid = P.gs[i].d_cells; hf[k].n_cells = (long long)P.cells; hf[k].value = P.pl.fill;
    hg[k] = int(scan_offset[i] - s0);
  }
  hg[count] = int(s1 - s0);
  ctx->pool.run(count, 32, [&](int k0, int k1) {
    for (int k = k0; k < k1; ++k) {
      const int i = first + k;
      for (int64_t sidx = scan_offset[i]; sidx < scan_offset[i + 1]; ++sidx)
        make_raster_scan(&P.gs[i], base[sidx].pose_world, base[sidx].d_pts, base[sidx].n, hs[sidx - s0]);
    }
  });
  std::memcpy(up + o_stamp, P.pl.stamp.data(), P.pl.stamp.size() * 4);
  CU(cudaMemcpyAsync(dw, up, up_bytes, cudaMemcpyHostToDevice, L->stream));
  L->stats.h2d_bytes += up_bytes;
  {
    Prof p(ctx, KC_RASTER, L);
    CU(launch_fill(count, 4, L->stream, reinterpret_cast<const FillJob*>(dw + o_fill)));
    if (s1 > s0)
      CU(enqueue_raster(P.pl, L->stream, int(s1 - s0), reinterpret_cast<const RasterScan*>(dw + o_scans),
                        reinterpret_cast<const int*>(dw + o_stamp), count, reinterpret_cast<const int*>(dw + o_groups)));
  }
  L->stats.kernel_launches += 2;
  return RSM_OK;
}

int loop_closure_core(rsm_ctx* ctx, int n, int grid_size, double resolution, float default_prob, double sigma,
                      double occu_offset, const double* centres_world, const int64_t* scan_offset,
                      const std::vector<BaseRef>& base, double* const* dp, const int* np, const rsm_pass_param* params,
                      bool use_fine, double* poses_world, double* covs, double* scores, double* responses,
                      const rsm_grid* pub_map, const double* const* pub_pts, const int32_t* pub_counts,
                      const rsm_map_check_param* check, const BackendOpt* opt = nullptr) {
  PhaseTimer pt_setup(&ctx->stats);
  GridPool fine;
  int rc = make_grid_pool(ctx, ctx->d_pool_grids, n, grid_size, resolution, default_prob, sigma, occu_offset, centres_world,
                          scan_offset, base, fine);
  if (rc) return rc;
  const std::function<int(Lane*, int, int)> pre = [&](Lane* L, int first, int count) -> int {
    return raster_pool(ctx, fine, L, first, count);
  };
  // ---- Gauss-Newton pre-step on the coarse maps (scan_matchers.h:205-232) -----------------------------------------
  std::vector<unsigned char> skip_coarse;
  std::vector<double> opt_cost;
  if (opt) {
    GridPool coarse;
    rc = make_grid_pool(ctx, ctx->d_pool_grids2, n, opt->grid_size, opt->resolution, default_prob, opt->sigma, occu_offset,
                        centres_world, scan_offset, *opt->base, coarse);
    if (rc) return rc;
    rc = raster_pool(ctx, coarse, &ctx->L0, 0, n);
    if (rc) return rc;
    rc = sync_stream(ctx);       // optimize_core reuses the staging buffers
    if (rc) return rc;
    std::vector<double> process(poses_world, poses_world + 3 * size_t(n));
    opt_cost.assign(n, 0.0);
    std::vector<const double*> cdp(n);
    for (int i = 0; i < n; ++i) cdp[i] = opt->dp[i];
    rc = optimize_core(ctx, n, coarse.gp.data(), cdp.data(), opt->np, opt->op, process.data(), opt_cost.data(), nullptr, false);   // :207
    if (rc) return rc;
    skip_coarse.assign(n, 0);
    for (int i = 0; i < n; ++i) {
      // the coarse correlative pass runs unless the fine passes follow and the optimiser succeeded (:224-226); it then
      // starts from the seed again (:232)
      if (use_fine && !(opt_cost[i] > opt->failed_cost)) {
        skip_coarse[i] = 1;
        for (int k = 0; k < 3; ++k) poses_world[3 * i + k] = process[3 * i + k];
      }
    }
  }
  pt_setup.lap(6);
  std::vector<double> resp3;
  double* r3 = responses;
  if (opt) { resp3.assign(3 * size_t(n), 0.0); r3 = resp3.data(); }
  rc = run_chain(ctx, n, fine.gp.data(), dp, np, params, true, use_fine, poses_world, covs, scores, r3, &pre,
                 opt ? skip_coarse.data() : nullptr);
  pt_setup.lap(7);
  if (rc) return rc;
  if (opt) {
    for (int i = 0; i < n; ++i) {
      // scan_match_score / scan_match_times (:211, :230-240, :247-263, :281)
      double score = skip_coarse[i] ? opt->failed_cost / (opt_cost[i] + opt->failed_cost) : r3[3 * i];
      int times = 1;
      if (use_fine) { score += r3[3 * i + 1]; times++; score += r3[3 * i + 2]; times++; }
      scores[i] = score / times;
      if (responses) { responses[4 * i] = opt_cost[i]; for (int k = 0; k < 3; ++k) responses[4 * i + 1 + k] = r3[3 * i + k]; }
    }
  }
  if (!pub_map) return RSM_OK;
  // MapCheckPenalize(pub_map_range_data, best_pose, true) on the matched poses (slam_processor.cpp:313-317)
  std::vector<double> coeff(n);
  rc = map_check_core(ctx, pub_map, n, poses_world, nullptr, 0, 0, nullptr, pub_pts, pub_counts, nullptr,
                      check->check_point_num, check->bound_tolerance, check->penalty_gain, check->use_logistic, coeff.data());
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    scores[i] *= coeff[i];
    scores[i] = (scores[i] > 1.0) ? (1.0) : (scores[i]);
  }
  return RSM_OK;
}


// BasedOptimizeScanMatch::ScanMatch (scan_match/optimize_scan_matcher.h:68-131) for n independent problems.
// Every iteration is one launch over the still-active problems (UpdateCost on the device, sums in point
// order); the 3x3 solve, the stopping rule and the clamped update run here with the host libm, as do the
// cos / sin of the next estimate.  poses_world in/out; costs out; iterations (nullable) = UpdateCost calls.
int optimize_core(rsm_ctx* ctx, int n, const rsm_grid* const* grids, const double* const* d_pts, const int* n_pts,
                  const rsm_optimize_param* op, double* poses_world, double* costs, int32_t* iterations,
                  bool map_coords) {   // map_coords: poses are given and returned in map cells (adapter path)
  const double kMaxCost = 1.0 * 1000;                         // :231-232
  if (op->iterate_max_times < 1)   // the reference would return the previous call's cost_ (a stale member)
    return fail(ctx, RSM_ERR_INVALID, "rsm_optimize: iterate_max_times must be >= 1");
  struct Prob { int i; double est[3]; double cost, last_cost; };
  std::vector<Prob> act;
  act.reserve(n);
  for (int i = 0; i < n; ++i) {
    if (iterations) iterations[i] = 0;
    if (!grids[i]->init || n_pts[i] == 0) { costs[i] = kMaxCost; continue; }    // :73-76, pose untouched
    Prob P; P.i = i; P.cost = 0.0; P.last_cost = 0.0;
    if (map_coords) { for (int k = 0; k < 3; ++k) P.est[k] = poses_world[3 * i + k]; }
    else grids[i]->tf.world_to_map(poses_world + 3 * i, P.est);                 // :80-81
    act.push_back(P);
  }
  const size_t out_stride = kOptSums + 1;
  auto finish = [&](const Prob& P) {
    double est[3] = {P.est[0], P.est[1], normalize_angle(P.est[2])};            // :125
    if (map_coords) { for (int k = 0; k < 3; ++k) poses_world[3 * P.i + k] = est[k]; }
    else grids[P.i]->tf.map_to_world(est, poses_world + 3 * P.i);                // :127
    costs[P.i] = P.cost;
  };
  for (int iter = 0; iter < op->iterate_max_times && !act.empty(); ++iter) {
    const int m = int(act.size());
    Layout dl;
    const size_t o_jobs = dl.take(sizeof(OptimizeJob) * size_t(m));
    const size_t up_bytes = dl.off;
    const size_t o_out = dl.take(size_t(m) * out_stride * 8, 16);
    int rc = ensure_dev(ctx, ctx->d_work, dl.off);
    if (rc) return rc;
    rc = ensure_pinned(ctx, ctx->h_up, up_bytes);
    if (rc) return rc;
    rc = ensure_pinned(ctx, ctx->h_down, size_t(m) * out_stride * 8);
    if (rc) return rc;
    char* dw = ctx->d_work.p;
    OptimizeJob* jobs = reinterpret_cast<OptimizeJob*>(ctx->h_up.p + o_jobs);
    for (int a = 0; a < m; ++a) {
      const Prob& P = act[a];
      const rsm_grid* g = grids[P.i];
      OptimizeJob& J = jobs[a];
      std::memset(&J, 0, sizeof J);
      J.grid = g->d_cells; J.pts = d_pts[P.i]; J.n_pts = n_pts[P.i];
      J.size_x = g->size_x; J.size_y = g->size_y; J.pitch = g->pitch; J.fixed = g->fixed ? 1 : 0;
      J.c = std::cos(P.est[2]); J.s = std::sin(P.est[2]);                        // :95-97
      J.tx = P.est[0]; J.ty = P.est[1];
      J.out = reinterpret_cast<double*>(dw + o_out) + size_t(a) * out_stride;
    }
    CU(cudaMemcpyAsync(dw + o_jobs, ctx->h_up.p + o_jobs, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_optimize(m, ctx->stream, reinterpret_cast<const OptimizeJob*>(dw + o_jobs)));
    CU(cudaMemcpyAsync(ctx->h_down.p, dw + o_out, size_t(m) * out_stride * 8, cudaMemcpyDeviceToHost, ctx->stream));
    rc = sync_stream(ctx);
    if (rc) return rc;
    ctx->stats.h2d_bytes += up_bytes; ctx->stats.d2h_bytes += size_t(m) * out_stride * 8; ctx->stats.kernel_launches++;
    const double* out = reinterpret_cast<const double*>(ctx->h_down.p);
    std::vector<Prob> next;
    next.reserve(m);
    for (int a = 0; a < m; ++a) {
      Prob P = act[a];
      const double* S = out + size_t(a) * out_stride;
      if (iterations) iterations[P.i] = iter + 1;
      const int valid_point = 1 + int(S[kOptSums]);                              // :160, :211
      P.last_cost = P.cost;                                                      // :87
      P.cost = S[9] * (1000.0 / valid_point);                                    // :218
      const double H[3][3] = {{S[0], S[1], S[2]}, {S[1], S[3], S[4]}, {S[2], S[4], S[5]}};
      const double b[3] = {S[6], S[7], S[8]};
      double det[3];
      ldlt3_solve(H, b, det);                                                    // :135-141
      if (std::isnan(det[0]) || std::isnan(det[1]) || std::isnan(det[2])) { costs[P.i] = kMaxCost; continue; }   // :103-106
      if (iter > 0 && (P.last_cost - P.cost < op->cost_decrease_threshold || P.cost < op->cost_min_threshold)) {  // :112-118
        finish(P);
        continue;
      }
      const double map_resolution = grids[P.i]->cell_len();                      // :83
      P.est[0] += max_abs_limit(det[0], op->max_update_distance / map_resolution);   // :143-152
      P.est[1] += max_abs_limit(det[1], op->max_update_distance / map_resolution);
      P.est[2] += max_abs_limit(det[2], op->max_update_angle);
      next.push_back(P);
    }
    act.swap(next);
  }
  for (const Prob& P : act) finish(P);     // ran out of iterations
  return RSM_OK;
}
}  // namespace

extern "C" {

int rsm_loop_closure_batch(rsm_ctx* ctx, int n, int grid_size, double resolution, float default_prob, double sigma,
                           double occu_offset, const double* centres_world, const int64_t* scan_offset,
                           const int32_t* base_n_pts, const double* base_pts_xy, const double* base_poses_world,
                           const double* pts_xy, const int64_t* pts_offset, const rsm_pass_param params[3], int use_fine,
                           double* poses_world, double* covs, double* scores, double* responses) {
  DeviceGuard device_guard(ctx);
  if (!ctx || n < 0 || grid_size <= 0 || !(resolution > 0) ||
      (n > 0 && (!centres_world || !scan_offset || !base_n_pts || !base_pts_xy || !base_poses_world || !pts_xy ||
                 !pts_offset || !params || !poses_world || !covs || !scores)))
    return fail(ctx, RSM_ERR_INVALID, "rsm_loop_closure_batch: bad arguments");
  if (n == 0) return RSM_OK;
  // points: [base scans | match scans] in one device buffer
  const int64_t n_base_scans = scan_offset[n];
  size_t base_pts_total = 0;
  for (int64_t s = 0; s < n_base_scans; ++s) base_pts_total += size_t(base_n_pts[s]);
  const size_t match_pts_total = size_t(pts_offset[n]);
  int rc = ensure_dev(ctx, ctx->d_pts, (base_pts_total + match_pts_total) * 16);
  if (rc) return rc;
  double* d_base = reinterpret_cast<double*>(ctx->d_pts.p);
  double* d_match = d_base + 2 * base_pts_total;
  if (base_pts_total) CU(cudaMemcpyAsync(d_base, base_pts_xy, base_pts_total * 16, cudaMemcpyHostToDevice, ctx->stream));
  if (match_pts_total) CU(cudaMemcpyAsync(d_match, pts_xy, match_pts_total * 16, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += (base_pts_total + match_pts_total) * 16;
  std::vector<BaseRef> base(n_base_scans);
  size_t poff = 0;
  for (int64_t s = 0; s < n_base_scans; ++s) {
    base[s] = BaseRef{d_base + 2 * poff, base_n_pts[s], base_poses_world + 3 * s};
    poff += base_n_pts[s];
  }
  std::vector<double*> dp(n);
  std::vector<int> np(n);
  for (int i = 0; i < n; ++i) { dp[i] = d_match + 2 * pts_offset[i]; np[i] = int(pts_offset[i + 1] - pts_offset[i]); }
  return loop_closure_core(ctx, n, grid_size, resolution, default_prob, sigma, occu_offset, centres_world, scan_offset,
                           base, dp.data(), np.data(), params, use_fine != 0, poses_world, covs, scores, responses,
                           nullptr, nullptr, nullptr, nullptr);
}

// ---- scan store ----------------------------------------------------------------------------------
int rsm_scan_store_create(rsm_ctx* ctx, rsm_scan_store** out) {
  if (!ctx || !out) return fail(ctx, RSM_ERR_INVALID, "rsm_scan_store_create: null argument");
  *out = new rsm_scan_store;
  return RSM_OK;
}

void rsm_scan_store_destroy(rsm_ctx* ctx, rsm_scan_store* store) {
  DeviceGuard device_guard(ctx);
  if (!store) return;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  for (void* c : store->chunks) cudaFree(c);
  delete store;
}

int rsm_scan_store_size(const rsm_scan_store* store) { return store ? int(store->scans.size()) : 0; }

int rsm_scan_store_add(rsm_ctx* ctx, rsm_scan_store* store, const double* pts_xy, int n_pts, const double pose_world[3],
                       int32_t* id_out) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !store || n_pts < 0 || (n_pts > 0 && !pts_xy) || !pose_world)
    return fail(ctx, RSM_ERR_INVALID, "rsm_scan_store_add: bad arguments");
  const size_t bytes = (size_t(n_pts) * 16 + 255) / 256 * 256;
  if (store->chunks.empty() || store->chunk_used + bytes > store->chunk_cap) {
    const size_t cap = std::max<size_t>(bytes, size_t(16) << 20);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess) return fail(ctx, RSM_ERR_CUDA, "cudaMalloc(scan store) failed: %s", cudaGetErrorString(e));
    store->chunks.push_back(p); store->chunk_used = 0; store->chunk_cap = cap;
  }
  rsm_scan_store::Entry E;
  E.d_pts = reinterpret_cast<double*>(static_cast<char*>(store->chunks.back()) + store->chunk_used);
  E.n = n_pts;
  for (int k = 0; k < 3; ++k) E.pose[k] = pose_world[k];
  if (n_pts > 0) {
    CU(cudaMemcpyAsync(E.d_pts, pts_xy, size_t(n_pts) * 16, cudaMemcpyHostToDevice, ctx->stream));
    int rc = sync_stream(ctx);   // the caller's array is only borrowed for the call
    if (rc) return rc;
    ctx->stats.h2d_bytes += size_t(n_pts) * 16;
  }
  store->chunk_used += bytes;
  store->scans.push_back(E);
  if (id_out) *id_out = int32_t(store->scans.size() - 1);
  return RSM_OK;
}

int rsm_scan_store_set_poses(rsm_ctx* ctx, rsm_scan_store* store, int n, const int32_t* ids, const double* poses_world) {
  if (!ctx || !store || n < 0 || (n > 0 && (!ids || !poses_world)))
    return fail(ctx, RSM_ERR_INVALID, "rsm_scan_store_set_poses: bad arguments");
  for (int i = 0; i < n; ++i)
    if (ids[i] < 0 || size_t(ids[i]) >= store->scans.size()) return fail(ctx, RSM_ERR_INVALID, "rsm_scan_store_set_poses: unknown scan id %d", ids[i]);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) store->scans[ids[i]].pose[k] = poses_world[3 * i + k];
  return RSM_OK;
}

int rsm_scan_store_get_pose(const rsm_scan_store* store, int32_t id, double pose_world[3]) {
  if (!store || !pose_world || id < 0 || size_t(id) >= store->scans.size()) return RSM_ERR_INVALID;
  for (int k = 0; k < 3; ++k) pose_world[k] = store->scans[id].pose[k];
  return RSM_OK;
}

}  // extern "C"
namespace {
// the argument checks and id lookups shared by the two batched back-end entry points
int interface_batch_core(rsm_ctx* ctx, const char* who, const rsm_scan_store* store, int n, int grid_size, double resolution,
                         float default_prob, double sigma, double occu_offset, const double* centres_world,
                         const int64_t* chain_offset, const int32_t* chain_ids, const int32_t* match_ids,
                         const rsm_pass_param params[3], int use_fine, double* poses_world, double* covs, double* scores,
                         double* responses, const rsm_grid* pub_map, const rsm_scan_store* pub_store,
                         const rsm_map_check_param* check, const rsm_scan_store* coarse_store, int coarse_grid_size,
                         double coarse_resolution, double coarse_sigma, const rsm_optimize_param* opt, double optimize_failed_cost) {
  if (!ctx || !store || n < 0 || grid_size <= 0 || !(resolution > 0) ||
      (n > 0 && (!centres_world || !chain_offset || !match_ids || !params || !poses_world || !covs || !scores)) ||
      (pub_map && (!pub_store || !check)) ||
      (coarse_store && (!opt || coarse_grid_size <= 0 || !(coarse_resolution > 0))))
    return fail(ctx, RSM_ERR_INVALID, "%s: bad arguments", who);
  if (n == 0) return RSM_OK;
  if (pub_map && !pub_map->d_occ) return fail(ctx, RSM_ERR_NOT_INIT, "%s: no occupancy uploaded for the publishing map", who);
  const int64_t n_entries = chain_offset[n];
  if (n_entries < 0 || (n_entries > 0 && !chain_ids)) return fail(ctx, RSM_ERR_INVALID, "%s: bad chain list", who);
  const size_t have = store->scans.size();
  std::vector<BaseRef> base(n_entries), cbase(coarse_store ? n_entries : 0);
  for (int64_t s = 0; s < n_entries; ++s) {
    if (chain_ids[s] < 0 || size_t(chain_ids[s]) >= have) return fail(ctx, RSM_ERR_INVALID, "%s: unknown scan id %d", who, chain_ids[s]);
    const rsm_scan_store::Entry& E = store->scans[chain_ids[s]];
    base[s] = BaseRef{E.d_pts, E.n, E.pose};
    if (coarse_store) {
      if (size_t(chain_ids[s]) >= coarse_store->scans.size()) return fail(ctx, RSM_ERR_INVALID, "%s: scan id %d is not in the coarse-map store", who, chain_ids[s]);
      const rsm_scan_store::Entry& C = coarse_store->scans[chain_ids[s]];
      cbase[s] = BaseRef{C.d_pts, C.n, E.pose};      // one pose per scan: the fine store's (UpdateRangeData updates every resolution alike)
    }
  }
  std::vector<double*> dp(n), cdp(coarse_store ? n : 0);
  std::vector<int> np(n), cnp(coarse_store ? n : 0);
  std::vector<const double*> pub_pts(pub_map ? n : 0);
  std::vector<int32_t> pub_counts(pub_map ? n : 0);
  for (int i = 0; i < n; ++i) {
    if (match_ids[i] < 0 || size_t(match_ids[i]) >= have) return fail(ctx, RSM_ERR_INVALID, "%s: unknown scan id %d", who, match_ids[i]);
    dp[i] = store->scans[match_ids[i]].d_pts; np[i] = store->scans[match_ids[i]].n;
    if (pub_map) {
      if (size_t(match_ids[i]) >= pub_store->scans.size()) return fail(ctx, RSM_ERR_INVALID, "%s: scan id %d is not in the publishing-map store", who, match_ids[i]);
      pub_pts[i] = pub_store->scans[match_ids[i]].d_pts; pub_counts[i] = pub_store->scans[match_ids[i]].n;
    }
    if (coarse_store) {
      if (size_t(match_ids[i]) >= coarse_store->scans.size()) return fail(ctx, RSM_ERR_INVALID, "%s: scan id %d is not in the coarse-map store", who, match_ids[i]);
      cdp[i] = coarse_store->scans[match_ids[i]].d_pts; cnp[i] = coarse_store->scans[match_ids[i]].n;
    }
  }
  BackendOpt bo;
  if (coarse_store) {
    bo.base = &cbase; bo.dp = cdp.data(); bo.np = cnp.data(); bo.grid_size = coarse_grid_size; bo.resolution = coarse_resolution;
    bo.sigma = coarse_sigma; bo.op = opt; bo.failed_cost = optimize_failed_cost;
  }
  return loop_closure_core(ctx, n, grid_size, resolution, default_prob, sigma, occu_offset, centres_world, chain_offset,
                           base, dp.data(), np.data(), params, use_fine != 0, poses_world, covs, scores, responses,
                           pub_map, pub_map ? pub_pts.data() : nullptr, pub_map ? pub_counts.data() : nullptr, check,
                           coarse_store ? &bo : nullptr);
}
}  // namespace
extern "C" {

int rsm_scan_match_interface_batch(rsm_ctx* ctx, const rsm_scan_store* store, int n, int grid_size, double resolution,
                                   float default_prob, double sigma, double occu_offset, const double* centres_world,
                                   const int64_t* chain_offset, const int32_t* chain_ids, const int32_t* match_ids,
                                   const rsm_pass_param params[3], int use_fine, double* poses_world, double* covs,
                                   double* scores, double* responses, const rsm_grid* pub_map,
                                   const rsm_scan_store* pub_store, const rsm_map_check_param* check) {
  DeviceGuard device_guard(ctx);
  return interface_batch_core(ctx, "rsm_scan_match_interface_batch", store, n, grid_size, resolution, default_prob, sigma, occu_offset,
                              centres_world, chain_offset, chain_ids, match_ids, params, use_fine, poses_world, covs, scores,
                              responses, pub_map, pub_store, check, nullptr, 0, 0.0, 0.0, nullptr, 0.0);
}

int rsm_scan_match_interface_batch_opt(rsm_ctx* ctx, const rsm_scan_store* fine_store, const rsm_scan_store* coarse_store, int n,
                                       int grid_size, double resolution, double sigma, int coarse_grid_size,
                                       double coarse_resolution, double coarse_sigma, float default_prob, double occu_offset,
                                       const double* centres_world, const int64_t* chain_offset, const int32_t* chain_ids,
                                       const int32_t* match_ids, const rsm_pass_param params[3],
                                       const rsm_optimize_param* optimize, double optimize_failed_cost, int use_fine,
                                       double* poses_world, double* covs, double* scores, double* responses,
                                       const rsm_grid* pub_map, const rsm_scan_store* pub_store,
                                       const rsm_map_check_param* check) {
  DeviceGuard device_guard(ctx);
  if (!coarse_store || !optimize) return fail(ctx, RSM_ERR_INVALID, "rsm_scan_match_interface_batch_opt: bad arguments");
  return interface_batch_core(ctx, "rsm_scan_match_interface_batch_opt", fine_store, n, grid_size, resolution, default_prob, sigma,
                              occu_offset, centres_world, chain_offset, chain_ids, match_ids, params, use_fine, poses_world, covs,
                              scores, responses, pub_map, pub_store, check, coarse_store, coarse_grid_size, coarse_resolution,
                              coarse_sigma, optimize, optimize_failed_cost);
}

// ---- Gauss-Newton matcher -----------------------------------------------------------------------
int rsm_optimize_batch(rsm_ctx* ctx, int n, const rsm_grid* const* grids, const double* pts_xy, const int64_t* pts_offset,
                       const rsm_optimize_param* param, double* poses_world, double* costs, int32_t* iterations) {
  DeviceGuard device_guard(ctx);
  if (!ctx || n < 0 || !param || (n > 0 && (!grids || !pts_offset || !poses_world || !costs)))
    return fail(ctx, RSM_ERR_INVALID, "rsm_optimize_batch: bad arguments");
  if (n == 0) return RSM_OK;
  for (int i = 0; i < n; ++i) if (!grids[i]) return fail(ctx, RSM_ERR_INVALID, "rsm_optimize_batch: null grid");
  const size_t total = size_t(pts_offset[n]);
  if (total && !pts_xy) return fail(ctx, RSM_ERR_INVALID, "rsm_optimize_batch: pts_xy is null");
  double* d_pts = nullptr;
  if (total) { int rc = upload_points(ctx, pts_xy, total, &d_pts); if (rc) return rc; }
  std::vector<const double*> dp(n);
  std::vector<int> np(n);
  for (int i = 0; i < n; ++i) { dp[i] = d_pts + 2 * pts_offset[i]; np[i] = int(pts_offset[i + 1] - pts_offset[i]); }
  return optimize_core(ctx, n, grids, dp.data(), np.data(), param, poses_world, costs, iterations, false);
}

int rsm_optimize(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_optimize_param* param,
                 double pose_world[3], double* cost, int32_t* iterations) {
  if (!grid || n_pts < 0) return fail(ctx, RSM_ERR_INVALID, "rsm_optimize: bad arguments");
  const int64_t off[2] = {0, n_pts};
  const rsm_grid* gs[1] = {grid};
  return rsm_optimize_batch(ctx, 1, gs, pts_xy, off, param, pose_world, cost, iterations);
}

int rsm_optimize_map(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_optimize_param* param,
                     double pose_map[3], double* cost, int32_t* iterations) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !param || !pose_map || !cost || n_pts < 0 || (n_pts > 0 && !pts_xy))
    return fail(ctx, RSM_ERR_INVALID, "rsm_optimize_map: bad arguments");
  double* d_pts = nullptr;
  if (n_pts > 0) { int rc = upload_points(ctx, pts_xy, size_t(n_pts), &d_pts); if (rc) return rc; }
  const rsm_grid* gs[1] = {grid};
  const double* dp[1] = {d_pts};
  const int np[1] = {n_pts};
  return optimize_core(ctx, 1, gs, dp, np, param, pose_map, cost, iterations, true);
}

int rsm_match_chain_opt(rsm_ctx* ctx, const rsm_grid* coarse_grid, const double* pts_coarse, int n_coarse,
                        const rsm_grid* fine_grid, const double* pts_fine, int n_fine, const rsm_pass_param params[3],
                        const rsm_optimize_param* opt, double optimize_failed_cost, int use_fine, double pose_world[3],
                        double cov[9], double* score, double responses[4]) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !coarse_grid || !fine_grid || !params || !opt || !pose_world || !cov || !score || n_coarse < 0 || n_fine < 0 ||
      (n_coarse > 0 && !pts_coarse) || (n_fine > 0 && !pts_fine))
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_chain_opt: bad arguments");
  // scan_matchers.h:179-289 with use_optimize_scan_match_ = true
  const double best_pose[3] = {pose_world[0], pose_world[1], pose_world[2]};
  double process_pose[3] = {best_pose[0], best_pose[1], best_pose[2]};
  double optimize_cost = 0.0;
  int rc = rsm_optimize(ctx, coarse_grid, pts_coarse, n_coarse, opt, process_pose, &optimize_cost, nullptr);   // :207
  if (rc) return rc;
  double scan_match_score = optimize_failed_cost / (optimize_cost + optimize_failed_cost);   // :211
  int scan_match_times = 1;
  double r[3] = {0.0, 0.0, 0.0};
  double* d_pts = nullptr;
  if (n_fine > 0) { rc = upload_points(ctx, pts_fine, size_t(n_fine), &d_pts); if (rc) return rc; }
  auto one_pass = [&](int k) -> int {
    std::vector<PassItem> items(1);
    items[0].grid = fine_grid; items[0].d_pts = d_pts; items[0].P = n_fine; items[0].param = params[k];
    items[0].pose_world = process_pose; items[0].cov = cov;
    if (!fine_grid->init || n_fine == 0) { r[k] = 0.0; return RSM_OK; }          // correlate_scan_matcher.h:792-795
    int e = run_pass(ctx, items, MODE_MATCH, nullptr, 0, nullptr);
    if (e) return e;
    r[k] = items[0].response;
    return RSM_OK;
  };
  if (!use_fine || optimize_cost > optimize_failed_cost) {                                     // :224-226
    scan_match_score = 0.0;                                                                    // :230-232
    scan_match_times--;
    for (int k = 0; k < 3; ++k) process_pose[k] = best_pose[k];
    rc = one_pass(0);                                                                          // :237-239
    if (rc) return rc;
    scan_match_score += r[0];
    scan_match_times++;
  }
  if (use_fine) {                                                                              // :247-263
    rc = one_pass(1);
    if (rc) return rc;
    scan_match_score += r[1]; scan_match_times++;
    rc = one_pass(2);
    if (rc) return rc;
    scan_match_score += r[2]; scan_match_times++;
  }
  scan_match_score /= scan_match_times;                                                        // :281
  for (int k = 0; k < 3; ++k) pose_world[k] = process_pose[k];
  *score = scan_match_score;
  if (responses) { responses[0] = optimize_cost; responses[1] = r[0]; responses[2] = r[1]; responses[3] = r[2]; }
  return RSM_OK;
}

int rsm_pass_scores(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_pass_param* param,
                    const double pose_world[3], int angle_begin, int angle_end, double* scores_out, int64_t capacity,
                    int64_t* n_written) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !param || !pose_world || !scores_out || n_pts <= 0 || !pts_xy)
    return fail(ctx, RSM_ERR_INVALID, "rsm_pass_scores: bad arguments");
  if (!grid->init) return fail(ctx, RSM_ERR_NOT_INIT, "grid has no content");
  double* d_pts = nullptr;
  int rc = upload_points(ctx, pts_xy, size_t(n_pts), &d_pts);
  if (rc) return rc;
  double pose[3] = {pose_world[0], pose_world[1], pose_world[2]};
  double cov[9];
  std::vector<PassItem> items(1);
  items[0].grid = grid; items[0].d_pts = d_pts; items[0].P = n_pts; items[0].param = *param;
  items[0].pose_world = pose; items[0].cov = cov;
  items[0].ang_begin = angle_begin; items[0].ang_end = angle_end;
  return run_pass(ctx, items, MODE_SCORES, scores_out, capacity, n_written);
}

int rsm_match_partial(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_pass_param* param,
                      const double pose_world[3], int angle_begin, int angle_end, void* partial) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !param || !pose_world || !partial || n_pts <= 0 || !pts_xy)
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_partial: bad arguments");
  if (!grid->init) return fail(ctx, RSM_ERR_NOT_INIT, "grid has no content");
  ctx->slice.valid = false;
  std::memset(partial, 0, sizeof(PartialHeader));
  double* d_pts = nullptr;
  int rc = upload_points(ctx, pts_xy, size_t(n_pts), &d_pts);
  if (rc) return rc;
  double pose[3] = {pose_world[0], pose_world[1], pose_world[2]};
  double cov[9];
  std::vector<PassItem> items(1);
  items[0].grid = grid; items[0].d_pts = d_pts; items[0].P = n_pts; items[0].param = *param;
  items[0].pose_world = pose; items[0].cov = cov;
  items[0].ang_begin = angle_begin; items[0].ang_end = angle_end;
  rc = run_pass(ctx, items, MODE_PARTIAL, reinterpret_cast<double*>(partial), 0, nullptr);
  if (rc) return rc;
  if (!ctx->slice.valid) {   // empty slice: still hand over a well-formed (empty) partial
    PartialHeader H;
    std::memset(&H, 0, sizeof H);
    H.magic = kPartialMagic; H.a0 = H.a1 = std::max(0, angle_begin);
    std::memcpy(partial, &H, sizeof H);
    double center[3];
    grid->tf.world_to_map(pose_world, center);
    SliceState& S = ctx->slice;
    S.valid = true; S.merged = false; S.exact_needed = false; S.param = *param; S.grid = grid;
    S.geo = make_geo(*param, n_pts, grid->cell_len(), center);
    S.a0 = S.a1 = H.a0; S.d_score = nullptr; S.d_gjob = nullptr; S.d_gout = nullptr;
  }
  return RSM_OK;
}

}  // extern "C"

namespace {
// Merge every rank's partial: global maximum, averaging set, best pose, global top list, the same-(x,y) columns the
// angular covariance needs.  Leaves the result in ctx->slice (exact_needed when the reference's sort order matters).
int merge_core(rsm_ctx* ctx, const void* const* partials, int n_partials) {
  SliceState& S = ctx->slice;
  if (!S.valid) return fail(ctx, RSM_ERR_INVALID, "rsm_match_merge: no rsm_match_partial pending on this context");
  const PassGeo& g = S.geo;
  unsigned long long best_key = 0ull;
  bool exact = false;
  for (int r = 0; r < n_partials; ++r) {
    PartialHeader H;
    std::memcpy(&H, partials[r], sizeof H);
    if (H.magic != kPartialMagic) return fail(ctx, RSM_ERR_INVALID, "partial %d is not an rsm partial", r);
    if (H.a1 > H.a0 && (H.n_ang != g.n_ang || H.n_xy != g.n_xy)) return fail(ctx, RSM_ERR_INVALID, "partial %d comes from another window", r);
    if (H.flags & 2) return fail(ctx, RSM_ERR_WINDOW, "search window + scan extent leaves the grid (slice %d)", r);
    if (H.flags & 1) exact = true;
    best_key = std::max(best_key, H.best_key);
  }
  const double top_score = key_to_score(best_key);
  std::vector<Cand> a_list;
  S.top.clear();
  for (int r = 0; r < n_partials && !exact; ++r) {
    PartialHeader H;
    std::memcpy(&H, partials[r], sizeof H);
    const Cand* e = reinterpret_cast<const Cand*>(static_cast<const char*>(partials[r]) + sizeof(PartialHeader));
    for (int i = 0; i < H.n_pool; ++i)
      if (DoubleEqual(e[i].score, top_score, kResponseFilterTolerance)) a_list.push_back(e[i]);
    for (int i = 0; i < H.n_top; ++i) S.top.push_back(e[H.n_pool + i]);
  }
  S.n_cols = 0;
  if (!exact) {
    std::sort(a_list.begin(), a_list.end(), by_score_desc);
    if (a_list.empty() || a_list[0].score != top_score || adjacent_tie(a_list, a_list.size())) exact = true;
  }
  if (!exact) {
    S.best = find_best(g, a_list.data(), a_list.size());
    if (S.top.size() > size_t(kTopK)) { std::nth_element(S.top.begin(), S.top.begin() + kTopK, S.top.end(), by_score_desc); S.top.resize(kTopK); }
    std::sort(S.top.begin(), S.top.end(), by_score_desc);
    const int type = S.param.type;
    if ((type == RSM_COARSE || type == RSM_SUPER) && !(S.best.score < kDoubleTolerance)) {
      int xs[8], ys[8], nx = 0, ny = 0;
      const double tol = g.factor;
      for (int ix = 0; ix < g.n_xy && nx < 8; ++ix) if (DoubleEqual(g.x_of(ix), S.best.x, tol)) xs[nx++] = ix;
      for (int iy = 0; iy < g.n_xy && ny < 8; ++iy) if (DoubleEqual(g.y_of(iy), S.best.y, tol)) ys[ny++] = iy;
      if (nx * ny > kMaxCols) exact = true;
      else for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) S.cols[S.n_cols++] = xs[i] * g.n_xy + ys[j];
    }
  }
  if (exact) S.n_cols = 0;
  S.exact_needed = exact;
  S.merged = true;
  return RSM_OK;
}

// one rank's same-(x,y) column scores: data[c * stride + (ia - a0)]
struct ColView { int a0, a1, n_cols; const int* cols; const double* data; int stride; };

// Everything after the columns are known; RSM_NEED_EXACT (nothing written) when ties decide.
int finish_core(rsm_ctx* ctx, const ColView* views, int n_views, double pose_world[3], double cov[9], double* response,
                rsm_pass_detail* detail) {
  SliceState& S = ctx->slice;
  *response = 0.0;
  if (detail) std::memset(detail, 0, sizeof *detail);
  // exact ties in a consumed set: the slice stays valid for rsm_match_slice_scores / rsm_match_finish_exact
  const char* need_exact = "exact score ties in the consumed candidate sets: gather the slices (rsm_match_slice_scores) and call rsm_match_finish_exact";
  if (S.exact_needed) return fail(ctx, RSM_NEED_EXACT, "%s", need_exact);
  const PassGeo& g = S.geo;
  const int type = S.param.type;
  const double bound = cov_score_bound(S.best);
  double cov_new[9];
  std::memcpy(cov_new, cov, sizeof cov_new);
  if (type == RSM_COARSE || type == RSM_FINE) {
    if (S.top.size() == size_t(kTopK) && S.top[kTopK - 1].score > bound && S.top[kTopK - 1].score == S.top[kTopK - 2].score)
      return fail(ctx, RSM_NEED_EXACT, "%s", need_exact);
    if (ctx->strict_ties && adjacent_tie(S.top, positional_prefix(S.top, bound))) return fail(ctx, RSM_NEED_EXACT, "%s", need_exact);
    positional_cov(g, S.param, S.best, S.top.data(), S.top.size(), cov_new);
  }
  if (type == RSM_COARSE || type == RSM_SUPER) {
    std::vector<Cand> xy;
    if (!(S.best.score < kDoubleTolerance)) {
      for (int r = 0; r < n_views; ++r) {
        const ColView& V = views[r];
        const int nang = V.a1 - V.a0;
        for (int c = 0; c < V.n_cols; ++c)
          for (int ia = 0; ia < nang; ++ia) {
            const double s = V.data[size_t(c) * V.stride + ia];
            if (s >= bound) xy.push_back(Cand{s, int64_t(V.a0 + ia) * g.n_xy * g.n_xy + V.cols[c]});
          }
      }
      if (xy.size() > size_t(kTopK)) { std::nth_element(xy.begin(), xy.begin() + kTopK, xy.end(), by_score_desc); xy.resize(kTopK); }
      std::sort(xy.begin(), xy.end(), by_score_desc);
      if (xy.size() > size_t(kMaxVarianceUsePointSize) && xy[kMaxVarianceUsePointSize].score == xy[kMaxVarianceUsePointSize - 1].score)
        return fail(ctx, RSM_NEED_EXACT, "%s", need_exact);
      if (ctx->strict_ties && adjacent_tie(xy, kMaxVarianceUsePointSize)) return fail(ctx, RSM_NEED_EXACT, "%s", need_exact);
    }
    angular_cov(g, S.param, S.best, xy.data(), xy.size(), cov_new);
  }
  S.valid = false;
  std::memcpy(cov, cov_new, sizeof cov_new);
  const double bs = S.best.score;
  *response = bs > 1.0 ? 1.0 : bs;
  rsm_pass_detail d;
  std::memset(&d, 0, sizeof d);
  d.best_score = bs;
  d.best_pose_map[0] = S.best.x; d.best_pose_map[1] = S.best.y; d.best_pose_map[2] = S.best.angle;
  d.n_candidates = g.n_cand(); d.n_avg = S.best.n_avg;
  d.n_ang = g.n_ang; d.n_xy = g.n_xy; d.visited = g.visited; d.divisor = g.divisor;
  if (*response > S.param.response_threshold) {
    const double b[3] = {S.best.x, S.best.y, S.best.angle};
    S.grid->tf.map_to_world(b, pose_world);
    d.pose_updated = 1;
  }
  if (detail) *detail = d;
  ctx->stats.passes++;
  return RSM_OK;
}

// the reference's own sort on the whole candidate array (correlate_scan_matcher.h:607-608), slices in angle order
int finish_exact_slices(rsm_ctx* ctx, const double* const* slices, const int64_t* counts, int n_slices, double pose_world[3],
                        double cov[9], double* response, rsm_pass_detail* detail) {
  SliceState& S = ctx->slice;
  const PassGeo& g = S.geo;
  int64_t total = 0;
  for (int r = 0; r < n_slices; ++r) {
    if (counts[r] < 0 || (counts[r] > 0 && !slices[r]) || counts[r] % (int64_t(g.n_xy) * g.n_xy) != 0)
      return fail(ctx, RSM_ERR_INVALID, "rsm_match_finish_exact: slice %d is not a whole number of angles", r);
    total += counts[r];
  }
  if (total != g.n_cand()) return fail(ctx, RSM_ERR_INVALID, "rsm_match_finish_exact: the slices hold %lld scores, the window has %lld candidates",
                                       (long long)total, (long long)g.n_cand());
  *response = 0.0;
  if (detail) std::memset(detail, 0, sizeof *detail);
  std::vector<Cand> c;
  c.resize(static_cast<size_t>(total));
  int64_t k = 0;
  for (int r = 0; r < n_slices; ++r)
    for (int64_t i = 0; i < counts[r]; ++i, ++k) { c[k].score = slices[r][i]; c[k].index = k; }
  BestPose best;
  finish_exact_core(g, S.param, c, best, cov);
  S.valid = false;
  const double bs = best.score;
  *response = bs > 1.0 ? 1.0 : bs;
  rsm_pass_detail d;
  std::memset(&d, 0, sizeof d);
  d.best_score = bs;
  d.best_pose_map[0] = best.x; d.best_pose_map[1] = best.y; d.best_pose_map[2] = best.angle;
  d.n_candidates = g.n_cand(); d.n_avg = best.n_avg; d.exact_sort_used = 1;
  d.n_ang = g.n_ang; d.n_xy = g.n_xy; d.visited = g.visited; d.divisor = g.divisor;
  if (*response > S.param.response_threshold) {
    const double b[3] = {best.x, best.y, best.angle};
    S.grid->tf.map_to_world(b, pose_world);
    d.pose_updated = 1;
  }
  if (detail) *detail = d;
  ctx->stats.passes++;
  ctx->stats.exact_sort_passes++;
  return RSM_OK;
}
}  // namespace

extern "C" {

int rsm_match_merge(rsm_ctx* ctx, const void* const* partials, int n_partials, void* columns) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !partials || n_partials < 1 || !columns) return fail(ctx, RSM_ERR_INVALID, "rsm_match_merge: bad arguments");
  int rc = merge_core(ctx, partials, n_partials);
  if (rc) return rc;
  SliceState& S = ctx->slice;
  const PassGeo& g = S.geo;
  // this rank's column scores
  ColumnsHeader CH;
  std::memset(&CH, 0, sizeof CH);
  CH.magic = kPartialMagic; CH.a0 = S.a0; CH.a1 = S.a1; CH.n_cols = S.n_cols;
  for (int c = 0; c < CH.n_cols; ++c) CH.cols[c] = S.cols[c];
  const int nang = S.a1 - S.a0;
  const size_t need = sizeof(ColumnsHeader) + size_t(CH.n_cols) * nang * 8;
  if (need > RSM_COLUMNS_BYTES) return fail(ctx, RSM_ERR_UNSUPPORTED, "angle slice too long for RSM_COLUMNS_BYTES (%d angles)", nang);
  std::memcpy(columns, &CH, sizeof CH);
  if (CH.n_cols > 0 && nang > 0) {
    GatherJob G;
    std::memset(&G, 0, sizeof G);
    G.score = S.d_score; G.out = S.d_gout; G.n_xy = g.n_xy; G.n_ang = nang; G.n_cols = CH.n_cols;
    for (int c = 0; c < CH.n_cols; ++c) G.cols[c] = CH.cols[c];
    rc = ensure_pinned(ctx, ctx->h_up, sizeof G);
    if (rc) return rc;
    rc = ensure_pinned(ctx, ctx->h_down, size_t(CH.n_cols) * nang * 8);
    if (rc) return rc;
    std::memcpy(ctx->h_up.p, &G, sizeof G);
    CU(cudaMemcpyAsync(S.d_gjob, ctx->h_up.p, sizeof G, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_gather(1, ctx->stream, reinterpret_cast<const GatherJob*>(S.d_gjob)));
    ctx->stats.kernel_launches++;
    CU(cudaMemcpyAsync(ctx->h_down.p, S.d_gout, size_t(CH.n_cols) * nang * 8, cudaMemcpyDeviceToHost, ctx->stream));
    rc = sync_stream(ctx);
    if (rc) return rc;
    std::memcpy(static_cast<char*>(columns) + sizeof(ColumnsHeader), ctx->h_down.p, size_t(CH.n_cols) * nang * 8);
  }
  return RSM_OK;
}

int rsm_match_finish(rsm_ctx* ctx, const void* const* partials, int n_partials, const void* const* columns,
                     double pose_world[3], double cov[9], double* response, rsm_pass_detail* detail) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !partials || !columns || n_partials < 1 || !pose_world || !cov || !response)
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_finish: bad arguments");
  SliceState& S = ctx->slice;
  if (!S.valid || !S.merged) return fail(ctx, RSM_ERR_INVALID, "rsm_match_finish: call rsm_match_partial and rsm_match_merge first");
  std::vector<ColView> views(n_partials);
  std::vector<ColumnsHeader> heads(n_partials);
  for (int r = 0; r < n_partials; ++r) {
    std::memcpy(&heads[r], columns[r], sizeof(ColumnsHeader));
    if (heads[r].magic != kPartialMagic) return fail(ctx, RSM_ERR_INVALID, "columns %d is not an rsm column block", r);
    views[r] = ColView{heads[r].a0, heads[r].a1, heads[r].n_cols, heads[r].cols,
                       reinterpret_cast<const double*>(static_cast<const char*>(columns[r]) + sizeof(ColumnsHeader)),
                       heads[r].a1 - heads[r].a0};
  }
  return finish_core(ctx, views.data(), n_partials, pose_world, cov, response, detail);
}

int rsm_match_slice_scores(rsm_ctx* ctx, double* scores_out, int64_t capacity, int64_t* n_slice) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !n_slice || capacity < 0 || (capacity > 0 && !scores_out)) return fail(ctx, RSM_ERR_INVALID, "rsm_match_slice_scores: bad arguments");
  SliceState& S = ctx->slice;
  if (!S.valid) return fail(ctx, RSM_ERR_INVALID, "rsm_match_slice_scores: no rsm_match_partial pending on this context");
  const int64_t n = int64_t(S.a1 - S.a0) * S.geo.n_xy * S.geo.n_xy;
  *n_slice = n;
  if (capacity == 0 || n == 0) return RSM_OK;
  if (capacity < n) return fail(ctx, RSM_ERR_INVALID, "rsm_match_slice_scores: need room for %lld scores", (long long)n);
  CU(cudaMemcpyAsync(scores_out, S.d_score, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  int rc = sync_stream(ctx);
  if (rc) return rc;
  ctx->stats.d2h_bytes += size_t(n) * 8;
  return RSM_OK;
}

int rsm_match_finish_exact(rsm_ctx* ctx, const double* const* slices, const int64_t* counts, int n_slices, double pose_world[3],
                           double cov[9], double* response, rsm_pass_detail* detail) {
  if (!ctx || !slices || !counts || n_slices < 1 || !pose_world || !cov || !response)
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_finish_exact: bad arguments");
  if (!ctx->slice.valid) return fail(ctx, RSM_ERR_INVALID, "rsm_match_finish_exact: no rsm_match_partial pending on this context");
  return finish_exact_slices(ctx, slices, counts, n_slices, pose_world, cov, response, detail);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time (no link dependency: a caller that never shards one window never loads it)
// ---------------------------------------------------------------------------------------------
namespace {
struct NcclId { char internal[128]; };      // ncclUniqueId
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy the process already holds (torch loads its own libnccl.so.2), else the system's
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather) api.lib = h;
  });
  return api.lib ? &api : nullptr;
}
constexpr int kNcclChar = 0;      // ncclInt8 / ncclChar

// items [begin, end) of `rank`: sizes differ by at most one, earlier ranks get the extra (roborts_edu_slam_b200/sharding.py)
void contiguous_range(int n_items, int rank, int world, int* begin, int* end) {
  const int base = n_items / world, extra = n_items % world;
  *begin = rank * base + std::min(rank, extra);
  *end = *begin + base + (rank < extra ? 1 : 0);
}
}  // namespace

extern "C" {

int rsm_comm_unique_id(void* id_out) {
  if (!id_out) return RSM_ERR_INVALID;
  NcclApi* N = nccl_api();
  if (!N) return RSM_ERR_UNSUPPORTED;
  NcclId id;
  if (N->GetUniqueId(&id) != 0) return RSM_ERR_CUDA;
  std::memcpy(id_out, &id, sizeof id);
  return RSM_OK;
}

int rsm_comm_init(rsm_ctx* ctx, int rank, int world_size, const void* id) {
  DeviceGuard device_guard(ctx);
  if (!ctx || world_size < 1 || rank < 0 || rank >= world_size || (world_size > 1 && !id))
    return fail(ctx, RSM_ERR_INVALID, "rsm_comm_init: bad arguments");
  if (ctx->comm) return fail(ctx, RSM_ERR_INVALID, "rsm_comm_init: this context already has a communicator");
  ctx->comm_rank = rank; ctx->comm_world = world_size;
  if (world_size == 1) return RSM_OK;
  NcclApi* N = nccl_api();
  if (!N) return fail(ctx, RSM_ERR_UNSUPPORTED, "rsm_comm_init: libnccl.so.2 could not be loaded");
  NcclId nid;
  std::memcpy(&nid, id, sizeof nid);
  const int r = N->CommInitRank(&ctx->comm, world_size, nid, rank);
  if (r != 0) { ctx->comm = nullptr; return fail(ctx, RSM_ERR_CUDA, "ncclCommInitRank failed: %s", N->GetErrorString ? N->GetErrorString(r) : "?"); }
  return RSM_OK;
}

int rsm_comm_destroy(rsm_ctx* ctx) {
  DeviceGuard device_guard(ctx);
  if (!ctx) return RSM_ERR_INVALID;
  if (ctx->comm) {
    cudaStreamSynchronize(ctx->stream);
    if (NcclApi* N = nccl_api()) N->CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->comm_rank = 0; ctx->comm_world = 1;
  return RSM_OK;
}

// The whole angle-sliced match on every rank of the communicator: score + select the own slice, pack the partial on
// the device, ncclAllGather, merge on the host, gather the same-(x,y) columns into the exchange buffer, ncclAllGather,
// finish.  Two stream synchronisations per match; the collectives run on the context's own stream, device to device.
int rsm_match_sliced(rsm_ctx* ctx, const rsm_grid* grid, const double* pts_xy, int n_pts, const rsm_pass_param* param,
                     double pose_world[3], double cov[9], double* response, rsm_pass_detail* detail) {
  DeviceGuard device_guard(ctx);
  if (!ctx || !grid || !param || !pose_world || !cov || !response || n_pts < 0 || (n_pts > 0 && !pts_xy))
    return fail(ctx, RSM_ERR_INVALID, "rsm_match_sliced: bad arguments");
  *response = 0.0;
  if (detail) std::memset(detail, 0, sizeof *detail);
  if (!grid->init || n_pts == 0) return RSM_OK;   // correlate_scan_matcher.h:792-795
  const int world = ctx->comm_world, rank = ctx->comm_rank;
  NcclApi* N = world > 1 ? nccl_api() : nullptr;
  if (world > 1 && (!N || !ctx->comm)) return fail(ctx, RSM_ERR_INVALID, "rsm_match_sliced: call rsm_comm_init first");
  ctx->slice.valid = false;
  double* d_pts = nullptr;
  int rc = upload_points(ctx, pts_xy, size_t(n_pts), &d_pts);
  if (rc) return rc;
  double pose[3] = {pose_world[0], pose_world[1], pose_world[2]};
  double cov_unused[9];
  std::vector<PassItem> items(1);
  PassItem& it = items[0];
  it.grid = grid; it.d_pts = d_pts; it.P = n_pts; it.param = *param; it.pose_world = pose; it.cov = cov_unused;
  std::vector<int> act;
  rc = pass_geometry(ctx, items, act);      // full window first: n_ang decides the slices
  if (rc) return rc;
  if (act.empty()) return fail(ctx, RSM_ERR_INVALID, "rsm_match_sliced: empty search window");
  const PassGeo geo = it.geo;
  int max_nang = 0;
  std::vector<int> a0s(world), a1s(world);
  for (int r = 0; r < world; ++r) { contiguous_range(geo.n_ang, r, world, &a0s[r], &a1s[r]); max_nang = std::max(max_nang, a1s[r] - a0s[r]); }
  it.ang_begin = a0s[rank]; it.ang_end = a1s[rank];
  const bool have_slice = a1s[rank] > a0s[rank];
  // exchange buffers: [own partial | all partials] then, reused, [own columns | all columns]
  const size_t col_bytes = size_t(kMaxCols) * max_nang * 8;
  const size_t unit = std::max<size_t>(kPartialDevBytes, (col_bytes + 255) / 256 * 256);
  rc = ensure_dev(ctx, ctx->d_xchg, unit * size_t(world + 1));
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_xchg, unit * size_t(world + 1));
  if (rc) return rc;
  char* dx = ctx->d_xchg.p;
  char* hx = ctx->h_xchg.p;
  Lane* lane = &ctx->L0;
  cudaStream_t st = ctx->stream;
  PassRun R;
  std::vector<char> own_partial(RSM_PARTIAL_BYTES);
  if (have_slice) {
    rc = pass_geometry(ctx, items, act);
    if (rc) return rc;
    rc = pass_begin(ctx, lane, items, act, MODE_PARTIAL, reinterpret_cast<double*>(own_partial.data()), 0, nullptr, R);
    if (rc) return rc;
    // the partial, packed where the selection left it
    char* dw = lane->d_work.p;
    PackJob PJ;
    std::memset(&PJ, 0, sizeof PJ);
    PJ.best_key = reinterpret_cast<const unsigned long long*>(dw + R.o_best);
    PJ.err = reinterpret_cast<const int*>(dw + R.o_err);
    PJ.pool_count = reinterpret_cast<const int*>(dw + R.o_poolcnt);
    PJ.pool = reinterpret_cast<const PoolEntry*>(dw + R.o_pool);
    PJ.ftop = reinterpret_cast<const Entry*>(dw + R.o_ftop);
    PJ.fcnt = reinterpret_cast<const int*>(dw + R.o_fcnt);
    PJ.base = int64_t(it.a0) * geo.n_xy * geo.n_xy;
    PJ.pool_cap = R.pool_cap; PJ.a0 = it.a0; PJ.a1 = it.a1; PJ.n_ang = geo.n_ang; PJ.n_xy = geo.n_xy;
    PJ.out = dx;
    CU(launch_pack_partial(st, PJ));
    lane->stats.kernel_launches++;
  } else {
    // more ranks than angles: a well-formed empty partial
    PartialHeader H;
    std::memset(&H, 0, sizeof H);
    H.magic = kPartialMagic; H.a0 = H.a1 = a0s[rank];
    std::memcpy(hx, &H, sizeof H);
    CU(cudaMemcpyAsync(dx, hx, sizeof H, cudaMemcpyHostToDevice, st));
  }
  if (world > 1) {
    const int r = N->AllGather(dx, dx + unit, kPartialDevBytes, kNcclChar, ctx->comm, st);
    if (r != 0) return fail(ctx, RSM_ERR_CUDA, "ncclAllGather(partials) failed: %s", N->GetErrorString ? N->GetErrorString(r) : "?");
  } else {
    CU(cudaMemcpyAsync(dx + unit, dx, kPartialDevBytes, cudaMemcpyDeviceToDevice, st));
  }
  CU(cudaMemcpyAsync(hx + unit, dx + unit, size_t(kPartialDevBytes) * world, cudaMemcpyDeviceToHost, st));
  // one wait for: scoring, selection, the own read-back, the exchange and the gathered partials
  if (have_slice) {
    R.blocking = false;
    rc = pass_end(ctx, R);        // sets up ctx->slice (scores stay on the device)
    if (rc) return rc;
  } else {
    rc = sync_stream(ctx);
    if (rc) return rc;
    SliceState& S0 = ctx->slice;
    S0.valid = true; S0.merged = false; S0.exact_needed = false; S0.param = *param; S0.grid = grid; S0.geo = geo;
    S0.a0 = S0.a1 = a0s[rank]; S0.d_score = nullptr; S0.d_gjob = nullptr; S0.d_gout = nullptr;
  }
  ctx->stats.d2h_bytes += size_t(kPartialDevBytes) * world;
  std::vector<const void*> parts(world);
  for (int r = 0; r < world; ++r) parts[r] = hx + unit + size_t(r) * kPartialDevBytes;
  rc = merge_core(ctx, parts.data(), world);
  if (rc) return rc;
  SliceState& S = ctx->slice;
  if (!S.exact_needed) {
    // same-(x,y) columns of every slice: [c * max_nang + ia] per rank; every rank derives the same column list
    const int nang = S.a1 - S.a0;
    if (S.n_cols > 0) {
      if (nang > 0) {
        GatherJob G;
        std::memset(&G, 0, sizeof G);
        G.score = S.d_score; G.out = reinterpret_cast<double*>(dx); G.n_xy = geo.n_xy; G.n_ang = nang; G.n_cols = S.n_cols;
        for (int c = 0; c < S.n_cols; ++c) G.cols[c] = S.cols[c];
        CU(launch_gather_columns(st, G, max_nang));
        ctx->stats.kernel_launches++;
      }
      const size_t msg = size_t(S.n_cols) * max_nang * 8;
      if (world > 1) {
        const int r = N->AllGather(dx, dx + unit, msg, kNcclChar, ctx->comm, st);
        if (r != 0) return fail(ctx, RSM_ERR_CUDA, "ncclAllGather(columns) failed: %s", N->GetErrorString ? N->GetErrorString(r) : "?");
      } else {
        CU(cudaMemcpyAsync(dx + unit, dx, msg, cudaMemcpyDeviceToDevice, st));
      }
      CU(cudaMemcpyAsync(hx + unit, dx + unit, msg * world, cudaMemcpyDeviceToHost, st));
      rc = sync_stream(ctx);
      if (rc) return rc;
      ctx->stats.d2h_bytes += msg * world;
    }
    std::vector<ColView> views(world);
    const size_t msg = size_t(S.n_cols) * max_nang * 8;
    for (int r = 0; r < world; ++r)
      views[r] = ColView{a0s[r], a1s[r], S.n_cols, S.cols, reinterpret_cast<const double*>(hx + unit + size_t(r) * msg), max_nang};
    rc = finish_core(ctx, views.data(), world, pose_world, cov, response, detail);
    if (rc != RSM_NEED_EXACT) return rc;
  }
  // exact ties: every rank gets every slice's scores (device to device), then runs the reference's sort itself
  const int64_t plane = int64_t(geo.n_xy) * geo.n_xy;
  const size_t slot = size_t(max_nang) * size_t(plane) * 8;
  rc = ensure_dev(ctx, ctx->d_xchg, slot * size_t(world + 1));
  if (rc) return rc;
  rc = ensure_pinned(ctx, ctx->h_xchg, slot * size_t(world));
  if (rc) return rc;
  dx = ctx->d_xchg.p; hx = ctx->h_xchg.p;
  const int64_t mine = int64_t(S.a1 - S.a0) * plane;
  if (mine > 0) CU(cudaMemcpyAsync(dx, S.d_score, size_t(mine) * 8, cudaMemcpyDeviceToDevice, st));
  if (world > 1) {
    const int r = N->AllGather(dx, dx + slot, slot, kNcclChar, ctx->comm, st);
    if (r != 0) return fail(ctx, RSM_ERR_CUDA, "ncclAllGather(scores) failed: %s", N->GetErrorString ? N->GetErrorString(r) : "?");
  } else {
    CU(cudaMemcpyAsync(dx + slot, dx, slot, cudaMemcpyDeviceToDevice, st));
  }
  CU(cudaMemcpyAsync(hx, dx + slot, slot * world, cudaMemcpyDeviceToHost, st));
  rc = sync_stream(ctx);
  if (rc) return rc;
  ctx->stats.d2h_bytes += slot * world;
  std::vector<const double*> sl(world);
  std::vector<int64_t> cnt(world);
  for (int r = 0; r < world; ++r) { sl[r] = reinterpret_cast<const double*>(hx + size_t(r) * slot); cnt[r] = int64_t(a1s[r] - a0s[r]) * plane; }
  return finish_exact_slices(ctx, sl.data(), cnt.data(), world, pose_world, cov, response, detail);
}

}  // extern "C"

// rsm_kernels.h -- launchers of the kernels in rsm_kernels.cu (host-callable).
#ifndef RSM_KERNELS_H_
#define RSM_KERNELS_H_

#include <cuda_runtime.h>
#include <stddef.h>

#include "rsm_device.h"

namespace rsm {

// Scoring kernel variants (rsm_score.cu).  fixed: int32 2^-25 fixed-point cells (else float32);
// affine: search step is an exact integer number of cells (lx in {8,16,32}), else lx in {4,8,16,32};
// ry in 1..8 rows per thread.  A CTA has score_threads(lx) threads and covers lx x score_rows(lx, ry)
// translations of one angle.
int score_threads(int lx);
inline int score_rows(int lx, int ry) { return (score_threads(lx) / lx) * ry; }
// const_pitch: kPitchSmall / kPitchLarge if every job has that row pitch and a unit search step, else 0
int score_occupancy(bool fixed, bool affine, int lx, int ry, int const_pitch);
cudaError_t launch_score(bool fixed, bool affine, int lx, int ry, int const_pitch, int n_cta, cudaStream_t st,
                         const ScoreJob* jobs, const int* cta_begin, int n_jobs);
// Flat variant (small windows, non-integer step): k candidates per thread, k * 256 consecutive candidates per CTA.
int score_flat_max_nxy();
int score_flat_k(int n_xy);
int score_flat_ctas(int n_local, int k);
cudaError_t launch_score_flat(bool fixed, int k, int n_cta, cudaStream_t st, const ScoreJob* jobs, const int* cta_begin, int n_jobs);
// Patch variant (batches of small unit-step windows, fixed point): one CTA = score_patch_angles() consecutive search
// angles of one job, the whole window (n_xy <= 16) per angle; grid cells reach shared memory as one box per 32 beams.
int score_patch_angles();
cudaError_t launch_score_patch(int n_xy, int n_cta, cudaStream_t st, const ScoreJob* jobs, const int* cta_begin, int n_jobs);
// Staged variant (shared-memory window filled by TMA bulk copies): fixed-point grid, unit search step.
// A CTA covers score_staged_tile(variant) translations of one angle and 1/n_split of the beams; the
// n_split CTAs of a tile are launched as one thread-block cluster and combine their sums through
// distributed shared memory.  score_staged_variant picks the tile shape for a window size.
int score_staged_variant(int n_xy);
void score_staged_tile(int variant, int* tile_x, int* tile_y);
void score_staged_boxes(int box_w[2], int box_h[2]);   // the two TMA box shapes (cells)
size_t score_staged_smem(int beams_per_split);
int score_staged_resident_ctas(int variant, int n_split, int max_beams_per_split);
cudaError_t launch_score_staged(int variant, int n_split, int n_cta, int max_beams_per_split, cudaStream_t st,
                                const ScoreJob* jobs, const int* cta_begin, int n_jobs);
// Stream plan of the staged variant: n_cta persistent CTAs, each with a contiguous share (StreamCta) of the launch's
// (job, angle, tile, beam) sequence; item_begin[j] = first item of job j (items = ang_count * tiles_x * tiles_y).
// variant: score_stream_variant(n_xy); tiles, TMA boxes and per-beam cost of a (partial) tile come from the variant.
int score_stream_variant(int n_xy);
void score_stream_tile(int variant, int* tile_x, int* tile_y);
void score_stream_boxes(int variant, int box_w[2], int box_h[2]);
int score_stream_weight(int variant, int ext_x, int ext_y);
size_t score_stream_partial_words(int variant);   // 64-bit words per partial slot
size_t score_stream_smem(int variant, int max_beams);
cudaError_t launch_score_stream(int variant, int n_cta, int max_beams, cudaStream_t st, const ScoreJob* jobs, const int* item_begin,
                                int n_jobs, const StreamCta* plan, unsigned long long* partials, int* tickets);
cudaError_t launch_select(int n_cta, cudaStream_t st, const SelectJob* jobs, const int* cta_begin,
                          int n_jobs, PoolEntry* pool, int pool_cap, int* pool_count);
cudaError_t launch_gather(int n_jobs, cudaStream_t st, const GatherJob* jobs);
cudaError_t launch_pack_partial(cudaStream_t st, const PackJob& job);
cudaError_t launch_gather_columns(cudaStream_t st, const GatherJob& job, int stride);
cudaError_t launch_fill(int n_jobs, int ctas_per_job, cudaStream_t st, const FillJob* jobs);
cudaError_t launch_penalty(int n_jobs, cudaStream_t st, const PenaltyJob* jobs, const unsigned char* occ, int size_x,
                           int size_y, double bound_tolerance);
cudaError_t launch_pub_update(cudaStream_t st, PubScan* scan);
cudaError_t launch_pub_occupancy(cudaStream_t st, const float* pass, const float* prob, long long n_cells, float occu_threshold,
                                 float min_pass_through, unsigned char* occ);
cudaError_t launch_fill_f32(cudaStream_t st, float* p, long long n, float v);
cudaError_t launch_optimize(int n_jobs, cudaStream_t st, const OptimizeJob* jobs);
cudaError_t launch_raster(int n_scans, cudaStream_t st, const RasterScan* scans, const int* stamp,
                          int half, int one);
// non-blur update (SET_CELL_OCCUPIED): one CTA per grid, its scans [group_begin[g], group_begin[g+1]) applied in order
cudaError_t launch_raster_occu(int n_groups, cudaStream_t st, const RasterScan* scans, const int* group_begin, int half, int fixed);
cudaError_t launch_microbench(cudaStream_t st, int mode, const int* g, unsigned int words, int iters,
                              int n_cta, unsigned long long* sink);
cudaError_t launch_flush(cudaStream_t st, void* buf, long long bytes, int v);

}  // namespace rsm
#endif

// rsm_kernels.h -- launchers of the kernels in rsm_kernels.cu (host-callable).
#ifndef RSM_KERNELS_H_
#define RSM_KERNELS_H_

#include <cuda_runtime.h>
#include <stddef.h>

#include "rsm_device.h"

namespace rsm {

size_t score_smem_bytes(int lx, int ry, int nt);
// fixed: int32 2^-25 fixed-point cells (else float32).  lx in {4,8,16,32}, ry in 1..8, nt in {128,256}.
cudaError_t launch_score(bool fixed, int lx, int ry, int nt, int n_cta, cudaStream_t st,
                         const ScoreJob* jobs, const int* cta_begin, int n_jobs);
cudaError_t launch_select(int n_cta, cudaStream_t st, const SelectJob* jobs, const int* cta_begin,
                          int n_jobs, PoolEntry* pool, int pool_cap, int* pool_count);
cudaError_t launch_gather(int n_jobs, cudaStream_t st, const GatherJob* jobs);
cudaError_t launch_fill(int n_jobs, int ctas_per_job, cudaStream_t st, const FillJob* jobs);
cudaError_t launch_raster(int n_scans, cudaStream_t st, const RasterScan* scans, const int* stamp,
                          int half, int one);
cudaError_t launch_flush(cudaStream_t st, void* buf, long long bytes, int v);

}  // namespace rsm
#endif

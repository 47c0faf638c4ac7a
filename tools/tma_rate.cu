// Probe: how fast does one SM ingest 2D TMA boxes of an L2-resident int32 grid, alone and with every SM doing the same?
//   tma_rate <ctas> <box_w> <box_h> <split> <depth> <align> [iters]
// Each CTA loads `iters` logical boxes of box_w x box_h cells from random places of a 1120 x 1120 grid (pitch 4128 cells,
// as BASELINE configs[1]); a logical box is issued as `split` TMA copies of box_h / split rows each; `depth` logical boxes
// are in flight (own buffer + mbarrier each); x is rounded down to a multiple of `align` cells (4 = 16 bytes is the least
// TMA takes).  Prints cycles per logical box (mean over CTAs) and bytes per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_rate tools/tma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void rate(const __grid_constant__ CUtensorMap tm, int box_w, int sub_h, int split, int depth, int align, int iters,
                     int size, long long* cycles, int* sink) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) uint64_t bar[8];
  int* buf = reinterpret_cast<int*>(dyn);
  const int box_cells = box_w * sub_h * split;
  if (threadIdx.x == 0) {
    for (int d = 0; d < depth; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[d])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int rng = 12345u + 977u * blockIdx.x;
    auto issue = [&](int d) {
      rng = rng * 1664525u + 1013904223u;
      const int x = int((rng >> 8) % (unsigned)(size - box_w)) / align * align;
      rng = rng * 1664525u + 1013904223u;
      const int y = int((rng >> 8) % (unsigned)(size - sub_h * split));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[d])), "r"(box_cells * 4) : "memory");
      for (int k = 0; k < split; ++k)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         s32(buf + d * box_cells + k * box_w * sub_h)),
                     "l"(&tm), "r"(x), "r"(y + k * sub_h), "r"(s32(&bar[d]))
                     : "memory");
    };
    for (int d = 0; d < depth; ++d) issue(d);
    const long long t0 = clock64();
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
      const int d = i % depth;
      const uint32_t parity = (i / depth) & 1;
      asm volatile(
          "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(
              s32(&bar[d])),
          "r"(parity)
          : "memory");
      acc += buf[d * box_cells + (i & 63)];
      if (i + depth < iters) issue(d);
    }
    cycles[blockIdx.x] = clock64() - t0;
    sink[blockIdx.x] = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 148, bw = argc > 2 ? atoi(argv[2]) : 160, bh = argc > 3 ? atoi(argv[3]) : 128;
  const int split = argc > 4 ? atoi(argv[4]) : 1, depth = argc > 5 ? atoi(argv[5]) : 1, align = argc > 6 ? atoi(argv[6]) : 4;
  const int iters = argc > 7 ? atoi(argv[7]) : 200;
  const int size = 1120, pitch = 4128;
  if (bh % split || depth > 8 || size_t(bw) * bh * 4 * depth > 220 * 1024) { printf("bad shape\n"); return 2; }
  int* d; cudaMalloc(&d, size_t(pitch) * size * 4); cudaMemset(d, 1, size_t(pitch) * size * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encoder\n"); return 2; }
  alignas(64) CUtensorMap tm;
  const cuuint64_t dims[2] = {cuuint64_t(size), cuuint64_t(size)}; const cuuint64_t strides[1] = {cuuint64_t(pitch) * 4};
  const cuuint32_t box[2] = {cuuint32_t(bw), cuuint32_t(bh / split)}, es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode rc %d\n", int(r)); return 3; }
  long long* cyc; int* sink;
  cudaMalloc(&cyc, ctas * 8); cudaMalloc(&sink, ctas * 4);
  const size_t smem = size_t(bw) * bh * 4 * depth;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  for (int rep = 0; rep < 2; ++rep) rate<<<ctas, 32, smem>>>(tm, bw, bh / split, split, depth, align, iters, size, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost);
  double mean = 0, mx = 0;
  for (long long v : h) { mean += double(v); mx = mx > double(v) ? mx : double(v); }
  mean /= ctas;
  printf("ctas %3d box %3dx%3d split %d depth %d align %2d: %8.0f cycles per box (slowest CTA %8.0f), %6.1f B/clk/SM\n", ctas, bw, bh, split,
         depth, align, mean / iters, mx / iters, double(bw) * bh * 4 * iters / mean);
  return 0;
}

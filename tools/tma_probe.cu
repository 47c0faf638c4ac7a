// Probe: 2D tensor-map TMA box loads of an int32 grid.  usage: tma_probe <mode> <box_w> <box_h>
//   mode 0: descriptor as __grid_constant__ parameter, 1: descriptor in global memory (+ acquire fence)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap pm, const void* gm, int use_global, int x, int y, int cells, int* out) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) uint64_t bar;
  int* buf = reinterpret_cast<int*>(dyn);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const void* tm = (use_global & 1) ? gm : (const void*)&pm;
    if (use_global == 2) {   // no copy at all
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
    } else if (use_global == 3) {   // plain 1D bulk copy of 256 B from the grid pointer handed in gm
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(256) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf)), "l"(gm), "r"(256), "r"(s32(&bar)) : "memory");
    } else {
    if (use_global) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(cells * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(buf)),
        "l"(tm), "r"(x), "r"(y), "r"(s32(&bar))
        : "memory");
    }
  }
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0, 0x989680;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(
          s32(&bar))
      : "memory");
  for (int i = threadIdx.x; i < cells; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, dt = argc > 4 ? atoi(argv[4]) : 0, bw = argc > 2 ? atoi(argv[2]) : 64, bh = argc > 3 ? atoi(argv[3]) : 64;
  const int sx = 500, sy = 400, pitch = 544, x0 = argc > 5 ? atoi(argv[5]) : 37, y0 = argc > 6 ? atoi(argv[6]) : 301;
  std::vector<int> h(size_t(pitch) * sy);
  for (size_t i = 0; i < h.size(); ++i) h[i] = int(i * 2654435761u >> 7);
  int* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encoder\n"); return 2; }
  alignas(64) CUtensorMap tm;
  const cuuint64_t dims[2] = {cuuint64_t(sx), cuuint64_t(sy)}; const cuuint64_t strides[1] = {cuuint64_t(pitch) * 4};
  const cuuint32_t box[2] = {cuuint32_t(bw), cuuint32_t(bh)}, es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&tm, dt == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, dt == 3 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d; desc:", int(r)); for (int i = 0; i < 16; ++i) printf(" %016llx", (unsigned long long)tm.opaque[i]); printf("\n");
  if (r) return 3;
  void* gtm; cudaMalloc(&gtm, 256); cudaMemcpy(gtm, &tm, 128, cudaMemcpyHostToDevice);
  const int cells = bw * bh;
  int* out; cudaMalloc(&out, cells * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, cells * 4);
  probe<<<1, 256, cells * 4>>>(tm, mode == 3 ? (void*)d : gtm, mode, x0, y0, cells, out);
  printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d box %dx%d: %s\n", mode, bw, bh, cudaGetErrorString(e));
  if (e) return 1;
  std::vector<int> o(cells); cudaMemcpy(o.data(), out, cells * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
    const int gx = x0 + xx, gy = y0 + yy;
    const int want = (gx < sx && gy < sy) ? h[size_t(gy) * pitch + gx] : 0;
    if (o[yy * bw + xx] != want) ++bad;
  }
  printf("mismatches %ld of %d\n", bad, cells);
  return bad ? 4 : 0;
}

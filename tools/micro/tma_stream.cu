// Micro-benchmark: how fast can TMA box loads stream a row-major fp32 matrix into shared memory?
// One producer thread + one consumer warp per CTA; the consumer only acknowledges stages.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t b, uint32_t ph) {
  uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory"); return ok; }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) { while (!mbar_try(b, ph)) {} }
__device__ __forceinline__ void tma2d(uint32_t dst, const void* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1) : "memory"); }

__global__ void stream_kernel(const __grid_constant__ CUtensorMap tm, int64_t tiles_total, int64_t tiles_per_block, int rows, int box_cols, int nboxes, int stages, unsigned* sink)
{
  extern __shared__ uint8_t smem[];
  uint32_t base = (s32(smem) + 1023u) & ~1023u;
  uint32_t stage_bytes = (uint32_t)rows * box_cols * 4 * nboxes;
  uint32_t bars = base + stages * stage_bytes;   // full[s], empty[s]
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bars + s * 8, 1); mbar_init(bars + (stages + s) * 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int64_t t0 = (int64_t)blockIdx.x * tiles_per_block, t1 = min(tiles_total, t0 + tiles_per_block);
  if (threadIdx.x == 0) {
    uint32_t s = 0, ph = 0;
    for (int64_t t = t0; t < t1; ++t) {
      mbar_wait(bars + (stages + s) * 8, ph ^ 1);
      mbar_expect(bars + s * 8, stage_bytes);
      for (int b = 0; b < nboxes; ++b) tma2d(base + s * stage_bytes + b * rows * box_cols * 4, &tm, b * box_cols, (int)(t * rows), bars + s * 8);
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x >= 32 && threadIdx.x < 64) {
    uint32_t s = 0, ph = 0; unsigned acc = 0;
    for (int64_t t = t0; t < t1; ++t) {
      mbar_wait(bars + s * 8, ph);
      unsigned v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(base + s * stage_bytes + (threadIdx.x - 32) * 4));
      acc += v;
      __syncwarp();
      if (threadIdx.x == 32) mbar_arrive(bars + (stages + s) * 8);
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
    if (acc == 0x12345678) sink[0] = acc;
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncFn enc = (EncFn)fp;
  const size_t bytes = 4ull << 30;   // 4 GiB
  float* X; CK(cudaMalloc(&X, bytes)); CK(cudaMemset(X, 1, bytes));
  unsigned* sink; CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("d  box_cols nboxes rows stages ctas/sm  inflight_KB/SM   GB/s\n");
  for (int d : {16, 32, 64, 128}) {
    int64_t n = bytes / (d * 4);
    for (int rows : {32, 64, 128, 256}) {
      for (int stages : {2, 4, 8}) {
        for (int cps : {1, 2, 4}) {
          int box_cols = d < 32 ? d : 32; int nboxes = d / box_cols;
          size_t stage_bytes = (size_t)rows * d * 4;
          size_t smem = stages * stage_bytes + stages * 16 + 1024 + 64;
          if (smem > 227 * 1024 / cps - 1024) continue;
          CUtensorMap tm; cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)n}; cuuint64_t str[1] = {(cuuint64_t)d * 4};
          cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
          CUtensorMapSwizzle sw = box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
          if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
          int64_t tiles = n / rows; int grid = 148 * cps; int64_t tpb = (tiles + grid - 1) / grid;
          float best = 1e9;
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            stream_kernel<<<grid, 64, smem>>>(tm, tiles, tpb, rows, box_cols, nboxes, stages, sink);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
          }
          CK(cudaGetLastError());
          printf("%3d %5d %5d %5d %5d %5d %10.0f %10.0f\n", d, box_cols, nboxes, rows, stages, cps, cps * stages * stage_bytes / 1024.0, bytes / (best * 1e-3) / 1e9);
        }
      }
    }
  }
  return 0;
}

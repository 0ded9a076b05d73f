// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, SS operands) as a function of N,
// kind (tf32 K = 8 / bf16 K = 16) and the mix the E-step issues per 32-feature K-block (4 bf16 + 4 tf32).
// It settles the MMA-issue floor of fused_l2_argmin_2cta_kernel: 17 MMAs per 256-row tile at C3.
// Uses the library's own PTX wrappers (cuml_b200/csrc/ptx.cuh).  Operand contents are ones; timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../cuml_b200/csrc -o mma_rate_pair mma_rate_pair.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "ptx.cuh"

using namespace cb2;

// mode 0: tf32 only, 1: bf16 only, 2: the E-step mix (4 bf16 K=16 then 4 tf32 K=8 per K-block)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_rate(int N, int nacc, int mode, int iters, long long* out)
{
  extern __shared__ uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t raw  = ptx::smem_u32(smem);
  const uint32_t base = (raw + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem + (base - raw))[i] = 0x3f803f80u;   // 1.0 as bf16 pairs; a small fp32 as tf32
  const bool leader = ptx::cluster_ctarank() == 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc_2cta(ptx::smem_u32(&tbase), 512);
    ptx::tmem_relinquish_2cta();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tm = tbase;
  if (threadIdx.x < 32) {
    const uint32_t id32 = ptx::umma_idesc_tf32(256, N);
    const uint32_t id16 = ptx::umma_idesc_bf16(256, N);
    // A: this CTA's 128 rows (16 KB, 128B swizzle; 8 KB tiles with 64B swizzle for bf16); B: N/2 rows per CTA
    const uint64_t a32 = ptx::umma_desc_sw128(base), b32 = ptx::umma_desc_sw128(base + 16384);
    const uint64_t a16 = ptx::umma_desc_sw64(base + 32768), b16 = ptx::umma_desc_sw64(base + 32768 + 8192);
    const long long t0 = clock64();
    if (leader) {
      if (ptx::elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t d = tm + static_cast<uint32_t>((i % nacc) * N);
          if (mode == 0) {
            ptx::mma_tf32_ss_2cta(d, a32 + static_cast<uint64_t>((i & 3) * 2), b32 + static_cast<uint64_t>((i & 3) * 2), id32, 1u);
          } else if (mode == 1) {
            ptx::mma_f16_ss_2cta(d, a16 + static_cast<uint64_t>((i & 1) * 2), b16 + static_cast<uint64_t>((i & 1) * 2), id16, 1u);
          } else {
            const int j = i & 7;   // 0..3 bf16, 4..7 tf32
            if (j < 4) ptx::mma_f16_ss_2cta(d, a16 + static_cast<uint64_t>((j & 1) * 2), b16 + static_cast<uint64_t>((j & 1) * 2), id16, 1u);
            else ptx::mma_tf32_ss_2cta(d, a32 + static_cast<uint64_t>((j & 3) * 2), b32 + static_cast<uint64_t>((j & 3) * 2), id32, 1u);
          }
        }
        ptx::mma_commit_2cta(ptx::smem_u32(&bar), 3);   // multicast: both CTAs' barriers complete
      }
      __syncwarp();
    }
    ptx::mbar_wait(ptx::smem_u32(&bar), 0u);
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  if (threadIdx.x < 32) ptx::tmem_dealloc_2cta(tm, 512);
}

int main()
{
  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(pair_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  printf("mode  N  nacc  grid  cycles/mma  TFLOP/s(148 SMs at 1.9 GHz)\n");
  const char* names[3] = {"tf32", "bf16", "mix "};
  for (int mode : {0, 1, 2})
    for (int N : {64, 128, 256})
      for (int nacc : {1, 2})
        for (int grid : {2, 148}) {
          if (nacc * N > 512) continue;
          const int iters = 4096;
          long long h = 0;
          for (int rep = 0; rep < 2; ++rep) {
            pair_rate<<<grid, 128, 66 * 1024>>>(N, nacc, mode, iters, out);
            cudaDeviceSynchronize();
          }
          cudaError_t e = cudaGetLastError();
          if (e != cudaSuccess) {
            printf("err %s\n", cudaGetErrorString(e));
            return 1;
          }
          cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
          const double cyc = static_cast<double>(h) / iters;
          const double kavg = mode == 0 ? 8.0 : (mode == 1 ? 16.0 : 12.0);
          // per SM: 128 rows x N columns x K per MMA
          printf("%s %4d %3d %4d %9.1f %10.1f\n", names[mode], N, nacc, grid, cyc, 2.0 * 128 * N * kavg / cyc * 1.9e9 * 148 / 1e12);
        }
  return 0;
}

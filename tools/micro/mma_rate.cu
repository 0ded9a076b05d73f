// Micro-benchmark: cycles per tcgen05.mma (kind::tf32, SS operands, M=128, cta_group::1) as a function of N
// and of how many TMEM accumulators consecutive MMAs rotate over (1 = fully dependent chain).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d; }
__device__ __forceinline__ uint64_t desc_sw64(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(512 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)4 << 61; return d; }
__device__ __forceinline__ bool elect_one() { uint32_t p; asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p)); return p != 0; }

__global__ void __launch_bounds__(128, 1) k(int N, int nacc, int kind_bf16, int iters, long long* out, int ts)
{
  extern __shared__ uint8_t smem[];
  __shared__ uint64_t bar; __shared__ uint32_t tbase;
  uint32_t base = (s32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)(smem + (base - s32(smem))))[i] = 0x3f800000u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(512) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = tbase;
  if (threadIdx.x < 32) {
    uint32_t idesc = (1u << 4) | ((kind_bf16 ? 1u : 2u) << 7) | ((kind_bf16 ? 1u : 2u) << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint64_t da = desc_sw128(base), db = desc_sw128(base + 16384);
    if (ts == 2) { da = desc_sw64(base); db = desc_sw64(base + 16384); }
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        uint32_t d = tm + (uint32_t)((i % nacc) * N);
        uint64_t adv = (uint64_t)((i & 3) * 2);
        if (ts == 2) adv = (uint64_t)((i & 1) * 2);
        if (ts == 1) {
          // A operand from TMEM (columns 448.. hold garbage; timing only)
          uint32_t at = tm + 448 + (uint32_t)((i & 3) * 8);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(at), "l"(db + adv), "r"(idesc), "r"(1u) : "memory");
        } else if (kind_bf16) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da + adv), "l"(db + adv), "r"(idesc), "r"(1u) : "memory");
        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da + adv), "l"(db + adv), "r"(idesc), "r"(1u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main()
{
  long long* out; cudaMalloc(&out, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("kind  N  nacc  grid  cycles/mma  TFLOP/s(all SMs, if grid=148)\n");
  for (int bf : {0, 1}) for (int N : {64, 128, 256}) for (int nacc : {1, 2, 4}) for (int grid : {1, 148}) {
    if (nacc * N > 512) continue;
    int iters = 4096; long long h = 0;
    for (int rep = 0; rep < 2; ++rep) { k<<<grid, 128, 50 * 1024>>>(N, nacc, bf, iters, out, 0); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h / iters; int K = bf ? 16 : 8;
    printf("%s %4d %3d %4d %9.1f %10.1f\n", bf ? "bf16" : "tf32", N, nacc, grid, cyc, 2.0 * 128 * N * K / cyc * 1.9e9 * 148 / 1e12);
  }
  printf("bf16 with 64-byte swizzled operand rows (SW64)\n");
  for (int N : {128, 256}) {
    int iters = 4096; long long h = 0;
    for (int rep = 0; rep < 2; ++rep) { k<<<148, 128, 50 * 1024>>>(N, 1, 1, iters, out, 2); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h / iters;
    printf("bf16-SW64 %4d %9.1f %10.1f\n", N, cyc, 2.0 * 128 * N * 16 / cyc * 1.9e9 * 148 / 1e12);
  }
  printf("TS mode (A in TMEM), tf32\n");
  for (int N : {64, 128, 256}) for (int nacc : {1}) {
    int iters = 4096; long long h = 0;
    for (int rep = 0; rep < 2; ++rep) { k<<<148, 128, 50 * 1024>>>(N, nacc, 0, iters, out, 1); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h / iters;
    printf("tf32-TS %4d %9.1f %10.1f\n", N, cyc, 2.0 * 128 * N * 8 / cyc * 1.9e9 * 148 / 1e12);
  }
  return 0;
}

// Micro-benchmark: issue-to-issue cost of ONE tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N, of the
// shared-memory operand layout (no-swizzle K-major planes as the kernels in csrc/ use them, or 128-byte swizzle), of where A comes
// from (shared memory or TMEM) and of how many accumulators the MMAs rotate over.  One CTA per SM, one issuing thread, 2048 MMAs
// back to back, one commit at the end.  Answers "why do the N = 16..160 kernels sit at 25 % tensor-pipe busy".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I shift-net_b200/csrc -o scripts/_bin/ubench_umma scripts/ubench_umma.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace gsn;

constexpr int ITERS = 2048;
constexpr int SMEM = 160 * 1024;

// layout: 0 = no swizzle (LBO = plane pitch, SBO = 128), 1 = SWIZZLE_128B (SBO = 1024)
// a_src : 0 = A from shared memory, 1 = A from TMEM (last 8 columns)
// nacc  : accumulators the MMAs rotate over (1 = one dependent chain)
// walk  : 1 = the A start address moves by 16 rows per MMA (tap-offset style), 0 = same operand every time
__global__ void __launch_bounds__(128) k_umma(long long *cyc, int N, int layout, int a_src, int nacc, int walk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < SMEM / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (threadIdx.x == 0) mbar_init(bar_a, 1);
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    // A: 128 rows, up to 64 K elements ; B right behind it
    const uint32_t a_off = 0, b_off = 64 * 1024;
    uint64_t ad, bd;
    if (layout == 0) {
      ad = make_smem_desc(sbase + a_off, 512 * 16, 128);      // planes of 512 rows (room for the walk)
      bd = make_smem_desc(sbase + b_off, 256 * 16, 128);
    } else {
      ad = make_smem_desc(sbase + a_off, 16, 1024) | ((uint64_t)2 << 61);
      bd = make_smem_desc(sbase + b_off, 16, 1024) | ((uint64_t)2 << 61);
    }
    const uint32_t step16 = layout == 0 ? 16 : 128;   // 16 rows further, in 16-byte units
    const uint32_t spacing = N <= 64 ? 64u : (N <= 128 ? 128u : 256u);   // disjoint accumulators
    if (nacc * spacing > 512u) nacc = 512 / spacing;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t d = tmem + (uint32_t)(j & (nacc - 1)) * spacing;
        const uint64_t a = ad + (uint64_t)(walk ? (uint32_t)(j * step16) : 0u);
        if (a_src == 0) {
          umma_f16(d, a, bd, idesc, 1u);
        } else {
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
              "}\n" ::"r"(d), "r"(tmem + 504u), "l"(bd), "r"(idesc), "r"(1u)
              : "memory");
        }
      }
    }
    umma_commit(bar_a);
    mbar_wait(bar_a, 0);
    const long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
}

int main() {
  long long *cyc;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int Ns[] = {8, 16, 32, 48, 64, 96, 128, 160, 192, 256};
  printf("tcgen05.mma cta_group::1 kind::f16 M=128 K=16: cycles per MMA (2048 back to back, 148 CTAs, median SM)\n");
  printf("%-34s", "config \\ N");
  for (int n : Ns) printf("%7d", n);
  printf("\n%-34s", "math floor 128*N/256");
  for (int n : Ns) printf("%7d", n / 2);
  printf("\n");
  struct Cfg { const char *name; int layout, a_src, nacc, walk; };
  const Cfg cfgs[] = {
      {"SS no-swizzle, 1 acc", 0, 0, 1, 0},       {"SS no-swizzle, 4 acc", 0, 0, 4, 0},
      {"SS no-swizzle, 4 acc, A walks", 0, 0, 4, 1}, {"SS swizzle-128B, 1 acc", 1, 0, 1, 0},
      {"SS swizzle-128B, 4 acc", 1, 0, 4, 0},      {"SS swizzle-128B, 4 acc, A walks", 1, 0, 4, 1},
      {"TS (A in TMEM) no-swizzle B, 1 acc", 0, 1, 1, 0}, {"TS (A in TMEM) swizzle B, 1 acc", 1, 1, 1, 0},
  };
  for (const Cfg &c : cfgs) {
    printf("%-34s", c.name);
    for (int n : Ns) {
      long long h[148];
      for (int rep = 0; rep < 2; ++rep) {
        k_umma<<<148, 128, SMEM>>>(cyc, n, c.layout, c.a_src, c.nacc, c.walk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  ERR(%s)", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      for (int i = 0; i < 148; ++i) for (int j = i + 1; j < 148; ++j) if (h[j] < h[i]) { long long t = h[i]; h[i] = h[j]; h[j] = t; }
      printf("%7.1f", (double)h[74] / ITERS);
    }
    printf("\n");
  }
  return 0;
}

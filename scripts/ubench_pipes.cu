// Micro-benchmark of the CUDA-core pipes the fused pass A leans on (sm_100a): warp-instructions per clock per SM for
//   HFMA2 (3 register operands), FFMA, HFMA2 interleaved with LDS.128, and LDS.128 alone,
// at the occupancy of the pass-A kernel (16 warps per SM, 1 CTA per SM) and at 32 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/ubench_pipes scripts/ubench_pipes.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

template <int ILP>
__global__ void k_hfma2(unsigned *out, long long *cyc, unsigned seed) {
  unsigned a[ILP], w0 = seed | 0x3c003c00u, w1 = seed ^ 0x38003800u;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(w0), "r"(w1));
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void k_ffma(float *out, long long *cyc, float seed) {
  float a[ILP], w0 = seed, w1 = seed * 0.5f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(w0), "f"(w1));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// per iteration: NL LDS.128 (conflict-free, lane-consecutive 16 B) + NF HFMA2 consuming the loaded values
template <int NL, int NF>
__global__ void k_mix(unsigned *out, long long *cyc, unsigned seed) {
  extern __shared__ uint4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
  unsigned acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = i;
  unsigned w = seed | 0x3c003c00u;
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    uint4 v[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) v[l] = sm[(idx + l * 512) & 4095];
    idx = (idx + 32) & 4095;
    if (NF > 0) {
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const uint4 &vv = v[f % NL];
        unsigned x = (f & 3) == 0 ? vv.x : (f & 3) == 1 ? vv.y : (f & 3) == 2 ? vv.z : vv.w;
        asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(acc[f & 7]) : "r"(x), "r"(w));
      }
    } else {
#pragma unroll
      for (int l = 0; l < NL; ++l) acc[l & 7] ^= v[l].x ^ v[l].y ^ v[l].z ^ v[l].w;
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mma.sync m16n8k16 f16 -> f32, ILP independent accumulator sets per warp
template <int ILP>
__global__ void k_hmma(float *out, long long *cyc, unsigned seed) {
  float acc[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  unsigned a0 = seed | 0x3c003c00u, a1 = a0 + threadIdx.x, a2 = a0 ^ 0x100u, a3 = a1 ^ 0x1u, b0 = 0x38003800u, b1 = 0x34003400u;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ldmatrix.x4 (conflict-free rows at a 48-byte pitch) feeding 2 or 6 HMMAs (the dense-conv inner loop shapes)
template <int NMMA>
__global__ void k_ldm_hmma(float *out, long long *cyc, unsigned seed) {
  extern __shared__ uint4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_uint4(0x3c003c00u, 0x38003800u, 0x34003400u, 0x30003000u);
  float acc[NMMA][4];
#pragma unroll
  for (int i = 0; i < NMMA; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  unsigned b0 = 0x38003800u | seed, b1 = 0x34003400u;
  __syncthreads();
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + ((threadIdx.x & 15) * 48 + (threadIdx.x >> 4 & 1) * 16) + (threadIdx.x >> 5) * 1536;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    unsigned a0, a1, a2, a3;
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(base + (it & 7) * 96));
#pragma unroll
    for (int i = 0; i < NMMA; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NMMA; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// MUFU: tanh.approx.f32 (one element per op) vs tanh.approx.f16x2 (two) -- the sigmoid gate of pass A is MUFU-bound
template <int ILP, bool PACKED>
__global__ void k_tanh(unsigned *out, long long *cyc, unsigned seed) {
  unsigned a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 0x34003400u + threadIdx.x + i + seed;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (PACKED) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
      else asm volatile("tanh.approx.f32 %0, %0;" : "+f"(*reinterpret_cast<float *>(&a[i])));
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// do HFMA2 and FFMA share one pipe?  NH HFMA2 + NF FFMA per iteration, independent chains
template <int NH, int NF>
__global__ void k_hf_mix(unsigned *out, long long *cyc, unsigned seed) {
  unsigned h[NH > 0 ? NH : 1], w0 = seed | 0x3c003c00u, w1 = seed ^ 0x38003800u;
  float f[NF > 0 ? NF : 1], g0 = 1.0001f, g1 = 0.5f;
#pragma unroll
  for (int i = 0; i < NH; ++i) h[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < (NH > NF ? NH : NF); ++i) {
      if (i < NH) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(w0), "r"(w1));
      if (i < NF) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g0), "f"(g1));
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < NH; ++i) s ^= h[i];
#pragma unroll
  for (int i = 0; i < NF; ++i) s ^= __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char *name, F launch, int threads, double winstr_per_thread_iter) {
  unsigned *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  launch(out, cyc);
  launch(out, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double warps = threads / 32.0;
  printf("%-34s threads=%4d  cycles=%9.0f  warp-instr/clk/SM = %.3f   (%s)\n", name, threads, avg,
         warps * ITERS * winstr_per_thread_iter / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  cudaFuncSetAttribute(k_mix<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k_mix<3, 36>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k_mix<3, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k_mix<2, 36>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k_ldm_hmma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k_ldm_hmma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int threads : {256, 512, 1024}) {
    run("HMMA m16n8k16 ILP8 (x2048 MAC)", [&](unsigned *o, long long *c) { k_hmma<8><<<148, threads>>>((float *)o, c, 1); }, threads, 8);
    run("HMMA m16n8k16 ILP16", [&](unsigned *o, long long *c) { k_hmma<16><<<148, threads>>>((float *)o, c, 1); }, threads, 16);
    run("ldmatrix.x4 + 2 HMMA", [&](unsigned *o, long long *c) { k_ldm_hmma<2><<<148, threads, 65536>>>((float *)o, c, 0); }, threads, 3);
    run("ldmatrix.x4 + 6 HMMA", [&](unsigned *o, long long *c) { k_ldm_hmma<6><<<148, threads, 65536>>>((float *)o, c, 0); }, threads, 7);
    run("HFMA2 ILP8", [&](unsigned *o, long long *c) { k_hfma2<8><<<148, threads>>>(o, c, 1); }, threads, 8);
    run("HFMA2 ILP16", [&](unsigned *o, long long *c) { k_hfma2<16><<<148, threads>>>(o, c, 1); }, threads, 16);
    run("FFMA ILP8", [&](unsigned *o, long long *c) { k_ffma<8><<<148, threads>>>((float *)o, c, 1.f); }, threads, 8);
    run("FFMA ILP16", [&](unsigned *o, long long *c) { k_ffma<16><<<148, threads>>>((float *)o, c, 1.f); }, threads, 16);
    run("MUFU tanh.approx.f32 ILP8", [&](unsigned *o, long long *c) { k_tanh<8, false><<<148, threads>>>(o, c, 1); }, threads, 8);
    run("MUFU tanh.approx.f16x2 ILP8", [&](unsigned *o, long long *c) { k_tanh<8, true><<<148, threads>>>(o, c, 1); }, threads, 8);
    run("8 HFMA2 + 8 FFMA (co-issue?)", [&](unsigned *o, long long *c) { k_hf_mix<8, 8><<<148, threads>>>(o, c, 1); }, threads, 16);
    run("8 HFMA2 + 4 FFMA", [&](unsigned *o, long long *c) { k_hf_mix<8, 4><<<148, threads>>>(o, c, 1); }, threads, 12);
    run("LDS.128 x4 only (xor)", [&](unsigned *o, long long *c) { k_mix<4, 0><<<148, threads, 65536>>>(o, c, 1); }, threads, 4);
    run("LDS.128 x3 + 36 HFMA2 (dw3x3 mix)", [&](unsigned *o, long long *c) { k_mix<3, 36><<<148, threads, 65536>>>(o, c, 1); }, threads, 39);
    run("LDS.128 x3 + 12 HFMA2", [&](unsigned *o, long long *c) { k_mix<3, 12><<<148, threads, 65536>>>(o, c, 1); }, threads, 15);
    run("LDS.128 x2 + 36 HFMA2", [&](unsigned *o, long long *c) { k_mix<2, 36><<<148, threads, 65536>>>(o, c, 1); }, threads, 38);
  }
  return 0;
}

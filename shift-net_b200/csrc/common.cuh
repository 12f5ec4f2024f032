// Shared device/host helpers for the shiftnet_b200 kernels (sm_100a).
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/shiftnet_b200.h"

namespace gsn {

// ---- host-side error plumbing -------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch();
int check_launch(const char *what);   // cudaGetLastError -> GSN_E_CUDA + message
// 4-D fp16 NHWC tensor map (dims innermost first: C, W, H, T) with a (bc, bw, bh, 1) box, zero OOB fill. false = unavailable.
// swizzle128: CU_TENSOR_MAP_SWIZZLE_128B (box rows of exactly 128 bytes; 16-byte chunk c of box row r sits at chunk c ^ (r & 7)).
bool encode_tmap_nhwc(CUtensorMap *tm, const void *base, int C, int W, int H, int T, int bc, int bw, int bh, bool swizzle128 = false);
// 4-D fp16 tensor map over a k-chunk planar tensor [T][KC][H][W][8] (dims innermost first: W*8, H, KC, T) with a
// (bw*8, bh, KC, 1) box: the box lands in shared memory as KC planes of bh*bw 16-byte pixel vectors = the no-swizzle
// K-major UMMA operand layout.  Zero OOB fill.
bool encode_tmap_planar(CUtensorMap *tm, const void *base, int W, int H, int KC, int T, int bw, int bh);

// Kernel function attributes (cudaFuncSetAttribute) and the SM count are per DEVICE, not per process: one-time setup is
// keyed by the current device.  The flag is published only after the setup ran, so a second host thread either repeats the
// (idempotent) setup or sees it complete -- it never launches ahead of it.
int current_device();
int sm_count();                                    // multiprocessors of the current device (cached per device)
bool device_needs_setup(unsigned long long *mask); // atomically reads the current device's bit
void device_setup_done(unsigned long long *mask);
#define GSN_ONCE_PER_DEVICE(...)                          \
  do {                                                    \
    static unsigned long long once_mask_ = 0;             \
    if (gsn::device_needs_setup(&once_mask_)) {           \
      __VA_ARGS__;                                        \
      gsn::device_setup_done(&once_mask_);                \
    }                                                     \
  } while (0)

#define GSN_REQUIRE(cond, ...)             \
  do {                                     \
    if (!(cond)) {                         \
      gsn::set_error(__VA_ARGS__);         \
      return GSN_E_BADARG;                 \
    }                                      \
  } while (0)

// ---- small device helpers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid) {
  // 16-byte async copy; src-size 0 => zero fill (used for image borders / out-of-range frames)
  uint32_t d = smem_u32(smem_dst);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2 *>(&v));
}
__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
  float2 a = unpack_half2(v.x), b = unpack_half2(v.y), c = unpack_half2(v.z), d = unpack_half2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_half2(f[0], f[1]); v.y = pack_half2(f[2], f[3]);
  v.z = pack_half2(f[4], f[5]); v.w = pack_half2(f[6], f[7]);
  return v;
}
__device__ __forceinline__ float sigmoidf_fast(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace gsn

// Second 1x1 (C -> 2C) + SimpleGate2 + per-tile channel sums of the Ours+ gated blocks (C = 80; gshift_deblur1.py:186-258:
// body.4 then z = a * sigmoid(b), and the spatial sums CALayer2 pools) as a streaming TMA + tcgen05 kernel (sm_100a):
// [pixels x 80] . [80 x 160], the second genuinely dense GEMM of the block.
//
// Same structure as ln_pw_tc.cu: persistent CTAs, 128-pixel tiles, 3-stage ring of k-chunk planes landed by TMA boxes
// {8 channels, 128 pixels}, warp 1 issues 5 x tcgen05.mma (M=128, N=160, K=16) per tile into one of two TMEM accumulators,
// warps 2..5 run the gate on the accumulators (thread = pixel).  The weights are held at HALF scale (exact in fp16), so that
// a * sigmoid(b) = (a/2) * tanh(b/2) + a/2 is one MUFU and one FFMA per element.  The per-tile sums of z (fp32, in-image pixels
// only, fixed order: butterfly over the 32 pixels of a warp, then the four warps) go to chan_partial[t][tile][C] with the same
// 128-pixel tiling as the mma.sync kernel it replaces (gsn_cab_tiles_linear).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace gsn {

constexpr int kPgThreads = 192;

struct PgCfg {
  static constexpr int C = 80, N = 160, MP = 128, NST = 3, KC = 10;
  static constexpr int PLANE = MP * 16;
  static constexpr int STAGE = KC * PLANE;                   // 20 KB
  static constexpr int W_BYTES = KC * N * 16;                // [10 planes][160][8] fp16 = 25.6 KB
  static constexpr int S_W = NST * STAGE;
  static constexpr int S_RED = S_W + W_BYTES;                // [4 warps][80] fp32
  static constexpr int S_BAR = S_RED + 4 * C * 4;
  static constexpr int SMEM = S_BAR + 128;
};

__global__ void __launch_bounds__(kPgThreads, 1) pw_gate_tc_kernel(const __grid_constant__ CUtensorMap tm_u, const unsigned char *__restrict__ w2half,
                                                                 __half *__restrict__ z, float *__restrict__ chan_partial, int T, long long hw) {
  using K = PgCfg;
  constexpr int C = K::C;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_f = (int)((hw + K::MP - 1) / K::MP), total = tiles_f * T;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + K::S_BAR;
  auto full = [&](int s) { return bar0 + 8 * s; };
  auto empty = [&](int s) { return bar0 + 8 * (K::NST + s); };
  auto tmem_full = [&](int a) { return bar0 + 8 * (2 * K::NST + a); };
  auto tmem_empty = [&](int a) { return bar0 + 8 * (2 * K::NST + 2 + a); };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + K::S_BAR + 8 * (2 * K::NST + 4));

  if (tid == 0) {
    for (int s = 0; s < K::NST; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);           // tcgen05.commit: the MMAs of the tile have read the stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full(a), 1);
      mbar_init(tmem_empty(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  for (int i = tid; i < K::W_BYTES / 16; i += kPgThreads) cp_async16(smem + K::S_W + i * 16, w2half + (size_t)i * 16, true);
  cp_async_commit();
  cp_async_wait<0>();
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int s = i % K::NST, ph = (i / K::NST) & 1;
        const int t = tile / tiles_f, p0 = (tile - t * tiles_f) * K::MP;
        mbar_wait(empty(s), ph ^ 1);
        const uint32_t st = sbase + s * K::STAGE, fb = full(s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(K::STAGE) : "memory");
        for (int c = 0; c < K::KC; ++c)
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                  "r"(st + c * K::PLANE), "l"(reinterpret_cast<uint64_t>(&tm_u)), "r"(c * 8), "r"(p0), "r"(t), "r"(fb)
              : "memory");
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, K::N);
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int s = i % K::NST, ph = (i / K::NST) & 1, acc = i & 1, aph = (i >> 1) & 1;
        mbar_wait(tmem_empty(acc), aph ^ 1);
        mbar_wait(full(s), ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < K::KC / 2; ++k) {
          const uint64_t ad = smem_desc_at(sbase >> 4, s * K::STAGE + 2 * k * K::PLANE, K::PLANE, 128);
          const uint64_t bd = smem_desc_at(sbase >> 4, K::S_W + 2 * k * (K::N * 16), K::N * 16, 128);
          umma_f16(tmem + acc * 256, ad, bd, idesc, k > 0);
        }
        umma_commit(tmem_full(acc));
        umma_commit(empty(s));
      }
    }
  } else {
    // ---- gate warps (2..5): TMEM lane quarter = warp % 4, thread = pixel --------------------------------------------------
    const int q = warp & 3, r = q * 32 + lane, et = tid - 64;
    float *red = reinterpret_cast<float *>(smem + K::S_RED);
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int acc = i & 1, aph = (i >> 1) & 1;
      const int t = tile / tiles_f, tf = tile - t * tiles_f, p0 = tf * K::MP;
      const long long pixel = (long long)p0 + r;
      const bool valid = pixel < hw;
      mbar_wait(tmem_full(acc), aph);
      tc_fence_after();
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + acc * 256;
      __half *zp = z + ((size_t)t * hw + (valid ? pixel : 0)) * C;
      float csum[5];
#pragma unroll
      for (int cg = 0; cg < 5; ++cg) {             // 16 channels at a time: a = columns 16 cg.., b = columns 80 + 16 cg..
        uint32_t a[16], b[16];
        tmem_ld16_nowait(ta + cg * 16, a);
        tmem_ld16_nowait(ta + C + cg * 16, b);
        tmem_ld_wait();
        if (cg == 4) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(acc));
        }
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float th;
          asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(__uint_as_float(b[e])));
          v[e] = fmaf(__uint_as_float(a[e]), th, __uint_as_float(a[e]));      // half-scale weights: (a/2) tanh(b/2) + a/2
        }
        if (valid) {
          *reinterpret_cast<uint4 *>(zp + cg * 16) = pack8(*reinterpret_cast<float(*)[8]>(&v[0]));
          *reinterpret_cast<uint4 *>(zp + cg * 16 + 8) = pack8(*reinterpret_cast<float(*)[8]>(&v[8]));
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0.f;
        }
        // butterfly over the warp's 32 pixels: lane l ends with the sum of channel 16 cg + (l >> 1)
#pragma unroll
        for (int off = 16, n = 8; off >= 2; off >>= 1, n >>= 1) {
          const bool hi = lane & off;
#pragma unroll
          for (int e = 0; e < n; ++e) {
            const float send = hi ? v[e] : v[e + n], keep = hi ? v[e + n] : v[e];
            v[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        csum[cg] = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
      }
      asm volatile("bar.sync 1, 128;\n" ::: "memory");             // the previous tile's reduction has read `red`
      if (!(lane & 1)) {
#pragma unroll
        for (int cg = 0; cg < 5; ++cg) red[q * C + cg * 16 + (lane >> 1)] = csum[cg];
      }
      asm volatile("bar.sync 2, 128;\n" ::: "memory");
      if (et < C) chan_partial[((size_t)t * tiles_f + tf) * C + et] = (red[et] + red[C + et]) + (red[2 * C + et] + red[3 * C + et]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
  }
}

bool encode_tmap_chunk128(CUtensorMap *tm, const void *base, int C, long long hw, int T);   // ln_pw_tc.cu

}  // namespace gsn

extern "C" int gsn_pw_gate_tc(const void *u, const void *w2half, void *z, float *chan_partial, int T, int H, int W, int C, void *stream) {
  using namespace gsn;
  using K = PgCfg;
  GSN_REQUIRE(u && w2half && z && chan_partial, "pw_gate_tc: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "pw_gate_tc: empty shape");
  if (C != 80) { set_error("pw_gate_tc: C=%d unsupported (80)", C); return GSN_E_UNSUPPORTED; }
  const long long hw = (long long)H * W;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (!encode_tmap_chunk128(&tm, u, C, hw, T)) {
    set_error("pw_gate_tc: cuTensorMapEncodeTiled failed (H*W=%lld T=%d)", hw, T);
    return GSN_E_CUDA;
  }
  GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(pw_gate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  const long long total = (hw + K::MP - 1) / K::MP * T;
  const int sms = sm_count();
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  pw_gate_tc_kernel<<<grid, kPgThreads, K::SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      tm, reinterpret_cast<const unsigned char *>(w2half), reinterpret_cast<__half *>(z), chan_partial, T, hw);
  count_launch();
  return check_launch("pw_gate_tc");
}

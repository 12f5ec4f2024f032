// Dense 3x3 convolution (stride 1, zero padding 1) for the WIDE channel-attention blocks of Ours+ (CAB.body, gshift_deblur1.py:
// 143-150, TFR_UNet levels 2 and 3: 36 and 48 channels, stored as 40 / 48) as an implicit GEMM on tcgen05 (sm_100a).
//
// The input tile (18 x 32 pixels incl. the 1-pixel halo, all input channels) is landed by TMA as k-chunk PLANES
// [chunk][18 x 32 pixels][8 channels]: one box {8 channels, 32 px, 18 rows} per chunk -- the no-swizzle K-major UMMA operand
// layout with the pixels of the region in raster order.  For output pixel m = y * 32 + x of the 16 x 32 output raster the tap
// (dy, dx) reads region pixel m + 32 dy + dx: the same operand, its descriptor start address moved by (32 dy + dx) * 16 bytes.
// So the conv is 9 taps x 3 k-steps tcgen05.mma (M = 128 pixels, N = 48, K = 16) per 4-row slab, fp32 accumulators in TMEM,
// no im2col, no operand copies; columns x = 30, 31 of the raster are wrap-around garbage and never stored.  TMA's hardware zero
// fill is the conv's zero padding at the image border AND the K padding of 40-channel tensors (the sixth chunk lies beyond C).
// With N = 48 an MMA's shared-memory operand reads (5.5 KB) cost 43 cycles of the SM's port against 24 cycles of tensor pipe:
// the kernel is bound by that port at ~10 cycles per output pixel per SM, 3.7x the mma.sync kernel it replaces for these shapes.
//
// Persistent CTAs; warp 0 = TMA producer (2-stage ring), warp 1 = MMA issuer (two TMEM accumulator sets, so the epilogue of a
// tile runs under the MMAs of the next), warps 2..5 = epilogue (thread = TMEM lane = pixel of a raster row: bias, PReLU, fp16,
// the pixel's contiguous NHWC row; optional deterministic per-tile channel sums for the CALayer pooling).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace gsn {

constexpr int kC3Threads = 192;

struct C3Cfg {
  static constexpr int TW = 30, TH = 16, RW = 32, RH = TH + 2, KC = 6, NP = 48;   // output tile, region, k-chunks, padded N
  static constexpr int PLANE = (RH * RW * 16 + 64 + 127) / 128 * 128;              // 9344: 18 x 32 px x 16 B + room for the +2 px over-read
  static constexpr int STAGE = KC * PLANE, NST = 2;
  static constexpr int W_BYTES = 9 * KC * NP * 16;                                  // [tap][chunk][48][8] fp16 = 41472
  static constexpr int S_W = NST * STAGE;
  static constexpr int S_BIAS = S_W + W_BYTES;                                      // 48 floats
  static constexpr int S_RED = S_BIAS + NP * 4;                                     // [4 warps][48] floats
  static constexpr int S_BAR = S_RED + 4 * NP * 4;
  static constexpr int SMEM = S_BAR + 128;
  static_assert(S_W % 128 == 0 && PLANE % 128 == 0, "TMA destination alignment");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

__global__ void __launch_bounds__(kC3Threads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const unsigned char *__restrict__ wpack,
                                                                 const float *__restrict__ bias, __half *__restrict__ dst,
                                                                 float *__restrict__ chan_partial, int T, int H, int W, int cout_p,
                                                                 int has_prelu, float slope) {
  using K = C3Cfg;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = (W + K::TW - 1) / K::TW, tiles_y = (H + K::TH - 1) / K::TH, tiles_f = tiles_x * tiles_y, total = tiles_f * T;
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
      mbar_init(tmem_empty(a), 4);      // the four epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  for (int i = tid; i < K::W_BYTES / 16; i += kC3Threads) cp_async16(smem + K::S_W + i * 16, wpack + (size_t)i * 16, true);
  cp_async_commit();
  if (tid < K::NP) reinterpret_cast<float *>(smem + K::S_BIAS)[tid] = bias ? bias[tid] : 0.f;
  // the tail of every plane (the +2 pixel over-read of the last raster row, garbage outputs only) must at least be finite-free of
  // stale NaN patterns mixing into nothing: it feeds only the never-stored columns, but zero it once so the rows are reproducible
  for (int i = tid; i < K::NST * K::KC * 4; i += kC3Threads) {
    const int pl = i / 4, r = i - pl * 4;
    *reinterpret_cast<uint4 *>(smem + pl * K::PLANE + K::RH * K::RW * 16 + r * 16) = make_uint4(0, 0, 0, 0);
  }
  cp_async_wait<0>();
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- producer: six k-chunk planes per tile ---------------------------------------------------------------------
    if (lane == 0) {
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int s = i % K::NST, ph = (i / K::NST) & 1;
        const int t = tile / tiles_f, r = tile - t * tiles_f, ty = r / tiles_x, tx = r - ty * tiles_x;
        mbar_wait(empty(s), ph ^ 1);
        const uint32_t st = sbase + s * K::STAGE, fb = full(s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(K::KC * K::RH * K::RW * 16) : "memory");
        for (int c = 0; c < K::KC; ++c)
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
                  "r"(st + c * K::PLANE), "l"(reinterpret_cast<uint64_t>(&tm_in)), "r"(c * 8), "r"(tx * K::TW - 1), "r"(ty * K::TH - 1), "r"(t),
              "r"(fb)
              : "memory");
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: 4 slabs of 4 raster rows x 9 taps x 3 k-steps ----------------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, K::NP);
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int s = i % K::NST, ph = (i / K::NST) & 1, acc = i & 1, aph = (i >> 1) & 1;
        mbar_wait(tmem_empty(acc), aph ^ 1);
        mbar_wait(full(s), ph);
        tc_fence_after();
        // tap-major order: the four slabs' MMAs of one (tap, k-step) go out back to back -- four independent accumulators in flight
        // (consecutive MMAs into ONE accumulator issue ~90 cycles apart here) and the same B operand four times in a row
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int toff = (tap / 3) * K::RW + (tap % 3);
#pragma unroll
          for (int k = 0; k < K::KC / 2; ++k) {
            const uint64_t bd = smem_desc_at(sbase >> 4, K::S_W + (tap * K::KC + 2 * k) * (K::NP * 16), K::NP * 16, 128);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
              const uint64_t ad = smem_desc_at(sbase >> 4, s * K::STAGE + 2 * k * K::PLANE + (mt * 128 + toff) * 16, K::PLANE, 128);
              umma_f16(tmem + acc * 256 + mt * 64, ad, bd, idesc, (tap | k) ? 1u : 0u);
            }
          }
        }
        umma_commit(tmem_full(acc));
        umma_commit(empty(s));
      }
    }
  } else {
    // ---- epilogue warps (2..5): TMEM lane quarter q = warp % 4; slab mt, quarter q = raster row 4 mt + q, lane = x ----------------
    const int q = warp & 3, et = tid - 64;
    const float *bs = reinterpret_cast<const float *>(smem + K::S_BIAS);
    float *red = reinterpret_cast<float *>(smem + K::S_RED);
    const int nch = cout_p / 8;
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int acc = i & 1, aph = (i >> 1) & 1;
      const int t = tile / tiles_f, r = tile - t * tiles_f, ty = r / tiles_x, tx = r - ty * tiles_x;
      const int gx = tx * K::TW + lane;
      const bool xok = lane < K::TW && gx < W;
      float csum[3] = {0.f, 0.f, 0.f};       // lane l: channels 16 g + (l >> 1), g = 0..2, summed over this warp's rows of the tile
      mbar_wait(tmem_full(acc), aph);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < 4; ++mt) {
        const int gy = ty * K::TH + mt * 4 + q;
        const bool valid = xok && gy < H;
        uint32_t v[32], v2[16];
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + acc * 256 + mt * 64;
        tmem_ld32_nowait(ta, v);
        tmem_ld16_nowait(ta + 32, v2);
        tmem_ld_wait();
        if (mt == 3) {                         // everything of this accumulator set is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(acc));
        }
        float o[48];
#pragma unroll
        for (int e = 0; e < 48; ++e) {
          float a = __uint_as_float(e < 32 ? v[e] : v2[e - 32]) + bs[e];
          if (has_prelu) a = a >= 0.f ? a : a * slope;
          o[e] = a;
        }
        if (valid) {
          __half *dp = dst + (((size_t)t * H + gy) * W + gx) * cout_p;
#pragma unroll
          for (int c = 0; c < 6; ++c)
            if (c < nch) *reinterpret_cast<uint4 *>(dp + c * 8) = pack8(*reinterpret_cast<float(*)[8]>(&o[c * 8]));
        }
        if (chan_partial) {
          // butterfly over the 32 pixels of the row, 16 channels at a time: lane l ends with the row sum of channel 16 g + (l >> 1)
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            float z[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) z[e] = valid ? o[g * 16 + e] : 0.f;
#pragma unroll
            for (int off = 16, n = 8; off >= 2; off >>= 1, n >>= 1) {
              const bool hi = lane & off;
#pragma unroll
              for (int e = 0; e < n; ++e) {
                const float send = hi ? z[e] : z[e + n], keep = hi ? z[e + n] : z[e];
                z[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            z[0] += __shfl_xor_sync(0xffffffffu, z[0], 1);
            csum[g] += z[0];
          }
        }
      }
      if (chan_partial) {
        // per-tile sums, fixed order: rows of a warp in sequence (above), then the four warps
        asm volatile("bar.sync 1, 128;\n" ::: "memory");           // the previous tile's reduction has read `red`
        if (!(lane & 1)) {
#pragma unroll
          for (int g = 0; g < 3; ++g) red[q * K::NP + g * 16 + (lane >> 1)] = csum[g];
        }
        asm volatile("bar.sync 2, 128;\n" ::: "memory");
        if (et < cout_p) chan_partial[((size_t)t * tiles_f + r) * cout_p + et] = (red[et] + red[K::NP + et]) + (red[2 * K::NP + et] + red[3 * K::NP + et]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
  }
}

}  // namespace gsn

extern "C" int gsn_conv3x3_tc_tiles(int H, int W) { return ((H + 15) / 16) * ((W + 29) / 30); }

extern "C" int gsn_conv3x3_tc(const GsnConvDesc *dp, void *stream) {
  using namespace gsn;
  using K = C3Cfg;
  GSN_REQUIRE(dp != nullptr, "conv3x3_tc: null descriptor");
  const GsnConvDesc &d = *dp;
  GSN_REQUIRE(d.n_src == 1 && d.src[0] && d.wpack && d.dst, "conv3x3_tc: one source, weights and destination are required");
  GSN_REQUIRE(d.T > 0 && d.Hin > 0 && d.Win > 0 && d.Hout == d.Hin && d.Wout == d.Win, "conv3x3_tc: stride-1 same-size conv only");
  GSN_REQUIRE(d.ks == 3 && d.stride == 1 && d.pad == 1 && !d.residual && !d.pixel_shuffle, "conv3x3_tc: 3x3 / stride 1 / pad 1, no residual or shuffle");
  if ((d.src_c[0] != 40 && d.src_c[0] != 48) || (d.cout_p != 40 && d.cout_p != 48)) {
    set_error("conv3x3_tc: %d -> %d stored channels unsupported (40 or 48 each)", d.src_c[0], d.cout_p);
    return GSN_E_UNSUPPORTED;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (!encode_tmap_nhwc(&tm, d.src[0], d.src_c[0], d.Win, d.Hin, d.T, 8, K::RW, K::RH)) {
    set_error("conv3x3_tc: cuTensorMapEncodeTiled failed (W=%d H=%d T=%d)", d.Win, d.Hin, d.T);
    return GSN_E_CUDA;
  }
  GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  const long long total = (long long)gsn_conv3x3_tc_tiles(d.Hout, d.Wout) * d.T;
  const int sms = sm_count();
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  conv3x3_tc_kernel<<<grid, kC3Threads, K::SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      tm, reinterpret_cast<const unsigned char *>(d.wpack), d.bias, reinterpret_cast<__half *>(d.dst), d.chan_partial, d.T, d.Hout, d.Wout,
      d.cout_p, d.has_prelu, d.prelu_slope);
  count_launch();
  return check_launch("conv3x3_tc");
}

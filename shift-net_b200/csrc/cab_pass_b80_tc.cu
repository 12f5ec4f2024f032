// Pass B of the Ours+ gated blocks (C = 80) on TMA + tcgen05:  out = shortcut + Weff_t . z + beff_t
// (gshift_deblur1.py: last 1x1 of CAB1 / CAB2, CALayer2 scale and beta folded into Weff_t by cab_fold; shortcut = the ROLLED
// stream for CAB2, gshift_deblur1.py:253,257).  The third dense GEMM of the block: [pixels x 80] . [80 x 80] per frame.
//
// Structure as cab_pass_b_tc.cu (C = 64), with k-chunk planar planes instead of swizzled rows (80 channels are not a power-of-two
// row): persistent CTAs, 128-pixel tiles, 3-stage ring; per tile the producer lands ten {8 channels, 128 pixels} boxes of z
// (= the no-swizzle K-major UMMA operand), ten boxes of the shortcut (two rolled channel halves of two frames: the temporal
// roll is just a different box), the frame's folded weight [10][80][8] and bias by bulk copies; warp 1 issues 5 x tcgen05.mma
// (M=128, N=80, K=16) into one of two TMEM accumulators; warps 2..5 add shortcut and bias and write the pixel's 160-byte row.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "shift_common.cuh"
#include "tc_common.cuh"

namespace gsn {

constexpr int kPb80Threads = 192;

struct Pb80Cfg {
  static constexpr int C = 80, MP = 128, NST = 3, KC = 10;
  static constexpr int PLANE = MP * 16;
  static constexpr int Z_BYTES = KC * PLANE, SC_BYTES = KC * PLANE;       // 20 KB each
  static constexpr int W_BYTES = KC * C * 16;                             // 12.8 KB
  static constexpr int B_BYTES = C * 4;
  static constexpr int OFF_SC = Z_BYTES, OFF_W = OFF_SC + SC_BYTES, OFF_B = OFF_W + W_BYTES;
  static constexpr int STAGE = (OFF_B + B_BYTES + 1023) / 1024 * 1024;    // 54 KB
  static constexpr int TX_BYTES = Z_BYTES + SC_BYTES + W_BYTES + B_BYTES;
  static constexpr int S_BAR = NST * STAGE;
  static constexpr int SMEM = S_BAR + 128;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

__global__ void __launch_bounds__(kPb80Threads, 1) cab_pass_b80_tc_kernel(const GsnCabPassB d, const __grid_constant__ CUtensorMap tm_z,
                                                                        const __grid_constant__ CUtensorMap tm_x) {
  using K = Pb80Cfg;
  constexpr int C = K::C;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long hw = (long long)d.H * d.W;
  const int tiles_f = (int)((hw + K::MP - 1) / K::MP), total = tiles_f * d.T;
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
      mbar_init(empty(s), 1 + 4);       // tcgen05.commit (z and W consumed) + the four epilogue warps (shortcut / bias consumed)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full(a), 1);
      mbar_init(tmem_empty(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
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
        const RollSrc rs = roll_source(d.mode, d.circular, t, d.T, C);
        mbar_wait(empty(s), ph ^ 1);
        const uint32_t st = sbase + s * K::STAGE, fb = full(s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(K::TX_BYTES) : "memory");
        auto tma3 = [&](uint32_t dst, const CUtensorMap *tm, int c0, int f) {
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                  "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(p0), "r"(f), "r"(fb)
              : "memory");
        };
        for (int c = 0; c < K::KC; ++c) tma3(st + c * K::PLANE, &tm_z, c * 8, t);
        for (int c = 0; c < 5; ++c) tma3(st + K::OFF_SC + c * K::PLANE, &tm_x, rs.c_lo + c * 8, rs.f_lo);
        for (int c = 0; c < 5; ++c) tma3(st + K::OFF_SC + (5 + c) * K::PLANE, &tm_x, rs.c_hi + c * 8, rs.f_hi);
        const unsigned char *wg = reinterpret_cast<const unsigned char *>(d.weff) + (size_t)t * K::W_BYTES;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                         "r"(st + K::OFF_W), "l"(wg), "r"(K::W_BYTES), "r"(fb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                         "r"(st + K::OFF_B), "l"(d.beff + (size_t)t * C), "r"(K::B_BYTES), "r"(fb) : "memory");
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, C);
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int s = i % K::NST, ph = (i / K::NST) & 1, acc = i & 1, aph = (i >> 1) & 1;
        mbar_wait(tmem_empty(acc), aph ^ 1);
        mbar_wait(full(s), ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < K::KC / 2; ++k) {
          const uint64_t ad = smem_desc_at(sbase >> 4, s * K::STAGE + 2 * k * K::PLANE, K::PLANE, 128);
          const uint64_t bd = smem_desc_at(sbase >> 4, s * K::STAGE + K::OFF_W + 2 * k * (C * 16), C * 16, 128);
          umma_f16(tmem + acc * 128, ad, bd, idesc, k > 0);
        }
        umma_commit(tmem_full(acc));
        umma_commit(empty(s));
      }
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    __half *outp = reinterpret_cast<__half *>(d.out);
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int s = i % K::NST, ph = (i / K::NST) & 1, acc = i & 1, aph = (i >> 1) & 1;
      const int t = tile / tiles_f, p0 = (tile - t * tiles_f) * K::MP;
      const long long pixel = (long long)p0 + r;
      const bool valid = pixel < hw;
      mbar_wait(full(s), ph);             // shortcut + bias of this stage are in shared memory
      mbar_wait(tmem_full(acc), aph);
      tc_fence_after();
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + acc * 128;
      const unsigned char *st = smem + s * K::STAGE;
      const float *be = reinterpret_cast<const float *>(st + K::OFF_B);
      __half *op = outp + ((size_t)t * hw + (valid ? pixel : 0)) * C;
#pragma unroll
      for (int cg = 0; cg < 5; ++cg) {             // 16 channels = two k-chunks of the shortcut at a time
        uint32_t v[16];
        tmem_ld16_nowait(ta + cg * 16, v);
        tmem_ld_wait();
        if (cg == 4) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(acc));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float sc[8], o[8];
          unpack8(*reinterpret_cast<const uint4 *>(st + K::OFF_SC + (2 * cg + h) * K::PLANE + r * 16), sc);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = sc[e] + __uint_as_float(v[h * 8 + e]) + be[cg * 16 + h * 8 + e];
          if (valid) *reinterpret_cast<uint4 *>(op + cg * 16 + h * 8) = pack8(o);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(s));   // this warp is done with the stage's shortcut and bias
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256));
  }
}

bool encode_tmap_chunk128(CUtensorMap *tm, const void *base, int C, long long hw, int T);   // ln_pw_tc.cu

int cab_pass_b80_tc_dispatch(const GsnCabPassB &d, cudaStream_t st) {
  using K = Pb80Cfg;
  const long long hw = (long long)d.H * d.W;
  CUtensorMap tm_z, tm_x;
  memset(&tm_z, 0, sizeof(tm_z));
  memset(&tm_x, 0, sizeof(tm_x));
  if (!encode_tmap_chunk128(&tm_z, d.z, 80, hw, d.T) || !encode_tmap_chunk128(&tm_x, d.x, 80, hw, roll_frames(d.circular, d.T))) {
    set_error("cab_pass_b (C=80): cuTensorMapEncodeTiled failed (H*W=%lld T=%d)", hw, d.T);
    return GSN_E_CUDA;
  }
  GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(cab_pass_b80_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  const long long total = (hw + K::MP - 1) / K::MP * d.T;
  const int sms = sm_count();
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  cab_pass_b80_tc_kernel<<<grid, kPb80Threads, K::SMEM, st>>>(d, tm_z, tm_x);
  count_launch();
  return check_launch("cab_pass_b80_tc");
}

}  // namespace gsn

// Pass B of the fused shift + NAF block on tcgen05 / TMA (sm_100a, C = 64):  out = shortcut + Weff_t . z (+ beff)
// (gshift_deblur2.py:239,257: last 1x1, CALayer2 scale and beta folded into Weff_t by cab_fold; shortcut = the ROLLED stream
// for CAB2), optionally followed by the LayerNorm of the next CAB1 (gshift_deblur2.py:209) written in the k-chunk planar
// layout of GsnCabPassA.a1_pre.
//
// An HBM-bound streaming kernel built so that no compute warp ever issues a load:
//   * persistent CTAs (one per SM), 128-pixel tiles, a 4-stage shared-memory ring;
//   * warp 0 = producer: per tile three TMA tile loads (z tile [128 px][64 ch] with the 128-byte swizzle = the K-major UMMA
//     operand layout; the two rolled halves of the shortcut, [128 px][32 ch] each, 64-byte swizzle) and one bulk copy of the
//     frame's folded weight + bias, all completing on the stage's `full` mbarrier;
//   * warp 1 = MMA issuer: 4 x tcgen05.mma (M=128, N=64, K=16) per tile into one of two TMEM accumulators, tcgen05.commit to
//     the accumulator's `tmem_full` barrier and to the stage's `empty` barrier;
//   * warps 2..5 = epilogue: thread = pixel (TMEM lane): 64 fp32 accumulators + the pixel's 128 shortcut bytes -> fp16 `out`
//     row in a swizzled staging tile (one TMA tile store per tile), LayerNorm statistics without any cross-lane traffic,
//     normalised row straight to the planar a1 (lanes = consecutive pixels -> 512-byte contiguous stores).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "shift_common.cuh"
#include "tc_common.cuh"

namespace gsn {

constexpr int kPbThreads = 192;

struct PbCfg {
  static constexpr int C = 64, MP = 128, NST = 4;
  static constexpr int Z_BYTES = MP * C * 2;             // 16 KB, 128-byte rows, SWIZZLE_128B
  static constexpr int SC_HALF = MP * (C / 2) * 2;       // 8 KB, 64-byte rows, SWIZZLE_64B
  static constexpr int W_BYTES = (C / 8) * C * 16;       // 8 KB, [K/8][N][8] (no-swizzle K-major UMMA operand)
  static constexpr int B_BYTES = C * 4;                  // beff: 64 floats
  static constexpr int OFF_SC = Z_BYTES, OFF_W = OFF_SC + 2 * SC_HALF, OFF_B = OFF_W + W_BYTES;
  static constexpr int STAGE = (OFF_B + B_BYTES + 1023) / 1024 * 1024;                     // 41 KB
  static constexpr int TX_BYTES = Z_BYTES + 2 * SC_HALF + W_BYTES + B_BYTES;
  static constexpr int S_OUT = NST * STAGE;              // 2 x 16 KB out staging (128-byte rows, SWIZZLE_128B)
  static constexpr int S_LN = S_OUT + 2 * Z_BYTES;       // gamma[64], beta[64] of the next LayerNorm
  static constexpr int S_BAR = S_LN + 2 * C * 4;         // full[NST], empty[NST], tmem_full[2], tmem_empty[2], tmem slot
  static constexpr int SMEM = S_BAR + 256;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// K-major operand with the 128-byte swizzle (rows of 64 fp16 = 128 bytes, 8-row atoms of 1024 bytes): SBO = 1024 bytes,
// LBO unused, layout type 2 (SWIZZLE_128B), descriptor version 1.  The k-th 16-element slice starts 32 bytes further.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr_bytes) {
  const uint32_t lo = ((saddr_bytes >> 4) & 0x3FFF) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(kPbThreads, 1) cab_pass_b_tc_kernel(const GsnCabPassB d, const __grid_constant__ CUtensorMap tm_z,
                                                                      const __grid_constant__ CUtensorMap tm_x,
                                                                      const __grid_constant__ CUtensorMap tm_out) {
  using K = PbCfg;
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (d.a1_next) {
    float *lnp = reinterpret_cast<float *>(smem + K::S_LN);
    for (int i = tid; i < 2 * C; i += kPbThreads) lnp[i] = d.ln_next[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- producer ------------------------------------------------------------------------------------------------
    if (lane == 0) {
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int s = i % K::NST, ph = (i / K::NST) & 1;
        const int t = tile / tiles_f, p0 = (tile - t * tiles_f) * K::MP;
        const RollSrc rs = roll_source(d.mode, d.circular, t, d.T, C);
        mbar_wait(empty(s), ph ^ 1);      // a fresh barrier passes the wait on the "previous" phase
        const uint32_t st = sbase + s * K::STAGE, fb = full(s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(K::TX_BYTES) : "memory");
        auto tma3 = [&](uint32_t dst, const CUtensorMap *tm, int c0, int f) {
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                  "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(p0), "r"(f), "r"(fb)
              : "memory");
        };
        tma3(st, &tm_z, 0, t);
        tma3(st + K::OFF_SC, &tm_x, rs.c_lo, rs.f_lo);
        tma3(st + K::OFF_SC + K::SC_HALF, &tm_x, rs.c_hi, rs.f_hi);
        const unsigned char *wg = reinterpret_cast<const unsigned char *>(d.weff) + (size_t)t * K::W_BYTES;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                         "r"(st + K::OFF_W), "l"(wg), "r"(K::W_BYTES), "r"(fb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                         "r"(st + K::OFF_B), "l"(d.beff + (size_t)t * C), "r"(K::B_BYTES), "r"(fb) : "memory");
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, C);
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int s = i % K::NST, ph = (i / K::NST) & 1, acc = i & 1, aph = (i >> 1) & 1;
        mbar_wait(tmem_empty(acc), aph ^ 1);
        mbar_wait(full(s), ph);
        tc_fence_after();
        const uint32_t st = sbase + s * K::STAGE;
#pragma unroll
        for (int k = 0; k < C / 16; ++k) {
          const uint64_t ad = smem_desc_sw128(st + k * 32);
          const uint64_t bd = smem_desc_at(sbase >> 4, s * K::STAGE + K::OFF_W + 2 * k * (C * 16), C * 16, 128);
          umma_f16(tmem + acc * C, ad, bd, idesc, k > 0);
        }
        umma_commit(tmem_full(acc));
        umma_commit(empty(s));
      }
    }
  } else {
    // ---- epilogue warps (2..5): TMEM lane quarter = warp % 4 --------------------------------------------------------------
    const int q = warp & 3, r = q * 32 + lane, et = tid - 64;    // r: pixel row of the tile; et: 0..127
    const float *lnp = reinterpret_cast<const float *>(smem + K::S_LN);
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int s = i % K::NST, ph = (i / K::NST) & 1, acc = i & 1, aph = (i >> 1) & 1, ob = i & 1;
      const int t = tile / tiles_f, p0 = (tile - t * tiles_f) * K::MP;
      const long long pixel = (long long)p0 + r;
      const bool valid = pixel < hw;
      mbar_wait(full(s), ph);             // shortcut + bias of this stage are in shared memory
      mbar_wait(tmem_full(acc), aph);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + acc * C;
      tmem_ld32(ta, v0);
      tmem_ld32(ta + 32, v1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty(acc));
      const unsigned char *st = smem + s * K::STAGE;
      const float *be = reinterpret_cast<const float *>(st + K::OFF_B);
      uint4 o[8];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 sv = *reinterpret_cast<const uint4 *>(st + K::OFF_SC + (c >> 2) * K::SC_HALF + r * 64 + (((c & 3) ^ ((r >> 1) & 3)) << 4));
        float sc[8];
        unpack8(sv, sc);
        const float4 b0 = *reinterpret_cast<const float4 *>(be + c * 8), b1 = *reinterpret_cast<const float4 *>(be + c * 8 + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int idx = c * 8 + e;
          const float a = __uint_as_float(idx < 32 ? v0[idx] : v1[idx - 32]);
          f[e] = sc[e] + a + bb[e];
        }
        o[c] = pack8(f);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(s));   // this warp is done with the stage's shortcut and bias
      // out row -> swizzled staging; the TMA store of two tiles ago must have left this staging buffer
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
      asm volatile("bar.sync 1, 128;\n" ::: "memory");
      unsigned char *og = smem + K::S_OUT + ob * K::Z_BYTES + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4 *>(og + ((c ^ (r & 7)) << 4)) = o[c];
      fence_async_proxy();
      asm volatile("bar.sync 2, 128;\n" ::: "memory");
      if (et == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];\n" ::
                         "l"(reinterpret_cast<uint64_t>(&tm_out)), "r"(0), "r"(p0), "r"(t), "r"(sbase + K::S_OUT + ob * K::Z_BYTES)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
      if (d.a1_next) {
        // LayerNorm of the fp16-rounded row (what the next kernel would read back from HBM)
        float x[64];
#pragma unroll
        for (int c = 0; c < 8; ++c) unpack8(o[c], *reinterpret_cast<float(*)[8]>(&x[c * 8]));
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 64; ++e) { s4[e & 3] += x[e]; q4[e & 3] = fmaf(x[e], x[e], q4[e & 3]); }
        s1 = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        s2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        const float mu = s1 * (1.f / C);
        const float rstd = rsqrtf(fmaxf(s2 * (1.f / C) - mu * mu, 0.f) + 1e-6f);
        const float nmr = -mu * rstd;
        if (valid) {
          __half *ag = reinterpret_cast<__half *>(d.a1_next) + ((size_t)t * (C / 8) * hw + pixel) * 8;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaf(fmaf(x[c * 8 + e], rstd, nmr), lnp[c * 8 + e], lnp[C + c * 8 + e]);
            *reinterpret_cast<uint4 *>(ag + (size_t)c * hw * 8) = pack8(f);
          }
        }
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128));
  }
}

// 3-D fp16 tensor map over a (T, HW, C) pixel-major tensor (dims innermost first: C, HW, T) with a (bc, bp, 1) box
static bool encode_tmap_pix(CUtensorMap *tm, const void *base, int C, long long hw, int T, int bc, int bp, CUtensorMapSwizzle sw) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn enc = [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)hw, (cuuint64_t)T};
  const cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)hw * C * 2};
  const cuuint32_t box[3] = {(cuuint32_t)bc, (cuuint32_t)bp, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  // half-line boxes (the rolled shortcut halves, 64-byte rows) must not be promoted to 128-byte L2 fetches: that doubled the
  // DRAM reads of the shortcut (ncu: 1.70 GB read per level-1 launch for 1.18 GB of operands)
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             bc * 2 <= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int cab_pass_b_tc_dispatch(const GsnCabPassB &d, cudaStream_t st) {
  using K = PbCfg;
  const long long hw = (long long)d.H * d.W;
  CUtensorMap tm_z, tm_x, tm_out;
  memset(&tm_z, 0, sizeof(tm_z));
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_out, 0, sizeof(tm_out));
  if (!encode_tmap_pix(&tm_z, d.z, 64, hw, d.T, 64, K::MP, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !encode_tmap_pix(&tm_x, d.x, 64, hw, roll_frames(d.circular, d.T), 32, K::MP, CU_TENSOR_MAP_SWIZZLE_64B) ||
      !encode_tmap_pix(&tm_out, d.out, 64, hw, d.T, 64, K::MP, CU_TENSOR_MAP_SWIZZLE_128B)) {
    set_error("cab_pass_b: cuTensorMapEncodeTiled failed (H*W=%lld T=%d)", hw, d.T);
    return GSN_E_CUDA;
  }
  GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(cab_pass_b_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  const int num_sms = sm_count();
  const long long total = (hw + K::MP - 1) / K::MP * d.T;
  const unsigned grid = (unsigned)(total < num_sms ? total : num_sms);
  cab_pass_b_tc_kernel<<<grid, kPbThreads, K::SMEM, st>>>(d, tm_z, tm_x, tm_out);
  count_launch();
  return check_launch("cab_pass_b_tc");
}

}  // namespace gsn

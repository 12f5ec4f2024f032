// LayerNorm + first 1x1 of the Ours+ gated blocks (C = 80; gshift_deblur1.py:186-258: CAB1.norm / CAB2.norm then body.0) as
// a streaming TMA + tcgen05 kernel (sm_100a).  The 1x1 convs are the genuinely dense GEMMs of the path: this one is
// [pixels x CIN] . [CIN x 2C] with CIN = 80 (CAB1) or 120 (CAB2: rolled stream 80 | conv1(shifted half) 40), N = 160.
//
// The LayerNorm is FOLDED around the GEMM so that the tensor core can eat the raw fp16 activations straight from the TMA tiles:
//     W . LN(x) = W . (gamma * (x - mu) * rstd + beta) = rstd * (W' . x  -  mu * rowsum(W'))  +  W . beta,      W' = W diag(gamma)
// W' (rounded to fp16), rowsum(W') (of the ROUNDED weights, so the mean term cancels exactly) and W.beta are packed on the host;
// mu and rstd of every pixel come from the same shared-memory tile the tensor core reads (thread = pixel, fp32).
//
// Structure (as cab_pass_b_tc.cu): persistent CTAs, 128-pixel tiles, a 3-stage shared-memory ring;
//   warp 0 = producer: one TMA box {8 channels, 128 pixels} per k-chunk -> k-chunk planar planes = the no-swizzle K-major UMMA
//            operand layout; CAB2's three sources (two rolled channel halves of two frames, the shifted half) are just different
//            boxes of the same ring stage -- the temporal roll costs nothing here;
//   warp 1 = MMA issuer: CIN/16 x tcgen05.mma (M=128, N=160, K=16) per tile into one of two TMEM accumulators;
//   warps 2..5 = epilogue: per-pixel statistics from the stage, then rstd * (acc - mu * wsum) + bias -> fp16 -> the pixel's
//            contiguous 160-byte rows of the two outputs (a | b halves).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "shift_common.cuh"
#include "tc_common.cuh"

namespace gsn {

constexpr int kLpThreads = 192;

struct LpCfg {
  static constexpr int C = 80, N = 160, MP = 128, NST = 3, KCMAX = 16;
  static constexpr int PLANE = MP * 16;                      // one k-chunk of a tile: [128 px][16 B]
  static constexpr int STAGE = KCMAX * PLANE;                // 32 KB
  static constexpr int W_BYTES = KCMAX * N * 16;             // folded weights [16 planes][160][8] fp16 = 40 KB (zero-padded K)
  static constexpr int S_W = NST * STAGE;
  static constexpr int S_VEC = S_W + W_BYTES;                // wsum[160], bias[160] fp32
  static constexpr int S_BAR = S_VEC + 2 * N * 4;            // full[NST], empty[NST], tmem_full[2], tmem_empty[2], tmem slot
  static constexpr int SMEM = S_BAR + 128;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

__global__ void __launch_bounds__(kLpThreads, 1) ln_pw_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hw,
                                                               const unsigned char *__restrict__ wfold, const float *__restrict__ wvec,
                                                               __half *__restrict__ ga, __half *__restrict__ gb, int T, long long hw, int mode,
                                                               int circular) {
  using K = LpCfg;
  constexpr int C = K::C;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool shift = mode != GSN_MODE_CAB1;
  const int kc = shift ? 15 : 10, ksteps = shift ? 8 : 5, cin = kc * 8;      // k-chunks, MMA k-steps (CAB2: chunk 15 is a zero plane)
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
      mbar_init(empty(s), 1 + 4);       // tcgen05.commit (operand consumed) + the four epilogue warps (statistics read)
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
  // resident: folded weights, the two epilogue vectors; the padding plane of every stage (CAB2's 16th k-chunk) is zero for good
  for (int i = tid; i < K::W_BYTES / 16; i += kLpThreads) cp_async16(smem + K::S_W + i * 16, wfold + (size_t)i * 16, true);
  cp_async_commit();
  for (int i = tid; i < 2 * K::N; i += kLpThreads) reinterpret_cast<float *>(smem + K::S_VEC)[i] = wvec[i];
  for (int i = tid; i < K::NST * (K::PLANE / 16); i += kLpThreads) {
    const int s = i / (K::PLANE / 16), r = i - s * (K::PLANE / 16);
    *reinterpret_cast<uint4 *>(smem + s * K::STAGE + 15 * K::PLANE + r * 16) = make_uint4(0, 0, 0, 0);
  }
  cp_async_wait<0>();
  fence_async_proxy();     // generic-proxy writes (weights, zero planes) -> visible to the tensor core
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
        const RollSrc rs = roll_source(mode, circular, t, T, C);
        mbar_wait(empty(s), ph ^ 1);      // a fresh barrier passes the wait on the "previous" phase
        const uint32_t st = sbase + s * K::STAGE, fb = full(s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(kc * K::PLANE) : "memory");
        auto tma3 = [&](uint32_t dst, const CUtensorMap *tm, int c0, int f) {
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                  "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(p0), "r"(f), "r"(fb)
              : "memory");
        };
        if (!shift) {
          for (int c = 0; c < 10; ++c) tma3(st + c * K::PLANE, &tm_x, c * 8, t);
        } else {
          // LN input of CAB2 = [rolled stream (low half | high half) | conv1(shifted half)] (gshift_deblur1.py:250-254)
          for (int c = 0; c < 5; ++c) tma3(st + c * K::PLANE, &tm_x, rs.c_lo + c * 8, rs.f_lo);
          for (int c = 0; c < 5; ++c) tma3(st + (5 + c) * K::PLANE, &tm_x, rs.c_hi + c * 8, rs.f_hi);
          for (int c = 0; c < 5; ++c) tma3(st + (10 + c) * K::PLANE, &tm_hw, c * 8, t);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, K::N);
      int i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int s = i % K::NST, ph = (i / K::NST) & 1, acc = i & 1, aph = (i >> 1) & 1;
        mbar_wait(tmem_empty(acc), aph ^ 1);
        mbar_wait(full(s), ph);
        tc_fence_after();
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ad = smem_desc_at(sbase >> 4, s * K::STAGE + 2 * k * K::PLANE, K::PLANE, 128);
          const uint64_t bd = smem_desc_at(sbase >> 4, K::S_W + 2 * k * (K::N * 16), K::N * 16, 128);
          umma_f16(tmem + acc * 256, ad, bd, idesc, k > 0);
        }
        umma_commit(tmem_full(acc));
        umma_commit(empty(s));
      }
    }
  } else {
    // ---- epilogue warps (2..5): TMEM lane quarter = warp % 4, thread = pixel -------------------------------------------
    const int q = warp & 3, r = q * 32 + lane;
    const float *wsum = reinterpret_cast<const float *>(smem + K::S_VEC), *bias = wsum + K::N;
    const float inv_cin = 1.f / (float)cin;
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int s = i % K::NST, ph = (i / K::NST) & 1, acc = i & 1, aph = (i >> 1) & 1;
      const int t = tile / tiles_f, p0 = (tile - t * tiles_f) * K::MP;
      const long long pixel = (long long)p0 + r;
      const bool valid = pixel < hw;
      mbar_wait(full(s), ph);
      // per-pixel LayerNorm statistics over the CIN raw channels (biased variance, eps 1e-6; gshift_deblur1.py:19-28)
      float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
      const unsigned char *pp = smem + s * K::STAGE + r * 16;
      for (int c = 0; c < kc; ++c) {
        float v[8];
        unpack8(*reinterpret_cast<const uint4 *>(pp + c * K::PLANE), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s4[e & 3] += v[e]; q4[e & 3] = fmaf(v[e], v[e], q4[e & 3]); }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(s));        // this warp is done with the stage
      const float mu = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * inv_cin;
      const float rstd = rsqrtf(fmaxf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * inv_cin - mu * mu, 0.f) + 1e-6f);
      const float nmr = -mu * rstd;
      mbar_wait(tmem_full(acc), aph);
      tc_fence_after();
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + acc * 256;
      __half *gpa = ga + ((size_t)t * hw + (valid ? pixel : 0)) * C, *gpb = gb + ((size_t)t * hw + (valid ? pixel : 0)) * C;
#pragma unroll 1
      for (int cg = 0; cg < 5; ++cg) {             // 32 accumulator columns at a time
        uint32_t v[32];
        tmem_ld32(ta + cg * 32, v);
        if (cg == 4) {                             // the accumulator is in registers: hand the TMEM slot back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(acc));
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n0 = cg * 32 + j * 8;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)            // rstd * (acc - mu * wsum) + bias
              o[e] = fmaf(__uint_as_float(v[j * 8 + e]), rstd, fmaf(nmr, wsum[n0 + e], bias[n0 + e]));
            __half *dst = n0 < C ? gpa + n0 : gpb + (n0 - C);
            *reinterpret_cast<uint4 *>(dst) = pack8(o);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
  }
}

// 3-D fp16 tensor map over a (T, HW, C) pixel-major tensor with an (8 channels, 128 pixels, 1) box: one k-chunk plane of a tile
bool encode_tmap_chunk128(CUtensorMap *tm, const void *base, int C, long long hw, int T) {
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
  const cuuint32_t box[3] = {8, (cuuint32_t)LpCfg::MP, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gsn

extern "C" int gsn_ln_pw_tc(const void *x, const void *hw_pre, int T, int H, int W, int C, int mode, int circular, const void *wfold,
                            const float *wvec, void *ga, void *gb, void *stream) {
  using namespace gsn;
  using K = LpCfg;
  GSN_REQUIRE(x && wfold && wvec && ga && gb, "ln_pw_tc: null pointer");
  GSN_REQUIRE(T > 0 && H > 0 && W > 0, "ln_pw_tc: empty shape");
  GSN_REQUIRE(mode >= GSN_MODE_CAB1 && mode <= GSN_MODE_CAB2_REV, "ln_pw_tc: mode=%d", mode);
  GSN_REQUIRE(mode == GSN_MODE_CAB1 || hw_pre, "ln_pw_tc: CAB2 modes need the gsn_shift_conv1 output");
  if (C != 80) { set_error("ln_pw_tc: C=%d unsupported (80)", C); return GSN_E_UNSUPPORTED; }
  const long long hw = (long long)H * W;
  CUtensorMap tm_x, tm_hw;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_hw, 0, sizeof(tm_hw));
  if (!encode_tmap_chunk128(&tm_x, x, C, hw, roll_frames(circular, T)) || (hw_pre && !encode_tmap_chunk128(&tm_hw, hw_pre, C / 2, hw, T))) {
    set_error("ln_pw_tc: cuTensorMapEncodeTiled failed (H*W=%lld T=%d)", hw, T);
    return GSN_E_CUDA;
  }
  GSN_ONCE_PER_DEVICE(cudaFuncSetAttribute(ln_pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  const long long total = (hw + K::MP - 1) / K::MP * T;
  const int sms = sm_count();
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  ln_pw_tc_kernel<<<grid, kLpThreads, K::SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(
      tm_x, tm_hw, reinterpret_cast<const unsigned char *>(wfold), wvec, reinterpret_cast<__half *>(ga), reinterpret_cast<__half *>(gb), T, hw,
      mode, circular);
  count_launch();
  return check_launch("ln_pw_tc");
}
